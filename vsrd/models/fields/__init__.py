from .hyper_distance_field import HyperDistanceField  # noqa: F401
