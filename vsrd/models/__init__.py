from .detectors import BoxParameters3D  # noqa: F401
from .encoders import SinusoidalEncoder  # noqa: F401
from .fields import HyperDistanceField  # noqa: F401
