from .box_parameters import BoxParameters3D, rotation_matrix_y  # noqa: F401
