from .geometric_operations import (  # noqa: F401
    clip_lines_to_front, expand_to_4x4, project_box_3d, rotation_matrix_x, rotation_matrix_y, rotation_matrix_z)
from .kitti360_operations import box_3d_iou, box_3d_iou_exact  # noqa: F401
