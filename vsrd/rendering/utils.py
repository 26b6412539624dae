"""`vsrd.rendering.ray_casting` (reference: vsrd/rendering/utils.py:5-18) on the ray kernel."""
import torch

from vsrd_b200 import ops


def ray_casting(image_size, intrinsic_matrices, extrinsic_matrices):
    """image_size (H, W); intrinsics [...,3,3]; extrinsics [...,4,4] (world -> camera).
    Returns camera positions [...,3] and unit ray directions [...,H,W,3]."""
    height, width = int(image_size[0]), int(image_size[1])
    inv_k = torch.linalg.inv(intrinsic_matrices)
    inv_e = torch.linalg.inv(extrinsic_matrices)
    inv_proj = inv_e[..., :3, :3] @ inv_k
    lead = inv_proj.shape[:-2]
    dirs = ops.ray_directions(inv_proj.reshape(-1, 3, 3).float(), height, width)
    return inv_e[..., :3, 3], dirs.reshape(*lead, height, width, 3).to(inv_proj.dtype)
