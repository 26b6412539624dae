"""`vsrd.operations` geometry used by the optimisation loop (API of
vsrd/operations/geometric_operations.py: expand_to_4x4 :9-14, rotation_matrix_{x,y,z} :28-64,
clip_lines_to_front :343-365, project_box_3d :368-389).  Plain PyTorch on tiny tensors; fusing the
per-step projection loop into one kernel is SURVEY.md §8f "next" row 1."""
import torch


def expand_to_4x4(matrices):
    out = torch.eye(4, dtype=matrices.dtype, device=matrices.device).repeat(*matrices.shape[:-2], 1, 1)
    out[..., :matrices.shape[-2], :matrices.shape[-1]] = matrices
    return out


def _axis_rotation(angles, axis):
    c, s = torch.cos(angles), torch.sin(angles)
    o, z = torch.ones_like(angles), torch.zeros_like(angles)
    rows = {
        "x": ((o, z, z), (z, c, -s), (z, s, c)),
        "y": ((c, z, s), (z, o, z), (-s, z, c)),
        "z": ((c, -s, z), (s, c, z), (z, z, o)),
    }[axis]
    return torch.stack([torch.stack(row, dim=-1) for row in rows], dim=-2)


def rotation_matrix_x(angles):
    return _axis_rotation(angles, "x")


def rotation_matrix_y(angles):
    return _axis_rotation(angles, "y")


def rotation_matrix_z(angles):
    return _axis_rotation(angles, "z")


def clip_lines_to_front(lines, epsilon=1e-6):
    """Clip 3D segments [..., 2, 3] against the plane z = 0, keeping the part in front of the camera.
    Returns the clipped segments (far end first) and a mask of segments with a visible part."""
    a, b = lines[..., 0, :], lines[..., 1, :]
    a_is_far = a[..., -1:] > b[..., -1:]
    far = torch.where(a_is_far, a, b)
    near = torch.where(a_is_far, b, a)
    z_far, z_near = far[..., -1:], near[..., -1:]
    t = (z_far / torch.clamp(z_far - z_near, min=epsilon)).clamp(max=1.0)
    near = far + (near - far) * t
    return torch.stack([far, near], dim=-2), far[..., -1] > 0


_BOX_EDGES = ((0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7))   # main.py:27-31


class _ProjectBox3D(torch.autograd.Function):
    """One launch forward, one backward (vsrd_project_box_3d*): scripts/main.py:346-355 calls project_box_3d once per view
    and instance (136 times per step); as ~25 small PyTorch ops with two host synchronisations each (`torch.any`, boolean
    indexing) that loop is most of the script's step time once the renderer is fused."""

    @staticmethod
    def forward(ctx, box_3d, intrinsic_matrix, epsilon):
        from vsrd_b200 import ops
        ctx.save_for_backward(box_3d, intrinsic_matrix)
        ctx.epsilon = epsilon
        return ops.project_box_3d(box_3d.detach(), intrinsic_matrix.detach(), epsilon).reshape(*box_3d.shape[:-2], 2, 2)

    @staticmethod
    def backward(ctx, grad):
        from vsrd_b200 import ops
        box_3d, intrinsic_matrix = ctx.saved_tensors
        return ops.project_box_3d_backward(box_3d, intrinsic_matrix, grad.contiguous(), ctx.epsilon).reshape(box_3d.shape), None, None


def project_box_3d(box_3d, line_indices, intrinsic_matrix, epsilon=1e-6):
    """2D bounding box [2,2] (min; max) of the visible part of the 12 edges of `box_3d` [8,3]
    (camera frame); zeros when the box is entirely behind the camera.  CUDA float32 boxes with main.py's LINE_INDICES
    run as one fused kernel (differentiable w.r.t. `box_3d`); anything else as the PyTorch ops below."""
    if (box_3d.is_cuda and box_3d.dtype == torch.float32 and intrinsic_matrix.shape == (3, 3) and not intrinsic_matrix.requires_grad
            and tuple(map(tuple, line_indices)) == _BOX_EDGES):
        return _ProjectBox3D.apply(box_3d, intrinsic_matrix, float(epsilon))
    lines, visible = clip_lines_to_front(box_3d[..., line_indices, :], epsilon)
    pixels = lines @ intrinsic_matrix.T
    pixels = pixels[..., :-1] / torch.clamp(pixels[..., -1:], min=epsilon)
    if not torch.any(visible):
        return box_3d.new_zeros(*box_3d.shape[:-2], 2, 2)
    points = pixels[visible, ...].flatten(-3, -2)
    return torch.stack([points.min(dim=-2).values, points.max(dim=-2).values], dim=-2)
