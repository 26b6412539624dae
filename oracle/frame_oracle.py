"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch / numpy, autograd for the adjoints) of the
per-step rows of the reference's hot path that sit around the renderer.  Nothing in the product path
imports this module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.

Pinned against the UNMODIFIED reference by tests/golden/make_golden_frame.py -> tests/golden/frame.npz
(project_box_3d / clip_lines_to_front imported from the reference checkout, torchvision + scipy as the
reference calls them, SoftRasterizer.make_distance_map compiled from the reference's source text).

    a14  project_box_3d, clip_lines_to_front   vsrd/operations/geometric_operations.py:343-389
         multi_view_boxes_2d                   scripts/main.py:339-367
    a15  matching, projection_losses           scripts/main.py:374-415
    a2   select_rays (sequential weighted draw without replacement), gather_targets   scripts/main.py:620-627, 656
         soft_mask distance map                vsrd/transforms/geometric_transforms.py:267-309
         annealing schedule                    scripts/main.py:420-431, 677
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LINE_INDICES = [                      # scripts/main.py:26-30
    [0, 1], [1, 2], [2, 3], [3, 0],
    [4, 5], [5, 6], [6, 7], [7, 4],
    [0, 4], [1, 5], [2, 6], [3, 7],
]


# ------------------------------------------------------------------------------------------------
# a14
# ------------------------------------------------------------------------------------------------

def clip_lines_to_front(lines, epsilon=1e-6):
    """geometric_operations.py:343-365: order each segment by depth, pull the far-behind end onto z=0."""
    p1, p2 = torch.unbind(lines, dim=-2)
    d1, d2 = p1[..., -1:], p2[..., -1:]
    swap = d1 > d2
    p1, p2 = torch.where(swap, p1, p2), torch.where(swap, p2, p1)
    d1, d2 = torch.where(swap, d1, d2), torch.where(swap, d2, d1)
    w = torch.clamp(d1 / torch.clamp(d1 - d2, min=epsilon), max=1.0)
    p2 = p1 + (p2 - p1) * w
    return torch.stack([p1, p2], dim=-2), p1[..., -1] > 0


def project_box_3d(box_3d, intrinsic_matrix, epsilon=1e-6):
    """geometric_operations.py:368-389 for one box [8,3] in the camera frame -> [2,2] (min, max)."""
    lines, masks = clip_lines_to_front(box_3d[..., LINE_INDICES, :], epsilon)
    lines = lines @ intrinsic_matrix.T
    lines = lines[..., :-1] / torch.clamp(lines[..., -1:], min=epsilon)
    if torch.any(masks):
        points = lines[masks, ...].flatten(-3, -2)
        return torch.stack([points.min(dim=-2).values, points.max(dim=-2).values], dim=-2)
    return box_3d.new_zeros(*box_3d.shape[:-2], 2, 2)


def multi_view_boxes_2d(world_boxes_3d, extrinsic_matrices, intrinsic_matrices, image_size):
    """scripts/main.py:339-362: world corners [N,8,3] -> per-view clipped 2D boxes [V,N,4] (x1 y1 x2 y2)."""
    h, w = image_size
    homog = F.pad(world_boxes_3d, (0, 1), mode="constant", value=1.0)
    out = []
    for e, k in zip(extrinsic_matrices, intrinsic_matrices):
        cam = torch.einsum("mn,...n->...m", e, homog)
        cam = cam[..., :-1] / cam[..., -1:]
        boxes = torch.stack([project_box_3d(b, k) for b in cam], dim=0).flatten(-2, -1)
        bx = boxes[..., 0::2].clamp(min=0, max=w)          # torchvision.ops.clip_boxes_to_image
        by = boxes[..., 1::2].clamp(min=0, max=h)
        out.append(torch.stack([bx, by], dim=-1).reshape(boxes.shape))
    return torch.stack(out, dim=0)


# ------------------------------------------------------------------------------------------------
# a15
# ------------------------------------------------------------------------------------------------

def distance_box_iou(boxes1, boxes2, eps=1e-7):
    """torchvision.ops.distance_box_iou (boxes.py::_box_diou_iou): pairwise [N,M]."""
    area1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    area2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    lt = torch.max(boxes1[:, None, :2], boxes2[None, :, :2])
    rb = torch.min(boxes1[:, None, 2:], boxes2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    iou = inter / (area1[:, None] + area2[None, :] - inter)
    lti = torch.min(boxes1[:, None, :2], boxes2[None, :, :2])
    rbi = torch.max(boxes1[:, None, 2:], boxes2[None, :, 2:])
    whi = (rbi - lti).clamp(min=0)
    diag = whi[..., 0] ** 2 + whi[..., 1] ** 2 + eps
    cx = (boxes1[:, 0] + boxes1[:, 2])[:, None] / 2 - (boxes2[:, 0] + boxes2[:, 2])[None, :] / 2
    cy = (boxes1[:, 1] + boxes1[:, 3])[:, None] / 2 - (boxes2[:, 1] + boxes2[:, 3])[None, :] / 2
    return iou - (cx ** 2 + cy ** 2) / diag


def distance_box_iou_loss(boxes1, boxes2, eps=1e-7):
    """torchvision.ops.distance_box_iou_loss(reduction='none') (diou_loss.py::_diou_iou_loss)."""
    x1, y1, x2, y2 = boxes1.unbind(dim=-1)
    x1g, y1g, x2g, y2g = boxes2.unbind(dim=-1)
    xk1, yk1 = torch.max(x1, x1g), torch.max(y1, y1g)
    xk2, yk2 = torch.min(x2, x2g), torch.min(y2, y2g)
    mask = (yk2 > yk1) & (xk2 > xk1)
    inter = torch.where(mask, (xk2 - xk1) * (yk2 - yk1), torch.zeros_like(x1))
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter
    iou = inter / (union + eps)
    xc1, yc1 = torch.min(x1, x1g), torch.min(y1, y1g)
    xc2, yc2 = torch.max(x2, x2g), torch.max(y2, y2g)
    diag = (xc2 - xc1) ** 2 + (yc2 - yc1) ** 2 + eps
    cen = ((x2 + x1) / 2 - (x1g + x2g) / 2) ** 2 + ((y2 + y1) / 2 - (y1g + y2g) / 2) ** 2
    return 1 - iou + cen / diag


def matching(pd_boxes_target, gt_boxes_target):
    """scripts/main.py:374-386: Hungarian assignment on -DIoU of the target view -> (pd_indices, gt_indices)."""
    from scipy.optimize import linear_sum_assignment
    cost = -distance_box_iou(pd_boxes_target.detach(), gt_boxes_target.detach())
    rows, cols = linear_sum_assignment(cost.cpu().numpy())
    return torch.as_tensor(rows, dtype=torch.int64), torch.as_tensor(cols, dtype=torch.int64)


def projection_losses(pd_boxes_2d, gt_boxes_2d, visible_masks, pd_indices, gt_indices):
    """scripts/main.py:391-415: means over the visible matched pairs of every view.
    pd/gt boxes [V,N,4], visible_masks [V,N] bool."""
    iou_terms, l1_terms = [], []
    for pd, gt, vis in zip(pd_boxes_2d, gt_boxes_2d, visible_masks):
        keep = vis[gt_indices]
        a, b = pd[pd_indices[keep]], gt[gt_indices[keep]]
        iou_terms.append(distance_box_iou_loss(a, b))
        l1_terms.append(F.smooth_l1_loss(a, b, reduction="none"))
    return torch.mean(torch.cat(iou_terms, dim=0)), torch.mean(torch.cat(l1_terms, dim=0))


def projection_step(world_boxes_3d, extrinsic_matrices, intrinsic_matrices, image_size, gt_boxes_2d, visible_masks,
                    target_view, fixed_gt_indices=None):
    """a14 + a15 in one call -> (boxes_2d [V,N,4], gt_indices [N], iou_loss, l1_loss)."""
    boxes = multi_view_boxes_2d(world_boxes_3d, extrinsic_matrices, intrinsic_matrices, image_size)
    if fixed_gt_indices is None:
        pd_idx, gt_idx = matching(boxes[target_view], gt_boxes_2d[target_view])
    else:
        gt_idx = torch.as_tensor(fixed_gt_indices, dtype=torch.int64)
        pd_idx = torch.arange(gt_idx.numel())
    iou, l1 = projection_losses(boxes, gt_boxes_2d, visible_masks, pd_idx, gt_idx)
    return boxes, gt_idx, iou, l1


# ------------------------------------------------------------------------------------------------
# a2
# ------------------------------------------------------------------------------------------------

def ray_weights(soft_masks):
    """scripts/main.py:621-624: per-pixel weight = max over instances; soft_masks [..., N] -> [P]."""
    return soft_masks.reshape(-1, soft_masks.shape[-1]).max(dim=-1).values


def select_rays(weights, num_rays, uniforms):
    """Sequential weighted sampling without replacement (what torch.multinomial(replacement=False) draws
    from, main.py:620-627), driven by injected uniforms: each uniform picks a pixel by inverse CDF of the
    FULL distribution; a pixel already taken is rejected and the next uniform is used.  Rejecting repeats
    of i.i.d. draws is exactly drawing from the renormalised remainder.  Returns int64 [num_rays]."""
    cdf = np.cumsum(np.asarray(weights, dtype=np.float64))
    total = cdf[-1]
    taken, out = set(), []
    for u in np.asarray(uniforms, dtype=np.float64):
        x = min(u * total, total * (1.0 - 2.0 ** -53))
        i = int(np.searchsorted(cdf, x, side="right"))
        i = min(i, cdf.size - 1)
        if i in taken:
            continue
        taken.add(i)
        out.append(i)
        if len(out) == num_rays:
            break
    if len(out) < num_rays:
        raise RuntimeError(f"only {len(out)} of {num_rays} rays could be drawn from {len(uniforms)} uniforms")
    return torch.as_tensor(out, dtype=torch.int64)


def gather_targets(soft_masks, pixel_indices, gt_indices=None):
    """scripts/main.py:656: soft_masks.flatten(0,-2)[rays][..., gt_indices]."""
    flat = soft_masks.reshape(-1, soft_masks.shape[-1])[pixel_indices]
    return flat if gt_indices is None else flat[..., gt_indices]


# ------------------------------------------------------------------------------------------------
# soft masks (SoftRasterizer) and the schedule
# ------------------------------------------------------------------------------------------------

def polygon_distance_map(polygon, image_size):
    """geometric_transforms.py:267-290 make_distance_map: unsigned pixel distance to a closed polygon [P,2]."""
    positions = list(reversed(torch.meshgrid(*map(torch.arange, image_size), indexing="ij")))
    positions = torch.stack(list(map(torch.flatten, positions)), dim=-1)
    prev_v, next_v = polygon, torch.roll(polygon, shifts=-1, dims=-2)
    sides = next_v.unsqueeze(-3) - prev_v.unsqueeze(-3)
    positions = positions.unsqueeze(-2) - prev_v.unsqueeze(-3)
    ratios = (sides * positions).sum(-1, keepdim=True) / ((sides * sides).sum(-1, keepdim=True) + 1e-6)
    normals = positions - sides * torch.clamp(ratios, 0.0, 1.0)
    return torch.linalg.norm(normals, dim=-1).min(dim=-1).values.unflatten(-1, image_size)


def polygon_inside(polygon, image_size):
    """Even-odd point-in-polygon test at integer pixel coordinates (stands in for cv.fillPoly, :256-265)."""
    h, w = image_size
    ys, xs = torch.meshgrid(torch.arange(h, dtype=polygon.dtype), torch.arange(w, dtype=polygon.dtype), indexing="ij")
    inside = torch.zeros(h, w, dtype=torch.bool)
    a, b = polygon, torch.roll(polygon, shifts=-1, dims=0)
    for (ax, ay), (bx, by) in zip(a.tolist(), b.tolist()):
        crosses = (ay > ys) != (by > ys)
        if by == ay:
            continue
        xint = ax + (ys - ay) * (bx - ax) / (by - ay)
        inside ^= crosses & (xs < xint)
    return inside


def soft_mask(polygon, image_size, temperature=10.0):
    """geometric_transforms.py:301-309: sigmoid(signed distance / temperature)."""
    dist = polygon_distance_map(polygon, image_size)
    return torch.sigmoid(torch.where(polygon_inside(polygon, image_size), dist, -dist) / temperature)


def schedule(step, num_steps, warmup_steps, temperature=(1.0, 0.1), std_deviation=(1.0, 0.1), eikonal_weight=0.01):
    """scripts/main.py:420-431, 677 (numpy double arithmetic, as the script computes it)."""
    anneal = lambda x, a, b: (np.cos(np.pi * x) + 1.0) / 2.0 * (a - b) + b
    x = step / num_steps
    return dict(temperature=float(anneal(x, *temperature)), std_deviation=float(anneal(x, *std_deviation)),
                cosine_ratio=x, eikonal_weight=eikonal_weight if step >= warmup_steps else 0.0)
