"""Synthetic KITTI-360-shaped frames for tests, smoke and benchmarks.

World frame = rectified target camera (x right, y down, z forward).  A frame holds N ground-truth
boxes, V pinhole views along a forward-moving trajectory, and helpers to draw per-step ray batches
near the instances (the reference draws them with a multinomial over the soft masks,
scripts/main.py:620-627; the draw itself is a "next" row, so batches are pre-generated here).
All tensors are created on the host with a seeded generator and moved to the device by the caller.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch

# KITTI-360 perspective camera (calibration constants, treated as synthetic)
KITTI360_INTRINSICS = (552.554, 552.554, 682.049, 238.770)
KITTI360_IMAGE_SIZE = (376, 1408)
DIMENSION_RANGE = ((0.75, 0.75, 1.5), (1.00, 1.00, 2.5))      # box_parameters.py:27-30 (half extents)
LOCATION_RANGE = ((-50.0, 1.55 - 1.75 / 2.0 - 5.0, 0.0), (50.0, 1.55 - 1.75 / 2.0 + 5.0, 100.0))


def rotation_y(yaw: torch.Tensor) -> torch.Tensor:
    c, s = torch.cos(yaw), torch.sin(yaw)
    o, z = torch.ones_like(c), torch.zeros_like(c)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


@dataclasses.dataclass
class SyntheticFrame:
    image_size: tuple                 # (H, W)
    intrinsics: torch.Tensor          # [V,3,3]
    extrinsics: torch.Tensor          # [V,4,4] world -> camera
    gt_locations: torch.Tensor        # [N,3]
    gt_half_extents: torch.Tensor     # [N,3]
    gt_yaws: torch.Tensor             # [N]

    @property
    def num_views(self):
        return self.intrinsics.shape[0]

    @property
    def num_instances(self):
        return self.gt_locations.shape[0]

    @property
    def gt_rotations(self):
        return rotation_y(self.gt_yaws)

    def inverse_projections(self):
        """inv(E)[:3,:3] @ inv(K) and camera centres, as `ray_casting` forms them (rendering/utils.py:8-17)."""
        inv_e = torch.linalg.inv(self.extrinsics)
        inv_k = torch.linalg.inv(self.intrinsics)
        return (inv_e[:, :3, :3] @ inv_k).contiguous(), inv_e[:, :3, 3].contiguous()

    def project(self, points: torch.Tensor, view: torch.Tensor) -> torch.Tensor:
        """World points [P,3] into pixel coordinates of `view` [P] -> [P,2] (u, v) and depth [P]."""
        e = self.extrinsics[view]
        cam = torch.einsum("pmn,pn->pm", e[:, :3, :3], points) + e[:, :3, 3]
        k = self.intrinsics[view]
        uvw = torch.einsum("pmn,pn->pm", k, cam)
        return uvw[:, :2] / uvw[:, 2:].clamp_min(1e-3), cam[:, 2]

    def draw_pixel_indices(self, num_rays: int, gen: torch.Generator, spread: float = 1.3) -> torch.Tensor:
        """Flat indices into [V,H,W] of pixels scattered around the projected instances."""
        h, w = self.image_size
        out = []
        need = num_rays
        while need > 0:
            m = need * 2
            view = torch.randint(0, self.num_views, (m,), generator=gen)
            inst = torch.randint(0, self.num_instances, (m,), generator=gen)
            centre = self.gt_locations[inst]
            uv, depth = self.project(centre, view)
            f = self.intrinsics[view, 0, 0]
            radius = spread * f * self.gt_half_extents[inst].norm(dim=-1) / depth.clamp_min(1.0)
            off = (torch.rand(m, 2, generator=gen) * 2 - 1) * radius[:, None]
            px = (uv + off).round().long()
            ok = (depth > 1.0) & (px[:, 0] >= 0) & (px[:, 0] < w) & (px[:, 1] >= 0) & (px[:, 1] < h)
            flat = (view * h + px[:, 1]) * w + px[:, 0]
            flat = flat[ok][:need]
            out.append(flat)
            need -= flat.numel()
        return torch.cat(out)


def make_frame(num_instances: int = 8, num_views: int = 17, image_size=KITTI360_IMAGE_SIZE,
               seed: int = 0, layout: str = "street", intrinsics_scale: float = 1.0) -> SyntheticFrame:
    gen = torch.Generator().manual_seed(seed)
    n = num_instances
    if layout == "street":
        x = torch.rand(n, generator=gen) * 12.0 - 6.0
        z = torch.rand(n, generator=gen) * 22.0 + 8.0
    elif layout == "parking":   # cfg3: dense grid with 0.5 m gaps so frusta overlap
        cols = 6
        col, row = torch.arange(n) % cols, torch.arange(n) // cols
        x = (col.float() - (cols - 1) / 2.0) * 2.5
        z = 10.0 + row.float() * 5.5
    else:
        raise ValueError(f"unknown layout {layout!r}")
    y = torch.full((n,), 0.675)
    lo, hi = torch.tensor(DIMENSION_RANGE[0]), torch.tensor(DIMENSION_RANGE[1])
    half = lo + torch.rand(n, 3, generator=gen) * (hi - lo)
    yaw = torch.rand(n, generator=gen) * 2 * math.pi - math.pi
    if layout == "parking":
        yaw = yaw * 0.05

    fx, fy, cx, cy = (v * intrinsics_scale for v in KITTI360_INTRINSICS)
    k = torch.tensor([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])
    rel = torch.arange(num_views, dtype=torch.float32) - (num_views // 2)     # -8 .. +8
    drift = (torch.rand(num_views, generator=gen) * 2 - 1) * math.radians(2.0)
    drift[num_views // 2] = 0.0
    cam_rot = rotation_y(drift)                                               # camera -> world
    cam_pos = torch.stack([torch.zeros_like(rel), torch.zeros_like(rel), rel * 1.0], dim=-1)
    e = torch.eye(4).repeat(num_views, 1, 1)
    e[:, :3, :3] = cam_rot.transpose(-2, -1)
    e[:, :3, 3] = -torch.einsum("vmn,vn->vm", cam_rot.transpose(-2, -1), cam_pos)
    return SyntheticFrame(tuple(image_size), k.repeat(num_views, 1, 1), e,
                          torch.stack([x, y, z], -1), half, yaw)


def logit_range(value: torch.Tensor, lo, hi) -> torch.Tensor:
    lo, hi = torch.as_tensor(lo), torch.as_tensor(hi)
    return torch.logit(((value - lo) / (hi - lo)).clamp(1e-4, 1 - 1e-4))


def perturbed_raw_parameters(frame: SyntheticFrame, seed: int = 0, position_noise: float = 0.5,
                             yaw_noise: float = 0.15):
    """Raw (pre-sigmoid) BoxParameters3D values for an initial guess near the ground truth."""
    gen = torch.Generator().manual_seed(seed + 1000)
    n = frame.num_instances
    loc = frame.gt_locations + torch.randn(n, 3, generator=gen) * torch.tensor([position_noise, 0.05, position_noise])
    yaw = frame.gt_yaws + torch.randn(n, generator=gen) * yaw_noise
    raw_loc = logit_range(loc, LOCATION_RANGE[0], LOCATION_RANGE[1])
    raw_dim = torch.zeros(n, 3)
    raw_ori = torch.stack([torch.cos(yaw), torch.sin(yaw)], dim=-1)
    return raw_loc, raw_dim, raw_ori


# ------------------------------------------------------------------------------------------------
# Per-frame supervision of a synthetic frame: what KITTI360Dataset + the mask transforms hand to
# scripts/main.py: soft masks, 2D boxes and source-view visibility.
# ------------------------------------------------------------------------------------------------
_CORNER_SIGNS = torch.tensor([
    (-1.0, -1.0, +1.0), (+1.0, -1.0, +1.0), (+1.0, -1.0, -1.0), (-1.0, -1.0, -1.0),
    (-1.0, +1.0, +1.0), (+1.0, +1.0, +1.0), (+1.0, +1.0, -1.0), (-1.0, +1.0, -1.0),
])
MAX_POLYGON_VERTICES = 8


def gt_corners(frame: SyntheticFrame) -> torch.Tensor:
    """Ground-truth corners [N,8,3] in the corner order of box_parameters.py:77-86."""
    local = _CORNER_SIGNS[None] * frame.gt_half_extents[:, None]
    return local @ frame.gt_rotations.transpose(-2, -1) + frame.gt_locations[:, None]


def _convex_hull(points):
    """Andrew's monotone chain on a handful of 2D points (list of (x, y)); counter-clockwise hull."""
    pts = sorted(set(points))
    if len(pts) < 3:
        return pts

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return lower[:-1] + upper[:-1]


@dataclasses.dataclass
class FrameSupervision:
    polygons: torch.Tensor        # [V,N,8,2] silhouette polygons (x, y) of the GT boxes, zero padded
    polygon_sizes: torch.Tensor   # [V,N] int32, 0 where the instance is not visible in the view
    boxes_2d: torch.Tensor        # [V,N,4] x1 y1 x2 y2 (bounding box of the visible polygon, clipped to the image)
    visible: torch.Tensor         # [V,N] bool
    target_view: int


def frame_supervision(frame: SyntheticFrame, min_depth: float = 1.0) -> FrameSupervision:
    """Projects the GT boxes into every view: the instance silhouette is the convex hull of its 8
    projected corners.  An instance is visible in a view when all corners are in front of the camera
    and its box overlaps the image."""
    h, w = frame.image_size
    v, n = frame.num_views, frame.num_instances
    corners = gt_corners(frame)                                                    # [N,8,3]
    cam = torch.einsum("vmk,nck->vncm", frame.extrinsics[:, :3, :3], corners) + frame.extrinsics[:, None, None, :3, 3]
    uvw = torch.einsum("vmk,vnck->vncm", frame.intrinsics, cam)
    uv = uvw[..., :2] / uvw[..., 2:].clamp_min(1e-6)
    polygons = torch.zeros(v, n, MAX_POLYGON_VERTICES, 2)
    sizes = torch.zeros(v, n, dtype=torch.int32)
    boxes = torch.zeros(v, n, 4)
    visible = torch.zeros(v, n, dtype=torch.bool)
    for vi in range(v):
        for ni in range(n):
            if float(cam[vi, ni, :, 2].min()) < min_depth:
                continue
            hull = _convex_hull([(float(x), float(y)) for x, y in uv[vi, ni].tolist()])
            if len(hull) < 3:
                continue
            p = torch.tensor(hull)
            x1, y1 = float(p[:, 0].min()), float(p[:, 1].min())
            x2, y2 = float(p[:, 0].max()), float(p[:, 1].max())
            if x2 <= 0 or y2 <= 0 or x1 >= w or y1 >= h:
                continue
            polygons[vi, ni, :len(hull)] = p
            sizes[vi, ni] = len(hull)
            boxes[vi, ni] = torch.tensor([max(x1, 0.0), max(y1, 0.0), min(x2, float(w)), min(y2, float(h))])
            visible[vi, ni] = True
    return FrameSupervision(polygons, sizes, boxes, visible, target_view=v // 2)
