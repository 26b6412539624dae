"""Inference / logging renderers on the B200 kernels (scripts/main.py:1011-1041):

  * `union_field`     the composed soft-union field at arbitrary points: distance, spatial gradient, weights
  * `sphere_trace`    vsrd.rendering.sphere_tracing (rendering/renderers.py:21-76) for a `UnionField` scene
  * `surface_normals` vsrd.rendering.surface_normal (renderers.py:79-113): the normalised union gradient,
                      analytic (the kernels return d and grad d together) instead of an autograd call
  * `render_image`    the full-image two-pass volumetric render main.py runs row by row at image_intervals
                      (main.py:1011-1024), here in ray chunks sized for the [N, R*M] field buffer

Everything runs under no_grad: these paths only feed TensorBoard images and the exported masks.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import functional as F
from . import ops
from .ops import SceneArgs


def _scene(field, temperature=None) -> SceneArgs:
    """`field` is a vsrd.rendering.renderers.UnionField (or anything with its attributes)."""
    w = field.mlp_weights
    return SceneArgs(field.locations.detach().float(), field.rotations.detach().float(), field.half_extents.detach().float(),
                     None if w is None else w.detach().float(),
                     field.temperature if temperature is None else temperature, field.scale)


@torch.no_grad()
def union_field(field, points: torch.Tensor, want_weights: bool = False):
    """points [..., 3] -> (distance [..., 1], gradient [..., 3], weights [..., N] or None)."""
    scene = _scene(field)
    lead = points.shape[:-1]
    pts = points.detach().float().reshape(-1, 3).contiguous()
    per_instance = ops.field_points(scene, pts)
    out, weights = ops.union_points(scene, per_instance, want_weights)
    return (out[:, :1].reshape(*lead, 1), out[:, 1:].reshape(*lead, 3),
            None if weights is None else weights.reshape(*lead, -1))


def sphere_intersection(ray_positions, ray_directions, bounding_radius):
    """rendering/renderers.py:10-18."""
    a = torch.sum(ray_directions * ray_directions, dim=-1, keepdim=True)
    b = torch.sum(ray_directions * ray_positions, dim=-1, keepdim=True)
    c = torch.sum(ray_positions * ray_positions, dim=-1, keepdim=True) - bounding_radius ** 2.0
    d = b ** 2.0 - a * c
    masks = d >= 0.0
    return (-b - torch.sqrt(d)) / a, (-b + torch.sqrt(d)) / a, masks


@torch.no_grad()
def sphere_trace(field, ray_positions: torch.Tensor, ray_directions: torch.Tensor, num_iterations: int,
                 convergence_criteria: float, foreground_masks: Optional[torch.Tensor] = None,
                 bounding_radius: Optional[float] = None, initialization: bool = True,
                 differentiable: bool = False, poll_every: int = 16) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (surface positions [..., 3], convergence masks [..., 1] bool) exactly as the reference loop
    does, including its global early exit.  The loop runs on the device: per iteration one field launch, one
    union launch and one update launch; the host only polls the active-ray counter every `poll_every`
    iterations to stop enqueueing work (the reference synchronises every iteration, renderers.py:55)."""
    scene = _scene(field)
    lead = torch.broadcast_shapes(ray_positions.shape[:-1], ray_directions.shape[:-1])
    dev = ray_directions.device
    positions = ray_positions.detach().float().expand(*lead, 3)
    directions = ray_directions.detach().float().expand(*lead, 3)
    if foreground_masks is None:
        foreground_masks = torch.all(torch.isfinite(positions), dim=-1, keepdim=True)
    foreground_masks = foreground_masks.expand(*lead, 1)
    if bounding_radius and initialization:
        near, _, hit = sphere_intersection(positions, directions, bounding_radius)
        positions = torch.where(hit, positions + directions * near, positions)
        foreground_masks = foreground_masks & hit

    pos = positions.reshape(-1, 3).contiguous().clone()
    dirs = directions.reshape(-1, 3).contiguous()
    fg = foreground_masks.reshape(-1).to(torch.uint8).contiguous().clone()
    conv = torch.zeros_like(fg)
    active = torch.zeros(max(int(num_iterations), 1), dtype=torch.int32, device=dev)
    for it in range(int(num_iterations)):
        per_instance = ops.field_points(scene, pos)
        out, _ = ops.union_points(scene, per_instance)
        ops.sphere_trace_step(out, dirs, pos, fg, conv, active, it, convergence_criteria, bounding_radius)
        if (it + 1) % poll_every == 0 and int(active[it]) == 0:
            break
    if differentiable:
        # renderers.py:57-71: one Newton step along the ray at the converged positions.  The kernels return the
        # spatial gradient with the value, so no autograd call is needed.  The result is DETACHED: main.py's photometric
        # branch (:742-754, differentiable=True, loss weight 0.0 in every shipped config) would get no gradient, so
        # vsrd.rendering.sphere_tracing refuses that call when a field parameter requires grad; the logging call
        # (:1026-1038) uses differentiable=False.
        per_instance = ops.field_points(scene, pos)
        out, _ = ops.union_points(scene, per_instance)
        step = -out[:, :1] / torch.sum(out[:, 1:] * dirs, dim=-1, keepdim=True)
        pos = torch.where(conv.bool()[:, None], pos + dirs * step, pos)
    return pos.reshape(*lead, 3), conv.bool().reshape(*lead, 1)


@torch.no_grad()
def surface_normals(field, surface_positions: torch.Tensor, finite_difference_epsilon: Optional[float] = None) -> torch.Tensor:
    """renderers.py:79-113.  The default branch is the analytic union gradient; the finite-difference branch
    evaluates the six shifted positions like the reference."""
    if finite_difference_epsilon:
        eye = torch.eye(3, device=surface_positions.device, dtype=torch.float32) * float(finite_difference_epsilon)
        parts = [union_field(field, surface_positions + e)[0] - union_field(field, surface_positions - e)[0] for e in eye]
        normals = torch.cat(parts, dim=-1)
    else:
        normals = union_field(field, surface_positions)[1]
    return torch.nn.functional.normalize(normals, dim=-1)


@torch.no_grad()
def render_image(field, camera_position: torch.Tensor, ray_directions: torch.Tensor, *, distance_range=(0.0, 100.0),
                 num_samples: int = 100, std_deviation: float, cosine_ratio: float = 1.0, epsilon: float = 1e-6,
                 seed: int = 0, max_rays_per_chunk: int = 1 << 16, jitter: Optional[torch.Tensor] = None,
                 sorted_uniforms: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Soft instance labels of every pixel: ray_directions [H,W,3] (or [...,3]) -> [H,W,N]
    (main.py:1011-1024 renders the same thing one image row per call).  Chunked so that the per-instance
    field buffer [N, chunk * (2S-1)] float4 stays bounded (N=8, S=100, 65 536 rays: 1.7 GB)."""
    scene = _scene(field)
    lead = ray_directions.shape[:-1]
    dirs = ray_directions.detach().float().reshape(-1, 3).contiguous()
    origin = camera_position.detach().float().reshape(-1, 3)
    if origin.shape[0] not in (1, dirs.shape[0]):
        raise RuntimeError("vsrd_b200: camera_position must be [3] or match ray_directions")
    bins = F.distance_bins(distance_range, num_samples, dirs.device)
    out = torch.empty(dirs.shape[0], scene.num_instances, device=dirs.device, dtype=torch.float32)
    for start in range(0, dirs.shape[0], max_rays_per_chunk):
        sl = slice(start, min(start + max_rays_per_chunk, dirs.shape[0]))
        o = origin if origin.shape[0] == 1 else origin[sl]
        labels, *_ = F.two_pass_render(
            scene.locations, scene.rotations, scene.half_extents, scene.mlp_weights, o.expand(sl.stop - sl.start, 3).contiguous(),
            dirs[sl], num_samples=num_samples, temperature=scene.temperature, std_deviation=std_deviation,
            cosine_ratio=cosine_ratio, epsilon=epsilon, scale=scene.scale, bins=bins,
            jitter=None if jitter is None else jitter[sl], sorted_uniforms=None if sorted_uniforms is None else sorted_uniforms[sl],
            seed=seed + start)
        out[sl] = labels
    return out.reshape(*lead, -1)
