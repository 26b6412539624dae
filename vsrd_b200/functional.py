"""Autograd-level API of the B200 silhouette renderer.

`render_pass` is the differentiable core: everything `hierarchical_volumetric_rendering` does after
sample placement (vsrd/rendering/renderers.py:212-270) for a soft union of box(+residual) instances
(scripts/main.py:433-618), as one custom autograd Function whose forward and backward are the
hand-written kernels.  The backward replaces the reference's double-backward graph replay
(renderers.py:226 `create_graph=True`) with an analytic adjoint.

Layouts are ray-major ([R, M, ...]); `vsrd.rendering` exposes the reference's sample-major views.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops
from .ops import RayArgs, SceneArgs


class _RenderPass(torch.autograd.Function):

    @staticmethod
    def forward(ctx, locations, rotations, half_extents, mlp_weights, origins, directions, distances,
                temperature, scale, std_deviation, cosine_ratio, epsilon, step_state=None, cull=None):
        scene = SceneArgs(locations.detach(), rotations.detach(), half_extents.detach(),
                          None if mlp_weights is None else mlp_weights.detach(), temperature, scale, step_state)
        rays = RayArgs(origins.detach(), directions.detach(), distances.detach())
        field = ops.field_forward(scene, rays, cull=cull)
        labels, grads, weights, _ = ops.composite_forward(scene, rays, field, std_deviation, cosine_ratio, epsilon)
        ctx.scene, ctx.rays, ctx.field = scene, rays, field
        ctx.render = (std_deviation, cosine_ratio, epsilon)
        ctx.has_mlp = mlp_weights is not None
        return labels, grads, weights

    @staticmethod
    def backward(ctx, grad_labels, grad_gradients, grad_weights):
        std_deviation, cosine_ratio, epsilon = ctx.render
        adjoint = ops.composite_backward(ctx.scene, ctx.rays, ctx.field, std_deviation, cosine_ratio, epsilon,
                                         grad_labels=grad_labels, grad_gradients=grad_gradients,
                                         grad_weights=grad_weights)
        g_loc, g_rot, g_dim, g_w = ops.field_backward(ctx.scene, ctx.rays, adjoint)
        return g_loc, g_rot, g_dim, (g_w if ctx.has_mlp else None), None, None, None, None, None, None, None, None, None, None


def render_pass(locations, rotations, half_extents, mlp_weights, ray_positions, ray_directions, distances, *,
                temperature: float, std_deviation: float, cosine_ratio: float = 1.0, epsilon: float = 1e-6,
                scale: float = 100.0, step_state=None):
    """Union field + SDF->opacity + compositing at the given sample distances.

    locations [N,3], rotations [N,3,3], half_extents [N,3], mlp_weights [N,1617] or None,
    ray_positions [R,3] (or [3]), ray_directions [R,3], distances [R,M+1] (ascending, detached).
    Returns labels [R,N], gradients [R,M,3] (un-normalised union gradient), weights [R,M];
    differentiable w.r.t. the first four arguments.  `step_state` (ops.StepState) makes the kernels read
    temperature / std_deviation / cosine_ratio from device memory instead (CUDA-graph replay across steps).
    """
    # no-grad calls are the coarse placement passes of the two-pass wrapper (main.py:515-516): culling off there
    # (ops.field_forward); decided here because grad mode is always off inside Function.forward
    cull = None if torch.is_grad_enabled() else False
    return _RenderPass.apply(locations, rotations, half_extents, mlp_weights, ray_positions, ray_directions,
                             distances, float(temperature), float(scale), float(std_deviation),
                             float(cosine_ratio), float(epsilon), step_state, cull)


def distance_bins(distance_range: Sequence[float], num_samples: int, device) -> torch.Tensor:
    """`torch.linspace(*distance_range, num_samples + 1)` (renderers.py:191), computed on the host in
    float32 exactly as the CPU reference does, then moved to the device."""
    return torch.linspace(float(distance_range[0]), float(distance_range[1]), num_samples + 1,
                          dtype=torch.float32).to(device)


def two_pass_render(locations, rotations, half_extents, mlp_weights, ray_positions, ray_directions, *,
                    distance_range=(0.0, 100.0), num_samples: int, temperature: float, std_deviation: float,
                    cosine_ratio: float = 1.0, epsilon: float = 1e-6, scale: float = 100.0,
                    jitter: Optional[torch.Tensor] = None, sorted_uniforms: Optional[torch.Tensor] = None,
                    seed: int = 0, bins: Optional[torch.Tensor] = None):
    """`hierarchical_wrapper` (scripts/main.py:511-523): no-grad coarse pass, importance resampling,
    differentiable fine pass.  Returns (labels, gradients, coarse_distances, coarse_weights,
    fine_distances, fine_weights), all ray-major."""
    device = ray_directions.device
    num_rays = ray_directions.reshape(-1, 3).shape[0]
    if bins is None:
        bins = distance_bins(distance_range, num_samples, device)
    kw = dict(temperature=temperature, std_deviation=std_deviation, cosine_ratio=cosine_ratio,
              epsilon=epsilon, scale=scale)
    with torch.no_grad():
        coarse = ops.place_coarse(bins, num_rays, jitter, seed)
        _, _, coarse_w = render_pass(locations, rotations, half_extents, mlp_weights, ray_positions,
                                     ray_directions, coarse, **kw)
        fine = ops.place_fine(coarse, coarse_w, sorted_uniforms, seed)
    labels, grads, fine_w = render_pass(locations, rotations, half_extents, mlp_weights, ray_positions,
                                        ray_directions, fine, **kw)
    return labels, grads, coarse, coarse_w, fine, fine_w


class _FusedRenderLoss(torch.autograd.Function):
    """Fine pass with the silhouette BCE + eikonal reduction fused into the compositing kernels
    (scripts/main.py:653-687, 855).  Forward returns (loss, labels)."""

    @staticmethod
    def forward(ctx, locations, rotations, half_extents, mlp_weights, origins, directions, distances, targets,
                temperature, scale, std_deviation, cosine_ratio, epsilon, silhouette_weight, eikonal_weight,
                step_state=None):
        scene = SceneArgs(locations.detach(), rotations.detach(), half_extents.detach(),
                          None if mlp_weights is None else mlp_weights.detach(), temperature, scale, step_state)
        rays = RayArgs(origins.detach(), directions.detach(), distances.detach())
        field = ops.field_forward(scene, rays)
        labels, _, _, loss_parts = ops.composite_forward(
            scene, rays, field, std_deviation, cosine_ratio, epsilon,
            targets=targets, silhouette_weight=silhouette_weight, eikonal_weight=eikonal_weight)
        ctx.scene, ctx.rays, ctx.field, ctx.labels, ctx.targets = scene, rays, field, labels, targets
        ctx.render = (std_deviation, cosine_ratio, epsilon, silhouette_weight, eikonal_weight)
        ctx.has_mlp = mlp_weights is not None
        ctx.mark_non_differentiable(labels, loss_parts)
        return loss_parts.sum(), labels, loss_parts

    @staticmethod
    def backward(ctx, grad_loss, _grad_labels, _grad_parts):
        std_deviation, cosine_ratio, epsilon, sil_w, eik_w = ctx.render
        adjoint = ops.composite_backward(ctx.scene, ctx.rays, ctx.field, std_deviation, cosine_ratio, epsilon,
                                         targets=ctx.targets, labels=ctx.labels,
                                         silhouette_weight=sil_w, eikonal_weight=eik_w)
        g_loc, g_rot, g_dim, g_w = ops.field_backward(ctx.scene, ctx.rays, adjoint)
        scale = grad_loss
        return (g_loc * scale, g_rot * scale, g_dim * scale, (g_w * scale if ctx.has_mlp else None),
                None, None, None, None, None, None, None, None, None, None, None, None)


def fused_render_loss(locations, rotations, half_extents, mlp_weights, ray_positions, ray_directions, distances,
                      targets, *, temperature: float, std_deviation: float, cosine_ratio: float = 1.0,
                      epsilon: float = 1e-6, scale: float = 100.0, silhouette_weight: float = 1.0,
                      eikonal_weight: float = 0.01, step_state=None):
    """loss = silhouette_weight * mean BCE(clamp(labels, 1e-6, 1-1e-6), targets)
            + eikonal_weight * mean((|grad| - 1)^2), computed inside the compositing kernel.
    Returns (loss, labels [R,N], loss_parts [2])."""
    return _FusedRenderLoss.apply(locations, rotations, half_extents, mlp_weights, ray_positions, ray_directions,
                                  distances, targets, float(temperature), float(scale), float(std_deviation),
                                  float(cosine_ratio), float(epsilon), float(silhouette_weight), float(eikonal_weight),
                                  step_state)


def render_step(locations, rotations, half_extents, mlp_weights, ray_positions, ray_directions, targets, *,
                bins: torch.Tensor, temperature: float, std_deviation: float, cosine_ratio: float,
                epsilon: float = 1e-6, scale: float = 100.0, silhouette_weight: float = 1.0,
                eikonal_weight: Optional[float] = None, jitter=None, sorted_uniforms=None, seed: int = 0,
                step_state=None):
    """One renderer step of the optimisation loop (scripts/main.py:629-687): coarse pass, resampling,
    fine pass and the fused loss.  Returns (loss, labels, loss_parts)."""
    if eikonal_weight is None:
        eikonal_weight = 0.01 if mlp_weights is not None else 0.0   # main.py:677: only with the residual field
    num_rays = ray_directions.reshape(-1, 3).shape[0]
    kw = dict(temperature=temperature, std_deviation=std_deviation, cosine_ratio=cosine_ratio,
              epsilon=epsilon, scale=scale, step_state=step_state)
    with torch.no_grad():
        coarse = ops.place_coarse(bins, num_rays, jitter, seed, step_state)
        _, _, coarse_w = render_pass(locations, rotations, half_extents, mlp_weights, ray_positions,
                                     ray_directions, coarse, **kw)
        fine = ops.place_fine(coarse, coarse_w, sorted_uniforms, seed, step_state)
    return fused_render_loss(locations, rotations, half_extents, mlp_weights, ray_positions, ray_directions,
                             fine, targets, silhouette_weight=silhouette_weight,
                             eikonal_weight=eikonal_weight, **kw)
