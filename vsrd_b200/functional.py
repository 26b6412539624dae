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
                temperature, scale, std_deviation, cosine_ratio, epsilon, step_state=None, differentiable=True):
        scene = SceneArgs(locations.detach(), rotations.detach(), half_extents.detach(),
                          None if mlp_weights is None else mlp_weights.detach(), temperature, scale, step_state)
        rays = RayArgs(origins.detach(), directions.detach(), distances.detach())
        field = ops.field_forward(scene, rays, backward=differentiable)
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
    # no-grad calls are the coarse placement passes of the two-pass wrapper (main.py:515-516): no backward tile marks
    # there (ops.field_forward); decided here because grad mode is always off inside Function.forward
    differentiable = torch.is_grad_enabled()
    return _RenderPass.apply(locations, rotations, half_extents, mlp_weights, ray_positions, ray_directions,
                             distances, float(temperature), float(scale), float(std_deviation),
                             float(cosine_ratio), float(epsilon), step_state, differentiable)


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


# ---- models under autograd (a3 / a4): the nn.Modules of `vsrd.models` call these on CUDA parameters -----------------
_fused_modules = True


def set_fused_modules(enabled: bool) -> None:
    """`vsrd.models.BoxParameters3D` / `HyperDistanceField` on CUDA run through the model kernels (default) or as plain
    PyTorch ops (the reference's own formulation; the yardstick of tests/test_gpu_models.py)."""
    global _fused_modules
    _fused_modules = bool(enabled)


def fused_modules() -> bool:
    return _fused_modules


def _hyper_layers(module):
    """[(linear, layer_norm or None)] of a HyperDistanceField.hypernetwork the kernels can run, else None."""
    from . import _lib
    blocks = list(module.hypernetwork)
    if not 2 <= len(blocks) <= _lib.HYPER_MAX_LAYERS:
        return None
    layers = []
    for index, block in enumerate(blocks):
        linear = block[0]
        norm = block[1] if len(block) > 1 else None
        last = index == len(blocks) - 1
        ok = (isinstance(linear, torch.nn.Linear) and hasattr(linear, "weight_v") and hasattr(linear, "weight_g")
              and linear.in_features == _lib.HYPER_WIDTH and linear.bias is not None
              and (last or (isinstance(norm, torch.nn.LayerNorm) and linear.out_features == _lib.HYPER_WIDTH
                            and norm.elementwise_affine and len(block) == 3
                            and isinstance(block[2], torch.nn.GELU) and getattr(block[2], "approximate", "none") == "none"))
              and (not last or len(block) == 1)
              and linear.weight_v.dtype == torch.float32)
        if not ok:
            return None
        layers.append((linear, norm))
    return layers


def _hyper_tables(tensors, grads=None):
    """VsrdHyperNet (and VsrdHyperNetGrads) over per-layer tuples (weight_v, weight_g, bias, ln_weight, ln_bias)."""
    from . import _lib
    net = _lib.VsrdHyperNet()
    net.num_layers = len(tensors)
    table = None
    if grads is not None:
        table = _lib.VsrdHyperNetGrads()
        table.num_layers = len(tensors)
    for l, (v, g, b, lw, lb) in enumerate(tensors):
        L = net.layers[l]
        L.weight_v, L.weight_g, L.bias = v.data_ptr(), g.data_ptr(), b.data_ptr()
        L.in_features, L.out_features = int(v.shape[1]), int(v.shape[0])
        if lw is not None:
            L.ln_weight, L.ln_bias = lw.data_ptr(), lb.data_ptr()
        if table is not None:
            gv, gg, gb, glw, glb = grads[l]
            G = table.layers[l]
            G.weight_v, G.weight_g, G.bias = gv.data_ptr(), gg.data_ptr(), gb.data_ptr()
            if glw is not None:
                G.ln_weight, G.ln_bias = glw.data_ptr(), glb.data_ptr()
    return net, table


class _Hypernetwork(torch.autograd.Function):
    """HyperDistanceField.forward (hyper_distance_field.py:75-77) as 5 launches, its backward as 6 (csrc/vsrd_model.cu),
    instead of ~110 ATen launches.  Inputs: embeddings [..., 256], then per layer weight_v, weight_g, bias[, ln.weight, ln.bias]."""

    @staticmethod
    def forward(ctx, embeddings, has_norm, *flat):
        tensors, k = [], 0
        for norm in has_norm:
            v, g, b = flat[k], flat[k + 1], flat[k + 2]
            lw, lb = (flat[k + 3], flat[k + 4]) if norm else (None, None)
            k += 5 if norm else 3
            tensors.append(tuple(None if t is None else t.detach().contiguous() for t in (v, g, b, lw, lb)))
        emb = embeddings.detach().contiguous().reshape(-1, embeddings.shape[-1])
        net, _ = _hyper_tables(tensors)
        ctx.lead = embeddings.shape[:-1]
        out = torch.empty(*ctx.lead, int(tensors[-1][0].shape[0]), device=emb.device, dtype=torch.float32)   # final shape, see _DecodeBoxes
        _, activations = ops.hyper_forward(net, emb, mlp_weights=out)
        ctx.tensors, ctx.has_norm, ctx.emb, ctx.activations = tensors, has_norm, emb, activations
        return out

    @staticmethod
    def backward(ctx, grad_weights):
        grads = [tuple(None if t is None else torch.empty_like(t) for t in layer) for layer in ctx.tensors]
        net, table = _hyper_tables(ctx.tensors, grads)
        g_emb = torch.empty_like(ctx.emb)
        ops.hyper_backward(net, table, ctx.emb, ctx.activations, grad_weights.contiguous().reshape(ctx.emb.shape[0], -1), g_emb)
        flat = []
        for norm, (gv, gg, gb, glw, glb) in zip(ctx.has_norm, grads):
            flat += [gv, gg, gb] + ([glw, glb] if norm else [])
        return (g_emb.reshape(*ctx.lead, g_emb.shape[-1]), None, *flat)


def hypernetwork(module, embeddings):
    """`module.hypernetwork(embeddings)` through the kernels, or None when the module / input is not what they are
    compiled for (other widths, more than 32 embedding rows, non-fp32): the caller then runs the nn.Sequential."""
    from . import _lib
    layers = _hyper_layers(module)
    rows = embeddings.numel() // max(1, embeddings.shape[-1])
    if (layers is None or embeddings.dtype != torch.float32 or embeddings.shape[-1] != _lib.HYPER_WIDTH
            or not 1 <= rows <= _lib.MAX_INSTANCES):
        return None
    flat, has_norm = [], []
    for linear, norm in layers:
        flat += [linear.weight_v, linear.weight_g, linear.bias]
        has_norm.append(norm is not None)
        if norm is not None:
            flat += [norm.weight, norm.bias]
    return _Hypernetwork.apply(embeddings, tuple(has_norm), *flat)


class _DecodeBoxes(torch.autograd.Function):
    """BoxParameters3D.forward (box_parameters.py:124-146) and its backward as one launch each."""

    @staticmethod
    def forward(ctx, raw_locations, raw_dimensions, raw_orientations, ranges):
        lead = raw_locations.shape[:-1]
        raw = [t.detach().contiguous().reshape(-1, t.shape[-1]) for t in (raw_locations, raw_dimensions, raw_orientations)]
        # outputs are allocated in their final shape: a re-viewed output would carry a `_base` outside the autograd graph,
        # which the renderer's closure matcher (vsrd.rendering.renderers._gather_rows) relies on
        loc, dim, rot, boxes = ops.decode_boxes(ranges, *raw, lead=lead)
        ctx.raw, ctx.ranges, ctx.lead, ctx.decoded = raw, ranges, lead, (dim, rot)
        return loc, dim, rot, boxes

    @staticmethod
    def backward(ctx, g_loc, g_dim, g_rot, g_boxes):
        n = ctx.raw[0].shape[0]
        dev = ctx.raw[0].device
        zeros = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)
        g_loc = zeros(n, 3) if g_loc is None else g_loc.contiguous().reshape(n, 3)
        g_dim = zeros(n, 3) if g_dim is None else g_dim.contiguous().reshape(n, 3)
        g_rot = zeros(n, 9) if g_rot is None else g_rot.contiguous().reshape(n, 9)
        pair = None
        if g_boxes is not None:
            pair = zeros(2, n, 24)
            pair[0].copy_(g_boxes.reshape(n, 24))
        out = [torch.empty_like(t) for t in ctx.raw]
        dim, rot = ctx.decoded
        ops.decode_boxes_backward(ctx.ranges, *ctx.raw, dim, rot, g_loc, g_dim, g_rot, pair, 1.0, 0.0, *out)
        return out[0].reshape(*ctx.lead, 3), out[1].reshape(*ctx.lead, 3), out[2].reshape(*ctx.lead, 2), None


def decode_boxes(raw_locations, raw_dimensions, raw_orientations, ranges):
    """Differentiable BoxParameters3D decode on CUDA tensors; `ranges` is a `_lib.VsrdBoxRanges`.
    Returns (locations [...,3], half extents [...,3], rotations [...,3,3], corners [...,8,3])."""
    return _DecodeBoxes.apply(raw_locations, raw_dimensions, raw_orientations, ranges)
