"""Tensor-level wrappers around the C ABI (include/vsrd_b200.h).  No autograd here.

Tensors are borrowed: every input must be a CUDA float32 tensor (contiguous copies are made when
needed) and outputs are allocated with torch so they live in its caching allocator.  Kernels are
enqueued on torch's current stream, so the calls compose with CUDA graphs and stream contexts.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import VsrdLoss, VsrdRays, VsrdRenderParams, VsrdSchedule, VsrdScene, VsrdStepState, VsrdViews

MLP_WEIGHTS = _lib.MLP_WEIGHTS
GRAD_STRIDE = _lib.GRAD_STRIDE


def _stream() -> ctypes.c_void_p:
    # the raw handle of torch's current stream: ~1 us, where torch.cuda.current_stream() builds a Stream object through
    # several Python frames (~20 us; 2 x 136 project_box_3d calls per step of scripts/main.py made that 5 ms per step)
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"vsrd_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"vsrd_b200: {name} must be float32, got {t.dtype}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class SceneArgs:
    """Keeps the tensors alive and exposes the C struct."""

    def __init__(self, locations, rotations, half_extents, mlp_weights, temperature, scale=100.0, step_state=None):
        self.locations = _f32(locations, "locations").reshape(-1, 3)
        n = self.locations.shape[0]
        self.rotations = _f32(rotations, "rotations").reshape(n, 3, 3)
        self.half_extents = _f32(half_extents, "half_extents").reshape(n, 3)
        self.mlp_weights = None
        if mlp_weights is not None:
            self.mlp_weights = _f32(mlp_weights, "mlp_weights")
            if tuple(self.mlp_weights.shape) != (n, MLP_WEIGHTS):
                raise RuntimeError(
                    f"vsrd_b200: mlp_weights must be [{n}, {MLP_WEIGHTS}] (48-16-16-16-16-1 residual field), "
                    f"got {tuple(self.mlp_weights.shape)}")
        if not 1 <= n <= _lib.MAX_INSTANCES:
            raise RuntimeError(f"vsrd_b200: number of instances must be in [1, {_lib.MAX_INSTANCES}], got {n}")
        self.num_instances = n
        self.temperature = float(temperature)
        self.scale = float(scale)
        self.step_state = step_state          # StepState or None: device-resident schedule overrides the scalars
        self.struct = VsrdScene(n, 0, _ptr(self.locations), _ptr(self.rotations), _ptr(self.half_extents),
                                _ptr(self.mlp_weights), self.temperature, self.scale,
                                None if step_state is None else step_state.ptr)


class RayArgs:
    def __init__(self, origins, directions, distances):
        self.directions = _f32(directions, "ray_directions").reshape(-1, 3)
        r = self.directions.shape[0]
        origins = _f32(origins, "ray_positions")
        self.origins = origins.expand(r, 3).contiguous() if origins.numel() == 3 else origins.reshape(r, 3)
        self.distances = _f32(distances, "distances")
        if self.distances.dim() != 2 or self.distances.shape[0] != r or self.distances.shape[1] < 2:
            raise RuntimeError(f"vsrd_b200: distances must be [R, M+1] with R={r}, got {tuple(self.distances.shape)}")
        self.num_rays = r
        self.num_intervals = self.distances.shape[1] - 1
        if self.num_intervals > _lib.MAX_INTERVALS:
            raise RuntimeError(f"vsrd_b200: at most {_lib.MAX_INTERVALS} intervals per ray, got {self.num_intervals}")
        self.forward_samples = None   # live-sample lists of the forward kernel and backward tile marks (enable_culling)
        self.live_tiles = None
        self.struct = VsrdRays(r, self.num_intervals, _ptr(self.origins), _ptr(self.directions), _ptr(self.distances), None, None, None)

    def enable_backward_culling(self, scene: "SceneArgs") -> None:
        """Attaches the tile marks of the backward kernels (include/vsrd_b200.h, VsrdRays::live_tiles): vsrd_composite_backward
        zeroes the adjoints of instances whose soft-min weight is below exp(-VSRD_CULL_LOG_EPS) and marks the tiles that
        still carry one; vsrd_field_backward visits the marked tiles only.  No extra launch."""
        if self.num_rays == 0 or self.live_tiles is not None:
            return
        size = _lib.load().vsrd_live_tiles_bytes(scene.num_instances, self.num_rays, self.num_intervals)
        if size < 1:
            _lib.check(1)
        self.live_tiles = torch.zeros(size, device=self.directions.device, dtype=torch.uint8)
        self.struct.live_tiles = _ptr(self.live_tiles)
        self.struct.cull_stats = _ptr(_cull_stats(self.directions.device))

    def cull_forward(self, scene: "SceneArgs", field: torch.Tensor) -> None:
        """Runs the culling pre-pass for `field` (one launch of ~15-25 us at R = 1000: the box field of the (sample,
        instance) pairs whose soft-min weight is provably < exp(-20) is written here, the other samples are listed per
        instance) and attaches the lists: the residual field kernel then evaluates the listed samples only
        (include/vsrd_b200.h, VsrdRays::forward_samples)."""
        if self.num_rays == 0:
            return
        dev = self.directions.device
        samples = self.num_rays * self.num_intervals
        if self.forward_samples is None:
            self.forward_samples = torch.empty(_lib.CULL_HEADER_INTS + scene.num_instances * samples, device=dev, dtype=torch.int32)
            self.struct.cull_stats = _ptr(_cull_stats(dev))
        _lib.check(_lib.load().vsrd_cull_samples(ctypes.byref(scene.struct), ctypes.byref(self.struct), _ptr(field),
                                                 _ptr(self.forward_samples), _stream()))
        self.struct.forward_samples = _ptr(self.forward_samples)
        self.num_instances = scene.num_instances

    def live_pairs(self):
        """(live, total) (sample, instance) pairs of the last culling pre-pass on these rays (synchronises); None
        without culling."""
        if self.forward_samples is None:
            return None
        live = int(self.forward_samples[:_lib.CULL_HEADER_INTS:_lib.CULL_COUNT_STRIDE][:self.num_instances].sum())
        return live, self.num_instances * self.num_rays * self.num_intervals


# ---- instance culling -----------------------------------------------------------------
_culling = os.environ.get("VSRD_CULL", "1") != "0"
_cull_counters = {}
# Culling costs ~25 us (backward: marks, census, tile lists) + ~35 us (forward pre-passes) per two-pass step at R = 1000
# (profiles/r02_culling_sweep.txt) and culls nothing while 1 + 20 T exceeds the distances between the boxes of a street
# scene (break-even at T ~ 0.5 on the synthetic KITTI-360 frames): it is active once the temperature is at most
CULL_MAX_TEMPERATURE = 0.5


def set_culling(enabled: bool, max_temperature: Optional[float] = None) -> None:
    """Instance culling in the residual field kernels (on by default; VSRD_CULL=0 turns it off at import);
    `max_temperature` replaces CULL_MAX_TEMPERATURE (tests pass inf to cull at every temperature)."""
    global _culling, CULL_MAX_TEMPERATURE
    _culling = bool(enabled)
    if max_temperature is not None:
        CULL_MAX_TEMPERATURE = float(max_temperature)


def culling_enabled() -> bool:
    return _culling


def _cull_stats(device) -> torch.Tensor:
    key = torch.device(device).index or 0
    if key not in _cull_counters:
        _cull_counters[key] = torch.zeros(4, dtype=torch.int64, device=device)
    return _cull_counters[key]


def culling_counters(device="cuda", reset: bool = False, forward: bool = False):
    """(skipped, visited) accumulated on `device` (synchronises): tiles of the backward field kernel
    (vsrd_backward_tile_rows() samples each), or with `forward` the (sample, instance) pairs of the culling pre-pass."""
    t = _cull_stats(torch.device(device))
    v = [int(x) for x in t.cpu()]
    if reset:
        t.zero_()
    return (v[2], v[3]) if forward else (v[0], v[1])


def _state_ptr(step_state) -> Optional[int]:
    return None if step_state is None else step_state.ptr


def _params(std_deviation, cosine_ratio, epsilon) -> VsrdRenderParams:
    return VsrdRenderParams(float(std_deviation), float(cosine_ratio), float(epsilon), 0.0)


def _loss(targets: Optional[torch.Tensor], silhouette_weight: float, eikonal_weight: float, r: int, n: int):
    if targets is None:
        return None, None
    targets = _f32(targets, "targets")
    if tuple(targets.shape) != (r, n):
        raise RuntimeError(f"vsrd_b200: targets must be [{r}, {n}], got {tuple(targets.shape)}")
    return targets, VsrdLoss(_ptr(targets), float(silhouette_weight), float(eikonal_weight))


# ---- a1 ---------------------------------------------------------------------------------------

def ray_directions(inv_projection: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """inv_projection [V,3,3] = inv(E)[:3,:3] @ inv(K)  ->  unit directions [V,H,W,3]."""
    p = _f32(inv_projection, "inv_projection").reshape(-1, 3, 3)
    out = torch.empty(p.shape[0], height, width, 3, device=p.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_ray_directions(_ptr(p), p.shape[0], height, width, _ptr(out), _stream()))
    return out


def gather_rays(inv_projection, camera_positions, pixel_indices, height, width) -> Tuple[torch.Tensor, torch.Tensor]:
    p = _f32(inv_projection, "inv_projection").reshape(-1, 3, 3)
    c = _f32(camera_positions, "camera_positions").reshape(-1, 3)
    if pixel_indices.dtype != torch.int64 or not pixel_indices.is_cuda:
        raise RuntimeError("vsrd_b200: pixel_indices must be a CUDA int64 tensor")
    idx = pixel_indices.contiguous().reshape(-1)
    origins = torch.empty(idx.numel(), 3, device=p.device, dtype=torch.float32)
    dirs = torch.empty_like(origins)
    _lib.check(_lib.load().vsrd_gather_rays(_ptr(p), _ptr(c), _ptr(idx), idx.numel(), p.shape[0], height, width,
                                            _ptr(origins), _ptr(dirs), _stream()))
    return origins, dirs


# ---- a9 / a10 ---------------------------------------------------------------------------------

def place_coarse(bins: torch.Tensor, num_rays: int, jitter: Optional[torch.Tensor] = None, seed: int = 0,
                 step_state=None) -> torch.Tensor:
    bins = _f32(bins, "bins").reshape(-1)
    s = bins.numel() - 1
    if jitter is not None:
        jitter = _f32(jitter, "jitter").reshape(num_rays, s)
    out = torch.empty(num_rays, s, device=bins.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_place_coarse(_ptr(bins), _ptr(jitter), seed & (2 ** 64 - 1), _state_ptr(step_state),
                                             num_rays, s, _ptr(out), _stream()))
    return out


def place_fine(coarse_distances, coarse_weights, sorted_uniforms: Optional[torch.Tensor] = None, seed: int = 0,
               step_state=None) -> torch.Tensor:
    t = _f32(coarse_distances, "coarse_distances")
    w = _f32(coarse_weights, "coarse_weights")
    r, s = t.shape
    if tuple(w.shape) != (r, s - 1):
        raise RuntimeError(f"vsrd_b200: coarse_weights must be [{r}, {s - 1}], got {tuple(w.shape)}")
    if sorted_uniforms is not None:
        sorted_uniforms = _f32(sorted_uniforms, "sorted_uniforms").reshape(r, s)
    out = torch.empty(r, 2 * s, device=t.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_place_fine(_ptr(t), _ptr(w), _ptr(sorted_uniforms), seed & (2 ** 64 - 1),
                                           _state_ptr(step_state), r, s, _ptr(out), _stream()))
    return out


# ---- field + compositing ----------------------------------------------------------------------

def field_forward(scene: SceneArgs, rays: RayArgs, cull: Optional[bool] = None, backward: bool = True,
                  forward_cull: Optional[bool] = None) -> torch.Tensor:
    """`cull`: instance culling for this pass: True / False, or None = the module default (see set_culling) AND a low
    enough temperature -- the scene's own, or the decision its StepState carries (CULL_MAX_TEMPERATURE); `backward`: the
    pass will be differentiated (attach the backward kernels' tile marks); `forward_cull`: run the forward culling
    pre-pass (None = whenever culling is on)."""
    field = torch.empty(scene.num_instances, rays.num_rays * rays.num_intervals, 4,
                        device=rays.directions.device, dtype=torch.float32)
    if cull is None:
        cull = _culling and (scene.temperature <= CULL_MAX_TEMPERATURE if scene.step_state is None else scene.step_state.cull)
    if cull and scene.mlp_weights is not None and scene.num_instances > 1:
        if backward:
            rays.enable_backward_culling(scene)
        if forward_cull is None or forward_cull:
            rays.cull_forward(scene, field)
    _lib.check(_lib.load().vsrd_field_forward(ctypes.byref(scene.struct), ctypes.byref(rays.struct), _ptr(field), _stream()))
    return field


def composite_forward(scene: SceneArgs, rays: RayArgs, field, std_deviation, cosine_ratio, epsilon=1e-6,
                      targets=None, silhouette_weight=1.0, eikonal_weight=0.0):
    dev = rays.directions.device
    r, m, n = rays.num_rays, rays.num_intervals, scene.num_instances
    labels = torch.empty(r, n, device=dev, dtype=torch.float32)
    grads = torch.empty(r, m, 3, device=dev, dtype=torch.float32)
    weights = torch.empty(r, m, device=dev, dtype=torch.float32)
    targets, loss = _loss(targets, silhouette_weight, eikonal_weight, r, n)
    loss_out = torch.zeros(2, device=dev, dtype=torch.float32) if loss is not None else None
    params = _params(std_deviation, cosine_ratio, epsilon)
    _lib.check(_lib.load().vsrd_composite_forward(
        ctypes.byref(scene.struct), ctypes.byref(rays.struct), ctypes.byref(params), _ptr(field),
        _ptr(labels), _ptr(grads), _ptr(weights), ctypes.byref(loss) if loss is not None else None,
        _ptr(loss_out), _stream()))
    return labels, grads, weights, loss_out


def composite_backward(scene: SceneArgs, rays: RayArgs, field, std_deviation, cosine_ratio, epsilon=1e-6,
                       grad_labels=None, grad_gradients=None, grad_weights=None,
                       targets=None, labels=None, silhouette_weight=1.0, eikonal_weight=0.0) -> torch.Tensor:
    r, m, n = rays.num_rays, rays.num_intervals, scene.num_instances
    if grad_labels is not None:
        grad_labels = _f32(grad_labels, "grad_labels").reshape(r, n)
    if grad_gradients is not None:
        grad_gradients = _f32(grad_gradients, "grad_gradients").reshape(r, m, 3)
    if grad_weights is not None:
        grad_weights = _f32(grad_weights, "grad_weights").reshape(r, m)
    targets, loss = _loss(targets, silhouette_weight, eikonal_weight, r, n)
    if loss is not None:
        labels = _f32(labels, "labels").reshape(r, n)
    adjoint = torch.empty_like(field)
    params = _params(std_deviation, cosine_ratio, epsilon)
    if rays.live_tiles is not None:
        rays.live_tiles.zero_()              # marks of an earlier backward through the same pass
    _lib.check(_lib.load().vsrd_composite_backward(
        ctypes.byref(scene.struct), ctypes.byref(rays.struct), ctypes.byref(params), _ptr(field),
        _ptr(grad_labels), _ptr(grad_gradients), _ptr(grad_weights),
        ctypes.byref(loss) if loss is not None else None, _ptr(labels) if loss is not None else None,
        _ptr(adjoint), _stream()))
    return adjoint


_partials_cache = {}


def field_backward(scene: SceneArgs, rays: RayArgs, adjoint: torch.Tensor, *, _entry: str = "vsrd_field_backward"):
    dev = rays.directions.device
    n = scene.num_instances
    blocks = _lib.load().vsrd_backward_blocks_per_instance(n, rays.num_rays, rays.num_intervals)
    if blocks < 1:
        _lib.check(1)
    partials = torch.empty(n * blocks * GRAD_STRIDE, device=dev, dtype=torch.float32)
    g_loc = torch.empty(n, 3, device=dev, dtype=torch.float32)
    g_rot = torch.empty(n, 3, 3, device=dev, dtype=torch.float32)
    g_dim = torch.empty(n, 3, device=dev, dtype=torch.float32)
    g_w = torch.empty(n, MLP_WEIGHTS, device=dev, dtype=torch.float32) if scene.mlp_weights is not None else None
    _lib.check(getattr(_lib.load(), _entry)(
        ctypes.byref(scene.struct), ctypes.byref(rays.struct), _ptr(adjoint), _ptr(partials),
        _ptr(g_loc), _ptr(g_rot), _ptr(g_dim), _ptr(g_w), _stream()))
    return g_loc, g_rot, g_dim, g_w


def experimental_field_backward_tcgen05(scene: SceneArgs, rays: RayArgs, adjoint: torch.Tensor):
    """The tcgen05 / TMEM field backward (a measured dead end, NOT used by anything in the package;
    tools/compare_backward.py and tests/test_gpu_experimental.py call it)."""
    return field_backward(scene, rays, adjoint, _entry="vsrd_experimental_field_backward_tcgen05")


# ---- device-resident schedule -------------------------------------------------------------------

class StepState:
    """`VsrdStepState` in device memory plus the host-side `VsrdSchedule` that drives it
    (scripts/main.py:420-431, 677).  `advance()` / `set_step()` enqueue a one-thread kernel, so they can
    be captured into the same CUDA graph as the step they parameterise."""

    def __init__(self, *, num_steps: int, warmup_steps: int = 0, temperature=(1.0, 0.1), std_deviation=(1.0, 0.1),
                 eikonal_weight: float = 0.01, seed: int = 0, device="cuda"):
        self.schedule = VsrdSchedule(int(num_steps), int(warmup_steps), float(temperature[0]), float(temperature[1]),
                                     float(std_deviation[0]), float(std_deviation[1]), float(eikonal_weight), 0.0,
                                     int(seed) & (2 ** 64 - 1))
        self.buffer = torch.zeros(ctypes.sizeof(VsrdStepState), dtype=torch.uint8, device=device)
        self.cull = False             # host-side: whether the field kernels cull (the owner decides per phase, CULL_MAX_TEMPERATURE)
        self.set_step(0)

    def temperature_at(self, step: int) -> float:
        """Host copy of the annealed temperature of `step` (scripts/main.py:420-431)."""
        s = self.schedule
        x = (math.cos(math.pi * step / max(s.num_steps, 1)) + 1.0) / 2.0
        return x * (s.max_temperature - s.min_temperature) + s.min_temperature

    @property
    def ptr(self) -> int:
        return self.buffer.data_ptr()

    def set_step(self, step: int) -> None:
        _lib.check(_lib.load().vsrd_step_state_update(self.ptr, ctypes.byref(self.schedule), int(step), _stream()))

    def advance(self) -> None:
        _lib.check(_lib.load().vsrd_step_state_update(self.ptr, ctypes.byref(self.schedule), -1, _stream()))

    def read(self) -> dict:
        """Host copy (synchronises); for tests and logging only."""
        raw = bytes(self.buffer.cpu().numpy().tobytes())
        st = VsrdStepState.from_buffer_copy(raw)
        return dict(temperature=st.temperature, std_deviation=st.std_deviation, cosine_ratio=st.cosine_ratio,
                    eikonal_weight=st.eikonal_weight, seed=st.seed, step=st.step)


# ---- a14 / a15: projection, matching, projection losses -----------------------------------------

class ViewArgs:
    def __init__(self, extrinsics, intrinsics, image_size, target_view: int):
        self.extrinsics = _f32(extrinsics, "extrinsic_matrices").reshape(-1, 4, 4)
        v = self.extrinsics.shape[0]
        self.intrinsics = _f32(intrinsics, "intrinsic_matrices").reshape(v, 3, 3)
        self.num_views, self.target_view = v, int(target_view)
        self.height, self.width = int(image_size[0]), int(image_size[1])
        self.struct = VsrdViews(v, self.target_view, self.height, self.width, _ptr(self.extrinsics), _ptr(self.intrinsics))


def projection_step(views: ViewArgs, world_boxes, gt_boxes_2d=None, visible=None, fixed_gt_indices=None,
                    need_grad: bool = True):
    """world_boxes [N,8,3] -> boxes_2d [V,N,4]; with gt_boxes_2d [V,N,4] also (gt_indices [N] int64,
    losses [2], grad_world_boxes [2,N,8,3] or None)."""
    wb = _f32(world_boxes, "world_boxes").reshape(-1, 8, 3)
    n, v, dev = wb.shape[0], views.num_views, wb.device
    boxes = torch.empty(v, n, 4, device=dev, dtype=torch.float32)
    if gt_boxes_2d is None:
        _lib.check(_lib.load().vsrd_projection_step(ctypes.byref(views.struct), n, _ptr(wb), None, None, None,
                                                    _ptr(boxes), None, None, None, None, _stream()))
        return boxes
    gt = _f32(gt_boxes_2d, "gt_boxes_2d").reshape(v, n, 4)
    if visible is not None:
        if visible.dtype not in (torch.uint8, torch.bool) or not visible.is_cuda:
            raise RuntimeError("vsrd_b200: visible must be a CUDA bool/uint8 tensor")
        visible = visible.to(torch.uint8).contiguous().reshape(v, n)
    if fixed_gt_indices is not None:
        if fixed_gt_indices.dtype != torch.int64 or not fixed_gt_indices.is_cuda:
            raise RuntimeError("vsrd_b200: fixed_gt_indices must be a CUDA int64 tensor")
        fixed_gt_indices = fixed_gt_indices.contiguous().reshape(n)
    gt_indices = torch.empty(n, device=dev, dtype=torch.int64)
    losses = torch.empty(2, device=dev, dtype=torch.float32)
    grad = torch.empty(2, n, 8, 3, device=dev, dtype=torch.float32) if need_grad else None
    scratch = torch.empty(_lib.load().vsrd_projection_scratch_floats(v, n), device=dev, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_projection_step(
        ctypes.byref(views.struct), n, _ptr(wb), _ptr(gt), _ptr(visible), _ptr(fixed_gt_indices),
        _ptr(boxes), _ptr(gt_indices), _ptr(losses), _ptr(grad), _ptr(scratch), _stream()))
    return boxes, gt_indices, losses, grad


# ---- a2: ray selection ---------------------------------------------------------------------------

def project_box_3d(boxes_3d: torch.Tensor, intrinsic_matrix: torch.Tensor, epsilon: float = 1e-6) -> torch.Tensor:
    """Camera-frame boxes [B,8,3] + one intrinsic matrix [3,3] -> 2D boxes [B,4] (min u, min v, max u, max v)."""
    boxes = _f32(boxes_3d, "boxes_3d").reshape(-1, 8, 3)
    k = _f32(intrinsic_matrix, "intrinsic_matrix").reshape(3, 3)
    out = torch.empty(boxes.shape[0], 4, device=boxes.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_project_box_3d(_ptr(boxes), boxes.shape[0], _ptr(k), float(epsilon), _ptr(out), _stream()))
    return out


def project_box_3d_backward(boxes_3d, intrinsic_matrix, grad_boxes_2d, epsilon: float = 1e-6) -> torch.Tensor:
    boxes = _f32(boxes_3d, "boxes_3d").reshape(-1, 8, 3)
    k = _f32(intrinsic_matrix, "intrinsic_matrix").reshape(3, 3)
    grad = _f32(grad_boxes_2d, "grad_boxes_2d").reshape(boxes.shape[0], 4)
    out = torch.empty_like(boxes)
    _lib.check(_lib.load().vsrd_project_box_3d_backward(_ptr(boxes), boxes.shape[0], _ptr(k), float(epsilon), _ptr(grad),
                                                        _ptr(out), _stream()))
    return out


def ray_cdf_build(soft_masks: torch.Tensor) -> torch.Tensor:
    """soft_masks [..., N] (flattened to [P,N]) -> inclusive CDF [P] (float64) of max_n soft_masks."""
    m = _f32(soft_masks, "soft_masks")
    n = m.shape[-1]
    m = m.reshape(-1, n)
    p = m.shape[0]
    cdf = torch.empty(p, device=m.device, dtype=torch.float64)
    scratch = torch.empty(_lib.load().vsrd_ray_cdf_scratch_doubles(p), device=m.device, dtype=torch.float64)
    _lib.check(_lib.load().vsrd_ray_cdf_build(_ptr(m), p, n, _ptr(cdf), _ptr(scratch), _stream()))
    return cdf


def select_rays(cdf: torch.Tensor, num_rays: int, uniforms: Optional[torch.Tensor] = None, seed: int = 0,
                step_state=None, status: Optional[torch.Tensor] = None):
    """Weighted draw of `num_rays` distinct pixels (scripts/main.py:620-627).  Returns (indices [R] int64,
    status [1] int32 on the device: 0 = ok, else the number of rays that could not be drawn)."""
    if not cdf.is_cuda or cdf.dtype != torch.float64:
        raise RuntimeError("vsrd_b200: cdf must be a CUDA float64 tensor (ops.ray_cdf_build)")
    cdf = cdf.contiguous()
    draws = 0
    if uniforms is not None:
        if not uniforms.is_cuda or uniforms.dtype != torch.float64:
            raise RuntimeError("vsrd_b200: uniforms must be a CUDA float64 tensor")
        uniforms = uniforms.contiguous().reshape(-1)
        draws = uniforms.numel()
    out = torch.empty(num_rays, device=cdf.device, dtype=torch.int64)
    if status is None:
        status = torch.zeros(1, device=cdf.device, dtype=torch.int32)
    _lib.check(_lib.load().vsrd_select_rays(_ptr(cdf), cdf.numel(), _ptr(uniforms), draws, seed & (2 ** 64 - 1),
                                            _state_ptr(step_state), num_rays, _ptr(out), _ptr(status), _stream()))
    return out, status


def gather_targets(soft_masks: torch.Tensor, pixel_indices: torch.Tensor, gt_indices: Optional[torch.Tensor] = None):
    m = _f32(soft_masks, "soft_masks")
    n = m.shape[-1]
    m = m.reshape(-1, n)
    if pixel_indices.dtype != torch.int64 or not pixel_indices.is_cuda:
        raise RuntimeError("vsrd_b200: pixel_indices must be a CUDA int64 tensor")
    pix = pixel_indices.contiguous().reshape(-1)
    if gt_indices is not None:
        gt_indices = gt_indices.contiguous().reshape(n)
    out = torch.empty(pix.numel(), n, device=m.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_gather_targets(_ptr(m), _ptr(pix), _ptr(gt_indices), pix.numel(), n, _ptr(out), _stream()))
    return out


def soft_masks(polygons: torch.Tensor, polygon_sizes: torch.Tensor, image_size, temperature: float = 10.0):
    """polygons [V,N,PV,2] (x, y pixel coordinates), polygon_sizes [V,N] int32 -> soft masks [V,H,W,N]."""
    poly = _f32(polygons, "polygons")
    v, n, pv, _ = poly.shape
    if polygon_sizes.dtype != torch.int32 or not polygon_sizes.is_cuda:
        raise RuntimeError("vsrd_b200: polygon_sizes must be a CUDA int32 tensor")
    sizes = polygon_sizes.contiguous().reshape(v, n)
    h, w = int(image_size[0]), int(image_size[1])
    out = torch.empty(v, h, w, n, device=poly.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_soft_masks(_ptr(poly), _ptr(sizes), v, n, pv, h, w, float(temperature), _ptr(out), _stream()))
    return out


# ---- inference / logging renderers -------------------------------------------

def field_points(scene: SceneArgs, points: torch.Tensor) -> torch.Tensor:
    """Per-instance field (d_i, grad d_i) at arbitrary points [P,3] -> [N,P,4]."""
    pts = _f32(points, "points").reshape(-1, 3)
    field = torch.empty(scene.num_instances, pts.shape[0], 4, device=pts.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_field_points(ctypes.byref(scene.struct), _ptr(pts), pts.shape[0], _ptr(field), _stream()))
    return field


def union_points(scene: SceneArgs, field: torch.Tensor, want_weights: bool = False):
    """Soft union (main.py:477-492) of field [N,P,4] -> union [P,4] (d, grad d) and, optionally, the softmin
    weights [P,N] (the soft instance labels)."""
    f = _f32(field, "field")
    n, p = scene.num_instances, f.shape[1]
    if tuple(f.shape) != (n, p, 4):
        raise RuntimeError(f"vsrd_b200: field must be [{n}, P, 4], got {tuple(f.shape)}")
    out = torch.empty(p, 4, device=f.device, dtype=torch.float32)
    weights = torch.empty(p, n, device=f.device, dtype=torch.float32) if want_weights else None
    _lib.check(_lib.load().vsrd_union_points(ctypes.byref(scene.struct), _ptr(f), p, _ptr(out), _ptr(weights), _stream()))
    return out, weights


def sphere_trace_step(union_out, directions, positions, foreground, converged, active, iteration: int,
                      convergence_criteria: float, bounding_radius: Optional[float]) -> None:
    """One iteration of sphere_tracing (renderers.py:45-55), in place on positions / foreground / converged."""
    p = positions.shape[0]
    if directions.shape[0] not in (1, p):
        raise RuntimeError("vsrd_b200: directions must be [P,3] or [1,3]")
    for t, name, dt in ((foreground, "foreground", torch.uint8), (converged, "converged", torch.uint8), (active, "active", torch.int32)):
        if t.dtype != dt or not t.is_cuda or not t.is_contiguous():
            raise RuntimeError(f"vsrd_b200: {name} must be a contiguous CUDA {dt} tensor")
    if iteration >= active.numel():
        raise RuntimeError("vsrd_b200: `active` holds fewer counters than iterations")
    _lib.check(_lib.load().vsrd_sphere_trace_step(
        _ptr(union_out), _ptr(directions), int(directions.shape[0] == p and p > 0), p, float(convergence_criteria),
        float(bounding_radius or 0.0), _ptr(positions), _ptr(foreground), _ptr(converged), _ptr(active), int(iteration), _stream()))


# ---- a3 / a4 / a16: models and optimiser (csrc/vsrd_model.cu) --------------------------------------

def hyper_forward(net: "_lib.VsrdHyperNet", embeddings: torch.Tensor, activations: Optional[torch.Tensor] = None,
                  mlp_weights: Optional[torch.Tensor] = None):
    """HyperDistanceField.forward: embeddings [N,256] -> (mlp_weights [N,out], activations [L-1,N,256])."""
    emb = _f32(embeddings, "embeddings").reshape(-1, _lib.HYPER_WIDTH)
    n, layers = emb.shape[0], net.num_layers
    out_features = net.layers[layers - 1].out_features
    if activations is None:
        activations = torch.empty(layers - 1, n, _lib.HYPER_WIDTH, device=emb.device, dtype=torch.float32)
    if mlp_weights is None:
        mlp_weights = torch.empty(n, out_features, device=emb.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_hyper_forward(ctypes.byref(net), _ptr(emb), n, _ptr(activations), _ptr(mlp_weights), _stream()))
    return mlp_weights, activations


def hyper_backward(net, grads, embeddings: torch.Tensor, activations: torch.Tensor, grad_mlp_weights: torch.Tensor,
                   grad_embeddings: torch.Tensor) -> None:
    """Backward of `hyper_forward`: fills every tensor `grads` points at and grad_embeddings [N,256]."""
    emb = _f32(embeddings, "embeddings").reshape(-1, _lib.HYPER_WIDTH)
    n = emb.shape[0]
    gw = _f32(grad_mlp_weights, "grad_mlp_weights").reshape(n, -1)
    scratch = torch.empty(_lib.load().vsrd_hyper_scratch_floats(n), device=emb.device, dtype=torch.float32)
    _lib.check(_lib.load().vsrd_hyper_backward(ctypes.byref(net), ctypes.byref(grads), _ptr(emb), n, _ptr(activations),
                                               _ptr(gw), _ptr(grad_embeddings), _ptr(scratch), _stream()))


def decode_boxes(ranges, raw_locations, raw_dimensions, raw_orientations, lead=None):
    """BoxParameters3D.forward on raw [N,3], [N,3], [N,2] -> (locations, half_extents, rotations [N,3,3], boxes_3d [N,8,3]).
    `lead`: leading shape of the outputs (default (N,)); they are allocated in that shape, not re-viewed."""
    loc = _f32(raw_locations, "raw_locations").reshape(-1, 3)
    n, dev = loc.shape[0], loc.device
    dim = _f32(raw_dimensions, "raw_dimensions").reshape(n, 3)
    ori = _f32(raw_orientations, "raw_orientations").reshape(n, 2)
    lead = (n,) if lead is None else tuple(lead)
    out = [torch.empty(*lead, *s, device=dev, dtype=torch.float32) for s in ((3,), (3,), (3, 3), (8, 3))]
    _lib.check(_lib.load().vsrd_decode_boxes(ctypes.byref(ranges), _ptr(loc), _ptr(dim), _ptr(ori), n,
                                             *[_ptr(t) for t in out], _stream()))
    return tuple(out)


def decode_boxes_backward(ranges, raw_locations, raw_dimensions, raw_orientations, half_extents, rotations,
                          grad_locations, grad_half_extents, grad_rotations, grad_boxes_3d, iou_weight, l1_weight,
                          grad_raw_locations, grad_raw_dimensions, grad_raw_orientations,
                          render_loss_parts=None, projection_losses=None, losses=None) -> None:
    n = raw_locations.reshape(-1, 3).shape[0]
    _lib.check(_lib.load().vsrd_decode_boxes_backward(
        ctypes.byref(ranges), _ptr(raw_locations), _ptr(raw_dimensions), _ptr(raw_orientations), n,
        _ptr(half_extents), _ptr(rotations), _ptr(grad_locations), _ptr(grad_half_extents), _ptr(grad_rotations),
        _ptr(grad_boxes_3d), float(iou_weight), float(l1_weight), _ptr(grad_raw_locations), _ptr(grad_raw_dimensions),
        _ptr(grad_raw_orientations), _ptr(render_loss_parts), _ptr(projection_losses), _ptr(losses), _stream()))


def adam_step(params, grads, exp_avg, exp_avg_sq, groups, step_state=None, step: int = 0) -> None:
    """torch.optim.Adam + ExponentialLR over a flat arena (see include/vsrd_b200.h VsrdAdamGroups)."""
    for t, name in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != params.numel():
            raise RuntimeError(f"vsrd_b200: {name} must be a contiguous CUDA float32 tensor of the arena's size")
    _lib.check(_lib.load().vsrd_adam_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(),
                                          ctypes.byref(groups), _state_ptr(step_state), int(step), _stream()))
