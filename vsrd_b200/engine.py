"""A streamlined renderer step on raw kernels (no autograd bookkeeping), optionally replayed as a
CUDA graph: gather rays -> coarse placement -> coarse field/compositing -> importance placement ->
fine field/compositing with the fused loss -> compositing adjoint -> field adjoint + reduction.

This is the work scripts/main.py:629-687 + the renderer part of `backward` (main.py:859) perform per
optimisation step, down to the gradients of the decoded parameters (locations, rotations, half
extents, residual-MLP weights).  The autograd-facing API (vsrd_b200.functional, vsrd.rendering) runs
the same kernels; this class exists so throughput can be measured without Python dispatch between
launches and so a per-frame driver can overlap several frames on streams.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from .functional import distance_bins


class SilhouetteStep:
    # gather, place_coarse, cull, field, composite, place_fine, cull, field, composite, composite_bwd, field_bwd, reduce
    KERNELS_PER_STEP = 12

    def __init__(self, *, inv_projection, camera_positions, image_size, num_rays: int, num_samples: int,
                 distance_range=(0.0, 100.0), scale: float = 100.0, epsilon: float = 1e-6,
                 silhouette_weight: float = 1.0, eikonal_weight: float = 0.01, device="cuda"):
        self.device = torch.device(device)
        self.inv_projection = inv_projection.to(self.device, torch.float32).contiguous()
        self.camera_positions = camera_positions.to(self.device, torch.float32).contiguous()
        self.height, self.width = int(image_size[0]), int(image_size[1])
        self.num_rays, self.num_samples = int(num_rays), int(num_samples)
        self.bins = distance_bins(distance_range, num_samples, self.device)
        self.scale, self.epsilon = float(scale), float(epsilon)
        self.silhouette_weight, self.eikonal_weight = float(silhouette_weight), float(eikonal_weight)
        # static inputs (graph-friendly): overwritten in place every step
        self.pixel_indices = torch.zeros(num_rays, dtype=torch.int64, device=self.device)
        self.targets = None
        self.params: Dict[str, Optional[torch.Tensor]] = {}
        self.schedule = dict(temperature=1.0, std_deviation=1.0, cosine_ratio=0.0)
        self.seed = 0
        self.out: Dict[str, torch.Tensor] = {}
        self.timers = None
        self._graph = None

    # ---- inputs --------------------------------------------------------------------------------
    def set_parameters(self, locations, rotations, half_extents, mlp_weights):
        new = dict(locations=locations, rotations=rotations, half_extents=half_extents, mlp_weights=mlp_weights)
        for k, v in new.items():
            if v is None:
                self.params[k] = None
                continue
            v = v.detach().to(self.device, torch.float32)
            if self.params.get(k) is not None and self.params[k].shape == v.shape:
                self.params[k].copy_(v)          # keep addresses stable for graph replay
            else:
                self.params[k] = v.contiguous().clone()

    def set_batch(self, pixel_indices, targets):
        self.pixel_indices.copy_(pixel_indices, non_blocking=True)
        if self.targets is None or self.targets.shape != targets.shape:
            self.targets = torch.empty(targets.shape, dtype=torch.float32, device=self.device)
        self.targets.copy_(targets, non_blocking=True)

    def set_schedule(self, *, temperature, std_deviation, cosine_ratio):
        changed = self.schedule != dict(temperature=temperature, std_deviation=std_deviation, cosine_ratio=cosine_ratio)
        self.schedule = dict(temperature=float(temperature), std_deviation=float(std_deviation),
                             cosine_ratio=float(cosine_ratio))
        if changed:
            self._graph = None   # scalars are baked into the captured launches

    # ---- the step ------------------------------------------------------------------------------
    def _mark(self, name):
        if self.timers is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timers.append((name, ev))

    def run_eager(self, backward: bool = True):
        p, s = self.params, self.schedule
        residual = p["mlp_weights"] is not None
        eik_w = self.eikonal_weight if residual else 0.0
        self._mark("start")
        origins, dirs = ops.gather_rays(self.inv_projection, self.camera_positions, self.pixel_indices,
                                        self.height, self.width)
        self._mark("gather_rays")
        scene = ops.SceneArgs(p["locations"], p["rotations"], p["half_extents"], p["mlp_weights"],
                              s["temperature"], self.scale)
        coarse = ops.place_coarse(self.bins, self.num_rays, None, self.seed)
        self._mark("place_coarse")
        rays_c = ops.RayArgs(origins, dirs, coarse)
        field_c = ops.field_forward(scene, rays_c, backward=False)
        self._mark("field_forward_coarse")
        _, _, coarse_w, _ = ops.composite_forward(scene, rays_c, field_c, s["std_deviation"], s["cosine_ratio"], self.epsilon)
        self._mark("composite_forward_coarse")
        fine = ops.place_fine(coarse, coarse_w, None, self.seed)
        self._mark("place_fine")
        rays_f = ops.RayArgs(origins, dirs, fine)
        field_f = ops.field_forward(scene, rays_f)
        self._mark("field_forward_fine")
        labels, grads, weights, loss_parts = ops.composite_forward(
            scene, rays_f, field_f, s["std_deviation"], s["cosine_ratio"], self.epsilon,
            targets=self.targets, silhouette_weight=self.silhouette_weight, eikonal_weight=eik_w)
        self._mark("composite_forward_fine")
        self.out = dict(labels=labels, loss_parts=loss_parts, fine_distances=fine, coarse_weights=coarse_w)
        self.rays = dict(coarse=rays_c, fine=rays_f)     # culling lists of the last step (RayArgs.live_pairs)
        if backward:
            adjoint = ops.composite_backward(
                scene, rays_f, field_f, s["std_deviation"], s["cosine_ratio"], self.epsilon,
                targets=self.targets, labels=labels, silhouette_weight=self.silhouette_weight, eikonal_weight=eik_w)
            self._mark("composite_backward")
            g_loc, g_rot, g_dim, g_w = ops.field_backward(scene, rays_f, adjoint)
            self._mark("field_backward")
            self.out.update(grad_locations=g_loc, grad_rotations=g_rot, grad_half_extents=g_dim, grad_mlp_weights=g_w)
        return self.out

    def capture(self):
        """Capture one step into a CUDA graph (inputs are the static buffers set by set_*)."""
        self.timers = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.run_eager()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.run_eager()
        self._graph = graph
        return graph

    def run(self):
        if self._graph is None:
            self.capture()
        self._graph.replay()
        return self.out

    # ---- bookkeeping ---------------------------------------------------------------------------
    def ray_samples_per_step(self) -> int:
        """R * ((S-1) + (2S-1)): coarse + fine field evaluations of the union SDF."""
        return self.num_rays * (3 * self.num_samples - 2)
