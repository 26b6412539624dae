"""Per-frame auto-labeling loop on the B200 kernels: what `scripts/main.py:train()` does for ONE target
frame (main.py:106-865), device-resident.

The reference runs, per optimisation step, ~21 k ATen calls, 136 Python projection calls, a scipy
assignment on the host and a 9 M-way multinomial.  Here one step is

    schedule (1 thread)  ->  BoxParameters3D decode (PyTorch, [N,.] tensors)  ->  projection + matching +
    projection losses (1 launch)  ->  ray draw + target gather + ray generation (3 launches)  ->
    hypernetwork (PyTorch/cuBLAS, after warm-up)  ->  coarse pass, resampling, fine pass with the fused
    silhouette/eikonal loss (6 launches)  ->  autograd backward (compositing + field adjoint kernels, then
    PyTorch through the decode / hypernetwork)  ->  Adam

with no host synchronisation, so the whole step (forward, backward, optimizer) is captured once into a
CUDA graph per phase (warm-up: box only; main: box + residual field) and replayed for the remaining
steps.  Every per-step scalar (annealed temperature / std deviation, cosine ratio, eikonal switch,
sampler seed, learning-rate decay) lives in device memory (`ops.StepState`).

Losses and weights: main.py:653-687, 391-415, 855 and config.json:120-127; optimizer and scheduler:
config.json:177-215 (Adam, five groups, ExponentialLR with gamma = 0.01 ** (1 / num_steps)).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional

import torch

from . import functional as F
from . import ops

LOSS_WEIGHTS = dict(silhouette_loss=1.0, eikonal_loss=0.01, iou_projection_loss=0.1, l1_projection_loss=1.0)
LEARNING_RATES = dict(locations=1e-2, dimensions=1e-2, orientations=1e-2, embeddings=1e-3, hyper_distance_field=1e-4)


@dataclasses.dataclass
class FrameInputs:
    """Device tensors of one target frame, as main.py:204-316 stacks them."""
    soft_masks: torch.Tensor        # [V,H,W,N] float32, instance-minor, target instance order
    boxes_2d: torch.Tensor          # [V,N,4] x1 y1 x2 y2
    visible: torch.Tensor           # [V,N] bool
    extrinsics: torch.Tensor        # [V,4,4] world -> camera
    intrinsics: torch.Tensor        # [V,3,3]
    target_view: int

    @property
    def image_size(self):
        return int(self.soft_masks.shape[1]), int(self.soft_masks.shape[2])

    @property
    def num_instances(self):
        return int(self.soft_masks.shape[-1])


def synthetic_frame_inputs(frame, device, temperature: float = 10.0) -> FrameInputs:
    """Supervision of a `synthetic.SyntheticFrame`, rasterised on the device (soft_masks kernel)."""
    from . import synthetic
    sup = synthetic.frame_supervision(frame)
    masks = ops.soft_masks(sup.polygons.to(device), sup.polygon_sizes.to(device), frame.image_size, temperature)
    return FrameInputs(masks, sup.boxes_2d.to(device), sup.visible.to(device), frame.extrinsics.to(device),
                       frame.intrinsics.to(device), sup.target_view)


class _ProjectionLosses(torch.autograd.Function):
    """a14 + a15 (main.py:339-415) as one launch; the kernel returns d losses / d corners with the forward."""

    @staticmethod
    def forward(ctx, world_boxes, views, gt_boxes_2d, visible):
        _, gt_indices, losses, grad = ops.projection_step(views, world_boxes.detach(), gt_boxes_2d, visible)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(gt_indices)
        return losses, gt_indices

    @staticmethod
    def backward(ctx, grad_losses, _):
        grad, = ctx.saved_tensors
        return (grad * grad_losses.reshape(2, 1, 1, 1)).sum(dim=0), None, None, None


class FrameLabeler:
    """Optimises the boxes (and residual fields) of one target frame.

        labeler = FrameLabeler(inputs, num_steps=3000, warmup_steps=1000)
        result = labeler.run()            # dict(boxes_3d [N,8,3], locations, dimensions, orientations, losses)

    `rays` selects where a step's ray batch comes from:
      "draw"     weighted draw without replacement on the device (main.py:620-627) + target gather (the product path);
      "indices"  `step(pixel_indices)` — the caller injects the pixel indices (parity tests), targets gathered on device;
      "batches"  `step(pixel_indices, targets)` — indices and silhouette targets copied from (pinned) host memory every
                 step (bench.py's end-to-end leg).
    `inject_samples=True` additionally takes the stratified jitter [R,S] and sorted importance uniforms [R,S] of each
    step from the caller instead of the counter-based generator (parity tests against the CPU oracle)."""

    def __init__(self, inputs: FrameInputs, *, num_steps: int = 3000, warmup_steps: int = 1000, num_rays: int = 1000,
                 num_samples: int = 100, distance_range=(0.0, 100.0), seed: int = 0, loss_weights: Optional[Dict] = None,
                 learning_rates: Optional[Dict] = None, temperature=(1.0, 0.1), std_deviation=(1.0, 0.1),
                 use_graph: bool = True, rays: str = "draw", inject_samples: bool = False, initial_parameters=None,
                 model_seed: Optional[int] = None, models: str = "fused"):
        import vsrd
        dev = inputs.soft_masks.device
        if dev.type != "cuda":
            raise RuntimeError("vsrd_b200: FrameLabeler needs CUDA tensors (there is no CPU path)")
        self.device, self.inputs = dev, inputs
        self.num_steps, self.warmup_steps = int(num_steps), int(warmup_steps)
        self.num_rays, self.num_samples = int(num_rays), int(num_samples)
        self.weights = dict(LOSS_WEIGHTS, **(loss_weights or {}))
        if rays not in ("draw", "indices", "batches"):
            raise ValueError(f"rays must be 'draw', 'indices' or 'batches', got {rays!r}")
        if models not in ("fused", "torch"):
            raise ValueError(f"models must be 'fused' or 'torch', got {models!r}")
        self.models = models
        self.use_graph, self.rays, self.inject_samples = bool(use_graph), rays, bool(inject_samples)
        h, w = inputs.image_size
        n = inputs.num_instances
        self.height, self.width, self.num_instances = h, w, n

        # ---- per-frame constants
        inv_e = torch.linalg.inv(inputs.extrinsics.double())
        inv_k = torch.linalg.inv(inputs.intrinsics.double())
        self.inv_projection = (inv_e[:, :3, :3] @ inv_k).float().contiguous()       # rendering/utils.py:8-17
        self.camera_positions = inv_e[:, :3, 3].float().contiguous()
        self.views = ops.ViewArgs(inputs.extrinsics, inputs.intrinsics, (h, w), inputs.target_view)
        self.gt_boxes = inputs.boxes_2d.float().contiguous()
        self.visible = inputs.visible.to(torch.uint8).contiguous()
        self.bins = F.distance_bins(distance_range, num_samples, dev)
        self.scale = float(max(distance_range))
        self.cdf = ops.ray_cdf_build(inputs.soft_masks) if rays == "draw" else None      # once per frame
        self.state = ops.StepState(num_steps=num_steps, warmup_steps=warmup_steps, temperature=temperature,
                                   std_deviation=std_deviation, eikonal_weight=self.weights["eikonal_loss"],
                                   seed=seed, device=dev)
        self._step_view = self.state.buffer[24:32].view(torch.int64)                  # VsrdStepState.step

        # ---- fresh models per frame (main.py:174-199)
        if model_seed is not None:
            torch.manual_seed(model_seed)
        self.detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n).to(dev)
        self.hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16, 16, 16, 16],
                                                    hyper_in_channels=256, hyper_out_channels_list=[256] * 4).to(dev)
        if initial_parameters is not None:
            with torch.no_grad():
                for name, value in initial_parameters.items():
                    getattr(self.detector, name).copy_(value.reshape(getattr(self.detector, name).shape))
        lrs = dict(LEARNING_RATES, **(learning_rates or {}))
        self._base_lrs = [lrs["locations"], lrs["dimensions"], lrs["orientations"], lrs["embeddings"], lrs["hyper_distance_field"]]
        self._log_gamma = math.log(0.01) / float(num_steps)                            # ExponentialLR, config.json:211-215
        if models == "fused":
            # decode, hypernetwork, their backward and Adam as a dozen launches over one flat arena (csrc/vsrd_model.cu)
            from .models import ParameterArena
            self.arena = ParameterArena(self.detector, self.hyper, self._base_lrs, num_steps=num_steps, warmup_steps=warmup_steps)
            self.optimizer = None
            self.side_stream = torch.cuda.Stream(device=dev)
            self.side_stream2 = torch.cuda.Stream(device=dev)
        else:
            # the reference's own formulation: nn.Modules under autograd + torch.optim.Adam (~150 launches per step)
            self.arena = None
            groups = [
                dict(params=[self.detector.locations]), dict(params=[self.detector.dimensions]),
                dict(params=[self.detector.orientations]), dict(params=[self.detector.embeddings]),
                dict(params=list(self.hyper.parameters())),
            ]
            for g, lr in zip(groups, self._base_lrs):
                g["lr"] = torch.tensor(lr, dtype=torch.float32, device=dev)            # tensor lr: updated in-graph
            self.optimizer = torch.optim.Adam(groups, lr=1e-2, capturable=True, fused=True)

        # ---- static I/O of the step
        self.pixel_indices = torch.zeros(self.num_rays, dtype=torch.int64, device=dev)
        self.targets = torch.zeros(self.num_rays, n, dtype=torch.float32, device=dev)
        self.jitter = torch.zeros(self.num_rays, self.num_samples, device=dev) if inject_samples else None
        self.sorted_uniforms = torch.zeros(self.num_rays, self.num_samples, device=dev) if inject_samples else None
        self.draw_failures = torch.zeros(1, dtype=torch.int32, device=dev)
        self.losses = torch.zeros(5, dtype=torch.float32, device=dev)    # total, silhouette, eikonal, iou, l1 (weighted)
        self.step_index = 0
        # all of the frame's work is enqueued on its own stream, after the set-up above
        self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream())
        # Fused path, on-device draws: the ray batch of step k+1 is drawn DURING step k (it only depends on the CDF
        # and on the seed of step k+1), so the ~50 us single-CTA draw never sits on the critical path.  A second
        # device-resident schedule runs one step ahead to supply that seed.
        self.state_ahead = None
        if self.arena is not None and rays == "draw":
            self.state_ahead = ops.StepState(num_steps=num_steps, warmup_steps=warmup_steps, temperature=temperature,
                                             std_deviation=std_deviation, eikonal_weight=self.weights["eikonal_loss"],
                                             seed=seed, device=dev)
            self.next_pixel_indices = torch.zeros(self.num_rays, dtype=torch.int64, device=dev)
            with torch.cuda.stream(self.stream):
                self._draw_ahead(self.state)          # batch of step 0
                self.state_ahead.set_step(1)
        # phases: 0 = warm-up (box only), 1 = residual field, 2 = residual field with instance culling (late schedule)
        self._graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self._eager_done: Dict[int, int] = {0: 0, 1: 0, 2: 0}
        # advance(): STEPS_PER_REPLAY consecutive steps of one phase captured as ONE graph (everything a step needs --
        # schedule, ray draw, learning rates -- is device-resident), so the host issues one launch per 8 steps
        self._multi_graphs: Dict[int, torch.cuda.CUDAGraph] = {}

    # ---- one optimisation step (main.py:328-865), enqueued on the current stream (= self.stream) ----
    def _step_body(self, residual: bool) -> None:
        if self.arena is not None:
            return self._step_body_fused(residual)
        st = self.state
        # learning rates of this step: lr0 * gamma^step (scheduler.step() follows optimizer.step(), main.py:863-865)
        decay = torch.exp(self._step_view.double() * self._log_gamma).float()
        for group, base in zip(self.optimizer.param_groups, self._base_lrs):
            group["lr"].copy_(decay[0] * base)
        self.optimizer.zero_grad(set_to_none=True)

        world = self.detector()
        boxes_3d = world["boxes_3d"][0]
        proj, gt_indices = _ProjectionLosses.apply(boxes_3d, self.views, self.gt_boxes, self.visible)

        if self.rays == "draw":
            pix, status = ops.select_rays(self.cdf, self.num_rays, step_state=st)
            self.draw_failures.add_(status)
            self.pixel_indices.copy_(pix)
        if self.rays != "batches":
            self.targets.copy_(ops.gather_targets(self.inputs.soft_masks, self.pixel_indices, gt_indices))
        origins, directions = ops.gather_rays(self.inv_projection, self.camera_positions, self.pixel_indices,
                                              self.height, self.width)
        mlp_weights = self.hyper(world["embeddings"])[0] if residual else None
        render_loss, _, parts = F.render_step(
            world["locations"][0], world["orientations"][0], world["dimensions"][0], mlp_weights,
            origins, directions, self.targets, bins=self.bins, temperature=1.0, std_deviation=1.0, cosine_ratio=0.0,
            scale=self.scale, silhouette_weight=self.weights["silhouette_loss"],
            eikonal_weight=self.weights["eikonal_loss"] if residual else 0.0, step_state=st,
            jitter=self.jitter, sorted_uniforms=self.sorted_uniforms)
        w_iou, w_l1 = self.weights["iou_projection_loss"], self.weights["l1_projection_loss"]
        loss = render_loss + w_iou * proj[0] + w_l1 * proj[1]
        loss.backward()
        self.optimizer.step()
        with torch.no_grad():
            self.losses.copy_(torch.stack([loss.detach(), parts[0], parts[1], w_iou * proj[0].detach(), w_l1 * proj[1].detach()]))
        st.advance()

    def _step_body_fused(self, residual: bool) -> None:
        """The same step with the models, their backward and the optimiser as hand-written launches over the
        parameter arena: no autograd, ~30 launches in total.  The projection / matching kernel (one CTA, ~50 us
        of latency) and the ray draw run on a side stream next to the hypernetwork forward; captured into the
        step's CUDA graph they become parallel branches."""
        st, arena, main = self.state, self.arena, torch.cuda.current_stream()
        side, side2 = self.side_stream, self.side_stream2
        w = self.weights
        loc, dim, rot, boxes_3d = arena.decode()
        # branch 1 (side): projection + matching, then the target gather (needs the matched ground-truth order);
        # its results are first needed by the fine compositing pass (targets) and the decode backward (gradients)
        if self.rays == "draw":
            self.pixel_indices.copy_(self.next_pixel_indices)         # drawn during the previous step
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _, gt_indices, proj_losses, proj_grad = ops.projection_step(self.views, boxes_3d, self.gt_boxes, self.visible)
            if self.rays != "batches":
                self.targets.copy_(ops.gather_targets(self.inputs.soft_masks, self.pixel_indices, gt_indices))
        origins, directions = ops.gather_rays(self.inv_projection, self.camera_positions, self.pixel_indices,
                                              self.height, self.width)
        mlp_weights = arena.hyper_forward() if residual else None

        eik_w = w["eikonal_loss"] if residual else 0.0
        scene = ops.SceneArgs(loc, rot, dim, mlp_weights, 1.0, self.scale, st)
        coarse = ops.place_coarse(self.bins, self.num_rays, self.jitter, 0, st)
        rays = ops.RayArgs(origins, directions, coarse)
        field = ops.field_forward(scene, rays, backward=False)
        _, _, coarse_w, _ = ops.composite_forward(scene, rays, field, 1.0, 0.0, 1e-6)
        fine = ops.place_fine(coarse, coarse_w, self.sorted_uniforms, 0, st)
        rays = ops.RayArgs(origins, directions, fine)
        field = ops.field_forward(scene, rays)
        main.wait_stream(side)                                        # targets / projection results from here on
        if self.state_ahead is not None:
            # branch 2 (side2): the NEXT step's ray batch, next to this step's backward
            side2.wait_stream(main)
            with torch.cuda.stream(side2):
                self._draw_ahead(self.state_ahead)
                self.state_ahead.advance()
        labels, _, _, parts = ops.composite_forward(scene, rays, field, 1.0, 0.0, 1e-6, targets=self.targets,
                                                    silhouette_weight=w["silhouette_loss"], eikonal_weight=eik_w)
        adjoint = ops.composite_backward(scene, rays, field, 1.0, 0.0, 1e-6, targets=self.targets, labels=labels,
                                         silhouette_weight=w["silhouette_loss"], eikonal_weight=eik_w)
        g_loc, g_rot, g_dim, g_w = ops.field_backward(scene, rays, adjoint)
        if residual:
            arena.hyper_backward(g_w)
        if self.state_ahead is not None:
            main.wait_stream(side2)
        arena.decode_backward(dim, rot, g_loc, g_dim, g_rot, proj_grad, w["iou_projection_loss"], w["l1_projection_loss"],
                              parts, proj_losses, self.losses)
        arena.adam_step(st)
        st.advance()

    def _draw_ahead(self, state) -> None:
        """Draws the ray batch keyed by `state`'s seed into `next_pixel_indices` (current stream)."""
        pix, status = ops.select_rays(self.cdf, self.num_rays, step_state=state)
        self.draw_failures.add_(status)
        self.next_pixel_indices.copy_(pix)

    def seek(self, step: int) -> None:
        """Continue from optimisation step `step` (schedule, learning-rate decay, sampler streams); parameters and
        optimiser moments are left as they are."""
        with torch.cuda.stream(self.stream):
            self.state.set_step(step)
            if self.state_ahead is not None:
                self._draw_ahead(self.state)
                self.state_ahead.set_step(step + 1)
        self.step_index = int(step)

    def phase_of(self, step: int) -> int:
        """0 = warm-up (box only, main.py:582-618), 1 = residual field, 2 = residual field with instance culling: worth
        its launches once the temperature has dropped to ops.CULL_MAX_TEMPERATURE."""
        if step < self.warmup_steps:
            return 0
        late = self.state.temperature_at(step) <= ops.CULL_MAX_TEMPERATURE
        return 2 if late and ops.culling_enabled() and self.inputs.num_instances > 1 else 1

    def _run_phase(self, phase: int) -> None:
        self.state.cull = phase == 2
        self._step_body(phase >= 1)

    def _capture(self, phase: int, count: int = 1) -> None:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=self.stream):
            for _ in range(count):
                self._run_phase(phase)
        (self._graphs if count == 1 else self._multi_graphs)[phase] = graph

    STEPS_PER_REPLAY = 8

    def advance(self) -> int:
        """Enqueues the next optimisation step(s) and returns how many: STEPS_PER_REPLAY at once, replayed as one CUDA
        graph, when the labeler draws its own rays and samples (nothing comes from the host between steps) and those
        steps lie in one phase of the schedule; a single step() otherwise.  The kernels and their order are exactly
        those of step() called that many times -- only the number of host launches changes (the frames/hour of several
        frames in flight is otherwise sensitive to how fast the host thread turns)."""
        k = self.STEPS_PER_REPLAY
        phase = self.phase_of(self.step_index)
        if (self.use_graph and self.rays == "draw" and not self.inject_samples and phase in self._graphs
                and self.step_index + k <= self.num_steps and self.phase_of(self.step_index + k - 1) == phase):
            with torch.cuda.stream(self.stream):
                if phase not in self._multi_graphs:
                    self._capture(phase, k)
                self._multi_graphs[phase].replay()
            self.step_index += k
            return k
        self.step()
        return 1

    def step(self, pixel_indices: Optional[torch.Tensor] = None, targets: Optional[torch.Tensor] = None,
             jitter: Optional[torch.Tensor] = None, sorted_uniforms: Optional[torch.Tensor] = None) -> None:
        """Enqueues optimisation step `self.step_index` on the labeler's own stream (asynchronous: several
        labelers can be stepped round-robin so their frames overlap on one GPU).  Injected inputs (see `rays`,
        `inject_samples`) are copied from the given (pinned host or device) tensors into the step's static
        buffers first."""
        if self.step_index >= self.num_steps:
            raise RuntimeError("vsrd_b200: FrameLabeler.step() called after the last optimisation step")
        if self.rays != "draw" and (pixel_indices is None or (self.rays == "batches" and targets is None)):
            raise RuntimeError(f"vsrd_b200: rays={self.rays!r} needs the ray batch of every step")
        if self.inject_samples and (jitter is None or sorted_uniforms is None):
            raise RuntimeError("vsrd_b200: inject_samples=True needs jitter and sorted_uniforms for every step")
        phase = self.phase_of(self.step_index)
        # Injected DEVICE tensors were produced on the caller's stream and may be freed by the caller as soon as this
        # call returns: the labeler's stream waits for their producer, and the caching allocator is told that this
        # stream still reads them (without both, a host that runs ahead of the GPU feeds a step the wrong batch).
        on_device = [t for t in (pixel_indices, targets, jitter, sorted_uniforms) if t is not None and t.is_cuda]
        if on_device:
            self.stream.wait_stream(torch.cuda.current_stream(self.stream.device))
        with torch.cuda.stream(self.stream):
            if self.rays != "draw":
                self.pixel_indices.copy_(pixel_indices.reshape(-1), non_blocking=True)
                if self.rays == "batches":
                    self.targets.copy_(targets, non_blocking=True)
            if self.inject_samples:
                self.jitter.copy_(jitter.reshape(self.jitter.shape), non_blocking=True)
                self.sorted_uniforms.copy_(sorted_uniforms.reshape(self.sorted_uniforms.shape), non_blocking=True)
            for t in on_device:
                t.record_stream(self.stream)
            if self.use_graph and phase not in self._graphs and self._eager_done[phase] >= 3:
                self._capture(phase)
            if self.use_graph and phase in self._graphs:
                self._graphs[phase].replay()
            else:
                # without graphs, or for the first steps of each phase (allocator / cuBLAS warm-up before capture)
                self._run_phase(phase)
                self._eager_done[phase] += 1
        self.step_index += 1

    def synchronize(self) -> None:
        self.stream.synchronize()

    # ---- results -----------------------------------------------------------------------------------
    @torch.no_grad()
    def boxes(self) -> Dict[str, torch.Tensor]:
        """Decoded boxes after the steps enqueued so far (waits for the labeler's stream)."""
        torch.cuda.current_stream().wait_stream(self.stream)
        world = self.detector()
        return dict(boxes_3d=world["boxes_3d"][0], locations=world["locations"][0],
                    dimensions=world["dimensions"][0], orientations=world["orientations"][0])

    # ---- checkpoints (scripts/main.py:1109-1121, 134-136) ----------------------------------------------
    def checkpoint(self, metrics: Optional[Dict] = None) -> Dict:
        """What the reference saves as `step_{step}.pt` for a frame (scripts/main.py:1109-1121), key for key:
          step       index of the last completed optimisation step
          models     state dicts under the config's model names (`detector` is what
                     tools/kitti_360/make_predictions.py:50-58 loads to produce the pseudo labels)
          optimizer  torch.optim.Adam.state_dict() over the config's five parameter groups (per-parameter `step`,
                     `exp_avg`, `exp_avg_sq`; `lr` = the decayed rate) -- loadable by torch.optim.Adam.load_state_dict
          scheduler  torch.optim.lr_scheduler.ExponentialLR.state_dict() (`last_epoch` = completed steps)
          metrics    the caller's metric dict (main.py:888-924 fills it with IoU metrics when the frame has 3D ground
                     truth, else leaves it empty)
        plus one extra key, `losses` = (total, silhouette, eikonal, iou, l1) of the last step.  Host tensors; waits for
        the labeler's stream."""
        import vsrd
        self.stream.synchronize()
        cpu = lambda sd: {k: v.detach().cpu().clone() for k, v in sd.items()}
        ckpt = dict(step=self.step_index - 1,
                    models=dict(detector=cpu(self.detector.state_dict()),
                                hyper_distance_field=cpu(self.hyper.state_dict()),
                                positional_encoder=cpu(vsrd.models.SinusoidalEncoder(num_frequencies=8).state_dict())),
                    metrics=dict(metrics or {}),
                    losses=self.losses.detach().cpu().clone())
        if self.arena is not None:
            ckpt["optimizer"], ckpt["scheduler"] = self.arena.torch_optimizer_state(self.step_index)
        else:
            ckpt["optimizer"] = self.optimizer.state_dict()
            gamma = math.exp(self._log_gamma)
            lrs = [base * gamma ** self.step_index for base in self._base_lrs]
            ckpt["scheduler"] = dict(gamma=gamma, base_lrs=list(self._base_lrs), last_epoch=self.step_index,
                                     _step_count=self.step_index + 1, _get_lr_called_within_step=False, _last_lr=lrs)
        return ckpt

    def load_checkpoint(self, ckpt: Dict) -> None:
        """Restore parameters, Adam moments and the schedule position from `checkpoint()` or from a checkpoint written by
        the reference's main.py with the same config (same key set).  The schedule position comes from `step`; the
        optimizer's per-group update counts must agree with it (`step + 1 - first step of the group`), which is how
        torch counts them for this schedule -- anything else is refused rather than resumed with a wrong bias correction."""
        opt = ckpt.get("optimizer")
        completed = int(ckpt["step"]) + 1
        with torch.cuda.stream(self.stream), torch.no_grad():
            for module, key in ((self.detector, "detector"), (self.hyper, "hyper_distance_field")):
                state = ckpt["models"].get(key)
                if state is None:
                    continue
                own = module.state_dict()
                for name, value in state.items():
                    own[name].copy_(torch.as_tensor(value).reshape(own[name].shape))     # in place: parameters stay in the arena
            if isinstance(opt, dict) and "state" in opt:
                if self.arena is not None:
                    counts = self.arena.torch_update_counts(opt)
                    for k, count in enumerate(counts):
                        expected = completed - int(self.arena.adam_groups.first_step[k])
                        if (count or 0) != max(expected, 0):
                            raise ValueError(f"vsrd_b200: optimizer group {k} made {count} Adam updates, the schedule says "
                                             f"{max(expected, 0)} after step {completed - 1} (different warmup_steps?)")
                    self.arena.load_torch_optimizer_state(opt)
                else:
                    self.optimizer.load_state_dict(opt)
            elif opt is not None:
                raise ValueError("vsrd_b200: `optimizer` is not a torch.optim.Adam state_dict")
        self.seek(completed)

    def run(self) -> Dict[str, torch.Tensor]:
        if self.rays != "draw" or self.inject_samples:
            raise RuntimeError("vsrd_b200: run() draws its own rays and samples; drive step(...) yourself when injecting them")
        while self.step_index < self.num_steps:
            self.advance()
        out = self.boxes()
        out["losses"] = self.losses.clone()
        failures = int(self.draw_failures)           # the only host sync of the frame
        if failures:
            raise RuntimeError(f"vsrd_b200: {failures} rays could not be drawn (fewer weighted pixels than num_rays; "
                               "torch.multinomial raises in the reference, main.py:620-627)")
        return out
