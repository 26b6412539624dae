"""Flat parameter arena of one target frame's models (scripts/main.py:174-199).

`BoxParameters3D` and `HyperDistanceField` stay what the caller sees (same `nn.Module`s, same `state_dict` keys as
the reference), but their parameters are re-pointed at slices of ONE contiguous fp32 buffer so that

  * the hypernetwork forward/backward kernels (csrc/vsrd_model.cu) read weights and write gradients in place,
  * Adam + ExponentialLR for all 27 tensors (config.json:177-215) is a single launch over the arena.

Group order = the reference's optimizer groups: locations, dimensions, orientations, embeddings, hypernetwork.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn as nn

from . import _lib, ops


class ParameterArena:
    def __init__(self, detector: nn.Module, hyper: nn.Module, learning_rates: Sequence[float], *, num_steps: int,
                 warmup_steps: int, betas=(0.9, 0.999), eps: float = 1e-8, final_lr_ratio: float = 0.01):
        groups: List[List[nn.Parameter]] = [[detector.locations], [detector.dimensions], [detector.orientations],
                                            [detector.embeddings], list(hyper.parameters())]
        if len(learning_rates) != len(groups):
            raise ValueError("one learning rate per optimizer group (locations, dimensions, orientations, embeddings, hypernetwork)")
        device = detector.locations.device
        if device.type != "cuda":
            raise RuntimeError("vsrd_b200: ParameterArena needs CUDA parameters (there is no CPU path)")
        total = sum(p.numel() for g in groups for p in g)
        self.params = torch.empty(total, device=device, dtype=torch.float32)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self._grad_views: Dict[int, torch.Tensor] = {}
        offset, ends = 0, []
        for group in groups:
            for p in group:
                n = p.numel()
                view = self.params[offset:offset + n].view(p.shape)
                view.copy_(p.data)
                p.data = view                                   # the module now reads/writes the arena
                self._grad_views[id(p)] = self.grads[offset:offset + n].view(p.shape)
                offset += n
            ends.append(offset)
        self.detector, self.hyper = detector, hyper
        self.num_instances = int(detector.locations.shape[-2])
        self._groups = groups
        self._num_steps = int(num_steps)

        g = _lib.VsrdAdamGroups()
        g.num_groups = len(groups)
        for k, (end, lr) in enumerate(zip(ends, learning_rates)):
            g.group_end[k] = end
            g.base_lr[k] = float(lr)
            # Adam skips parameters without a gradient: embeddings / hypernetwork first get one after the warm-up
            g.first_step[k] = int(warmup_steps) if k >= 3 else 0
        g.beta1, g.beta2, g.eps = float(betas[0]), float(betas[1]), float(eps)
        g.log_gamma = math.log(final_lr_ratio) / float(num_steps)     # ExponentialLR gamma = 0.01 ** (1 / num_steps)
        self.adam_groups = g

        self.ranges = _lib.VsrdBoxRanges()
        lo, hi = detector.location_range.detach().cpu().tolist()
        dlo, dhi = detector.dimension_range.detach().cpu().tolist()
        for k in range(3):
            self.ranges.location_min[k], self.ranges.location_max[k] = lo[k], hi[k]
            self.ranges.dimension_min[k], self.ranges.dimension_max[k] = dlo[k], dhi[k]

        self.net, self.net_grads = self._hyper_tables(hyper)
        n = self.num_instances
        self.activations = torch.empty(self.net.num_layers - 1, n, _lib.HYPER_WIDTH, device=device, dtype=torch.float32)
        self.mlp_weights = torch.empty(n, self.net.layers[self.net.num_layers - 1].out_features, device=device, dtype=torch.float32)

    def grad(self, p: nn.Parameter) -> torch.Tensor:
        """The arena slice that receives d loss / d p."""
        return self._grad_views[id(p)]

    def _hyper_tables(self, hyper: nn.Module):
        blocks = list(hyper.hypernetwork)
        if not 2 <= len(blocks) <= _lib.HYPER_MAX_LAYERS:
            raise RuntimeError(f"vsrd_b200: the hypernetwork kernels take 2..{_lib.HYPER_MAX_LAYERS} Linear layers, got {len(blocks)}")
        net, grads = _lib.VsrdHyperNet(), _lib.VsrdHyperNetGrads()
        net.num_layers = grads.num_layers = len(blocks)
        for l, block in enumerate(blocks):
            linear = block[0]
            if not (hasattr(linear, "weight_v") and hasattr(linear, "weight_g")):
                raise RuntimeError("vsrd_b200: the hypernetwork kernels expect weight-normed Linear layers (weight_g / weight_v)")
            norm = block[1] if len(block) > 1 else None
            if norm is not None and not isinstance(norm, nn.LayerNorm):
                raise RuntimeError("vsrd_b200: the hypernetwork kernels expect Linear -> LayerNorm -> GELU blocks")
            if len(block) > 2 and getattr(block[2], "approximate", "none") != "none":
                raise RuntimeError("vsrd_b200: the hypernetwork kernels implement the exact (erf) GELU")
            L, G = net.layers[l], grads.layers[l]
            L.weight_v, L.weight_g, L.bias = linear.weight_v.data_ptr(), linear.weight_g.data_ptr(), linear.bias.data_ptr()
            G.weight_v, G.weight_g, G.bias = (self.grad(linear.weight_v).data_ptr(), self.grad(linear.weight_g).data_ptr(),
                                              self.grad(linear.bias).data_ptr())
            L.in_features, L.out_features = linear.in_features, linear.out_features
            if norm is not None:
                L.ln_weight, L.ln_bias = norm.weight.data_ptr(), norm.bias.data_ptr()
                G.ln_weight, G.ln_bias = self.grad(norm.weight).data_ptr(), self.grad(norm.bias).data_ptr()
            if linear.in_features != _lib.HYPER_WIDTH or (norm is not None and linear.out_features != _lib.HYPER_WIDTH):
                raise RuntimeError("vsrd_b200: the hypernetwork kernels are compiled for 256-wide layers (configs/kitti_360); "
                                   "rebuild csrc/vsrd_model.cu for other widths")
        return net, grads

    # ---- the model side of one optimisation step -----------------------------------------------------
    def decode(self):
        d = self.detector
        return ops.decode_boxes(self.ranges, d.locations.data, d.dimensions.data, d.orientations.data)

    def hyper_forward(self) -> torch.Tensor:
        ops.hyper_forward(self.net, self.detector.embeddings.data, self.activations, self.mlp_weights)
        return self.mlp_weights

    def hyper_backward(self, grad_mlp_weights: torch.Tensor) -> None:
        emb = self.detector.embeddings
        ops.hyper_backward(self.net, self.net_grads, emb.data, self.activations, grad_mlp_weights, self.grad(emb))

    def decode_backward(self, half_extents, rotations, g_loc, g_dim, g_rot, g_boxes, iou_weight, l1_weight,
                        render_loss_parts=None, projection_losses=None, losses=None) -> None:
        d = self.detector
        ops.decode_boxes_backward(self.ranges, d.locations.data, d.dimensions.data, d.orientations.data, half_extents,
                                  rotations, g_loc, g_dim, g_rot, g_boxes, iou_weight, l1_weight,
                                  self.grad(d.locations), self.grad(d.dimensions), self.grad(d.orientations),
                                  render_loss_parts, projection_losses, losses)

    def adam_step(self, step_state=None, step: int = 0) -> None:
        ops.adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.adam_groups, step_state, step)

    # ---- interchange with torch.optim.Adam / ExponentialLR (scripts/main.py:182-190, 1118-1119) -------------------
    def _slices(self):
        """(group index, parameter, arena offset) in torch's parameter order: groups in order, parameters within."""
        offset = 0
        for k, group in enumerate(self._groups):
            for p in group:
                yield k, p, offset
                offset += p.numel()

    def torch_optimizer_state(self, completed_steps: int):
        """`optimizer.state_dict()` and `scheduler.state_dict()` as the reference's torch.optim.Adam over the config's five
        parameter groups and its ExponentialLR would hold them after `completed_steps` optimisation steps.  Built with
        real torch objects on host copies, so the key set is the installed torch's own.  A group that has not received
        a gradient yet (embeddings / hypernetwork before `warmup_steps`) has no `state` entry, as in torch."""
        g = self.adam_groups
        host = [[torch.nn.Parameter(p.detach().cpu().clone()) for p in group] for group in self._groups]
        opt = torch.optim.Adam([dict(params=ps, lr=float(g.base_lr[k])) for k, ps in enumerate(host)],
                               lr=float(g.base_lr[0]), betas=(float(g.beta1), float(g.beta2)), eps=float(g.eps))
        gamma = math.exp(float(g.log_gamma))
        sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=gamma)
        exp_avg, exp_avg_sq = self.exp_avg.cpu(), self.exp_avg_sq.cpu()
        flat = [p for ps in host for p in ps]
        for (k, p, offset), hp in zip(self._slices(), flat):
            updates = completed_steps - int(g.first_step[k])
            if updates > 0:
                n = p.numel()
                opt.state[hp] = dict(step=torch.tensor(float(updates)),
                                     exp_avg=exp_avg[offset:offset + n].view(p.shape).clone(),
                                     exp_avg_sq=exp_avg_sq[offset:offset + n].view(p.shape).clone())
        lrs = [float(g.base_lr[k]) * gamma ** completed_steps for k in range(len(host))]
        for group, lr in zip(opt.param_groups, lrs):
            group["lr"] = lr
        sched.last_epoch = int(completed_steps)
        sched._step_count = int(completed_steps) + 1
        sched._last_lr = lrs
        return opt.state_dict(), sched.state_dict()

    def load_torch_optimizer_state(self, state_dict: Dict) -> None:
        """Adam moments from a torch.optim.Adam `state_dict()` over the same five groups (a checkpoint written by the
        reference's main.py, or by `torch_optimizer_state`).  Parameters without a state entry get zero moments.  The
        bias-correction count is not stored here: the kernel derives it from the schedule step and `warmup_steps`, which
        is what torch's per-parameter `step` equals for this schedule (checked)."""
        groups = state_dict["param_groups"]
        if [len(gr["params"]) for gr in groups] != [len(gr) for gr in self._groups]:
            raise ValueError("vsrd_b200: optimizer state_dict does not have the config's parameter groups "
                             "(locations, dimensions, orientations, embeddings, hypernetwork)")
        state = state_dict["state"]
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        index = 0
        for k, p, offset in self._slices():
            entry = state.get(index, state.get(str(index)))
            index += 1
            if entry is None:
                continue
            n = p.numel()
            self.exp_avg[offset:offset + n].copy_(torch.as_tensor(entry["exp_avg"]).reshape(-1))
            self.exp_avg_sq[offset:offset + n].copy_(torch.as_tensor(entry["exp_avg_sq"]).reshape(-1))

    def torch_update_counts(self, state_dict: Dict):
        """Per-group Adam update counts of a torch state_dict (None for a group without state)."""
        out, index = [], 0
        for gr in state_dict["param_groups"]:
            entry = state_dict["state"].get(gr["params"][0], state_dict["state"].get(str(gr["params"][0])))
            out.append(None if entry is None else int(float(entry["step"])))
        return out
