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
