"""TEST / BASELINE INFRASTRUCTURE ONLY — one optimisation step of scripts/main.py:328-865 executed by the UNMODIFIED
reference modules on the CPU (bench.py `--impl reference` and `cpu_baseline`; never on a product path).

What runs is the reference's own code, imported from the checkout or from the copy staged under baseline/_ref
(oracle/ref_import.py): `BoxParameters3D`, `HyperDistanceField`, `SinusoidalEncoder`,
`rendering.hierarchical_volumetric_rendering`, `rendering.sdfs.*`, `operations.project_box_3d`, and the field closures
compiled verbatim from the script's AST (`residual_distance_field` ... `hierarchical_wrapper`, main.py:433-523).
The lines of main.py that are not importable (they are inline in `train()`) are restated here one for one and cite
their source: projection loop :339-367, matching :374-386, projection losses :391-415, annealing :420-431, field
composition :530-578, renderer call :629-651, silhouette :653-671, eikonal :679-687, total :855, step :859-865.
The ray batch (pixel indices + matched silhouette targets) is an input, exactly as for the native end-to-end leg.
"""
from __future__ import annotations

import functools
import os

import numpy as np
import scipy.optimize
import torch
import torch.nn as nn
import torchvision

from . import ref_import

LINE_INDICES = [[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]]  # main.py:27-31


class _AttrDict(dict):
    __getattr__ = dict.__getitem__


class ReferenceStep:
    """Holds the reference models / optimiser of one frame and runs optimisation steps on given ray batches."""

    def __init__(self, num_instances, extrinsics, intrinsics, image_size, gt_boxes_2d, visible, num_steps=3000,
                 num_samples=100, distance_range=(0.0, 100.0), raw_parameters=None, seed=0):
        if not ref_import.available():
            raise RuntimeError("reference modules not available (no checkout and no staged baseline/_ref)")
        with ref_import.reference_modules() as ref:
            self.ref = ref
            torch.manual_seed(seed)
            detector = ref.box_parameters.BoxParameters3D(batch_size=1, num_instances=num_instances, num_features=256)
            hyper = ref.fields.HyperDistanceField(in_channels=48, out_channels_list=[16, 16, 16, 16],
                                                  hyper_in_channels=256, hyper_out_channels_list=[256, 256, 256, 256])
            encoder = ref.encoders.SinusoidalEncoder(num_frequencies=8)
        if raw_parameters is not None:
            with torch.no_grad():
                detector.locations.copy_(raw_parameters[0][None])
                detector.dimensions.copy_(raw_parameters[1][None])
                detector.orientations.copy_(raw_parameters[2][None])
        self.models = _AttrDict(detector=detector, hyper_distance_field=hyper, positional_encoder=encoder)
        self.config = _AttrDict(volume_rendering=_AttrDict(distance_range=list(distance_range)))
        self.num_instances, self.num_steps, self.num_samples = num_instances, num_steps, num_samples
        self.closures = ref_import.main_closures(dict(torch=torch, nn=nn, config=self.config, models=self.models,
                                                      num_instances=num_instances))
        self.extrinsics, self.intrinsics, self.image_size = extrinsics, intrinsics, tuple(image_size)
        self.gt_boxes_2d, self.visible = gt_boxes_2d, visible            # [V,N,4] x1y1x2y2, [V,N] bool
        self.target_view = extrinsics.shape[0] // 2
        self.optimizer = torch.optim.Adam([                              # config.json:177-205
            dict(params=[detector.locations], lr=1e-2), dict(params=[detector.dimensions], lr=1e-2),
            dict(params=[detector.orientations], lr=1e-2), dict(params=[detector.embeddings], lr=1e-3),
            dict(params=list(hyper.parameters()), lr=1e-4)], lr=1e-2)
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, gamma=0.01 ** (1.0 / num_steps))

    def step(self, step, ray_positions, ray_directions, targets, warmup=False):
        ref, models, closures = self.ref, self.models, self.closures
        sdfs = ref.rendering.sdfs
        self.optimizer.zero_grad()
        world = models.detector()
        # ---- multi-view projection, main.py:339-367 (one python call per view and instance, as in the script)
        world_boxes = nn.functional.pad(world["boxes_3d"], (0, 1), mode="constant", value=1.0)
        boxes_2d = []
        for extrinsic, intrinsic in zip(self.extrinsics, self.intrinsics):
            camera_boxes = torch.einsum("mn,b...n->b...m", extrinsic, world_boxes)
            camera_boxes = camera_boxes[..., :-1] / camera_boxes[..., -1:]
            projected = torch.stack([
                ref.geometric_operations.project_box_3d(box_3d=box, line_indices=LINE_INDICES, intrinsic_matrix=intrinsic)
                for box in camera_boxes[0]], dim=0)
            boxes_2d.append(torchvision.ops.clip_boxes_to_image(projected.flatten(-2, -1), self.image_size))
        # ---- matching on the target view, main.py:374-386
        cost = -torchvision.ops.distance_box_iou(boxes_2d[self.target_view], self.gt_boxes_2d[self.target_view])
        pd_indices, gt_indices = map(torch.from_numpy, scipy.optimize.linear_sum_assignment(cost.detach().numpy()))
        # ---- projection losses over the visible (view, instance) pairs, main.py:391-415
        pd, gt = [], []
        for view in range(len(boxes_2d)):
            keep = self.visible[view][gt_indices]
            pd.append(boxes_2d[view][pd_indices[keep]])
            gt.append(self.gt_boxes_2d[view][gt_indices[keep]])
        pd, gt = torch.cat(pd), torch.cat(gt)
        iou_projection_loss = torchvision.ops.distance_box_iou_loss(pd, gt, reduction="none").mean()
        l1_projection_loss = nn.functional.smooth_l1_loss(pd, gt, reduction="none").mean()
        # ---- annealing, main.py:420-431
        anneal = lambda x, a, b: (np.cos(np.pi * x) + 1.0) / 2.0 * (a - b) + b
        ratio = step / self.num_steps
        temperature, std_deviation = anneal(ratio, 1.0, 0.1), anneal(ratio, 1.0, 0.1)
        # ---- field composition, main.py:525-618
        locations, dimensions, orientations = world["locations"][0], world["dimensions"][0], world["orientations"][0]
        weights = None if warmup else models.hyper_distance_field(world["embeddings"])[0]
        fields = []
        for label in range(self.num_instances):
            inner = sdfs.box(dimensions[label])
            if not warmup:
                inner = closures["residual_composition"](
                    distance_field=inner,
                    residual_distance_field=closures["residual_distance_field"](
                        distance_field=functools.partial(models.hyper_distance_field.distance_field, weights[label])))
            inst = closures["instance_field"](distance_field=inner,
                                              instance_label=dimensions.new_tensor(label, dtype=torch.long))
            fields.append(sdfs.translation(sdfs.rotation(inst, orientations[label]), locations[label]))
        field = closures["soft_union"](distance_fields=fields, temperature=temperature)
        # ---- two-pass renderer, main.py:511-523, 629-651
        labels, gradients = closures["hierarchical_wrapper"](ref.rendering.hierarchical_volumetric_rendering)(
            distance_field=field, ray_positions=ray_positions, ray_directions=ray_directions,
            distance_range=self.config.volume_rendering.distance_range, num_samples=self.num_samples,
            sdf_std_deviation=std_deviation, cosine_ratio=ratio)
        # ---- losses, main.py:653-687, 855 (weights config.json:120-127)
        silhouette = nn.functional.binary_cross_entropy(labels[..., pd_indices].clamp(1.0e-6, 1.0 - 1.0e-6),
                                                        targets[..., gt_indices], reduction="none").mean()
        loss = 1.0 * silhouette + 0.1 * iou_projection_loss + 1.0 * l1_projection_loss
        if not warmup:
            eikonal = nn.functional.mse_loss(torch.norm(gradients, dim=-1), gradients.new_ones(*gradients.shape[:-1]))
            loss = loss + 0.01 * eikonal
        torch.autograd.backward(loss)                                    # main.py:859
        self.optimizer.step()
        self.scheduler.step()
        return float(loss)
