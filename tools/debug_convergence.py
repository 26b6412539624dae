#!/usr/bin/env python
"""Trajectory of the per-frame optimisation on a synthetic frame: centre / yaw / size error against the
ground truth and the loss terms every `--every` steps, from the GT itself or from a perturbed start.

    python tools/debug_convergence.py --start gt --steps 300 --warmup 300
    python tools/debug_convergence.py --start perturbed --steps 600 --warmup 200
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from vsrd_b200 import synthetic  # noqa: E402
from vsrd_b200.frame import FrameLabeler, synthetic_frame_inputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--start", choices=["gt", "perturbed"], default="perturbed")
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--instances", type=int, default=4)
    ap.add_argument("--views", type=int, default=7)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--seed", type=int, default=6)
    ap.add_argument("--every", type=int, default=50)
    ap.add_argument("--position-noise", type=float, default=0.5)
    ap.add_argument("--yaw-noise", type=float, default=0.15)
    ap.add_argument("--no-projection", action="store_true")
    ap.add_argument("--no-silhouette", action="store_true")
    args = ap.parse_args()

    frame = synthetic.make_frame(num_instances=args.instances, num_views=args.views, seed=args.seed)
    dev = torch.device("cuda", 0)
    if args.start == "gt":
        raw = (synthetic.logit_range(frame.gt_locations, *synthetic.LOCATION_RANGE),
               synthetic.logit_range(frame.gt_half_extents, *synthetic.DIMENSION_RANGE),
               torch.stack([torch.cos(frame.gt_yaws), torch.sin(frame.gt_yaws)], -1))
    else:
        raw = synthetic.perturbed_raw_parameters(frame, seed=args.seed, position_noise=args.position_noise,
                                                 yaw_noise=args.yaw_noise)
    weights = {}
    if args.no_projection:
        weights.update(iou_projection_loss=0.0, l1_projection_loss=0.0)
    if args.no_silhouette:
        weights.update(silhouette_loss=0.0)
    labeler = FrameLabeler(synthetic_frame_inputs(frame, dev), num_steps=args.steps, warmup_steps=args.warmup,
                           num_rays=1000, num_samples=args.samples, seed=1, model_seed=0, loss_weights=weights,
                           initial_parameters=dict(locations=raw[0].to(dev), dimensions=raw[1].to(dev),
                                                   orientations=raw[2].to(dev)))
    gt = synthetic.gt_corners(frame)

    def report(tag):
        b = labeler.boxes()
        loc = b["locations"].cpu()
        err = (loc - frame.gt_locations)
        rot = b["orientations"].cpu()
        yaw = torch.atan2(rot[:, 0, 2], rot[:, 0, 0])
        dyaw = torch.remainder(yaw - frame.gt_yaws + math.pi / 2, math.pi) - math.pi / 2     # box symmetry: mod pi
        ddim = b["dimensions"].cpu() - frame.gt_half_extents
        corner = (b["boxes_3d"].cpu() - gt).norm(dim=-1).mean(dim=-1)
        print(f"{tag:>6}  |dloc| {err.norm(dim=-1).tolist()}  dx {err[:, 0].tolist()} dz {err[:, 2].tolist()}")
        print(f"        dyaw {dyaw.tolist()}  ddim {ddim.abs().max(dim=-1).values.tolist()}  corner {corner.tolist()}")
        print(f"        losses(total, sil, eik, iou, l1) {labeler.losses.tolist()}")

    report("start")
    while labeler.step_index < args.steps:
        labeler.step()
        if labeler.step_index % args.every == 0:
            report(str(labeler.step_index))


if __name__ == "__main__":
    main()
