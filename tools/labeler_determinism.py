#!/usr/bin/env python
"""Runs the optimisation-parity case of tests/test_gpu_labeler.py (cached oracle inputs, tests/optim_cases.py) several
times and checks that the final boxes are BIT-identical from run to run (the CUDA path is deterministic; a difference
means a race between streams).   python tools/labeler_determinism.py [--case cfg1] [--runs 6]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import optim_cases as oc
from vsrd_b200.frame import FrameLabeler, synthetic_frame_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="cfg1")
ap.add_argument("--runs", type=int, default=6)
a = ap.parse_args()
c = oc.get_case(a.case)
frame, steps, warm, r, s = c["frame"], c["steps"], c["warmup"], c["num_rays"], c["num_samples"]
init = dict(locations=c["raw"][0], dimensions=c["raw"][1], orientations=c["raw"][2])
first, bad = None, 0
for run in range(a.runs):
    inputs = synthetic_frame_inputs(frame, torch.device("cuda", 0))
    inputs.soft_masks = c["soft"].cuda().contiguous()
    labeler = FrameLabeler(inputs, initial_parameters={k: v.cuda() for k, v in init.items()}, model_seed=oc.MODEL_SEED,
                           num_steps=steps, warmup_steps=warm, num_rays=r, num_samples=s, rays="indices",
                           inject_samples=True, use_graph=True)
    for step in range(steps):
        labeler.step(c["pix"][step].cuda(), jitter=c["jitter"][step].cuda(), sorted_uniforms=c["uniforms"][step].cuda())
    got = labeler.boxes()["boxes_3d"].cpu()
    if first is None:
        first = got
    elif not torch.equal(first, got):
        bad += 1
        print(f"run {run}: max difference {float((first - got).abs().max()):.3e} m")
print(f"{a.case}: {a.runs} runs of {steps} steps, {bad} differ from the first")
sys.exit(1 if bad else 0)
