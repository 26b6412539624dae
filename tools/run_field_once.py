#!/usr/bin/env python
"""Run the field forward / backward kernels a few times at a BASELINE shape (cfg2 / cfg3, fine pass: R = 1000 rays,
199 intervals), for `ncu -k regex:...` captures and quick CUDA-event timings.

    python tools/run_field_once.py [--cfg cfg2] [--sched mid] [--reps 20] [--backward]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tests import fullsize_cases as fc  # noqa: E402  (scene generator only; no oracle call)
from vsrd_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--sched", default="mid")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--backward", action="store_true")
ap.add_argument("--cull", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
inp = fc.scene_inputs(a.cfg)
s = fc.SCHEDULES[a.sched]
r = fc.NUM_RAYS
scene = ops.SceneArgs(*[inp[k].to(dev) for k in fc.GRAD_NAMES], s["temperature"], 100.0)
o, d = inp["origins"][:r].to(dev), inp["directions"][:r].to(dev)
gen = torch.Generator().manual_seed(0)
dist = torch.sort(torch.rand(r, 2 * fc.NUM_SAMPLES, generator=gen) * 60.0, dim=-1).values.to(dev)
rays = ops.RayArgs(o, d, dist)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = {"forward": [], "backward": []}
for it in range(a.reps):
    flush.zero_()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    field = ops.field_forward(scene, rays, cull=bool(a.cull))
    e1.record()
    if a.backward:
        adj = torch.randn_like(field) * 1e-3
        e1.record()
        ops.field_backward(scene, rays, adj)
    e2.record()
    torch.cuda.synchronize()
    if it >= 3:
        times["forward"].append(e0.elapsed_time(e1))
        times["backward"].append(e1.elapsed_time(e2))
med = lambda v: sorted(v)[len(v) // 2] if v else float("nan")
print(f"{a.cfg}/{a.sched} impl={os.environ.get('VSRD_FIELD_IMPL', 'default')}: forward {med(times['forward']):.4f} ms"
      + (f", backward {med(times['backward']):.4f} ms" if a.backward else "") + f", field checksum {float(field.double().sum()):.6f}")
