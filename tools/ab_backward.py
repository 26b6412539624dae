#!/usr/bin/env python
"""Times the shipped field-backward launch (ranges + kernel + row reduction) on one scene with CUDA events and an L2
flush between launches -- the loop used for the A/B of kernel variants; also the ncu target for the
kernel's profile:
    python tools/ab_backward.py [--cfg cfg2] [--rays 1000] [--reps 20]
    ncu --set full -k regex:field_backward_mma -s 3 -c 1 -o out python tools/ab_backward.py --reps 3"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import fullsize_cases as fc  # noqa: E402
from vsrd_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--rays", type=int, default=1000)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
inp = fc.scene_inputs(a.cfg)
s = fc.SCHEDULES["mid"]
r = a.rays
scene = ops.SceneArgs(*[inp[k].to(dev) for k in fc.GRAD_NAMES], s["temperature"], 100.0)
gen = torch.Generator().manual_seed(0)
dist = torch.sort(torch.rand(r, 2 * fc.NUM_SAMPLES, generator=gen) * 60.0, dim=-1).values.to(dev)
rays = ops.RayArgs(inp["origins"][:r].to(dev), inp["directions"][:r].to(dev), dist)
field = ops.field_forward(scene, rays, cull=False)
adj = (torch.randn(field.shape, generator=gen) * 1e-3).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

g = ops.field_backward(scene, rays, adj)
torch.cuda.synchronize()
print("gradient norms:", " ".join(f"{float(t.norm()):.6e}" for t in g), "finite", all(bool(torch.isfinite(t).all()) for t in g))
ts = []
for _ in range(a.reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.field_backward(scene, rays, adj)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"{a.cfg} R={r}: median {ts[len(ts) // 2]:.4f} ms  best {ts[0]:.4f} ms (ranges + kernel + row reduction launches)")
