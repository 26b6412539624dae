#!/usr/bin/env python
"""A/B of the shipped field-backward kernel ("mma": mma.sync, fragments in registers) and the experimental tcgen05 /
TMEM one ("umma", ops.experimental_field_backward_tcgen05) on one scene: relative differences of the four gradients
and CUDA-event timings.   python tools/compare_backward.py [--cfg cfg2] [--rays 1000] [--reps 10]"""
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
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--sparse", type=float, default=0.0, help="fraction of rays whose adjoints are zeroed")
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
if a.sparse > 0:
    keep = (torch.rand(r, generator=gen) >= a.sparse).to(dev)
    adj = (adj.reshape(adj.shape[0], r, -1, 4) * keep[None, :, None, None]).reshape(adj.shape)
out, times = {}, {}
for impl, backward in (("mma", ops.field_backward), ("umma", ops.experimental_field_backward_tcgen05)):
    g = backward(scene, rays, adj)
    torch.cuda.synchronize()
    out[impl] = [t.clone() for t in g]
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        backward(scene, rays, adj)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    times[impl] = sorted(ts)[len(ts) // 2]
for name, x, y in zip(["locations", "rotations", "half_extents", "mlp_weights"], out["mma"], out["umma"]):
    rel = float((x - y).norm() / x.norm().clamp_min(1e-30))
    print(f"{name:12s} |mma| {float(x.norm()):.4e}  rel diff umma vs mma {rel:.3e}  finite {bool(torch.isfinite(y).all())}")
w_m, w_u = out["mma"][3], out["umma"][3]
for lo, hi, label in ((0, 784, "layer0"), (784, 1056, "layer1"), (1056, 1328, "layer2"), (1328, 1600, "layer3"), (1600, 1617, "layer4")):
    print(f"  {label}: rel diff {float((w_m[:, lo:hi] - w_u[:, lo:hi]).norm() / w_m[:, lo:hi].norm().clamp_min(1e-30)):.3e}")
print(f"time mma {times['mma']:.4f} ms, umma {times['umma']:.4f} ms")
for l, base in ((1, 784), (2, 1056), (3, 1328)):
    m = w_m[:, base:base + 272].reshape(-1, 16, 17)
    u = w_u[:, base:base + 272].reshape(-1, 16, 17)
    rel = lambda x, y: float((x - y).norm() / x.norm().clamp_min(1e-30))
    print(f"  layer{l}: weights rel {rel(m[..., :16], u[..., :16]):.3e}  bias rel {rel(m[..., 16], u[..., 16]):.3e}   |w| {float(m[..., :16].norm()):.3e} |b| {float(m[..., 16].norm()):.3e}")
    print("     mma  w[0,0,:4]", [f"{v:.4e}" for v in m[0, 0, :4].tolist()], "b[0,:3]", [f"{v:.4e}" for v in m[0, :3, 16].tolist()])
    print("     umma w[0,0,:4]", [f"{v:.4e}" for v in u[0, 0, :4].tolist()], "b[0,:3]", [f"{v:.4e}" for v in u[0, :3, 16].tolist()])
m0 = w_m[:, :784].reshape(-1, 16, 49); u0 = w_u[:, :784].reshape(-1, 16, 49)
print(f"  layer0: weights rel {rel(m0[..., :48], u0[..., :48]):.3e}  bias rel {rel(m0[..., 48], u0[..., 48]):.3e}")
