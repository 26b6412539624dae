#!/usr/bin/env python
"""Where the culled forward pass spends its time: dense forward, culling pre-pass alone, pre-pass + list-walking forward,
each captured as a CUDA graph and replayed (no launch gaps), on the bench scene at several temperatures.

    python tools/time_forward_culling.py [--samples 199] [--rays 1000]
"""
import argparse
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsrd_b200 import _lib, ops, synthetic  # noqa: E402
from oracle import vsrd_oracle as oracle  # noqa: E402  (hypernetwork weights for the synthetic scene only)

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=1000)
ap.add_argument("--intervals", type=int, default=199)
ap.add_argument("--instances", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
frame = synthetic.make_frame(a.instances, 17, seed=0)
gen = torch.Generator().manual_seed(100)
inv_proj, cam = frame.inverse_projections()
pix = frame.draw_pixel_indices(a.rays, gen)
h, w = frame.image_size
o, d = ops.gather_rays(inv_proj.to(dev), cam.to(dev), pix.to(dev), h, w)
with torch.no_grad():
    weights = oracle.HyperNetwork()(torch.rand(a.instances, 256, generator=gen)).to(dev)
loc, rot, half = (t.to(dev) for t in (frame.gt_locations, frame.gt_rotations, frame.gt_half_extents))
# fine-pass-like distances: half stratified, half clustered around the nearest box centre
base = torch.rand(a.rays, a.intervals + 1, generator=gen) * 100.0
dist = torch.sort(base, dim=-1).values.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for frac in (0.34, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0):
    T = (math.cos(math.pi * frac) + 1) / 2 * 0.9 + 0.1
    scene = ops.SceneArgs(loc, rot, half, weights, T, 100.0)
    rays = ops.RayArgs(o, d, dist)
    field = torch.empty(a.instances, a.rays * a.intervals, 4, device=dev)
    lib = _lib.load()
    dense = timed(lambda: ops.field_forward(scene, ops.RayArgs(o, d, dist), cull=False))
    rays.cull_forward(scene, field)                            # allocates the lists, leaves them attached
    def prepass():
        _lib.check(lib.vsrd_cull_samples(ctypes.byref(scene.struct), ctypes.byref(rays.struct), field.data_ptr(),
                                         rays.forward_samples.data_ptr(), torch.cuda.current_stream().cuda_stream))
    pre = timed(prepass)
    def culled():
        prepass()
        _lib.check(lib.vsrd_field_forward(ctypes.byref(scene.struct), ctypes.byref(rays.struct), field.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    both = timed(culled)
    live, total = rays.live_pairs()
    print(f"frac {frac:.2f} T={T:.3f}: dense {dense:.4f} ms | pre-pass {pre:.4f} | pre-pass + listed forward {both:.4f} "
          f"(forward alone ~{both - pre:.4f}) | live pairs {live / total:.3f} -> ideal {dense * live / total:.4f}")
