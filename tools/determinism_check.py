#!/usr/bin/env python
"""Launches the field forward / backward kernels 40 times on one full-size scene (cfg2, cfg3) and checks that every
repeat is BIT-identical to the first (accumulation order is a function of the inputs only: no float atomics).
    python tools/determinism_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import fullsize_cases as fc
from vsrd_b200 import ops
dev = torch.device("cuda", 0)
for cfg in ("cfg2", "cfg3"):
    inp = fc.scene_inputs(cfg)
    s = fc.SCHEDULES["mid"]
    r = 1000
    scene = ops.SceneArgs(*[inp[k].to(dev) for k in fc.GRAD_NAMES], s["temperature"], 100.0)
    gen = torch.Generator().manual_seed(0)
    dist = torch.sort(torch.rand(r, 2 * fc.NUM_SAMPLES, generator=gen) * 60.0, dim=-1).values.to(dev)
    rays = ops.RayArgs(inp["origins"][:r].to(dev), inp["directions"][:r].to(dev), dist)
    field = ops.field_forward(scene, rays, cull=False)
    adj = (torch.randn(field.shape, generator=gen) * 1e-3).to(dev)
    base = [t.clone() for t in ops.field_backward(scene, rays, adj)]
    f0 = field.clone()
    bad = 0; badf = 0
    for it in range(40):
        g = ops.field_backward(scene, rays, adj)
        torch.cuda.synchronize()
        if not all(torch.equal(a, b) for a, b in zip(base, g)): bad += 1
        f = ops.field_forward(scene, rays, cull=False)
        if not torch.equal(f, f0): badf += 1
    print(cfg, "backward mismatching repeats:", bad, "of 40; forward:", badf)
