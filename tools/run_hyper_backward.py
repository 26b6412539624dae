#!/usr/bin/env python
"""Runs the hypernetwork forward + backward launches a few times, eagerly (a target for `ncu -k regex:hyper_layer`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vsrd
from vsrd_b200.models import ParameterArena

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n).to(dev)
hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=256,
                                       hyper_out_channels_list=[256] * 4).to(dev)
arena = ParameterArena(detector, hyper, [1e-2, 1e-2, 1e-2, 1e-3, 1e-4], num_steps=3000, warmup_steps=0)
gw = torch.randn(n, 1617, device=dev)
for _ in range(4):
    arena.hyper_forward()
    arena.hyper_backward(gw)
torch.cuda.synchronize()
