#!/usr/bin/env python
"""CUDA-event timings of the model-side launches of one optimisation step (csrc/vsrd_model.cu): hypernetwork forward
(5 launches), hypernetwork backward (5 + 1), Adam.  Warm L2, as inside the step's CUDA graph.

    python tools/time_models.py [--instances 8] [--reps 200]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vsrd  # noqa: E402
from vsrd_b200.models import ParameterArena  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=8)
ap.add_argument("--reps", type=int, default=200)
a = ap.parse_args()
dev = "cuda:0"
torch.manual_seed(0)
detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=a.instances).to(dev)
hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=256,
                                       hyper_out_channels_list=[256] * 4).to(dev)
arena = ParameterArena(detector, hyper, [1e-2, 1e-2, 1e-2, 1e-3, 1e-4], num_steps=3000, warmup_steps=0)
gw = torch.randn(a.instances, 1617, device=dev)


def timed(fn):
    for _ in range(10):
        fn()
    graph = torch.cuda.CUDAGraph()                      # graph replay: launch gaps as in the labeler's step
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            for _ in range(10):
                fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps // 10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 100.0)          # us per call
    return sorted(ts)[len(ts) // 2]


print(f"N={a.instances}: hyper_forward {timed(arena.hyper_forward):.1f} us, "
      f"hyper_backward {timed(lambda: arena.hyper_backward(gw)):.1f} us, adam {timed(lambda: arena.adam_step(step=5)):.1f} us")
