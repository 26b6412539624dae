#!/usr/bin/env python
"""Times the model-side launches of one optimisation step (csrc/vsrd_model.cu) with CUDA events, L2-warm, each as a
replayed CUDA graph of 20 back-to-back calls: decode, hypernetwork forward / backward, decode backward, Adam.
    python tools/time_models.py [--instances 8]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vsrd
from vsrd_b200.models import ParameterArena

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
n = a.instances
detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n).to(dev)
hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=256,
                                       hyper_out_channels_list=[256] * 4).to(dev)
arena = ParameterArena(detector, hyper, [1e-2, 1e-2, 1e-2, 1e-3, 1e-4], num_steps=3000, warmup_steps=0)
loc, dim, rot, boxes = arena.decode()
gw = torch.randn(n, 1617, device=dev)
g3, g9, gb = torch.randn(n, 3, device=dev), torch.randn(n, 3, 3, device=dev), torch.randn(2, n, 8, 3, device=dev)
parts, proj, losses = torch.zeros(2, device=dev), torch.zeros(2, device=dev), torch.zeros(5, device=dev)
arena.hyper_forward()

cases = {
    "decode": lambda: arena.decode(),
    "hyper_forward": lambda: arena.hyper_forward(),
    "hyper_backward": lambda: arena.hyper_backward(gw),
    "decode_backward": lambda: arena.decode_backward(dim, rot, g3, g3, g9, gb, 0.1, 1.0, parts, proj, losses),
    "adam_step": lambda: arena.adam_step(step=10),
}
out = {}
reps = 20
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(reps):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        out[name + "_us"] = s.elapsed_time(e) * 1e3 / (10 * reps)
print(json.dumps(dict(instances=n, **out)))
