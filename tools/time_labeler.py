#!/usr/bin/env python
"""Times FrameLabeler on KITTI-360-shaped synthetic frames (configs[1]): per-phase ms/step for one frame, and
frames/hour with K frames in flight on one GPU (round-robin stepping, one stream + CUDA graphs per frame)."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsrd_b200 import synthetic
from vsrd_b200.frame import FrameLabeler, synthetic_frame_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3000)
ap.add_argument("--warmup-steps", type=int, default=1000)
ap.add_argument("--instances", type=int, default=8)
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--in-flight", type=int, default=1)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--single-steps", action="store_true", help="frames in flight: step() per host iteration instead of advance()")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ADVANCE = not a.single_steps


def make(fid):
    frame = synthetic.make_frame(a.instances, 17, seed=fid)
    inputs = synthetic_frame_inputs(frame, dev)
    raw = synthetic.perturbed_raw_parameters(frame, seed=fid)
    lab = FrameLabeler(inputs, num_steps=a.steps, warmup_steps=a.warmup_steps, use_graph=not a.no_graph, seed=fid,
                       initial_parameters=dict(locations=raw[0].to(dev), dimensions=raw[1].to(dev), orientations=raw[2].to(dev)))
    return frame, lab


t0 = time.perf_counter()
frame, lab = make(0)
torch.cuda.synchronize(); t2 = time.perf_counter()
marks = {}
w = a.warmup_steps
p2 = next((k for k in range(w, a.steps) if lab.phase_of(k) == 2), a.steps)       # first step with instance culling
for k in range(a.steps):
    if k in (10, w, w + 10, p2, p2 + 10):
        torch.cuda.synchronize(); marks[k] = time.perf_counter()
    lab.step()
torch.cuda.synchronize(); t3 = time.perf_counter()
out = lab.boxes()
gt = synthetic.gt_corners(frame)
err = float((out["boxes_3d"].cpu().mean(1) - gt.mean(1)).norm(dim=-1).mean())
from vsrd_b200 import ops as _ops
culled, visited = _ops.culling_counters(dev, reset=True)
res = dict(culled_tile_fraction_first_frame=(culled / visited) if visited else 0.0, first_frame_setup_s=t2 - t0, first_frame_steps_s=t3 - t2,
           warmup_ms_per_step=(marks[w] - marks[10]) / (w - 10) * 1e3,
           main_ms_per_step=(t3 - marks[w + 10]) / (a.steps - w - 10) * 1e3,
           residual_no_culling_ms_per_step=(marks[p2] - marks[w + 10]) / max(p2 - w - 10, 1) * 1e3 if p2 in marks else None,
           residual_culling_ms_per_step=(t3 - marks[p2 + 10]) / max(a.steps - p2 - 10, 1) * 1e3 if p2 + 10 in marks else None,
           first_culling_step=p2,
           centre_error_m=err, losses=lab.losses.tolist(), draw_failures=int(lab.draw_failures))
if a.frames > 1:
    del lab
    t4 = time.perf_counter()
    done, next_id, active = 0, 1, []
    setup_s = 0.0
    while done < a.frames - 1:
        while len(active) < a.in_flight and next_id < a.frames:
            ts = time.perf_counter()
            active.append(make(next_id)[1]); next_id += 1
            setup_s += time.perf_counter() - ts
        for lab in active:
            lab.advance() if ADVANCE else lab.step()
        for lab in [l for l in active if l.step_index >= a.steps]:
            lab.boxes()["boxes_3d"].cpu()
            active.remove(lab); done += 1
    torch.cuda.synchronize(); t5 = time.perf_counter()
    res.update(steady_frames=a.frames - 1, in_flight=a.in_flight, steady_s_per_frame=(t5 - t4) / (a.frames - 1),
               steady_setup_s_per_frame=setup_s / (a.frames - 1), steady_frames_per_hour=3600.0 * (a.frames - 1) / (t5 - t4))
print(json.dumps(res))
