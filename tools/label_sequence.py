#!/usr/bin/env python
"""Frame-parallel auto-labeling of a synthetic sequence (BASELINE.json configs[4]).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/label_sequence.py --frames F

Every rank labels its DistributedSampler slice of the F frames (seed = frame id, N ~ Poisson(6) clipped to
[1, 24]) with FrameLabeler, `--in-flight` frames at a time on its GPU, and the final boxes are exchanged with
one NCCL all_gather.  Rank 0 prints one JSON line (frames/hour over the whole job, max-over-ranks time)."""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsrd_b200 import sequence, synthetic
from vsrd_b200.frame import FrameLabeler, synthetic_frame_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--steps", type=int, default=3000)
ap.add_argument("--in-flight", type=int, default=2)
ap.add_argument("--fixed-instances", type=int, default=0, help="0: N ~ Poisson(6) clipped to [1,24] per frame")
ap.add_argument("--ckpt-dir", default=None, help="write <dir>/frame_<id>/step_<last>.pt per frame (the reference's per-frame "
                "checkpoint, main.py:1109-1121) and skip frames whose final checkpoint exists (main.py:134-136)")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def instances_of(fid):
    if a.fixed_instances:
        return a.fixed_instances
    g = torch.Generator().manual_seed(fid)
    return int(torch.poisson(torch.tensor(6.0), generator=g).clamp(1, 24))


def make(fid):
    frame = synthetic.make_frame(instances_of(fid), 17, seed=fid)
    raw = synthetic.perturbed_raw_parameters(frame, seed=fid)
    return FrameLabeler(synthetic_frame_inputs(frame, dev), num_steps=a.steps, warmup_steps=a.steps // 3, seed=fid,
                        initial_parameters=dict(locations=raw[0].to(dev), dimensions=raw[1].to(dev), orientations=raw[2].to(dev)))


def ckpt_path(fid):
    return os.path.join(a.ckpt_dir, f"frame_{fid:06d}", f"step_{a.steps - 1}.pt")


mine = sequence.partition_frames(a.frames, rank, world, seed=0, drop_duplicates=True)
resumed = {}
if a.ckpt_dir:
    import vsrd
    for fid in list(mine):
        if os.path.exists(ckpt_path(fid)):                      # already labelled by an earlier run: reuse its boxes
            state = torch.load(ckpt_path(fid), map_location="cpu")["models"]["detector"]
            det = vsrd.models.BoxParameters3D(*state["embeddings"].shape)
            det.load_state_dict(state)
            with torch.no_grad():
                resumed[fid] = dict(boxes_3d=det()["boxes_3d"][0].to(dev))
            mine.remove(fid)
make(mine[0] if mine else 0).step()          # process warm-up (context, cuBLAS, first-launch module loads) outside the clock
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
results, queue, active = {}, list(mine), []
num_resumed = len(resumed)
while queue or active:
    while queue and len(active) < a.in_flight:
        fid = queue.pop(0)
        active.append((fid, make(fid)))
    for _, lab in active:
        lab.advance()
    for fid, lab in [x for x in active if x[1].step_index >= a.steps]:
        out = lab.boxes()
        if int(lab.draw_failures):
            raise RuntimeError(f"frame {fid}: ray draw failed")
        results[fid] = dict(boxes_3d=out["boxes_3d"])
        if a.ckpt_dir:
            os.makedirs(os.path.dirname(ckpt_path(fid)), exist_ok=True)
            torch.save(lab.checkpoint(), ckpt_path(fid))
        active.remove((fid, lab))
results.update(resumed)
merged = sequence.gather_labels(results, device=dev)
torch.cuda.synchronize()
t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = sorted(merged) == list(range(a.frames)) and all(bool(torch.isfinite(b).all()) for b in merged.values()) \
    and all(merged[f].shape[0] == instances_of(f) for f in merged)
if rank == 0:
    print(json.dumps(dict(metric="target_frames_labeled_per_hour", value=a.frames * 3600.0 / float(t[0]), unit="frames/hour",
                          n_gpus=world, frames=a.frames, steps_per_frame=a.steps, in_flight_per_gpu=a.in_flight,
                          seconds=float(t[0]), gathered_frames=len(merged), resumed_frames_rank0=num_resumed, all_frames_gathered_and_finite=ok,
                          instances=[instances_of(f) for f in range(a.frames)])), flush=True)
if world > 1:
    dist.destroy_process_group()
if not ok:
    sys.exit(1)
