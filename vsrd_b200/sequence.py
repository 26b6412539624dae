"""Frame-parallel auto-labeling of a sequence.

Target frames are independent optimisations (README.md:122-128) and the reference never averages
gradients across processes: `DistributedDataLoader` (vsrd/distributed/loader.py:6-9) hands each rank a
strided slice of a seeded permutation of the frames and every rank writes its own checkpoints
(main.py:1109-1121).  So the data path has NO collective; one process per GPU labels its frames and the
only exchange is a single `all_gather` of the final pseudo-label boxes at the end (NCCL on GPUs, gloo
in the CPU tests).

    partition_frames   the frames of one rank, exactly as torch's DistributedSampler draws them
    partition_balanced the frames of one rank under longest-processing-time-first balancing on a per-frame cost
                       (instance count): the only scaling loss of this workload is load imbalance
    label_frames       runs a per-frame labeler over this rank's frames (skip-if-done like main.py:134-136)
    label_frames_in_flight  the same with several frames in flight on one GPU (round-robin over their CUDA graphs)
    gather_labels      padded all_gather of (frame id, instance count, boxes [N_max,8,3]) -> dict on every rank
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

MAX_INSTANCES = 32


def partition_frames(num_frames: int, rank: int, world_size: int, *, seed: int = 0, epoch: int = 0,
                     shuffle: bool = True, drop_duplicates: bool = False) -> List[int]:
    """Indices `torch.utils.data.distributed.DistributedSampler(dataset, world_size, rank, shuffle, seed)`
    yields for `rank` (what the reference's loader uses, vsrd/distributed/loader.py:6-9): a permutation
    seeded with seed + epoch, padded by wrapping around to a multiple of world_size, strided by rank.
    `drop_duplicates` removes the wrap-around repeats (the reference recomputes them and then skips them
    through its already-optimised check, main.py:134-136)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    if num_frames <= 0:
        return []
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(num_frames, generator=g).tolist()
    else:
        order = list(range(num_frames))
    total = math.ceil(num_frames / world_size) * world_size
    pad = total - len(order)
    if pad:
        reps = math.ceil(pad / len(order))
        order = order + (order * reps)[:pad]
    mine = order[rank:total:world_size]
    if drop_duplicates:
        first_owner = {}
        for pos, f in enumerate(order):
            first_owner.setdefault(f, pos)
        mine = [f for pos, f in zip(range(rank, total, world_size), mine) if first_owner[f] == pos]
    return mine


def partition_balanced(costs: Sequence[float], rank: int, world_size: int) -> List[int]:
    """Longest-processing-time-first assignment: frames sorted by descending cost (ties by id) go one by one to the
    currently least-loaded rank.  Deterministic and identical on every rank (no communication); every frame is owned
    exactly once, so there are no wrap-around duplicates.  The per-step cost of a frame is linear in its instance
    count (field kernels: N MLP evaluations per sample), so `costs = [N_f + c0]` is the natural model."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    loads = [0.0] * world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for fid in sorted(range(len(costs)), key=lambda f: (-float(costs[f]), f)):
        r = min(range(world_size), key=lambda k: (loads[k], k))
        loads[r] += float(costs[fid])
        owned[r].append(fid)
    return owned[rank]


def label_frames_in_flight(frame_ids: Iterable[int], make_labeler: Callable[[int], object], num_steps: int,
                           in_flight: int = 4, on_done: Optional[Callable[[int, object], None]] = None
                           ) -> Dict[int, Dict[str, torch.Tensor]]:
    """Labels this rank's frames with up to `in_flight` of them resident on the GPU at once: their per-step CUDA
    graphs are replayed round-robin from one host thread, so one frame's launch gaps are filled by the others
    (a single frame leaves the GPU idle ~15 % of a step).  `make_labeler(frame_id)` returns an object with
    `.step()`, `.step_index`, `.boxes()` (vsrd_b200.frame.FrameLabeler)."""
    results: Dict[int, Dict[str, torch.Tensor]] = {}
    queue, active = [int(f) for f in frame_ids], []
    while queue or active:
        while queue and len(active) < in_flight:
            fid = queue.pop(0)
            active.append((fid, make_labeler(fid)))
        for _, labeler in active:
            # FrameLabeler.advance() enqueues several steps per host launch where it can; anything with .step() works
            (labeler.advance if hasattr(labeler, "advance") else labeler.step)()
        for fid, labeler in [item for item in active if item[1].step_index >= num_steps]:
            results[fid] = dict(boxes_3d=labeler.boxes()["boxes_3d"])
            if on_done is not None:
                on_done(fid, labeler)
            active.remove((fid, labeler))
    return results


def label_frames(frame_ids: Iterable[int], label_one: Callable[[int], Dict[str, torch.Tensor]],
                 done: Optional[Dict[int, Dict[str, torch.Tensor]]] = None) -> Dict[int, Dict[str, torch.Tensor]]:
    """Labels this rank's frames one after another.  `label_one(frame_id)` returns at least `boxes_3d`
    [N,8,3].  Frames already in `done` are skipped (main.py:134-136: idempotent restarts)."""
    results: Dict[int, Dict[str, torch.Tensor]] = dict(done or {})
    for fid in frame_ids:
        if fid in results:
            continue
        out = label_one(int(fid))
        boxes = out["boxes_3d"]
        if boxes.dim() != 3 or tuple(boxes.shape[1:]) != (8, 3) or boxes.shape[0] > MAX_INSTANCES:
            raise RuntimeError(f"frame {fid}: boxes_3d must be [N<=32, 8, 3], got {tuple(boxes.shape)}")
        results[int(fid)] = out
    return results


def gather_labels(results: Dict[int, Dict[str, torch.Tensor]], device=None,
                  group: Optional[dist.ProcessGroup] = None) -> Dict[int, torch.Tensor]:
    """All ranks end up with {frame id: boxes_3d [N,8,3]} for every labelled frame of the sequence.

    Ranks may hold different numbers of frames and frames different numbers of instances, so the payload
    is padded: counts are exchanged first (one tiny all_gather), then one all_gather of
    [F_max, 2 + 32*8*3] float32 rows (frame id, instance count, corners).  With no process group
    initialised this is the identity (single-GPU runs)."""
    local = {int(k): v["boxes_3d"].detach() for k, v in results.items()}
    if not (dist.is_available() and dist.is_initialized()):
        return {k: v.cpu() for k, v in local.items()}
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    count = torch.tensor([len(local)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    f_max = max(int(c) for c in counts)
    row = 2 + MAX_INSTANCES * 24
    payload = torch.zeros(max(f_max, 1), row, dtype=torch.float32, device=device)
    for i, (fid, boxes) in enumerate(sorted(local.items())):
        n = boxes.shape[0]
        payload[i, 0] = float(fid)                       # exact for frame ids < 2^24
        payload[i, 1] = float(n)
        payload[i, 2:2 + n * 24] = boxes.to(device=device, dtype=torch.float32).reshape(-1)
    gathered = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    merged: Dict[int, torch.Tensor] = {}
    for r, (chunk, c) in enumerate(zip(gathered, counts)):
        chunk = chunk.cpu()
        for i in range(int(c)):
            fid, n = int(chunk[i, 0]), int(chunk[i, 1])
            merged.setdefault(fid, chunk[i, 2:2 + n * 24].reshape(n, 8, 3).clone())   # duplicates: first rank wins
    return merged


def my_frames(num_frames: int, *, costs: Optional[Sequence[float]] = None, seed: int = 0, shuffle: bool = True) -> List[int]:
    """This rank's frames: the reference's DistributedSampler slice, or the balanced partition when per-frame
    `costs` are given."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    if costs is not None:
        if len(costs) != num_frames:
            raise ValueError("one cost per frame")
        return partition_balanced(costs, rank, world)
    return partition_frames(num_frames, rank, world, seed=seed, shuffle=shuffle, drop_duplicates=True)


def label_sequence(num_frames: int, label_one: Optional[Callable[[int], Dict[str, torch.Tensor]]] = None, *,
                   seed: int = 0, shuffle: bool = True, device=None, costs: Optional[Sequence[float]] = None,
                   make_labeler: Optional[Callable[[int], object]] = None, num_steps: Optional[int] = None,
                   in_flight: int = 4) -> Dict[int, torch.Tensor]:
    """Frame-parallel driver: partition -> label -> gather.  Call from every rank (torchrun).  Either `label_one`
    (one frame at a time) or `make_labeler` + `num_steps` (several frames in flight per GPU)."""
    mine = my_frames(num_frames, costs=costs, seed=seed, shuffle=shuffle)
    if make_labeler is not None:
        if num_steps is None:
            raise ValueError("make_labeler needs num_steps")
        results = label_frames_in_flight(mine, make_labeler, num_steps, in_flight)
    elif label_one is not None:
        results = label_frames(mine, label_one)
    else:
        raise ValueError("label_sequence needs label_one or make_labeler")
    return gather_labels(results, device=device)
