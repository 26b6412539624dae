"""Frame-parallel driver on CPU: the partitioner against torch's DistributedSampler (what the reference's
DistributedDataLoader uses, vsrd/distributed/loader.py:6-9) and the final label gather on a 2-rank gloo world."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch.utils.data.distributed import DistributedSampler

from vsrd_b200 import sequence


@pytest.mark.parametrize("num_frames,world", [(10, 2), (2562, 8), (7, 4), (3, 8), (1, 1), (64, 8)])
@pytest.mark.parametrize("shuffle", [True, False])
def test_partition_matches_distributed_sampler(num_frames, world, shuffle):
    dataset = list(range(num_frames))
    seen = []
    for rank in range(world):
        ref = list(DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=shuffle, seed=0))
        got = sequence.partition_frames(num_frames, rank, world, seed=0, shuffle=shuffle)
        assert got == ref
        seen += sequence.partition_frames(num_frames, rank, world, seed=0, shuffle=shuffle, drop_duplicates=True)
    assert sorted(seen) == dataset            # without the wrap-around repeats: every frame exactly once


def test_partition_edge_cases():
    assert sequence.partition_frames(0, 0, 4) == []
    with pytest.raises(ValueError):
        sequence.partition_frames(4, 4, 4)
    # survey probe: 10 frames on 2 ranks
    assert sequence.partition_frames(10, 0, 2) == [4, 7, 3, 0, 6]
    assert sequence.partition_frames(10, 1, 2) == [1, 5, 9, 8, 2]


def test_label_frames_skips_done_and_validates():
    calls = []

    def label_one(fid):
        calls.append(fid)
        return dict(boxes_3d=torch.full((2, 8, 3), float(fid)))

    done = {3: dict(boxes_3d=torch.zeros(1, 8, 3))}
    out = sequence.label_frames([1, 3, 5], label_one, done=done)
    assert calls == [1, 5] and sorted(out) == [1, 3, 5]
    with pytest.raises(RuntimeError):
        sequence.label_frames([0], lambda f: dict(boxes_3d=torch.zeros(2, 4, 3)))
    assert {k: v.shape for k, v in sequence.gather_labels(out).items()} == {1: (2, 8, 3), 3: (1, 8, 3), 5: (2, 8, 3)}


def _fake_boxes(fid):
    n = 1 + fid % 5                               # ragged instance counts
    g = torch.Generator().manual_seed(fid)
    return torch.rand(n, 8, 3, generator=g) * 50.0


def _worker(rank, world, port, num_frames, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        labelled = []

        def label_one(fid):
            labelled.append(fid)
            return dict(boxes_3d=_fake_boxes(fid))

        merged = sequence.label_sequence(num_frames, label_one, seed=0)
        ret[rank] = (labelled, {k: v.clone() for k, v in merged.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_frames", [7, 1])
def test_two_rank_gloo_label_gather(num_frames):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker, args=(2, port, num_frames, ret), nprocs=2, join=True)
    (l0, m0), (l1, m1) = ret[0], ret[1]
    assert sorted(l0 + l1) == list(range(num_frames))                  # each frame labelled exactly once
    assert l0 == sequence.partition_frames(num_frames, 0, 2, drop_duplicates=True)
    for merged in (m0, m1):                                            # every rank holds every frame's boxes
        assert sorted(merged) == list(range(num_frames))
        for fid, boxes in merged.items():
            assert torch.equal(boxes, _fake_boxes(fid))


def test_balanced_partition_owns_every_frame_once_and_levels_the_load():
    gen = torch.Generator().manual_seed(0)
    counts = torch.poisson(torch.full((64,), 6.0), generator=gen).clamp(1, 24).tolist()   # cfg5 instance counts
    costs = [n + 2.0 for n in counts]
    for world in (1, 2, 4, 8):
        owned = [sequence.partition_balanced(costs, r, world) for r in range(world)]
        assert sorted(f for part in owned for f in part) == list(range(64))
        loads = [sum(costs[f] for f in part) for part in owned]
        sampler = [sum(costs[f] for f in sequence.partition_frames(64, r, world, drop_duplicates=True)) for r in range(world)]
        assert max(loads) - min(loads) <= max(costs)                      # LPT bound
        assert max(loads) <= max(sampler) + 1e-9                          # never worse than the sampler's stride
    with pytest.raises(ValueError):
        sequence.partition_balanced(costs, 2, 2)


def test_in_flight_scheduler_runs_every_frame_to_completion():
    class FakeLabeler:
        live = 0
        peak = 0

        def __init__(self, fid):
            self.fid, self.step_index = fid, 0
            FakeLabeler.live += 1
            FakeLabeler.peak = max(FakeLabeler.peak, FakeLabeler.live)

        def step(self):
            self.step_index += 1

        def boxes(self):
            FakeLabeler.live -= 1
            return dict(boxes_3d=torch.full((1, 8, 3), float(self.fid)))

    done = []
    out = sequence.label_frames_in_flight([5, 2, 9, 4, 7], FakeLabeler, num_steps=3, in_flight=2,
                                          on_done=lambda fid, lab: done.append((fid, lab.step_index)))
    assert sorted(out) == [2, 4, 5, 7, 9] and FakeLabeler.peak == 2 and FakeLabeler.live == 0
    assert done == [(5, 3), (2, 3), (9, 3), (4, 3), (7, 3)]
    assert all(float(out[f]["boxes_3d"][0, 0, 0]) == f for f in out)


def test_in_flight_scheduler_prefers_multi_step_advance():
    """A labeler that offers advance() (FrameLabeler: several steps per host launch) is driven through it; frames whose
    step counts differ still finish exactly at num_steps."""
    class Advancing:
        def __init__(self, fid):
            self.fid, self.step_index, self.calls = fid, 0, 0

        def step(self):
            raise AssertionError("advance() must be preferred")

        def advance(self):
            self.calls += 1
            n = min(8, 20 - self.step_index)
            self.step_index += n
            return n

        def boxes(self):
            return dict(boxes_3d=torch.full((2, 8, 3), float(self.fid)))

    seen = []
    out = sequence.label_frames_in_flight([1, 3, 2], Advancing, num_steps=20, in_flight=2,
                                          on_done=lambda fid, lab: seen.append((fid, lab.step_index, lab.calls)))
    assert sorted(out) == [1, 2, 3] and seen == [(1, 20, 3), (3, 20, 3), (2, 20, 3)]
