"""CPU tests of the pure-Python glue scripts/main.py needs around the renderer:
`vsrd.configuration.Configurator` (configurator.py:116-164) and `vsrd.distributed` (loader.py:4-9,
utils.py:36-69) on a 2-rank gloo world."""
import json
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vsrd

REFERENCE = "/root/reference"


def _write(path, obj):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(obj, f)


@pytest.fixture
def config_tree(tmp_path):
    root = tmp_path / "configs"
    _write(str(root / "config.json"), dict(optimization=dict(num_steps=3000), models=dict(detector=dict(num_features=256))))
    _write(str(root / "kitti" / "config.json"), dict(optimization=dict(warmup_steps=1000), datasets=dict(image_size=[376, 1408])))
    _write(str(root / "kitti" / "drive_0" / "config.json"), dict(datasets=dict(filenames=["a.txt"]), optimization=dict(num_steps=3000)))
    return str(root / "kitti" / "drive_0" / "config.json")


def test_configurator_merges_parent_configs(config_tree):
    config = vsrd.configuration.Configurator.load(config_tree)
    assert config == dict(
        optimization=dict(num_steps=3000, warmup_steps=1000),
        models=dict(detector=dict(num_features=256)),
        datasets=dict(image_size=[376, 1408], filenames=["a.txt"]),
    )
    # main.py:38-40 wraps it in the attribute dict and reads nested keys
    cfg = vsrd.utils.Dict.apply(config)
    assert cfg.optimization.warmup_steps == 1000 and cfg.datasets.filenames == ["a.txt"]


def test_configurator_rejects_conflicting_leaves(tmp_path):
    _write(str(tmp_path / "config.json"), dict(a=1))
    _write(str(tmp_path / "x" / "config.json"), dict(a=2))
    with pytest.raises(AssertionError):
        vsrd.configuration.Configurator.load(str(tmp_path / "x" / "config.json"))
    assert vsrd.configuration.Configurator.merge(dict(a=dict(b=1)), dict(a=dict(c=2)), dict(d=3)) == dict(a=dict(b=1, c=2), d=3)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not mounted")
def test_configurator_equals_reference(config_tree):
    import importlib
    from oracle import ref_import
    shipped = os.path.join(REFERENCE, "configs/kitti_360/vsrd/2013_05_28_drive_0000_sync/config.json")
    ours = [vsrd.configuration.Configurator.load(f) for f in (config_tree, shipped)]
    with ref_import.reference_modules():
        ref = importlib.import_module("vsrd.configuration")
        theirs = [ref.Configurator.load(f) for f in (config_tree, shipped)]
    assert ours == theirs
    assert ours[1]["volume_rendering"]["num_rays"] == 1000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1")
    os.environ.pop("MASTER_PORT", None)
    vsrd.distributed.init_process_group("gloo", port=port)
    try:
        assert dist.get_rank() == rank and dist.get_world_size() == world
        with vsrd.distributed.barrier():
            pass
        loader = vsrd.distributed.DistributedDataLoader(list(range(10)), batch_size=1, collate_fn=lambda b: b[0])
        frames = [int(f) for f in loader]
        it = vsrd.distributed.tqdm(range(3), disable=True)
        assert (rank == 0) == (not isinstance(it, range))          # progress bar on rank 0 only
        torch.save(frames, os.path.join(out_dir, f"{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_loader_partitions_frames(tmp_path):
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="fork")
    frames = [torch.load(os.path.join(str(tmp_path), f"{r}.pt")) for r in range(world)]
    # the survey's probe of the reference's loader: DistributedSampler(shuffle=True, seed=0)
    assert frames[0] == [4, 7, 3, 0, 6] and frames[1] == [1, 5, 9, 8, 2]
    from vsrd_b200 import sequence
    assert frames[0] == sequence.partition_frames(10, 0, 2) and frames[1] == sequence.partition_frames(10, 1, 2)


def test_init_process_group_needs_a_launcher(monkeypatch):
    for key in ("RANK", "WORLD_SIZE"):
        monkeypatch.delenv(key, raising=False)
    with pytest.raises(RuntimeError, match="RANK"):
        vsrd.distributed.init_process_group("gloo", port=12345)


def test_get_device_id_requires_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        vsrd.distributed.get_device_id()
