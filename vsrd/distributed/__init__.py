"""`vsrd.distributed` — the part of vsrd/distributed/* that scripts/main.py touches:
process-group start-up, device selection, the start-up barrier context, rank-0 progress bars and the
frame-partitioning data loader.  VSRD is frame-parallel: "gradients are not averaged between processes"
(README.md:128), so the gradient-averaging helpers, the DDP wrapper and the trainer of the reference
(distributed/utils.py:11-34, parallel.py, trainer.py) are never called by main.py and are not provided.
The multi-GPU driver of this repository is vsrd_b200/sequence.py (same partition, one final NCCL gather)."""
import contextlib
import os

import torch

from .. import utils


def init_process_group(backend, port=None):
    """distributed/initialization.py:7-27 bootstraps MASTER_ADDR/PORT/RANK/WORLD_SIZE through mpi4py for the slurm
    launcher.  Here the rendezvous comes from the environment (torchrun, or a launcher that exports the same four
    variables); `port` fills in MASTER_PORT when the launcher did not set one."""
    if port is not None:
        os.environ.setdefault("MASTER_PORT", str(port))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    missing = [key for key in ("RANK", "WORLD_SIZE") if key not in os.environ]
    if missing:
        raise RuntimeError(f"vsrd.distributed.init_process_group: {', '.join(missing)} not set; launch with torchrun "
                           "(`--launcher torchrun`) or export RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT")
    torch.distributed.init_process_group(backend=backend)


def get_device_id(num_devices_per_process=1, device_id_offset=0):
    """distributed/utils.py:36-41: the local rank among the processes of this node, times the devices each owns."""
    local_processes = torch.cuda.device_count() // num_devices_per_process
    if local_processes < 1:
        raise RuntimeError("vsrd.distributed.get_device_id: no CUDA device visible (there is no CPU path)")
    return (torch.distributed.get_rank() % local_processes) * num_devices_per_process + device_id_offset


def get_logger(*args, rank=0, **kwargs):
    logger = utils.get_logger(*args, **kwargs)
    logger.addFilter(lambda _record: torch.distributed.get_rank() == rank)
    return logger


def tqdm(iterable, *args, **kwargs):
    """Progress bar on rank 0 only (distributed/utils.py:59-60)."""
    if torch.distributed.get_rank():
        return iterable
    import tqdm as _tqdm
    return _tqdm.tqdm(iterable, *args, **kwargs)


@contextlib.contextmanager
def barrier():
    """Barrier on entry and on exit (distributed/utils.py:63-69; main.py:53-57 prints rank by rank inside it)."""
    torch.distributed.barrier()
    try:
        yield
    finally:
        torch.distributed.barrier()


class DistributedDataLoader(torch.utils.data.DataLoader):
    """DataLoader whose default sampler strides the dataset across ranks: `DistributedSampler(dataset)` with its
    defaults (shuffle=True, seed=0; main.py never calls set_epoch) — distributed/loader.py:4-9.  This is the whole
    multi-GPU strategy of the reference: each rank optimises its own target frames."""

    def __init__(self, dataset, *args, sampler=None, batch_sampler=None, **kwargs):
        if sampler is None and batch_sampler is None:
            sampler = torch.utils.data.distributed.DistributedSampler(dataset)
        super().__init__(dataset, *args, sampler=sampler, batch_sampler=batch_sampler, **kwargs)
