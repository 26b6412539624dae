"""The reference's UNMODIFIED scripts/main.py (SHA-256 checked, read from the checkout / the staged baseline/_ref at test
time) executed on the drop-in `vsrd` package with the CUDA kernels underneath: 60 optimisation steps of one synthetic
frame (20 box-only warm-up steps, then the residual field), six `scalar_intervals` events, one `image_intervals` event
(full-image two-pass render + sphere tracing + the drawing helpers) and two checkpoints.  VERDICT r1 "missing" #1."""
import glob
import json
import os
import re

import pytest
import torch

from tools import run_main, stage_reference

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = os.path.join(ROOT, "configs", "synthetic", "vsrd", "drive_0000_synthetic", "config.json")

SHORT = [
    "datasets.train.kwargs.num_frames=1", "datasets.train.kwargs.num_source_frames=4",
    "datasets.train.kwargs.image_size=[188,704]", "datasets.train.kwargs.intrinsics_scale=0.5",
    "datasets.train.kwargs.num_instances=4", "datasets.train.kwargs.seed=5",
    "optimization.num_steps=60", "optimization.warmup_steps=20",
    "scheduler.kwargs.gamma=\"eval:0.01 ** (1.0 / 60.0)\"",
    "volume_rendering.num_rays=512", "volume_rendering.num_fine_samples=64",
    "logging.scalar_intervals=10", "logging.image_intervals=60", "logging.ckpt_intervals=30",
]


def _scalars(text):
    """[(step, {name: value})] from the '[Training] ... scalars: {...}' records main.py:939-946 logs."""
    out = []
    for match in re.finditer(r"Step: (\d+), Progress: [^\n]*?scalars: (\{.*?\n\})", text, flags=re.S):
        out.append((int(match.group(1)), json.loads(match.group(2))))
    return out


@pytest.mark.skipif(stage_reference.reference_root() is None, reason="reference checkout not available")
def test_unmodified_main_py_runs_on_the_cuda_kernels(tmp_path, monkeypatch):
    from vsrd_b200 import ops
    monkeypatch.setenv("MASTER_PORT", "29548")
    main_path = run_main.find_main()
    assert stage_reference.sha256(main_path) == stage_reference.MAIN_PY_SHA256
    ops.culling_counters("cuda", reset=True)
    try:
        config = run_main.run(CONFIG, SHORT, workdir=str(tmp_path))
    finally:
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
    _, visited = ops.culling_counters("cuda")
    assert visited > 0, "the residual field kernels never ran: the script did not go through vsrd_b200"
    base = os.path.dirname(config)
    ckpts = sorted(glob.glob(os.path.join(base.replace("configs", "ckpts"), "**", "step_*.pt"), recursive=True))
    assert [os.path.basename(c) for c in ckpts] == ["step_29.pt", "step_59.pt"]
    ckpt = torch.load(ckpts[-1], weights_only=False)
    assert ckpt["step"] == 59 and "iou_3d" in ckpt["metrics"]
    log, = glob.glob(os.path.join(base.replace("configs", "logs"), "**", "log.txt"), recursive=True)
    records = _scalars(open(log).read())
    assert [step for step, _ in records] == [9, 19, 29, 39, 49, 59]
    for step, scalars in records:
        assert all(v == v and abs(v) < 1e6 for v in scalars.values()), (step, scalars)
        assert ("losses/eikonal_loss" in scalars) == (step >= 20)
    first, last = records[0][1], records[-1][1]
    print("step 9:", {k: round(v, 5) for k, v in first.items() if k.startswith(("losses", "metrics/iou"))})
    print("step 59:", {k: round(v, 5) for k, v in last.items() if k.startswith(("losses", "metrics/iou"))})
    assert last["losses/l1_projection_loss"] < first["losses/l1_projection_loss"]
    assert last["learning_rates/detector/locations"] == pytest.approx(0.01 * 0.01 ** (60.0 / 60.0), rel=1e-6)
    events = glob.glob(os.path.join(os.path.dirname(log), "events.out.tfevents.*"))
    assert events and os.path.getsize(events[0]) > 100_000      # the image_intervals branch wrote its renders
