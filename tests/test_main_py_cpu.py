"""The reference's UNMODIFIED scripts/main.py driven through every piece of host-side glue of the drop-in package on the
CPU: configuration, distributed start-up (gloo), the synthetic dataset + DistributedDataLoader + collate, the
instance-id association (reversed_pad), models under `import_module`, the optimiser / scheduler from the config, the
meters, the logging branch (`scalar_intervals`, `image_intervals` with vsrd.visualization) and the checkpoint saver.

There is no CPU renderer in this repository, so for THIS test only `vsrd.rendering` is pointed at the reference's own
rendering package (imported unmodified from the checkout); the GPU twin (tests/test_gpu_main_py.py) runs the same
script on the CUDA kernels.  Skipped where neither /root/reference nor the staged baseline/_ref exists."""
import glob
import json
import os
import sys

import pytest
import torch

from tools import run_main, stage_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = os.path.join(ROOT, "configs", "synthetic", "vsrd", "drive_0000_synthetic", "config.json")

TINY = [
    "distributed.backend=\"gloo\"",
    "datasets.train.kwargs.num_frames=1", "datasets.train.kwargs.num_source_frames=2",
    "datasets.train.kwargs.image_size=[24,88]", "datasets.train.kwargs.intrinsics_scale=0.0625",
    "datasets.train.kwargs.mean_instances=2.0", "datasets.train.kwargs.seed=3",
    "optimization.num_steps=6", "optimization.warmup_steps=3",
    "scheduler.kwargs.gamma=\"eval:0.01 ** (1.0 / 6.0)\"",
    "volume_rendering.num_rays=48", "volume_rendering.num_fine_samples=12",
    "surface_rendering.num_iterations=8",
    "logging.scalar_intervals=2", "logging.image_intervals=6", "logging.ckpt_intervals=3",
]


@pytest.mark.skipif(stage_reference.reference_root() is None, reason="reference checkout not available")
def test_unmodified_main_py_runs_on_the_drop_in_glue(tmp_path, monkeypatch):
    import vsrd
    from oracle import ref_import
    monkeypatch.setenv("MASTER_PORT", "29547")
    monkeypatch.setenv("VSRD_REFERENCE_ROOT", stage_reference.reference_root())
    monkeypatch.setattr(ref_import, "REFERENCE_ROOT", stage_reference.reference_root())
    with ref_import.reference_modules() as ref:
        reference_rendering = ref.rendering
    monkeypatch.setattr(vsrd, "rendering", reference_rendering)
    monkeypatch.setattr(vsrd.distributed, "get_device_id", lambda *a, **k: "cpu")
    try:
        config = run_main.run(CONFIG, TINY, workdir=str(tmp_path))
    finally:
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
    base = os.path.dirname(config)
    ckpts = sorted(glob.glob(os.path.join(base.replace("configs", "ckpts"), "**", "step_*.pt"), recursive=True))
    assert [os.path.basename(c) for c in ckpts] == ["step_2.pt", "step_5.pt"]
    ckpt = torch.load(ckpts[-1], weights_only=False)
    assert ckpt["step"] == 5 and set(ckpt["models"]) == {"detector", "hyper_distance_field", "positional_encoder"}
    assert {"optimizer", "scheduler", "metrics"} <= set(ckpt)
    assert "iou_3d" in ckpt["metrics"]                       # the synthetic frames carry GT boxes -> box_3d_iou ran
    logs = glob.glob(os.path.join(base.replace("configs", "logs"), "**", "log.txt"), recursive=True)
    assert len(logs) == 1
    text = open(logs[0]).read()
    assert text.count("[Training]") == 6 and "runtimes" in text and "losses/eikonal_loss" in text
    events = glob.glob(os.path.join(os.path.dirname(logs[0]), "events.out.tfevents.*"))
    assert events, "tensorboard scalars / images were not written"

    # ---- checkpoint -> pseudo-label JSON with confidences (tools/make_predictions.py = tools/kitti_360/make_predictions.py)
    import subprocess
    ckpt_root = base.replace("configs", "ckpts")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_predictions.py"), "--config", config,
                           "--ckpt-root", ckpt_root, "--out", str(tmp_path / "predictions")], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert json.loads(proc.stdout.strip().splitlines()[-1])["prediction_files"] == 3          # target + 2 source frames
    record = json.load(open(glob.glob(str(tmp_path / "predictions" / "*" / "+0.json"))[0]))
    assert len(record["boxes_3d"]["car"]) == len(record["confidences"]["car"]) == len(ckpt["models"]["detector"]["locations"][0])
