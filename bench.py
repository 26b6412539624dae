#!/usr/bin/env python
"""Throughput of the VSRD silhouette-renderer hot path (BASELINE.json metric: ray-samples/s fwd+bwd).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one optimisation step's renderer work on one synthetic KITTI-360-shaped target frame
(configs[1]: 376x1408, 17 views, 8 instances, 1000 rays, 100+100 samples, residual field on):
ray gather -> coarse pass -> importance resampling -> fine pass + silhouette/eikonal loss ->
backward to the gradients of (locations, rotations, half extents, residual-MLP weights).
Unit of work: ray-samples = R * ((S-1) + (2S-1)) per step.

  value       device-resident inputs, raw kernel sequence replayed as a CUDA graph, CUDA-event timed,
              L2 flushed between steps, max over ranks.
  e2e         same metric through the public API a labeling job calls (vsrd_b200.frame.FrameLabeler.step: the
              whole scripts/main.py optimisation step -- decode, projection/matching losses, hypernetwork,
              two-pass render, loss, backward, Adam -- replayed as one CUDA graph), the step's ray batch coming
              from pinned host memory and the losses read back to the host every step.
  eager_api   (informational) the drop-in vsrd.* API composed from main.py's closures every step, no graph.
  main_py     (N=1) the reference's UNMODIFIED scripts/main.py run on the drop-in package (tools/run_main.py), ms/step from
              the script's own log timestamps: what a user of the reference gets without changing a line.
  frames      whole frames labelled per hour through vsrd_b200.sequence.label_sequence over a cfg5 frame list
              (N ~ Poisson(6) clipped to [1,24], 4 frames per rank, several in flight per GPU) INCLUDING the final NCCL
              all_gather of the boxes; per-rank min / max seconds for the DistributedSampler and the balanced partition.
  roofline    dominant kernel (field backward, mma.sync 3xTF32) against the FP32-FMA peak; `forward_fine` = the tcgen05 forward.
  cpu_baseline  the reference's own modules (staged under baseline/_ref, oracle/reference_step.py) on the host cores,
              bounded sample; the oracle port when the staged reference is absent.  The same leg checks the device
              leg's labels / loss of one batch against the oracle.

`--impl reference` times that CPU reference alone at the full 1000 rays/step (rank 0 only under torchrun).
Multi-GPU: frame-parallel, no collective on the data path (weak scaling).
"""
from __future__ import annotations

import argparse
import functools
import json
import math
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_MLP = 2 * (16 * 49 + 3 * 16 * 17 + 17)       # 3234 contraction flop per MLP value pass


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rays", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--instances", type=int, default=8)
    ap.add_argument("--views", type=int, default=17)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-sample-rays", type=int, default=0, help="rays per CPU reference step (0 = the full --rays)")
    ap.add_argument("--cpu-baseline-steps", type=int, default=4, help="timed CPU reference steps of the cpu_baseline leg")
    ap.add_argument("--main-py-steps", type=int, default=90, help="steps of the unmodified scripts/main.py leg (N=1; 0 = skip)")
    ap.add_argument("--frames-per-rank", type=int, default=4)
    ap.add_argument("--in-flight", type=int, default=4)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--frames", type=int, default=1, help="0 = skip the frames/hour leg")
    ap.add_argument("--frame-steps", type=int, default=3000)
    ap.add_argument("--schedule-frac", type=float, default=0.5,
                    help="operating point of the device-resident leg on the annealing schedule (0.5 = step 1500 of 3000)")
    ap.add_argument("--compare-torch-models", action="store_true",
                    help="also time the end-to-end step with the models as nn.Modules under autograd (informational)")
    return ap.parse_args()


# Mid-schedule operating point (step 1500 of 3000, residual field on): scripts/main.py:420-431
def schedule_at(frac=0.5):
    anneal = lambda hi, lo: (math.cos(math.pi * frac) + 1.0) / 2.0 * (hi - lo) + lo
    return dict(temperature=anneal(1.0, 0.1), std_deviation=anneal(1.0, 0.1), cosine_ratio=frac)


def workload_name(args):
    return (f"configs[1]: synthetic KITTI-360 frame 376x1408, V={args.views}, N={args.instances}, "
            f"R={args.rays} rays/step, S={args.samples}+{args.samples} samples/ray, residual MLP 48-16x4-1")


def bench_config(args, world, graph=True):
    """One config dict for both arms (the driver compares them)."""
    return {"workload": workload_name(args), "l2": "flushed (256 MiB memset) between timed steps",
            "schedule": schedule_at(args.schedule_frac), "graph": graph,
            "parallelism": f"frame-parallel x{world}, no collective"}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (the reference's own modules; oracle port as fallback) -- also used for cpu_baseline
# ------------------------------------------------------------------------------------------------
def _cpu_ray_batch(frame, num_rays, num_instances, gen, inv_proj, cam):
    h, w = frame.image_size
    pix = frame.draw_pixel_indices(num_rays, gen)
    view, v, u = pix // (h * w), (pix // w) % h, pix % w
    d = torch.einsum("rmn,rn->rm", inv_proj[view], torch.stack([u, v, torch.ones_like(u)], -1).float())
    return cam[view], torch.nn.functional.normalize(d, dim=-1), torch.rand(num_rays, num_instances, generator=gen)


def cpu_reference_run(args, steps, warmup, num_rays):
    """One optimisation step of scripts/main.py on the host cores, timed: detector decode + multi-view projection /
    matching / projection losses + hypernetwork + two-pass renderer + BCE/eikonal + backward + Adam -- the same step
    the native e2e leg runs, at the same mid-schedule operating point, on the same frame shape.

    kind "reference": the reference's UNMODIFIED modules and main.py's own closures (oracle/reference_step.py, imported
    from the checkout or from the copy staged under baseline/_ref).  kind "port": the oracle restatement
    (oracle/vsrd_oracle.py + frame_oracle.py), only when no reference files are present."""
    from oracle import ref_import
    from vsrd_b200 import synthetic

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    frame = synthetic.make_frame(args.instances, args.views, seed=0)
    sup = synthetic.frame_supervision(frame)
    gen = torch.Generator().manual_seed(0)
    inv_proj, cam = frame.inverse_projections()
    raw = synthetic.perturbed_raw_parameters(frame, seed=0)
    mid_step = int(round(3000 * args.schedule_frac))
    kind = "reference" if ref_import.available() else "port"
    if kind == "reference":
        from oracle import reference_step
        runner = reference_step.ReferenceStep(args.instances, frame.extrinsics, frame.intrinsics, frame.image_size,
                                              sup.boxes_2d, sup.visible, num_steps=3000, num_samples=args.samples,
                                              raw_parameters=raw, seed=0)

        def one_step(it, o, d, targets):
            return runner.step(mid_step + it, o, d, targets)
    else:
        from oracle import frame_oracle
        from oracle import vsrd_oracle as oracle
        leaves = [t.clone().requires_grad_(True) for t in raw]
        emb = torch.rand(256, generator=gen).repeat(args.instances, 1).requires_grad_(True)
        torch.manual_seed(0)
        hyper = oracle.HyperNetwork()
        optimizer = torch.optim.Adam([dict(params=[leaves[0]], lr=1e-2), dict(params=[leaves[1]], lr=1e-2),
                                      dict(params=[leaves[2]], lr=1e-2), dict(params=[emb], lr=1e-3),
                                      dict(params=list(hyper.parameters()), lr=1e-4)], lr=1e-2)
        sched = schedule_at(args.schedule_frac)
        h, w = frame.image_size

        def one_step(it, o, d, targets):
            optimizer.zero_grad(set_to_none=True)
            loc, dim, rot = oracle.decode_box_parameters(*leaves)
            _, _, iou, l1 = frame_oracle.projection_step(oracle.box_corners(loc, dim, rot), frame.extrinsics, frame.intrinsics,
                                                         (h, w), sup.boxes_2d, sup.visible, sup.target_view)
            scene = oracle.Scene(loc, rot, dim, hyper(emb), sched["temperature"])
            loss, _ = oracle.render_loss(scene, o, d, targets, num_samples=args.samples, distance_range=[0.0, 100.0],
                                         sdf_std_deviation=sched["std_deviation"], cosine_ratio=sched["cosine_ratio"])
            loss = loss + 0.1 * iou + 1.0 * l1
            loss.backward()
            optimizer.step()
            return float(loss.detach())

    times = []
    for it in range(warmup + steps):
        o, d, targets = _cpu_ray_batch(frame, num_rays, args.instances, gen, inv_proj, cam)
        t0 = time.perf_counter()
        loss = one_step(it, o, d, targets)
        dt = time.perf_counter() - t0
        if not math.isfinite(loss):
            raise RuntimeError("bench.py: non-finite loss from the CPU reference")
        if it >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    units = num_rays * (3 * args.samples - 2)
    what = ("unmodified reference modules + main.py closures (oracle/reference_step.py)" if kind == "reference"
            else "oracle port (no reference files present)")
    return dict(value=units / per_step, ms_per_step=per_step * 1e3, cores=threads, kind=kind,
                sample=f"{num_rays} of {args.rays} rays/step (whole optimisation step, same frame shape and schedule point), "
                       f"{warmup} warm-up + {steps} timed steps; {what}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    num_rays = args.cpu_sample_rays or args.rays
    res = cpu_reference_run(args, max(1, args.steps), max(1, args.warmup), num_rays)
    line = {
        "impl": "reference", "metric": "ray_samples_per_sec_fwd_bwd", "value": res["value"], "unit": "ray-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, args.gpus),
        "note": "the reference's CPU implementation of the step on the host cores (see cpu_baseline.kind / sample)",
        "cpu_baseline": {"value": res["value"], "unit": "ray-samples/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return False

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "note": "NVML unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def build_frame_inputs(args, rank, device, steps_total):
    """Synthetic frame, per-step ray batches and silhouette targets (rendered from the GT boxes)."""
    from vsrd_b200 import functional as F
    from vsrd_b200 import ops, synthetic

    frame = synthetic.make_frame(args.instances, args.views, seed=rank)
    gen = torch.Generator().manual_seed(100 + rank)
    inv_proj, cam = frame.inverse_projections()
    pool = torch.stack([frame.draw_pixel_indices(args.rays, gen) for _ in range(steps_total)])   # [T,R] int64
    h, w = frame.image_size
    # targets: hard-ish silhouettes of the ground-truth boxes through the same renderer
    targets = []
    gt = [t.to(device) for t in (frame.gt_locations, frame.gt_rotations, frame.gt_half_extents)]
    for k in range(steps_total):
        o, d = ops.gather_rays(inv_proj.to(device), cam.to(device), pool[k].to(device), h, w)
        with torch.no_grad():
            lab, *_ = F.two_pass_render(*gt, None, o, d, num_samples=args.samples, temperature=0.1,
                                        std_deviation=0.1, cosine_ratio=1.0, seed=k)
        targets.append(lab.clamp(0.0, 1.0).cpu())
    return frame, inv_proj, cam, pool, torch.stack(targets)


def make_models(args, frame, rank, device):
    import vsrd
    from vsrd_b200 import synthetic

    torch.manual_seed(rank)
    detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=args.instances)
    hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16, 16, 16, 16],
                                           hyper_in_channels=256, hyper_out_channels_list=[256, 256, 256, 256])
    encoder = vsrd.models.SinusoidalEncoder(num_frequencies=8)
    raw_loc, raw_dim, raw_ori = synthetic.perturbed_raw_parameters(frame, seed=rank)
    with torch.no_grad():
        detector.locations.copy_(raw_loc[None])
        detector.dimensions.copy_(raw_dim[None])
        detector.orientations.copy_(raw_ori[None])
    return detector.to(device), hyper.to(device), encoder.to(device)


def main_style_step(vsrd, model_tuple, config, rays_o, rays_d, targets, sched, num_instances, return_outputs=False):
    """The renderer part of one scripts/main.py step, written against the drop-in API exactly as the
    script composes it (closures of main.py:433-523, 530-578, 629-687)."""
    import torch.nn as nn
    detector, hyper, encoder = model_tuple

    class _NS:
        pass
    models = _NS()   # what main.py calls `models` (attribute access)
    models.positional_encoder, models.hyper_distance_field, models.detector = encoder, hyper, detector
    world = detector()
    weights = hyper(world["embeddings"])

    def residual_distance_field(distance_field):
        def wrapper(positions):
            x, y, z = torch.unbind(positions, dim=-1)
            positions = torch.stack([torch.abs(x), y, z], dim=-1) / max(config.volume_rendering.distance_range)
            return torch.sigmoid(distance_field(models.positional_encoder(positions)) - 1.0)
        return wrapper

    def residual_composition(distance_field, residual_distance_field):
        def wrapper(positions):
            return distance_field(positions) + residual_distance_field(positions)
        return wrapper

    def instance_field(distance_field, instance_label):
        def wrapper(positions):
            distances = distance_field(positions)
            labels = nn.functional.one_hot(instance_label, num_instances)
            return distances, labels.expand(*distances.shape[:-1], -1)
        return wrapper

    def soft_union(distance_fields, temperature):
        def wrapper(positions):
            distances, labels = map(torch.stack, zip(*[f(positions) for f in distance_fields]))
            w = nn.functional.softmin(distances / temperature, dim=0)
            return torch.sum(distances * w, dim=0), torch.sum(labels * w, dim=0)
        return wrapper

    field = soft_union(
        distance_fields=[
            vsrd.rendering.sdfs.translation(
                vsrd.rendering.sdfs.rotation(
                    instance_field(
                        distance_field=residual_composition(
                            distance_field=vsrd.rendering.sdfs.box(dimension),
                            residual_distance_field=residual_distance_field(
                                distance_field=functools.partial(hyper.distance_field, w_i)),
                        ),
                        instance_label=dimension.new_tensor(i, dtype=torch.long),     # main.py:543
                    ),
                    orientation),
                location)
            for i, (location, dimension, orientation, w_i) in enumerate(zip(
                world["locations"][0], world["dimensions"][0], world["orientations"][0], weights[0]))
        ],
        temperature=sched["temperature"],
    )
    render = vsrd.rendering.hierarchical_volumetric_rendering
    kwargs = dict(distance_field=field, ray_positions=rays_o, ray_directions=rays_d,
                  distance_range=config.volume_rendering.distance_range, num_samples=config.volume_rendering.num_fine_samples,
                  sdf_std_deviation=sched["std_deviation"], cosine_ratio=sched["cosine_ratio"])
    with torch.no_grad():
        *_, sampled_distances, sampled_weights = render(**kwargs)
    labels, gradients, fine_distances, _ = render(**kwargs, sampled_distances=sampled_distances, sampled_weights=sampled_weights)
    silhouette = nn.functional.binary_cross_entropy(labels.clamp(1.0e-6, 1.0 - 1.0e-6), targets, reduction="none").mean()
    eikonal = nn.functional.mse_loss(torch.norm(gradients, dim=-1), gradients.new_ones(*gradients.shape[:-1]))
    loss = silhouette + 0.01 * eikonal
    return (loss, labels, gradients, fine_distances) if return_outputs else loss


def oracle_check(step, frame, inv_proj, cam, pix, targets, sched, num_samples):
    """CHECKER (cpu_baseline leg only): the device leg's labels and loss for one batch against the CPU oracle evaluated
    on the very fine-pass sample distances the kernels placed (renderers.py:212-270 + main.py:653-687)."""
    from oracle import vsrd_oracle as oracle
    step.set_batch(pix, targets)
    out = step.run_eager(backward=False)
    torch.cuda.synchronize()
    h, w = frame.image_size
    p = pix.cpu()
    view, v, u = p // (h * w), (p // w) % h, p % w
    d = torch.nn.functional.normalize(
        torch.einsum("rmn,rn->rm", inv_proj[view], torch.stack([u, v, torch.ones_like(u)], -1).float()), dim=-1)
    prm = {k: t.detach().cpu() for k, t in step.params.items()}
    scene = oracle.Scene(prm["locations"], prm["rotations"], prm["half_extents"], prm["mlp_weights"], sched["temperature"])
    fine = out["fine_distances"].cpu()
    keep = fine.max(dim=1).values < 1e3
    with torch.no_grad():
        ref = oracle.render_pass(scene.field(), cam[view][keep], d[keep], fine[keep].t()[..., None].contiguous(),
                                 sched["std_deviation"], sched["cosine_ratio"])
    label_err = float((out["labels"].cpu()[keep] - ref[0]).abs().max())
    if not label_err < 1e-4:
        raise RuntimeError(f"bench.py: device-leg labels differ from the oracle by {label_err:.3e} (> 1e-4)")
    return dict(rays_checked=int(keep.sum()), max_label_error=label_err, tolerance=1e-4,
                what="labels of one full-size batch of the device leg vs the CPU oracle on the kernels' own fine samples")


def main_py_leg(args):
    """The reference's unmodified scripts/main.py on the drop-in package, one synthetic cfg2-shaped frame, in a child
    process (the script initialises its own process group).  ms/step from the timestamps of the script's own
    `[Training]` log records over the residual-field phase."""
    import datetime
    import glob
    import re
    import subprocess
    import tempfile
    from tools import stage_reference
    if stage_reference.reference_root() is None:
        return {"unavailable": "scripts/main.py neither mounted nor staged under baseline/_ref"}
    steps = args.main_py_steps
    warm = steps // 3
    every = max((steps - warm) // 3, 1)
    work = tempfile.mkdtemp(prefix="vsrd_main_py_")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_main.py"), "--train", "--workdir", work,
           "--config", os.path.join(ROOT, "configs", "synthetic", "vsrd", "drive_0000_synthetic", "config.json")]
    for item in (f"datasets.train.kwargs.num_frames=1", f"datasets.train.kwargs.num_instances={args.instances}",
                 f"datasets.train.kwargs.num_source_frames={args.views - 1}",
                 f"optimization.num_steps={steps}", f"optimization.warmup_steps={warm}",
                 f"volume_rendering.num_rays={args.rays}", f"volume_rendering.num_fine_samples={args.samples}",
                 f"logging.scalar_intervals={every}", "logging.image_intervals=1000000", "logging.ckpt_intervals=1000000"):
        cmd += ["--set", item]
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29561")
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    if proc.returncode != 0:
        raise RuntimeError(f"bench.py: scripts/main.py failed on the drop-in package:\n{proc.stderr[-3000:]}")
    log, = glob.glob(os.path.join(work, "logs", "**", "log.txt"), recursive=True)
    stamps = {}
    for m in re.finditer(r"INFO: (\d{4}-\d\d-\d\d \d\d:\d\d:\d\d,\d{3}): \[Training\] Rank: 0, Step: (\d+),[^\n]*runtimes", open(log).read()):
        stamps[int(m.group(2))] = datetime.datetime.strptime(m.group(1), "%Y-%m-%d %H:%M:%S,%f")
    residual = sorted(k for k in stamps if k >= warm + every - 1)
    if len(residual) < 2:
        raise RuntimeError("bench.py: too few log records from scripts/main.py")
    a, b = residual[0], residual[-1]
    ms = (stamps[b] - stamps[a]).total_seconds() * 1e3 / (b - a)
    units = args.rays * (3 * args.samples - 2)
    return {"ms_per_step": ms, "value": units / (ms * 1e-3), "unit": "ray-samples/s", "steps_timed": b - a,
            "script_sha256": stage_reference.MAIN_PY_SHA256,
            "what": "UNMODIFIED scripts/main.py (SHA-256 checked) via tools/run_main.py on the drop-in vsrd package: the "
                    "script's own per-step Python (136 project_box_3d calls, scipy matching, 9 M-way multinomial, closure "
                    "composition, autograd, torch Adam) around the fused renderer; residual-field phase"}


def run_native(args):
    import vsrd
    from vsrd_b200 import _lib, ops
    from vsrd_b200.engine import SilhouetteStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on the C-level stdout at communicator creation; the driver reads ONE JSON
        # line from stdout, so point fd 1 at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    _lib.load()

    K, W = args.steps, max(args.warmup, 3)
    total = K + W
    frame, inv_proj, cam, pool, targets = build_frame_inputs(args, rank, device, total)
    detector, hyper, encoder = make_models(args, frame, rank, device)
    sched = schedule_at(args.schedule_frac)
    units = args.rays * (3 * args.samples - 2)

    # ---------------- device-resident leg ("value") ----------------
    with torch.no_grad():
        world_out = detector()
        mlp_w = hyper(world_out["embeddings"])[0]
    step = SilhouetteStep(inv_projection=inv_proj, camera_positions=cam, image_size=frame.image_size,
                          num_rays=args.rays, num_samples=args.samples, device=device)
    step.set_parameters(world_out["locations"][0], world_out["orientations"][0], world_out["dimensions"][0], mlp_w)
    step.set_schedule(**sched)
    pool_dev, targets_dev = pool.to(device), targets.to(device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2
    step.set_batch(pool_dev[0], targets_dev[0])
    use_graph = not args.no_graph
    if use_graph:
        step.capture()
    runner = step.run if use_graph else step.run_eager

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for k in range(W):
        step.set_batch(pool_dev[k], targets_dev[k])
        runner()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    with ClockSampler(local_rank) as clocks:
        for k in range(K):
            step.set_batch(pool_dev[W + k], targets_dev[W + k])
            flush.zero_()                                  # evict L2 between timed steps (outside the events)
            starts[k].record()
            runner()
            ends[k].record()
        barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    culled_tiles, visited_tiles = ops.culling_counters(device, reset=True)
    loss_value = float(step.out["loss_parts"].sum())
    if not math.isfinite(loss_value):
        raise RuntimeError("bench.py: non-finite loss from the device-resident leg")

    # per-kernel breakdown (eager, event per launch) for the roofline of the dominant kernel
    per_kernel = {}
    reps = 10
    for k in range(reps):
        step.set_batch(pool_dev[(W + k) % total], targets_dev[(W + k) % total])
        flush.zero_()
        step.timers = []
        step.run_eager()
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(step.timers[:-1], step.timers[1:]):
            per_kernel.setdefault(n1, []).append(e0.elapsed_time(e1))
    step.timers = None
    per_kernel = {k: statistics.median(v) for k, v in per_kernel.items()}
    fwd_pairs = {name: rays.live_pairs() for name, rays in step.rays.items()}       # of the last eager step
    fwd_live = {name: (p[0] / p[1] if p else 1.0) for name, p in fwd_pairs.items()}

    # ---------------- end-to-end leg through the public API ----------------
    # vsrd_b200.frame.FrameLabeler = scripts/main.py's optimisation step for one frame (decode, projection +
    # matching + projection losses, hypernetwork, two-pass render, silhouette/eikonal loss, backward, Adam, LR
    # decay), replayed as one CUDA graph.  Every step: the ray batch (pixel indices + silhouette targets) comes
    # from pinned host memory and the step's losses are read back to the host.
    from vsrd_b200.frame import FrameLabeler, synthetic_frame_inputs
    inputs = synthetic_frame_inputs(frame, device)
    from vsrd_b200 import synthetic as _syn
    raw_loc, raw_dim, raw_ori = _syn.perturbed_raw_parameters(frame, seed=rank)
    labeler = FrameLabeler(inputs, num_steps=3000, warmup_steps=1000, num_rays=args.rays, num_samples=args.samples,
                           rays="batches", use_graph=use_graph, seed=rank, model_seed=rank,
                           initial_parameters=dict(locations=raw_loc.to(device), dimensions=raw_dim.to(device),
                                                   orientations=raw_ori.to(device)))
    start_step = max(1000, min(3000 - (K + W), int(round(3000 * args.schedule_frac)) - (K + W) // 2))   # the device leg's operating point
    labeler.seek(start_step)
    pool_pin, targets_pin = pool.pin_memory(), targets.pin_memory()
    loss_pin = [torch.zeros(5, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def e2e_enqueue(k):
        """Step k: H2D copies of its ray batch + graph replay + D2H copy of its losses, all on the labeler's stream."""
        labeler.step(pool_pin[k], targets_pin[k])
        with torch.cuda.stream(labeler.stream):
            loss_pin[k & 1].copy_(labeler.losses, non_blocking=True)
            loss_ready[k & 1].record()

    def e2e_read(k):
        loss_ready[k & 1].synchronize()
        return float(loss_pin[k & 1][0])

    def e2e_run(first, count):
        """Every step's loss is read on the host; the read of step k overlaps the execution of step k+1 (the
        reference itself reads losses every 50 steps only, main.py:872)."""
        last = float("nan")
        for k in range(first, first + count):
            e2e_enqueue(k)
            if k > first:
                last = e2e_read(k - 1)
                if not math.isfinite(last):
                    raise RuntimeError("bench.py: non-finite loss from the end-to-end leg")
        return e2e_read(first + count - 1)

    # Setup, not warm-up: the labeler runs the first three steps of a phase eagerly (allocator / cuBLAS warm-up) and
    # captures the step's CUDA graph on the fourth.  Do that here, so that no timed step can contain a capture
    # whatever W is, then put the schedule back where the measurement starts.
    if use_graph:
        for first in sorted({start_step, start_step + K + W - 1}, key=labeler.phase_of):
            if labeler.phase_of(first) in labeler._graphs:
                continue                                   # (a long run may cross into the phase with the forward culling pre-pass)
            labeler.seek(first)
            for k in range(4):
                e2e_enqueue(k)
            e2e_read(3)
        labeler.seek(start_step)
    e2e_run(0, W)
    barrier()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(labeler.stream):
        e_start.record()
    last = e2e_run(W, K)
    with torch.cuda.stream(labeler.stream):
        e_end.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_end)
    if not math.isfinite(last):
        raise RuntimeError("bench.py: non-finite loss from the end-to-end leg")

    # informational: the same step with the models as nn.Modules under autograd + torch.optim.Adam (~150 more launches)
    torch_models_ms = None
    if args.compare_torch_models:
        lab2 = FrameLabeler(inputs, num_steps=3000, warmup_steps=1000, num_rays=args.rays, num_samples=args.samples,
                            rays="batches", use_graph=use_graph, seed=rank, model_seed=rank, models="torch",
                            initial_parameters=dict(locations=raw_loc.to(device), dimensions=raw_dim.to(device),
                                                    orientations=raw_ori.to(device)))
        lab2.seek(start_step)

        def torch_step(k):
            lab2.step(pool_pin[k], targets_pin[k])
            with torch.cuda.stream(lab2.stream):
                loss_pin[0].copy_(lab2.losses, non_blocking=True)
            lab2.stream.synchronize()

        for k in range(W + 2):
            torch_step(k % total)
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(lab2.stream):
            t_start.record()
        reps2 = min(K, 20)
        for k in range(reps2):
            torch_step((W + k) % total)
        with torch.cuda.stream(lab2.stream):
            t_end.record()
        torch.cuda.synchronize()
        torch_models_ms = t_start.elapsed_time(t_end) / reps2
        del lab2

    # ---------------- eager drop-in API leg (informational): main.py's closures -> vsrd.rendering ----------------
    config = vsrd.utils.Dict.apply(dict(volume_rendering=dict(distance_range=[0.0, 100.0], num_fine_samples=args.samples)))
    params = [
        dict(params=[detector.locations], lr=1e-2), dict(params=[detector.dimensions], lr=1e-2),
        dict(params=[detector.orientations], lr=1e-2), dict(params=[detector.embeddings], lr=1e-3),
        dict(params=list(hyper.parameters()), lr=1e-4),
    ]
    optimizer = torch.optim.Adam(params, lr=1e-2)
    inv_proj_dev, cam_dev = inv_proj.to(device), cam.to(device)
    h, w = frame.image_size

    def api_step(k):
        pix = pool_pin[k].to(device, non_blocking=True)
        tgt = targets_pin[k].to(device, non_blocking=True)
        rays_o, rays_d = ops.gather_rays(inv_proj_dev, cam_dev, pix, h, w)
        optimizer.zero_grad(set_to_none=True)
        loss = main_style_step(vsrd, (detector, hyper, encoder), config, rays_o, rays_d, tgt, sched, args.instances)
        loss.backward()
        optimizer.step()
        return float(loss)

    api_steps = min(K, 20)
    for k in range(3):
        api_step(k)
    torch.cuda.synchronize()
    a_start, a_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_start.record()
    for k in range(api_steps):
        api_step(3 + k)
    a_end.record()
    torch.cuda.synchronize()
    api_ms = a_start.elapsed_time(a_end) / api_steps

    # ---------------- whole frames: the real multi-GPU path (vsrd_b200.sequence.label_sequence) ----------------
    # cfg5: frames with N ~ Poisson(6) clipped to [1,24] (seed = frame id), `frames_per_rank` x world of
    # them, partitioned across the ranks, `in_flight` frames resident per GPU, ONE padded NCCL all_gather of the final
    # boxes at the end -- all inside the clock.  Run twice: the reference's DistributedSampler partition
    # (distributed/loader.py:6-9) and the cost-balanced one; per-rank seconds expose the load imbalance.
    frames_info = None
    if args.frames > 0:
        from vsrd_b200 import sequence
        num_frames = args.frames_per_rank * world

        def instances_of(fid):
            g = torch.Generator().manual_seed(7919 + fid)
            return int(torch.poisson(torch.tensor(6.0), generator=g).clamp(1, 24))

        def make_labeler(fid):
            fr = _syn.make_frame(instances_of(fid), args.views, seed=1000 + fid)
            raw = _syn.perturbed_raw_parameters(fr, seed=fid)
            return FrameLabeler(synthetic_frame_inputs(fr, device), num_steps=args.frame_steps,
                                warmup_steps=args.frame_steps // 3, num_rays=args.rays, num_samples=args.samples, seed=fid,
                                initial_parameters=dict(locations=raw[0].to(device), dimensions=raw[1].to(device),
                                                        orientations=raw[2].to(device)))

        counts = [instances_of(f) for f in range(num_frames)]
        frames_info = dict(frames=num_frames, frames_per_rank=args.frames_per_rank, steps_per_frame=args.frame_steps,
                           in_flight=args.in_flight, instances=counts)
        for name, costs in (("sampler", None), ("balanced", [n + 2.0 for n in counts])):
            barrier()
            t0 = time.perf_counter()
            mine = sequence.my_frames(num_frames, costs=costs, seed=0)
            results = sequence.label_frames_in_flight(mine, make_labeler, args.frame_steps, args.in_flight)
            torch.cuda.synchronize()
            t_label = time.perf_counter() - t0
            merged = sequence.gather_labels(results, device=device)          # NCCL all_gather (identity at N=1)
            torch.cuda.synchronize()
            t_total = time.perf_counter() - t0
            if sorted(merged) != list(range(num_frames)) or not all(bool(torch.isfinite(b).all()) for b in merged.values()) \
                    or any(merged[f].shape[0] != counts[f] for f in merged):
                raise RuntimeError("bench.py: the frame leg did not gather every frame's finite boxes")
            ts = torch.tensor([t_label, -t_label, t_total], device=device, dtype=torch.float64)
            if world > 1:
                torch.distributed.all_reduce(ts, op=torch.distributed.ReduceOp.MAX)
            frames_info[name] = dict(seconds=float(ts[2]), rank_seconds_max=float(ts[0]), rank_seconds_min=-float(ts[1]),
                                     gather_seconds=t_total - t_label,
                                     frames_per_hour=num_frames * 3600.0 / float(ts[2]))
        frames_info["seconds"] = frames_info["balanced"]["seconds"]

    # ---------------- reduce over ranks ----------------
    t = torch.tensor([dev_ms, e2e_ms, frames_info["seconds"] if frames_info else 0.0], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if frames_info:
        frames_info["frames_per_hour"] = frames_info["balanced"]["frames_per_hour"]
        frames_info["note"] = ("whole-job frames/hour through sequence.label_sequence's pieces incl. per-frame set-up (soft masks, "
                               "CDF, models, graph capture) and the final all_gather; 3000 steps/frame as configs/kitti_360, "
                               "logging/checkpoint paths off; headline = balanced partition, `sampler` = the reference's "
                               "DistributedSampler partition")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        fma_peak_tflops = sms * 128 * 2 * sm_max * 1e6 / 1e12
        m_fine = 2 * args.samples - 1
        bwd_ms = per_kernel.get("field_backward", float("nan"))
        # algorithmic contraction flops of the backward field kernel: 4F per (fine sample, instance)
        # (reverse-over-reverse beyond the 2F forward; the kernel's own 2F recompute earns no credit)
        bwd_flops = 4 * F_MLP * args.instances * args.rays * m_fine
        if visited_tiles:      # executed = nominal * (1 - culled fraction), counted in-kernel
            bwd_flops = int(bwd_flops * (1.0 - culled_tiles / visited_tiles))
        achieved = bwd_flops / (bwd_ms * 1e-3) / 1e12 if bwd_ms == bwd_ms and bwd_ms > 0 else None
        # kernel name + DRAM bytes per launch of the CURRENT backward kernel, written by tools/ncu_summary.py --roofline
        # from the round's `ncu --set full` capture of this very command
        capture = {}
        try:
            capture = json.load(open(os.path.join(ROOT, "profiles", "roofline_kernel.json")))
        except Exception:
            pass
        same_shape = capture.get("shape") == [args.rays, args.samples, args.instances]
        hbm_bytes = 80 * args.instances * args.rays * m_fine   # 80*N B per fine ray-sample (3-kernel split)
        line = {
            "metric": "ray_samples_per_sec_fwd_bwd",
            "value": world * units * K / (dev_ms * 1e-3),
            "unit": "ray-samples/s",
            "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, world, use_graph),
            "e2e": {"value": world * units * K / (e2e_ms * 1e-3), "unit": "ray-samples/s",
                    "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": int(pool[0].numel() * 8 + targets[0].numel() * 4),
                    "d2h_bytes_per_step": 20,
                    "api": "vsrd_b200.frame.FrameLabeler.step(pixel_indices, targets): decode + projection/matching losses + "
                           "hypernetwork + two-pass render + loss + backward + Adam as one CUDA graph; models, their "
                           "backward and Adam are the vsrd_model.cu launches (no autograd)",
                    "ms_per_step_autograd_models": torch_models_ms},
            "eager_api": {"ms_per_step": api_ms, "value": units / (api_ms * 1e-3), "unit": "ray-samples/s (per GPU)",
                          "api": "vsrd.models + vsrd.rendering.hierarchical_volumetric_rendering composed from main.py's "
                                 "closures every step + autograd + Adam (no graph: Python dispatch bound)"},
            "frames": frames_info,
            "gpu_launches": SilhouetteStep.KERNELS_PER_STEP * K,
            "roofline": {"bound": "fp32_fma", "kernel": capture.get("kernel", "field_backward (see kernel_ms)"), "achieved": achieved,
                         "peak": fma_peak_tflops, "unit": "TFLOP/s",
                         "frac": (achieved / fma_peak_tflops) if achieved else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this exact shape, from the
                         # `ncu --set full` capture named in `traffic_unit` (one pass over the 25.5 MB adjoint buffer;
                         # everything else stays in L2 / shared / tensor memory)
                         "traffic": capture.get("traffic_bytes") if same_shape else None,
                         "traffic_unit": f"dram bytes read + written per launch (ncu --set full, {capture.get('source', 'no capture')})",
                         "peak_source": f"{sms} SMs x 128 lanes x 2 x sm_max_mhz {sm_max:.0f} (MEASURED_PEAKS.json clock)",
                         "bound_note": "issue-bound FP32 work between mma.sync m16n8k16 (bf16 hi + lo) contractions, stash and weight-gradient "
                                       "accumulators in tensor memory (tcgen05.st / tcgen05.ld): DRAM traffic is one pass over the adjoint "
                                       "buffer (26 MB per launch, 0.7 % of the HBM roofline) and the tensor pipe is ~32 % busy, so the kernel "
                                       "is rated against the FP32 FMA peak (DESIGN.md section 3)",
                         "forward_fine": {"kernel": ("cull_samples_kernel + field_forward_umma_kernel<4, true>" if fwd_pairs.get("fine")
                                                     else "field_forward_umma_kernel<4, false>") + " (tcgen05 / TMEM, vsrd_field_umma.cu)",
                                          "kernel_ms": per_kernel.get("field_forward_fine"),
                                          "executed_fraction": fwd_live["fine"],
                                          "achieved": (2 * F_MLP * args.instances * args.rays * m_fine * fwd_live["fine"]
                                                       / (per_kernel["field_forward_fine"] * 1e-3) / 1e12)
                                          if per_kernel.get("field_forward_fine") else None,
                                          "frac": (2 * F_MLP * args.instances * args.rays * m_fine * fwd_live["fine"]
                                                   / (per_kernel["field_forward_fine"] * 1e-3) / 1e12 / fma_peak_tflops)
                                          if per_kernel.get("field_forward_fine") else None,
                                          "note": "2F credited per executed (sample, instance); with culling the time includes the pre-pass launch"},
                         "kernel_ms": bwd_ms, "algorithmic_flops_per_launch": bwd_flops,
                         "algorithmic_hbm_bytes_per_step": hbm_bytes,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "kernel_ms": per_kernel,
            "culling": {"enabled": visited_tiles > 0, "backward_tiles_skipped": culled_tiles, "backward_tiles_visited": visited_tiles,
                        "skipped_fraction": (culled_tiles / visited_tiles) if visited_tiles else 0.0,
                        "forward_pairs_skipped_fraction": {k: 1.0 - v for k, v in fwd_live.items()},
                        "note": "instance culling (SURVEY 8d): (sample, instance) pairs whose soft-min weight is < exp(-20) "
                                "(value and gradient terms < 2^-24) skip the residual MLP (forward: per pair, both passes; backward: 16-sample tiles); "
                                "counted in-kernel, backward over warm-up + timed steps of the device-resident leg, forward "
                                "on its last step; `value` counts nominal ray-samples"},
            "clocks": clocks.summary(),
            "loss": loss_value,
        }
        if world == 1 and args.main_py_steps > 0:
            line["main_py"] = main_py_leg(args)
        if world == 1 and not args.skip_cpu_baseline:
            line["oracle_check"] = oracle_check(step, frame, inv_proj, cam, pool_dev[0], targets_dev[0], sched, args.samples)
            res = cpu_reference_run(args, steps=args.cpu_baseline_steps, warmup=1, num_rays=args.cpu_sample_rays or args.rays)
            line["cpu_baseline"] = {"value": res["value"], "unit": "ray-samples/s", "cores": res["cores"],
                                    "kind": res["kind"], "sample": res["sample"], "ms_per_step": res["ms_per_step"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
