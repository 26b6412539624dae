#!/usr/bin/env python
"""cfg3 / cfg4 of BASELINE.json: fwd+bwd renderer microbenchmark over rays/step, samples/ray and instance
count on the raw kernel sequence (vsrd_b200.engine.SilhouetteStep, eager launches, CUDA events per kernel).
Prints one JSON line per point: ray-samples/s, per-kernel ms, and the achieved GB/s of the compositing kernels
against their algorithmic traffic (16 N B read per sample forward, 32 N B read+write backward).

    python tools/sweep.py [--rays 1000,4000,...] [--samples 100] [--instances 8,24] [--reps 5]
"""
import argparse, json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsrd_b200 import synthetic
from vsrd_b200.engine import SilhouetteStep
import vsrd

ap = argparse.ArgumentParser()
ap.add_argument("--rays", default="1000,4000,16000,64000,256000")
ap.add_argument("--samples", default="100")
ap.add_argument("--instances", default="8,24")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
F_MLP = 3234
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in [int(x) for x in a.instances.split(",")]:
    frame = synthetic.make_frame(n, 17, seed=0, layout="parking" if n > 8 else "street")
    inv_proj, cam = frame.inverse_projections()
    torch.manual_seed(0)
    detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n).to(dev)
    hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=256,
                                           hyper_out_channels_list=[256] * 4).to(dev)
    with torch.no_grad():
        mlp_w = hyper(detector.embeddings)[0]
    gt = (frame.gt_locations, frame.gt_rotations, frame.gt_half_extents)
    for s in [int(x) for x in a.samples.split(",")]:
        for r in [int(x) for x in a.rays.split(",")]:
            gen = torch.Generator().manual_seed(r)
            pix = frame.draw_pixel_indices(r, gen).to(dev)
            targets = torch.rand(r, n, generator=gen).to(dev)
            step = SilhouetteStep(inv_projection=inv_proj, camera_positions=cam, image_size=frame.image_size,
                                  num_rays=r, num_samples=s, device=dev)
            step.set_parameters(*gt, mlp_w)
            step.set_schedule(temperature=0.55, std_deviation=0.55, cosine_ratio=0.5)
            step.set_batch(pix, targets)
            for _ in range(2):
                step.run_eager()
            torch.cuda.synchronize()
            per, totals = {}, []
            for k in range(a.reps):
                flush.zero_()
                step.timers = []
                step.run_eager()
                torch.cuda.synchronize()
                t = step.timers
                totals.append(t[0][1].elapsed_time(t[-1][1]))
                for (n0, e0), (n1, e1) in zip(t[:-1], t[1:]):
                    per.setdefault(n1, []).append(e0.elapsed_time(e1))
            step.timers = None
            per = {k: statistics.median(v) for k, v in per.items()}
            ms = statistics.median(totals)
            m_f, m_c = 2 * s - 1, s - 1
            units = r * (3 * s - 2)
            gb = lambda bytes_, t_ms: bytes_ / (t_ms * 1e-3) / 1e9
            line = dict(
                workload=f"N={n} ({'parking' if n > 8 else 'street'}) R={r} S={s}", ms_per_step=ms, ray_samples_per_s=units / (ms * 1e-3),
                kernel_ms=per,
                composite_forward_fine_gbs=gb((16 * n + 16) * r * m_f, per["composite_forward_fine"]),
                composite_backward_gbs=gb(32 * n * r * m_f, per["composite_backward"]),
                field_backward_tflops=4 * F_MLP * n * r * m_f / (per["field_backward"] * 1e-3) / 1e12,
                field_forward_fine_tflops=2 * F_MLP * n * r * m_f / (per["field_forward_fine"] * 1e-3) / 1e12,
                hbm_peak_gbs=peaks.get("hbm_gbs"), loss=float(step.out["loss_parts"].sum()))
            print(json.dumps(line), flush=True)
            del step
            torch.cuda.empty_cache()
