"""Optimisation-parity cases (test infrastructure): the whole per-frame loop of scripts/main.py:328-865 run by the CPU
oracle in fp32 (the reference's precision) AND fp64 (ground truth) on seeded draws, cached under tests/golden/_cache
(the fp64 run of the full-size case takes minutes).  The GPU test replays the same draws through FrameLabeler.

The yardstick matters: the loop is chaotic (importance resampling turns a 1-ulp CDF difference into a different sample,
Adam normalises tiny gradients), so after 100 steps even the reference's own fp32 arithmetic has drifted from fp64 by
centimetres.  `get_case` therefore returns both trajectories; the test asks the CUDA boxes to agree with the fp32
oracle to >= 0.99 3D IoU (BASELINE north_star) where fp32 and fp64 themselves agree that well, and otherwise to be no
further from fp64 than the fp32 reference is.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import frame_oracle as fo
from oracle import vsrd_oracle as oracle

VERSION = 2
CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "_cache")
CASES = {
    # name: (frame kwargs, seed, steps, warm-up, rays, samples)
    "small": (dict(num_instances=3, num_views=3, image_size=(94, 352), intrinsics_scale=0.25), 4, 36, 12, 160, 20),
    "cfg1": (dict(num_instances=4, num_views=2, image_size=(94, 352), intrinsics_scale=0.25), 4, 100, 33, 96, 16),
    # BASELINE.json configs[0] at its stated size
    "cfg1_full": (dict(num_instances=4, num_views=2, image_size=(94, 352), intrinsics_scale=0.25), 4, 100, 33, 1000, 100),
}
MODEL_SEED = 0


def case_inputs(name):
    """Frame, initial raw parameters and the per-step draws (pixel indices from the ORACLE's soft masks)."""
    from vsrd_b200 import synthetic
    kwargs, seed, steps, warm, r, s = CASES[name]
    frame = synthetic.make_frame(seed=seed, **kwargs)
    raw = synthetic.perturbed_raw_parameters(frame, seed=seed)
    sup = synthetic.frame_supervision(frame)
    v, n = frame.num_views, frame.num_instances
    soft = torch.zeros(v, *frame.image_size, n)
    for vi in range(v):
        for ni in range(n):
            k = int(sup.polygon_sizes[vi, ni])
            if k >= 3:
                soft[vi, :, :, ni] = fo.soft_mask(sup.polygons[vi, ni, :k], frame.image_size)
    gen = torch.Generator().manual_seed(0)
    weights = fo.ray_weights(soft)
    pix = torch.stack([torch.multinomial(weights, r, replacement=False, generator=gen) for _ in range(steps)])
    jit = torch.rand(steps, r, s, generator=gen)
    uni = torch.sort(torch.rand(steps, r, s, generator=gen), dim=-1).values
    return dict(frame=frame, sup=sup, raw=raw, soft=soft, pix=pix, jitter=jit, uniforms=uni, steps=steps, warmup=warm,
                num_rays=r, num_samples=s)




def run_oracle(name, dtype):
    """The optimisation loop on the CPU oracle in `dtype`; returns final corners [N,8,3] and the losses of step 0 / warm-up."""
    import vsrd
    torch.set_num_threads(os.cpu_count() or 1)
    c = case_inputs(name)
    frame, sup, steps, warm, s = c["frame"], c["sup"], c["steps"], c["warmup"], c["num_samples"]
    n, (h, w) = frame.num_instances, frame.image_size
    torch.manual_seed(MODEL_SEED)                          # same construction order as FrameLabeler.__init__
    detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n)
    hyper_module = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16, 16, 16, 16],
                                                  hyper_in_channels=256, hyper_out_channels_list=[256] * 4)
    raw = [t.clone().to(dtype).requires_grad_(True) for t in c["raw"]]
    emb = detector.embeddings.detach()[0].clone().to(dtype).requires_grad_(True)
    hyper = oracle.HyperNetwork().to(dtype)
    hyper.load_state_dict({k: v.detach().to(dtype) for k, v in hyper_module.state_dict().items()})
    opt = torch.optim.Adam([dict(params=[raw[0]], lr=1e-2), dict(params=[raw[1]], lr=1e-2), dict(params=[raw[2]], lr=1e-2),
                            dict(params=[emb], lr=1e-3), dict(params=list(hyper.parameters()), lr=1e-4)], lr=1e-2)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.01 ** (1.0 / steps))
    inv_proj, cam = frame.inverse_projections()
    inv_proj, cam = inv_proj.to(dtype), cam.to(dtype)
    ext, intr, gt2d = frame.extrinsics.to(dtype), frame.intrinsics.to(dtype), sup.boxes_2d.to(dtype)
    soft = c["soft"].to(dtype)
    marks = {}
    previous_default = torch.get_default_dtype()
    torch.set_default_dtype(dtype)                         # the renderer builds its bin edges in the default dtype
    try:
        boxes = _optimise(c, raw, emb, hyper, opt, sched, inv_proj, cam, ext, intr, gt2d, soft, dtype, marks)
    finally:
        torch.set_default_dtype(previous_default)
    return boxes.double(), marks


def _optimise(c, raw, emb, hyper, opt, sched, inv_proj, cam, ext, intr, gt2d, soft, dtype, marks):
    frame, sup, steps, warm, s = c["frame"], c["sup"], c["steps"], c["warmup"], c["num_samples"]
    h, w = frame.image_size
    for step in range(steps):
        sc = fo.schedule(step, steps, warm)
        loc, dim, rot = oracle.decode_box_parameters(*raw)
        corners = oracle.box_corners(loc, dim, rot)
        _, gt_idx, iou, l1 = fo.projection_step(corners, ext, intr, (h, w), gt2d, sup.visible, sup.target_view)
        p = c["pix"][step]
        view, v, u = p // (h * w), (p // w) % h, p % w
        d = torch.nn.functional.normalize(
            torch.einsum("rmn,rn->rm", inv_proj[view], torch.stack([u, v, torch.ones_like(u)], -1).to(dtype)), dim=-1)
        targets = fo.gather_targets(soft, p, gt_idx)
        mlp = hyper(emb) if step >= warm else None
        scene = oracle.Scene(loc, rot, dim, mlp, sc["temperature"])
        loss, _ = oracle.render_loss(scene, cam[view], d, targets, num_samples=s, distance_range=[0.0, 100.0],
                                     sdf_std_deviation=sc["std_deviation"], cosine_ratio=sc["cosine_ratio"], eikonal_weight=0.01,
                                     jitter=c["jitter"][step][:, None, :].to(dtype), sorted_uniforms=c["uniforms"][step][:, None, :].to(dtype))
        loss = loss + 0.1 * iou + 1.0 * l1
        if step in (0, warm):
            marks[step] = float(loss.detach())
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if step + 1 == steps // 2:                         # boxes half way: before the chaotic amplification sets in
            with torch.no_grad():
                marks["half"] = oracle.box_corners(*[oracle.decode_box_parameters(*raw)[i] for i in (0, 1, 2)]).double().clone()
    with torch.no_grad():
        loc, dim, rot = oracle.decode_box_parameters(*raw)
        return oracle.box_corners(loc, dim, rot)


def get_case(name):
    path = os.path.join(CACHE_DIR, f"optim_v{VERSION}_{name}.npz")
    if os.path.exists(path):
        data = dict(np.load(path))
    else:
        b32, m32 = run_oracle(name, torch.float32)
        b64, m64 = run_oracle(name, torch.float64)
        warm = CASES[name][3]
        data = dict(boxes_f32=b32.numpy(), boxes_f64=b64.numpy(), half_f32=m32["half"].numpy(), half_f64=m64["half"].numpy(),
                    loss_first=np.float64(m32[0]), loss_warm=np.float64(m32[warm]))
        os.makedirs(CACHE_DIR, exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, **data)
        os.replace(tmp, path)
    case = case_inputs(name)
    case.update({k: torch.from_numpy(np.asarray(v)) for k, v in data.items()})
    with torch.no_grad():
        loc, dim, rot = oracle.decode_box_parameters(*case["raw"])
        case["boxes_init"] = oracle.box_corners(loc, dim, rot).double()
    return case


if __name__ == "__main__":
    import time
    import vsrd
    rot = vsrd.operations.rotation_matrix_x(torch.tensor(-np.pi / 2.0)).double()
    for name in CASES:
        t0 = time.time()
        c = get_case(name)
        ious = [float(vsrd.operations.box_3d_iou_exact(a @ rot.T, b @ rot.T)[0]) for a, b in zip(c["boxes_f32"], c["boxes_f64"])]
        print(f"{name}: half way: fp32 vs fp64 max corner difference {float((c['half_f32'] - c['half_f64']).abs().max()):.5f} m")
        print(f"{name}: {time.time() - t0:.0f} s; fp32 oracle vs fp64 oracle after {c['steps']} steps: max corner difference "
              f"{float((c['boxes_f32'] - c['boxes_f64']).abs().max()):.5f} m, 3D IoU {[round(v, 5) for v in ious]}; boxes moved "
              f"{float((c['boxes_f32'] - c['boxes_init']).abs().max()):.3f} m", flush=True)
