"""Full-size parity cases (test infrastructure): BASELINE.json configs[1] / configs[2] at their stated sizes
(N=8 street frame; cfg3: N=24 parking grid; 17 views at 376x1408, R=1000 rays, S=100 samples)
evaluated by the CPU oracle in fp32 (the reference's own precision) and fp64 (ground truth).

The oracle needs ~10-60 s per case on a 16-core host, so results are cached under tests/golden/_cache/ (git-ignored,
rebuilt from the seeds below when absent; `python -m tests.fullsize_cases` pre-builds them).  Everything is derived from
seeded generators; the cache key carries VERSION so stale files are never used.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from oracle import vsrd_oracle as oracle

VERSION = 4
CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "_cache")
NUM_RAYS, NUM_SAMPLES = 1000, 100
SCHEDULES = {           # annealing endpoints + mid-point (config.json:231-234)
    "early": dict(temperature=1.0, std_deviation=1.0, cosine_ratio=0.33),
    "mid": dict(temperature=0.55, std_deviation=0.55, cosine_ratio=0.67),
    "late": dict(temperature=0.1, std_deviation=0.1, cosine_ratio=1.0),
}
CONFIGS = {
    "cfg2": dict(num_instances=8, layout="street", seed=2),
    "cfg3": dict(num_instances=24, layout="parking", seed=3),
}
CASES = [("cfg2", "early"), ("cfg2", "mid"), ("cfg2", "late"), ("cfg3", "mid"), ("cfg3", "late")]
GRAD_NAMES = ["locations", "rotations", "half_extents", "mlp_weights"]


def scene_inputs(cfg: str):
    """Scene parameters, rays and targets of one configuration (fp32, host)."""
    from vsrd_b200 import synthetic
    c = CONFIGS[cfg]
    frame = synthetic.make_frame(num_instances=c["num_instances"], num_views=17, layout=c["layout"], seed=c["seed"])
    gen = torch.Generator().manual_seed(100 + c["seed"])
    n = frame.num_instances
    # predicted boxes: the ground truth disturbed by ~0.3 m / 0.1 rad (what the optimiser sees mid-run)
    loc = frame.gt_locations + torch.randn(n, 3, generator=gen) * torch.tensor([0.3, 0.05, 0.3])
    yaw = frame.gt_yaws + torch.randn(n, generator=gen) * 0.1
    rot = synthetic.rotation_y(yaw)
    lo, hi = torch.tensor(synthetic.DIMENSION_RANGE[0]), torch.tensor(synthetic.DIMENSION_RANGE[1])
    dim = lo + torch.rand(n, 3, generator=gen) * (hi - lo)
    torch.manual_seed(7)
    hyper = oracle.HyperNetwork()
    emb = torch.rand(1, 256, generator=gen) + 0.1 * torch.randn(n, 256, generator=gen)
    with torch.no_grad():
        weights = hyper(emb)
    pix = frame.draw_pixel_indices(NUM_RAYS + NUM_RAYS // 2, gen)
    h, w = frame.image_size
    view, v, u = pix // (h * w), (pix // w) % h, pix % w
    inv_proj, cam = frame.inverse_projections()
    dirs = torch.nn.functional.normalize(
        torch.einsum("rmn,rn->rm", inv_proj[view], torch.stack([u, v, torch.ones_like(u)], -1).float()), dim=-1)
    origins = cam[view].contiguous()
    r = pix.numel()
    targets = torch.rand(r, n, generator=gen) * (torch.rand(r, n, generator=gen) < 0.3).float()
    jitter = torch.rand(r, 1, NUM_SAMPLES, generator=gen)
    uniforms = torch.sort(torch.rand(r, 1, NUM_SAMPLES, generator=gen), dim=-1).values
    return dict(locations=loc, rotations=rot, half_extents=dim, mlp_weights=weights, origins=origins,
                directions=dirs, targets=targets, jitter=jitter, sorted_uniforms=uniforms)


def training_loss(labels, grads, targets):
    """BCE + 0.01 eikonal, scripts/main.py:653-687, 855."""
    return oracle.silhouette_loss(labels, targets) + 0.01 * oracle.eikonal_loss(grads)


def linear_coefficients(r, m, n):
    gen = torch.Generator().manual_seed(11)
    return (torch.randn(r, n, generator=gen, dtype=torch.float64),
            torch.randn(m, r, 3, generator=gen, dtype=torch.float64) * 0.1,
            torch.randn(m, r, 1, generator=gen, dtype=torch.float64))


def linear_loss(labels, grads, weights, coeffs):
    """Random linear functional of the renderer's three outputs (sample-major): isolates the kernels' adjoint."""
    cl, cg, cw = (c.to(labels.device, labels.dtype) for c in coeffs)
    return (labels * cl).sum() + (grads * cg).sum() + (weights * cw).sum()


def _oracle_pass(inp, sched, keep, fine_sm, dtype, with_linear):
    leaves = [inp[k][...].to(dtype).clone().requires_grad_(True) for k in GRAD_NAMES]
    scene = oracle.Scene(*leaves[:3], leaves[3], sched["temperature"])
    out = oracle.render_pass(scene.field(), inp["origins"][keep].to(dtype), inp["directions"][keep].to(dtype),
                             fine_sm.to(dtype), sched["std_deviation"], sched["cosine_ratio"])
    labels, grads, _, weights = out
    res = dict(labels=labels.detach(), gradients=grads.detach(), weights=weights.detach())
    loss = training_loss(labels, grads, inp["targets"][keep].to(dtype))
    g = torch.autograd.grad(loss, leaves, retain_graph=with_linear)
    res["loss"] = loss.detach()
    res.update({f"grad_{k}": v for k, v in zip(GRAD_NAMES, g)})
    if with_linear:
        coeffs = linear_coefficients(labels.shape[0], grads.shape[0], labels.shape[1])
        lin = linear_loss(labels, grads, weights, coeffs)
        g = torch.autograd.grad(lin, leaves)
        res["lin_loss"] = lin.detach()
        res.update({f"lin_grad_{k}": v for k, v in zip(GRAD_NAMES, g)})
    return res


def compute_case(cfg: str, sched_name: str):
    torch.set_num_threads(os.cpu_count() or 1)
    inp = scene_inputs(cfg)
    sched = SCHEDULES[sched_name]
    scene = oracle.Scene(inp["locations"], inp["rotations"], inp["half_extents"], inp["mlp_weights"], sched["temperature"])
    with torch.no_grad():     # the reference's own coarse pass + importance resampling, fp32
        _, _, cd, cw, fd, _ = oracle.two_pass_render(
            scene.field(), inp["origins"], inp["directions"], [0.0, 100.0], NUM_SAMPLES, sched["std_deviation"],
            sched["cosine_ratio"], jitter=inp["jitter"], sorted_uniforms=inp["sorted_uniforms"])
    hit = fd.squeeze(-1).max(dim=0).values < 1e3       # rays whose resampling did not extrapolate
    keep = torch.nonzero(hit).squeeze(-1)[:NUM_RAYS]
    assert keep.numel() == NUM_RAYS, f"{cfg}/{sched_name}: only {keep.numel()} hit rays"
    miss = torch.nonzero(~hit).squeeze(-1)
    fine_sm = fd[:, keep].contiguous()
    out = dict(keep=keep, miss=miss, coarse_distances=cd.squeeze(-1).t().contiguous(),
               coarse_weights=cw.squeeze(-1).t().contiguous(), fine_all=fd.squeeze(-1).t().contiguous())
    f32 = _oracle_pass(inp, sched, keep, fine_sm, torch.float32, with_linear=True)
    f64 = _oracle_pass(inp, sched, keep, fine_sm, torch.float64, with_linear=True)
    out.update({f"f32_{k}": v for k, v in f32.items()})
    out.update({f"f64_{k}": v for k, v in f64.items()})
    if miss.numel():      # labels of the missed rays (extrapolated samples), fp32 oracle, no gradients
        with torch.no_grad():
            o = oracle.render_pass(scene.field(), inp["origins"][miss], inp["directions"][miss],
                                   fd[:, miss].contiguous(), sched["std_deviation"], sched["cosine_ratio"])
        out["miss_labels"] = o[0]
    return {k: v.numpy() for k, v in out.items()}


def get_case(cfg: str, sched_name: str):
    """dict of host tensors: inputs (`scene_inputs`) + oracle outputs (`f32_*`, `f64_*`, placement)."""
    path = os.path.join(CACHE_DIR, f"fullsize_v{VERSION}_{cfg}_{sched_name}.npz")
    if os.path.exists(path):
        data = dict(np.load(path))
    else:
        data = compute_case(cfg, sched_name)
        os.makedirs(CACHE_DIR, exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, **data)
        os.replace(tmp, path)
    case = {k: torch.from_numpy(v) for k, v in data.items()}
    case.update(scene_inputs(cfg))
    case["schedule"] = SCHEDULES[sched_name]
    return case


if __name__ == "__main__":
    import time
    for cfg, s in CASES:
        t0 = time.time()
        c = get_case(cfg, s)
        e = {k: float((c[f"f32_grad_{k}"].double() - c[f"f64_grad_{k}"]).norm() / c[f"f64_grad_{k}"].norm()) for k in GRAD_NAMES}
        print(f"{cfg}/{s}: {time.time() - t0:.1f} s, misses {c['miss'].numel()}, "
              f"label err fp32 vs fp64 {float((c['f32_labels'].double() - c['f64_labels']).abs().max()):.2e}, "
              f"fp32 grad rel err vs fp64 { {k: f'{v:.1e}' for k, v in e.items()} }", flush=True)
