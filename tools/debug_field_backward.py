#!/usr/bin/env python
"""Compare the tensor-core field backward (default) with the SIMT cross-check kernel
(VSRD_FIELD_IMPL=simt) on random adjoints, parameter group by parameter group.

    python tools/debug_field_backward.py [--rays R] [--intervals M] [--instances N]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vsrd_b200 import ops, synthetic  # noqa: E402

GROUPS = [("W0", 0, 784), ("W1", 784, 1056), ("W2", 1056, 1328), ("W3", 1328, 1600), ("W4", 1600, 1617)]


def scene_and_rays(n, r, m, seed, dev):
    gen = torch.Generator().manual_seed(seed)
    frame = synthetic.make_frame(num_instances=n, num_views=2, image_size=(94, 352), seed=seed, intrinsics_scale=0.25)
    loc, rot, dim = frame.gt_locations, frame.gt_rotations, frame.gt_half_extents
    w = torch.randn(n, ops.MLP_WEIGHTS, generator=gen) * 0.3
    # rays through the boxes so that samples land near / inside them
    pick = torch.randint(0, n, (r,), generator=gen)
    target = loc[pick] + (torch.rand(r, 3, generator=gen) * 2 - 1) * torch.tensor([2.0, 1.5, 3.0])
    origin = torch.zeros(r, 3)
    dirs = torch.nn.functional.normalize(target - origin, dim=-1)
    depth = (target - origin).norm(dim=-1, keepdim=True)
    dist = depth + torch.sort(torch.rand(r, m + 1, generator=gen) * 8 - 4, dim=-1).values
    adj = torch.randn(n, r * m, 4, generator=gen)
    adj[:, :, 1:] *= 0.1
    scene = ops.SceneArgs(loc.to(dev), rot.to(dev), dim.to(dev), w.to(dev), 0.5)
    rays = ops.RayArgs(origin.to(dev), dirs.to(dev), dist.to(dev))
    return scene, rays, adj.to(dev)


def run(scene, rays, adj, impl):
    if impl == "simt":
        os.environ["VSRD_FIELD_IMPL"] = "simt"
    else:
        os.environ.pop("VSRD_FIELD_IMPL", None)
    out = ops.field_backward(scene, rays, adj)
    torch.cuda.synchronize()
    os.environ.pop("VSRD_FIELD_IMPL", None)
    return [t.double().cpu() for t in out]


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def compare(n, r, m, seed=0, verbose=True):
    dev = torch.device("cuda", 0)
    scene, rays, adj = scene_and_rays(n, r, m, seed, dev)
    got = run(scene, rays, adj, "mma")
    want = run(scene, rays, adj, "simt")
    worst = 0.0
    rows = []
    for name, a, b in zip(["loc", "rot", "dim"], got[:3], want[:3]):
        rows.append((name, rel(a, b)))
    for name, lo, hi in GROUPS:
        rows.append((name, rel(got[3][:, lo:hi], want[3][:, lo:hi])))
        if name != "W4":
            fan = 49 if name == "W0" else 17
            ga = got[3][:, lo:hi].reshape(n, 16, fan)
            gb = want[3][:, lo:hi].reshape(n, 16, fan)
            rows.append((name + ".bias", rel(ga[..., -1], gb[..., -1])))
    for name, e in rows:
        worst = max(worst, e)
        if verbose:
            print(f"  N={n} R={r} M={m}  {name:8s} rel-L2 {e:.3e}")
    finite = all(torch.isfinite(t).all() for t in got)
    return worst, finite


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=0)
    ap.add_argument("--intervals", type=int, default=31)
    ap.add_argument("--instances", type=int, default=3)
    args = ap.parse_args()
    cases = [(args.instances, args.rays, args.intervals)] if args.rays else [
        (1, 1, 1), (3, 64, 31), (8, 257, 199), (24, 100, 63), (5, 1000, 199)]
    bad = 0
    for n, r, m in cases:
        worst, finite = compare(n, r, m)
        print(f"N={n} R={r} M={m}: worst rel-L2 {worst:.3e} finite={finite}")
        bad += (worst > 2e-4) or not finite
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
