#!/usr/bin/env python
"""Times the inference / logging renderers (scripts/main.py:1011-1041) on one KITTI-360-shaped
view: the full-image volumetric silhouette render (the reference: 376 calls x 2 passes, one image row each) and
full-image sphere tracing + surface normals (the reference: up to 1000 iterations with a host sync each)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vsrd
from vsrd_b200 import ops, surface, synthetic
from vsrd.rendering.renderers import UnionField

dev = torch.device("cuda", 0)
n = 8
frame = synthetic.make_frame(n, 17, seed=0)
torch.manual_seed(0)
hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=256,
                                       hyper_out_channels_list=[256] * 4).to(dev)
with torch.no_grad():
    weights = hyper(torch.rand(1, n, 256, device=dev))[0]
field = UnionField(frame.gt_locations.to(dev), frame.gt_rotations.to(dev), frame.gt_half_extents.to(dev), weights, 0.3, 100.0)
inv_proj, cam = frame.inverse_projections()
h, w = frame.image_size
view = frame.num_views // 2
dirs = ops.ray_directions(inv_proj[view:view + 1].to(dev), h, w)[0]          # [H,W,3]
origin = cam[view].to(dev)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


with torch.no_grad():
    t_render, labels = timed(lambda: surface.render_image(field, origin, dirs, num_samples=100, std_deviation=0.3, cosine_ratio=0.5))
    t_trace, traced = timed(lambda: surface.sphere_trace(field, origin.expand(h * w, 3).contiguous(), dirs.reshape(-1, 3), 64, 1e-3, bounding_radius=100.0))
rays = h * w
print(json.dumps(dict(image=[h, w], instances=n, render_image_s=t_render, render_ray_samples_per_s=rays * 298 / t_render,
                      sphere_trace_64_iterations_s=t_trace, label_mass=float(labels.sum(-1).mean()))))
