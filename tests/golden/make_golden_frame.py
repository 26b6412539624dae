"""Golden fixtures for the frame-level rows (a14 projection, a15 matching + projection losses, soft-mask
distance map) from the UNMODIFIED reference.  Build container only (needs /root/reference):

    python tests/golden/make_golden_frame.py        ->  tests/golden/frame.npz

What runs: the reference's own `vsrd.operations.project_box_3d` / `clip_lines_to_front`
(geometric_operations.py:343-389) inside the loop of scripts/main.py:339-415 re-typed here verbatim in
structure (it is inline script code and cannot be imported), with torchvision.ops.clip_boxes_to_image /
distance_box_iou / distance_box_iou_loss, scipy.optimize.linear_sum_assignment and
nn.functional.smooth_l1_loss exactly as the script calls them; autograd supplies the gradients w.r.t.
the world-frame corners.  `SoftRasterizer.make_distance_map` (transforms/geometric_transforms.py:267-290)
is compiled from the reference's source text (the module itself imports skimage, absent here).
"""
from __future__ import annotations

import ast
import math
import os
import sys

import numpy as np
import scipy as sp
import scipy.optimize  # noqa: F401
import torch
import torch.nn as nn
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402

LINE_INDICES = [
    [0, 1], [1, 2], [2, 3], [3, 0],
    [4, 5], [5, 6], [6, 7], [7, 4],
    [0, 4], [1, 5], [2, 6], [3, 7],
]


def reference_make_distance_map():
    path = os.path.join(ref_import.REFERENCE_ROOT, "vsrd", "transforms", "geometric_transforms.py")
    tree = ast.parse(open(path).read(), filename=path)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "make_distance_map":
            module = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(module)
            scope = {"torch": torch}
            exec(compile(module, path, "exec"), scope)
            return lambda polygons, image_size: scope["make_distance_map"](None, polygons, image_size)
    raise RuntimeError("make_distance_map not found")


def scene(out_dtype, seed=0):
    # built in float64 and cast, so the f32 and f64 variants of a case describe the SAME scene
    dtype = torch.float64
    gen = torch.Generator().manual_seed(seed)
    n, v = 6, 5
    h, w = 94, 352
    k = torch.tensor([[552.554 / 4, 0.0, 682.049 / 4], [0.0, 552.554 / 4, 238.770 / 4], [0.0, 0.0, 1.0]], dtype=dtype)
    # boxes in front of the target camera; one very close (straddles z=0 in later views), one off to the side
    centres = torch.stack([
        torch.tensor([-3.0, 2.5, 0.5, -1.0, 6.5, 1.2], dtype=dtype),
        torch.full((n,), 0.675, dtype=dtype),
        torch.tensor([12.0, 18.0, 3.2, 25.0, 9.0, 1.4], dtype=dtype),
    ], dim=-1)
    half = torch.tensor([0.8, 0.8, 2.0], dtype=dtype) + torch.rand(n, 3, generator=gen, dtype=dtype) * 0.2
    yaw = torch.rand(n, generator=gen, dtype=dtype) * 2 * math.pi
    c, s = torch.cos(yaw), torch.sin(yaw)
    o, z = torch.ones_like(c), torch.zeros_like(c)
    rot = torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)
    signs = torch.tensor([[-1, -1, +1], [+1, -1, +1], [+1, -1, -1], [-1, -1, -1],
                          [-1, +1, +1], [+1, +1, +1], [+1, +1, -1], [-1, +1, -1]], dtype=dtype)
    boxes = (signs * half[:, None, :]) @ rot.transpose(-2, -1) + centres[:, None, :]
    # cameras moving forward 1.5 m per view with a small yaw drift; view 1 is the target
    extr = []
    for i in range(v):
        ang = torch.tensor((i - 1) * 0.03, dtype=dtype)
        r = torch.stack([torch.stack([torch.cos(ang), z[0], torch.sin(ang)]), torch.stack([z[0], o[0], z[0]]),
                         torch.stack([-torch.sin(ang), z[0], torch.cos(ang)])])
        pos = torch.tensor([0.1 * (i - 1), 0.0, 1.5 * (i - 1)], dtype=dtype)
        e = torch.eye(4, dtype=dtype)
        e[:3, :3] = r.T
        e[:3, 3] = -(r.T @ pos)
        extr.append(e)
    extr = torch.stack(extr)
    intr = k.repeat(v, 1, 1)
    return boxes.to(out_dtype), extr.to(out_dtype), intr.to(out_dtype), (h, w)


def run_projection(ref, dtype, shuffle, seed):
    boxes_gt, extr, intr, image_size = scene(dtype, seed)
    gen = torch.Generator().manual_seed(seed + 1)
    n, v = boxes_gt.shape[0], extr.shape[0]
    target_view = 1

    def project_all(world_boxes_3d):            # scripts/main.py:339-367
        world = nn.functional.pad(world_boxes_3d[None], (0, 1), mode="constant", value=1.0)
        out = []
        for e, kmat in zip(extr, intr):
            camera_boxes_3d = torch.einsum("bmn,b...n->b...m", e[None], world)
            camera_boxes_3d = camera_boxes_3d[..., :-1] / camera_boxes_3d[..., -1:]
            camera_boxes_2d = torch.stack([
                torch.stack([
                    ref.geometric_operations.project_box_3d(box_3d=b, line_indices=LINE_INDICES, intrinsic_matrix=km)
                    for b in cb
                ], dim=0)
                for cb, km in zip(camera_boxes_3d, kmat[None])
            ], dim=0)
            camera_boxes_2d = torchvision.ops.clip_boxes_to_image(
                boxes=camera_boxes_2d.flatten(-2, -1), size=image_size).unflatten(-1, (2, 2))
            out.append(camera_boxes_2d[0])
        return torch.stack(out, dim=0)          # [V,N,2,2]

    with torch.no_grad():
        gt_boxes_2d = project_all(boxes_gt)
    perm = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
    gt_boxes_2d = gt_boxes_2d[:, perm]          # ground truth listed in a different instance order
    visible = torch.rand(v, n, generator=gen) > 0.25
    visible[target_view] = True

    noise = torch.randn(n, 1, 3, generator=gen, dtype=torch.float64) * torch.tensor([0.4, 0.05, 0.6], dtype=torch.float64)
    world = boxes_gt + noise.to(dtype)
    world = world.clone().requires_grad_(True)
    pd_boxes_2d = project_all(world)

    cost = -torchvision.ops.distance_box_iou(boxes1=pd_boxes_2d[target_view].flatten(-2, -1),
                                             boxes2=gt_boxes_2d[target_view].flatten(-2, -1))
    pd_indices, gt_indices = map(torch.as_tensor, sp.optimize.linear_sum_assignment(cost.detach().numpy()))

    iou_projection_loss = torch.mean(torch.cat([                    # scripts/main.py:391-402
        torchvision.ops.distance_box_iou_loss(
            boxes1=pd[pd_indices[vis[gt_indices]], ...].flatten(-2, -1),
            boxes2=gt[gt_indices[vis[gt_indices]], ...].flatten(-2, -1),
            reduction="none")
        for pd, gt, vis in zip(pd_boxes_2d, gt_boxes_2d, visible)
    ], dim=0))
    l1_projection_loss = torch.mean(torch.cat([                     # scripts/main.py:404-415
        nn.functional.smooth_l1_loss(
            input=pd[pd_indices[vis[gt_indices]], ...].flatten(-2, -1),
            target=gt[gt_indices[vis[gt_indices]], ...].flatten(-2, -1),
            reduction="none")
        for pd, gt, vis in zip(pd_boxes_2d, gt_boxes_2d, visible)
    ], dim=0))
    g_iou, = torch.autograd.grad(iou_projection_loss, world, retain_graph=True)
    g_l1, = torch.autograd.grad(l1_projection_loss, world)
    return dict(world_boxes=world.detach(), extrinsics=extr, intrinsics=intr,
                image_size=torch.tensor(image_size), target_view=torch.tensor(target_view),
                gt_boxes_2d=gt_boxes_2d.flatten(-2, -1), visible=visible,
                boxes_2d=pd_boxes_2d.detach().flatten(-2, -1), cost=cost.detach(),
                pd_indices=pd_indices, gt_indices=gt_indices,
                iou_loss=iou_projection_loss.detach(), l1_loss=l1_projection_loss.detach(),
                grad_iou=g_iou, grad_l1=g_l1)


def main():
    if not ref_import.available():
        raise SystemExit("reference checkout not found")
    out = {}
    with ref_import.reference_modules() as ref:
        for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
            for case, shuffle, seed in (("ordered", False, 0), ("shuffled", True, 3)):
                res = run_projection(ref, dtype, shuffle, seed)
                for k, val in res.items():
                    out[f"proj_{case}_{tag}.{k}"] = val.numpy()
    make_distance_map = reference_make_distance_map()
    polygon = torch.tensor([[20.3, 10.2], [58.7, 14.9], [66.1, 40.4], [41.0, 55.5], [15.5, 38.0]])
    out["soft.polygon"] = polygon.numpy()
    out["soft.image_size"] = np.array([64, 80])
    out["soft.distance_map"] = make_distance_map(polygon, (64, 80)).numpy()
    path = os.path.join(HERE, "frame.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("proj_shuffled_f32")})
    print("gt_indices (shuffled f32):", out["proj_shuffled_f32.gt_indices"], "losses:",
          out["proj_shuffled_f32.iou_loss"], out["proj_shuffled_f32.l1_loss"])


if __name__ == "__main__":
    main()
