"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs the upstream checkout at /root/reference):

    python tests/golden/make_golden.py

What runs here is the reference's own code: `vsrd.rendering.hierarchical_volumetric_rendering`,
`vsrd.rendering.sdfs.*`, `HyperDistanceField`, `SinusoidalEncoder`, `BoxParameters3D`, and the field
closures compiled verbatim from `scripts/main.py:433-523` (see oracle/ref_import.py).  The only
intervention is that `torch.rand` / `torch.rand_like` are patched for the duration of the renderer
call so the draws are known and can be replayed into the oracle and the CUDA path.

Outputs (all small, committed):
    render_<case>.npz   inputs + every renderer output + loss + parameter gradients
    units.npz           ray_casting, BoxParameters3D decode, distance_field MLP, project_box_3d
"""
from __future__ import annotations

import functools
import math
import os
import sys
from unittest import mock

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402


class _AttrDict(dict):
    __getattr__ = dict.__getitem__


def synthetic_rays(ref, num_rays, gen, dtype, boxes_center_z=(8.0, 30.0)):
    """Pinhole camera at the origin (KITTI-360-like intrinsics scaled 1/4), random pixels."""
    h, w = 94, 352
    k = torch.tensor([[552.554 / 4, 0.0, 682.049 / 4], [0.0, 552.554 / 4, 238.770 / 4], [0.0, 0.0, 1.0]], dtype=dtype)
    e = torch.eye(4, dtype=dtype)
    cam, dirs = ref.rendering.ray_casting((h, w), k[None], e[None])
    # draw pixels from the band of the image the boxes project into, so most rays hit something
    rows = torch.randint(52, 84, (num_rays,), generator=gen)
    cols = torch.randint(40, 312, (num_rays,), generator=gen)
    picked = dirs[0, rows, cols]
    return cam.expand(num_rays, 3).contiguous(), picked.contiguous(), k, e


def make_scene(num_instances, gen, dtype):
    raw_loc = torch.zeros(num_instances, 3, dtype=dtype)
    # place boxes in front of the camera: invert the sigmoid-lerp of BoxParameters3D for the centres
    centres = torch.stack([
        torch.rand(num_instances, generator=gen, dtype=dtype) * 8.0 - 4.0,
        torch.full((num_instances,), 0.675, dtype=dtype),
        torch.rand(num_instances, generator=gen, dtype=dtype) * 14.0 + 8.0,
    ], dim=-1)
    lo = torch.tensor([-50.0, 1.55 - 1.75 / 2.0 - 5.0, 0.0], dtype=dtype)
    hi = torch.tensor([+50.0, 1.55 - 1.75 / 2.0 + 5.0, 100.0], dtype=dtype)
    raw_loc = torch.logit((centres - lo) / (hi - lo))
    raw_dim = torch.randn(num_instances, 3, generator=gen, dtype=dtype)
    yaw = torch.rand(num_instances, generator=gen, dtype=dtype) * 2 * math.pi - math.pi
    raw_ori = torch.stack([torch.cos(yaw), torch.sin(yaw)], dim=-1) * (0.5 + torch.rand(num_instances, 1, generator=gen, dtype=dtype))
    return raw_loc, raw_dim, raw_ori


def run_case(name, *, num_instances, num_rays, num_samples, residual, dtype, seed,
             temperature, std_deviation, cosine_ratio, weight_scale=1.0):
    # Build the scene in float64 so the f32 and f64 variants of a case describe the SAME scene
    # (the f32 case is then the reference's own rounding error relative to the f64 case).
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    distance_range = [0.0, 100.0]
    f64 = torch.float64

    with ref_import.reference_modules() as ref:
        detector = ref.box_parameters.BoxParameters3D(batch_size=1, num_instances=num_instances)
        hdf = ref.fields.HyperDistanceField(
            in_channels=48, out_channels_list=[16, 16, 16, 16],
            hyper_in_channels=256, hyper_out_channels_list=[256, 256, 256, 256])
        raw_loc, raw_dim, raw_ori = make_scene(num_instances, gen, f64)
        with torch.no_grad():
            detector.locations.copy_(raw_loc[None])
            detector.dimensions.copy_(raw_dim[None])
            detector.orientations.copy_(raw_ori[None])
            # distinct embeddings per instance so the residual fields differ
            detector.embeddings.copy_(torch.rand(1, num_instances, 256, generator=gen, dtype=f64))

        rays_o, rays_d, _, _ = synthetic_rays(ref, num_rays, gen, f64)
        targets = torch.rand(num_rays, num_instances, generator=gen, dtype=f64)
        jitter = torch.rand(num_rays, 1, num_samples, generator=gen, dtype=f64)
        uniforms = torch.rand(num_rays, 1, num_samples, generator=gen, dtype=f64)
        world = detector()
        w64 = (hdf(world["embeddings"])[0] * weight_scale).detach() if residual else None

        # ---- from here on everything runs in the case's dtype ----
        torch.set_default_dtype(dtype)
        rays_o, rays_d, targets, jitter, uniforms = (t.to(dtype) for t in (rays_o, rays_d, targets, jitter, uniforms))
        encoder = ref.encoders.SinusoidalEncoder(num_frequencies=8)
        assert encoder.frequencies.dtype == dtype
        models = _AttrDict(detector=detector, hyper_distance_field=hdf, positional_encoder=encoder)
        config = _AttrDict(volume_rendering=_AttrDict(distance_range=distance_range))
        closures = ref_import.main_closures(dict(torch=torch, nn=nn, config=config, models=models,
                                                 num_instances=num_instances))
        sdfs = ref.rendering.sdfs

        # decoded parameters as explicit leaves so their gradients can be read back
        loc = world["locations"][0].detach().to(dtype).requires_grad_(True)
        dim = world["dimensions"][0].detach().to(dtype).requires_grad_(True)
        rot = world["orientations"][0].detach().to(dtype).requires_grad_(True)
        w = w64.to(dtype).requires_grad_(True) if residual else None

        def build_field():
            fields = []
            for i in range(num_instances):
                box = sdfs.box(dim[i])
                if residual:
                    inner = closures["residual_composition"](
                        distance_field=box,
                        residual_distance_field=closures["residual_distance_field"](
                            distance_field=functools.partial(hdf.distance_field, w[i])))
                else:
                    inner = box
                inst = closures["instance_field"](distance_field=inner,
                                                  instance_label=dim.new_tensor(i, dtype=torch.long))
                fields.append(sdfs.translation(sdfs.rotation(inst, rot[i]), loc[i]))
            return closures["soft_union"](distance_fields=fields, temperature=temperature)

        field = build_field()
        renderer = ref.rendering.hierarchical_volumetric_rendering
        kwargs = dict(distance_field=field, ray_positions=rays_o, ray_directions=rays_d,
                      distance_range=distance_range, num_samples=num_samples,
                      sdf_std_deviation=std_deviation, cosine_ratio=cosine_ratio)

        # pass 1 (no grad): rand_like -> jitter
        with torch.no_grad(), mock.patch.object(torch, "rand_like", lambda t, *a, **k: jitter.to(t).reshape(t.shape)):
            *_, coarse_d, coarse_w = renderer(**kwargs)
        # pass 2: torch.rand -> uniforms (the reference sorts them itself)
        with mock.patch.object(torch, "rand", lambda *shape, **k: uniforms.reshape(shape)):
            labels, grads, fine_d, fine_w = renderer(**kwargs, sampled_distances=coarse_d, sampled_weights=coarse_w)

        sil = nn.functional.binary_cross_entropy(labels.clamp(1.0e-6, 1.0 - 1.0e-6), targets, reduction="none").mean()
        loss = sil
        out = dict(silhouette_loss=sil.detach())
        if residual:
            eik = nn.functional.mse_loss(torch.norm(grads, dim=-1), grads.new_ones(*grads.shape[:-1]), reduction="mean")
            loss = loss + 0.01 * eik
            out.update(eikonal_loss=eik.detach())
        leaves = [loc, dim, rot] + ([w] if residual else [])
        g = torch.autograd.grad(loss, leaves)
        out.update(grad_locations=g[0], grad_half_extents=g[1], grad_rotations=g[2])
        if residual:
            out.update(grad_mlp_weights=g[3])

    out.update(
        locations=loc, half_extents=dim, rotations=rot,
        ray_positions=rays_o, ray_directions=rays_d, targets=targets,
        jitter=jitter, sorted_uniforms=torch.sort(uniforms, dim=-1).values,
        coarse_distances=coarse_d, coarse_weights=coarse_w,
        labels=labels, sampled_gradients=grads, fine_distances=fine_d, fine_weights=fine_w,
        loss=loss,
        temperature=torch.tensor(temperature), std_deviation=torch.tensor(std_deviation),
        cosine_ratio=torch.tensor(cosine_ratio), num_samples=torch.tensor(num_samples),
        eikonal_weight=torch.tensor(0.01),
    )
    if residual:
        out.update(mlp_weights=w)
    arrays = {k: v.detach().cpu().numpy() for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), **arrays)
    print(f"render_{name}: loss={float(loss):.9g} labels.sum={float(labels.sum()):.6g} "
          f"hit_rays={int((coarse_w.sum(0) > 1e-3).sum())}/{num_rays}")
    torch.set_default_dtype(torch.float32)


def run_units():
    torch.set_default_dtype(torch.float32)
    gen = torch.Generator().manual_seed(7)
    out = {}
    with ref_import.reference_modules() as ref:
        # ray_casting on a tiny image with a non-trivial pose
        k = torch.tensor([[120.0, 0.0, 10.3], [0.0, 118.0, 6.1], [0.0, 0.0, 1.0]])
        ang = 0.3
        e = torch.eye(4)
        e[:3, :3] = torch.tensor([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
        e[:3, 3] = torch.tensor([0.4, -0.2, 1.5])
        cam, dirs = ref.rendering.ray_casting((12, 20), k[None], e[None])
        out.update(rc_intrinsic=k, rc_extrinsic=e, rc_camera_positions=cam, rc_ray_directions=dirs)

        # BoxParameters3D decode
        det = ref.box_parameters.BoxParameters3D(batch_size=1, num_instances=5)
        with torch.no_grad():
            det.locations.copy_(torch.randn(1, 5, 3, generator=gen))
            det.dimensions.copy_(torch.randn(1, 5, 3, generator=gen))
            det.orientations.copy_(torch.randn(1, 5, 2, generator=gen))
        world = det()
        out.update(bp_raw_locations=det.locations, bp_raw_dimensions=det.dimensions,
                   bp_raw_orientations=det.orientations, bp_locations=world["locations"],
                   bp_dimensions=world["dimensions"], bp_orientations=world["orientations"],
                   bp_boxes_3d=world["boxes_3d"])
        loc2, dim2, rot2 = det.encode_box_3d(world["boxes_3d"])
        out.update(bp_enc_locations=loc2, bp_enc_dimensions=dim2, bp_enc_orientations=rot2)

        # residual MLP + positional encoding
        hdf = ref.fields.HyperDistanceField(48, [16, 16, 16, 16], 256, [256, 256, 256, 256])
        enc = ref.encoders.SinusoidalEncoder(8)
        emb = torch.rand(2, 256, generator=gen)
        w = hdf(emb)
        x = torch.rand(7, 3, generator=gen) * 2 - 1
        pe = enc(x)
        out.update(mlp_embeddings=emb, mlp_weights=w, mlp_points=x, mlp_encoding=pe,
                   mlp_out0=hdf.distance_field(w[0], pe), mlp_out1=hdf.distance_field(w[1], pe))
        # state-dict contract of the hypernetwork (names + shapes only; the values are random init)
        out.update(hyper_state_keys=np.array(sorted(hdf.state_dict().keys())),
                   hyper_state_numel=torch.tensor([hdf.state_dict()[k].numel() for k in sorted(hdf.state_dict().keys())]))

        # project_box_3d: in front, straddling the image plane, behind
        line_indices = [[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]]
        boxes = world["boxes_3d"][0].detach().clone()
        boxes[0] += torch.tensor([0.0, 0.0, 60.0])
        boxes[1] += torch.tensor([3.0, 0.0, 51.0])
        boxes[2] += torch.tensor([0.0, 0.0, -500.0])
        boxes[3] += torch.tensor([-4.0, 1.0, 70.0])
        boxes[4] += torch.tensor([1.0, 0.0, 50.5])
        kk = torch.tensor([[552.554, 0.0, 682.049], [0.0, 552.554, 238.770], [0.0, 0.0, 1.0]])
        proj = torch.stack([ref.geometric_operations.project_box_3d(b, line_indices, kk) for b in boxes])
        out.update(pb_boxes_3d=boxes, pb_intrinsic=kk, pb_boxes_2d=proj,
                   pb_line_indices=torch.tensor(line_indices))
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "units.npz"), **arrays)
    print("units:", sorted(arrays))


if __name__ == "__main__":
    common = dict(num_instances=3, num_rays=48, num_samples=12)
    run_case("box_f32", residual=False, dtype=torch.float32, seed=1, temperature=0.7,
             std_deviation=0.6, cosine_ratio=0.25, **common)
    run_case("residual_f32", residual=True, dtype=torch.float32, seed=2, temperature=0.55,
             std_deviation=0.5, cosine_ratio=0.6, **common)
    run_case("residual_f64", residual=True, dtype=torch.float64, seed=2, temperature=0.55,
             std_deviation=0.5, cosine_ratio=0.6, **common)
    # late-schedule endpoint (T = sigma = 0.1, rho ~ 1) with amplified residual weights
    run_case("late_f64", residual=True, dtype=torch.float64, seed=3, temperature=0.1,
             std_deviation=0.1, cosine_ratio=0.97, weight_scale=3.0, **common)
    run_case("late_f32", residual=True, dtype=torch.float32, seed=3, temperature=0.1,
             std_deviation=0.1, cosine_ratio=0.97, weight_scale=3.0, **common)
    run_units()
