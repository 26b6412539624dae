"""Golden fixtures of the inference / logging renderers, produced by the UNMODIFIED reference
(`vsrd.rendering.sphere_tracing`, `vsrd.rendering.surface_normal`, rendering/renderers.py:21-113) on the
union field compiled verbatim from scripts/main.py:433-492.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_surface.py

The scenes are those of the committed render_*.npz cases (their decoded parameters and residual-MLP weights
are read back from the fixtures), looked at through a 20 x 72 pixel window of the 94 x 352 pinhole camera.
Output: surface.npz (small, committed).
"""
from __future__ import annotations

import functools
import operator
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402


class _AttrDict(dict):
    __getattr__ = dict.__getitem__


def build_field(ref, closures, hdf, loc, dim, rot, w, temperature):
    sdfs = ref.rendering.sdfs
    fields = []
    for i in range(loc.shape[0]):
        box = sdfs.box(dim[i])
        if w is not None:
            inner = closures["residual_composition"](
                distance_field=box,
                residual_distance_field=closures["residual_distance_field"](
                    distance_field=functools.partial(hdf.distance_field, w[i])))
        else:
            inner = box
        inst = closures["instance_field"](distance_field=inner, instance_label=torch.tensor(i, dtype=torch.long))
        fields.append(sdfs.translation(sdfs.rotation(inst, rot[i]), loc[i]))
    return closures["soft_union"](distance_fields=fields, temperature=temperature)


def run_case(ref, case, out, num_iterations, criteria):
    z = np.load(os.path.join(HERE, f"render_{case}.npz"))
    loc, dim, rot = (torch.from_numpy(z[k]) for k in ("locations", "half_extents", "rotations"))
    w = torch.from_numpy(z["mlp_weights"]) if "mlp_weights" in z.files else None
    temperature = float(z["temperature"])
    n = loc.shape[0]
    hdf = ref.fields.HyperDistanceField(in_channels=48, out_channels_list=[16, 16, 16, 16],
                                        hyper_in_channels=256, hyper_out_channels_list=[256, 256, 256, 256])
    encoder = ref.encoders.SinusoidalEncoder(num_frequencies=8)
    models = _AttrDict(hyper_distance_field=hdf, positional_encoder=encoder)
    config = _AttrDict(volume_rendering=_AttrDict(distance_range=[0.0, 100.0]))
    closures = ref_import.main_closures(dict(torch=torch, nn=nn, config=config, models=models, num_instances=n))
    field = build_field(ref, closures, hdf, loc, dim, rot, w, temperature)
    distance = ref.utils.compose(field, operator.itemgetter(0))          # main.py:1030

    k = torch.tensor([[552.554 / 4, 0.0, 682.049 / 4], [0.0, 552.554 / 4, 238.770 / 4], [0.0, 0.0, 1.0]])
    cam, dirs = ref.rendering.ray_casting((94, 352), k[None], torch.eye(4)[None])
    window = dirs[0, 56:76, 140:212].contiguous()                          # [20,72,3]: looks at the boxes
    camera_position = cam[0]

    with torch.no_grad():
        # exactly the call of main.py:1028-1040
        positions, converged = ref.rendering.sphere_tracing(
            distance_field=distance, ray_positions=camera_position, ray_directions=window,
            num_iterations=num_iterations, convergence_criteria=criteria, bounding_radius=100.0,
            initialization=False, differentiable=False)
        # a short run that ends on the iteration cap, not on the global early exit
        positions_cap, converged_cap = ref.rendering.sphere_tracing(
            distance_field=distance, ray_positions=camera_position, ray_directions=window,
            num_iterations=7, convergence_criteria=criteria, bounding_radius=100.0,
            initialization=False, differentiable=False)
        # bounding-sphere initialisation from a camera outside a 40 m sphere
        far_camera = torch.tensor([0.0, 0.0, -45.0])
        positions_init, converged_init = ref.rendering.sphere_tracing(
            distance_field=distance, ray_positions=far_camera, ray_directions=window,
            num_iterations=num_iterations, convergence_criteria=criteria, bounding_radius=40.0,
            initialization=True, differentiable=False)
    positions_newton, converged_newton = ref.rendering.sphere_tracing(
        distance_field=distance, ray_positions=camera_position, ray_directions=window,
        num_iterations=num_iterations, convergence_criteria=criteria, bounding_radius=100.0,
        initialization=False, differentiable=True)
    normals = ref.rendering.surface_normal(distance, positions.clone())
    normals_fd = ref.rendering.surface_normal(distance, positions.clone(), finite_difference_epsilon=1e-2)
    gen = torch.Generator().manual_seed(11)
    points = loc[torch.randint(0, n, (256,), generator=gen)] + torch.randn(256, 3, generator=gen) * 2.0
    with torch.no_grad():
        point_distances, point_labels = field(points)

    out.update({f"{case}.{k}": v.detach().numpy() for k, v in dict(
        camera_position=camera_position, far_camera=far_camera, ray_directions=window,
        positions=positions, converged=converged, positions_cap=positions_cap, converged_cap=converged_cap,
        positions_init=positions_init, converged_init=converged_init,
        positions_newton=positions_newton, converged_newton=converged_newton,
        normals=normals, normals_fd=normals_fd, points=points, point_distances=point_distances,
        point_labels=point_labels, num_iterations=torch.tensor(num_iterations), criteria=torch.tensor(criteria),
    ).items()})
    print(f"{case}: converged {int(converged.sum())}/{converged.numel()} (cap: {int(converged_cap.sum())}, "
          f"init: {int(converged_init.sum())}), |normal|={float(normals.norm(dim=-1).mean()):.4f}")


def main():
    torch.set_default_dtype(torch.float32)
    torch.manual_seed(0)
    out = {}
    with ref_import.reference_modules() as ref:
        for case in ("box_f32", "residual_f32", "late_f32"):
            run_case(ref, case, out, num_iterations=200, criteria=0.01)
    np.savez_compressed(os.path.join(HERE, "surface.npz"), **out)
    print("wrote surface.npz", os.path.getsize(os.path.join(HERE, "surface.npz")), "bytes")


if __name__ == "__main__":
    main()
