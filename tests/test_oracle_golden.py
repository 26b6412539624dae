"""Pin the CPU oracle (oracle/vsrd_oracle.py) to the fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import vsrd_oracle as oracle
from tests.helpers import GOLDEN_DIR, RENDER_CASES, load_golden, render_kwargs, scene_from_golden


@pytest.mark.parametrize("case", RENDER_CASES)
def test_render_matches_reference(case):
    g = load_golden(case)
    dtype = g["labels"].dtype
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        scene = scene_from_golden(g, requires_grad=True)
        loss, parts = oracle.render_loss(
            scene, g["ray_positions"], g["ray_directions"], g["targets"],
            jitter=g["jitter"], sorted_uniforms=g["sorted_uniforms"],
            eikonal_weight=float(g["eikonal_weight"]), **render_kwargs(g))
        leaves = [scene.locations, scene.half_extents, scene.rotations]
        names = ["grad_locations", "grad_half_extents", "grad_rotations"]
        if scene.mlp_weights is not None:
            leaves.append(scene.mlp_weights)
            names.append("grad_mlp_weights")
        grads = torch.autograd.grad(loss, leaves)
    finally:
        torch.set_default_dtype(prev)
    # same ops, same order, same machine => exact (allow a few ulp for thread-count dependent reductions)
    tol = 1e-6 if dtype == torch.float32 else 1e-12
    for key in ["coarse_distances", "coarse_weights", "fine_distances", "fine_weights", "labels", "sampled_gradients"]:
        torch.testing.assert_close(parts[key], g[key], rtol=tol, atol=tol, msg=lambda m, k=key: f"{k}: {m}")
    torch.testing.assert_close(loss.detach(), g["loss"], rtol=tol, atol=tol)
    # the late-schedule f32 case is ill-conditioned w.r.t. its own rounding (op order alone moves its
    # gradients by 1e-3..2e-2, SURVEY.md §7 hard part 4); the f64 cases pin the gradient formulas.
    gtol = (5e-2 if case.startswith("late") else 1e-5) if dtype == torch.float32 else 1e-8
    for name, got in zip(names, grads):
        err = float((got - g[name]).norm() / g[name].norm().clamp_min(1e-30))
        assert err < gtol, f"{name}: rel-L2 {err}"


def test_units_match_reference():
    u = np.load(f"{GOLDEN_DIR}/units.npz")
    t = lambda k: torch.from_numpy(u[k])
    cam, dirs = oracle.ray_casting((12, 20), t("rc_intrinsic")[None], t("rc_extrinsic")[None])
    torch.testing.assert_close(cam, t("rc_camera_positions"), rtol=0, atol=0)
    torch.testing.assert_close(dirs, t("rc_ray_directions"), rtol=0, atol=1e-7)

    loc, dim, rot = oracle.decode_box_parameters(t("bp_raw_locations"), t("bp_raw_dimensions"), t("bp_raw_orientations"))
    torch.testing.assert_close(loc, t("bp_locations"), rtol=0, atol=0)
    torch.testing.assert_close(dim, t("bp_dimensions"), rtol=0, atol=0)
    torch.testing.assert_close(rot, t("bp_orientations"), rtol=0, atol=0)
    torch.testing.assert_close(oracle.box_corners(loc, dim, rot), t("bp_boxes_3d"), rtol=0, atol=1e-6)

    pe = oracle.sinusoidal_encoding(t("mlp_points"))
    torch.testing.assert_close(pe, t("mlp_encoding"), rtol=0, atol=0)
    w = t("mlp_weights")
    torch.testing.assert_close(oracle.residual_mlp(w[0], pe), t("mlp_out0"), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(oracle.residual_mlp(w[1], pe), t("mlp_out1"), rtol=1e-6, atol=1e-6)

    net = oracle.HyperNetwork()
    keys = sorted(net.state_dict().keys())
    assert keys == list(u["hyper_state_keys"])
    assert [net.state_dict()[k].numel() for k in keys] == list(u["hyper_state_numel"])


@pytest.mark.parametrize("case", ["box_f32", "residual_f32", "late_f32"])
def test_surface_renderers_match_reference(case):
    """oracle/surface_oracle.py against the reference's own sphere_tracing / surface_normal outputs."""
    from oracle import surface_oracle as so
    from tests.helpers import load_surface_golden
    g = load_surface_golden(case)
    scene = scene_from_golden(load_golden(case))
    it, crit = int(g["num_iterations"]), float(g["criteria"])
    d, labels = scene.field()(g["points"])
    torch.testing.assert_close(d, g["point_distances"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(labels, g["point_labels"], rtol=1e-6, atol=1e-6)
    runs = [
        ("positions", "converged", dict(ray_positions=g["camera_position"], num_iterations=it, bounding_radius=100.0, initialization=False)),
        ("positions_cap", "converged_cap", dict(ray_positions=g["camera_position"], num_iterations=7, bounding_radius=100.0, initialization=False)),
        ("positions_init", "converged_init", dict(ray_positions=g["far_camera"], num_iterations=it, bounding_radius=40.0, initialization=True)),
        ("positions_newton", "converged_newton", dict(ray_positions=g["camera_position"], num_iterations=it, bounding_radius=100.0,
                                                      initialization=False, differentiable=True)),
    ]
    for pk, ck, kw in runs:
        pos, conv = so.sphere_tracing(scene, ray_directions=g["ray_directions"], convergence_criteria=crit, **kw)
        assert torch.equal(conv, g[ck]), (pk, int((conv != g[ck]).sum()))
        torch.testing.assert_close(pos.detach(), g[pk], rtol=1e-6, atol=1e-5, msg=lambda m, k=pk: f"{k}: {m}")
    torch.testing.assert_close(so.surface_normal(scene, g["positions"]), g["normals"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(so.surface_normal(scene, g["positions"], 1e-2), g["normals_fd"], rtol=1e-4, atol=1e-4)
