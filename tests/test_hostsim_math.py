"""The per-thread device math (vsrd_b200/csrc/vsrd_math.cuh), compiled for the host, against the
autograd oracle.  Runs without a GPU; the `-m gpu` tests repeat the same comparisons through the
CUDA kernels and the C-ABI."""
import shutil

import pytest
import torch

from oracle import vsrd_oracle as oracle
from tests.helpers import load_golden, rel_l2, render_kwargs, scene_from_golden

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def hs():
    from tests import hostsim
    hostsim.build()
    return hostsim


def _points(g, n=257, seed=0):
    """Sample positions near the boxes (inside, near faces, outside) and far away."""
    gen = torch.Generator().manual_seed(seed)
    loc = g["locations"].double()
    i = torch.randint(0, loc.shape[0], (n,), generator=gen)
    spread = torch.tensor([3.0, 2.0, 5.0], dtype=torch.float64)
    x = loc[i] + (torch.rand(n, 3, generator=gen, dtype=torch.float64) * 2 - 1) * spread
    x[: n // 8] += torch.randn(n // 8, 3, generator=gen, dtype=torch.float64) * 30.0
    return x.float()


@pytest.mark.parametrize("residual", [False, True])
def test_field_forward_matches_autograd(hs, residual):
    g = load_golden("residual_f32")
    x = _points(g)
    for i in range(g["locations"].shape[0]):
        w = g["mlp_weights"][i] if residual else None
        got = hs.field_forward(x, g["locations"][i], g["rotations"][i], g["half_extents"][i], w)
        xd = x.double().requires_grad_(True)
        d = oracle.instance_sdf(xd, g["locations"][i].double(), g["rotations"][i].double(),
                                g["half_extents"][i].double(), None if w is None else w.double(), 100.0)
        grad, = torch.autograd.grad(d.sum(), xd)
        assert (got[:, 0].double() - d.squeeze(-1)).abs().max() < 2e-5
        assert (got[:, 1:].double() - grad).abs().max() < 2e-4


@pytest.mark.parametrize("residual", [False, True])
def test_field_backward_matches_autograd(hs, residual):
    g = load_golden("residual_f32")
    x = _points(g, n=193, seed=1)
    gen = torch.Generator().manual_seed(5)
    adj = torch.randn(x.shape[0], 4, generator=gen)
    for i in range(g["locations"].shape[0]):
        w = g["mlp_weights"][i] if residual else None
        got = hs.field_backward(x, g["locations"][i], g["rotations"][i], g["half_extents"][i], w, adj)
        leaves = dict(locations=g["locations"][i].double().requires_grad_(True),
                      rotations=g["rotations"][i].double().requires_grad_(True),
                      half_extents=g["half_extents"][i].double().requires_grad_(True))
        if residual:
            leaves["mlp_weights"] = w.double().requires_grad_(True)
        xd = x.double().requires_grad_(True)
        d = oracle.instance_sdf(xd, leaves["locations"], leaves["rotations"], leaves["half_extents"],
                                leaves.get("mlp_weights"), 100.0)
        G, = torch.autograd.grad(d.sum(), xd, create_graph=True)
        phi = (adj[:, 0].double() * d.squeeze(-1)).sum() + (adj[:, 1:].double() * G).sum()
        want = torch.autograd.grad(phi, list(leaves.values()))
        for (name, _), ref in zip(leaves.items(), want):
            err = rel_l2(got[name], ref)
            assert err < 2e-4, f"instance {i} {name}: rel-L2 {err}"


def _fine_inputs(g):
    """Ray-major distances / per-instance fields for the golden's fine pass."""
    dist = g["fine_distances"].squeeze(-1).t().contiguous()             # [R, M+1]
    mid = (dist[:, :-1] + dist[:, 1:]) / 2.0
    pos = g["ray_positions"][:, None, :] + g["ray_directions"][:, None, :] * mid[..., None]  # [R,M,3]
    return dist, pos


@pytest.mark.parametrize("case", ["box_f32", "residual_f32", "late_f32"])
def test_forward_pipeline_matches_golden(hs, case):
    g = load_golden(case)
    dist, pos = _fine_inputs(g)
    R, M = pos.shape[:2]
    N = g["locations"].shape[0]
    field = torch.stack([
        hs.field_forward(pos.reshape(-1, 3), g["locations"][i], g["rotations"][i], g["half_extents"][i],
                         g["mlp_weights"][i] if "mlp_weights" in g else None).reshape(R, M, 4)
        for i in range(N)])
    kw = render_kwargs(g)
    labels, grads, weights = hs.composite_forward(dist, g["ray_directions"], field, float(g["temperature"]),
                                                  kw["sdf_std_deviation"], kw["cosine_ratio"])
    # north_star tolerance: rendered silhouettes within 1e-4 abs
    assert (labels - g["labels"]).abs().max() < 1e-4
    assert (weights.t()[..., None] - g["fine_weights"]).abs().max() < 1e-4
    # spatial gradients: exclude samples the miss-ray quirk throws to >1e3 m
    near = (g["fine_distances"][1:] < 1e3).expand_as(g["sampled_gradients"])
    diff = (grads.permute(1, 0, 2) - g["sampled_gradients"]).abs()
    assert diff[near].max() < 2e-3


@pytest.mark.parametrize("case", ["box_f32", "residual_f32"])
def test_backward_pipeline_matches_fp64_oracle(hs, case):
    """Param grads of (BCE + 0.01 eikonal) through composite-backward + field-backward vs the fp64
    oracle evaluated on the same sample distances (north_star tolerance: 1e-3 rel)."""
    g = load_golden(case)
    dist, pos = _fine_inputs(g)
    R, M = pos.shape[:2]
    N = g["locations"].shape[0]
    residual = "mlp_weights" in g
    # keep rays whose samples stay within the scene (the miss-ray extrapolation to 1e6 m makes the
    # fp32 reference itself disagree with fp64 by O(1); measured separately in the GPU tests)
    keep = dist.max(dim=1).values < 1e3
    dist, pos = dist[keep], pos[keep]
    dirs, origins, targets = g["ray_directions"][keep], g["ray_positions"][keep], g["targets"][keep]
    R = int(keep.sum())
    field = torch.stack([
        hs.field_forward(pos.reshape(-1, 3), g["locations"][i], g["rotations"][i], g["half_extents"][i],
                         g["mlp_weights"][i] if residual else None).reshape(R, M, 4)
        for i in range(N)])
    kw = render_kwargs(g)
    T = float(g["temperature"])
    labels, grads, _ = hs.composite_forward(dist, dirs, field, T, kw["sdf_std_deviation"], kw["cosine_ratio"])
    labels.requires_grad_(True)
    grads.requires_grad_(True)
    loss = oracle.silhouette_loss(labels, targets)
    if residual:
        loss = loss + 0.01 * oracle.eikonal_loss(grads)
    gl, gg = torch.autograd.grad(loss, [labels, grads], allow_unused=True)
    adj = hs.composite_backward(dist, dirs, field, T, kw["sdf_std_deviation"], kw["cosine_ratio"], gl=gl, gg=gg)
    got = dict(locations=[], rotations=[], half_extents=[], mlp_weights=[])
    for i in range(N):
        gi = hs.field_backward(pos.reshape(-1, 3), g["locations"][i], g["rotations"][i], g["half_extents"][i],
                               g["mlp_weights"][i] if residual else None, adj[i].reshape(-1, 4))
        for k in got:
            got[k].append(gi[k])
    got = {k: torch.stack(v) for k, v in got.items()}

    # oracle on identical distances, in fp64 (ground truth) and fp32 (the reference's own precision)
    names = ["locations", "half_extents", "rotations"] + (["mlp_weights"] if residual else [])
    want = {}
    for dt in (torch.float64, torch.float32):
        scene = scene_from_golden(g, dtype=dt, requires_grad=True)
        dist_sm = dist.to(dt).t()[..., None].contiguous()      # sample-major [M+1, R, 1]
        out = oracle.render_pass(scene.field(), origins.to(dt), dirs.to(dt), dist_sm,
                                 kw["sdf_std_deviation"], kw["cosine_ratio"])
        ref_loss = oracle.silhouette_loss(out[0], targets.to(dt))
        if residual:
            ref_loss = ref_loss + 0.01 * oracle.eikonal_loss(out[1])
        want[dt] = (ref_loss.detach(), torch.autograd.grad(ref_loss, [getattr(scene, n) for n in names]))
    assert abs(float(loss.detach()) - float(want[torch.float32][0])) < 1e-5
    for k, n in enumerate(names):
        ref64, ref32 = want[torch.float64][1][k], want[torch.float32][1][k].double()
        # north_star: parameter gradients within 1e-3 rel.  Against the fp32 reference that holds
        # outright; against fp64 the fp32 reference is itself 4e-3..7e-3 away on this scene (BCE on
        # labels near the 1e-6 clamp amplifies rounding), so there the bar is the reference's own error.
        assert rel_l2(got[n], ref32) < 1e-3, f"{n} vs fp32 oracle: {rel_l2(got[n], ref32)}"
        assert rel_l2(got[n], ref64) < max(1e-3, 1.5 * rel_l2(ref32, ref64)), f"{n} vs fp64 oracle"
