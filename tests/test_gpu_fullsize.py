"""CUDA-vs-oracle parity at the BASELINE shapes: cfg2 (N=8) and cfg3 (N=24) at R=1000 rays, S=100
samples (199 fine intervals), 17 views of 376x1408, at the annealing endpoints and mid-point, culling on and off.

This is the regime where the persistent-CTA tile split, instance-segment crossing, tile pairing and the culling
compaction of the field kernels are exercised.  Tolerances are BASELINE.json's: silhouettes (labels) and compositing
weights within 1e-4 abs; parameter gradients within 1e-3 rel of the fp64 oracle for a well-conditioned (linear)
upstream, and for the training loss (BCE at the clamp + eikonal, where the fp32 REFERENCE is itself 1e-3..1e-1 from
fp64) within max(1e-3, 1.5x the fp32 reference's own error), which is printed alongside.
"""
import pytest
import torch

from tests import fullsize_cases as fc
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def F():
    from vsrd_b200 import functional
    return functional


@pytest.fixture(scope="module")
def ops():
    from vsrd_b200 import ops
    return ops


@pytest.fixture(params=[True, False], ids=["cull", "nocull"])
def culling(request, ops):
    default = ops.CULL_MAX_TEMPERATURE
    ops.set_culling(request.param, max_temperature=float("inf"))      # cull at every temperature of the schedule
    yield request.param
    ops.set_culling(True, max_temperature=default)


def _leaves(case, requires_grad):
    return [case[k].to(DEV).clone().requires_grad_(requires_grad) for k in fc.GRAD_NAMES]


def _render(F, case, leaves, rays, distances):
    s = case["schedule"]
    return F.render_pass(*leaves, case["origins"][rays].to(DEV), case["directions"][rays].to(DEV), distances.to(DEV),
                         temperature=s["temperature"], std_deviation=s["std_deviation"], cosine_ratio=s["cosine_ratio"])


@pytest.mark.parametrize("cfg,sched", fc.CASES)
def test_fine_pass_matches_oracle_at_full_size(F, culling, cfg, sched):
    case = fc.get_case(cfg, sched)
    keep = case["keep"]
    assert keep.numel() == fc.NUM_RAYS and case["fine_all"].shape[1] == 2 * fc.NUM_SAMPLES
    fine = case["fine_all"][keep]
    # grad mode on so that the culled forward path is the one under test (no-grad calls are the un-culled coarse pass)
    labels, grads, weights = _render(F, case, _leaves(case, True), keep, fine)
    labels, grads, weights = labels.detach().cpu(), grads.detach().cpu(), weights.detach().cpu()
    e32 = float((labels - case["f32_labels"]).abs().max())
    e64 = float((labels.double() - case["f64_labels"]).abs().max())
    ref = float((case["f32_labels"].double() - case["f64_labels"]).abs().max())
    w32 = float((weights - case["f32_weights"].squeeze(-1).t()).abs().max())
    w64 = float((weights.double() - case["f64_weights"].squeeze(-1).t()).abs().max())
    wref = float((case["f32_weights"].double() - case["f64_weights"]).abs().max())
    g64 = float((grads.double().permute(1, 0, 2) - case["f64_gradients"]).abs().max())
    gref = float((case["f32_gradients"].double() - case["f64_gradients"]).abs().max())
    print(f"{cfg}/{sched} cull={culling}: labels vs fp32 {e32:.2e} vs fp64 {e64:.2e} (fp32 reference's own error {ref:.2e}); "
          f"weights vs fp32 {w32:.2e} vs fp64 {w64:.2e} (reference {wref:.2e}); union gradient vs fp64 {g64:.2e} (reference {gref:.2e})")
    # north_star: within 1e-4 abs of the reference renderer (fp32); against fp64 the bound is the reference's own error
    assert e32 < 1e-4 and e64 < max(1e-4, 1.5 * ref)
    assert w32 < 1e-4 and w64 < max(1e-4, 1.5 * wref)
    assert g64 < max(1e-3, 2.0 * gref)
    if case["miss"].numel():          # rays whose importance samples extrapolated to 1e3..1e6 m: labels only
        miss = case["miss"]
        with torch.no_grad():
            lm, _, _ = _render(F, case, _leaves(case, False), miss, case["fine_all"][miss])
        assert (lm.cpu() - case["miss_labels"]).abs().max() < 1e-4


@pytest.mark.parametrize("cfg,sched", fc.CASES)
def test_parameter_gradients_match_fp64_oracle_at_full_size(F, culling, cfg, sched):
    case = fc.get_case(cfg, sched)
    keep = case["keep"]
    fine = case["fine_all"][keep]
    targets = case["targets"][keep].to(DEV)
    r, m, n = fc.NUM_RAYS, fine.shape[1] - 1, case["locations"].shape[0]

    # (a) linear upstream: the kernels' adjoint alone, 1e-3 rel vs fp64 (or the fp32 reference's own error where that
    #     is larger: a ray grazing the relu / 1/(cdf + eps) kinks of renderers.py:248 is ill-conditioned in fp32)
    leaves = _leaves(case, True)
    labels, grads, weights = _render(F, case, leaves, keep, fine)
    lin = fc.linear_loss(labels, grads.permute(1, 0, 2), weights.t()[..., None], fc.linear_coefficients(r, m, n))
    got = torch.autograd.grad(lin, leaves)
    assert abs(float(lin) - float(case["f64_lin_loss"])) < 1e-3 * max(1.0, abs(float(case["f64_lin_loss"])))
    for name, g in zip(fc.GRAD_NAMES, got):
        err = rel_l2(g.cpu().double(), case[f"f64_lin_grad_{name}"])
        e_ref = rel_l2(case[f"f32_lin_grad_{name}"].double(), case[f"f64_lin_grad_{name}"])
        e_32 = rel_l2(g.cpu().double(), case[f"f32_lin_grad_{name}"].double())
        print(f"{cfg}/{sched} cull={culling} linear  {name:12s} rel-L2 vs fp64 {err:.2e} (fp32 reference's own error {e_ref:.2e}), vs fp32 {e_32:.2e}")
        assert err < max(1e-3, 1.5 * e_ref), f"{name}: {err} (reference {e_ref})"

    # (b) the loss main.py optimises, unfused (autograd through labels / gradients) and fused
    leaves = _leaves(case, True)
    labels, grads, _ = _render(F, case, leaves, keep, fine)
    loss = fc.training_loss(labels, grads, targets)
    got = torch.autograd.grad(loss, leaves)
    s = case["schedule"]
    leaves_f = _leaves(case, True)
    loss_f, _, _ = F.fused_render_loss(*leaves_f, case["origins"][keep].to(DEV), case["directions"][keep].to(DEV),
                                       fine.to(DEV), targets, eikonal_weight=0.01, **s)
    got_f = torch.autograd.grad(loss_f, leaves_f)
    assert abs(float(loss) - float(case["f32_loss"])) < 1e-5 * max(1.0, abs(float(case["f32_loss"])))
    assert abs(float(loss_f) - float(loss)) < 2e-6 * max(1.0, abs(float(loss)))
    for name, g, gf in zip(fc.GRAD_NAMES, got, got_f):
        want64, want32 = case[f"f64_grad_{name}"], case[f"f32_grad_{name}"].double()
        e_ref = rel_l2(want32, want64)
        e_64 = rel_l2(g.cpu().double(), want64)
        e_f = rel_l2(gf.cpu().double(), g.cpu().double())
        print(f"{cfg}/{sched} cull={culling} training {name:12s} rel-L2 vs fp64 {e_64:.2e} (fp32 reference's own error {e_ref:.2e}), fused vs unfused {e_f:.1e}")
        assert e_64 < max(1e-3, 1.5 * e_ref), f"{name}: {e_64} (reference {e_ref})"
        assert e_f < 1e-5, f"{name}: fused vs unfused {e_f}"


@pytest.mark.parametrize("cfg,sched", [("cfg2", "mid"), ("cfg3", "late")])
def test_two_pass_placement_at_full_size(F, ops, cfg, sched):
    """Own coarse pass + importance resampling with the oracle's draws injected: stratified placement bit-exact,
    coarse weights within 1e-4, and the fine samples equal to the oracle's up to rounding except where a 1-ulp CDF
    difference flips a searchsorted bin (SURVEY §7 hard part 3)."""
    case = fc.get_case(cfg, sched)
    s = case["schedule"]
    with torch.no_grad():
        _, _, cd, cw, fd, _ = F.two_pass_render(
            *_leaves(case, False), case["origins"].to(DEV), case["directions"].to(DEV), num_samples=fc.NUM_SAMPLES,
            jitter=case["jitter"].squeeze(1).to(DEV), sorted_uniforms=case["sorted_uniforms"].squeeze(1).to(DEV), **s)
    assert torch.equal(cd.cpu(), case["coarse_distances"])
    assert (cw.cpu() - case["coarse_weights"]).abs().max() < 1e-4
    ref = case["fine_all"]
    close = (fd.cpu() - ref).abs() <= 1e-4 * (1.0 + ref.abs())
    print(f"{cfg}/{sched}: fine samples within 1e-4 rel of the oracle's: {float(close.float().mean()):.4f}")
    assert close.float().mean() > 0.97
    assert bool((fd[:, 1:] >= fd[:, :-1]).all())
