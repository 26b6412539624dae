"""GPU parity tests of the frame-level kernels (soft masks, schedule)
through the C ABI, against the reference-generated goldens (tests/golden/frame.npz) and the CPU oracle
(oracle/frame_oracle.py).  Index outputs are compared exactly, floating point within the stated tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import frame_oracle as fo
from tests.helpers import GOLDEN_DIR, rel_l2

pytestmark = pytest.mark.gpu


def load_case(name):
    data = np.load(os.path.join(GOLDEN_DIR, "frame.npz"))
    pre = f"proj_{name}."
    return {k[len(pre):]: torch.from_numpy(data[k]) for k in data.files if k.startswith(pre)}


@pytest.fixture(scope="module")
def ops():
    from vsrd_b200 import _lib, ops as _ops
    _lib.load()
    return _ops


# ---- a14 / a15 -----------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["ordered_f32", "shuffled_f32"])
def test_projection_step_matches_reference(ops, case):
    g, g64 = load_case(case), load_case(case.replace("f32", "f64"))
    dev = "cuda"
    size = tuple(int(x) for x in g["image_size"])
    views = ops.ViewArgs(g["extrinsics"].to(dev), g["intrinsics"].to(dev), size, int(g["target_view"]))
    boxes, gt_idx, losses, grad = ops.projection_step(views, g["world_boxes"].to(dev), g["gt_boxes_2d"].to(dev),
                                                      g["visible"].to(dev))
    assert torch.allclose(boxes.cpu(), g["boxes_2d"], rtol=1e-5, atol=1e-3)
    assert torch.equal(gt_idx.cpu(), g["gt_indices"])                       # Hungarian assignment: exact
    assert abs(float(losses[0]) - float(g["iou_loss"])) < 1e-5 * max(1.0, float(g["iou_loss"]))
    assert abs(float(losses[1]) - float(g["l1_loss"])) < 1e-5 * max(1.0, float(g["l1_loss"]))
    assert rel_l2(grad[0].cpu().double(), g64["grad_iou"]) < 1e-3            # fp32 kernel vs fp64 reference autograd
    assert rel_l2(grad[1].cpu().double(), g64["grad_l1"]) < 1e-3
    # injected assignment (the reference re-uses the matching between logging steps) gives the same result
    boxes2, gt2, losses2, grad2 = ops.projection_step(views, g["world_boxes"].to(dev), g["gt_boxes_2d"].to(dev),
                                                      g["visible"].to(dev), fixed_gt_indices=g["gt_indices"].to(dev))
    assert torch.equal(boxes2, boxes) and torch.equal(gt2, gt_idx)
    assert torch.equal(losses2, losses) and torch.equal(grad2, grad)
    # projection only
    only = ops.projection_step(views, g["world_boxes"].to(dev))
    assert torch.equal(only, boxes)


def test_projection_assignment_random_costs_match_scipy(ops):
    """The in-kernel assignment against scipy.optimize.linear_sum_assignment on random box sets (N = 1..32)."""
    from scipy.optimize import linear_sum_assignment
    gen = torch.Generator().manual_seed(0)
    k = torch.tensor([[500.0, 0, 320], [0, 500.0, 120], [0, 0, 1]])
    e = torch.eye(4)
    for n in [1, 2, 3, 5, 8, 13, 24, 32]:
        centre = torch.stack([torch.rand(n, generator=gen) * 16 - 8, torch.full((n,), 0.7),
                              torch.rand(n, generator=gen) * 30 + 6], -1)
        half = torch.rand(n, 3, generator=gen) * 0.8 + 0.7
        signs = torch.tensor([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=torch.float32)
        world = centre[:, None] + signs[None] * half[:, None]
        size = (240, 640)
        pd = fo.multi_view_boxes_2d(world, e[None], k[None], size)
        perm = torch.randperm(n, generator=gen)
        gt = (pd[:, perm] + torch.randn(1, n, 4, generator=gen) * 3.0)
        cost = -fo.distance_box_iou(pd[0], gt[0])
        expect = linear_sum_assignment(cost.numpy())[1]
        views = ops.ViewArgs(e[None].cuda(), k[None].cuda(), size, 0)
        _, gt_idx, losses, _ = ops.projection_step(views, world.cuda(), gt.cuda(), None)
        got = gt_idx.cpu().numpy()
        # equal assignment, or (ties) equal optimal cost
        if not np.array_equal(got, expect):
            assert sorted(got.tolist()) == list(range(n))
            c = cost.double().numpy()
            assert abs(c[np.arange(n), got].sum() - c[np.arange(n), expect].sum()) < 1e-5
        assert torch.isfinite(losses).all()


# ---- a2 ------------------------------------------------------------------------------------------

def _masks(p, n, seed, zero_fraction=0.6):
    gen = torch.Generator().manual_seed(seed)
    m = torch.rand(p, n, generator=gen) ** 4
    m[torch.rand(p, generator=gen) < zero_fraction] = 0.0
    return m


@pytest.mark.parametrize("p,n", [(1, 1), (4095, 3), (4096, 8), (4097, 8), (376 * 1408 + 13, 5)])
def test_ray_cdf_matches_numpy_cumsum(ops, p, n):
    m = _masks(p, n, seed=p)
    cdf = ops.ray_cdf_build(m.cuda()).cpu().numpy()
    ref = np.cumsum(m.max(dim=-1).values.double().numpy())
    assert cdf.shape == ref.shape
    assert np.all(np.diff(cdf) >= 0)
    assert np.allclose(cdf, ref, rtol=1e-12, atol=1e-12)


def test_select_rays_with_injected_uniforms_is_exact(ops):
    p, n, r = 20000, 4, 1000
    m = _masks(p, n, seed=7)
    cdf = ops.ray_cdf_build(m.cuda())
    u = torch.rand(4096, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    got, status = ops.select_rays(cdf, r, uniforms=u.cuda())
    assert int(status) == 0
    weights = m.max(dim=-1).values
    ref = fo.select_rays(weights.numpy(), r, u.numpy())
    got = got.cpu()
    assert len(set(got.tolist())) == r and bool((weights[got] > 0).all())
    # the oracle accumulates its CDF in a different order: a draw landing within an ulp of a CDF edge
    # may legitimately resolve to the neighbouring pixel; everything else must be identical, in order
    assert (got == ref).float().mean() > 0.999
    targets = ops.gather_targets(m.cuda(), got.cuda())
    assert torch.equal(targets.cpu(), fo.gather_targets(m, got))
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(2))
    targets = ops.gather_targets(m.cuda(), got.cuda(), perm.cuda())
    assert torch.equal(targets.cpu(), fo.gather_targets(m, got, perm))


def test_select_rays_generator_draws_follow_the_weights(ops):
    p, r = 64, 16
    w = torch.linspace(0.0, 1.0, p) ** 2
    cdf = ops.ray_cdf_build(w[:, None].cuda())
    counts = torch.zeros(p)
    trials = 400
    for seed in range(trials):
        idx, status = ops.select_rays(cdf, r, seed=seed)
        idx = idx.cpu()
        assert int(status) == 0 and len(set(idx.tolist())) == r
        counts[idx] += 1
    assert counts[0] == 0                                      # zero-weight pixel is never drawn
    # inclusion frequencies against torch.multinomial without replacement (same sampling design)
    ref = torch.zeros(p)
    gen = torch.Generator().manual_seed(0)
    for _ in range(trials):
        ref[torch.multinomial(w, r, replacement=False, generator=gen)] += 1
    assert float((counts - ref).abs().max()) / trials < 0.12
    assert abs(float(counts[p // 2:].sum() - ref[p // 2:].sum())) / (trials * r) < 0.03
    # different seeds -> different draws; same seed -> same draw
    a, _ = ops.select_rays(cdf, r, seed=5)
    b, _ = ops.select_rays(cdf, r, seed=5)
    c, _ = ops.select_rays(cdf, r, seed=6)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_select_rays_reports_exhaustion(ops):
    w = torch.zeros(100)
    w[[3, 50, 77]] = 1.0
    cdf = ops.ray_cdf_build(w[:, None].cuda())
    idx, status = ops.select_rays(cdf, 8, seed=0)              # torch.multinomial raises here
    assert int(status) == 5
    assert sorted(idx.cpu().tolist()[:3]) == [3, 50, 77] and idx.cpu().tolist()[3:] == [-1] * 5
    cdf0 = ops.ray_cdf_build(torch.zeros(100, 1).cuda())
    idx, status = ops.select_rays(cdf0, 4, seed=0)
    assert int(status) == 4 and idx.cpu().tolist() == [-1] * 4


# ---- soft masks ----------------------------------------------------------------------------------

def test_soft_masks_match_oracle(ops):
    size = (48, 72)
    polys = torch.zeros(2, 3, 6, 2)
    sizes = torch.zeros(2, 3, dtype=torch.int32)
    polys[0, 0, :4] = torch.tensor([[10.3, 8.2], [40.7, 6.1], [45.2, 30.9], [12.8, 35.5]]); sizes[0, 0] = 4
    polys[0, 1, :5] = torch.tensor([[50.0, 10.0], [70.5, 12.0], [80.0, 30.0], [60.0, 44.0], [48.0, 30.0]]); sizes[0, 1] = 5
    polys[1, 2, :3] = torch.tensor([[-5.0, 20.0], [30.0, -4.0], [33.0, 52.0]]); sizes[1, 2] = 3
    polys[1, 0, :6] = torch.tensor([[20.0, 20.0], [30.0, 15.0], [40.0, 20.0], [40.0, 30.0], [30.0, 36.0], [20.0, 30.0]]); sizes[1, 0] = 6
    out = ops.soft_masks(polys.cuda(), sizes.cuda(), size, temperature=10.0).cpu()
    assert out.shape == (2, 48, 72, 3)
    for v in range(2):
        for n in range(3):
            c = int(sizes[v, n])
            if c < 3:
                assert float(out[v, ..., n].abs().max()) == 0.0       # absent instance: empty mask
                continue
            ref = fo.soft_mask(polys[v, n, :c], size, temperature=10.0)
            assert torch.allclose(out[v, ..., n], ref, atol=1e-5), (v, n, float((out[v, ..., n] - ref).abs().max()))


# ---- schedule ------------------------------------------------------------------------------------

def test_step_state_follows_the_schedule(ops):
    st = ops.StepState(num_steps=3000, warmup_steps=1000, seed=11)
    seeds = set()
    for step in [0, 1, 999, 1000, 1500, 2999]:
        st.set_step(step)
        got, ref = st.read(), fo.schedule(step, 3000, 1000)
        assert got["step"] == step
        for k in ("temperature", "std_deviation", "cosine_ratio", "eikonal_weight"):
            assert abs(got[k] - ref[k]) <= 1e-7 * max(1.0, abs(ref[k])), (step, k)
        seeds.add(got["seed"])
    assert len(seeds) == 6
    st.set_step(41)
    st.advance()
    assert st.read()["step"] == 42


def test_step_state_drives_the_kernels(ops):
    """A non-NULL step_state overrides the by-value scalars: same outputs as passing them by value."""
    from vsrd_b200 import functional as F
    from tests.helpers import load_golden
    g = load_golden("residual_f32")
    dev = "cuda"
    st = ops.StepState(num_steps=100, warmup_steps=0, seed=3)
    st.set_step(37)
    s = st.read()
    scene_v = ops.SceneArgs(g["locations"].to(dev), g["rotations"].to(dev), g["half_extents"].to(dev),
                            g["mlp_weights"].to(dev), s["temperature"])
    scene_s = ops.SceneArgs(g["locations"].to(dev), g["rotations"].to(dev), g["half_extents"].to(dev),
                            g["mlp_weights"].to(dev), 123.0, step_state=st)
    r, ns = g["ray_directions"].shape[0], int(g["num_samples"])
    bins = F.distance_bins([0.0, 100.0], ns, dev)
    a = ops.place_coarse(bins, r, None, seed=s["seed"])
    b = ops.place_coarse(bins, r, None, seed=999, step_state=st)
    assert torch.equal(a, b)
    rays = ops.RayArgs(g["ray_positions"].to(dev), g["ray_directions"].to(dev), a)
    fa, fb = ops.field_forward(scene_v, rays), ops.field_forward(scene_s, rays)
    la, ga, wa, _ = ops.composite_forward(scene_v, rays, fa, s["std_deviation"], s["cosine_ratio"])
    lb, gb, wb, _ = ops.composite_forward(scene_s, rays, fb, 55.0, 0.123)
    assert torch.equal(la, lb) and torch.equal(ga, gb) and torch.equal(wa, wb)


def test_vsrd_losses_projection_losses_match_the_frame_oracle():
    """`vsrd.losses.projection_losses` (the public face of projection_step_kernel) against the restatement of
    main.py:339-415, values and gradient w.r.t. the world boxes."""
    import vsrd
    from vsrd_b200 import synthetic
    frame = synthetic.make_frame(num_instances=5, num_views=5, seed=12)
    sup = synthetic.frame_supervision(frame)
    raw = synthetic.perturbed_raw_parameters(frame, seed=12)
    loc, dim, rot = oracle_decode(raw)
    from oracle import vsrd_oracle as vo
    boxes = vo.box_corners(loc, dim, rot)
    want_boxes = boxes.clone().double().requires_grad_(True)
    _, gt_idx, iou, l1 = fo.projection_step(want_boxes, frame.extrinsics.double(), frame.intrinsics.double(), frame.image_size,
                                            sup.boxes_2d.double(), sup.visible, sup.target_view)
    want_grad, = torch.autograd.grad(0.1 * iou + l1, want_boxes)
    got_boxes = boxes.cuda().requires_grad_(True)
    g_iou, g_l1, g_idx = vsrd.losses.projection_losses(got_boxes, frame.extrinsics.cuda(), frame.intrinsics.cuda(), frame.image_size,
                                                      sup.boxes_2d.cuda(), sup.visible.cuda(), sup.target_view)
    got_grad, = torch.autograd.grad(0.1 * g_iou + g_l1, got_boxes)
    assert g_idx.cpu().tolist() == gt_idx.tolist()
    assert abs(float(g_iou) - float(iou)) < 1e-5 and abs(float(g_l1) - float(l1)) < 1e-4 * max(1.0, float(l1))
    assert float((got_grad.cpu().double() - want_grad).norm() / want_grad.norm()) < 1e-4


def oracle_decode(raw):
    from oracle import vsrd_oracle as vo
    return vo.decode_box_parameters(*raw)


def test_fused_project_box_3d_matches_the_pytorch_ops_and_the_reference_golden():
    """`vsrd.operations.project_box_3d` on CUDA (one launch + one backward launch) against its own PyTorch-op form on
    the CPU (pinned to the reference's function by tests/golden/units.npz in test_vsrd_api.py): boxes in front of the
    camera, straddling the image plane and behind it; values and the gradient w.r.t. the corners."""
    import vsrd
    u = np.load(os.path.join(GOLDEN_DIR, "units.npz"))
    boxes = torch.from_numpy(u["pb_boxes_3d"])
    k = torch.from_numpy(u["pb_intrinsic"])
    lines = [[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]]
    got = torch.stack([vsrd.operations.project_box_3d(b.cuda(), lines, k.cuda()) for b in boxes])
    torch.testing.assert_close(got.cpu(), torch.from_numpy(u["pb_boxes_2d"]), rtol=1e-5, atol=1e-3)
    assert torch.equal(got[2].cpu(), torch.zeros(2, 2))                      # entirely behind the camera
    gen = torch.Generator().manual_seed(3)
    weights = torch.randn(len(boxes), 2, 2, generator=gen)
    for i, box in enumerate(boxes):
        a = box.clone().requires_grad_(True)
        b = box.clone().cuda().requires_grad_(True)
        ref = vsrd.operations.project_box_3d(a, lines, k)                    # CPU tensors -> the PyTorch-op path
        out = vsrd.operations.project_box_3d(b, lines, k.cuda())
        torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-3)
        if ref.requires_grad:
            (ref * weights[i]).sum().backward()
            (out * weights[i].cuda()).sum().backward()
            torch.testing.assert_close(b.grad.cpu(), a.grad, rtol=1e-4, atol=1e-4)
