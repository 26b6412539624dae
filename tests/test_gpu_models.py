"""GPU parity tests of the model-side kernels (csrc/vsrd_model.cu) through the C ABI:
BoxParameters3D decode (box_parameters.py:60-146), HyperDistanceField hypernetwork forward/backward
(hyper_distance_field.py:30-55, 75-77), Adam + ExponentialLR (config.json:177-215), each against the plain
PyTorch fp32/fp64 formulation of the same op (the reference's own nn.Modules under autograd), and the fused
FrameLabeler step against the autograd + torch.optim.Adam step it replaces."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(autouse=True)
def plain_torch_modules():
    """In this file the nn.Modules are the yardstick: run them as plain PyTorch ops, not through the model kernels."""
    from vsrd_b200 import functional
    functional.set_fused_modules(False)
    yield
    functional.set_fused_modules(True)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _models(n, seed=0, randomise_ln=True):
    import vsrd
    torch.manual_seed(seed)
    detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n).to(DEV)
    hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=256,
                                           hyper_out_channels_list=[256] * 4).to(DEV)
    with torch.no_grad():
        detector.locations.normal_(0.0, 0.3)
        detector.dimensions.normal_(0.0, 0.5)
        detector.orientations.normal_(0.0, 1.0)
        detector.embeddings.copy_(torch.rand_like(detector.embeddings))          # distinct rows per instance
        if randomise_ln:
            for m in hyper.modules():
                if isinstance(m, torch.nn.LayerNorm):
                    m.weight.uniform_(0.5, 1.5)
                    m.bias.uniform_(-0.3, 0.3)
    return detector, hyper


def _clone_models(detector, hyper, dtype=torch.float32):
    """Independent copies (deepcopy does not work on weight-normed modules: `weight` is a non-leaf attribute)."""
    import vsrd
    n = detector.locations.shape[1]
    det = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n).to(DEV)
    hyp = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=256,
                                         hyper_out_channels_list=[256] * 4).to(DEV)
    det.load_state_dict({k: v.clone() for k, v in detector.state_dict().items()})
    hyp.load_state_dict({k: v.clone() for k, v in hyper.state_dict().items()})
    return det.to(dtype), hyp.to(dtype)


def _arena(detector, hyper, steps=100, warm=0):
    from vsrd_b200.models import ParameterArena
    return ParameterArena(detector, hyper, [1e-2, 1e-2, 1e-2, 1e-3, 1e-4], num_steps=steps, warmup_steps=warm)


@pytest.mark.parametrize("n", [1, 8, 32])
def test_decode_boxes_matches_module(n):
    detector, hyper = _models(n)
    arena = _arena(detector, hyper)
    loc, dim, rot, boxes = arena.decode()
    ref = detector()
    assert torch.allclose(loc, ref["locations"][0], atol=1e-5, rtol=1e-6)
    assert torch.allclose(dim, ref["dimensions"][0], atol=1e-6, rtol=1e-6)
    assert torch.allclose(rot, ref["orientations"][0], atol=1e-6)
    assert torch.allclose(boxes, ref["boxes_3d"][0], atol=2e-5, rtol=1e-6)


@pytest.mark.parametrize("n", [1, 8, 32])
def test_decode_boxes_backward_matches_autograd(n):
    detector, hyper = _models(n, seed=1)
    arena = _arena(detector, hyper)
    loc, dim, rot, boxes = arena.decode()
    gen = torch.Generator(device=DEV).manual_seed(2)
    g_loc, g_dim = torch.randn(n, 3, device=DEV, generator=gen), torch.randn(n, 3, device=DEV, generator=gen)
    g_rot = torch.randn(n, 3, 3, device=DEV, generator=gen)
    g_boxes = torch.randn(2, n, 8, 3, device=DEV, generator=gen)
    w_iou, w_l1 = 0.1, 1.0
    parts = torch.tensor([0.25, 0.5], device=DEV)
    proj = torch.tensor([2.0, 3.0], device=DEV)
    losses = torch.zeros(5, device=DEV)
    arena.decode_backward(dim, rot, g_loc, g_dim, g_rot, g_boxes, w_iou, w_l1, parts, proj, losses)
    ref = detector()
    scalar = ((ref["locations"][0] * g_loc).sum() + (ref["dimensions"][0] * g_dim).sum() + (ref["orientations"][0] * g_rot).sum()
              + (ref["boxes_3d"][0] * (w_iou * g_boxes[0] + w_l1 * g_boxes[1])).sum())
    scalar.backward()
    for p in (detector.locations, detector.dimensions, detector.orientations):
        assert _rel(arena.grad(p), p.grad) < 1e-5, _rel(arena.grad(p), p.grad)
    assert torch.allclose(losses.cpu(), torch.tensor([0.25 + 0.5 + 0.2 + 3.0, 0.25, 0.5, 0.2, 3.0]), atol=1e-6)


@pytest.mark.parametrize("n", [1, 5, 8, 11, 32])
def test_hypernetwork_forward_backward_match_autograd(n):
    detector, hyper = _models(n, seed=3)
    arena = _arena(detector, hyper)
    w = arena.hyper_forward()
    ref = hyper(detector.embeddings)[0]
    assert w.shape == (n, 1617)
    assert torch.allclose(w, ref, atol=2e-5, rtol=1e-4), float((w - ref).abs().max())
    gw = torch.randn(n, 1617, device=DEV, generator=torch.Generator(device=DEV).manual_seed(4))
    arena.hyper_backward(gw)
    # fp64 autograd of the same module as the yardstick; the fp32 module's own error sets the tolerance
    _, hyper64 = _clone_models(detector, hyper, torch.float64)
    emb64 = detector.embeddings.detach().double().requires_grad_(True)
    (hyper64(emb64)[0] * gw.double()).sum().backward()
    (ref * gw).sum().backward()
    pairs = [(detector.embeddings, emb64.grad)] + list(zip(hyper.parameters(), [p.grad for p in hyper64.parameters()]))
    assert len(pairs) == 24
    for p, g64 in pairs:
        ours, theirs = _rel(arena.grad(p), g64), _rel(p.grad, g64)
        assert ours < max(1e-4, 3.0 * theirs), (tuple(p.shape), ours, theirs)


def test_hypernetwork_backward_is_deterministic():
    detector, hyper = _models(8, seed=5)
    arena = _arena(detector, hyper)
    arena.hyper_forward()
    gw = torch.randn(8, 1617, device=DEV)
    arena.hyper_backward(gw)
    first = arena.grads.clone()
    arena.grads.zero_()
    arena.hyper_backward(gw)
    assert torch.equal(first, arena.grads)


def test_adam_step_matches_torch_adam_with_exponential_lr():
    steps, warm = 12, 4
    detector, hyper = _models(4, seed=6)
    det_ref, hyp_ref = _clone_models(detector, hyper)
    arena = _arena(detector, hyper, steps=steps, warm=warm)
    groups = [[det_ref.locations], [det_ref.dimensions], [det_ref.orientations], [det_ref.embeddings], list(hyp_ref.parameters())]
    opt = torch.optim.Adam([dict(params=g, lr=lr) for g, lr in zip(groups, [1e-2, 1e-2, 1e-2, 1e-3, 1e-4])], lr=1e-2)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.01 ** (1.0 / steps))
    ours = [detector.locations, detector.dimensions, detector.orientations, detector.embeddings, *hyper.parameters()]
    refs = [p for g in groups for p in g]
    gen = torch.Generator(device=DEV).manual_seed(7)
    for step in range(steps):
        opt.zero_grad(set_to_none=True)
        for k, (p, q) in enumerate(zip(ours, refs)):
            if k >= 3 and step < warm:        # embeddings / hypernetwork have no gradient during the warm-up
                continue
            g = torch.randn(p.shape, device=DEV, generator=gen) * (1.0 + k)
            arena.grad(p).copy_(g)
            q.grad = g.clone()
        arena.adam_step(step=step)
        opt.step()
        sched.step()
    for p, q in zip(ours, refs):
        assert torch.allclose(p.data, q.data, atol=2e-6, rtol=1e-5), (tuple(p.shape), float((p.data - q.data).abs().max()))
    assert float((detector.locations.data - det_ref.locations.data).abs().max()) < 2e-6


def test_arena_keeps_module_api_and_state_dict():
    detector, hyper = _models(3, seed=8)
    before = {k: v.clone() for k, v in hyper.state_dict().items()}
    arena = _arena(detector, hyper)
    after = hyper.state_dict()
    assert list(before) == list(after)
    assert all(torch.equal(before[k], after[k]) for k in before)
    assert detector.locations.data_ptr() == arena.params.data_ptr()
    # the module forward reads the arena: an in-place arena update is visible through the module
    ref0 = hyper(detector.embeddings)
    arena.params.mul_(1.0)
    assert torch.equal(ref0, hyper(detector.embeddings))


SMALL = dict(num_instances=3, num_views=3, image_size=(94, 352), intrinsics_scale=0.25)


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_labeler_step_equals_autograd_step(use_graph):
    """The same frame optimised with models='fused' (vsrd_model.cu) and models='torch' (nn.Modules + autograd +
    torch.optim.Adam) on identical rays and samples: parameters and losses must track each other."""
    from vsrd_b200 import synthetic
    from vsrd_b200.frame import FrameLabeler, synthetic_frame_inputs
    frame = synthetic.make_frame(seed=3, **SMALL)
    raw = synthetic.perturbed_raw_parameters(frame, seed=3)
    init = dict(locations=raw[0].cuda(), dimensions=raw[1].cuda(), orientations=raw[2].cuda())
    inputs = synthetic_frame_inputs(frame, torch.device(DEV))
    steps, warm, r, s = 16, 6, 192, 24
    kw = dict(num_steps=steps, warmup_steps=warm, num_rays=r, num_samples=s, rays="indices", inject_samples=True,
              use_graph=use_graph, initial_parameters=init, model_seed=0)
    a = FrameLabeler(inputs, models="fused", **kw)
    b = FrameLabeler(inputs, models="torch", **kw)
    gen = torch.Generator().manual_seed(0)
    h, w = frame.image_size
    for step in range(steps):
        pix = frame.draw_pixel_indices(r, gen).cuda()
        jit = torch.rand(r, s, generator=gen).cuda()
        uni = torch.sort(torch.rand(r, s, generator=gen), dim=-1).values.cuda()
        a.step(pix, jitter=jit, sorted_uniforms=uni)
        b.step(pix, jitter=jit, sorted_uniforms=uni)
        a.synchronize(); b.synchronize()
        # the first steps of each phase pin the kernels (same parameters on both sides); later the two Adam
        # trajectories drift apart slowly (sign-like early updates amplify 1e-6 gradient differences)
        tight = step < 2 or warm <= step < warm + 2
        assert torch.allclose(a.losses, b.losses, rtol=1e-4 if tight else 1e-2, atol=1e-6), (step, a.losses.tolist(), b.losses.tolist())
    for name in ("locations", "dimensions", "orientations", "embeddings"):
        pa, pb = getattr(a.detector, name).data, getattr(b.detector, name).data
        assert torch.allclose(pa, pb, atol=5e-3), (name, float((pa - pb).abs().max()))
    ba, bb = a.boxes()["boxes_3d"], b.boxes()["boxes_3d"]
    assert float((ba - bb).abs().max()) < 2e-2
    moved = float((getattr(a.detector, "locations").data.cpu() - raw[0].reshape(1, -1, 3)).abs().max())
    assert moved > 1e-2
    # the hypernetwork was optimised too (after the warm-up): both paths moved it the same way
    init_hyper = FrameLabeler(inputs, models="torch", **kw).hyper.state_dict()
    moved_any = False
    for (ka, va), (kb, vb) in zip(a.hyper.state_dict().items(), b.hyper.state_dict().items()):
        assert ka == kb
        travel = float((vb - init_hyper[kb]).norm())
        moved_any |= travel > 0.0
        assert float((va - vb).norm()) <= 0.25 * travel + 1e-6, (ka, float((va - vb).norm()), travel)
    assert moved_any


@pytest.mark.parametrize("n", [1, 8, 32])
def test_modules_under_autograd_use_the_kernels_and_match_plain_torch(n):
    """`vsrd.models.*` on CUDA call the model kernels through autograd Functions (what an unchanged scripts/main.py
    gets): outputs and every parameter gradient against the same modules run as plain PyTorch ops."""
    from vsrd_b200 import functional
    detector, hyper = _models(n, seed=9)
    gen = torch.Generator(device=DEV).manual_seed(1)
    cot = dict(boxes_3d=torch.randn(1, n, 8, 3, device=DEV, generator=gen), locations=torch.randn(1, n, 3, device=DEV, generator=gen),
               dimensions=torch.randn(1, n, 3, device=DEV, generator=gen), orientations=torch.randn(1, n, 3, 3, device=DEV, generator=gen))
    gw = torch.randn(1, n, 1617, device=DEV, generator=gen)
    params = [detector.locations, detector.dimensions, detector.orientations, detector.embeddings, *hyper.parameters()]
    results = []
    for fused in (True, False):
        functional.set_fused_modules(fused)
        world = detector()
        weights = hyper(world["embeddings"])
        scalar = sum((world[k] * cot[k]).sum() for k in cot) + (weights * gw).sum()
        grads = torch.autograd.grad(scalar, params)
        results.append((world, weights, grads))
    (wa, ha, ga), (wb, hb, gb) = results
    assert ha.grad_fn is not None and "Hypernetwork" in type(ha.grad_fn).__name__
    assert "DecodeBoxes" in type(wa["boxes_3d"].grad_fn).__name__
    for k in cot:
        assert torch.allclose(wa[k], wb[k], atol=2e-5, rtol=1e-6), k
    assert torch.allclose(ha, hb, atol=2e-5, rtol=1e-4)
    for p, a, b in zip(params, ga, gb):
        assert _rel(a, b) < 2e-4, (tuple(p.shape), _rel(a, b))


def test_modules_fall_back_to_plain_torch_for_other_architectures():
    """A hypernetwork the kernels are not compiled for (other widths) still runs — as the nn.Sequential it is."""
    import vsrd
    from vsrd_b200 import functional
    functional.set_fused_modules(True)
    hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=64,
                                           hyper_out_channels_list=[128] * 2).to(DEV)
    out = hyper(torch.rand(1, 3, 64, device=DEV))
    assert out.shape == (1, 3, 1617) and "Hypernetwork" not in type(out.grad_fn).__name__
