"""The drop-in path a user of the reference actually hits: scripts/main.py's own closure composition
(main.py:433-523, 530-578) around `vsrd.rendering.sdfs.*`, `vsrd.models.*` and
`vsrd.rendering.hierarchical_volumetric_rendering`, under PyTorch autograd, on the GPU.  The composed field is
recognised (`match_union_field`) and rendered by the kernels; the result is held against the CPU oracle evaluated on
the very sample distances the call returned (the placement RNG differs from ATen's, the renderer must not)."""
import pytest
import torch

from oracle import vsrd_oracle as oracle

pytestmark = pytest.mark.gpu


def test_main_py_composition_through_the_drop_in_api_matches_the_oracle():
    import bench
    import vsrd
    from vsrd_b200 import synthetic
    dev = torch.device("cuda", 0)
    n, r, s = 3, 96, 16
    frame = synthetic.make_frame(num_instances=n, num_views=2, image_size=(94, 352), seed=5, intrinsics_scale=0.25)
    gen = torch.Generator().manual_seed(0)
    pix = frame.draw_pixel_indices(r, gen)
    h, w = frame.image_size
    inv_proj, cam = frame.inverse_projections()
    view, v, u = pix // (h * w), (pix // w) % h, pix % w
    dirs = torch.nn.functional.normalize(
        torch.einsum("rmn,rn->rm", inv_proj[view], torch.stack([u, v, torch.ones_like(u)], -1).float()), dim=-1)
    origins = cam[view].contiguous()
    targets = torch.rand(r, n, generator=gen)

    torch.manual_seed(1)
    detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n)
    raw = synthetic.perturbed_raw_parameters(frame, seed=5)
    with torch.no_grad():
        detector.locations.copy_(raw[0].reshape(1, n, 3)); detector.dimensions.copy_(raw[1].reshape(1, n, 3))
        detector.orientations.copy_(raw[2].reshape(1, n, 2))
        detector.embeddings.copy_(torch.rand(1, n, 256))
    hyper = vsrd.models.HyperDistanceField(in_channels=48, out_channels_list=[16] * 4, hyper_in_channels=256,
                                           hyper_out_channels_list=[256] * 4)
    encoder = vsrd.models.SinusoidalEncoder(num_frequencies=8)
    state = ({k: v.clone() for k, v in detector.state_dict().items()}, {k: v.clone() for k, v in hyper.state_dict().items()})
    detector, hyper, encoder = detector.to(dev), hyper.to(dev), encoder.to(dev)
    config = vsrd.utils.Dict.apply(dict(volume_rendering=dict(distance_range=[0.0, 100.0], num_fine_samples=s)))
    sched = dict(temperature=0.6, std_deviation=0.5, cosine_ratio=0.4)

    loss, labels, gradients, fine = bench.main_style_step(
        vsrd, (detector, hyper, encoder), config, origins.to(dev), dirs.to(dev), targets.to(dev), sched, n, return_outputs=True)
    assert labels.shape == (r, n) and gradients.shape == (2 * s - 1, r, 3) and fine.shape == (2 * s, r, 1)   # sample-major, as the reference
    params = [detector.locations, detector.dimensions, detector.orientations, detector.embeddings, *hyper.parameters()]
    got = torch.autograd.grad(loss, params)

    # ---- CPU oracle on the same sample distances, same parameters (fp64)
    leaves = [state[0][k][0].double().requires_grad_(True) for k in ("locations", "dimensions", "orientations", "embeddings")]
    ref_hyper = oracle.HyperNetwork().double()
    ref_hyper.load_state_dict({k: v.double() for k, v in state[1].items()})
    loc, dim, rot = oracle.decode_box_parameters(*leaves[:3])
    scene = oracle.Scene(loc, rot, dim, ref_hyper(leaves[3]), sched["temperature"])
    out = oracle.render_pass(scene.field(), origins.double(), dirs.double(), fine.detach().cpu().double(),
                             sched["std_deviation"], sched["cosine_ratio"])
    ref_loss = oracle.silhouette_loss(out[0], targets.double()) + 0.01 * oracle.eikonal_loss(out[1])
    want = torch.autograd.grad(ref_loss, leaves + list(ref_hyper.parameters()))

    assert float((labels.detach().cpu().double() - out[0].detach()).abs().max()) < 1e-4
    assert float((gradients.detach().cpu().double() - out[1].detach()).abs().max()) < 1e-3
    assert abs(float(loss) - float(ref_loss)) < 1e-4
    assert len(got) == len(want) == 27
    for p, a, b in zip(params, got, want):
        denom = float(b.norm())
        if denom > 1e-9:
            rel = float((a.detach().cpu().double().reshape(b.shape) - b).norm()) / denom
            assert rel < 5e-3, (tuple(p.shape), rel)
