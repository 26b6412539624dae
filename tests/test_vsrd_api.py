"""CPU-side checks of the drop-in `vsrd` API: closure pattern-matcher, models, operations, utils."""
import functools
import operator
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import vsrd
from tests.helpers import GOLDEN_DIR
from vsrd.rendering import UnsupportedFieldError, match_union_field, sdfs


def _attr(**kw):
    return vsrd.utils.Dict.apply(kw)


def compose_like_main(locations, dimensions, orientations, weights, temperature, models, config, num_instances):
    """Closures with the same nesting and free-variable names as scripts/main.py:433-578 (written
    for this test; the GPU-box tests cannot read the reference script)."""

    def residual_distance_field(distance_field):
        def wrapper(positions):
            x, y, z = torch.unbind(positions, dim=-1)
            positions = torch.stack([torch.abs(x), y, z], dim=-1)                  # main.py:437-438
            positions = positions / max(config.volume_rendering.distance_range)
            return torch.sigmoid(distance_field(models.positional_encoder(positions)) - 1.0)
        return wrapper

    def residual_composition(distance_field, residual_distance_field):
        def wrapper(positions):
            return distance_field(positions) + residual_distance_field(positions)
        return wrapper

    def instance_field(distance_field, instance_label):
        def wrapper(positions):
            distances = distance_field(positions)
            labels = nn.functional.one_hot(instance_label, num_instances)
            return distances, labels.expand(*distances.shape[:-1], -1)
        return wrapper

    def soft_union(distance_fields, temperature):
        def wrapper(positions):
            distances, labels = map(torch.stack, zip(*[f(positions) for f in distance_fields]))
            w = nn.functional.softmin(distances / temperature, dim=0)
            return torch.sum(distances * w, dim=0), torch.sum(labels * w, dim=0)
        return wrapper

    fields = []
    for i in range(num_instances):
        inner = sdfs.box(dimensions[i])
        if weights is not None:
            inner = residual_composition(
                distance_field=inner,
                residual_distance_field=residual_distance_field(
                    distance_field=functools.partial(models.hyper_distance_field.distance_field, weights[i])))
        inst = instance_field(distance_field=inner, instance_label=dimensions.new_tensor(i, dtype=torch.long))   # main.py:543
        fields.append(sdfs.translation(sdfs.rotation(inst, orientations[i]), locations[i]))
    return soft_union(distance_fields=fields, temperature=temperature)


@pytest.fixture(scope="module")
def scene():
    torch.manual_seed(0)
    n = 3
    detector = vsrd.models.BoxParameters3D(batch_size=1, num_instances=n)
    hdf = vsrd.models.HyperDistanceField(48, [16, 16, 16, 16], 256, [256, 256, 256, 256])
    enc = vsrd.models.SinusoidalEncoder(8)
    models = _attr(detector=detector, hyper_distance_field=hdf, positional_encoder=enc)
    config = _attr(volume_rendering=dict(distance_range=[0.0, 100.0]))
    world = detector()
    return n, models, config, world, hdf(world["embeddings"])


@pytest.mark.parametrize("residual", [False, True])
def test_matcher_recovers_scene(scene, residual):
    n, models, config, world, weights = scene
    field = compose_like_main(world["locations"][0], world["dimensions"][0], world["orientations"][0],
                              weights[0] if residual else None, 0.37, models, config, n)
    u = match_union_field(field)
    assert u.temperature == pytest.approx(0.37) and u.scale == 100.0 and u.returns_features
    assert torch.equal(u.locations, world["locations"][0])
    assert torch.equal(u.rotations, world["orientations"][0])
    assert torch.equal(u.half_extents, world["dimensions"][0])
    if residual:
        assert torch.equal(u.mlp_weights, weights[0])
        # autograd connectivity: gradients reach the detector and the hypernetwork through the stack
        (u.locations.sum() + u.mlp_weights.sum()).backward()
        assert models.detector.locations.grad is not None
        assert all(p.grad is not None for p in models.hyper_distance_field.parameters())
    else:
        assert u.mlp_weights is None
    # logging path: compose(field, itemgetter(0))  (main.py:1030)
    u2 = match_union_field(vsrd.utils.compose(field, operator.itemgetter(0)))
    assert not u2.returns_features and torch.equal(u2.locations, u.locations)


def test_composed_field_still_evaluates_in_plain_pytorch(scene):
    """The tagged leaves keep the reference's call protocol (positions -> distances)."""
    n, models, config, world, weights = scene
    field = compose_like_main(world["locations"][0], world["dimensions"][0], world["orientations"][0],
                              weights[0], 0.5, models, config, n)
    sdf, labels = field(torch.randn(5, 7, 3) * 10)
    assert sdf.shape == (5, 7, 1) and labels.shape == (5, 7, n)
    torch.testing.assert_close(labels.sum(-1), torch.ones(5, 7))


def test_unrecognised_fields_raise(scene):
    n, models, config, world, weights = scene
    with pytest.raises(UnsupportedFieldError):
        match_union_field(lambda x: (x.norm(dim=-1, keepdim=True), x))
    with pytest.raises(UnsupportedFieldError):
        match_union_field(sdfs.box(torch.ones(3)))
    # a union whose members are not translation(rotation(...))
    def soft_union(distance_fields, temperature):
        def wrapper(positions):
            return distance_fields[0](positions) / temperature
        return wrapper
    with pytest.raises(UnsupportedFieldError, match="translation"):
        match_union_field(soft_union([sdfs.box(torch.ones(3))], 1.0))
    # unsupported MLP geometry
    small = vsrd.models.HyperDistanceField(48, [8, 8], 16, [16])
    with pytest.raises(RuntimeError, match="48-16-16-16-16-1"):
        small.check_fused_layout()


def test_models_match_reference_fixtures():
    u = np.load(os.path.join(GOLDEN_DIR, "units.npz"))
    t = lambda k: torch.from_numpy(u[k])
    det = vsrd.models.BoxParameters3D(batch_size=1, num_instances=5)
    assert sorted(det.state_dict()) == ["dimension_range", "dimensions", "embeddings", "location_range",
                                        "locations", "orientations"]
    with torch.no_grad():
        det.locations.copy_(t("bp_raw_locations"))
        det.dimensions.copy_(t("bp_raw_dimensions"))
        det.orientations.copy_(t("bp_raw_orientations"))
    world = det()
    torch.testing.assert_close(world["locations"], t("bp_locations"), rtol=0, atol=0)
    torch.testing.assert_close(world["dimensions"], t("bp_dimensions"), rtol=0, atol=0)
    torch.testing.assert_close(world["orientations"], t("bp_orientations"), rtol=0, atol=0)
    torch.testing.assert_close(world["boxes_3d"], t("bp_boxes_3d"), rtol=0, atol=1e-6)
    loc, dim, rot = det.encode_box_3d(world["boxes_3d"])
    torch.testing.assert_close(loc, t("bp_enc_locations"), rtol=0, atol=1e-6)
    torch.testing.assert_close(dim, t("bp_enc_dimensions"), rtol=0, atol=1e-6)
    torch.testing.assert_close(rot, t("bp_enc_orientations"), rtol=0, atol=1e-6)

    hdf = vsrd.models.HyperDistanceField(48, [16, 16, 16, 16], 256, [256, 256, 256, 256])
    keys = sorted(hdf.state_dict())
    assert keys == list(u["hyper_state_keys"])
    assert [hdf.state_dict()[k].numel() for k in keys] == list(u["hyper_state_numel"])
    assert hdf.num_neurons_list == [784, 272, 272, 272, 17]
    enc = vsrd.models.SinusoidalEncoder(8)
    torch.testing.assert_close(enc(t("mlp_points")), t("mlp_encoding"), rtol=0, atol=1e-6)
    w = t("mlp_weights")
    torch.testing.assert_close(hdf.distance_field(w[0], t("mlp_encoding")), t("mlp_out0"), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(hdf.distance_field(w[1], t("mlp_encoding")), t("mlp_out1"), rtol=1e-5, atol=1e-5)


def test_project_box_3d_matches_reference_fixtures():
    u = np.load(os.path.join(GOLDEN_DIR, "units.npz"))
    boxes, k = torch.from_numpy(u["pb_boxes_3d"]), torch.from_numpy(u["pb_intrinsic"])
    lines = u["pb_line_indices"].tolist()
    got = torch.stack([vsrd.operations.project_box_3d(b, lines, k) for b in boxes])
    torch.testing.assert_close(got, torch.from_numpy(u["pb_boxes_2d"]), rtol=1e-5, atol=1e-3)
    assert torch.equal(got[2], torch.zeros(2, 2))       # box behind the camera


def _random_box_pairs(count, seed=0, dtype=np.float32):
    """Box pairs exactly as scripts/main.py:892-899 hands them to `box_3d_iou`: corners in the order of
    `BoxParameters3D.decode_box_3d` (box_parameters.py:73-91; camera frame, y down) rotated by
    `rotation_matrix_x(-pi/2)` so that Z is up.  Ordinary overlaps, near-coincident pairs and distant pairs."""
    rng = np.random.default_rng(seed)
    signs = np.array([[-1, -1, 1], [1, -1, 1], [1, -1, -1], [-1, -1, -1], [-1, 1, 1], [1, 1, 1], [1, 1, -1], [-1, 1, -1]], float)
    to_z_up = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -1.0, 0.0]])            # rotation_matrix_x(-pi/2)

    def box(centre, half, yaw):
        c, s = np.cos(yaw), np.sin(yaw)
        rot_y = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        return (((signs * half) @ rot_y.T + centre) @ to_z_up.T).astype(dtype)

    pairs = []
    for k in range(count):
        half = rng.uniform([0.75, 0.75, 1.5], [1.0, 1.0, 2.5])
        centre, yaw = rng.uniform([-20, 0, 5], [20, 1.5, 60]), rng.uniform(-np.pi, np.pi)
        spread = [0.02, 0.5, 3.0][k % 3]
        other = box(centre + rng.normal(0, spread, 3) * [1, 0.2, 1], half * rng.uniform(0.9, 1.1, 3), yaw + rng.normal(0, 0.2 * spread))
        pairs.append((box(centre, half, yaw), other))
    return pairs


def test_box_3d_iou_is_value_identical_to_the_reference_and_exact_variant_is_exact():
    """tests/golden/box_iou.npz: the reference's `box_3d_iou` (kitti360_operations.py:84-117, imported unmodified by
    tests/golden/make_golden_box_iou.py) on 300 random pairs, float32 and float64 corners."""
    golden = np.load(os.path.join(GOLDEN_DIR, "box_iou.npz"))
    for dtype, key in ((np.float32, "f32"), (np.float64, "f64")):
        pairs = _random_box_pairs(300, seed=7, dtype=dtype)
        got = np.array([[float(v) for v in vsrd.operations.box_3d_iou(torch.from_numpy(a), torch.from_numpy(b))] for a, b in pairs])
        assert np.array_equal(got, golden[key]), np.abs(got - golden[key]).max()
    exact = np.array([[float(v) for v in vsrd.operations.box_3d_iou_exact(a, b)] for a, b in _random_box_pairs(300, seed=7, dtype=np.float64)])
    # the reference's +0.01 fudge (:29) is visible against the exact clip, worst on near-coincident pairs
    gap = np.abs(exact - golden["f64"])[:, 0]
    assert 0.01 < gap.max() and np.median(gap) < 0.02

    b = np.array([[-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1]], float)
    iou, bev = vsrd.operations.box_3d_iou_exact(b, b)
    assert float(iou) == pytest.approx(1.0) and float(bev) == pytest.approx(1.0)
    iou, bev = vsrd.operations.box_3d_iou_exact(b, b + np.array([1.0, 0, 0]))
    assert float(iou) == pytest.approx(1 / 3) and float(bev) == pytest.approx(1 / 3)
    iou, _ = vsrd.operations.box_3d_iou_exact(b, b + np.array([5.0, 0, 0]))
    assert float(iou) == 0.0
    c, s = np.cos(0.3), np.sin(0.3)
    rot = b @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]).T
    iou_rot, _ = vsrd.operations.box_3d_iou_exact(b, rot)
    assert 0.5 < float(iou_rot) < 1.0


def test_utils_smoke():

    d = vsrd.utils.Dict.apply({"a": {"b": 1}, "c": [{"d": 2}]})
    assert d.a.b == 1 and d.c[0].d == 2
    f = vsrd.utils.compose(lambda x: x + 1, lambda x: x * 2)
    assert f(3) == 8
    made = vsrd.utils.import_module({"function": "torch.zeros", "args": [2], "kwargs": {"dtype": "eval:torch.float64"}},
                                    globals())
    assert made.dtype == torch.float64 and made.shape == (2,)
    assert torch.equal(vsrd.operations.expand_to_4x4(torch.ones(2, 3, 3))[0, 3], torch.tensor([0.0, 0.0, 0.0, 1.0]))


def test_gather_rows_reviews_the_base_tensor_instead_of_stacking():
    """main.py iterates `world_outputs.locations[0]` etc.; the matcher must hand the kernels the base tensor (same
    values, same gradients) and fall back to torch.stack for anything else."""
    from vsrd.rendering.renderers import _gather_rows
    base = torch.rand(1, 5, 3, requires_grad=True)
    rows = list(base[0] * 1.0)                       # non-leaf base, unbind views — as `detector()` outputs are
    got = _gather_rows(rows)
    assert got.shape == (5, 3) and torch.equal(got, torch.stack(rows))
    assert got._base is rows[0]._base or got is rows[0]._base          # a view, not a copy
    (got * torch.arange(15.0).reshape(5, 3)).sum().backward()
    assert torch.equal(base.grad[0], torch.arange(15.0).reshape(5, 3))
    # not consecutive rows of one tensor: plain stack
    other = [torch.rand(3) for _ in range(4)]
    assert torch.equal(_gather_rows(other), torch.stack(other))
    rev = list(reversed(list(torch.rand(4, 3))))
    assert torch.equal(_gather_rows(rev), torch.stack(rev))
    part = list(torch.rand(6, 3))[:4]                # a prefix of a larger tensor
    assert torch.equal(_gather_rows(part), torch.stack(part))


def _verbatim_field(scene, residual, union="soft_union", temperature=0.37):
    """main.py's own closure factories, compiled verbatim from the script's AST (oracle/ref_import.py; test
    infrastructure), composed around THIS package's sdfs leaves and models exactly as main.py:530-578 does."""
    from oracle import ref_import
    n, models, config, world, weights = scene
    closures = ref_import.main_closures(dict(torch=torch, nn=nn, config=config, models=models, num_instances=n))
    locations, dimensions, orientations = world["locations"][0], world["dimensions"][0], world["orientations"][0]
    fields = []
    for label in range(n):
        inner = vsrd.rendering.sdfs.box(dimensions[label])
        if residual:
            inner = closures["residual_composition"](
                distance_field=inner,
                residual_distance_field=closures["residual_distance_field"](
                    distance_field=functools.partial(models.hyper_distance_field.distance_field, weights[0][label])))
        inst = closures["instance_field"](distance_field=inner, instance_label=dimensions.new_tensor(label, dtype=torch.long))
        fields.append(vsrd.rendering.sdfs.translation(vsrd.rendering.sdfs.rotation(inst, orientations[label]), locations[label]))
    if union == "hard_union":
        return closures["hard_union"](fields)
    return closures["soft_union"](distance_fields=fields, temperature=temperature)


def _reference_available():
    from oracle import ref_import
    return ref_import.available()


@pytest.mark.skipif(not _reference_available(), reason="scripts/main.py neither mounted nor staged")
@pytest.mark.parametrize("residual", [False, True])
def test_matcher_recovers_scene_from_the_verbatim_main_py_closures(scene, residual):
    n, models, config, world, weights = scene
    field = _verbatim_field(scene, residual)
    u = match_union_field(field)
    assert u.temperature == pytest.approx(0.37) and u.scale == 100.0 and not u.hard and len(u.code_key) >= 2
    assert torch.equal(u.locations, world["locations"][0]) and torch.equal(u.half_extents, world["dimensions"][0])
    assert torch.equal(u.rotations, world["orientations"][0])
    assert (u.mlp_weights is not None) == residual
    if residual:
        assert torch.equal(u.mlp_weights, weights[0])
    hard = match_union_field(_verbatim_field(scene, residual, union="hard_union"))
    assert hard.hard and hard.temperature == 1e-6
    logged = match_union_field(vsrd.utils.compose(field, operator.itemgetter(0)))      # main.py:1030
    assert not logged.returns_features
    # the verbatim closure itself evaluates on this package's leaves (plain PyTorch on the CPU)
    sdf, labels = field(torch.randn(5, 3) * 3.0 + world["locations"][0][0].detach())
    assert sdf.shape == (5, 1) and labels.shape == (5, n)
