"""CPU tests of the frame-level rows (a14 projection, a15 matching/losses, a2 ray selection, soft masks,
schedule): the oracle against the reference-generated goldens (tests/golden/frame.npz), and the device
math header (vsrd_frame_math.cuh) compiled for the host against the same goldens."""
import os
import shutil

import numpy as np
import pytest
import torch

from oracle import frame_oracle as fo
from tests.helpers import GOLDEN_DIR, rel_l2

CASES = ["ordered_f32", "shuffled_f32", "ordered_f64", "shuffled_f64"]


def load_case(name):
    data = np.load(os.path.join(GOLDEN_DIR, "frame.npz"))
    pre = f"proj_{name}."
    return {k[len(pre):]: torch.from_numpy(data[k]) for k in data.files if k.startswith(pre)}


def load_soft():
    data = np.load(os.path.join(GOLDEN_DIR, "frame.npz"))
    return {k[5:]: torch.from_numpy(data[k]) for k in data.files if k.startswith("soft.")}


@pytest.mark.parametrize("case", CASES)
def test_oracle_projection_matches_reference(case):
    g = load_case(case)
    tol = 1e-12 if case.endswith("f64") else 1e-5
    world = g["world_boxes"].clone().requires_grad_(True)
    size = tuple(int(x) for x in g["image_size"])
    boxes, gt_idx, iou, l1 = fo.projection_step(world, g["extrinsics"], g["intrinsics"], size, g["gt_boxes_2d"],
                                                g["visible"], int(g["target_view"]))
    assert torch.allclose(boxes, g["boxes_2d"], rtol=tol, atol=tol * 100)
    t = int(g["target_view"])
    cost = -fo.distance_box_iou(boxes[t].detach(), g["gt_boxes_2d"][t])
    assert torch.allclose(cost, g["cost"], rtol=tol, atol=tol)
    assert torch.equal(gt_idx, g["gt_indices"]) and torch.equal(g["pd_indices"], torch.arange(gt_idx.numel()))
    assert abs(float(iou) - float(g["iou_loss"])) <= tol * max(1.0, abs(float(iou)))
    assert abs(float(l1) - float(g["l1_loss"])) <= tol * max(1.0, abs(float(l1)))
    g_iou, = torch.autograd.grad(iou, world, retain_graph=True)
    g_l1, = torch.autograd.grad(l1, world)
    assert rel_l2(g_iou, g["grad_iou"]) < (1e-10 if case.endswith("f64") else 1e-4)
    assert rel_l2(g_l1, g["grad_l1"]) < (1e-10 if case.endswith("f64") else 1e-4)


def test_oracle_distance_map_matches_reference():
    g = load_soft()
    size = tuple(int(x) for x in g["image_size"])
    assert torch.equal(fo.polygon_distance_map(g["polygon"], size), g["distance_map"])


def test_select_rays_is_sequential_sampling_without_replacement():
    gen = np.random.default_rng(0)
    w = gen.random(50) ** 3
    w[[3, 17, 18]] = 0.0
    picks = fo.select_rays(w, 20, gen.random(200))
    assert len(set(picks.tolist())) == 20 and all(w[i] > 0 for i in picks.tolist())
    # first-draw frequencies follow the weights; the pair distribution follows the renormalised remainder
    first = np.zeros(50)
    second_given = np.zeros(50)
    trials = 20000
    for _ in range(trials):
        a, b = fo.select_rays(w, 2, gen.random(16)).tolist()
        first[a] += 1
        if a == 7:
            second_given[b] += 1
    assert np.abs(first / trials - w / w.sum()).max() < 0.01
    rest = w.copy(); rest[7] = 0.0
    assert np.abs(second_given / second_given.sum() - rest / rest.sum()).max() < 0.05
    with pytest.raises(RuntimeError):
        fo.select_rays(w, 48, gen.random(10000))          # only 47 pixels have weight


def test_select_rays_matches_torch_multinomial_distribution():
    """torch.multinomial(replacement=False) (main.py:620) and the oracle draw from the same distribution:
    compare inclusion frequencies of each pixel over many trials."""
    gen = np.random.default_rng(1)
    tg = torch.Generator().manual_seed(1)
    w = torch.tensor(gen.random(12) ** 2, dtype=torch.float32)
    trials, k = 6000, 4
    inc_ref, inc_ours = np.zeros(12), np.zeros(12)
    for _ in range(trials):
        inc_ref[torch.multinomial(w, k, replacement=False, generator=tg).numpy()] += 1
        inc_ours[fo.select_rays(w.numpy(), k, gen.random(64)).numpy()] += 1
    assert np.abs(inc_ref - inc_ours).max() / trials < 0.03


def test_schedule_endpoints():
    s0 = fo.schedule(0, 3000, 1000)
    s1 = fo.schedule(2999, 3000, 1000)
    assert s0["temperature"] == 1.0 and s0["eikonal_weight"] == 0.0 and s0["cosine_ratio"] == 0.0
    assert abs(s1["std_deviation"] - 0.1) < 1e-5 and s1["eikonal_weight"] == 0.01


# ---- device math compiled for the host -----------------------------------------------------------
needs_gxx = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def hs():
    from tests import hostsim
    hostsim.build()
    return hostsim


@needs_gxx
@pytest.mark.parametrize("case", ["ordered_f32", "shuffled_f32"])
def test_device_projection_math_matches_reference(hs, case):
    g = load_case(case)
    g64 = load_case(case.replace("f32", "f64"))
    size = tuple(int(x) for x in g["image_size"])
    t = int(g["target_view"])
    boxes, cost, losses, grad = hs.projection_step(g["extrinsics"], g["intrinsics"], g["world_boxes"], size,
                                                   g["gt_boxes_2d"], g["visible"], g["gt_indices"], target_view=t)
    assert torch.allclose(boxes, g["boxes_2d"], rtol=1e-5, atol=1e-3)
    assert torch.allclose(cost, g["cost"], rtol=1e-4, atol=1e-5)
    from scipy.optimize import linear_sum_assignment
    assert np.array_equal(linear_sum_assignment(cost.numpy())[1], g["gt_indices"].numpy())
    assert abs(float(losses[0]) - float(g["iou_loss"])) < 1e-5 * max(1.0, float(g["iou_loss"]))
    assert abs(float(losses[1]) - float(g["l1_loss"])) < 1e-5 * max(1.0, float(g["l1_loss"]))
    # gradients: against the reference's fp64 autograd (same scene), tolerance of fp32 arithmetic
    assert rel_l2(grad[0].double(), g64["grad_iou"]) < 1e-3
    assert rel_l2(grad[1].double(), g64["grad_l1"]) < 1e-3
    # projection-only mode
    only = hs.projection_step(g["extrinsics"], g["intrinsics"], g["world_boxes"], size)
    assert torch.equal(only, boxes)


@needs_gxx
def test_device_soft_mask_math_matches_reference(hs):
    g = load_soft()
    size = tuple(int(x) for x in g["image_size"])
    sd, mask = hs.soft_mask(g["polygon"], size)
    assert torch.allclose(sd.abs(), g["distance_map"], rtol=1e-5, atol=1e-4)
    ref_mask = fo.soft_mask(g["polygon"], size)
    assert torch.allclose(mask, ref_mask, atol=1e-5)
    inside = fo.polygon_inside(g["polygon"], size)
    assert torch.equal(sd > 0, inside)
