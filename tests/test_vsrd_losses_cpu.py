"""`vsrd.losses` against the reference package's outputs (tests/golden/losses.npz, written by
tests/golden/make_golden_losses.py from the unmodified /root/reference/vsrd/losses), plus the step losses of
scripts/main.py against the oracle."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import vsrd

HERE = os.path.dirname(os.path.abspath(__file__))


def test_losses_match_the_reference_package_outputs():
    spec = importlib.util.spec_from_file_location("make_golden_losses", os.path.join(HERE, "golden", "make_golden_losses.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    ours = module.cases(vsrd.losses)
    golden = np.load(os.path.join(HERE, "golden", "losses.npz"))
    assert set(ours) == set(golden.files)
    for key in golden.files:
        assert ours[key].shape == golden[key].shape and ours[key].dtype == golden[key].dtype, key
        assert np.allclose(ours[key], golden[key], rtol=2e-6, atol=1e-7), (key, np.abs(ours[key] - golden[key]).max())


def test_reduction_argument_is_validated():
    with pytest.raises(ValueError):
        vsrd.losses.focal_loss(torch.rand(2, 2), torch.rand(2, 2), reduction="median")


def test_step_losses_match_the_oracle():
    from oracle import vsrd_oracle as oracle
    gen = torch.Generator().manual_seed(1)
    labels, targets = torch.rand(50, 4, generator=gen), torch.rand(50, 4, generator=gen)
    labels[0, 0], labels[1, 1] = 0.0, 1.0                      # the clamp (main.py:655)
    grads = torch.randn(31, 50, 3, generator=gen)
    pd, gt = torch.tensor([2, 0, 3, 1]), torch.tensor([0, 1, 2, 3])
    assert torch.equal(vsrd.losses.silhouette_loss(labels, targets, pd, gt), oracle.silhouette_loss(labels, targets, pd, gt))
    assert torch.equal(vsrd.losses.eikonal_loss(grads), oracle.eikonal_loss(grads))
    with pytest.raises(RuntimeError, match="CUDA"):
        vsrd.losses.projection_losses(torch.zeros(2, 8, 3), torch.eye(4)[None], torch.eye(3)[None], (4, 4),
                                      torch.zeros(1, 2, 4), torch.ones(1, 2, dtype=torch.bool))
