"""Pin `vsrd.losses` to the UNMODIFIED reference package (build container only; needs /root/reference):

    python tests/golden/make_golden_losses.py   ->  tests/golden/losses.npz
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402


def cases(losses):
    gen = torch.Generator().manual_seed(0)
    p = torch.rand(2, 3, 6, 7, generator=gen)
    t = torch.rand(2, 3, 6, 7, generator=gen)
    q = torch.softmax(torch.randn(4, 5, generator=gen), dim=-1)
    r = torch.softmax(torch.randn(4, 5, generator=gen), dim=-1)
    e1 = torch.eye(4).repeat(3, 1, 1) + 0.1 * torch.randn(3, 4, 4, generator=gen)
    e2 = torch.eye(4).repeat(3, 1, 1) + 0.1 * torch.randn(3, 4, 4, generator=gen)
    k1, k2 = torch.rand(3, 9, 2, generator=gen) * 50, torch.rand(3, 9, 2, generator=gen) * 50
    fm = torch.randn(3, 3, 3, generator=gen)
    out = {}
    for name in ("binary_cross_entropy", "binary_kl_divergence", "binary_js_divergence", "focal_loss",
                 "quality_focal_loss", "tversky_loss", "focal_tversky_loss"):
        for reduction in ("none", "mean", "sum"):
            out[f"{name}.{reduction}"] = getattr(losses, name)(p, t, reduction=reduction)
    for name in ("cross_entropy", "kl_divergence", "js_divergence"):
        out[f"{name}.dim"] = getattr(losses, name)(q, r, dim=-1, reduction="none")
        out[f"{name}.mean"] = getattr(losses, name)(q, r)
    out["rotation_consistency_loss"] = losses.rotation_consistency_loss(e1, e2, reduction="none")
    out["translation_consistency_loss"] = losses.translation_consistency_loss(e1, e2, reduction="none")
    out["sampson_epipolar_distance"] = losses.sampson_epipolar_distance(k1, k2, fm[:, None], reduction="none")
    out["ssim_loss"] = losses.ssim_loss(p, t, reduction="none")
    out["photometric_loss"] = losses.photometric_loss(p, t)
    out["smoothness_loss"] = losses.smoothness_loss(p[:, :1], t, reduction="none")
    out["smoothness_loss.raw"] = losses.smoothness_loss(p[:, :1], t, normalize=False)
    out["motion_smoothness_loss"] = losses.motion_smoothness_loss(p, reduction="none")
    out["motion_sparsity_loss"] = losses.motion_sparsity_loss(p - 0.5, reduction="sum")
    out["gradient_x"] = losses.gradient_x(p)
    out["gradient_y"] = losses.gradient_y(p)
    return {k: v.numpy() for k, v in out.items()}


if __name__ == "__main__":
    with ref_import.reference_modules():
        data = cases(importlib.import_module("vsrd.losses"))
    np.savez(os.path.join(HERE, "losses.npz"), **data)
    print(f"wrote {len(data)} entries")
