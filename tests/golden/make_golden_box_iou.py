"""tests/golden/box_iou.npz: the reference's `vsrd.operations.box_3d_iou` (kitti360_operations.py, imported unmodified)
on the random box pairs of tests/test_vsrd_api.py::_random_box_pairs (build container only)."""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402
from tests.test_vsrd_api import _random_box_pairs  # noqa: E402

if __name__ == "__main__":
    out = {}
    with ref_import.reference_modules():
        ops = importlib.import_module("vsrd.operations.kitti360_operations")
        for dtype, key in ((np.float32, "f32"), (np.float64, "f64")):
            out[key] = np.array([[float(v) for v in ops.box_3d_iou(torch.from_numpy(a), torch.from_numpy(b))]
                                 for a, b in _random_box_pairs(300, seed=7, dtype=dtype)])
    np.savez(os.path.join(HERE, "box_iou.npz"), **out)
    print({k: (v.shape, float(v[:, 0].mean())) for k, v in out.items()})
