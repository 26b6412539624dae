"""vsrd_b200.predictions (checkpoint -> pseudo-label records with confidences) against tests/golden/predictions.npz,
produced by tests/golden/make_golden_predictions.py from the reference tool's arithmetic
(tools/kitti_360/make_predictions.py:58-192) around the unmodified reference modules."""
import importlib.util
import json
import os

import numpy as np
import torch

from vsrd_b200 import predictions

HERE = os.path.dirname(os.path.abspath(__file__))


def _group():
    spec = importlib.util.spec_from_file_location("make_golden_predictions", os.path.join(HERE, "golden", "make_golden_predictions.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module.synthetic_group()


def test_predictions_match_the_reference_tool_arithmetic():
    golden = np.load(os.path.join(HERE, "golden", "predictions.npz"))
    raw, views, target_extrinsic, target_ids, image_size = _group()
    state = {k[len("state."):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("state.")}
    boxes = predictions.boxes_from_checkpoint(dict(models=dict(detector=state)))
    records = predictions.make_frame_predictions(boxes, target_extrinsic, target_ids, views, image_size)
    assert len(records) == len(views)
    for k, record in enumerate(records):
        assert set(record) == {"boxes_3d", "boxes_2d", "confidences"} and set(record["boxes_3d"]) == {"car"}
        np.testing.assert_allclose(np.array(record["boxes_3d"]["car"]), golden[f"boxes_3d_{k}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(np.array(record["boxes_2d"]["car"]), golden[f"boxes_2d_{k}"], rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(np.array(record["confidences"]["car"]), golden["confidences"], rtol=1e-5, atol=1e-6)
    json.dumps(records[0])                                   # what the tool writes


def test_rectification_of_a_level_camera_is_the_identity():
    assert torch.allclose(predictions.rectification_matrix(torch.eye(4)), torch.eye(3))
