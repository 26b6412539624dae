"""Checkpoint -> pseudo-label predictions with confidences (tools/kitti_360/make_predictions.py:26-192).

For one optimised target frame: load the detector state of its final checkpoint, move the boxes into every frame of
the target's group (`source_extrinsic @ inv(target_extrinsic) @ rectification^T`, :108-115), project them
(`project_box_3d` + clip to the image, :117-139), score each predicted box against each annotated instance by the 2D
IoU with the instance's mask box averaged over the frames where the instance is annotated (:141-165), match
predictions to instances (Hungarian, maximising the mean IoU, :190) and attach the matched mean IoU as the confidence
of every per-frame prediction record `{boxes_3d, boxes_2d, confidences}` (:167-192).

The on-disk parts of the reference tool (KITTI-360 annotation JSON, pycocotools masks, MaskRefiner) are replaced by
plain tensors: each group view carries its intrinsic / extrinsic matrices, the 2D boxes of its annotated instances and
their instance ids.  tools/make_predictions.py feeds it from checkpoints + the synthetic dataset.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import scipy.optimize
import torch
import torchvision

LINE_INDICES = [[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]]


def rectification_matrix(target_extrinsic: torch.Tensor) -> torch.Tensor:
    """Rotation about x that levels the target camera (make_predictions.py:72-80, kitti_360_dataset.py:214-220)."""
    import vsrd
    x_axis, y_axis, _ = target_extrinsic[..., :3, :3]
    upright = torch.round(y_axis)
    angle = torch.acos(torch.dot(upright, y_axis).clamp(-1.0, 1.0)) * torch.sign(torch.dot(torch.linalg.cross(upright, y_axis), x_axis))
    return vsrd.operations.rotation_matrix_x(angle)


def boxes_from_checkpoint(checkpoint: Dict) -> torch.Tensor:
    """World boxes [N,8,3] of `checkpoint["models"]["detector"]` (make_predictions.py:58-62)."""
    import vsrd
    state = checkpoint["models"]["detector"]
    model = vsrd.models.BoxParameters3D(*torch.as_tensor(state["embeddings"]).shape)
    model.load_state_dict(state)
    with torch.no_grad():
        boxes, = model()["boxes_3d"]
    return boxes.detach().cpu().float()


def make_frame_predictions(world_boxes_3d: torch.Tensor, target_extrinsic: torch.Tensor, target_instance_ids: torch.Tensor,
                           group_views: Sequence[Dict], image_size, class_name: str = "car") -> List[Dict]:
    """`group_views[k]` = dict(intrinsic_matrix [3,3], extrinsic_matrix [4,4] (raw, un-rectified), boxes_2d [M_k,2,2],
    instance_ids [M_k]).  Returns one prediction record per view (the JSON the reference writes)."""
    import vsrd
    num_pd, num_gt = world_boxes_3d.shape[0], int(target_instance_ids.numel())
    homogeneous = torch.nn.functional.pad(world_boxes_3d, (0, 1), mode="constant", value=1.0)
    to_world = torch.linalg.inv(target_extrinsic) @ vsrd.operations.expand_to_4x4(rectification_matrix(target_extrinsic).T)
    iou_sum, iou_cnt = torch.zeros(num_pd, num_gt), torch.zeros(num_pd, num_gt)
    target_ids = target_instance_ids.tolist()
    records = []
    for view in group_views:
        extrinsic = view["extrinsic_matrix"] @ to_world
        boxes_3d = homogeneous @ extrinsic.T
        boxes_3d = boxes_3d[..., :-1] / boxes_3d[..., -1:]
        boxes_2d = torch.stack([vsrd.operations.project_box_3d(box_3d=box, line_indices=LINE_INDICES,
                                                               intrinsic_matrix=view["intrinsic_matrix"]) for box in boxes_3d])
        boxes_2d = torchvision.ops.clip_boxes_to_image(boxes_2d.flatten(-2, -1), tuple(image_size)).unflatten(-1, (2, 2))
        iou = torch.nan_to_num(torchvision.ops.box_iou(boxes_2d.flatten(-2, -1), view["boxes_2d"].flatten(-2, -1)))
        columns = torch.tensor([target_ids.index(i) if i in target_ids else -1 for i in view["instance_ids"].tolist()],
                               dtype=torch.long)
        known = columns >= 0
        iou_sum[:, columns[known]] += iou[:, known]
        iou_cnt[:, columns[known]] += 1
        records.append(dict(boxes_3d=boxes_3d, boxes_2d=boxes_2d))
    mean_iou = iou_sum / iou_cnt                        # NaN where an instance is annotated in no view, as upstream
    rows, cols = scipy.optimize.linear_sum_assignment(mean_iou.numpy(), maximize=True)
    confidences = mean_iou[torch.from_numpy(rows), torch.from_numpy(cols)]
    return [{"boxes_3d": {class_name: r["boxes_3d"].tolist()}, "boxes_2d": {class_name: r["boxes_2d"].tolist()},
             "confidences": {class_name: confidences.tolist()}} for r in records]
