"""tests/golden/predictions.npz: the arithmetic of tools/kitti_360/make_predictions.py:58-192 restated line by line
around the reference's own `BoxParameters3D`, `project_box_3d`, `rotation_matrix_x`, `expand_to_4x4` (imported
unmodified; the tool itself imports pycocotools at module level and cannot be imported here) on a seeded synthetic
group.  Build container only."""
import importlib
import os
import sys

import numpy as np
import scipy.optimize
import torch
import torch.nn as nn
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402

LINE_INDICES = [[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]]


def synthetic_group(seed=0, num_instances=5, num_views=6):
    """Detector state + a group of views with annotated 2D boxes (some instances missing / extra per view)."""
    from vsrd_b200 import synthetic
    frame = synthetic.make_frame(num_instances=num_instances, num_views=num_views, seed=seed)
    sup = synthetic.frame_supervision(frame)
    gen = torch.Generator().manual_seed(seed)
    raw = synthetic.perturbed_raw_parameters(frame, seed=seed, position_noise=0.2, yaw_noise=0.05)
    tilt = 0.03                                                     # a target camera that is not level
    c, s = np.cos(tilt), np.sin(tilt)
    level = torch.tensor([[1.0, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1.0]], dtype=torch.float32)
    views = []
    for v in range(num_views):
        vis = torch.nonzero(sup.visible[v]).squeeze(-1)
        vis = vis[torch.randperm(vis.numel(), generator=gen)]
        if v % 2 and vis.numel() > 1:
            vis = vis[:-1]                                          # an instance not annotated in this view
        ids = vis + 100
        boxes = sup.boxes_2d[v, vis].reshape(-1, 2, 2) + torch.randn(vis.numel(), 2, 2, generator=gen) * 2.0
        if v == 1:                                                  # an instance the target frame does not know
            ids = torch.cat([ids, torch.tensor([999])])
            boxes = torch.cat([boxes, torch.tensor([[[10.0, 10.0], [50.0, 40.0]]])])
        views.append(dict(intrinsic_matrix=frame.intrinsics[v], extrinsic_matrix=level @ frame.extrinsics[v] @ torch.linalg.inv(level) @ level,
                          boxes_2d=boxes, instance_ids=ids))
    target = num_views // 2
    return raw, views, views[target]["extrinsic_matrix"], torch.arange(num_instances) + 100, frame.image_size


if __name__ == "__main__":
    raw, views, target_extrinsic, target_ids, image_size = synthetic_group()
    with ref_import.reference_modules() as ref:
        ops = ref.geometric_operations
        model = ref.box_parameters.BoxParameters3D(batch_size=1, num_instances=raw[0].shape[0], num_features=256)
        with torch.no_grad():
            model.locations.copy_(raw[0][None]); model.dimensions.copy_(raw[1][None]); model.orientations.copy_(raw[2][None])
        state = {k: v.clone() for k, v in model.state_dict().items()}
        world, = model()["boxes_3d"]
        world = nn.functional.pad(world.detach(), (0, 1), mode="constant", value=1.0)                 # :61-62
        inverse_target = torch.linalg.inv(target_extrinsic)                                             # :70
        x_axis, y_axis, _ = target_extrinsic[..., :3, :3]                                                # :72-77
        angle = torch.acos(torch.dot(torch.round(y_axis), y_axis)) * torch.sign(torch.dot(torch.cross(torch.round(y_axis), y_axis), x_axis))
        rect = ops.rotation_matrix_x(angle)
        acc_iou, acc_cnt = torch.zeros(len(world), len(target_ids)), torch.zeros(len(world), len(target_ids))
        out = {}
        for k, view in enumerate(views):
            extrinsic = view["extrinsic_matrix"] @ inverse_target @ ops.expand_to_4x4(rect.T)         # :108-112
            pd3 = world @ extrinsic.T
            pd3 = pd3[..., :-1] / pd3[..., -1:]
            pd2 = torch.stack([ops.project_box_3d(box_3d=b, line_indices=LINE_INDICES, intrinsic_matrix=view["intrinsic_matrix"]) for b in pd3])
            pd2 = torchvision.ops.clip_boxes_to_image(pd2.flatten(-2, -1), image_size).unflatten(-1, (2, 2))   # :136-139
            iou = torch.nan_to_num(torchvision.ops.box_iou(pd2.flatten(-2, -1), view["boxes_2d"].flatten(-2, -1)))
            idx = view["instance_ids"].new_tensor([target_ids.tolist().index(i.item()) if i in target_ids else -1
                                                   for i in view["instance_ids"]])                        # :152-156
            acc_iou[..., idx[idx >= 0]] += iou[..., idx >= 0]
            acc_cnt[..., idx[idx >= 0]] += 1
            out[f"boxes_3d_{k}"], out[f"boxes_2d_{k}"] = pd3.numpy(), pd2.numpy()
        mean = acc_iou / acc_cnt
        rows, cols = scipy.optimize.linear_sum_assignment(mean.numpy(), maximize=True)                  # :190
        out["confidences"] = mean[torch.from_numpy(rows), torch.from_numpy(cols)].numpy()
    out.update({f"state.{k}": v.numpy() for k, v in state.items()})
    np.savez(os.path.join(HERE, "predictions.npz"), **out)
    print("confidences", out["confidences"])
