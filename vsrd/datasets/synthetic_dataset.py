"""Synthetic KITTI-360-shaped frames in the dict format of `KITTI360Dataset.__getitem__`
(datasets/kitti_360_dataset.py:128-168, 175-250) after the config's transform chain (Resizer ... BoxGenerator,
SoftRasterizer: transforms/geometric_transforms.py:139-177, 233-317).

One item = `{relative_index: view_dict}` for the target frame (0) and its source frames, every view holding

    image [3,H,W]  masks / hard_masks / soft_masks [N_v,H,W]  labels [N_v]  boxes_3d [N_v,8,3] (camera frame)
    boxes_2d [N_v,2,2]  instance_ids [N_v]  intrinsic_matrix [3,3]  extrinsic_matrix [4,4] (world = rectified target
    camera -> this camera)  rectification_matrix [3,3]  filename

Only the instances visible in a view are listed there (N_v <= N), in a per-view shuffled order, so main.py's
instance-id association (main.py:204-262) is exercised.  Soft masks are rasterised by the `soft_masks` kernel when a
CUDA device is present and by the same formula in PyTorch otherwise.
"""
import functools
import os

import torch

from vsrd_b200 import synthetic


class SyntheticKITTI360Dataset(torch.utils.data.Dataset):

    def __init__(self, num_frames=8, num_source_frames=16, image_size=synthetic.KITTI360_IMAGE_SIZE,
                 intrinsics_scale=1.0, mean_instances=6.0, min_instances=1, max_instances=24, num_instances=None,
                 layout="street", seed=0, soft_mask_temperature=10.0, root_dirname="datasets/SYNTHETIC-360",
                 sequence="2013_05_28_drive_0000_sync", shuffle_instances=True, device=None):
        super().__init__()
        self.num_frames = int(num_frames)
        self.num_views = int(num_source_frames) + 1
        self.image_size = tuple(image_size)
        self.intrinsics_scale = float(intrinsics_scale)
        self.mean_instances, self.min_instances, self.max_instances = float(mean_instances), int(min_instances), int(max_instances)
        self.fixed_instances = num_instances
        self.layout, self.seed = layout, int(seed)
        self.soft_mask_temperature = float(soft_mask_temperature)
        self.root_dirname, self.sequence = root_dirname, sequence
        self.shuffle_instances = shuffle_instances
        self.device = device

    # ---- the static helpers main.py / the tools call on the dataset (kitti_360_dataset.py:51-59) -------------------
    @staticmethod
    def get_root_dirname(image_filename):
        return functools.reduce(lambda x, f: f(x), [os.path.dirname] * 5, image_filename)

    @staticmethod
    def get_sequence_dirname(image_filename):
        return functools.reduce(lambda x, f: f(x), [os.path.dirname] * 3, image_filename)

    def __len__(self):
        return self.num_frames

    def frame_filename(self, index, relative_index=0):
        return os.path.join(self.root_dirname, "data_2d_raw", self.sequence, "image_00", "data_rect",
                            f"{index * 10 + relative_index + 1000:010}.png")

    def num_instances_of(self, index):
        """N ~ Poisson(mean) clipped to [min, max], seeded by the frame id."""
        if self.fixed_instances is not None:
            return int(self.fixed_instances)
        gen = torch.Generator().manual_seed(self.seed * 100003 + index)
        n = int(torch.poisson(torch.tensor([self.mean_instances]), generator=gen))
        return max(self.min_instances, min(self.max_instances, n))

    def scene(self, index):
        n = self.num_instances_of(index)
        return synthetic.make_frame(num_instances=n, num_views=self.num_views, image_size=self.image_size,
                                    seed=self.seed * 100003 + index, layout=self.layout,
                                    intrinsics_scale=self.intrinsics_scale)

    def _soft_masks(self, sup):
        """[V,H,W,N] sigmoid(signed polygon distance / temperature); zero polygons give all-zero masks."""
        device = self.device if self.device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        if torch.device(device).type == "cuda":
            from vsrd_b200 import ops
            return ops.soft_masks(sup.polygons.to(device), sup.polygon_sizes.to(device), self.image_size,
                                  self.soft_mask_temperature).cpu()
        h, w = self.image_size
        v, n = sup.polygon_sizes.shape
        out = torch.zeros(v, h, w, n)
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        pixels = torch.stack([xs, ys], dim=-1).reshape(-1, 1, 2)
        for vi in range(v):
            for ni in range(n):
                k = int(sup.polygon_sizes[vi, ni])
                if k < 3:
                    continue
                a = sup.polygons[vi, ni, :k]
                b = torch.roll(a, shifts=-1, dims=0)
                side, rel = (b - a)[None], pixels - a[None]
                t = ((side * rel).sum(-1, keepdim=True) / ((side * side).sum(-1, keepdim=True) + 1e-6)).clamp(0.0, 1.0)
                dist = torch.linalg.norm(rel - side * t, dim=-1).min(dim=-1).values.reshape(h, w)
                crossing = ((a[:, 1] > ys[..., None]) != (b[:, 1] > ys[..., None]))
                slope = (b[:, 0] - a[:, 0]) / torch.where(b[:, 1] == a[:, 1], torch.ones(()), b[:, 1] - a[:, 1])
                inside = (crossing & (xs[..., None] < a[:, 0] + (ys[..., None] - a[:, 1]) * slope)).sum(-1) % 2 == 1
                out[vi, :, :, ni] = torch.sigmoid(torch.where(inside, dist, -dist) / self.soft_mask_temperature)
        return out

    def __getitem__(self, index):
        frame = self.scene(index)
        sup = synthetic.frame_supervision(frame)
        soft = self._soft_masks(sup)                                           # [V,H,W,N]
        corners = synthetic.gt_corners(frame)                                  # [N,8,3] world
        h, w = self.image_size
        gen = torch.Generator().manual_seed(self.seed * 7919 + index)
        target = sup.target_view
        ramp = torch.linspace(0.2, 0.8, w).expand(3, h, w)
        multi_inputs = {}
        for view in range(frame.num_views):
            relative_index = view - target
            visible = torch.nonzero(sup.visible[view]).squeeze(-1)
            if view == target:
                pass                                                           # main.py takes N (and the order) from here
            elif self.shuffle_instances and visible.numel() > 1:
                visible = visible[torch.randperm(visible.numel(), generator=gen)]
            e = frame.extrinsics[view]
            soft_v = soft[view].permute(2, 0, 1)[visible].contiguous()         # [N_v,H,W]
            hard_v = (soft_v > 0.5).float()
            boxes_2d = sup.boxes_2d[view, visible].reshape(-1, 2, 2)
            multi_inputs[relative_index] = dict(
                image=(ramp * (0.9 + 0.1 * view / max(frame.num_views - 1, 1))).contiguous(),
                masks=hard_v,
                hard_masks=hard_v,
                soft_masks=soft_v,
                labels=torch.zeros(visible.numel(), dtype=torch.long),
                boxes_3d=(corners[visible] @ e[:3, :3].T + e[:3, 3]),
                boxes_2d=boxes_2d,
                instance_ids=(visible + 26001).long(),
                intrinsic_matrix=frame.intrinsics[view].clone(),
                extrinsic_matrix=e.clone(),
                rectification_matrix=torch.eye(3),
                filename=self.frame_filename(index, relative_index),
            )
        return dict(sorted(multi_inputs.items()))
