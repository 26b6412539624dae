"""`BoxParameters3D` — API of vsrd/models/detectors/box_parameters.py:16-146.

Learnable raw parameters `locations` / `dimensions` / `orientations` / `embeddings` of shape
[B, N, {3, 3, 2, F}] with buffers `location_range` / `dimension_range`; `forward()` returns the
decoded dict main.py consumes (`boxes_3d`, `locations`, `dimensions`, `orientations`, `embeddings`)."""
import torch
import torch.nn as nn

_CAMERA_HEIGHT = 1.55
_CAR_HEIGHT = 1.75
# corner order of the KITTI-360 evaluation format (box_parameters.py:77-86)
_CORNERS = (
    (-1.0, -1.0, +1.0), (+1.0, -1.0, +1.0), (+1.0, -1.0, -1.0), (-1.0, -1.0, -1.0),
    (-1.0, +1.0, +1.0), (+1.0, +1.0, +1.0), (+1.0, +1.0, -1.0), (-1.0, +1.0, -1.0),
)


_corner_cache = {}


def _corner_signs(like):
    """The corner sign table on `like`'s device/dtype, uploaded once (a host->device copy per call would
    break CUDA-graph capture of the optimisation step)."""
    key = (like.device, like.dtype)
    if key not in _corner_cache:
        _corner_cache[key] = torch.tensor(_CORNERS, device=like.device, dtype=like.dtype)
    return _corner_cache[key]


def rotation_matrix_y(cos, sin):
    zero, one = torch.zeros_like(cos), torch.ones_like(cos)
    rows = [
        torch.stack([cos, zero, sin], dim=-1),
        torch.stack([zero, one, zero], dim=-1),
        torch.stack([-sin, zero, cos], dim=-1),
    ]
    return torch.stack(rows, dim=-2)


class BoxParameters3D(nn.Module):

    def __init__(
        self,
        batch_size,
        num_instances,
        num_features=256,
        location_range=(
            (-50.0, _CAMERA_HEIGHT - _CAR_HEIGHT / 2.0 - 5.0, 0.0),
            (+50.0, _CAMERA_HEIGHT - _CAR_HEIGHT / 2.0 + 5.0, 100.0),
        ),
        dimension_range=((0.75, 0.75, 1.5), (1.00, 1.00, 2.5)),
    ):
        super().__init__()
        shape = (batch_size, num_instances)
        self.locations = nn.Parameter(torch.zeros(*shape, 3))
        self.dimensions = nn.Parameter(torch.zeros(*shape, 3))
        self.orientations = nn.Parameter(torch.tensor([1.0, 0.0]).repeat(*shape, 1))
        # one random embedding shared by every instance at init (box_parameters.py:46-49)
        self.embeddings = nn.Parameter(torch.rand(num_features).repeat(*shape, 1))
        self.register_buffer("location_range", torch.as_tensor(location_range, dtype=torch.float32))
        self.register_buffer("dimension_range", torch.as_tensor(dimension_range, dtype=torch.float32))

    def decode_location(self, locations):
        return torch.lerp(self.location_range[0], self.location_range[1], torch.sigmoid(locations))

    def decode_dimension(self, dimensions):
        return torch.lerp(self.dimension_range[0], self.dimension_range[1], torch.sigmoid(dimensions))

    def decode_orientation(self, orientations):
        unit = nn.functional.normalize(orientations, dim=-1)
        return rotation_matrix_y(unit[..., 0], unit[..., 1])

    @staticmethod
    def decode_box_3d(locations, dimensions, orientations):
        corners = _corner_signs(dimensions) * dimensions.unsqueeze(-2)
        return corners @ orientations.transpose(-2, -1) + locations.unsqueeze(-2)

    @staticmethod
    def encode_box_3d(boxes_3d):
        def mean_edge(a, b):
            return (boxes_3d[..., a, :] - boxes_3d[..., b, :]).norm(dim=-1).mean(dim=-1)

        locations = boxes_3d.mean(dim=-2)
        widths = mean_edge([1, 2, 6, 5], [0, 3, 7, 4])
        heights = mean_edge([4, 5, 6, 7], [0, 1, 2, 3])
        lengths = mean_edge([1, 0, 4, 5], [2, 3, 7, 6])
        dimensions = torch.stack([widths, heights, lengths], dim=-1) / 2.0
        forward = (boxes_3d[..., [1, 0, 4, 5], :] - boxes_3d[..., [2, 3, 7, 6], :]).mean(dim=-2)
        unit = nn.functional.normalize(forward[..., [2, 0]], dim=-1)
        return locations, dimensions, rotation_matrix_y(unit[..., 0], unit[..., 1])

    def _box_ranges(self):
        """`location_range` / `dimension_range` as the C struct (read back from the buffers once)."""
        from vsrd_b200 import _lib
        key = (self.location_range._version, self.dimension_range._version, self.location_range.data_ptr())
        cached = getattr(self, "_ranges_cache", None)
        if cached is None or cached[0] != key:
            ranges = _lib.VsrdBoxRanges()
            lo, hi = self.location_range.detach().cpu().tolist()
            dlo, dhi = self.dimension_range.detach().cpu().tolist()
            for k in range(3):
                ranges.location_min[k], ranges.location_max[k] = lo[k], hi[k]
                ranges.dimension_min[k], ranges.dimension_max[k] = dlo[k], dhi[k]
            cached = (key, ranges)
            object.__setattr__(self, "_ranges_cache", cached)
        return cached[1]

    def forward(self):
        rows = self.locations.numel() // 3
        fused = self.locations.is_cuda and self.locations.dtype == torch.float32 and 1 <= rows <= 32
        if fused:
            from vsrd_b200 import functional
            fused = functional.fused_modules()
        if fused:
            # one launch forward, one backward (csrc/vsrd_model.cu) instead of ~40 ATen launches
            locations, dimensions, orientations, boxes_3d = functional.decode_boxes(
                self.locations, self.dimensions, self.orientations, self._box_ranges())
            return dict(boxes_3d=boxes_3d, locations=locations, dimensions=dimensions, orientations=orientations,
                        embeddings=self.embeddings)
        locations = self.decode_location(self.locations)
        dimensions = self.decode_dimension(self.dimensions)
        orientations = self.decode_orientation(self.orientations)
        return dict(
            boxes_3d=self.decode_box_3d(locations, dimensions, orientations),
            locations=locations,
            dimensions=dimensions,
            orientations=orientations,
            embeddings=self.embeddings,
        )
