"""`vsrd.visualization` — the four drawing helpers scripts/main.py calls at `image_intervals`
(main.py:975-1069; reference: visualization/drawers.py).  Images are [3,H,W] tensors (float in [0,1] or uint8); the
OpenCV drawing runs on the host exactly as in the reference (it is logging, not the hot path)."""
import contextlib

import cv2 as cv
import numpy as np
import torch

from .. import operations
from .. import utils


@contextlib.contextmanager
def _canvas(image):
    """[3,H,W] ndarray (float or uint8) -> a contiguous uint8 HWC canvas to draw on -> back, in a one-slot list."""
    is_float = image.dtype.kind == "f"
    hwc = np.ascontiguousarray(np.transpose(image, (1, 2, 0)))
    if is_float:
        hwc = np.ascontiguousarray(np.round(np.clip(hwc, 0.0, 1.0) * 255.0).astype(np.uint8))
    slot = [hwc]
    yield slot
    out = np.transpose(slot[0], (2, 0, 1))
    slot[0] = (out.astype(np.float32) / 255.0) if is_float else out


def _point(p):
    return tuple(int(v) for v in p)


@utils.torch_function
def draw_boxes_3d(image, boxes_3d, line_indices, intrinsic_matrix, *args, **kwargs):
    """Wireframes of camera-frame boxes [B,8,3]; edges are clipped to the half space in front of the camera."""
    with _canvas(image) as slot:
        for box_3d in boxes_3d:
            lines, visible = utils.numpy_function(operations.clip_lines_to_front)(box_3d[np.asarray(line_indices)])
            pixels = lines @ intrinsic_matrix.T
            pixels = pixels[..., :-1] / np.clip(pixels[..., -1:], 1e-3, None)
            for start, end in pixels[visible]:
                slot[0] = cv.line(slot[0], _point(start), _point(end), *args, **kwargs)
    return slot[0]


@utils.torch_function
def draw_boxes_bev(image, boxes_3d, extents=((-50.0, 100.0), (50.0, 0.0)), *args, **kwargs):
    """Bird's-eye footprints (mean of the top and bottom faces, x-z plane) on a 10x10 grid."""
    with _canvas(image) as slot:
        height, width = slot[0].shape[:2]
        footprints = np.mean(np.reshape(boxes_3d, (-1, 2, 4, 3)), axis=1)[..., [0, 2]]
        footprints = (footprints - extents[0]) / -np.subtract(*extents) * (width, height)
        for corners in footprints:
            for start, end in zip(corners, np.roll(corners, -1, axis=0)):
                slot[0] = cv.line(slot[0], _point(start), _point(end), *args, **kwargs)
        for y in range(0, height, max(height // 10, 1)):
            slot[0] = cv.line(slot[0], (0, y), (width, y), color=(128, 128, 128))
        for x in range(0, width, max(width // 10, 1)):
            slot[0] = cv.line(slot[0], (x, 0), (x, height), color=(128, 128, 128))
    return slot[0]


@utils.torch_function
def draw_boxes_2d(image, boxes_2d, *args, **kwargs):
    with _canvas(image) as slot:
        for top_left, bottom_right in boxes_2d:
            slot[0] = cv.rectangle(slot[0], _point(top_left), _point(bottom_right), *args, **kwargs)
    return slot[0]


@utils.torch_function
def draw_points_2d(image, points_2d, *args, **kwargs):
    with _canvas(image) as slot:
        for point in points_2d:
            slot[0] = cv.circle(slot[0], _point(point), *args, **kwargs)
    return slot[0]


def draw_masks(image, masks, colors=None, weight=0.5):
    """Blend per-instance masks [N,H,W] into the image with (random, max-normalised) colours."""
    is_uint8 = image.dtype == torch.uint8
    if is_uint8:
        image = image.float() / 255.0
    if colors is None:
        colors = torch.rand(len(masks), 3).to(masks)
        colors = colors / colors.max(dim=-1, keepdim=True).values
    overlay = torch.einsum("nhw,nc->chw", masks, colors.to(masks))
    image = torch.clamp(image + overlay * weight, 0.0, 1.0)
    return (image * 255.0).byte() if is_uint8 else image
