"""3D / BEV IoU of two yaw-only boxes given as 8 corners (contract of
vsrd/operations/kitti360_operations.py:84-114: corners [8,3] with Z up, corners 0-3 the top face in
order, 4-7 the bottom face).  CPU metric (numpy); returns (iou_3d, iou_bev)."""
import numpy as np
import torch


def _polygon_area(poly):
    x, y = poly[:, 0], poly[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1))))


def _ccw(poly):
    x, y = poly[:, 0], poly[:, 1]
    signed = float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    return poly if signed >= 0 else poly[::-1]


def _clip_convex(subject, clip):
    """Sutherland-Hodgman: intersect polygon `subject` with convex CCW polygon `clip`."""
    out = [tuple(p) for p in subject]
    for k in range(len(clip)):
        a, b = clip[k - 1], clip[k]
        edge = (b[0] - a[0], b[1] - a[1])
        side = lambda p: edge[0] * (p[1] - a[1]) - edge[1] * (p[0] - a[0])
        inp, out = out, []
        if not inp:
            return np.zeros((0, 2))
        prev = inp[-1]
        for cur in inp:
            sp, sc = side(prev), side(cur)
            if (sc >= 0) != (sp >= 0):
                t = sp / (sp - sc)
                out.append((prev[0] + t * (cur[0] - prev[0]), prev[1] + t * (cur[1] - prev[1])))
            if sc >= 0:
                out.append(cur)
            prev = cur
    return np.asarray(out).reshape(-1, 2)


def _box_3d_iou_numpy(corners1, corners2):
    c1, c2 = np.asarray(corners1, dtype=np.float64), np.asarray(corners2, dtype=np.float64)
    r1, r2 = _ccw(c1[:4, :2]), _ccw(c2[:4, :2])
    a1, a2 = _polygon_area(r1), _polygon_area(r2)
    inter_poly = _clip_convex(r1, r2)
    inter = min(_polygon_area(inter_poly) if len(inter_poly) >= 3 else 0.0, a1, a2)
    iou_bev = inter / max(a1 + a2 - inter, 1e-12)
    top = min(max(c1[0, 2], c1[4, 2]), max(c2[0, 2], c2[4, 2]))
    bottom = max(min(c1[0, 2], c1[4, 2]), min(c2[0, 2], c2[4, 2]))
    inter_vol = inter * max(0.0, top - bottom)
    v1 = a1 * abs(c1[0, 2] - c1[4, 2])
    v2 = a2 * abs(c2[0, 2] - c2[4, 2])
    return inter_vol / max(v1 + v2 - inter_vol, 1e-12), iou_bev


def box_3d_iou(corners1, corners2):
    to_np = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t
    iou, iou_bev = _box_3d_iou_numpy(to_np(corners1), to_np(corners2))
    return torch.as_tensor(iou), torch.as_tensor(iou_bev)
