"""3D / BEV IoU of two yaw-only boxes given as 8 corners (vsrd/operations/kitti360_operations.py: corners [8,3] with Z
up, corners 0-3 the top face in order, 4-7 the bottom face).  CPU metrics (numpy); both return (iou_3d, iou_bev).

    box_3d_iou        VALUE-IDENTICAL to the reference's `box3dIou` (kitti360_operations.py:84-114), which scripts/main.py
                      logs as metrics/iou_3d (main.py:892-899) and tools/ report.  That function is the classic
                      Sutherland-Hodgman clip with two quirks, reproduced here on purpose: the edge-intersection
                      denominator carries a `+ 0.01` fudge (:29), which misplaces intersection vertices by up to a few
                      centimetres on metre-sized boxes (up to 0.05 IoU on ordinary pairs, far more on near-coincident
                      ones), and the intersection area is the area of the convex hull of the clipped vertices
                      (scipy.spatial.ConvexHull(...).volume, :70).  Pinned to the reference on 300 random pairs
                      (tests/golden/box_iou.npz).
    box_3d_iou_exact  the same contract with an exact clip and the shoelace area: what the parity gates of this
                      repository use ("boxes agree to >= 0.99 3D IoU"), since a gate should not inherit the fudge.
"""
import numpy as np
import scipy.spatial
import torch


# ---- reference-identical metric ---------------------------------------------------------------------------------------

def _shoelace(x, y):
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


def _clip_with_fudge(subject, clip):
    """Sutherland-Hodgman clip of `subject` by convex CCW `clip` in the reference's arithmetic (scalar type preserved:
    float32 corners stay float32), including the `+ 0.01` in the intersection denominator.  None when empty."""
    kept = subject
    start = clip[-1]
    for end in clip:
        def is_inside(p):
            return (end[0] - start[0]) * (p[1] - start[1]) > (end[1] - start[1]) * (p[0] - start[0])

        def crossing(s, e):
            dc = [start[0] - end[0], start[1] - end[1]]
            dp = [s[0] - e[0], s[1] - e[1]]
            n1 = start[0] * end[1] - start[1] * end[0]
            n2 = s[0] * e[1] - s[1] * e[0]
            n3 = 1.0 / (dc[0] * dp[1] - dc[1] * dp[0] + 0.01)
            return [(n1 * dp[0] - n2 * dc[0]) * n3, (n1 * dp[1] - n2 * dc[1]) * n3]

        previous, kept = kept, []
        s = previous[-1]
        for e in previous:
            if is_inside(e):
                if not is_inside(s):
                    kept.append(crossing(s, e))
                kept.append(e)
            elif is_inside(s):
                kept.append(crossing(s, e))
            s = e
        start = end
        if not kept:
            return None
    return kept


def _box_3d_iou_reference(corners1, corners2):
    top1 = [(corners1[i, 0], corners1[i, 1]) for i in (3, 2, 1, 0)]
    top2 = [(corners2[i, 0], corners2[i, 1]) for i in (3, 2, 1, 0)]
    area1 = _shoelace(np.array(top1)[:, 0], np.array(top1)[:, 1])
    area2 = _shoelace(np.array(top2)[:, 0], np.array(top2)[:, 1])
    clipped = _clip_with_fudge(top1, top2)
    overlap = scipy.spatial.ConvexHull(clipped).volume if clipped is not None else 0.0
    overlap = min(min(area1, area2), overlap)
    iou_bev = overlap / (area1 + area2 - overlap)
    z_top = min(corners1[0, 2], corners2[0, 2])
    z_bottom = max(corners1[4, 2], corners2[4, 2])
    shared = overlap * max(0.0, z_top - z_bottom)

    def volume(c):
        edge = lambda i, j: np.sqrt(np.sum((c[i, :] - c[j, :]) ** 2))
        return edge(0, 1) * edge(1, 2) * edge(0, 4)

    return shared / (volume(corners1) + volume(corners2) - shared), iou_bev


def _as_numpy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def box_3d_iou(corners1, corners2):
    # like the reference's `utils.torch_function(box3dIou)`: tensors in, the two numpy scalars out
    return _box_3d_iou_reference(_as_numpy(corners1), _as_numpy(corners2))


# ---- exact metric (the gate of the parity tests) ----------------------------------------------------------------------

def _polygon_area(poly):
    x, y = poly[:, 0], poly[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1))))


def _ccw(poly):
    x, y = poly[:, 0], poly[:, 1]
    signed = float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    return poly if signed >= 0 else poly[::-1]


def _clip_convex(subject, clip):
    """Sutherland-Hodgman: intersect polygon `subject` with convex CCW polygon `clip` (exact parametrisation)."""
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


def _box_3d_iou_exact_numpy(corners1, corners2):
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


def box_3d_iou_exact(corners1, corners2):
    iou, iou_bev = _box_3d_iou_exact_numpy(_as_numpy(corners1), _as_numpy(corners2))
    return torch.as_tensor(iou), torch.as_tensor(iou_bev)
