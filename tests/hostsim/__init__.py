"""TEST INFRASTRUCTURE ONLY: host build of the per-thread device math (see hostsim.cpp)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libhostsim.so")
GRAD_STRIDE = 1632
NUM_W = 1617


def build(force=False):
    src = os.path.join(HERE, "hostsim.cpp")
    hdrs = [os.path.join(HERE, "..", "..", "vsrd_b200", "csrc", h) for h in ("vsrd_math.cuh", "vsrd_frame_math.cuh")]
    os.makedirs(BUILD, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) > max(os.path.getmtime(src), *map(os.path.getmtime, hdrs))):
        return LIB
    # -ffp-contract=off: keep the host arithmetic un-fused so it is a clean fp32 restatement
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++", src, "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(t):
    t = t.detach().to(torch.float32).contiguous()
    return t, ctypes.c_void_p(t.data_ptr())


def field_forward(x, loc, rot, dim, w, scale=100.0):
    x, px = _f(x); loc, pl = _f(loc); rot, pr = _f(rot); dim, pd = _f(dim)
    pw = None
    if w is not None:
        w, pw = _f(w)
    out = torch.empty(x.shape[0], 4)
    lib().hs_field_forward(px, pl, pr, pd, pw, ctypes.c_float(scale), ctypes.c_int(x.shape[0]),
                           ctypes.c_void_p(out.data_ptr()))
    return out


def field_backward(x, loc, rot, dim, w, adj, scale=100.0):
    x, px = _f(x); loc, pl = _f(loc); rot, pr = _f(rot); dim, pd = _f(dim); adj, pa = _f(adj)
    pw = None
    if w is not None:
        w, pw = _f(w)
    grads = torch.zeros(GRAD_STRIDE, dtype=torch.float64)
    lib().hs_field_backward(px, pl, pr, pd, pw, ctypes.c_float(scale), ctypes.c_int(x.shape[0]), pa,
                            ctypes.c_void_p(grads.data_ptr()))
    return dict(mlp_weights=grads[:NUM_W], locations=grads[NUM_W:NUM_W + 3],
                half_extents=grads[NUM_W + 3:NUM_W + 6], rotations=grads[NUM_W + 6:NUM_W + 15].reshape(3, 3))


def composite_forward(t, dirs, field, T, sigma, rho, eps=1e-6):
    """t [R,M+1], dirs [R,3], field [N,R,M,4] -> labels [R,N], grads [R,M,3], weights [R,M]."""
    t, pt = _f(t); dirs, pdir = _f(dirs); field, pf = _f(field)
    N, R, M, _ = field.shape
    labels = torch.empty(R, N); grads = torch.empty(R, M, 3); weights = torch.empty(R, M)
    lib().hs_composite_forward(pt, pdir, pf, R, M, N, ctypes.c_float(T), ctypes.c_float(sigma),
                               ctypes.c_float(rho), ctypes.c_float(eps),
                               ctypes.c_void_p(labels.data_ptr()), ctypes.c_void_p(grads.data_ptr()),
                               ctypes.c_void_p(weights.data_ptr()))
    return labels, grads, weights


def composite_backward(t, dirs, field, T, sigma, rho, gl=None, gg=None, gw=None, eps=1e-6):
    t, pt = _f(t); dirs, pdir = _f(dirs); field, pf = _f(field)
    N, R, M, _ = field.shape
    ptrs = []
    keep = []
    for g in (gl, gg, gw):
        if g is None:
            ptrs.append(None)
        else:
            g, p = _f(g)
            keep.append(g)
            ptrs.append(p)
    adj = torch.empty(N, R, M, 4)
    lib().hs_composite_backward(pt, pdir, pf, R, M, N, ctypes.c_float(T), ctypes.c_float(sigma),
                                ctypes.c_float(rho), ctypes.c_float(eps), ptrs[0], ptrs[1], ptrs[2],
                                ctypes.c_void_p(adj.data_ptr()))
    return adj


def projection_step(extrinsics, intrinsics, world_boxes, image_size, gt_boxes=None, visible=None, gt_indices=None,
                    target_view=0):
    """Serial host build of projection_step_kernel's math.  Returns boxes [V,N,4] or
    (boxes, cost [N,N] on the target view, losses [2], grad_world [2,N,8,3])."""
    e, pe = _f(extrinsics); k, pk = _f(intrinsics); w, pw = _f(world_boxes)
    v, n = e.shape[0], w.shape[0]
    boxes = torch.empty(v, n, 4)
    h, wd = image_size
    if gt_boxes is None:
        lib().hs_projection_step(pe, pk, pw, None, None, None, v, n, ctypes.c_float(h), ctypes.c_float(wd),
                                 ctypes.c_void_p(boxes.data_ptr()), None, None, None)
        return boxes
    # run once to get the boxes, then the cost on the target view, then losses with the given assignment
    gt, pg = _f(gt_boxes)
    vis = None if visible is None else visible.to(torch.uint8).contiguous()
    idx = gt_indices.to(torch.int64).contiguous()
    cost = torch.empty(n, n); losses = torch.empty(2); grad = torch.empty(2, n, 8, 3)
    lib().hs_projection_step(pe, pk, pw, pg, None if vis is None else ctypes.c_void_p(vis.data_ptr()),
                             ctypes.c_void_p(idx.data_ptr()), v, n, ctypes.c_float(h), ctypes.c_float(wd),
                             ctypes.c_void_p(boxes.data_ptr()), None, ctypes.c_void_p(losses.data_ptr()),
                             ctypes.c_void_p(grad.data_ptr()))
    # cost matrix of the target view: call the 1-view variant on that view's slices
    tb = boxes[target_view].contiguous(); tg = gt[target_view].contiguous()
    e1 = e[target_view:target_view + 1].contiguous(); k1 = k[target_view:target_view + 1].contiguous()
    l1 = torch.empty(2); g1 = torch.empty(2, n, 8, 3); b1 = torch.empty(1, n, 4)
    lib().hs_projection_step(ctypes.c_void_p(e1.data_ptr()), ctypes.c_void_p(k1.data_ptr()), pw,
                             ctypes.c_void_p(tg.data_ptr()), None, ctypes.c_void_p(idx.data_ptr()), 1, n,
                             ctypes.c_float(h), ctypes.c_float(wd), ctypes.c_void_p(b1.data_ptr()),
                             ctypes.c_void_p(cost.data_ptr()), ctypes.c_void_p(l1.data_ptr()),
                             ctypes.c_void_p(g1.data_ptr()))
    return boxes, cost, losses, grad


def soft_mask(polygon, image_size, temperature=10.0):
    p, pp = _f(polygon)
    h, w = image_size
    sd = torch.empty(h, w); mask = torch.empty(h, w)
    lib().hs_soft_mask(pp, p.shape[0], h, w, ctypes.c_float(temperature), ctypes.c_void_p(sd.data_ptr()),
                       ctypes.c_void_p(mask.data_ptr()))
    return sd, mask
