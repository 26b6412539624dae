// Per-thread math of the frame-level rows around the renderer (compiled for the device by
// vsrd_frame.cu and for the host by tests/hostsim so every formula is checked against autograd on CPU):
//   a14  multi-view box projection      scripts/main.py:339-367, operations/geometric_operations.py:343-389
//   a15  DIoU / smooth-L1 box losses    scripts/main.py:374-415 (torchvision.ops.distance_box_iou[_loss])
//   soft-mask signed distance           transforms/geometric_transforms.py:267-309 (SoftRasterizer)
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef VSRD_HD
#ifdef __CUDACC__
#define VSRD_HD __host__ __device__ __forceinline__
#else
#define VSRD_HD inline
#endif
#endif

namespace vsrd {

// un-fused products/sums where the reference evaluates separate ATen ops (keeps min/max ties stable)
#ifdef __CUDA_ARCH__
#define VSRD_MUL(a, b) __fmul_rn((a), (b))
#define VSRD_ADD(a, b) __fadd_rn((a), (b))
#else
#define VSRD_MUL(a, b) ((a) * (b))
#define VSRD_ADD(a, b) ((a) + (b))
#endif

// scripts/main.py:26-30 LINE_INDICES: 4 top-face edges, 4 bottom-face edges, 4 verticals
VSRD_HD int box_edge_a(int e) { return e < 8 ? e : e - 8; }
VSRD_HD int box_edge_b(int e) { return e < 4 ? ((e + 1) & 3) : (e < 8 ? 4 + ((e + 1) & 3) : e - 4); }

struct BoxProjection {
    float hom[8][4];   // E @ [x y z 1]
    float cam[8][3];   // camera-frame corners (main.py:343-344)
    float raw[4];      // x1 y1 x2 y2 before clipping to the image
    float box[4];      // after clip_boxes_to_image (main.py:359-362)
    int ext[4];        // which of the 24 clipped-edge points defines each side (edge*2 + end), -1 if none
};

struct ClippedEdge {
    int i1, i2;        // corner indices after the depth ordering of clip_lines_to_front (:348-355)
    float p1[3], q2[3];
    float w0, w, den, gap;
    bool front;
};

VSRD_HD void clip_edge(const float cam[8][3], int e, float eps, ClippedEdge& o) {
    const int ia = box_edge_a(e), ib = box_edge_b(e);
    const bool first = cam[ia][2] > cam[ib][2];
    o.i1 = first ? ia : ib;
    o.i2 = first ? ib : ia;
    const float d1 = cam[o.i1][2], d2 = cam[o.i2][2];
    o.gap = d1 - d2;
    o.den = fmaxf(o.gap, eps);
    o.w0 = d1 / o.den;
    o.w = fminf(o.w0, 1.0f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.p1[k] = cam[o.i1][k];
        o.q2[k] = VSRD_ADD(o.p1[k], VSRD_MUL(cam[o.i2][k] - o.p1[k], o.w));
    }
    o.front = o.p1[2] > 0.0f;
}

VSRD_HD void pinhole(const float* K, const float p[3], float eps, float h[3], float uv[2]) {
#pragma unroll
    for (int m = 0; m < 3; ++m) h[m] = p[0] * K[3 * m] + p[1] * K[3 * m + 1] + p[2] * K[3 * m + 2];
    const float hz = fmaxf(h[2], eps);
    uv[0] = h[0] / hz;
    uv[1] = h[1] / hz;
}

// One (view, instance) pair: world corners [8,3] -> clipped 2D box.
// `clip_to_image = false` stops after the min / max over the clipped edge points (geometric_operations.py:343-389 on
// its own; main.py applies torchvision's clip_boxes_to_image separately): box == raw.
VSRD_HD void project_box(const float* E, const float* K, const float* world, float height, float width,
                         float eps, BoxProjection& o, bool clip_to_image = true) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
            o.hom[c][m] = E[4 * m] * world[3 * c] + E[4 * m + 1] * world[3 * c + 1] + E[4 * m + 2] * world[3 * c + 2] + E[4 * m + 3];
#pragma unroll
        for (int k = 0; k < 3; ++k) o.cam[c][k] = o.hom[c][k] / o.hom[c][3];
    }
    float lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
    o.ext[0] = o.ext[1] = o.ext[2] = o.ext[3] = -1;
    for (int e = 0; e < 12; ++e) {
        ClippedEdge ce;
        clip_edge(o.cam, e, eps, ce);
        if (!ce.front) continue;
#pragma unroll
        for (int end = 0; end < 2; ++end) {
            float h[3], uv[2];
            pinhole(K, end ? ce.q2 : ce.p1, eps, h, uv);
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                if (uv[a] < lo[a]) { lo[a] = uv[a]; o.ext[a] = 2 * e + end; }
                if (uv[a] > hi[a]) { hi[a] = uv[a]; o.ext[2 + a] = 2 * e + end; }
            }
        }
    }
    if (o.ext[0] < 0) {   // every edge behind the camera: the reference returns a zero box (:384-387)
        o.raw[0] = o.raw[1] = o.raw[2] = o.raw[3] = 0.0f;
    } else {
        o.raw[0] = lo[0]; o.raw[1] = lo[1]; o.raw[2] = hi[0]; o.raw[3] = hi[1];
    }
    if (!clip_to_image) {
        o.box[0] = o.raw[0]; o.box[1] = o.raw[1]; o.box[2] = o.raw[2]; o.box[3] = o.raw[3];
        return;
    }
    o.box[0] = fminf(fmaxf(o.raw[0], 0.0f), width);
    o.box[1] = fminf(fmaxf(o.raw[1], 0.0f), height);
    o.box[2] = fminf(fmaxf(o.raw[2], 0.0f), width);
    o.box[3] = fminf(fmaxf(o.raw[3], 0.0f), height);
}

// Adjoint of project_box: upstream gradient of the clipped box -> gradient of the world corners [8,3]
// (accumulated into `gworld`).  Mirrors autograd: clamp passes the gradient inside the closed range,
// min/max send it to the selected point, torch.clamp(max=1) on the clip weight blocks it when active.
VSRD_HD void project_box_backward(const float* E, const float* K, const BoxProjection& o, float height, float width,
                                  float eps, const float gbox[4], float gworld[24], bool clip_to_image = true) {
    float gcam[8][3];
#pragma unroll
    for (int c = 0; c < 8; ++c) gcam[c][0] = gcam[c][1] = gcam[c][2] = 0.0f;
    const float bound[4] = {width, height, width, height};
    for (int s = 0; s < 4; ++s) {
        if (o.ext[s] < 0) continue;
        const float g = (!clip_to_image || (o.raw[s] >= 0.0f && o.raw[s] <= bound[s])) ? gbox[s] : 0.0f;
        if (g == 0.0f) continue;
        const int e = o.ext[s] >> 1, end = o.ext[s] & 1, a = s & 1;
        ClippedEdge ce;
        clip_edge(o.cam, e, eps, ce);
        const float* p = end ? ce.q2 : ce.p1;
        float h[3], uv[2];
        pinhole(K, p, eps, h, uv);
        float gh[3] = {0.0f, 0.0f, 0.0f};
        if (h[2] >= eps) {
            gh[a] = g / h[2];
            gh[2] = -g * h[a] / (h[2] * h[2]);
        } else {
            gh[a] = g / eps;
        }
        float gp[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) gp[n] = gh[0] * K[n] + gh[1] * K[3 + n] + gh[2] * K[6 + n];
        if (!end) {
#pragma unroll
            for (int n = 0; n < 3; ++n) gcam[ce.i1][n] += gp[n];
        } else {
            float gw = 0.0f;
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                gcam[ce.i1][n] += gp[n] * (1.0f - ce.w);
                gcam[ce.i2][n] += gp[n] * ce.w;
                gw += gp[n] * (o.cam[ce.i2][n] - ce.p1[n]);
            }
            if (ce.w0 <= 1.0f) {                       // weight not clamped
                const float d1 = ce.p1[2];
                gcam[ce.i1][2] += gw / ce.den;
                const float gden = -gw * d1 / (ce.den * ce.den);
                if (ce.gap >= eps) { gcam[ce.i1][2] += gden; gcam[ce.i2][2] -= gden; }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float w = o.hom[c][3];
        float gh[4];
        gh[3] = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { gh[k] = gcam[c][k] / w; gh[3] -= gcam[c][k] * o.hom[c][k] / (w * w); }
#pragma unroll
        for (int n = 0; n < 3; ++n)
            gworld[3 * c + n] += gh[0] * E[n] + gh[1] * E[4 + n] + gh[2] * E[8 + n] + gh[3] * E[12 + n];
    }
}

// -torchvision.ops.distance_box_iou (boxes.py::_box_diou_iou): the matching cost of main.py:374-381.
VSRD_HD float diou_pair(const float* a, const float* b, float eps = 1e-7f) {
    const float area1 = (a[2] - a[0]) * (a[3] - a[1]), area2 = (b[2] - b[0]) * (b[3] - b[1]);
    const float w = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.0f);
    const float h = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.0f);
    const float inter = w * h;
    const float iou = inter / (area1 + area2 - inter);
    const float ew = fmaxf(fmaxf(a[2], b[2]) - fminf(a[0], b[0]), 0.0f);
    const float eh = fmaxf(fmaxf(a[3], b[3]) - fminf(a[1], b[1]), 0.0f);
    const float diag = ew * ew + eh * eh + eps;
    const float cx = (a[0] + a[2]) / 2.0f - (b[0] + b[2]) / 2.0f;
    const float cy = (a[1] + a[3]) / 2.0f - (b[1] + b[3]) / 2.0f;
    return iou - (cx * cx + cy * cy) / diag;
}

// gradient share of torch.max(a, b) / torch.min(a, b) w.r.t. a (ties split evenly like ATen)
VSRD_HD float max_share(float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); }
VSRD_HD float min_share(float a, float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.0f); }

// torchvision.ops.distance_box_iou_loss (diou_loss.py::_diou_iou_loss) for one box pair and its
// gradient w.r.t. the predicted box `p` (x1 y1 x2 y2).
VSRD_HD float diou_loss_pair(const float* p, const float* g, float grad[4], float eps = 1e-7f) {
    const float x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
    const float xk1 = fmaxf(x1, g[0]), yk1 = fmaxf(y1, g[1]), xk2 = fminf(x2, g[2]), yk2 = fminf(y2, g[3]);
    const bool overlap = (yk2 > yk1) && (xk2 > xk1);
    const float inter = overlap ? (xk2 - xk1) * (yk2 - yk1) : 0.0f;
    const float uni = (x2 - x1) * (y2 - y1) + (g[2] - g[0]) * (g[3] - g[1]) - inter;
    const float den = uni + eps;
    const float iou = inter / den;
    const float xc1 = fminf(x1, g[0]), yc1 = fminf(y1, g[1]), xc2 = fmaxf(x2, g[2]), yc2 = fmaxf(y2, g[3]);
    const float ew = xc2 - xc1, eh = yc2 - yc1;
    const float diag = ew * ew + eh * eh + eps;
    const float cx = (x2 + x1) / 2.0f - (g[0] + g[2]) / 2.0f;
    const float cy = (y2 + y1) / 2.0f - (g[1] + g[3]) / 2.0f;
    const float cen = cx * cx + cy * cy;
    // d inter / d p
    float gi[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (overlap) {
        gi[0] = -(yk2 - yk1) * max_share(x1, g[0]);
        gi[1] = -(xk2 - xk1) * max_share(y1, g[1]);
        gi[2] = (yk2 - yk1) * min_share(x2, g[2]);
        gi[3] = (xk2 - xk1) * min_share(y2, g[3]);
    }
    const float ga[4] = {-(y2 - y1), -(x2 - x1), (y2 - y1), (x2 - x1)};                  // d area_p
    const float gew[4] = {-min_share(x1, g[0]), 0.0f, max_share(x2, g[2]), 0.0f};        // d ew
    const float geh[4] = {0.0f, -min_share(y1, g[1]), 0.0f, max_share(y2, g[3])};        // d eh
    const float gcx[4] = {0.5f, 0.0f, 0.5f, 0.0f}, gcy[4] = {0.0f, 0.5f, 0.0f, 0.5f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float guni = ga[k] - gi[k];
        const float giou = gi[k] / den - inter * guni / (den * den);
        const float gcen = 2.0f * cx * gcx[k] + 2.0f * cy * gcy[k];
        const float gdiag = 2.0f * ew * gew[k] + 2.0f * eh * geh[k];
        grad[k] = -giou + gcen / diag - cen * gdiag / (diag * diag);
    }
    return 1.0f - iou + cen / diag;
}

// nn.functional.smooth_l1_loss(beta=1, reduction="none") summed over the 4 coordinates, with gradient
VSRD_HD float smooth_l1_pair(const float* p, const float* g, float grad[4]) {
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float d = p[k] - g[k], a = fabsf(d);
        if (a < 1.0f) { sum += 0.5f * d * d; grad[k] = d; }
        else { sum += a - 0.5f; grad[k] = d > 0.0f ? 1.0f : -1.0f; }
    }
    return sum;
}

// Signed pixel distance to a closed polygon (SoftRasterizer.make_distance_map, :267-290, with the
// sign of :306 taken from an even-odd point-in-polygon test instead of cv.fillPoly).
VSRD_HD float polygon_signed_distance(const float* poly, int count, float px, float py) {
    float best = INFINITY;
    bool inside = false;
    for (int k = 0; k < count; ++k) {
        const int kn = (k + 1 == count) ? 0 : k + 1;
        const float ax = poly[2 * k], ay = poly[2 * k + 1], bx = poly[2 * kn], by = poly[2 * kn + 1];
        const float sx = bx - ax, sy = by - ay, qx = px - ax, qy = py - ay;
        float ratio = (sx * qx + sy * qy) / (sx * sx + sy * sy + 1e-6f);
        ratio = fminf(fmaxf(ratio, 0.0f), 1.0f);
        const float nx = qx - sx * ratio, ny = qy - sy * ratio;
        best = fminf(best, sqrtf(nx * nx + ny * ny));
        if ((ay > py) != (by > py)) {
            const float xint = ax + (py - ay) * sx / sy;
            if (px < xint) inside = !inside;
        }
    }
    return inside ? best : -best;
}

}  // namespace vsrd
