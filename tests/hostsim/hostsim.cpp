// TEST INFRASTRUCTURE ONLY.
// Compiles the per-thread device math (vsrd_b200/csrc/vsrd_math.cuh) for the host so the formulas the
// CUDA kernels execute can be checked against the autograd oracle in a container without a GPU.
// The warp-parallel glue of the kernels is replaced by plain serial loops here; nothing in the
// product path links against this file.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../vsrd_b200/csrc/vsrd_math.cuh"
#include "../../vsrd_b200/csrc/vsrd_frame_math.cuh"

using namespace vsrd;

namespace {

struct HostSink {
    double* acc;   // [kGradStride], original (reference) weight layout then pose
    void layer0(const float* hbar, const float* hdbar, const float* e, const float* ed) {
        for (int o = 0; o < kHid; ++o) {
            for (int j = 0; j < kEnc; ++j) acc[kW0 + o * (kEnc + 1) + j] += (double)hbar[o] * e[j] + (double)hdbar[o] * ed[j];
            acc[kW0 + o * (kEnc + 1) + kEnc] += hbar[o];
        }
    }
    void hidden(int l, const float* hbar, const float* hdbar, const float* g, const float* gd) {
        const int base = kW1 + (l - 1) * kWStride;
        for (int o = 0; o < kHid; ++o) {
            for (int i = 0; i < kHid; ++i) acc[base + o * (kHid + 1) + i] += (double)hbar[o] * g[i] + (double)hdbar[o] * gd[i];
            acc[base + o * (kHid + 1) + kHid] += hbar[o];
        }
    }
    void last(float obar, float odbar, const float* g, const float* gd) {
        for (int i = 0; i < kHid; ++i) acc[kW4 + i] += (double)obar * g[i] + (double)odbar * gd[i];
        acc[kW4 + kHid] += obar;
    }
    void pose(const float* tbar, const float* dimbar, const float* Rbar) {
        for (int k = 0; k < 3; ++k) acc[kNumW + k] += tbar[k];
        for (int k = 0; k < 3; ++k) acc[kNumW + 3 + k] += dimbar[k];
        for (int k = 0; k < 9; ++k) acc[kNumW + 6 + k] += Rbar[k];
    }
};

Instance make_instance(const float* t, const float* R, const float* dim) {
    Instance I;
    for (int k = 0; k < 3; ++k) { I.t[k] = t[k]; I.dim[k] = dim[k]; }
    for (int k = 0; k < 9; ++k) I.R[k] = R[k];
    return I;
}

void stage(const float* W, std::vector<float>& Wt) {
    Wt.resize(kNumW);
    for (int f = 0; f < kNumW; ++f) Wt[staged_index(f)] = W[f];
}

}  // namespace

extern "C" {

// x [S,3]; out [S,4] = (d, Gx, Gy, Gz).  W may be null (box-only).
void hs_field_forward(const float* x, const float* t, const float* R, const float* dim, const float* W,
                      float scale, int S, float* out) {
    Instance I = make_instance(t, R, dim);
    std::vector<float> Wt;
    if (W) stage(W, Wt);
    for (int s = 0; s < S; ++s) {
        float d, G[3], act[kFwdActRows];
        if (W) field_forward_looped<true>(x + 3 * s, I, Wt.data(), scale, act, 1, d, G);
        else field_forward_looped<false>(x + 3 * s, I, nullptr, scale, act, 1, d, G);
        out[4 * s] = d; out[4 * s + 1] = G[0]; out[4 * s + 2] = G[1]; out[4 * s + 3] = G[2];
    }
}

// adj [S,4] = (dd, dGx, dGy, dGz); grads [kGradStride] (double, zeroed here).
void hs_field_backward(const float* x, const float* t, const float* R, const float* dim, const float* W,
                       float scale, int S, const float* adj, double* grads) {
    Instance I = make_instance(t, R, dim);
    std::vector<float> Wt;
    if (W) stage(W, Wt);
    std::memset(grads, 0, sizeof(double) * kGradStride);
    HostSink sink{grads};
    for (int s = 0; s < S; ++s) {
        const float* a = adj + 4 * s;
        const float dG[3] = {a[1], a[2], a[3]};
        if (W) field_backward<true>(x + 3 * s, I, Wt.data(), scale, a[0], dG, sink);
        else field_backward<false>(x + 3 * s, I, nullptr, scale, a[0], dG, sink);
    }
}

// Serial restatement of the compositing kernels for R rays with M intervals.
//   t [R, M+1]; dirs [R,3]; F [N, R*M, 4]; labels [R,N]; grads [R,M,3]; weights [R,M]
void hs_composite_forward(const float* t, const float* dirs, const float* F, int R, int M, int N,
                          float T, float sigma, float rho, float eps,
                          float* labels, float* grads, float* weights) {
    const size_t stride = (size_t)R * M;
    for (int r = 0; r < R; ++r) {
        float trans = 1.0f;
        for (int n = 0; n < N; ++n) labels[r * N + n] = 0.0f;
        for (int j = 0; j < M; ++j) {
            const size_t idx = (size_t)r * M + j;
            auto load = [&](int i) { const float* f = F + (i * stride + idx) * 4; return Vec4{f[0], f[1], f[2], f[3]}; };
            UnionEval u;
            union_forward(load, N, T, u);
            OpacityEval o;
            const float delta = t[r * (M + 1) + j + 1] - t[r * (M + 1) + j];
            opacity_forward(u, dirs + 3 * r, delta, sigma, rho, eps, o);
            const float w = trans * o.alpha;
            trans *= (1.0f - o.alpha);
            weights[idx] = w;
            for (int c = 0; c < 3; ++c) grads[idx * 3 + c] = u.g[c];
            for (int n = 0; n < N; ++n) labels[r * N + n] += w * (expf(-(load(n).x / T) - u.mneg) / u.Z);
        }
    }
}

// Upstream grads: gl [R,N] (labels), gg [R,M,3] (sampled gradients), gw [R,M] (weights); any may be null.
// adj [N, R*M, 4].
void hs_composite_backward(const float* t, const float* dirs, const float* F, int R, int M, int N,
                           float T, float sigma, float rho, float eps,
                           const float* gl, const float* gg, const float* gw, float* adj) {
    const size_t stride = (size_t)R * M;
    std::vector<UnionEval> us(M);
    std::vector<OpacityEval> os(M);
    std::vector<float> trans(M), omega(M), a(M);
    for (int r = 0; r < R; ++r) {
        float tr = 1.0f;
        for (int j = 0; j < M; ++j) {
            const size_t idx = (size_t)r * M + j;
            auto load = [&](int i) { const float* f = F + (i * stride + idx) * 4; return Vec4{f[0], f[1], f[2], f[3]}; };
            union_forward(load, N, T, us[j]);
            const float delta = t[r * (M + 1) + j + 1] - t[r * (M + 1) + j];
            opacity_forward(us[j], dirs + 3 * r, delta, sigma, rho, eps, os[j]);
            trans[j] = tr;
            omega[j] = tr * os[j].alpha;
            tr *= (1.0f - os[j].alpha);
            float aj = gw ? gw[idx] : 0.0f;
            if (gl) for (int n = 0; n < N; ++n) aj += gl[r * N + n] * (expf(-(load(n).x / T) - us[j].mneg) / us[j].Z);
            a[j] = aj;
        }
        float suffix = 0.0f;   // sum_{k>j} a_k omega_k
        for (int j = M - 1; j >= 0; --j) {
            const size_t idx = (size_t)r * M + j;
            auto load = [&](int i) { const float* f = F + (i * stride + idx) * 4; return Vec4{f[0], f[1], f[2], f[3]}; };
            const float alpha_bar = a[j] * trans[j] - suffix / (1.0f - os[j].alpha);
            suffix += a[j] * omega[j];
            const float delta = t[r * (M + 1) + j + 1] - t[r * (M + 1) + j];
            float dbar_adj, gbar[3];
            opacity_backward(os[j], dirs + 3 * r, delta, sigma, rho, eps, alpha_bar, dbar_adj, gbar);
            if (gg) for (int c = 0; c < 3; ++c) gbar[c] += gg[idx * 3 + c];
            const float om = omega[j];
            auto wbar = [&](int i) { return gl ? om * gl[r * N + i] : 0.0f; };
            auto store = [&](int i, const Vec4& v) { float* o = adj + (i * stride + idx) * 4; o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; };
            union_backward(load, wbar, store, N, T, us[j], dbar_adj, gbar);
        }
    }
}

// ---- frame-level rows (vsrd_frame_math.cuh): serial stand-in for projection_step_kernel with the
// assignment supplied by the caller (the warp-parallel Hungarian solver only exists on the device).
void hs_projection_step(const float* E, const float* K, const float* world, const float* gt, const uint8_t* visible,
                        const int64_t* gt_idx, int V, int N, float height, float width,
                        float* boxes, float* cost, float* losses, float* grad_world /*[2,N,24]*/) {
    const float eps = 1e-6f;
    std::vector<BoxProjection> bp((size_t)V * N);
    for (int v = 0; v < V; ++v)
        for (int n = 0; n < N; ++n) {
            project_box(E + 16 * v, K + 9 * v, world + 24 * n, height, width, eps, bp[v * N + n]);
            for (int k = 0; k < 4; ++k) boxes[(v * N + n) * 4 + k] = bp[v * N + n].box[k];
        }
    if (!gt) return;
    if (cost)   // target view passed as view 0 of `cost` by the caller's choice of pointer offsets
        for (int a = 0; a < N; ++a)
            for (int b = 0; b < N; ++b) cost[a * N + b] = -diou_pair(boxes + 4 * a, gt + 4 * b);
    float iou = 0.0f, l1 = 0.0f, count = 0.0f;
    std::vector<float> gbox((size_t)V * N * 8, 0.0f);
    for (int v = 0; v < V; ++v)
        for (int k = 0; k < N; ++k) {
            const int g = (int)gt_idx[k];
            if (visible && !visible[v * N + g]) continue;
            iou += diou_loss_pair(boxes + (v * N + k) * 4, gt + (v * N + g) * 4, &gbox[(v * N + k) * 8]);
            l1 += smooth_l1_pair(boxes + (v * N + k) * 4, gt + (v * N + g) * 4, &gbox[(v * N + k) * 8 + 4]);
            count += 1.0f;
        }
    losses[0] = iou / count;
    losses[1] = l1 / (4.0f * count);
    for (int i = 0; i < 2 * N * 24; ++i) grad_world[i] = 0.0f;
    for (int v = 0; v < V; ++v)
        for (int n = 0; n < N; ++n)
            for (int which = 0; which < 2; ++which) {
                float g[4];
                const float scale = which ? 1.0f / (4.0f * count) : 1.0f / count;
                for (int c = 0; c < 4; ++c) g[c] = gbox[(v * N + n) * 8 + 4 * which + c] * scale;
                project_box_backward(E + 16 * v, K + 9 * v, bp[v * N + n], height, width, eps, g,
                                     grad_world + (which * N + n) * 24);
            }
}

void hs_soft_mask(const float* polygon, int count, int H, int W, float temperature, float* signed_distance, float* mask) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float d = polygon_signed_distance(polygon, count, (float)x, (float)y);
            signed_distance[y * W + x] = d;
            mask[y * W + x] = sigmoidf_(d / temperature);
        }
}

}  // extern "C"
