// sm_100a kernels of the VSRD hot path, part 5: the per-frame MODELS around the renderer.
//
// One optimisation step of scripts/main.py:328-865 touches, besides the renderer,
//   a3   BoxParameters3D.forward            (models/detectors/box_parameters.py:60-146)
//   a4   HyperDistanceField.forward         (models/fields/hyper_distance_field.py:30-55, 75-77):
//        weight-normed Linear(256,256) -> LayerNorm -> GELU, four times, then weight-normed Linear(256,1617)
//   a16  autograd.backward through both, Adam (5 groups) and ExponentialLR (main.py:859-865)
// on [N,.] tensors with N <= 32.  In PyTorch that is ~150 kernel launches of 2-5 us each per step — a
// third of the step once the renderer runs in ~1.3 ms (profiles/r01_v9_labeler_launches_eager.txt).
// Here it is 5 + 6 + 2 + 1 launches, every one latency-bound on KBs of data:
//
//   decode_boxes_kernel            raw parameters -> locations, half extents, rotations, 8 corners
//   hyper_layer_forward_kernel     one launch per Linear: prologue = LayerNorm + GELU of the previous layer's
//                                  output (recomputed per CTA, N x 256 values), then one warp per output row:
//                                  w = g v / |v| (weight norm, never materialised), y[n][o] = w . a[n] + b
//   hyper_layer_backward_kernel    one launch per Linear, last to first: prologue = input activations (as
//                                  above) and, for the hidden layers, dy = LayerNorm/GELU adjoint of the summed
//                                  partials of the layer above (+ LayerNorm affine gradients); one warp per
//                                  output row: dW = dy^T a, weight-norm adjoint (dv, dg), db, and the row's
//                                  contribution to da = dy W, reduced per CTA in a fixed order into one
//                                  partial (deterministic, no atomics)
//   sum_partials_kernel            embedding gradient = sum of layer 0's partials
//   decode_boxes_backward_kernel   render + projection gradients -> raw parameter gradients, loss record
//   adam_step_kernel               torch.optim.Adam over ONE flat parameter arena with per-group learning
//                                  rates lr0 * gamma^step (ExponentialLR) read from the device step state
//
// Everything is fp32 like the reference; reductions run in a fixed order, so a step is bit-reproducible.
#include "vsrd_common.cuh"

namespace vsrd {
namespace model {

constexpr int kHW = VSRD_HYPER_WIDTH;        // hypernetwork width == embedding size (config.json:142-162)
constexpr int kThreadsM = 256;
constexpr int kWarpsM = kThreadsM / 32;
constexpr int kPL = kHW / 32;                // channels per lane of a 256-wide row
constexpr int kHiddenCtas = 32;              // CTAs (= partial sums) of a hidden layer's backward
constexpr int kMaxRowsPerCta = 64;           // rows of one Linear a backward CTA may own
constexpr int kLastCtas = 64;                // ... of the output layer's backward

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ float gelu_value(float z) { return 0.5f * z * (1.0f + erff(z * kInvSqrt2)); }
__device__ __forceinline__ float gelu_slope(float z) {
    return 0.5f * (1.0f + erff(z * kInvSqrt2)) + z * kInvSqrt2Pi * expf(-0.5f * z * z);
}

// LayerNorm (biased variance, eps 1e-5, affine) of row `x` held as kPL values per lane.
__device__ __forceinline__ void layer_norm_row(const float (&x)[kPL], float (&xhat)[kPL], float& rstd) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kPL; ++k) s += x[k];
    const float mean = warp_sum(s) * (1.0f / kHW);
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < kPL; ++k) { xhat[k] = x[k] - mean; q += xhat[k] * xhat[k]; }
    rstd = rsqrtf(warp_sum(q) * (1.0f / kHW) + kLnEps);
#pragma unroll
    for (int k = 0; k < kPL; ++k) xhat[k] *= rstd;
}

// sA[n][i] = input activations of a Linear: the embeddings themselves (ln_w == NULL) or
// GELU(LayerNorm(x) * ln_w + ln_b) of the previous Linear's output x [N,256].
__device__ __forceinline__ void stage_activations(const float* __restrict__ x, const float* __restrict__ ln_w,
                                                  const float* __restrict__ ln_b, int N, float* sA) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < N; n += kWarpsM) {
        float v[kPL];
#pragma unroll
        for (int k = 0; k < kPL; ++k) v[k] = __ldg(x + (size_t)n * kHW + lane + 32 * k);
        if (ln_w != nullptr) {
            float xhat[kPL], rstd;
            layer_norm_row(v, xhat, rstd);
#pragma unroll
            for (int k = 0; k < kPL; ++k)
                v[k] = gelu_value(fmaf(xhat[k], __ldg(ln_w + lane + 32 * k), __ldg(ln_b + lane + 32 * k)));
        }
#pragma unroll
        for (int k = 0; k < kPL; ++k) sA[n * kHW + lane + 32 * k] = v[k];
    }
}

// ---- a4 forward ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsM) hyper_layer_forward_kernel(
        const float* __restrict__ x, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
        const float* __restrict__ wv, const float* __restrict__ wg, const float* __restrict__ bias,
        int N, int O, float* __restrict__ y) {
    extern __shared__ __align__(16) float smem_m[];
    float* sA = smem_m;
    stage_activations(x, ln_w, ln_b, N, sA);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = blockIdx.x * kWarpsM + warp; o < O; o += gridDim.x * kWarpsM) {
        float v[kPL], q = 0.0f;
#pragma unroll
        for (int k = 0; k < kPL; ++k) { v[k] = __ldg(wv + (size_t)o * kHW + lane + 32 * k); q = fmaf(v[k], v[k], q); }
        const float scale = __ldg(wg + o) / sqrtf(warp_sum(q));       // torch._weight_norm: v * (g / |v|)
        const float b = __ldg(bias + o);
        float mine = 0.0f;                                             // lane n keeps y[n][o]
        for (int n = 0; n < N; ++n) {
            float d = 0.0f;
#pragma unroll
            for (int k = 0; k < kPL; ++k) d = fmaf(v[k], sA[n * kHW + lane + 32 * k], d);
            d = warp_sum(d);
            if (lane == n) mine = fmaf(d, scale, b);
        }
        if (lane < N) y[(size_t)lane * O + o] = mine;
    }
}

// ---- a4 backward ------------------------------------------------------------------------------------
struct LayerBackwardArgs {
    int N, O;
    const float* wv;            // [O,256] this Linear's direction, magnitude
    const float* wg;            // [O]
    const float* x_in;          // [N,256] previous Linear's output, or the embeddings (ln_in_w == NULL)
    const float* ln_in_w;
    const float* ln_in_b;
    const float* dy;            // [N,O] upstream gradient of this Linear's output, or NULL:
    const float* partials_in;   // [num_partials][N][256] partial gradients of the activations after this layer's LN+GELU
    int num_partials;
    const float* y_out;         // [N,256] this Linear's output (input of the LayerNorm that follows)
    const float* ln_out_w;
    const float* ln_out_b;
    float* g_ln_out_w;          // gradients of that LayerNorm's affine
    float* g_ln_out_b;
    float* g_wv;                // outputs
    float* g_wg;
    float* g_bias;
    float* partials_out;        // [gridDim.x][N][256]
};

__global__ void __launch_bounds__(kThreadsM) hyper_layer_backward_kernel(LayerBackwardArgs a) {
    extern __shared__ __align__(16) float smem_m[];
    const int N = a.N, O = a.O;
    float* sA = smem_m;                   // [N][256] input activations
    float* sDA = sA + N * kHW;            // [N][256] xhat scratch of the LayerNorm adjoint
    float* sDY = sDA + N * kHW;           // [N][256] upstream gradient (hidden layers only)
    __shared__ float sRstd[VSRD_MAX_INSTANCES];
    __shared__ float sScaled[kMaxRowsPerCta * 32];   // [local row][instance] dy g / |v|
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    stage_activations(a.x_in, a.ln_in_w, a.ln_in_b, N, sA);
    if (a.dy == nullptr) {
        // sDY <- sum of the partials of the layer above, in partial order (deterministic).  Latency-bound on L2: every
        // thread keeps 2 chunks x 16 partials = 32 independent 128-bit loads in flight (8 in flight cost 11 of the
        // kernel's 20 us at 64 partials).
        {
            const float4* src = reinterpret_cast<const float4*>(a.partials_in);
            float4* dst = reinterpret_cast<float4*>(sDY);
            const int chunks = N * (kHW / 4), P = a.num_partials;
            for (int e0 = threadIdx.x; e0 < chunks; e0 += 2 * kThreadsM) {
                const int e1 = e0 + kThreadsM;
                const bool two = e1 < chunks;
                float4 s0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), s1 = s0;
                for (int p0 = 0; p0 < P; p0 += 16) {
                    float4 v0[16], v1[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const bool in = p0 + u < P;
                        v0[u] = in ? __ldg(src + (size_t)(p0 + u) * chunks + e0) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        v1[u] = in && two ? __ldg(src + (size_t)(p0 + u) * chunks + e1) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        if (p0 + u < P) {
                            s0.x += v0[u].x; s0.y += v0[u].y; s0.z += v0[u].z; s0.w += v0[u].w;
                            s1.x += v1[u].x; s1.y += v1[u].y; s1.z += v1[u].z; s1.w += v1[u].w;
                        }
                    }
                }
                dst[e0] = s0;
                if (two) dst[e1] = s1;
            }
        }
        __syncthreads();
        // dz = (sum of partials) * gelu'(z), z = LayerNorm(y_out) affine; keep dz in sDY and xhat in sDA
        for (int n = warp; n < N; n += kWarpsM) {
            float y[kPL], xhat[kPL], rstd;
#pragma unroll
            for (int k = 0; k < kPL; ++k) y[k] = __ldg(a.y_out + (size_t)n * kHW + lane + 32 * k);
            layer_norm_row(y, xhat, rstd);
            if (lane == 0) sRstd[n] = rstd;
#pragma unroll
            for (int k = 0; k < kPL; ++k) {
                const int i = lane + 32 * k;
                const float z = fmaf(xhat[k], __ldg(a.ln_out_w + i), __ldg(a.ln_out_b + i));
                sDY[n * kHW + i] *= gelu_slope(z);
                sDA[n * kHW + i] = xhat[k];
            }
        }
        __syncthreads();
        if (blockIdx.x == 0) {            // LayerNorm affine gradients: sums over the instances, fixed order
            const int i = threadIdx.x;
            float gw = 0.0f, gb = 0.0f;
            for (int n = 0; n < N; ++n) { const float dz = sDY[n * kHW + i]; gw = fmaf(dz, sDA[n * kHW + i], gw); gb += dz; }
            a.g_ln_out_w[i] = gw;
            a.g_ln_out_b[i] = gb;
        }
        __syncthreads();
        // dy = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), dxhat = dz * ln_w
        for (int n = warp; n < N; n += kWarpsM) {
            float dx[kPL], xh[kPL], s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int k = 0; k < kPL; ++k) {
                const int i = lane + 32 * k;
                xh[k] = sDA[n * kHW + i];
                dx[k] = sDY[n * kHW + i] * __ldg(a.ln_out_w + i);
                s0 += dx[k];
                s1 = fmaf(dx[k], xh[k], s1);
            }
            const float m0 = warp_sum(s0) * (1.0f / kHW), m1 = warp_sum(s1) * (1.0f / kHW);
            const float rstd = sRstd[n];
#pragma unroll
            for (int k = 0; k < kPL; ++k) sDY[n * kHW + lane + 32 * k] = rstd * (dx[k] - m0 - xh[k] * m1);
        }
    }
    __syncthreads();

    // phase 1, one warp per output row: dW = dy^T a with the weight-norm adjoint, db; the row's scaled upstream
    // gradient dy[n][o] g / |v| is parked in shared memory for phase 2
    const int row_stride = gridDim.x * kWarpsM;
    int r = warp;                                  // local row index: o = blockIdx.x * 8 + (r & 7) + (r >> 3) * row_stride
    for (int o = blockIdx.x * kWarpsM + warp; o < O; o += row_stride, r += kWarpsM) {
        float v[kPL], q = 0.0f;
#pragma unroll
        for (int k = 0; k < kPL; ++k) { v[k] = __ldg(a.wv + (size_t)o * kHW + lane + 32 * k); q = fmaf(v[k], v[k], q); }
        const float inv = 1.0f / sqrtf(warp_sum(q));
        const float g = __ldg(a.wg + o);
        const float scale = g * inv;
        float mine = 0.0f;                         // lane n holds dy[n][o]
        if (lane < N) mine = a.dy != nullptr ? __ldg(a.dy + (size_t)lane * O + o) : sDY[lane * kHW + o];
        sScaled[r * 32 + lane] = mine * scale;
        float dw[kPL], db = 0.0f;
#pragma unroll
        for (int k = 0; k < kPL; ++k) dw[k] = 0.0f;
        for (int n = 0; n < N; ++n) {
            const float d = __shfl_sync(kFull, mine, n);
            db += d;
#pragma unroll
            for (int k = 0; k < kPL; ++k) dw[k] = fmaf(d, sA[n * kHW + lane + 32 * k], dw[k]);
        }
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < kPL; ++k) s = fmaf(dw[k], v[k], s);
        s = warp_sum(s);
        // w = g v / |v|:  dg = (dw . v) / |v|,  dv = (g / |v|) dw - (g (dw . v) / |v|^3) v
        const float back = g * s * inv * inv * inv;
#pragma unroll
        for (int k = 0; k < kPL; ++k) a.g_wv[(size_t)o * kHW + lane + 32 * k] = fmaf(scale, dw[k], -back * v[k]);
        if (lane == 0) { a.g_wg[o] = s * inv; a.g_bias[o] = db; }
    }
    __syncthreads();
    // phase 2, one warp per 32 input channels: this CTA's partial of da[n][i] = sum_o dy[n][o] w[o][i] over its rows,
    // accumulated in row order by the lane that owns channel i (no cross-warp reduction, deterministic)
    const int i = 32 * warp + lane;
    const int rows = (O - blockIdx.x * kWarpsM + row_stride - 1) / row_stride * kWarpsM;    // upper bound on local rows
    float* out = a.partials_out + (size_t)blockIdx.x * N * kHW;
    for (int c = 0; c < N; c += 8) {               // instances in chunks of 8 accumulators
        float acc[8];
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) acc[nn] = 0.0f;
#pragma unroll 4
        for (int rr = 0; rr < rows; ++rr) {
            const int o = blockIdx.x * kWarpsM + (rr & (kWarpsM - 1)) + (rr >> 3) * row_stride;
            if (o < O) {
                const float wv = __ldg(a.wv + (size_t)o * kHW + i);
#pragma unroll
                for (int nn = 0; nn < 8; ++nn) acc[nn] = fmaf(sScaled[rr * 32 + ((c + nn) & 31)], wv, acc[nn]);
            }
        }
#pragma unroll
        for (int nn = 0; nn < 8; ++nn)
            if (c + nn < N) out[(size_t)(c + nn) * kHW + i] = acc[nn];
    }
}

__global__ void sum_partials_kernel(const float4* __restrict__ partials, int num_partials, int chunks, float4* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= chunks) return;
    float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 8
    for (int p = 0; p < num_partials; ++p) {
        const float4 v = __ldg(partials + (size_t)p * chunks + e);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[e] = s;
}

// ---- a3: BoxParameters3D --------------------------------------------------------------------------
// torch.lerp(start, end, w): w < 0.5 ? start + w (end - start) : end - (end - start)(1 - w)
__device__ __forceinline__ float lerp_torch(float lo, float hi, float w) {
    const float diff = hi - lo;
    return w < 0.5f ? lo + w * diff : hi - diff * (1.0f - w);
}

__device__ __constant__ float kCornerSigns[8][3] = {       // box_parameters.py:77-86
    {-1.0f, -1.0f, +1.0f}, {+1.0f, -1.0f, +1.0f}, {+1.0f, -1.0f, -1.0f}, {-1.0f, -1.0f, -1.0f},
    {-1.0f, +1.0f, +1.0f}, {+1.0f, +1.0f, +1.0f}, {+1.0f, +1.0f, -1.0f}, {-1.0f, +1.0f, -1.0f}};

struct DecodeArgs {
    int N;
    const float* raw_loc;       // [N,3]
    const float* raw_dim;       // [N,3]
    const float* raw_ori;       // [N,2]
    float loc_lo[3], loc_hi[3], dim_lo[3], dim_hi[3];
    float* loc;                 // [N,3]
    float* dim;                 // [N,3]
    float* rot;                 // [N,9]
    float* boxes;               // [N,8,3]
};

__global__ void decode_boxes_kernel(DecodeArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.N) return;
    float t[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        t[k] = lerp_torch(a.loc_lo[k], a.loc_hi[k], sigmoidf_(a.raw_loc[3 * n + k]));
        d[k] = lerp_torch(a.dim_lo[k], a.dim_hi[k], sigmoidf_(a.raw_dim[3 * n + k]));
        a.loc[3 * n + k] = t[k];
        a.dim[3 * n + k] = d[k];
    }
    const float o0 = a.raw_ori[2 * n], o1 = a.raw_ori[2 * n + 1];
    const float nrm = fmaxf(sqrtf(o0 * o0 + o1 * o1), 1e-12f);       // F.normalize eps
    const float c = o0 / nrm, s = o1 / nrm;
    const float R[9] = {c, 0.0f, s, 0.0f, 1.0f, 0.0f, -s, 0.0f, c};   // rotation_matrix_y
#pragma unroll
    for (int k = 0; k < 9; ++k) a.rot[9 * n + k] = R[k];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            float v = 0.0f;
#pragma unroll
            for (int j = 0; j < 3; ++j) v += kCornerSigns[k][j] * d[j] * R[3 * m + j];     // corners @ R^T
            a.boxes[24 * n + 3 * k + m] = v + t[m];
        }
}

struct DecodeBackwardArgs {
    int N;
    const float* raw_loc;
    const float* raw_dim;
    const float* raw_ori;
    float loc_lo[3], loc_hi[3], dim_lo[3], dim_hi[3];
    const float* dim;           // decoded half extents [N,3]
    const float* rot;           // decoded rotations [N,9]
    const float* g_loc;         // render gradients w.r.t. the decoded values
    const float* g_dim;
    const float* g_rot;
    const float* g_boxes;       // [2][N][24] projection-loss gradients w.r.t. the corners, or NULL
    float w_iou, w_l1;
    float* g_raw_loc;           // outputs
    float* g_raw_dim;
    float* g_raw_ori;
    // loss record: losses[5] = total, silhouette, eikonal, iou, l1 (weighted)
    const float* render_parts;  // [2] weighted silhouette / eikonal
    const float* proj_losses;   // [2] unweighted
    float* losses;
};

__global__ void decode_boxes_backward_kernel(DecodeBackwardArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0 && a.losses != nullptr) {
        const float sil = a.render_parts[0], eik = a.render_parts[1];
        const float iou = a.proj_losses ? a.w_iou * a.proj_losses[0] : 0.0f;
        const float l1 = a.proj_losses ? a.w_l1 * a.proj_losses[1] : 0.0f;
        a.losses[0] = sil + eik + iou + l1;
        a.losses[1] = sil; a.losses[2] = eik; a.losses[3] = iou; a.losses[4] = l1;
    }
    if (n >= a.N) return;
    float gt[3], gd[3], gR[9], d[3], R[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) { gt[k] = a.g_loc[3 * n + k]; gd[k] = a.g_dim[3 * n + k]; d[k] = a.dim[3 * n + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) { gR[k] = a.g_rot[9 * n + k]; R[k] = a.rot[9 * n + k]; }
    if (a.g_boxes != nullptr) {
        const float* g0 = a.g_boxes + 24 * n;
        const float* g1 = a.g_boxes + 24 * ((size_t)a.N + n);
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const float gb = a.w_iou * g0[3 * k + m] + a.w_l1 * g1[3 * k + m];
                gt[m] += gb;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    gR[3 * m + j] = fmaf(gb, kCornerSigns[k][j] * d[j], gR[3 * m + j]);
                    gd[j] = fmaf(gb, kCornerSigns[k][j] * R[3 * m + j], gd[j]);
                }
            }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float sl = sigmoidf_(a.raw_loc[3 * n + k]), sd = sigmoidf_(a.raw_dim[3 * n + k]);
        a.g_raw_loc[3 * n + k] = gt[k] * (a.loc_hi[k] - a.loc_lo[k]) * sl * (1.0f - sl);
        a.g_raw_dim[3 * n + k] = gd[k] * (a.dim_hi[k] - a.dim_lo[k]) * sd * (1.0f - sd);
    }
    // R = [[c,0,s],[0,1,0],[-s,0,c]], (c, s) = o / max(|o|, eps)
    const float gc = gR[0] + gR[8], gs = gR[2] - gR[6];
    const float o0 = a.raw_ori[2 * n], o1 = a.raw_ori[2 * n + 1];
    const float len = sqrtf(o0 * o0 + o1 * o1);
    if (len > 1e-12f) {
        const float c = o0 / len, s = o1 / len, dot = gc * c + gs * s;
        a.g_raw_ori[2 * n] = (gc - c * dot) / len;
        a.g_raw_ori[2 * n + 1] = (gs - s * dot) / len;
    } else {                                   // clamp_min branch of F.normalize: o / eps
        a.g_raw_ori[2 * n] = gc / 1e-12f;
        a.g_raw_ori[2 * n + 1] = gs / 1e-12f;
    }
}

// ---- a16: Adam + ExponentialLR over the flat arena ---------------------------------------------------
struct AdamArgs {
    float* params;
    const float* grads;
    float* exp_avg;
    float* exp_avg_sq;
    long long numel;
    VsrdAdamGroups groups;
    const VsrdStepState* state;
    long long host_step;
};

__global__ void adam_step_kernel(AdamArgs a) {
    __shared__ float s_step_size[VSRD_MAX_PARAM_GROUPS], s_inv_bc2_sqrt[VSRD_MAX_PARAM_GROUPS];
    if (threadIdx.x < a.groups.num_groups) {
        const int gi = threadIdx.x;
        const long long step = a.state != nullptr ? (long long)a.state->step : a.host_step;
        const long long t = step - a.groups.first_step[gi] + 1;          // this group's own step count (1-based)
        float step_size = 0.0f, inv = 0.0f;
        if (t >= 1) {
            const double lr = (double)a.groups.base_lr[gi] * exp(a.groups.log_gamma * (double)step);
            const double bc1 = 1.0 - pow((double)a.groups.beta1, (double)t);
            const double bc2 = 1.0 - pow((double)a.groups.beta2, (double)t);
            step_size = (float)(lr / bc1);
            inv = (float)(1.0 / sqrt(bc2));
        }
        s_step_size[gi] = step_size;
        s_inv_bc2_sqrt[gi] = inv;
    }
    __syncthreads();
    const float b1 = a.groups.beta1, b2 = a.groups.beta2, eps = a.groups.eps;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < a.numel; e += (long long)gridDim.x * blockDim.x) {
        int gi = 0;
        while (gi + 1 < a.groups.num_groups && e >= a.groups.group_end[gi]) ++gi;
        const float step_size = s_step_size[gi];
        if (step_size == 0.0f) continue;                                 // group not optimised yet (no gradient)
        const float g = a.grads[e];
        const float m = fmaf(g - a.exp_avg[e], 1.0f - b1, a.exp_avg[e]);
        const float v = fmaf(b2, a.exp_avg_sq[e], (1.0f - b2) * g * g);
        a.exp_avg[e] = m;
        a.exp_avg_sq[e] = v;
        const float denom = fmaf(sqrtf(v), s_inv_bc2_sqrt[gi], eps);
        a.params[e] -= step_size * m / denom;
    }
}

static int g_model_ready = 0;
static int g_model_sms = 148;

static int setup() {
    if (g_model_ready) return 0;
    const int max_smem = 3 * VSRD_MAX_INSTANCES * kHW * (int)sizeof(float);
    if (cudaFuncSetAttribute(hyper_layer_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess)
        return fail("vsrd_b200: cannot reserve %s of shared memory for hyper_layer_backward_kernel", "96 KB");
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) g_model_sms = prop.multiProcessorCount;
    g_model_ready = 1;
    return 0;
}

static int check_net(const VsrdHyperNet* net) {
    VSRD_CHECK_ARG(net != nullptr, "hypernetwork is NULL");
    VSRD_CHECK_ARG(net->num_layers >= 2 && net->num_layers <= VSRD_HYPER_MAX_LAYERS, "hypernetwork must have 2..5 Linear layers");
    for (int l = 0; l < net->num_layers; ++l) {
        const VsrdHyperLayer& L = net->layers[l];
        VSRD_CHECK_ARG(L.weight_v && L.weight_g && L.bias, "hypernetwork layer pointers must not be NULL");
        VSRD_CHECK_ARG(L.in_features == kHW, "the hypernetwork kernels are compiled for 256 input features per layer");
        VSRD_CHECK_ARG(l == net->num_layers - 1 ? L.out_features >= 1 : L.out_features == kHW,
                       "hidden hypernetwork layers must be 256 wide");
        VSRD_CHECK_ARG(l == net->num_layers - 1 || (L.ln_weight && L.ln_bias), "hidden layers need LayerNorm parameters");
    }
    return 0;
}

}  // namespace model
}  // namespace vsrd

using namespace vsrd;
using namespace vsrd::model;

extern "C" {

size_t vsrd_hyper_scratch_floats(int num_instances) {
    return (size_t)(kLastCtas + kHiddenCtas) * (size_t)num_instances * kHW;
}

int vsrd_hyper_forward(const VsrdHyperNet* net, const float* embeddings, int num_instances,
                       float* activations, float* mlp_weights, void* stream) {
    if (check_net(net)) return 1;
    VSRD_CHECK_ARG(num_instances >= 1 && num_instances <= VSRD_MAX_INSTANCES, "num_instances must be in [1, 32]");
    VSRD_CHECK_ARG(embeddings && activations && mlp_weights, "embeddings / activations / mlp_weights must not be NULL");
    if (model::setup()) return 1;
    const int N = num_instances, L = net->num_layers;
    const size_t smem = (size_t)N * kHW * sizeof(float);
    for (int l = 0; l < L; ++l) {
        const VsrdHyperLayer& P = net->layers[l];
        const float* x = l == 0 ? embeddings : activations + (size_t)(l - 1) * N * kHW;
        float* y = l == L - 1 ? mlp_weights : activations + (size_t)l * N * kHW;
        const int want = (P.out_features + kWarpsM - 1) / kWarpsM;
        const int grid = want < g_model_sms ? want : g_model_sms;
        hyper_layer_forward_kernel<<<grid, kThreadsM, smem, (cudaStream_t)stream>>>(
            x, l == 0 ? nullptr : net->layers[l - 1].ln_weight, l == 0 ? nullptr : net->layers[l - 1].ln_bias,
            P.weight_v, P.weight_g, P.bias, N, P.out_features, y);
        VSRD_CHECK_LAUNCH();
    }
    return 0;
}

int vsrd_hyper_backward(const VsrdHyperNet* net, const VsrdHyperNetGrads* grads, const float* embeddings,
                        int num_instances, const float* activations, const float* grad_mlp_weights,
                        float* grad_embeddings, float* scratch, void* stream) {
    if (check_net(net)) return 1;
    VSRD_CHECK_ARG(grads != nullptr && grads->num_layers == net->num_layers, "gradient table must match the hypernetwork");
    VSRD_CHECK_ARG(num_instances >= 1 && num_instances <= VSRD_MAX_INSTANCES, "num_instances must be in [1, 32]");
    VSRD_CHECK_ARG(embeddings && activations && grad_mlp_weights && grad_embeddings && scratch, "hyper backward pointers must not be NULL");
    if (model::setup()) return 1;
    const int N = num_instances, L = net->num_layers;
    float* part[2] = {scratch, scratch + (size_t)kLastCtas * N * kHW};     // ping (output layer's partials first) / pong
    int prev_partials = 0;
    for (int l = L - 1; l >= 0; --l) {
        const VsrdHyperLayer& P = net->layers[l];
        const VsrdHyperLayerGrads& G = grads->layers[l];
        VSRD_CHECK_ARG(G.weight_v && G.weight_g && G.bias, "gradient pointers must not be NULL");
        const bool last = l == L - 1;
        VSRD_CHECK_ARG(last || (G.ln_weight && G.ln_bias), "LayerNorm gradient pointers must not be NULL");
        LayerBackwardArgs a;
        a.N = N; a.O = P.out_features;
        a.wv = P.weight_v; a.wg = P.weight_g;
        a.x_in = l == 0 ? embeddings : activations + (size_t)(l - 1) * N * kHW;
        a.ln_in_w = l == 0 ? nullptr : net->layers[l - 1].ln_weight;
        a.ln_in_b = l == 0 ? nullptr : net->layers[l - 1].ln_bias;
        a.dy = last ? grad_mlp_weights : nullptr;
        a.partials_in = last ? nullptr : part[(L - 2 - l) & 1];
        a.num_partials = prev_partials;
        a.y_out = last ? nullptr : activations + (size_t)l * N * kHW;
        a.ln_out_w = last ? nullptr : P.ln_weight;
        a.ln_out_b = last ? nullptr : P.ln_bias;
        a.g_ln_out_w = last ? nullptr : G.ln_weight;
        a.g_ln_out_b = last ? nullptr : G.ln_bias;
        a.g_wv = G.weight_v; a.g_wg = G.weight_g; a.g_bias = G.bias;
        a.partials_out = part[(L - 1 - l) & 1];
        const int grid = last ? kLastCtas : kHiddenCtas;
        VSRD_CHECK_ARG((P.out_features + grid * kWarpsM - 1) / (grid * kWarpsM) * kWarpsM <= kMaxRowsPerCta,
                       "output layer too wide for the backward kernel's row table");
        const size_t smem = (size_t)(last ? 2 : 3) * N * kHW * sizeof(float);
        hyper_layer_backward_kernel<<<grid, kThreadsM, smem, (cudaStream_t)stream>>>(a);
        VSRD_CHECK_LAUNCH();
        prev_partials = grid;
    }
    VSRD_CHECK_ARG((reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad_embeddings) & 15) == 0,
                   "scratch and grad_embeddings must be 16-byte aligned");
    const int chunks = N * kHW / 4;
    sum_partials_kernel<<<(chunks + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(part[(L - 1) & 1]), prev_partials, chunks, reinterpret_cast<float4*>(grad_embeddings));
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_decode_boxes(const VsrdBoxRanges* ranges, const float* raw_locations, const float* raw_dimensions,
                      const float* raw_orientations, int num_instances, float* locations, float* half_extents,
                      float* rotations, float* boxes_3d, void* stream) {
    VSRD_CHECK_ARG(ranges != nullptr, "ranges is NULL");
    VSRD_CHECK_ARG(num_instances >= 1 && num_instances <= VSRD_MAX_INSTANCES, "num_instances must be in [1, 32]");
    VSRD_CHECK_ARG(raw_locations && raw_dimensions && raw_orientations && locations && half_extents && rotations && boxes_3d,
                   "decode pointers must not be NULL");
    DecodeArgs a;
    a.N = num_instances;
    a.raw_loc = raw_locations; a.raw_dim = raw_dimensions; a.raw_ori = raw_orientations;
    for (int k = 0; k < 3; ++k) {
        a.loc_lo[k] = ranges->location_min[k]; a.loc_hi[k] = ranges->location_max[k];
        a.dim_lo[k] = ranges->dimension_min[k]; a.dim_hi[k] = ranges->dimension_max[k];
    }
    a.loc = locations; a.dim = half_extents; a.rot = rotations; a.boxes = boxes_3d;
    decode_boxes_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_decode_boxes_backward(const VsrdBoxRanges* ranges, const float* raw_locations, const float* raw_dimensions,
                               const float* raw_orientations, int num_instances, const float* half_extents,
                               const float* rotations, const float* grad_locations, const float* grad_half_extents,
                               const float* grad_rotations, const float* grad_boxes_3d, float iou_weight, float l1_weight,
                               float* grad_raw_locations, float* grad_raw_dimensions, float* grad_raw_orientations,
                               const float* render_loss_parts, const float* projection_losses, float* losses, void* stream) {
    VSRD_CHECK_ARG(ranges != nullptr, "ranges is NULL");
    VSRD_CHECK_ARG(num_instances >= 1 && num_instances <= VSRD_MAX_INSTANCES, "num_instances must be in [1, 32]");
    VSRD_CHECK_ARG(raw_locations && raw_dimensions && raw_orientations && half_extents && rotations, "decode pointers must not be NULL");
    VSRD_CHECK_ARG(grad_locations && grad_half_extents && grad_rotations, "render gradients must not be NULL");
    VSRD_CHECK_ARG(grad_raw_locations && grad_raw_dimensions && grad_raw_orientations, "output gradients must not be NULL");
    VSRD_CHECK_ARG(losses == nullptr || render_loss_parts != nullptr, "the loss record needs render_loss_parts");
    DecodeBackwardArgs a;
    a.N = num_instances;
    a.raw_loc = raw_locations; a.raw_dim = raw_dimensions; a.raw_ori = raw_orientations;
    for (int k = 0; k < 3; ++k) {
        a.loc_lo[k] = ranges->location_min[k]; a.loc_hi[k] = ranges->location_max[k];
        a.dim_lo[k] = ranges->dimension_min[k]; a.dim_hi[k] = ranges->dimension_max[k];
    }
    a.dim = half_extents; a.rot = rotations;
    a.g_loc = grad_locations; a.g_dim = grad_half_extents; a.g_rot = grad_rotations; a.g_boxes = grad_boxes_3d;
    a.w_iou = iou_weight; a.w_l1 = l1_weight;
    a.g_raw_loc = grad_raw_locations; a.g_raw_dim = grad_raw_dimensions; a.g_raw_ori = grad_raw_orientations;
    a.render_parts = render_loss_parts; a.proj_losses = projection_losses; a.losses = losses;
    decode_boxes_backward_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t numel,
                   const VsrdAdamGroups* groups, const VsrdStepState* step_state, int64_t step, void* stream) {
    VSRD_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "Adam buffers must not be NULL");
    VSRD_CHECK_ARG(groups != nullptr && groups->num_groups >= 1 && groups->num_groups <= VSRD_MAX_PARAM_GROUPS,
                   "1..8 parameter groups");
    VSRD_CHECK_ARG(numel >= 0 && groups->group_end[groups->num_groups - 1] == numel, "the last group must end at numel");
    if (numel == 0) return 0;
    if (model::setup()) return 1;
    AdamArgs a{params, grads, exp_avg, exp_avg_sq, (long long)numel, *groups, step_state, (long long)step};
    const long long want = (numel + 255) / 256;
    const int grid = (int)(want < 4LL * g_model_sms ? want : 4LL * g_model_sms);
    adam_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
