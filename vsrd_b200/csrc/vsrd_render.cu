// sm_100a kernels of the VSRD silhouette-renderer hot path, part 1/3: ray generation, sample
// placement and the CTA-per-ray compositing kernels (forward and adjoint) + their C entry points.
// See include/vsrd_b200.h for the ABI and DESIGN.md for the kernel map and rooflines.
#include "vsrd_common.cuh"

namespace vsrd {

thread_local char g_error[512] = "";

// =============================================================================================
// a1: rays
// =============================================================================================
__device__ __forceinline__ void pixel_direction(const float* __restrict__ P, float u, float v, float d[3]) {
    float n2 = 0.0f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        d[m] = P[3 * m] * u + P[3 * m + 1] * v + P[3 * m + 2];
        n2 += d[m] * d[m];
    }
    const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);   // F.normalize
#pragma unroll
    for (int m = 0; m < 3; ++m) d[m] *= inv;
}

__global__ void ray_directions_kernel(const float* __restrict__ inv_proj, int V, int H, int W, float* __restrict__ out) {
    const size_t total = (size_t)V * H * W;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int u = (int)(idx % W);
        const int v = (int)((idx / W) % H);
        const int view = (int)(idx / ((size_t)W * H));
        float d[3];
        pixel_direction(inv_proj + 9 * view, (float)u, (float)v, d);
        out[3 * idx] = d[0]; out[3 * idx + 1] = d[1]; out[3 * idx + 2] = d[2];
    }
}

__global__ void gather_rays_kernel(const float* __restrict__ inv_proj, const float* __restrict__ cam,
                                   const int64_t* __restrict__ pix, int R, int H, int W,
                                   float* __restrict__ origins, float* __restrict__ dirs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t idx = pix[r];
    const int u = (int)(idx % W);
    const int v = (int)((idx / W) % H);
    const int view = (int)(idx / ((int64_t)W * H));
    float d[3];
    pixel_direction(inv_proj + 9 * view, (float)u, (float)v, d);
#pragma unroll
    for (int c = 0; c < 3; ++c) { dirs[3 * r + c] = d[c]; origins[3 * r + c] = cam[3 * view + c]; }
}

// =============================================================================================
// Counter-based generator (Philox4x32-10) for in-kernel draws when no uniforms are injected.
// =============================================================================================
__device__ __forceinline__ uint4 philox4x32(uint2 key, uint4 ctr) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ float uniform01(uint64_t seed, uint32_t stream, uint32_t a, uint32_t b) {
    const uint4 x = philox4x32(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), make_uint4(a, b, stream, 0u));
    return (float)(x.x >> 8) * (1.0f / 16777216.0f);   // [0, 1)
}

// =============================================================================================
// a9: stratified placement
// =============================================================================================
__global__ void place_coarse_kernel(const float* __restrict__ bins, const float* __restrict__ jitter, uint64_t seed,
                                    const VsrdStepState* __restrict__ state, int R, int S, float* __restrict__ out) {
    if (state != nullptr) seed = state->seed;
    const size_t total = (size_t)R * S;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx % S);
        const int r = (int)(idx / S);
        const float u = jitter ? __ldg(jitter + idx) : uniform01(seed, 1u, (uint32_t)r, (uint32_t)j);
        out[idx] = lerpf_(__ldg(bins + j), __ldg(bins + j + 1), u);
    }
}

// =============================================================================================
// a10: importance placement + merge.  One warp per ray; shared memory: 4*S floats per warp.
// =============================================================================================
__device__ __forceinline__ int lower_bound_f(const float* a, int n, float v) {   // first i with a[i] >= v
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int upper_bound_f(const float* a, int n, float v) {   // first i with a[i] > v
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(kThreads) place_fine_kernel(
        const float* __restrict__ coarse_t, const float* __restrict__ coarse_w, const float* __restrict__ uniforms,
        uint64_t seed, const VsrdStepState* __restrict__ state, int R, int S, float* __restrict__ out) {
    extern __shared__ float smem[];
    if (state != nullptr) seed = state->seed;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kWarps + warp;
    if (r >= R) return;
    float* tc = smem + (size_t)warp * 4 * S;   // coarse distances
    float* cdf = tc + S;                        // [0, cumsum(pdf)]
    float* tf = cdf + S;                        // importance samples
    float* us = tf + S;                         // sorted uniforms

    float l1 = 0.0f;
    for (int k = lane; k < S; k += 32) {
        tc[k] = __ldg(coarse_t + (size_t)r * S + k);
        if (k < S - 1) l1 += fabsf(__ldg(coarse_w + (size_t)r * (S - 1) + k));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) l1 += __shfl_xor_sync(kFull, l1, o);
    const float denom = fmaxf(l1, 1e-12f);      // F.normalize(p=1)

    if (uniforms) {
        for (int k = lane; k < S; k += 32) us[k] = __ldg(uniforms + (size_t)r * S + k);
    } else {
        // Sorted uniforms without a sort: normalised partial sums of S+1 exponentials are distributed
        // as the order statistics of S uniforms.
        float carry = 0.0f;
        for (int base = 0; base <= S; base += 32) {
            const int k = base + lane;
            float e = 0.0f;
            if (k <= S) e = -logf(1.0f - uniform01(seed, 2u, (uint32_t)r, (uint32_t)k));
            float inc = e;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const float n = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += n; }
            if (k < S) us[k] = carry + inc;
            carry += __shfl_sync(kFull, inc, 31);
        }
        __syncwarp();
        for (int k = lane; k < S; k += 32) us[k] = fminf(us[k] / carry, 0.99999994f);
    }
    __syncwarp();
    // torch.cumsum on CPU accumulates float32 inputs in double and rounds each prefix (verified);
    // follow it so bin boundaries agree with the oracle.
    // Here: each lane sums a contiguous chunk in double, the chunk totals are scanned across the warp in double,
    // every prefix is rounded to float once.  (Double re-association moves a prefix by <= 1e-16 relative, far
    // below the float rounding of the result.)
    {
        const int n = S - 1, chunk = (n + 31) / 32, k0 = lane * chunk;
        double local = 0.0;
        for (int k = k0; k < min(k0 + chunk, n); ++k) local += (double)(__ldg(coarse_w + (size_t)r * n + k) / denom);
        double incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double up = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += up; }
        double acc = incl - local;                  // sum of all earlier chunks
        if (lane == 0) cdf[0] = 0.0f;
        for (int k = k0; k < min(k0 + chunk, n); ++k) {
            acc += (double)(__ldg(coarse_w + (size_t)r * n + k) / denom);
            cdf[k + 1] = (float)acc;
        }
    }
    __syncwarp();
    for (int k = lane; k < S; k += 32) {
        const float u = us[k];
        int idx = lower_bound_f(cdf, S, u);                 // searchsorted(right=False)
        idx = min(max(idx, 1), S - 1);
        const float c0 = cdf[idx - 1], c1 = cdf[idx];
        const float frac = (u - c0) / (c1 - c0 + 1e-6f);
        tf[k] = lerpf_(tc[idx - 1], tc[idx], frac);
    }
    __syncwarp();
    // merge two ascending runs (== sort of the concatenation, renderers.py:201-210)
    float* o = out + (size_t)r * 2 * S;
    for (int k = lane; k < S; k += 32) {
        o[k + lower_bound_f(tf, S, tc[k])] = tc[k];
        o[k + upper_bound_f(tc, S, tf[k])] = tf[k];
    }
}

// =============================================================================================
// a8/a11: compositing forward.  One CTA per ray, one thread per sample: the per-sample work (soft union
// over the instances, SDF -> opacity) is fully parallel, transmittance is a multiplicative scan over the
// CTA (warp shuffles + one shared-memory hop), labels / loss terms are block reductions.
// =============================================================================================
constexpr int kMaxRayWarps = VSRD_MAX_INTERVALS / 32;    // 16 warps cover 512 intervals

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Warp sums of K values at once (K a power of two <= 32): after each exchange a lane keeps half of its values, so the
// butterfly costs K - 1 + log2(32 / K) shuffles instead of 5 K.  Lane L ends up with the sum of value L / (32 / K); the
// additions are those of warp_sum()'s butterfly in the same order (bit-identical results).
template <int K>
__device__ __forceinline__ float warp_sum_multi(float (&v)[K], int lane) {
    static_assert(K >= 1 && K <= 32 && (K & (K - 1)) == 0, "K must be a power of two <= 32");
    int o = 16;
#pragma unroll
    for (int w = K; w > 1; w >>= 1, o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < w / 2; ++i) {
            const float send = upper ? v[i] : v[i + w / 2];
            const float keep = upper ? v[i + w / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, o);
        }
    }
#pragma unroll
    for (; o; o >>= 1) v[0] += __shfl_xor_sync(kFull, v[0], o);
    return v[0];
}

// inclusive multiplicative scan across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float n = __shfl_up_sync(kFull, v, o); if (lane >= o) v *= n; }
    return v;
}

// Exclusive product of (1 - alpha) over all earlier samples of the ray (transmittance).
__device__ __forceinline__ float block_transmittance(float alpha, int warp, int lane, int num_warps, float* s_tot) {
    const float inc = warp_scan_mul(1.0f - alpha, lane);
    float excl = __shfl_up_sync(kFull, inc, 1);
    if (lane == 0) excl = 1.0f;
    if (lane == 31) s_tot[warp] = inc;
    __syncthreads();
    float carry = 1.0f;
    for (int w = 0; w < warp; ++w) carry *= s_tot[w];     // sequential like the per-ray cumprod
    (void)num_warps;
    return carry * excl;
}

struct LossDev {
    const float* targets;
    float sil_w;
    float eik_w;
};

constexpr int kRegsMaxInstances = 16;

template <int NMAX, int TMAX = VSRD_MAX_INTERVALS, int MINB = (NMAX <= 8 ? 2 : 1)>
__global__ void __launch_bounds__(TMAX, MINB) composite_forward_kernel(
        SceneDev scene, RaysDev rays, float sigma, float rho, float eps, const float4* __restrict__ field,
        float* __restrict__ labels, float* __restrict__ grads, float* __restrict__ weights,
        LossDev loss, float* __restrict__ loss_out) {
    __shared__ float s_tot[kMaxRayWarps];
    __shared__ float s_lab[kMaxRayWarps][NMAX + 1];     // per-warp label partials, [NMAX] = eikonal partial
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_warps = blockDim.x >> 5;
    const int r = blockIdx.x;
    const int N = scene.N, M = rays.M;
    const size_t stride = (size_t)rays.R * M;
    float T = scene.T;
    if (scene.state != nullptr) {     // device-resident schedule (graph replay across optimisation steps)
        T = scene.state->temperature; sigma = scene.state->std_deviation; rho = scene.state->cosine_ratio;
        loss.eik_w = scene.state->eikonal_weight;
    }
    const int j = threadIdx.x;
    const bool valid = j < M;
    const size_t idx = (size_t)r * M + (valid ? j : 0);
    auto load = [&](int i) { return ld4(field + (size_t)i * stride + idx); };

    UnionEval u;
    OpacityEval o;
    float alpha = 0.0f;
    constexpr bool kRegs = NMAX <= kRegsMaxInstances;   // all instances' field values live in registers
    UnionRegs<kRegs ? NMAX : 1> ur;
    if (valid) {
        const float dir[3] = {__ldg(rays.dirs + 3 * r), __ldg(rays.dirs + 3 * r + 1), __ldg(rays.dirs + 3 * r + 2)};
        const float* trow = rays.dist + (size_t)r * (M + 1);
        if constexpr (kRegs) {
#pragma unroll
            for (int i = 0; i < NMAX; ++i) ur.f[i] = i < N ? load(i) : Vec4{0.0f, 0.0f, 0.0f, 0.0f};
            union_forward_regs<NMAX>(ur, N, T, u);
        } else {
            union_forward(load, N, T, u);
        }
        const float delta = __ldg(trow + j + 1) - __ldg(trow + j);
        opacity_forward(u, dir, delta, sigma, rho, eps, o);
        alpha = o.alpha;
    }
    const float omega = block_transmittance(alpha, warp, lane, num_warps, s_tot) * alpha;
    float eik = 0.0f;
    if (valid) {
        weights[idx] = omega;
        grads[3 * idx] = u.g[0]; grads[3 * idx + 1] = u.g[1]; grads[3 * idx + 2] = u.g[2];
        const float e = o.gn - 1.0f;
        eik = e * e;
    }
    const float k = valid ? omega / u.Z : 0.0f;
    {                                                          // all label partials of the warp in one butterfly
        float part[NMAX];
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {                       // softmin numerators exp(-d_n/T - mneg)
            if constexpr (kRegs) part[n] = (valid && n < N) ? k * ur.e[n] : 0.0f;
            else part[n] = (valid && n < N) ? k * expf(-(load(n).x / T) - u.mneg) : 0.0f;
        }
        const float mine = warp_sum_multi<NMAX>(part, lane);
        constexpr int kLanesPer = 32 / NMAX;
        if ((lane & (kLanesPer - 1)) == 0 && lane / kLanesPer < N) s_lab[warp][lane / kLanesPer] = mine;
    }
    eik = warp_sum(eik);
    if (lane == 0) s_lab[warp][NMAX] = eik;
    __syncthreads();
    if (warp == 0) {
        float mine = 0.0f;
        if (lane < N) {
            for (int w = 0; w < num_warps; ++w) mine += s_lab[w][lane];
            labels[(size_t)r * N + lane] = mine;
        }
        if (loss.targets != nullptr && loss_out != nullptr) {
            float bce = 0.0f;
            if (lane < N) {
                const float y = __ldg(loss.targets + (size_t)r * N + lane);
                const float l = fminf(fmaxf(mine, 1.0e-6f), 1.0f - 1.0e-6f);
                bce = -(y * fmaxf(logf(l), -100.0f) + (1.0f - y) * fmaxf(logf(1.0f - l), -100.0f));
            }
            bce = warp_sum(bce);
            float e = lane < num_warps ? s_lab[lane][NMAX] : 0.0f;
            e = warp_sum(e);
            if (lane == 0) {
                atomicAdd(loss_out, loss.sil_w * bce / ((float)rays.R * (float)N));
                atomicAdd(loss_out + 1, loss.eik_w * e / ((float)rays.R * (float)M));
            }
        }
    }
}

// =============================================================================================
// compositing backward.  Same layout: one CTA per ray, one thread per sample;
// the forward quantities are recomputed once and stay in registers across the two block scans
// (transmittance prefix product, suffix sum of a_k omega_k).
// =============================================================================================
template <int NMAX>
__global__ void __launch_bounds__(VSRD_MAX_INTERVALS, NMAX <= 8 ? 2 : 1) composite_backward_kernel(
        SceneDev scene, RaysDev rays, float sigma, float rho, float eps, const float4* __restrict__ field,
        const float* __restrict__ grad_labels, const float* __restrict__ grad_grads, const float* __restrict__ grad_weights,
        LossDev loss, const float* __restrict__ labels, float4* __restrict__ adjoint, int tile_shift, int tiles_per_inst) {
    __shared__ float s_tot[kMaxRayWarps];
    __shared__ float s_suf[kMaxRayWarps];
    __shared__ float s_gl[NMAX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_warps = blockDim.x >> 5;
    const int r = blockIdx.x;
    const int N = scene.N, M = rays.M;
    const size_t stride = (size_t)rays.R * M;
    float T = scene.T;
    if (scene.state != nullptr) {     // device-resident schedule (graph replay across optimisation steps)
        T = scene.state->temperature; sigma = scene.state->std_deviation; rho = scene.state->cosine_ratio;
        loss.eik_w = scene.state->eikonal_weight;
    }
    const bool fused = loss.targets != nullptr;
    const float eik_scale = fused ? loss.eik_w * 2.0f / ((float)rays.R * (float)M) : 0.0f;

    // upstream gradient w.r.t. labels[r, :] (explicit + in-kernel BCE, main.py:653-671)
    if (threadIdx.x < NMAX) {
        const int n = threadIdx.x;
        float g = 0.0f;
        if (n < N) {
            if (grad_labels) g = __ldg(grad_labels + (size_t)r * N + n);
            if (fused) {
                const float lraw = __ldg(labels + (size_t)r * N + n);
                if (lraw >= 1.0e-6f && lraw <= 1.0f - 1.0e-6f) {   // clamp passes gradient inside the range
                    const float y = __ldg(loss.targets + (size_t)r * N + n);
                    g += loss.sil_w * (lraw - y) / (lraw * (1.0f - lraw)) / ((float)rays.R * (float)N);
                }
            }
        }
        s_gl[n] = g;
    }
    __syncthreads();

    const int j = threadIdx.x;
    const bool valid = j < M;
    const size_t idx = (size_t)r * M + (valid ? j : 0);
    auto load = [&](int i) { return ld4(field + (size_t)i * stride + idx); };
    const float dir[3] = {__ldg(rays.dirs + 3 * r), __ldg(rays.dirs + 3 * r + 1), __ldg(rays.dirs + 3 * r + 2)};
    UnionEval u;
    OpacityEval o;
    float alpha = 0.0f, a = 0.0f, delta = 0.0f;
    unsigned live_mask = 0;                             // instances with a non-zero adjoint at this sample
    constexpr float kCullWeight = 2.0611537e-9f;        // exp(-VSRD_CULL_LOG_EPS)
    static_assert(VSRD_CULL_LOG_EPS == 20.0f, "kCullWeight = exp(-VSRD_CULL_LOG_EPS)");
    constexpr bool kRegs = NMAX <= kRegsMaxInstances;   // all instances' field values live in registers
    UnionRegs<kRegs ? NMAX : 1> ur;
    if (valid) {
        const float* trow = rays.dist + (size_t)r * (M + 1);
        if constexpr (kRegs) {
#pragma unroll
            for (int i = 0; i < NMAX; ++i) ur.f[i] = i < N ? load(i) : Vec4{0.0f, 0.0f, 0.0f, 0.0f};
            union_forward_regs<NMAX>(ur, N, T, u);
        } else {
            union_forward(load, N, T, u);
        }
        delta = __ldg(trow + j + 1) - __ldg(trow + j);
        opacity_forward(u, dir, delta, sigma, rho, eps, o);
        alpha = o.alpha;
        if (grad_weights) a = __ldg(grad_weights + idx);
        if constexpr (kRegs) {
#pragma unroll
            for (int n = 0; n < NMAX; ++n) if (n < N) a += s_gl[n] * (ur.e[n] * ur.invZ);
        } else {
            for (int n = 0; n < N; ++n) a += s_gl[n] * (expf(-(load(n).x / T) - u.mneg) / u.Z);
        }
    }
    const float trans = block_transmittance(alpha, warp, lane, num_warps, s_tot);
    const float omega = trans * alpha;
    // exclusive suffix sum over later samples of a_k * omega_k
    const float v = a * omega;
    float inc = v;
#pragma unroll
    for (int sh = 1; sh < 32; sh <<= 1) { const float n = __shfl_down_sync(kFull, inc, sh); if (lane + sh < 32) inc += n; }
    if (lane == 0) s_suf[warp] = inc;
    __syncthreads();
    float suffix = inc - v;
    for (int w = num_warps - 1; w > warp; --w) suffix += s_suf[w];
    if (valid) {
        const float alpha_bar = a * trans - suffix / (1.0f - alpha);
        float dbar_adj, gbar[3];
        opacity_backward(o, dir, delta, sigma, rho, eps, alpha_bar, dbar_adj, gbar);
        if (grad_grads) {
#pragma unroll
            for (int c = 0; c < 3; ++c) gbar[c] += __ldg(grad_grads + 3 * idx + c);
        }
        if (fused && o.gn > 0.0f) {      // d/dg of eik_w * mean((|g| - 1)^2), main.py:679-687
            const float k = eik_scale * (o.gn - 1.0f) / o.gn;
#pragma unroll
            for (int c = 0; c < 3; ++c) gbar[c] += k * u.g[c];
        }
        auto wbar = [&](int i) { return omega * s_gl[i]; };
        auto store = [&](int i, const Vec4& q) {
            Vec4 v = q;
            if (rays.live != nullptr) {
                // culling: an instance whose soft-min weight is below exp(-VSRD_CULL_LOG_EPS) gets an exactly zero adjoint
                float wi;
                if constexpr (kRegs) wi = ur.e[i] * ur.invZ;
                else wi = expf(-(load(i).x / T) - u.mneg) / u.Z;
                if (wi < kCullWeight) v = Vec4{0.0f, 0.0f, 0.0f, 0.0f};
                if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) live_mask |= 1u << i;
            }
            adjoint[(size_t)i * stride + idx] = make_float4(v.x, v.y, v.z, v.w);
        };
        if constexpr (kRegs) union_backward_regs<NMAX>(ur, wbar, store, N, u, dbar_adj, gbar);
        else union_backward(load, wbar, store, N, T, u, dbar_adj, gbar);
    }
    if (rays.live != nullptr) {
        // mark the backward field kernel's warp tiles (2^tile_shift consecutive samples of the flat [R*M] index) that
        // received a non-zero adjoint; every writer stores the same byte, so no atomics are needed
        const int tile = valid ? (int)(idx >> tile_shift) : -1 - lane;
        const unsigned peers = __match_any_sync(kFull, tile);
        const bool leader = valid && lane == __ffs(peers) - 1;
        // ... and count the live samples per block of tiles (census_offset()), the weights backward_ranges_kernel balances
        // the field kernel's CTAs with; a warp's samples are attributed to the block of its first one
        int* census = reinterpret_cast<int*>(rays.live + census_offset(N, tiles_per_inst));
        const int nb = census_blocks(tiles_per_inst);
        const int blk = (int)(((size_t)r * M + 32 * warp) >> tile_shift) / VSRD_CENSUS_BLOCK_TILES;
        for (int i = 0; i < N; ++i) {
            const unsigned votes = __ballot_sync(kFull, (live_mask >> i) & 1u);
            if (leader && (votes & peers)) rays.live[(size_t)i * tiles_per_inst + tile] = 1;
            if (lane == 0 && votes) atomicAdd(census + i * nb + blk, __popc(votes));
        }
    }
}


static int g_num_sms_render = 0;
static int render_setup() {
    if (g_num_sms_render) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("vsrd_b200: no CUDA device%s");
    if (cudaDeviceGetAttribute(&g_num_sms_render, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return fail("vsrd_b200: cudaDeviceGetAttribute failed%s");
    return 0;
}

}  // namespace vsrd

using namespace vsrd;

extern "C" {

int vsrd_version(void) { return 1; }

const char* vsrd_last_error(void) { return g_error; }

int vsrd_ray_directions(const float* inv_projection, int num_views, int height, int width, float* directions, void* stream) {
    VSRD_CHECK_ARG(inv_projection && directions, "NULL pointer");
    VSRD_CHECK_ARG(num_views >= 0 && height >= 0 && width >= 0, "negative size");
    const size_t total = (size_t)num_views * height * width;
    if (total == 0) return 0;
    if (render_setup()) return 1;
    const int blocks = (int)((total + 255) / 256 < (size_t)g_num_sms_render * 16 ? (total + 255) / 256 : (size_t)g_num_sms_render * 16);
    ray_directions_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(inv_projection, num_views, height, width, directions);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_gather_rays(const float* inv_projection, const float* camera_positions, const int64_t* pixel_indices,
                     int num_rays, int num_views, int height, int width, float* origins, float* directions, void* stream) {
    VSRD_CHECK_ARG(num_rays >= 0 && num_views >= 1 && height >= 1 && width >= 1, "bad size");
    if (num_rays == 0) return 0;
    VSRD_CHECK_ARG(inv_projection && camera_positions && pixel_indices && origins && directions, "NULL pointer");
    gather_rays_kernel<<<(num_rays + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        inv_projection, camera_positions, pixel_indices, num_rays, height, width, origins, directions);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_place_coarse(const float* bins, const float* jitter, uint64_t seed, const VsrdStepState* step_state,
                      int num_rays, int num_samples, float* distances, void* stream) {
    VSRD_CHECK_ARG(num_rays >= 0 && num_samples >= 1, "bad size");
    const size_t total = (size_t)num_rays * num_samples;
    if (total == 0) return 0;                              // an empty ray batch is a no-op (buffers may be NULL)
    VSRD_CHECK_ARG(bins && distances, "NULL pointer");
    if (render_setup()) return 1;
    const size_t want = (total + 255) / 256, cap = (size_t)g_num_sms_render * 16;
    place_coarse_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(bins, jitter, seed, step_state, num_rays, num_samples, distances);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_place_fine(const float* coarse_distances, const float* coarse_weights, const float* sorted_uniforms,
                    uint64_t seed, const VsrdStepState* step_state, int num_rays, int num_samples,
                    float* distances, void* stream) {
    VSRD_CHECK_ARG(num_rays >= 0, "bad size");
    VSRD_CHECK_ARG(num_samples >= 2 && 2 * num_samples - 1 <= VSRD_MAX_INTERVALS, "num_samples must be in [2, 256]");
    if (num_rays == 0) return 0;
    VSRD_CHECK_ARG(coarse_distances && coarse_weights && distances, "NULL pointer");
    const size_t smem = (size_t)kWarps * 4 * num_samples * sizeof(float);
    place_fine_kernel<<<(num_rays + kWarps - 1) / kWarps, kThreads, smem, (cudaStream_t)stream>>>(
        coarse_distances, coarse_weights, sorted_uniforms, seed, step_state, num_rays, num_samples, distances);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_composite_forward(const VsrdScene* scene, const VsrdRays* rays, const VsrdRenderParams* params,
                           const float* field, float* labels, float* gradients, float* weights,
                           const VsrdLoss* loss, float* loss_out, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    VSRD_CHECK_ARG(params != nullptr, "params is NULL");
    VSRD_CHECK_ARG(s.state != nullptr || params->std_deviation > 0.0f, "std_deviation must be positive");
    if (r.R == 0) return 0;
    VSRD_CHECK_ARG(field && labels && gradients && weights, "NULL pointer");
    LossDev l{nullptr, 0.0f, 0.0f};
    if (loss && loss->targets) {
        VSRD_CHECK_ARG(loss_out != nullptr, "loss_out is NULL while loss targets are given");
        l = LossDev{loss->targets, loss->silhouette_weight, loss->eikonal_weight};
    }
    if (render_setup()) return 1;
    const int grid = r.R, block = 32 * ((r.M + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
#define VSRD_LAUNCH_CF(NMAX) composite_forward_kernel<NMAX><<<grid, block, 0, st>>>( \
        s, r, params->std_deviation, params->cosine_ratio, params->epsilon, (const float4*)field, labels, gradients, weights, l, loss_out)
    // Many rays per SM (HBM-bound regime): 48 registers -> 6 CTAs of <= 256 threads per SM instead of 4, 3.24 -> 3.59 TB/s at
    // R = 64k (55 % of the HBM peak); at R = 1000 (7 CTAs per SM in all) the spills of that variant cost 2 us, so the
    // 64-register one stays.  (40 registers / 7 CTAs: 3.17 TB/s.)
    if (s.N <= 8 && block <= 256 && grid >= 32 * g_num_sms_render)
        composite_forward_kernel<8, 256, 5><<<grid, block, 0, st>>>(
            s, r, params->std_deviation, params->cosine_ratio, params->epsilon, (const float4*)field, labels, gradients, weights, l, loss_out);
    else if (s.N <= 8) VSRD_LAUNCH_CF(8);
    else if (s.N <= 16) VSRD_LAUNCH_CF(16);
    else VSRD_LAUNCH_CF(32);
#undef VSRD_LAUNCH_CF
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_composite_backward(const VsrdScene* scene, const VsrdRays* rays, const VsrdRenderParams* params,
                            const float* field, const float* grad_labels, const float* grad_gradients,
                            const float* grad_weights, const VsrdLoss* loss, const float* labels,
                            float* adjoint, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    VSRD_CHECK_ARG(params != nullptr, "params is NULL");
    VSRD_CHECK_ARG(s.state != nullptr || params->std_deviation > 0.0f, "std_deviation must be positive");
    if (r.R == 0) return 0;
    VSRD_CHECK_ARG(field && adjoint, "NULL pointer");
    LossDev l{nullptr, 0.0f, 0.0f};
    if (loss && loss->targets) {
        VSRD_CHECK_ARG(labels != nullptr, "labels (forward output) required for the in-kernel loss gradient");
        l = LossDev{loss->targets, loss->silhouette_weight, loss->eikonal_weight};
    }
    const int grid = r.R, block = 32 * ((r.M + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
    int tile_shift = 0, tiles_per_inst = 0;
    if (r.live != nullptr) {
        const int rows = backward_mma_tile_rows();
        if (rows < 0) return 1;
        tile_shift = rows == 16 ? 4 : 5;
        tiles_per_inst = (int)(((size_t)r.R * r.M + rows - 1) / rows);
    }
#define VSRD_LAUNCH_CB(NMAX) composite_backward_kernel<NMAX><<<grid, block, 0, st>>>( \
        s, r, params->std_deviation, params->cosine_ratio, params->epsilon, (const float4*)field, \
        grad_labels, grad_gradients, grad_weights, l, labels, (float4*)adjoint, tile_shift, tiles_per_inst)
    if (s.N <= 8) VSRD_LAUNCH_CB(8);       // (48 registers / 6 CTAs per SM, as in the forward kernel: 232 B of spills, 3.63 -> 3.46 TB/s)
    else if (s.N <= 16) VSRD_LAUNCH_CB(16);
    else VSRD_LAUNCH_CB(32);
#undef VSRD_LAUNCH_CB
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
