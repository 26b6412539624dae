// sm_100a kernels of the VSRD silhouette-renderer hot path: per-(sample, instance) field forward entry points.
//
//   residual instances  -> field_forward_umma_kernel (vsrd_field_umma.cu: tcgen05 / TMEM, one thread == one sample)
//   box-only instances  -> field_forward_box_kernel below (warm-up steps, main.py:582-618: ~200 flop per sample)
//   cull_samples_kernel    instance culling: box field of the culled (sample, instance) pairs + per-instance lists of the live samples
#include "vsrd_common.cuh"

namespace vsrd {

// one thread per (sample, instance): box SDF value and world-frame gradient (sdfs.py:5-37).  grid = (ceil(R*M/128), N)
__global__ void __launch_bounds__(kThreads) field_forward_box_kernel(SceneDev scene, RaysDev rays, float4* __restrict__ field) {
    const int inst = blockIdx.y;
    const size_t total = (size_t)rays.R * rays.M;
    const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= total) return;
    const int r = (int)(idx / rays.M);
    const int j = (int)(idx - (size_t)r * rays.M);
    Instance I;
    load_instance(scene, inst, I);
    float x[3];
    sample_position(rays, r, j, x);
    BoxEval b;
    box_eval(x, I, b);
    field[(size_t)inst * total + idx] = make_float4(
        b.value,
        I.R[0] * b.gp[0] + I.R[1] * b.gp[1] + I.R[2] * b.gp[2],
        I.R[3] * b.gp[0] + I.R[4] * b.gp[1] + I.R[5] * b.gp[2],
        I.R[6] * b.gp[0] + I.R[7] * b.gp[1] + I.R[8] * b.gp[2]);
}

// Instance culling, forward side (VsrdRays::forward_samples): one CTA per ray, one thread per sample.  A thread evaluates
// the BOX SDF of every instance at its sample; an instance whose box SDF exceeds the lowest one + 1 (the residual's range)
// by more than VSRD_CULL_LOG_EPS * temperature has a soft-min weight < exp(-20) (include/vsrd_b200.h, VsrdRays).  For such
// a pair the field is the box field, written here, and the residual kernel never sees it; the other pairs are appended
// to the instance's list of live samples (one atomic per (CTA, instance) reserves a range, the order inside the range is
// the sample order; the counters sit one cache line apart).
// Along a ray the local position is affine in the distance, p = R^T (o - t) + mid * R^T d: the per-(ray, instance)
// coefficients are staged once per CTA, so a box evaluation costs three FMAs instead of a rotation (2.4x fewer
// instructions than evaluating box_eval() per pair; the pre-pass is pure overhead whenever little is culled).  These
// positions differ from the reference's o + d * mid by rounding (~1e-6 m), which only ever reaches the output through
// pairs whose weight is below exp(-20); the test itself keeps a 1e-3 m slack.
constexpr float kCullSlack = 1.0e-3f;
struct RayBox {                                            // per (ray, instance), 5 x float4 in shared memory
    float4 p0_dx;     // R^T (o - t), half extent x
    float4 pd_dy;     // R^T d,       half extent y
    float4 r0_dz;     // R row 0,     half extent z
    float4 r1, r2;    // R rows 1, 2
};

struct BoxValue { float p[3], q[3], a[3], nrm, mx, value; };
__device__ __forceinline__ void ray_box_value(const RayBox& B, float mid, BoxValue& v) {
    v.p[0] = fmaf(mid, B.pd_dy.x, B.p0_dx.x); v.p[1] = fmaf(mid, B.pd_dy.y, B.p0_dx.y); v.p[2] = fmaf(mid, B.pd_dy.z, B.p0_dx.z);
    v.q[0] = fabsf(v.p[0]) - B.p0_dx.w; v.q[1] = fabsf(v.p[1]) - B.pd_dy.w; v.q[2] = fabsf(v.p[2]) - B.r0_dz.w;
    float sumsq = 1e-6f;
#pragma unroll
    for (int k = 0; k < 3; ++k) { v.a[k] = fmaxf(v.q[k], 0.0f); sumsq = fmaf(v.a[k], v.a[k], sumsq); }
    v.nrm = sqrtf(sumsq);
    v.mx = fmaxf(v.q[0], fmaxf(v.q[1], v.q[2]));
    v.value = v.nrm - fmaxf(-v.mx, 0.0f);
}

template <int NMAX>     // NMAX <= 8: the instances' box values stay in registers between the two passes; else they are recomputed
__global__ void __launch_bounds__(VSRD_MAX_INTERVALS) cull_samples_kernel(SceneDev scene, RaysDev rays, float4* __restrict__ field,
                                                                        int* __restrict__ lists) {
    constexpr bool kRegs = NMAX <= 8;
    __shared__ RayBox s_box[kRegs ? NMAX : VSRD_MAX_INSTANCES];
    __shared__ unsigned s_ballot[kRegs ? NMAX : VSRD_MAX_INSTANCES][VSRD_MAX_INTERVALS / 32];
    __shared__ int s_base[kRegs ? NMAX : VSRD_MAX_INSTANCES];
    __shared__ int s_prefix[kRegs ? NMAX : VSRD_MAX_INSTANCES][VSRD_MAX_INTERVALS / 32];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, num_warps = blockDim.x >> 5;
    const int r = blockIdx.x, N = scene.N, M = rays.M;
    const int total = rays.R * M;
    if (t < N) {
        Instance I;
        load_instance(scene, t, I);
        float y[3], d[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) { y[m] = __ldg(rays.origins + 3 * r + m) - I.t[m]; d[m] = __ldg(rays.dirs + 3 * r + m); }
        RayBox B;
        B.p0_dx = make_float4(y[0] * I.R[0] + y[1] * I.R[3] + y[2] * I.R[6], y[0] * I.R[1] + y[1] * I.R[4] + y[2] * I.R[7],
                              y[0] * I.R[2] + y[1] * I.R[5] + y[2] * I.R[8], I.dim[0]);
        B.pd_dy = make_float4(d[0] * I.R[0] + d[1] * I.R[3] + d[2] * I.R[6], d[0] * I.R[1] + d[1] * I.R[4] + d[2] * I.R[7],
                              d[0] * I.R[2] + d[1] * I.R[5] + d[2] * I.R[8], I.dim[1]);
        B.r0_dz = make_float4(I.R[0], I.R[1], I.R[2], I.dim[2]);
        B.r1 = make_float4(I.R[3], I.R[4], I.R[5], 0.0f);
        B.r2 = make_float4(I.R[6], I.R[7], I.R[8], 0.0f);
        s_box[t] = B;
    }
    __syncthreads();
    const bool valid = t < M;
    const int j = valid ? t : M - 1;
    const int idx = r * M + j;
    const float* trow = rays.dist + (size_t)r * (M + 1);
    const float mid = __fadd_rn(__ldg(trow + j), __ldg(trow + j + 1)) / 2.0f;
    float val[kRegs ? NMAX : 1];
    float lowest = INFINITY;
    if constexpr (kRegs) {
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
            val[i] = INFINITY;
            if (i < N) { BoxValue v; ray_box_value(s_box[i], mid, v); val[i] = v.value; lowest = fminf(lowest, v.value); }
        }
    } else {
        for (int i = 0; i < N; ++i) { BoxValue v; ray_box_value(s_box[i], mid, v); lowest = fminf(lowest, v.value); }
    }
    const float threshold = lowest + 1.0f + kCullLogEps * scene_temperature(scene) + kCullSlack;
    unsigned live = 0;                                     // bit i: instance i needs its residual MLP at this sample
    auto decide = [&](int i, float value) {
        const bool far = value > threshold;
        if (valid && far) {                                // the box field of the pair (vsrd_math.cuh::box_eval, in the ray's affine form)
            const RayBox& B = s_box[i];
            BoxValue v;
            ray_box_value(B, mid, v);
            const float inv = 1.0f / v.nrm;
            const bool inside = v.mx < 0.0f;
            const int kmax = v.q[1] > v.q[0] ? (v.q[2] > v.q[1] ? 2 : 1) : (v.q[2] > v.q[0] ? 2 : 0);   // first maximal index
            float gp[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float sgn = (v.p[k] > 0.0f ? 1.0f : 0.0f) - (v.p[k] < 0.0f ? 1.0f : 0.0f);
                gp[k] = sgn * (v.a[k] * inv + ((inside && k == kmax) ? 1.0f : 0.0f));
            }
            field[(size_t)i * total + idx] = make_float4(
                v.value,
                B.r0_dz.x * gp[0] + B.r0_dz.y * gp[1] + B.r0_dz.z * gp[2],
                B.r1.x * gp[0] + B.r1.y * gp[1] + B.r1.z * gp[2],
                B.r2.x * gp[0] + B.r2.y * gp[1] + B.r2.z * gp[2]);
        }
        const unsigned votes = __ballot_sync(kFull, valid && !far);
        if (lane == 0) s_ballot[i][warp] = votes;
        live |= (valid && !far) ? 1u << i : 0u;
    };
    if constexpr (kRegs) {
#pragma unroll
        for (int i = 0; i < NMAX; ++i) if (i < N) decide(i, val[i]);
    } else {
        for (int i = 0; i < N; ++i) { BoxValue v; ray_box_value(s_box[i], mid, v); decide(i, v.value); }
    }
    __syncthreads();
    if (t < N) {                                           // votes -> exclusive prefix over the warps, range reserved by one atomic
        int count = 0;
        for (int w = 0; w < num_warps; ++w) { const int c = __popc(s_ballot[t][w]); s_prefix[t][w] = count; count += c; }
        s_base[t] = count ? atomicAdd(lists + t * VSRD_CULL_COUNT_STRIDE, count) : 0;
    }
    __syncthreads();
    for (int i = 0; i < N; ++i) {
        if (!((live >> i) & 1u)) continue;
        const int pos = s_base[i] + s_prefix[i][warp] + __popc(s_ballot[i][warp] & ((1u << lane) - 1u));
        lists[VSRD_CULL_HEADER_INTS + (size_t)i * total + pos] = idx;
    }
}

}  // namespace vsrd

using namespace vsrd;

extern "C" {

static int launch_field(const SceneDev& s, const RaysDev& r, float* field, void* stream) {
    const size_t total = (size_t)r.R * r.M;
    if (total == 0) return 0;
    VSRD_CHECK_ARG(field != nullptr, "field is NULL");
    VSRD_CHECK_ARG(total < (size_t)1 << 31, "R*M must be < 2^31");
    cudaStream_t st = (cudaStream_t)stream;
    if (s.W) {                                             // residual instances: tcgen05 kernel (vsrd_field_umma.cu)
        if (launch_field_forward_umma(s, r, field, total, st)) return 1;
    } else {                                               // box-only warm-up phase
        const dim3 grid((unsigned)((total + kThreads - 1) / kThreads), (unsigned)s.N);
        field_forward_box_kernel<<<grid, kThreads, 0, st>>>(s, r, (float4*)field);
    }
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_field_forward(const VsrdScene* scene, const VsrdRays* rays, float* field, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    return launch_field(s, r, field, stream);
}

int vsrd_cull_samples(const VsrdScene* scene, const VsrdRays* rays, float* field, int32_t* forward_samples, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    const size_t total = (size_t)r.R * r.M;
    if (total == 0) return 0;
    VSRD_CHECK_ARG(field != nullptr && forward_samples != nullptr, "NULL pointer");
    VSRD_CHECK_ARG(total < (size_t)1 << 31, "R*M must be < 2^31");
    VSRD_CHECK_ARG(r.dist != nullptr, "culling needs ray samples (not the points mode of vsrd_field_points)");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(forward_samples, 0, VSRD_CULL_HEADER_INTS * sizeof(int32_t), st) != cudaSuccess)
        return fail("vsrd_b200: cudaMemsetAsync failed%s");
    const int threads = max(32, (r.M + 31) / 32 * 32);     // >= VSRD_MAX_INSTANCES threads: one per instance
    if (s.N <= 8) cull_samples_kernel<8><<<r.R, threads, 0, st>>>(s, r, (float4*)field, forward_samples);
    else cull_samples_kernel<VSRD_MAX_INSTANCES><<<r.R, threads, 0, st>>>(s, r, (float4*)field, forward_samples);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_field_points(const VsrdScene* scene, const float* points, int num_points, float* field, void* stream) {
    SceneDev s;
    if (check_scene(scene, s)) return 1;
    VSRD_CHECK_ARG(num_points >= 0, "num_points must be non-negative");
    VSRD_CHECK_ARG(num_points == 0 || points != nullptr, "points is NULL");
    const RaysDev r{num_points, 1, points, nullptr, nullptr, nullptr, nullptr, nullptr};     // points mode of sample_position()
    return launch_field(s, r, field, stream);
}

}  // extern "C"
