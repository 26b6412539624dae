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

// Instance culling, forward side (VsrdRays::forward_samples): one thread per sample.  It evaluates the BOX SDF of every
// instance there; an instance whose box SDF exceeds the lowest one + 1 (the residual's range) by more than
// VSRD_CULL_LOG_EPS * temperature has a soft-min weight < exp(-20) (include/vsrd_b200.h, VsrdRays).  For such a pair the field is the box field,
// written here, and the residual kernel never sees it; the other pairs are appended to the instance's list of live
// samples (one atomic per (CTA, instance) reserves a range, the order inside the range is the sample order; the
// counters sit one cache line apart: 6 k atomics on ONE line cost 10 us, measured).
constexpr int kCullThreads = 256;
__global__ void __launch_bounds__(kCullThreads) cull_samples_kernel(SceneDev scene, RaysDev rays, float4* __restrict__ field,
                                                                  int* __restrict__ lists) {
    __shared__ Instance s_inst[VSRD_MAX_INSTANCES];
    __shared__ float s_val[VSRD_MAX_INSTANCES][kCullThreads];
    __shared__ unsigned s_ballot[VSRD_MAX_INSTANCES][kCullThreads / 32];
    __shared__ int s_base[VSRD_MAX_INSTANCES];
    for (int i = threadIdx.x; i < scene.N; i += blockDim.x) load_instance(scene, i, s_inst[i]);
    __syncthreads();
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int total = rays.R * rays.M;
    const bool in_range = blockIdx.x * kCullThreads + t < total;
    const int idx = min(blockIdx.x * kCullThreads + t, total - 1);
    const int r = idx / rays.M;
    const int j = idx - r * rays.M;
    float x[3];
    sample_position(rays, r, j, x);
    float lowest = INFINITY;
    for (int i = 0; i < scene.N; ++i) {
        BoxEval b;
        box_eval(x, s_inst[i], b);
        s_val[i][t] = b.value;
        lowest = fminf(lowest, b.value);
    }
    const float threshold = lowest + 1.0f + kCullLogEps * scene_temperature(scene);
    unsigned live = 0;                                     // bit i: instance i needs its residual MLP at this sample
    for (int i = 0; i < scene.N; ++i) {
        const bool far = s_val[i][t] > threshold;
        if (in_range && far) {                             // rare early in the schedule, most pairs late: the box field
            const Instance& I = s_inst[i];
            BoxEval b;
            box_eval(x, I, b);
            field[(size_t)i * total + idx] = make_float4(
                b.value,
                I.R[0] * b.gp[0] + I.R[1] * b.gp[1] + I.R[2] * b.gp[2],
                I.R[3] * b.gp[0] + I.R[4] * b.gp[1] + I.R[5] * b.gp[2],
                I.R[6] * b.gp[0] + I.R[7] * b.gp[1] + I.R[8] * b.gp[2]);
        }
        const unsigned votes = __ballot_sync(kFull, in_range && !far);
        if (lane == 0) s_ballot[i][warp] = votes;
        live |= (in_range && !far) ? 1u << i : 0u;
    }
    __syncthreads();
    if (t < scene.N) {
        int count = 0;
#pragma unroll
        for (int w = 0; w < kCullThreads / 32; ++w) count += __popc(s_ballot[t][w]);
        s_base[t] = count ? atomicAdd(lists + t * VSRD_CULL_COUNT_STRIDE, count) : 0;
    }
    if (t == VSRD_MAX_INSTANCES && rays.cull_stats != nullptr) {
        int kept = 0;
        for (int i = 0; i < scene.N; ++i)
#pragma unroll
            for (int w = 0; w < kCullThreads / 32; ++w) kept += __popc(s_ballot[i][w]);
        const int visited = scene.N * min(kCullThreads, total - (int)blockIdx.x * kCullThreads);
        atomicAdd(rays.cull_stats + 2, (unsigned long long)(visited - kept));
        atomicAdd(rays.cull_stats + 3, (unsigned long long)visited);
    }
    __syncthreads();
    for (int i = 0; i < scene.N; ++i) {
        if (!((live >> i) & 1u)) continue;
        int pos = s_base[i] + __popc(s_ballot[i][warp] & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += __popc(s_ballot[i][w]);
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
    static_assert(VSRD_MAX_INSTANCES < kCullThreads, "cull_samples_kernel: one thread per instance plus one for the statistics");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(forward_samples, 0, VSRD_CULL_HEADER_INTS * sizeof(int32_t), st) != cudaSuccess)
        return fail("vsrd_b200: cudaMemsetAsync failed%s");
    cull_samples_kernel<<<(unsigned)((total + kCullThreads - 1) / kCullThreads), kCullThreads, 0, st>>>(s, r, (float4*)field, forward_samples);
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
