// sm_100a kernels of the VSRD silhouette-renderer hot path: per-(sample, instance) field forward entry points.
//
//   residual instances  -> field_forward_umma_kernel (vsrd_field_umma.cu: tcgen05 / TMEM, one thread == one sample)
//   box-only instances  -> field_forward_box_kernel below (warm-up steps, main.py:582-618: ~200 flop per sample)
//   union_bound_kernel     the culling bound (min box SDF over the instances)
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

// min over the instances of the BOX SDF at every sample (the culling bound, see VsrdRays::union_bound)
__global__ void union_bound_kernel(SceneDev scene, RaysDev rays, float* __restrict__ bound) {
    __shared__ Instance s_inst[VSRD_MAX_INSTANCES];       // 15 floats per instance, staged once per CTA
    for (int i = threadIdx.x; i < scene.N; i += blockDim.x) load_instance(scene, i, s_inst[i]);
    __syncthreads();
    const size_t total = (size_t)rays.R * rays.M;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int r = (int)(idx / rays.M);
    const int j = (int)(idx - (size_t)r * rays.M);
    float x[3];
    sample_position(rays, r, j, x);
    float lowest = INFINITY;
    for (int i = 0; i < scene.N; ++i) {
        BoxEval b;
        box_eval(x, s_inst[i], b);
        lowest = fminf(lowest, b.value);
    }
    bound[idx] = lowest;
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

int vsrd_union_bound(const VsrdScene* scene, const VsrdRays* rays, float* union_bound, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    const size_t total = (size_t)r.R * r.M;
    if (total == 0) return 0;
    VSRD_CHECK_ARG(union_bound != nullptr, "union_bound is NULL");
    VSRD_CHECK_ARG(total < (size_t)1 << 31, "R*M must be < 2^31");
    union_bound_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s, r, union_bound);
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
