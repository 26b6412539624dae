// sm_100a kernels of the VSRD silhouette-renderer hot path, part 2/3: per-(sample, instance) field
// forward (box SDF + residual MLP, value and spatial gradient).
#include "vsrd_common.cuh"

namespace vsrd {

// =============================================================================================
// a5-a8: field forward.  grid = (ceil(R*M / 128), N)
// =============================================================================================
template <bool kResidual>
__global__ void __launch_bounds__(kThreads) field_forward_kernel(SceneDev scene, RaysDev rays, float4* __restrict__ field) {
    __shared__ __align__(16) float sW[kResidual ? kNumW : 4];
    __shared__ float sAct[kResidual ? kThreads * kFwdActRows : 1];   // [warp][row][lane]: lane-private, conflict-free
    const int inst = blockIdx.y;
    if (kResidual) {
        stage_weights(scene.W + (size_t)inst * kNumW, sW);
        __syncthreads();
    }
    const size_t total = (size_t)rays.R * rays.M;
    const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= total) return;
    const int r = (int)(idx / rays.M);
    const int j = (int)(idx - (size_t)r * rays.M);
    Instance I;
    load_instance(scene, inst, I);
    float x[3], d, G[3];
    sample_position(rays, r, j, x);
    float* act = sAct + (kResidual ? (threadIdx.x >> 5) * 32 * kFwdActRows + (threadIdx.x & 31) : 0);
    field_forward_looped<kResidual>(x, I, sW, scene.scale, act, 32, d, G);
    field[(size_t)inst * total + idx] = make_float4(d, G[0], G[1], G[2]);
}

}  // namespace vsrd

using namespace vsrd;

extern "C" {

int vsrd_field_forward(const VsrdScene* scene, const VsrdRays* rays, float* field, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    VSRD_CHECK_ARG(field != nullptr, "field is NULL");
    const size_t total = (size_t)r.R * r.M;
    if (total == 0) return 0;
    VSRD_CHECK_ARG(total < (size_t)1 << 31, "R*M must be < 2^31");
    const dim3 grid((unsigned)((total + kThreads - 1) / kThreads), (unsigned)s.N);
    if (s.W) field_forward_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(s, r, (float4*)field);
    else field_forward_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(s, r, (float4*)field);
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
