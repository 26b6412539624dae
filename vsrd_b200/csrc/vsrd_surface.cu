// sm_100a kernels of the VSRD hot path, part 5: the inference / logging renderers.
//
//   union_points_kernel        soft union of the per-instance field at arbitrary points
//                              (scripts/main.py:477-492): value, spatial gradient and instance weights
//   sphere_trace_step_kernel   one iteration of vsrd.rendering.sphere_tracing (rendering/renderers.py:41-55),
//                              including its GLOBAL early exit: the reference leaves the loop at the first
//                              iteration where every ray is either outside the foreground or converged,
//                              which freezes rays that converged in that very iteration one update earlier
//                              than a per-ray loop would.  A device-side counter per iteration reproduces it
//                              without a host round trip per iteration.
//
// The per-instance field itself comes from vsrd_field_points (vsrd_field_fwd.cu): the same tensor-core
// kernel as the training path, in points mode.  Everything here is a few bytes per ray: HBM-bound.
#include "vsrd_common.cuh"

namespace vsrd {

namespace {

constexpr int kSurfThreads = 256;

// field [N][P] float4 -> out [P] float4 (dbar, grad), weights [P][N] (softmin weights = soft instance labels)
__global__ void __launch_bounds__(kSurfThreads) union_points_kernel(SceneDev scene, const float4* __restrict__ field, int P,
                                                                    float4* __restrict__ out, float* __restrict__ weights) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float T = scene.T;
    if (scene.state != nullptr) T = scene.state->temperature;
    const int N = scene.N;
    auto load = [&](int i) { return ld4(field + (size_t)i * P + p); };
    UnionEval u;
    union_forward(load, N, T, u);
    if (out) out[p] = make_float4(u.dbar, u.g[0], u.g[1], u.g[2]);
    if (weights)
        for (int i = 0; i < N; ++i) weights[(size_t)p * N + i] = expf(-(load(i).x / T) - u.mneg) / u.Z;
}

// state: positions [P,3], foreground [P] u8, converged [P] u8 (both 0/1); active[it] counts the rays that are
// still foreground and not converged after iteration `it`.  Iteration it > 0 is a no-op when active[it-1] == 0
// (the reference has left its loop by then).
__global__ void __launch_bounds__(kSurfThreads) sphere_trace_step_kernel(
        const float4* __restrict__ union_out, const float* __restrict__ directions, int dir_stride, int P,
        float criteria, float bounding_radius, float* __restrict__ positions, uint8_t* __restrict__ foreground,
        uint8_t* __restrict__ converged, int32_t* __restrict__ active, int iteration) {
    if (iteration > 0 && active[iteration - 1] == 0) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    int still = 0;
    if (p < P) {
        const float sd = union_out[p].x;
        bool fg = foreground[p] != 0;
        const bool conv_prev = converged[p] != 0;
        float x[3] = {positions[3 * p], positions[3 * p + 1], positions[3 * p + 2]};
        if (fg && !conv_prev) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {          // ray_positions + ray_directions * signed_distances, no contraction
                x[c] = __fadd_rn(x[c], __fmul_rn(directions[(size_t)dir_stride * p + c], sd));
                positions[3 * p + c] = x[c];
            }
        }
        if (bounding_radius > 0.0f) {
            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x[0], x[0]), __fmul_rn(x[1], x[1])), __fmul_rn(x[2], x[2])));
            fg = fg && (nrm < bounding_radius);
            foreground[p] = fg ? 1 : 0;
        }
        const bool conv = fabsf(sd) < criteria;
        converged[p] = conv ? 1 : 0;
        still = (fg && !conv) ? 1 : 0;
    }
    const int total = __syncthreads_count(still);
    if (threadIdx.x == 0 && total) atomicAdd(active + iteration, total);
}

}  // namespace

}  // namespace vsrd

using namespace vsrd;

extern "C" {

int vsrd_union_points(const VsrdScene* scene, const float* field, int num_points, float* union_out, float* weights,
                      void* stream) {
    SceneDev s;
    if (check_scene(scene, s)) return 1;
    VSRD_CHECK_ARG(num_points >= 0, "num_points must be non-negative");
    if (num_points == 0) return 0;
    VSRD_CHECK_ARG(field != nullptr, "field is NULL");
    VSRD_CHECK_ARG(union_out != nullptr || weights != nullptr, "union_out and weights are both NULL");
    const int grid = (num_points + kSurfThreads - 1) / kSurfThreads;
    union_points_kernel<<<grid, kSurfThreads, 0, (cudaStream_t)stream>>>(s, (const float4*)field, num_points,
                                                                        (float4*)union_out, weights);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_sphere_trace_step(const float* union_out, const float* directions, int directions_per_ray, int num_rays,
                           float convergence_criteria, float bounding_radius, float* positions, uint8_t* foreground,
                           uint8_t* converged, int32_t* active, int iteration, void* stream) {
    VSRD_CHECK_ARG(num_rays >= 0, "num_rays must be non-negative");
    VSRD_CHECK_ARG(iteration >= 0, "iteration must be non-negative");
    if (num_rays == 0) return 0;
    VSRD_CHECK_ARG(union_out && directions && positions && foreground && converged && active, "NULL pointer argument");
    VSRD_CHECK_ARG(convergence_criteria > 0.0f, "convergence_criteria must be positive");
    const int grid = (num_rays + kSurfThreads - 1) / kSurfThreads;
    sphere_trace_step_kernel<<<grid, kSurfThreads, 0, (cudaStream_t)stream>>>(
        (const float4*)union_out, directions, directions_per_ray ? 3 : 0, num_rays, convergence_criteria,
        bounding_radius, positions, foreground, converged, active, iteration);
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
