// Shared declarations of the vsrd_b200 translation units (see vsrd_render.cu / vsrd_field_*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vsrd_b200.h"
#include "vsrd_math.cuh"

namespace vsrd {

static_assert(kNumW == VSRD_MLP_WEIGHTS, "header / kernel MLP size mismatch");
static_assert(kGradStride == VSRD_GRAD_STRIDE, "header / kernel gradient stride mismatch");

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

extern thread_local char g_error[512];   // defined in vsrd_render.cu

inline int fail(const char* fmt, const char* detail = "") {
    snprintf(g_error, sizeof(g_error), fmt, detail);
    return 1;
}

#define VSRD_CHECK_ARG(cond, msg) do { if (!(cond)) return fail("vsrd_b200: invalid argument: %s", msg); } while (0)
#define VSRD_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) return fail("vsrd_b200: CUDA launch failed: %s", cudaGetErrorString(e_)); } while (0)

struct SceneDev {
    int N;
    const float* loc;
    const float* rot;
    const float* dim;
    const float* W;
    float T;
    float scale;
    const VsrdStepState* state;   // device-resident schedule overriding T (and the render scalars) or NULL
};

struct RaysDev {
    int R;
    int M;
    const float* origins;
    const float* dirs;
    const float* dist;
    const int* fwd_samples;           // live samples of the forward field kernel per instance (vsrd_cull_samples), or NULL (no culling)
    unsigned long long* cull_stats;   // {backward tiles skipped, visited, forward pairs skipped, visited} or NULL
    unsigned char* live;              // [N][tiles] marks of backward tiles with a non-zero adjoint, or NULL
};

constexpr float kCullLogEps = VSRD_CULL_LOG_EPS;

__device__ __forceinline__ float scene_temperature(const SceneDev& s) {
    return s.state != nullptr ? s.state->temperature : s.T;
}

__device__ __forceinline__ Vec4 ld4(const float4* p) {
    const float4 v = __ldg(p);
    return Vec4{v.x, v.y, v.z, v.w};
}

__device__ __forceinline__ void load_instance(const SceneDev& s, int i, Instance& I) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { I.t[k] = __ldg(s.loc + 3 * i + k); I.dim[k] = __ldg(s.dim + 3 * i + k); }
#pragma unroll
    for (int k = 0; k < 9; ++k) I.R[k] = __ldg(s.rot + 9 * i + k);
}

// Global (reference layout) -> shared (transposed layout, vsrd_math.cuh::staged_index).
__device__ __forceinline__ void stage_weights(const float* __restrict__ W, float* sW) {
    for (int f = threadIdx.x; f < kNumW; f += blockDim.x) sW[staged_index(f)] = __ldg(W + f);
}

// Sample position exactly as the reference forms it: o + d * ((t0 + t1) / 2), no FMA contraction
// (renderers.py:213-216).
// POINTS mode (vsrd_field_points: sphere tracing / surface normals, renderers.py:21-113): dist == NULL, M = 1
// and `origins` holds the evaluation points themselves.
__device__ __forceinline__ void sample_position(const RaysDev& rays, int r, int j, float x[3]) {
    if (rays.dist == nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = __ldg(rays.origins + 3 * ((size_t)r * rays.M + j) + c);
        return;
    }
    const float t0 = __ldg(rays.dist + (size_t)r * (rays.M + 1) + j);
    const float t1 = __ldg(rays.dist + (size_t)r * (rays.M + 1) + j + 1);
    const float mid = __fadd_rn(t0, t1) / 2.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        x[c] = __fadd_rn(__ldg(rays.origins + 3 * r + c), __fmul_rn(__ldg(rays.dirs + 3 * r + c), mid));
}

inline int check_scene(const VsrdScene* s, SceneDev& d) {
    VSRD_CHECK_ARG(s != nullptr, "scene is NULL");
    VSRD_CHECK_ARG(s->num_instances >= 1 && s->num_instances <= VSRD_MAX_INSTANCES, "num_instances must be in [1, 32]");
    VSRD_CHECK_ARG(s->locations && s->rotations && s->half_extents, "scene pointers must not be NULL");
    VSRD_CHECK_ARG(s->step_state != nullptr || s->temperature > 0.0f, "temperature must be positive");
    VSRD_CHECK_ARG(s->scale > 0.0f, "scale must be positive");
    d = SceneDev{s->num_instances, s->locations, s->rotations, s->half_extents, s->mlp_weights, s->temperature, s->scale, s->step_state};
    return 0;
}

inline int check_rays(const VsrdRays* r, RaysDev& d) {
    VSRD_CHECK_ARG(r != nullptr, "rays is NULL");
    VSRD_CHECK_ARG(r->num_rays >= 0, "num_rays must be non-negative");
    VSRD_CHECK_ARG(r->num_intervals >= 1 && r->num_intervals <= VSRD_MAX_INTERVALS, "num_intervals must be in [1, 512]");
    VSRD_CHECK_ARG(r->num_rays == 0 || (r->origins && r->directions && r->distances), "ray pointers must not be NULL");
    d = RaysDev{r->num_rays, r->num_intervals, r->origins, r->directions, r->distances, r->forward_samples, r->cull_stats, r->live_tiles};
    return 0;
}

// vsrd_field_bwd_mma.cu: tensor-core field backward for residual instances.
// Rows of VSRD_GRAD_STRIDE floats the caller must provide in `partials` (-1 on error).
int backward_mma_partial_rows(int num_instances);
int backward_mma_tile_rows();         // samples per warp tile (16 or 32); -1 on error

// VsrdRays::live_tiles: [N][tiles_per_inst] marks (bytes), then -- at the next multiple of 16 bytes -- the census:
// int32 [N][census_blocks(tiles_per_inst)] = live samples per block of VSRD_CENSUS_BLOCK_TILES tiles, accumulated by
// composite_backward_kernel next to the marks, read by backward_ranges_kernel to balance the backward field kernel's CTAs.
__host__ __device__ inline int census_blocks(int tiles_per_inst) { return (tiles_per_inst + VSRD_CENSUS_BLOCK_TILES - 1) / VSRD_CENSUS_BLOCK_TILES; }
__host__ __device__ inline size_t census_offset(int N, int tiles_per_inst) { return ((size_t)N * tiles_per_inst + 15) / 16 * 16; }
int launch_field_backward_mma(const SceneDev& s, const RaysDev& r, const float* adjoint, float* partials,
                              float* gloc, float* grot, float* gdim, float* gW, cudaStream_t st);

// vsrd_field_bwd_umma.cu: tcgen05 / TMEM field backward for residual instances (same partial-row protocol)
int launch_field_backward_umma(const SceneDev& s, const RaysDev& r, const float* adjoint, float* partials,
                               float* gloc, float* grot, float* gdim, float* gW, cudaStream_t st);

// vsrd_field_umma.cu: tcgen05 / TMEM field forward for residual instances (one thread == one sample).
int launch_field_forward_umma(const SceneDev& s, const RaysDev& r, float* field, size_t total, cudaStream_t st);

}  // namespace vsrd
