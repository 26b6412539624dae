// sm_100a kernels of the VSRD silhouette-renderer hot path: field backward for BOX-ONLY instances (the warm-up phase,
// main.py:582-618), the parameter-gradient reduction, and the C entry point of the field backward (residual instances
// go to vsrd_field_bwd_mma.cu).
#include "vsrd_common.cuh"

namespace vsrd {

// =============================================================================================
// field backward.  grid = (G, N), persistent over 128-sample tiles of one instance.
// =============================================================================================
struct WarpSink {
    float* acc;   // this warp's [kGradStride] accumulators (shared memory)
    int lane;

    template <int S>
    __device__ __forceinline__ void butterfly_step(float (&v)[32]) const {
        const bool up = (lane & S) != 0;
#pragma unroll
        for (int k = 0; k < S; ++k) {
            // copy first: `c ? v[a] : v[b]` on lvalues selects the ADDRESS and forces v[] into local memory
            const float lo = v[k], hi = v[k + S];
            const float send = up ? lo : hi;
            const float keep = up ? hi : lo;
            v[k] = keep + __shfl_xor_sync(kFull, send, S);
        }
    }

    // Sum value(k) over the 32 lanes for k in [G*32, G*32+32) and add the totals to acc[base + k].
    // Butterfly "transpose-reduce": 31 shuffles per 32 outputs instead of 5 per output.
    template <int COUNT, int G, class F>
    __device__ __forceinline__ void reduce_group(int base, F& value) {
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = (G * 32 + k < COUNT) ? value(G * 32 + k) : 0.0f;
        butterfly_step<16>(v);
        butterfly_step<8>(v);
        butterfly_step<4>(v);
        butterfly_step<2>(v);
        butterfly_step<1>(v);
        const int f = G * 32 + lane;
        if (f < COUNT) acc[base + f] += v[0];
    }

    template <int COUNT, int G, class F>
    __device__ __forceinline__ void reduce_groups(int base, F& value) {
        if constexpr (G * 32 < COUNT) {
            reduce_group<COUNT, G>(base, value);
            reduce_groups<COUNT, G + 1>(base, value);
        }
    }

    template <int COUNT, class F>
    __device__ __forceinline__ void reduce_add(int base, F value) {
        reduce_groups<COUNT, 0>(base, value);
    }

    __device__ __forceinline__ void layer0(const float* hbar, const float* hdbar, const float* e, const float* ed) {
        reduce_add<kHid * (kEnc + 1)>(kW0, [&](int k) {
            const int o = k / (kEnc + 1), j = k % (kEnc + 1);
            return j < kEnc ? hbar[o] * e[j < kEnc ? j : 0] + hdbar[o] * ed[j < kEnc ? j : 0] : hbar[o];
        });
    }
    __device__ __forceinline__ void hidden(int l, const float* hbar, const float* hdbar, const float* g, const float* gd) {
        reduce_add<kWStride>(kW1 + (l - 1) * kWStride, [&](int k) {
            const int o = k / (kHid + 1), i = k % (kHid + 1);
            return i < kHid ? hbar[o] * g[i < kHid ? i : 0] + hdbar[o] * gd[i < kHid ? i : 0] : hbar[o];
        });
    }
    __device__ __forceinline__ void last(float obar, float odbar, const float* g, const float* gd) {
        reduce_add<kHid + 1>(kW4, [&](int k) {
            return k < kHid ? obar * g[k < kHid ? k : 0] + odbar * gd[k < kHid ? k : 0] : obar;
        });
    }
    __device__ __forceinline__ void pose(const float* tbar, const float* dimbar, const float* Rbar) {
        reduce_add<kNumPose>(kNumW, [&](int k) {
            return k < 3 ? tbar[k < 3 ? k : 0] : (k < 6 ? dimbar[(k >= 3 && k < 6) ? k - 3 : 0] : Rbar[(k >= 6 && k < 15) ? k - 6 : 0]);
        });
    }
};

// ---------------------------------------------------------------------------------------------
// Box-only instances (warm-up steps): only the 15 pose gradients, one butterfly per tile.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) field_backward_box_kernel(
        SceneDev scene, RaysDev rays, const float4* __restrict__ adjoint, float* __restrict__ partials) {
    __shared__ float sAcc[kWarps][kGradStride];
    const int inst = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int f = threadIdx.x; f < kWarps * kGradStride; f += kThreads) (&sAcc[0][0])[f] = 0.0f;
    __syncthreads();
    Instance I;
    load_instance(scene, inst, I);
    WarpSink sink{sAcc[warp], lane};
    const size_t total = (size_t)rays.R * rays.M;
    const size_t num_tiles = (total + kThreads - 1) / kThreads;
    for (size_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const size_t warp_base = tile * kThreads + (size_t)warp * 32;
        if (warp_base >= total) continue;                     // warp-uniform
        const size_t idx = warp_base + lane;
        float x[3] = {0.0f, 0.0f, 0.0f};
        float dd = 0.0f, dG[3] = {0.0f, 0.0f, 0.0f};          // zero adjoints contribute exactly zero
        if (idx < total) {
            const int r = (int)(idx / rays.M);
            const int j = (int)(idx - (size_t)r * rays.M);
            sample_position(rays, r, j, x);
            const float4 a = __ldg(adjoint + (size_t)inst * total + idx);
            dd = a.x; dG[0] = a.y; dG[1] = a.z; dG[2] = a.w;
        }
        field_backward<false>(x, I, nullptr, scene.scale, dd, dG, sink);
    }
    __syncthreads();
    float* out = partials + ((size_t)inst * gridDim.x + blockIdx.x) * kGradStride;
    for (int f = threadIdx.x; f < kGradStride; f += kThreads) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += sAcc[w][f];
        out[f] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Residual-MLP instances.  One lane = one (sample, instance); one warp = 32 samples.
//
//   phase 1  single-tangent ("dual") forward; LayerNorm outputs z_l and their tangents zd_l are
//            stashed in this warp's shared-memory rows (one row = one feature x 32 samples).
//   phase 2  reverse sweep layer by layer.  For each layer the adjoints (hbar, hdbar) and the layer
//            inputs (g, gd) -- written in place over the consumed z/zd rows -- form the operands of
//               dW_l[o][i] = sum_samples hbar[o] g[i] + hdbar[o] gd[i]
//            which is a [16 x 64] x [64 x 17] contraction per warp: done on the tensor cores with
//            mma.sync m16n8k8 TF32, error-compensated (3xTF32) so it stays fp32-accurate.
//            The MMA accumulators persist in registers across all tiles of the CTA.
//   phase 3  one flush of the accumulators per warp, one partial row per CTA.
//
// The positional encoding is rebuilt with the double-angle recurrence from one accurate sincosf per
// coordinate: its 1e-5 absolute error is irrelevant for gradients (tolerance 1e-3) and it removes 42
// sincosf evaluations and 48 live registers per sample.
// ---------------------------------------------------------------------------------------------
// partials[N][G][kGradStride] -> per-parameter gradients; grid = (ceil(kGradStride/128), N)
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int G, float* __restrict__ gloc,
                                       float* __restrict__ grot, float* __restrict__ gdim, float* __restrict__ gW) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int inst = blockIdx.y;
    if (f >= kNumW + kNumPose) return;
    const float* p = partials + (size_t)inst * G * kGradStride + f;
    float s = 0.0f;
    for (int g = 0; g < G; ++g) s += p[(size_t)g * kGradStride];
    if (f < kNumW) { if (gW) gW[(size_t)inst * kNumW + f] = s; }
    else if (f < kNumW + 3) gloc[3 * inst + (f - kNumW)] = s;
    else if (f < kNumW + 6) gdim[3 * inst + (f - kNumW - 3)] = s;
    else grot[9 * inst + (f - kNumW - 6)] = s;
}

static int g_num_sms = 0;
static int g_bwd_blocks_per_sm = 0;       // resident CTAs per SM of the box-only kernel

static int device_setup() {
    if (g_num_sms) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("vsrd_b200: no CUDA device%s");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail("vsrd_b200: cudaGetDeviceProperties failed%s");
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_bwd_blocks_per_sm, field_backward_box_kernel, kThreads, 0);
    if (cudaGetLastError() != cudaSuccess || g_bwd_blocks_per_sm < 1)
        return fail("vsrd_b200: kernels not loadable on this device (built for sm_100a)%s");
    g_num_sms = prop.multiProcessorCount;
    return 0;
}

static int backward_blocks(int N, int R, int M) {
    const long long tiles = ((long long)R * M + kThreads - 1) / kThreads;
    long long resident = (long long)g_num_sms * g_bwd_blocks_per_sm;
    long long g = resident / (N > 0 ? N : 1);
    if (g < 1) g = 1;
    if (g > tiles) g = tiles;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace vsrd

using namespace vsrd;

extern "C" {

int vsrd_backward_tile_rows(void) { return backward_mma_tile_rows(); }

size_t vsrd_live_tiles_bytes(int num_instances, int num_rays, int num_intervals) {
    const int rows = backward_mma_tile_rows();
    if (rows < 1 || num_instances < 1 || num_rays < 0 || num_intervals < 0) return 0;
    const int tiles_per_inst = (int)(((size_t)num_rays * num_intervals + rows - 1) / rows);
    return census_offset(num_instances, tiles_per_inst) + sizeof(int32_t) * (size_t)num_instances * census_blocks(tiles_per_inst);
}

int vsrd_backward_blocks_per_instance(int num_instances, int num_rays, int num_intervals) {
    if (device_setup()) return -1;
    // box-only kernel: one row per CTA of its (G, N) grid; residual kernel: one row per (CTA, instance) segment, at most
    // #SMs + N rows.  One buffer sized for the larger of the two serves both phases of the schedule.
    const int a = backward_blocks(num_instances, num_rays, num_intervals);
    const int rows = backward_mma_partial_rows(num_instances);
    if (rows < 0) return -1;
    const int n = num_instances > 0 ? num_instances : 1;
    const int c = (rows + n - 1) / n;
    return a > c ? a : c;
}

int vsrd_field_backward(const VsrdScene* scene, const VsrdRays* rays, const float* adjoint, float* partials,
                        float* grad_locations, float* grad_rotations, float* grad_half_extents,
                        float* grad_mlp_weights, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    VSRD_CHECK_ARG((adjoint || r.R == 0) && partials && grad_locations && grad_rotations && grad_half_extents, "NULL pointer");
    VSRD_CHECK_ARG(!s.W || grad_mlp_weights, "grad_mlp_weights is NULL while mlp_weights are given");
    if (device_setup()) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t total = (size_t)r.R * r.M;
    VSRD_CHECK_ARG(total < (size_t)1 << 31, "R*M must be < 2^31");
    if (total > 0 && s.W)
        return launch_field_backward_mma(s, r, adjoint, partials, grad_locations, grad_rotations, grad_half_extents,
                                         grad_mlp_weights, st);
    int G = 1;
    if (total == 0) {
        cudaMemsetAsync(partials, 0, (size_t)s.N * kGradStride * sizeof(float), st);
    } else {
        G = backward_blocks(s.N, r.R, r.M);
        const dim3 grid((unsigned)G, (unsigned)s.N);
        field_backward_box_kernel<<<grid, kThreads, 0, st>>>(s, r, (const float4*)adjoint, partials);
        VSRD_CHECK_LAUNCH();
    }
    const dim3 rgrid((kGradStride + 127) / 128, (unsigned)s.N);
    reduce_partials_kernel<<<rgrid, 128, 0, st>>>(partials, G, grad_locations, grad_rotations, grad_half_extents,
                                                 s.W ? grad_mlp_weights : nullptr);
    VSRD_CHECK_LAUNCH();
    return 0;
}

// EXPERIMENT, not on the product path: the same contract served by the tcgen05 / TMEM kernel of
// vsrd_field_bwd_umma.cu.  Slower than the shipped kernel and its weight gradients carry bf16 staging error (2e-3).
int vsrd_experimental_field_backward_tcgen05(const VsrdScene* scene, const VsrdRays* rays, const float* adjoint, float* partials,
                                             float* grad_locations, float* grad_rotations, float* grad_half_extents,
                                             float* grad_mlp_weights, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    VSRD_CHECK_ARG(adjoint && partials && grad_locations && grad_rotations && grad_half_extents && grad_mlp_weights, "NULL pointer");
    VSRD_CHECK_ARG(s.W != nullptr, "the tcgen05 backward serves residual instances only");
    if (device_setup()) return 1;
    const size_t total = (size_t)r.R * r.M;
    VSRD_CHECK_ARG(total > 0 && total < (size_t)1 << 31, "R*M must be in [1, 2^31)");
    return launch_field_backward_umma(s, r, adjoint, partials, grad_locations, grad_rotations, grad_half_extents,
                                      grad_mlp_weights, (cudaStream_t)stream);
}

}  // extern "C"
