// sm_100a field FORWARD kernel on the 5th-generation tensor cores (tcgen05 + TMEM): per-(sample, instance) box SDF +
// residual MLP, value and spatial gradient (hyper_distance_field.py:57-73,
// sinusoidal_encoder.py:14-19, sdfs.py:5-37, main.py:433-458).
//
// Layout: ONE THREAD == ONE SAMPLE.  A tile is 128 consecutive samples of one instance, owned by a group of four
// warps (TMEM lane == sample).  Per tile the group makes seven round trips
//
//     registers --tcgen05.st--> A operand in TMEM --tcgen05.mma (3xTF32, weights in shared memory)--> D in TMEM
//               --tcgen05.ld--> registers of the thread that owns the sample
//
//   L0  positional encoding e[48]      -> h0 = W0 e + b0 (16)  AND  g_c = dh0 / da_c (3 x 16, derivative weights W0'_c:
//                                         the chain through sin / cos is linear in e, so d/dx costs no extra sweep in)
//   L1..L3  a_l = gelu(LayerNorm(h))   -> h_l = W_l a_l + b_l          (bias = one extra k-step against a ones column)
//   (layer 4, 16 -> 1, and its adjoint stay on the SIMT pipes)
//   R3..R1  hbar = LayerNorm/GELU adjoint -> abar = W_l^T hbar         (reverse sweep for d out / d h0)
//   d out / d a_c = hbar0 . g_c
//
// LayerNorm / GELU are purely per-thread (16 channels in 16 registers): no shuffles, no fragment layouts, no MMA issue
// slots on the SIMT pipes.  Three groups per CTA (one CTA per SM) keep the tensor pipe and the LSU/TMEM paths busy
// while a group waits for its MMAs.  Measured building blocks: tools/umma_probe.cu (profiles/r02_umma_probe.txt).
#include "vsrd_umma_field.cuh"

namespace vsrd {
namespace fu {

constexpr int kMaxGroups = 4;                // tile groups (4 warps each) per CTA: 4 x 128 TMEM columns = all 512

// ---- TMEM columns of one group (128) ---------------------------------------------------------------------------
// L0 streams the three PE coordinates through two operand buffers (hi 16 + lo 16 columns each): coordinates 0 and 1
// go first, coordinate 2 reuses buffer 0 once their MMAs have completed (its encoding is computed meanwhile).
constexpr int kColBuf0 = 0, kColBuf1 = 32;   // after L0: A hi [0,16), A lo [16,32), D [32,48)
constexpr int kColAhi = 0, kColAlo = 16, kColD = 32;
constexpr int kColH0 = 64;                   // h0 [64,80)
constexpr int kColG = 80;                    // g_x, g_y, g_z [80,128)
constexpr int kColsPerGroup = 128;

// ---- shared-memory B operands (K-major, no swizzle; one [16 x 16] block = 256 floats, one [16 x 8] block = 128) ----
// every block is stored hi then lo
constexpr int kOffW0 = 0;                                 // [c] value weights of coordinate c:        3 x 2 x 256
constexpr int kOffW0d = kOffW0 + 3 * 2 * kBlk;            // [c] derivative weights W0'_c:             3 x 2 x 256
constexpr int kOffWl = kOffW0d + 3 * 2 * kBlk;            // [l-1] hidden weights W_l, l = 1..3:       3 x 2 x 256
constexpr int kOffWt = kOffWl + 3 * 2 * kBlk;             // [l-1] transposed hidden weights W_l^T:    3 x 2 x 256
constexpr int kOffBias = kOffWt + 3 * 2 * kBlk;           // biases of layers 0..3 (plain fp32, added on the SIMT side): 4 x 16
constexpr int kOffTail = kOffBias + 4 * 16;               // w4[16], b4
constexpr int kWeightFloats = kOffTail + 32;
constexpr int kStashFloats = 3 * 32;                      // per thread: layers 1..3 x (z[16], gelu'(z) / sigma [16])
constexpr size_t smem_bytes(int groups) { return 1024 + (size_t)kWeightFloats * 4 + (size_t)groups * kGroupThreads * kStashFloats * 4; }

// Global (reference layout: per layer [out][in + 1], bias last) -> the B operand blocks above.
//
// LayerNorm centring folded into the weights: every layer's output feeds a LayerNorm (no affine), whose first step
// subtracts the mean over the 16 channels.  h - mean(h) = (P W) a + P b with P = I - 11^T / 16, so the operands staged
// here are the CENTRED weights P W_l, P b_l (l = 0..3; column means over the output index), the contractions deliver
// centred pre-activations, and norm_gelu() / norm_gelu_adjoint() skip the mean and its adjoint (the adjoint w.r.t. the
// centred h times (P W)^T equals the full adjoint times W^T).  Saves 17 + 18 of ~260 instructions per (sample, layer).
constexpr int kNumMeans = (kEnc + 1) + 3 * (kHid + 1);       // column means: layer 0 (48 inputs + bias), layers 1..3 (16 + bias)
__device__ void stage_weights_umma(const float* __restrict__ W, float* sW, float* s_mean) {
    for (int i = threadIdx.x; i < kNumMeans; i += blockDim.x) {
        const bool first = i <= kEnc;
        const int l = first ? 0 : (i - (kEnc + 1)) / (kHid + 1), j = first ? i : (i - (kEnc + 1)) % (kHid + 1);
        const float* col = first ? W + kW0 + j : W + kW1 + l * kWStride + j;
        const int stride = first ? kEnc + 1 : kHid + 1;
        float sum = 0.0f;
#pragma unroll
        for (int o = 0; o < kHid; ++o) sum += __ldg(col + o * stride);
        s_mean[i] = sum * (1.0f / kHid);
    }
    __syncthreads();
    const float* mean0 = s_mean;                                                          // [kEnc + 1]
    const float* meanl = s_mean + (kEnc + 1);                                             // [3][kHid + 1]
    for (int i = threadIdx.x; i < kHid * kEnc; i += blockDim.x) {                         // layer 0
        const int o = i / kEnc, j = i % kEnc, c = j / 16, jj = j % 16, k = jj >> 1;
        const float w = __ldg(W + kW0 + o * (kEnc + 1) + j) - mean0[j];
        put_split(sW + kOffW0 + c * 2 * kBlk, o, jj, w);
        // e = (cos(2^k a), sin(2^k a)):  d/da of  w_cos cos + w_sin sin  =  2^k (w_sin cos - w_cos sin)
        const float f = (float)(1 << k);
        if (jj & 1) put_split(sW + kOffW0d + c * 2 * kBlk, o, jj - 1, f * w);             // sin weight multiplies cos
        else put_split(sW + kOffW0d + c * 2 * kBlk, o, jj + 1, -f * w);                   // cos weight multiplies -sin
    }
    for (int i = threadIdx.x; i < 4 * kHid; i += blockDim.x) {                            // biases of layers 0..3
        const int l = i / kHid, o = i % kHid;
        sW[kOffBias + i] = l == 0 ? __ldg(W + kW0 + o * (kEnc + 1) + kEnc) - mean0[kEnc]
                                  : __ldg(W + kW1 + (l - 1) * kWStride + o * (kHid + 1) + kHid) - meanl[(l - 1) * (kHid + 1) + kHid];
    }
    for (int i = threadIdx.x; i < 3 * kHid * kHid; i += blockDim.x) {                     // hidden layers and transposes
        const int l = i / (kHid * kHid), o = (i / kHid) % kHid, in = i % kHid;
        const float w = __ldg(W + kW1 + l * kWStride + o * (kHid + 1) + in) - meanl[l * (kHid + 1) + in];
        put_split(sW + kOffWl + l * 2 * kBlk, o, in, w);
        put_split(sW + kOffWt + l * 2 * kBlk, in, o, w);
    }
    for (int i = threadIdx.x; i < kHid + 1; i += blockDim.x) sW[kOffTail + i] = __ldg(W + kW4 + i);
}

// The MMAs of the eight round trips of a tile, issued by ONE thread of the group once the operands of the stage are in TMEM:
//   0: L0, coordinates 0 and 1    1: L0, coordinate 2    2..4: hidden layers 1..3    5..7: reverse sweep through layers 3..1
__device__ __forceinline__ void issue_stage(int stage, uint32_t tmem, uint64_t wdesc, uint32_t idesc, uint64_t* mbar) {
    fence_after_sync();
    if (stage == 0) {
        mma3_16x16(tmem + kColH0, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0, idesc, false);
        mma3_16x16(tmem + kColG, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0d, idesc, false);
        mma3_16x16(tmem + kColH0, tmem + kColBuf1, tmem + kColBuf1 + 16, wdesc, kOffW0 + 2 * kBlk, idesc, true);
        mma3_16x16(tmem + kColG + 16, tmem + kColBuf1, tmem + kColBuf1 + 16, wdesc, kOffW0d + 2 * kBlk, idesc, false);
    } else if (stage == 1) {
        mma3_16x16(tmem + kColH0, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0 + 4 * kBlk, idesc, true);
        mma3_16x16(tmem + kColG + 32, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0d + 4 * kBlk, idesc, false);
    } else if (stage <= 4) {
        mma3_16x16(tmem + kColD, tmem + kColAhi, tmem + kColAlo, wdesc, kOffWl + (stage - 2) * 2 * kBlk, idesc, false);
    } else {
        mma3_16x16(tmem + kColD, tmem + kColAhi, tmem + kColAlo, wdesc, kOffWt + (7 - stage) * 2 * kBlk, idesc, false);
    }
    mma_commit(mbar);
}

template <int kGroups, bool kCull>
__global__ void __launch_bounds__(kGroups * kGroupThreads, 1) field_forward_umma_kernel(
        SceneDev scene, RaysDev rays, float4* __restrict__ field, int tiles_per_inst) {
    static_assert(kGroups >= 1 && kGroups <= kMaxGroups, "TMEM columns");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t s_mbar[kMaxGroups];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_mean[kNumMeans];
    float* sW = reinterpret_cast<float*>(smem_raw);
    float* sStash = sW + kWeightFloats;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int group = warp >> 2, wq = warp & 3, gt = tid & (kGroupThreads - 1);
    // per group: [pair index][thread] of float2 (channel pairs): a warp's 32 lanes touch 256 contiguous bytes
    float* stash = sStash + (size_t)group * kGroupThreads * kStashFloats + 2 * gt;
    constexpr int kSS = kGroupThreads;                                                    // stride between pairs, in float2

    if (warp == 0) tmem_alloc<512>(&s_tmem_base);
    if (tid == 0) {
#pragma unroll
        for (int g = 0; g < kGroups; ++g) mbar_init(&s_mbar[g], 1);
        mbar_fence_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem_base + group * kColsPerGroup;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * wq) << 16);
    const uint32_t idesc = make_idesc_tf32(128, 16);
    const uint64_t wdesc = make_smem_desc(smem_u32(sW), kLbo, kSbo);      // descriptor of the weight arena's base
    uint64_t* mbar = &s_mbar[group];
    uint32_t parity = 0;

    const int total = rays.R * rays.M;
    // Tile source.  Dense: all N x tiles_per_inst tiles of 128 consecutive samples in (instance, tile) order.  Culled
    // (kCull): cull_samples_kernel has listed the live samples of every instance (and has already written the box field
    // of the others); a tile is then 128 consecutive ENTRIES of an instance's list -- a thread computes its sample from
    // (ray, interval) alone, so gathering costs nothing -- and every CTA gets the same number of live tiles.
    const int* counts = rays.fwd_samples;                  // [i * VSRD_CULL_COUNT_STRIDE] live samples of instance i
    const int* lists = kCull ? rays.fwd_samples + VSRD_CULL_HEADER_INTS : nullptr;   // [N][R * M] their flat indices
    auto live_tiles = [&](int i) { return (long long)((__ldg(counts + i * VSRD_CULL_COUNT_STRIDE) + kTile - 1) / kTile); };
    long long all_tiles = 0;
    if constexpr (kCull) {
        for (int i = 0; i < scene.N; ++i) all_tiles += live_tiles(i);
        if (blockIdx.x == 0 && tid == 0 && rays.cull_stats != nullptr) {     // {.., .., forward pairs skipped, visited}
            long long kept = 0;
            for (int i = 0; i < scene.N; ++i) kept += __ldg(counts + i * VSRD_CULL_COUNT_STRIDE);
            atomicAdd(rays.cull_stats + 2, (unsigned long long)((long long)scene.N * total - kept));
            atomicAdd(rays.cull_stats + 3, (unsigned long long)((long long)scene.N * total));
        }
    } else {
        all_tiles = (long long)scene.N * tiles_per_inst;
    }
    const long long begin = all_tiles * blockIdx.x / gridDim.x;
    const long long end = all_tiles * (blockIdx.x + 1) / gridDim.x;
    const float pi_scale = kPiF / scene.scale;
    int inst = 0;
    long long inst_first = 0;                              // index of the instance's first tile in the tile order

    for (long long seg = begin; seg < end;) {
        long long inst_tiles;
        if constexpr (kCull) {
            while (seg >= inst_first + live_tiles(inst)) { inst_first += live_tiles(inst); ++inst; }
            inst_tiles = live_tiles(inst);
        } else {
            inst = (int)(seg / tiles_per_inst);
            inst_first = (long long)inst * tiles_per_inst;
            inst_tiles = tiles_per_inst;
        }
        const long long seg_end = min(end, inst_first + inst_tiles);
        fence_before_sync();
        __syncthreads();                                   // every group is done with the previous instance's weights
        stage_weights_umma(scene.W + (size_t)inst * kNumW, sW, s_mean);
        fence_proxy_async_smem();
        __syncthreads();
        fence_after_sync();
        Instance I;
        load_instance(scene, inst, I);

#pragma unroll 1
        for (long long tile = seg + group; tile < seg_end; tile += kGroups) {
            const int base = (int)(tile - inst_first) * kTile;
            bool in_range;
            int idx;
            if constexpr (kCull) {
                const int count = __ldg(counts + inst * VSRD_CULL_COUNT_STRIDE);
                in_range = base + gt < count;
                idx = __ldg(lists + (size_t)inst * total + min(base + gt, count - 1));
            } else {
                in_range = base + gt < total;
                idx = min(base + gt, total - 1);
            }
            const int r = idx / rays.M;
            const int j = idx - r * rays.M;
            float x[3];
            sample_position(rays, r, j, x);
            BoxEval b;
            box_eval(x, I, b);
            // ------------------------------------------------------------ L0: positional encoding -> h0, g_c
            f2 h[8];
            {
                const float m[3] = {fabsf(b.p[0]), b.p[1], b.p[2]};
                f2 e[8];
                encode16(kPiF * (m[0] / scene.scale), e);
                store_operand(lane_base, kColBuf0, kColBuf0 + 16, e);
                encode16(kPiF * (m[1] / scene.scale), e);
                store_operand(lane_base, kColBuf1, kColBuf1 + 16, e);
                wait_st();
                fence_before_sync();
                named_barrier(1 + group, kGroupThreads);
                if (gt == 0) issue_stage(0, tmem, wdesc, idesc, mbar);
                encode16(kPiF * (m[2] / scene.scale), e);              // overlaps the MMAs of coordinates 0, 1
                mbar_wait(mbar, parity); parity ^= 1;                  // ... which must be done before buffer 0 is reused
                fence_after_sync();
                store_operand(lane_base, kColBuf0, kColBuf0 + 16, e);
                wait_st();
                fence_before_sync();
                named_barrier(1 + group, kGroupThreads);
                if (gt == 0) issue_stage(1, tmem, wdesc, idesc, mbar);
            }
            mbar_wait(mbar, parity); parity ^= 1;
            fence_after_sync();
            load_biased(lane_base + kColH0, sW + kOffBias, h);
            // ------------------------------------------------------------ L1..L3 + layer 4 on the SIMT pipes
            f2 abar[8];                                    // adjoint of h3 after the loop
            float out = 0.0f;
#pragma unroll
            for (int l = 1; l <= 4; ++l) {
                f2 z[8], a[8], g[8];
                norm_gelu(h, z, a, g);
                if (l < 4) {
                    float2* st = reinterpret_cast<float2*>(stash) + (l - 1) * 16 * kSS;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { st[i * kSS] = z[i]; st[(8 + i) * kSS] = g[i]; }
                    store_operand(lane_base, kColAhi, kColAlo, a);
                    wait_st();
                    fence_before_sync();
                    named_barrier(1 + group, kGroupThreads);
                    if (gt == 0) issue_stage(1 + l, tmem, wdesc, idesc, mbar);
                    mbar_wait(mbar, parity); parity ^= 1;
                    fence_after_sync();
                    load_biased(lane_base + kColD, sW + kOffBias + 16 * l, h);
                } else {
                    // out = w4 . a4 + b4;  abar4 = w4;  hbar3 = adjoint through LayerNorm/GELU of layer 4, right here
                    const float2* w4 = reinterpret_cast<const float2*>(sW + kOffTail);
                    f2 acc = bc(0.0f), wv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { wv[i] = w4[i]; acc = fma2(wv[i], a[i], acc); }
                    out = acc.x + acc.y + sW[kOffTail + 16];
                    norm_gelu_adjoint(wv, z, g, abar);                                    // = hbar3 (adjoint of h3)
                }
            }
            // ------------------------------------------------------------ R3..R1: abar_l = W_l^T hbar_l, then through layer l's norm/GELU
#pragma unroll
            for (int l = 3; l >= 1; --l) {
                store_operand(lane_base, kColAhi, kColAlo, abar);
                wait_st();
                fence_before_sync();
                named_barrier(1 + group, kGroupThreads);
                if (gt == 0) issue_stage(8 - l, tmem, wdesc, idesc, mbar);
                mbar_wait(mbar, parity); parity ^= 1;
                fence_after_sync();
                f2 ab[8], z[8], g[8];
                const float2* st = reinterpret_cast<const float2*>(stash) + (l - 1) * 16 * kSS;
#pragma unroll
                for (int i = 0; i < 8; ++i) { z[i] = st[i * kSS]; g[i] = st[(8 + i) * kSS]; }
                load_pairs(lane_base + kColD, ab);
                norm_gelu_adjoint(ab, z, g, abar);                                        // abar <- hbar_{l-1}
            }
            // ------------------------------------------------------------ d out / d a_c = hbar0 . g_c
            float ga[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                f2 g[8];
                load_pairs(lane_base + kColG + 16 * c, g);
                f2 acc = bc(0.0f);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fma2(abar[i], g[i], acc);
                ga[c] = acc.x + acc.y;
            }
            const float res = sigmoidf_(out - 1.0f);
            const float sp = res * (1.0f - res) * pi_scale;
            const float gp0 = b.gp[0] + sp * b.s[0] * ga[0];
            const float gp1 = b.gp[1] + sp * ga[1];
            const float gp2 = b.gp[2] + sp * ga[2];
            if (in_range)
                field[(size_t)inst * total + idx] = make_float4(
                    b.value + res,
                    I.R[0] * gp0 + I.R[1] * gp1 + I.R[2] * gp2,
                    I.R[3] * gp0 + I.R[4] * gp1 + I.R[5] * gp2,
                    I.R[6] * gp0 + I.R[7] * gp1 + I.R[8] * gp2);
            // the next tile's tcgen05.st must not overtake this tile's tcgen05.ld of the same columns
            fence_before_sync();
        }
        seg = seg_end;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<512>(s_tmem_base);
}

}  // namespace fu

static int g_umma_sms = 0;
constexpr int kUmmaGroups = 4;      // measured (cfg2 fine pass): 4 groups x 128 registers 0.227 ms, 3 groups x 160 registers 0.260 ms

template <int kGroups>
static void launch_umma(const SceneDev& s, const RaysDev& r, float* field, size_t total, cudaStream_t st) {
    const int tiles_per_inst = (int)((total + fu::kTile - 1) / fu::kTile);
    const long long all_tiles = (long long)s.N * tiles_per_inst;
    const long long want = (all_tiles + kGroups - 1) / kGroups;
    const int grid = (int)(want < g_umma_sms ? want : g_umma_sms);
    if (r.fwd_samples != nullptr)  // instance culling on: the variant that walks the lists of live samples
        fu::field_forward_umma_kernel<kGroups, true><<<grid, kGroups * fu::kGroupThreads, fu::smem_bytes(kGroups), st>>>(
            s, r, (float4*)field, tiles_per_inst);
    else
        fu::field_forward_umma_kernel<kGroups, false><<<grid, kGroups * fu::kGroupThreads, fu::smem_bytes(kGroups), st>>>(
            s, r, (float4*)field, tiles_per_inst);
}

int launch_field_forward_umma(const SceneDev& s, const RaysDev& r, float* field, size_t total, cudaStream_t st) {
    if (!g_umma_sms) {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
            return fail("vsrd_b200: no CUDA device%s");
        if (cudaFuncSetAttribute(fu::field_forward_umma_kernel<kUmmaGroups, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)fu::smem_bytes(kUmmaGroups)) != cudaSuccess ||
            cudaFuncSetAttribute(fu::field_forward_umma_kernel<kUmmaGroups, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)fu::smem_bytes(kUmmaGroups)) != cudaSuccess)
            return fail("vsrd_b200: cannot reserve shared memory for field_forward_umma_kernel (built for sm_100a)%s");
        g_umma_sms = prop.multiProcessorCount;
    }
    launch_umma<kUmmaGroups>(s, r, field, total, st);
    return 0;
}

}  // namespace vsrd
