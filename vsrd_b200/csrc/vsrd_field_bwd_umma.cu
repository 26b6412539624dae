// sm_100a field BACKWARD kernel on tcgen05 + TMEM (thread-per-sample layout of vsrd_field_umma.cu): given the adjoints
// (dd, dG) of every instance's field value and spatial gradient, accumulate the gradients of
//     phi = dd * d(p; theta) + (R^T dG) . grad_p d(p; theta)
// w.r.t. the instance pose (t, half extents, R) and the 1617 residual-MLP weights ("one tangent +
// one reverse sweep"; scalar statement: vsrd_math.cuh::field_backward / mlp_reverse).
//
// One thread == one sample, 128-sample tiles, three groups of four warps per CTA, one persistent CTA per SM.
// Per tile the group makes eleven TMEM round trips:
//   L0 (2)     e_c -> h0 = W0 e + b0 and g_c = dh0/da_c (derivative weights);  hd0 = sum_c adot_c g_c   (the tangent)
//   L1..L3     (a, ad) = dual gelu(LayerNorm) -> h = W a + b, hd = W ad
//   layer 4    on the SIMT pipes, with its adjoint and its weight gradient (registers)
//   R3..R1     (hbar, hdbar) -> (gbar, gdbar) = W_l^T (.)   +   dW_l += act^T adj   (see below)
//   RL0 (3)    per coordinate: e_c recomputed -> g_c, k_c = d g_c / d a_c -> abar_c, adbar_c;  dW0_c += act^T adj
// WEIGHT GRADIENT: dW_l[o][i] = sum_s hbar[s,o] a[s,i] + hdbar[s,o] ad[s,i] contracts over SAMPLES, which sit on TMEM
// lanes, so both operands go through shared memory with the samples along K: every thread writes its sample's column of
// the bf16, K-major, SWIZZLE_128B operands (act rows = features + a ones row for the bias, adj rows = outputs) and one
// SS-mode UMMA chain per (layer, value | tangent) accumulates D[feature][output] in TMEM columns that live for the whole
// (CTA, instance) segment and are read out once at its end.  bf16 rounding (RN, unbiased) averages out over the
// >= 10^5 samples of a segment (tools/umma_probe.cu T5; tests hold 1e-3 rel on the weight gradients).
// Stash: z, zd of layers 1 and 2 in shared memory (256 B per thread); layer 3's stay in registers across layer 4, layer
// 4 is reversed inline, GELU terms are recomputed in the reverse sweep, e_c is recomputed for RL0.
#include <cuda_bf16.h>

#include "vsrd_umma_field.cuh"

namespace vsrd {
namespace bu {

using namespace umma;
using namespace fu;

constexpr int kGroups = 3;
constexpr int kThreadsB = kGroups * kGroupThreads;

// ---- TMEM columns: 3 groups x 128 + 96 weight-gradient accumulators ------------------------------------------------
constexpr int kColBuf0 = 0, kColBuf1 = 32;     // L0 / RL0 operand buffers (hi 16 + lo 16)
constexpr int kColA = 0, kColAd = 32;          // hidden / reverse operands: value hi|lo [0,32), tangent hi|lo [32,64)
constexpr int kColD = 64, kColDd = 80;         // accumulators: h | gbar | g_c [64,80), hd | gdbar | k_c [80,96)
constexpr int kColG1 = 96;                     // second g_c of the forward L0 [96,112)
constexpr int kColPark = 96;                   // after L0: layer 3's (z, zd) parked across layer 4 [96,128)
constexpr int kColsPerGroup = 128;
constexpr int kColDw = kGroups * kColsPerGroup;          // 384: dW0_c at +16 c (c = 0..2), dW_l at +48 + 16 (l - 1) (l = 1..3)
static_assert(kColDw + 96 <= 512, "TMEM columns");

// ---- shared memory -----------------------------------------------------------------------------------------------------
// staging of the weight-gradient operands, per group (bf16, K-major SWIZZLE_128B, 64 samples per 128-byte row):
constexpr int kAtom = 1024;                    // 8 rows x 128 bytes
constexpr int kStageAv = 0;                    // act, value pass:   [2 k-blocks][3 atoms]  rows 0..15 features, row 16 ones
constexpr int kStageAt = kStageAv + 2 * 3 * kAtom;   // act, tangent pass: [2][3]  rows 0..15, atom 2 stays ZERO: the bias row (D lane 16)
                                                     //                             must receive nothing from the tangent pass
constexpr int kStageBv = kStageAt + 2 * 3 * kAtom;   // adj, value pass:   [2][2]
constexpr int kStageBt = kStageBv + 2 * 2 * kAtom;   // adj, tangent pass: [2][2]
constexpr int kStageBytes = kStageBt + 2 * 2 * kAtom;                 // 20 KB per group
// B operands of the TS-mode contractions (fp32, [16 x 16] K-major blocks, hi then lo):
constexpr int kOffW0 = 0;                                  // [c] W0_c
constexpr int kOffW0d = kOffW0 + 3 * 2 * kBlk;             // [c] W0'_c   (d/da)
constexpr int kOffW0dd = kOffW0d + 3 * 2 * kBlk;           // [c] W0''_c  (d2/da2)
constexpr int kOffWl = kOffW0dd + 3 * 2 * kBlk;            // [l-1] W_l
constexpr int kOffWt = kOffWl + 3 * 2 * kBlk;              // [l-1] W_l^T
constexpr int kOffBias = kOffWt + 3 * 2 * kBlk;            // biases of layers 0..3
constexpr int kOffTail = kOffBias + 4 * 16;                // w4[16], b4
constexpr int kWeightFloats = kOffTail + 32;
constexpr int kStashPairs = 2 * 16;                        // per thread: layers 1, 2 x (z 8 pairs, zd 8 pairs)
constexpr int kRedFloats = 32;                             // layer-4 weight gradient (17) + pose (15)
constexpr int kInstFloats = 16;                            // the segment's instance: t, half extents, R
constexpr int kParkFloats = 15;                            // per thread, across the MLP: dG, x - t, and the box part of dimbar, pbar, vbar
constexpr size_t kSmemBytes = 1024 + (size_t)kGroups * kStageBytes + (size_t)(kWeightFloats + kRedFloats + kInstFloats) * 4
                            + (size_t)kThreadsB * kStashPairs * 8 + (size_t)kThreadsB * kParkFloats * 4;

__device__ void stage_weights_bwd(const float* __restrict__ W, float* sW) {
    for (int i = threadIdx.x; i < kHid * kEnc; i += blockDim.x) {                         // layer 0 and its a-derivatives
        const int o = i / kEnc, j = i % kEnc, c = j / 16, jj = j % 16, k = jj >> 1;
        const float w = __ldg(W + kW0 + o * (kEnc + 1) + j);
        const float f = (float)(1 << k);
        put_split(sW + kOffW0 + c * 2 * kBlk, o, jj, w);
        put_split(sW + kOffW0dd + c * 2 * kBlk, o, jj, -f * f * w);                       // d2/da2 cos(f a) = -f^2 cos(f a)
        if (jj & 1) put_split(sW + kOffW0d + c * 2 * kBlk, o, jj - 1, f * w);             // sin weight multiplies cos
        else put_split(sW + kOffW0d + c * 2 * kBlk, o, jj + 1, -f * w);                   // cos weight multiplies -sin
    }
    for (int i = threadIdx.x; i < 4 * kHid; i += blockDim.x) {
        const int l = i / kHid, o = i % kHid;
        sW[kOffBias + i] = l == 0 ? __ldg(W + kW0 + o * (kEnc + 1) + kEnc) : __ldg(W + kW1 + (l - 1) * kWStride + o * (kHid + 1) + kHid);
    }
    for (int i = threadIdx.x; i < 3 * kHid * kHid; i += blockDim.x) {
        const int l = i / (kHid * kHid), o = (i / kHid) % kHid, in = i % kHid;
        const float w = __ldg(W + kW1 + l * kWStride + o * (kHid + 1) + in);
        put_split(sW + kOffWl + l * 2 * kBlk, o, in, w);
        put_split(sW + kOffWt + l * 2 * kBlk, in, o, w);
    }
    for (int i = threadIdx.x; i < kHid + 1; i += blockDim.x) sW[kOffTail + i] = __ldg(W + kW4 + i);
}

// ---- dual (value + one tangent) LayerNorm -> GELU, channel pairs ------------------------------------------------------
// z = LN(h), zd = its tangent along hd, a = gelu(z), ad = gelu'(z) zd;  rs = 1 / sigma, mz = mean(z * centred hd)
template <bool kTerms>
__device__ __forceinline__ void norm_gelu_dual(const f2 (&h)[8], const f2 (&hd)[8], f2 (&z)[8], f2 (&zd)[8], f2 (&a)[8], f2 (&ad)[8],
                                               float& rs, float& mz, f2 (&g1)[8], f2 (&g2)[8]) {
    f2 s0 = h[0], s1 = hd[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s0 = add2(s0, h[i]); s1 = add2(s1, hd[i]); }
    const f2 mean = bc((s0.x + s0.y) * (-1.0f / 16.0f)), mt = bc((s1.x + s1.y) * (-1.0f / 16.0f));
    f2 var = bc(0.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { z[i] = add2(h[i], mean); var = fma2(z[i], z[i], var); }
    rs = rsqrtf((var.x + var.y) * (1.0f / 16.0f) + kLnEps);
    const f2 r2 = bc(rs);
    f2 acc = bc(0.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { z[i] = mul2(z[i], r2); zd[i] = add2(hd[i], mt); acc = fma2(z[i], zd[i], acc); }
    mz = (acc.x + acc.y) * (1.0f / 16.0f);
    const f2 nmz = bc(-mz);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        zd[i] = mul2(fma2(z[i], nmz, zd[i]), r2);
        f2 Phi, phi;
        gelu_terms2(z[i], Phi, phi);
        a[i] = mul2(z[i], Phi);
        const f2 d1 = fma2(z[i], phi, Phi);
        ad[i] = mul2(d1, zd[i]);
        if (kTerms) { g1[i] = d1; g2[i] = mul2(phi, fma2(z[i], mul2(z[i], bc(-1.0f)), bc(2.0f))); }
    }
}

// adjoint of the dual LayerNorm -> GELU (vsrd_math.cuh::ln_gelu_reverse): (gbar, gdbar) w.r.t. (a, ad) -> (hbar, hdbar) w.r.t. (h, hd)
__device__ __forceinline__ void ln_gelu_reverse2(const f2 (&z)[8], const f2 (&zd)[8], float r, float m, const f2 (&gbar)[8], const f2 (&gdbar)[8],
                                                 const f2 (&g1)[8], const f2 (&g2)[8], f2 (&hbar)[8], f2 (&hdbar)[8]) {
    f2 s_zb = bc(0.0f), s_zzb = bc(0.0f), s_zdb = bc(0.0f), s_zzdb = bc(0.0f), s_zdzdb = bc(0.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const f2 zdbar = mul2(gdbar[i], g1[i]);
        const f2 zbar = fma2(gbar[i], g1[i], mul2(mul2(gdbar[i], g2[i]), zd[i]));
        hbar[i] = zbar; hdbar[i] = zdbar;
        s_zb = add2(s_zb, zbar);
        s_zzb = fma2(z[i], zbar, s_zzb);
        s_zdb = add2(s_zdb, zdbar);
        s_zzdb = fma2(z[i], zdbar, s_zzdb);
        s_zdzdb = fma2(zd[i], zdbar, s_zdzdb);
    }
    const float inv = 1.0f / 16.0f;
    const float a_zb = (s_zb.x + s_zb.y) * inv, a_zzb = (s_zzb.x + s_zzb.y) * inv, a_zdb = (s_zdb.x + s_zdb.y) * inv;
    const float a_zzdb = (s_zzdb.x + s_zzdb.y) * inv, a_zdzdb = (s_zdzdb.x + s_zdzdb.y) * inv;
    const f2 rr = bc(r), rm = bc(-r * m);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const f2 hd_ = mul2(rr, fma2(z[i], bc(-a_zzdb), add2(hdbar[i], bc(-a_zdb))));
        f2 hv = fma2(z[i], bc(-(a_zzb + a_zdzdb)), add2(hbar[i], bc(-a_zb)));
        hv = fma2(zd[i], bc(-a_zzdb), hv);
        hbar[i] = fma2(rm, hd_, mul2(rr, hv));
        hdbar[i] = hd_;
    }
}

// ---- weight-gradient operand staging ------------------------------------------------------------------------------------
// byte offset of (row r of an atom, this thread's sample) inside a k-block region, for r = 0..7
struct StageOffsets { uint32_t row[8]; uint32_t kb; };
__device__ __forceinline__ StageOffsets stage_offsets(int gt) {
    StageOffsets so;
    const int sp = gt & 63, chunk = sp >> 3, e = sp & 7;
#pragma unroll
    for (int r = 0; r < 8; ++r) so.row[r] = (uint32_t)(r * 128 + ((chunk ^ r) << 4) + e * 2);
    so.kb = (uint32_t)(gt >> 6);
    return so;
}
__device__ __forceinline__ void stage_put(unsigned char* region, int atoms, const StageOffsets& so, int row, float v) {
    *reinterpret_cast<__nv_bfloat16*>(region + so.kb * atoms * kAtom + (row >> 3) * kAtom + so.row[row & 7]) = __float2bfloat16_rn(v);
}
__device__ __forceinline__ void stage_pairs(unsigned char* region, int atoms, const StageOffsets& so, const f2 (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { stage_put(region, atoms, so, 2 * i, v[i].x); stage_put(region, atoms, so, 2 * i + 1, v[i].y); }
}

// dW accumulate: D[tmem_d][feature][output] += sum over the tile's 128 samples, value pass (A rows incl. the ones row) then
// tangent pass; 8 k-steps of 16 samples each
__device__ __forceinline__ void mma_wgrad(uint32_t tmem_d, uint32_t stage_addr, uint32_t idesc_bf16) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const uint32_t k_off = (ks & 3) * 32;
        mma_bf16_ss(tmem_d, make_smem_desc_sw128(stage_addr + kStageAv + (ks >> 2) * 3 * kAtom + k_off, kAtom),
                    make_smem_desc_sw128(stage_addr + kStageBv + (ks >> 2) * 2 * kAtom + k_off, kAtom), idesc_bf16, true);
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const uint32_t k_off = (ks & 3) * 32;
        mma_bf16_ss(tmem_d, make_smem_desc_sw128(stage_addr + kStageAt + (ks >> 2) * 3 * kAtom + k_off, kAtom),
                    make_smem_desc_sw128(stage_addr + kStageBt + (ks >> 2) * 2 * kAtom + k_off, kAtom), idesc_bf16, true);
    }
}

// stages: 0 L0 (c = 0, 1)   1 L0 (c = 2)   2..4 hidden l = 1..3   5..7 reverse l = 3..1   8..10 reverse L0 c = 0..2
// The weight-gradient chain of a stage is issued AFTER the commit the group waits on, so it runs on the tensor pipe under
// the group's next SIMT phase; its own commit (mbar_w) guards the staging buffers and the final readout.
__device__ __forceinline__ void issue_stage(int stage, uint32_t tmem, uint32_t tmem_dw, uint64_t wdesc, uint32_t stage_addr,
                                            uint32_t idesc, uint32_t idesc_bf16, uint64_t* mbar, uint64_t* mbar_w) {
    fence_after_sync();
    if (stage == 0) {
        mma3_16x16(tmem + kColD, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0, idesc, false);
        mma3_16x16(tmem + kColDd, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0d, idesc, false);
        mma3_16x16(tmem + kColD, tmem + kColBuf1, tmem + kColBuf1 + 16, wdesc, kOffW0 + 2 * kBlk, idesc, true);
        mma3_16x16(tmem + kColG1, tmem + kColBuf1, tmem + kColBuf1 + 16, wdesc, kOffW0d + 2 * kBlk, idesc, false);
    } else if (stage == 1) {
        mma3_16x16(tmem + kColD, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0 + 4 * kBlk, idesc, true);
        mma3_16x16(tmem + kColDd, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0d + 4 * kBlk, idesc, false);
    } else if (stage <= 4) {
        const int l = stage - 1;
        mma3_16x16(tmem + kColD, tmem + kColA, tmem + kColA + 16, wdesc, kOffWl + (l - 1) * 2 * kBlk, idesc, false);
        mma3_16x16(tmem + kColDd, tmem + kColAd, tmem + kColAd + 16, wdesc, kOffWl + (l - 1) * 2 * kBlk, idesc, false);
    } else if (stage <= 7) {
        const int l = 8 - stage;
        mma3_16x16(tmem + kColD, tmem + kColA, tmem + kColA + 16, wdesc, kOffWt + (l - 1) * 2 * kBlk, idesc, false);
        mma3_16x16(tmem + kColDd, tmem + kColAd, tmem + kColAd + 16, wdesc, kOffWt + (l - 1) * 2 * kBlk, idesc, false);
        mma_commit(mbar);
        mma_wgrad(tmem_dw + 48 + 16 * (l - 1), stage_addr, idesc_bf16);
        mma_commit(mbar_w);
        return;
    } else {
        const int c = stage - 8;
        mma3_16x16(tmem + kColD, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0d + c * 2 * kBlk, idesc, false);
        mma3_16x16(tmem + kColDd, tmem + kColBuf0, tmem + kColBuf0 + 16, wdesc, kOffW0dd + c * 2 * kBlk, idesc, false);
        mma_commit(mbar);
        mma_wgrad(tmem_dw + 16 * c, stage_addr, idesc_bf16);
        mma_commit(mbar_w);
        return;
    }
    mma_commit(mbar);
}

__device__ __forceinline__ float dot16(const f2 (&a)[8], const f2 (&b)[8]) {
    f2 acc = mul2(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) acc = fma2(a[i], b[i], acc);
    return acc.x + acc.y;
}

// Sum 16 per-lane values over the warp with a transposing butterfly (16 shuffles): afterwards lane L holds the warp total
// of value (L >> 1); the even lanes add theirs to red[0..15].  Keeps no accumulator alive across tiles.
__device__ __forceinline__ void warp_sum16_to_smem(float (&v)[16], int lane, float* red) {
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, bit);
        }
    }
    v[0] += __shfl_xor_sync(kFull, v[0], 1);
    if ((lane & 1) == 0) atomicAdd(red + (lane >> 1), v[0]);
}

__global__ void __launch_bounds__(kThreadsB, 1) field_backward_umma_kernel(
        SceneDev scene, RaysDev rays, const float4* __restrict__ adjoint, float* __restrict__ partials, int tiles_per_inst) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t s_mbar[kGroups], s_mbar_w[kGroups];
    __shared__ uint32_t s_tmem_base;
    unsigned char* sStage = smem_raw;                                                    // 1024-byte aligned atoms
    float* sW = reinterpret_cast<float*>(smem_raw + kGroups * kStageBytes);
    float* sRed = sW + kWeightFloats;
    float* sInst = sRed + kRedFloats;
    float2* sStash = reinterpret_cast<float2*>(sInst + kInstFloats);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = warp >> 2, wq = warp & 3, gt = tid & (kGroupThreads - 1);
    float2* stash = sStash + (size_t)group * kGroupThreads * kStashPairs + gt;            // [pair][thread]
    constexpr int kSS = kGroupThreads;
    unsigned char* stage = sStage + group * kStageBytes;
    const StageOffsets so = stage_offsets(gt);
    float* park = reinterpret_cast<float*>(sStash + (size_t)kThreadsB * kStashPairs) + tid;   // [value][thread]

    if (warp == 0) tmem_alloc<512>(&s_tmem_base);
    if (tid == 0) {
#pragma unroll
        for (int g = 0; g < kGroups; ++g) { mbar_init(&s_mbar[g], 1); mbar_init(&s_mbar_w[g], 1); }
        mbar_fence_init();
    }
    for (int i = tid; i < kGroups * 2 * kAtom / 4; i += kThreadsB) {                   // the zero atoms of the tangent-pass act operand
        const int g = i / (2 * kAtom / 4), rest = i % (2 * kAtom / 4), kb = rest / (kAtom / 4), w = rest % (kAtom / 4);
        reinterpret_cast<uint32_t*>(sStage + g * kStageBytes + kStageAt + kb * 3 * kAtom + 2 * kAtom)[w] = 0u;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem_base + group * kColsPerGroup;
    const uint32_t tmem_dw = s_tmem_base + kColDw;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * wq) << 16);
    const uint32_t idesc = make_idesc_tf32(128, 16), idesc_bf16 = make_idesc_bf16(128, 16);
    const uint64_t wdesc = make_smem_desc(smem_u32(sW), kLbo, kSbo);
    const uint32_t stage_addr = smem_u32(stage);
    uint64_t* mbar = &s_mbar[group];
    uint64_t* mbar_w = &s_mbar_w[group];
    uint32_t parity = 0, parity_w = 0;
    bool wgrad_pending = false;                            // group-uniform: a weight-gradient chain may still read the staging buffers

    const int total = rays.R * rays.M;
    const long long all_tiles = (long long)scene.N * tiles_per_inst;
    const long long begin = all_tiles * blockIdx.x / gridDim.x;
    const long long end = all_tiles * (blockIdx.x + 1) / gridDim.x;
    const float pi_scale = kPiF / scene.scale;
    unsigned tiles_visited = 0, tiles_culled = 0;

    for (long long seg = begin; seg < end;) {
        const int inst = (int)(seg / tiles_per_inst);
        const long long seg_end = min(end, (long long)(inst + 1) * tiles_per_inst);
        fence_before_sync();
        __syncthreads();                                   // every group is done with the previous instance
        stage_weights_bwd(scene.W + (size_t)inst * kNumW, sW);
        if (tid < kRedFloats) sRed[tid] = 0.0f;
        if (tid >= 32 && tid < 47) {
            const int k = tid - 32;
            sInst[k] = k < 3 ? __ldg(scene.loc + 3 * inst + k) : k < 6 ? __ldg(scene.dim + 3 * inst + k - 3) : __ldg(scene.rot + 9 * inst + k - 6);
        }
        if (warp == 0) {                                   // zero the weight-gradient accumulators (lanes 0..31 are all that is read)
            float zero[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zero[i] = 0.0f;
#pragma unroll
            for (int b = 0; b < 6; ++b) tmem_st16(s_tmem_base + kColDw + 16 * b, zero);
            wait_st();
        }
        fence_proxy_async_smem();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();

#pragma unroll 1
        for (long long tile = seg + group; tile < seg_end; tile += kGroups) {
            const int base = (int)(tile - (long long)inst * tiles_per_inst) * kTile;
            const bool in_range = base + gt < total;
            const int idx = min(base + gt, total - 1);
            float4 adj = __ldg(adjoint + (size_t)inst * total + idx);
            if (!in_range) adj = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            {   // a tile without a single non-zero adjoint (culled instance, samples behind the surface) contributes nothing
                const bool live = adj.x != 0.0f || adj.y != 0.0f || adj.z != 0.0f || adj.w != 0.0f;
                const bool warp_live = __any_sync(kFull, live);
                int tile_live;
                asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.b32 %0, 1, 0, q;\n\t}"
                             : "=r"(tile_live) : "r"((int)warp_live), "r"(1 + group), "r"(kGroupThreads) : "memory");
                if (lane == 0) { ++tiles_visited; tiles_culled += warp_live ? 0u : 1u; }
                if (!tile_live) continue;
            }
            const int r = idx / rays.M;
            const int j = idx - r * rays.M;
            const float dd = adj.x;
            float m[3], coef[3], adot[3];
            {
            float x[3];
            sample_position(rays, r, j, x);
            Instance I;                                    // read from shared memory where it is used, not held across the tile
#pragma unroll
            for (int k = 0; k < 3; ++k) { I.t[k] = sInst[k]; I.dim[k] = sInst[3 + k]; }
#pragma unroll
            for (int k = 0; k < 9; ++k) I.R[k] = sInst[6 + k];
            BoxEval b;
            box_eval(x, I, b);
            const float dG[3] = {adj.y, adj.z, adj.w};
            float v[3], pbar[3], vbar[3], dimbar[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k] = I.R[k] * dG[0] + I.R[3 + k] * dG[1] + I.R[6 + k] * dG[2];
            {   // box part: phi_box = dd * box + v . grad_p box
                float vs = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) vs += v[k] * b.s[k] * b.a[k];
                const float inv_n = 1.0f / b.nrm, inv_n3 = inv_n * inv_n * inv_n;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float act = b.q[k] > 0.0f ? 1.0f : 0.0f;
                    const float hess = act * v[k] * b.s[k] * inv_n - b.a[k] * vs * inv_n3;
                    pbar[k] = dd * b.gp[k] + b.s[k] * hess;
                    dimbar[k] = -(dd * (b.a[k] * inv_n + b.ind[k]) + hess);
                    vbar[k] = b.gp[k];
                }
            }
            m[0] = fabsf(b.p[0]); m[1] = b.p[1]; m[2] = b.p[2];
            coef[0] = b.s[0] * pi_scale; coef[1] = pi_scale; coef[2] = pi_scale;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                adot[k] = coef[k] * v[k];
                park[(0 + k) * kThreadsB] = dG[k];
                park[(3 + k) * kThreadsB] = b.y[k];
                park[(6 + k) * kThreadsB] = dimbar[k];
                park[(9 + k) * kThreadsB] = pbar[k];
                park[(12 + k) * kThreadsB] = vbar[k];
            }
            }
            // ------------------------------------------------------------ L0 (dual): h0, hd0 = sum_c adot_c g_c
            f2 h[8], hd[8];
            {
                f2 e[8], g[8];
                encode16(kPiF * (m[0] / scene.scale), e);
                store_operand(lane_base, kColBuf0, kColBuf0 + 16, e);
                encode16(kPiF * (m[1] / scene.scale), e);
                store_operand(lane_base, kColBuf1, kColBuf1 + 16, e);
                wait_st();
                fence_before_sync();
                named_barrier(1 + group, kGroupThreads);
                if (gt == 0) issue_stage(0, tmem, tmem_dw, wdesc, stage_addr, idesc, idesc_bf16, mbar, mbar_w);
                encode16(kPiF * (m[2] / scene.scale), e);
                mbar_wait(mbar, parity); parity ^= 1;
                fence_after_sync();
                load_pairs(lane_base + kColDd, g);
#pragma unroll
                for (int i = 0; i < 8; ++i) hd[i] = mul2(g[i], bc(adot[0]));
                load_pairs(lane_base + kColG1, g);
#pragma unroll
                for (int i = 0; i < 8; ++i) hd[i] = fma2(g[i], bc(adot[1]), hd[i]);
                // order the loads above before the next MMA overwrites kColDd; the operand buffer is free again
                store_operand(lane_base, kColBuf0, kColBuf0 + 16, e);
                wait_st();
                fence_before_sync();
                named_barrier(1 + group, kGroupThreads);
                if (gt == 0) issue_stage(1, tmem, tmem_dw, wdesc, stage_addr, idesc, idesc_bf16, mbar, mbar_w);
                mbar_wait(mbar, parity); parity ^= 1;
                fence_after_sync();
                load_biased(lane_base + kColD, sW + kOffBias, h);
                load_pairs(lane_base + kColDd, g);
#pragma unroll
                for (int i = 0; i < 8; ++i) hd[i] = fma2(g[i], bc(adot[2]), hd[i]);
            }
            // ------------------------------------------------------------ L1..L3 (dual)
            float rs_l[3], mz_l[3];
#pragma unroll
            for (int l = 1; l <= 3; ++l) {
                f2 z[8], zd[8], a[8], ad[8];
                norm_gelu_dual<false>(h, hd, z, zd, a, ad, rs_l[l - 1], mz_l[l - 1], a, a);
                if (l < 3) {
                    float2* st = stash + (l - 1) * 16 * kSS;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { st[i * kSS] = z[i]; st[(8 + i) * kSS] = zd[i]; }
                } else {                                   // layer 3's pair is parked in this group's spare TMEM columns
                    float t[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { t[2 * i] = z[i].x; t[2 * i + 1] = z[i].y; }
                    tmem_st16(lane_base + kColPark, t);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { t[2 * i] = zd[i].x; t[2 * i + 1] = zd[i].y; }
                    tmem_st16(lane_base + kColPark + 16, t);
                }
                store_operand(lane_base, kColA, kColA + 16, a);
                store_operand(lane_base, kColAd, kColAd + 16, ad);
                wait_st();
                fence_before_sync();
                named_barrier(1 + group, kGroupThreads);
                if (gt == 0) issue_stage(1 + l, tmem, tmem_dw, wdesc, stage_addr, idesc, idesc_bf16, mbar, mbar_w);
                mbar_wait(mbar, parity); parity ^= 1;
                fence_after_sync();
                load_biased(lane_base + kColD, sW + kOffBias + 16 * l, h);
                load_pairs(lane_base + kColDd, hd);
            }
            // ------------------------------------------------------------ layer 4 and its adjoint, on the SIMT pipes
            f2 hbar[8], hdbar[8];
            float db4;
            {
                f2 z[8], zd[8], a[8], ad[8], g1[8], g2[8];
                float rs4, mz4;
                norm_gelu_dual<true>(h, hd, z, zd, a, ad, rs4, mz4, g1, g2);
                const float2* w4 = reinterpret_cast<const float2*>(sW + kOffTail);
                f2 wv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) wv[i] = w4[i];
                const float out = dot16(wv, a) + sW[kOffTail + 16];
                const float outd = dot16(wv, ad);
                const float res = sigmoidf_(out - 1.0f);
                const float sp = res * (1.0f - res);
                const float spp = sp * (1.0f - 2.0f * res);
                const float obar = dd * sp + spp * outd;
                const float odbar = sp;
                db4 = obar;
                f2 gbar[8], gdbar[8];
                {
                    float dw4[16];                         // layer-4 weight gradient of this sample -> warp sum -> sRed[0..15]
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const f2 t = fma2(a[i], bc(obar), mul2(ad[i], bc(odbar)));
                        dw4[2 * i] = t.x; dw4[2 * i + 1] = t.y;
                    }
                    warp_sum16_to_smem(dw4, lane, sRed);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    gbar[i] = mul2(wv[i], bc(obar));
                    gdbar[i] = mul2(wv[i], bc(odbar));
                }
                ln_gelu_reverse2(z, zd, rs4, mz4, gbar, gdbar, g1, g2, hbar, hdbar);          // adjoints of (h3, hd3)
            }
            // ------------------------------------------------------------ R3..R1
#pragma unroll
            for (int l = 3; l >= 1; --l) {
                // dW_l operands: act = (a | 1), (ad);  adj = hbar, hdbar   (h_l = W_l a_l + b_l).  The adjoints go out first
                // (staging + MMA operands) so that they are dead while the layer's GELU terms are rebuilt.
                if (wgrad_pending) { mbar_wait(mbar_w, parity_w); parity_w ^= 1; }
                wgrad_pending = true;
                stage_pairs(stage + kStageBv, 2, so, hbar);
                stage_pairs(stage + kStageBt, 2, so, hdbar);
                store_operand(lane_base, kColA, kColA + 16, hbar);
                store_operand(lane_base, kColAd, kColAd + 16, hdbar);
                f2 z[8], zd[8], g1[8], g2[8];
                if (l == 3) {
                    load_pairs(lane_base + kColPark, z);
                    load_pairs(lane_base + kColPark + 16, zd);
                } else {
                    const float2* st = stash + (l - 1) * 16 * kSS;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { z[i] = st[i * kSS]; zd[i] = st[(8 + i) * kSS]; }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {              // a = z Phi, g1 = gelu'(z), g2 = gelu''(z) = phi (2 - z^2), ad = g1 zd
                    f2 Phi, phi;
                    gelu_terms2(z[i], Phi, phi);
                    const f2 a = mul2(z[i], Phi);
                    g1[i] = fma2(z[i], phi, Phi);
                    const f2 ad = mul2(g1[i], zd[i]);
                    g2[i] = mul2(phi, fma2(z[i], mul2(z[i], bc(-1.0f)), bc(2.0f)));
                    stage_put(stage + kStageAv, 3, so, 2 * i, a.x); stage_put(stage + kStageAv, 3, so, 2 * i + 1, a.y);
                    stage_put(stage + kStageAt, 3, so, 2 * i, ad.x); stage_put(stage + kStageAt, 3, so, 2 * i + 1, ad.y);
                }
                stage_put(stage + kStageAv, 3, so, 16, 1.0f);
                fence_proxy_async_smem();
                wait_st();
                fence_before_sync();
                named_barrier(1 + group, kGroupThreads);
                if (gt == 0) issue_stage(8 - l, tmem, tmem_dw, wdesc, stage_addr, idesc, idesc_bf16, mbar, mbar_w);
                mbar_wait(mbar, parity); parity ^= 1;
                fence_after_sync();
                f2 gbar[8], gdbar[8];
                load_pairs(lane_base + kColD, gbar);
                load_pairs(lane_base + kColDd, gdbar);
                ln_gelu_reverse2(z, zd, rs_l[l - 1], mz_l[l - 1], gbar, gdbar, g1, g2, hbar, hdbar);   // adjoints of (h_{l-1}, hd_{l-1})
            }
            // ------------------------------------------------------------ RL0: per coordinate
            mbar_wait(mbar_w, parity_w); parity_w ^= 1;                                   // R1's chain is done with the staging buffers
            wgrad_pending = false;
            stage_pairs(stage + kStageBv, 2, so, hbar);                                   // value-pass adjoint: the same for c = 0..2
            float pbar[3], vbar[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                f2 e[8];
                encode16(kPiF * (m[c] / scene.scale), e);
                {
                    f2 et[8], q[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {                                         // d e / d a: (cos, sin) -> f (-sin, cos)
                        const float f = (float)(1 << k);
                        et[k] = make_float2(-f * e[k].y, f * e[k].x);
                        q[k] = mul2(hdbar[k], bc(adot[c]));
                    }
                    if (wgrad_pending) { mbar_wait(mbar_w, parity_w); parity_w ^= 1; }
                    wgrad_pending = true;
                    stage_pairs(stage + kStageAv, 3, so, e);
                    stage_put(stage + kStageAv, 3, so, 16, c == 0 ? 1.0f : 0.0f);         // the bias row counts once
                    stage_pairs(stage + kStageAt, 3, so, et);
                    stage_pairs(stage + kStageBt, 2, so, q);
                }
                fence_proxy_async_smem();
                store_operand(lane_base, kColBuf0, kColBuf0 + 16, e);
                wait_st();
                fence_before_sync();
                named_barrier(1 + group, kGroupThreads);
                if (gt == 0) issue_stage(8 + c, tmem, tmem_dw, wdesc, stage_addr, idesc, idesc_bf16, mbar, mbar_w);
                mbar_wait(mbar, parity); parity ^= 1;
                fence_after_sync();
                f2 g[8], kk[8];
                load_pairs(lane_base + kColD, g);
                load_pairs(lane_base + kColDd, kk);
                const float gh = dot16(hbar, g), gd = dot16(hdbar, g), kd = dot16(hdbar, kk);
                const float abar = gh + adot[c] * kd;                                    // d phi / d a_c
                pbar[c] = abar * coef[c];
                vbar[c] = gd * coef[c];                                                 // d phi / d adot_c = hdbar . g_c
            }
            // ------------------------------------------------------------ pose: p = R^T (x - t), v = R^T dG
            {
                float pose[16];                            // sRed[16] = layer-4 bias, sRed[17..31] = pose
                pose[0] = db4;
#pragma unroll
                for (int k = 0; k < 3; ++k) { pbar[k] += park[(9 + k) * kThreadsB]; vbar[k] += park[(12 + k) * kThreadsB]; }
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) {
                    pose[1 + mm] = -(sInst[6 + 3 * mm] * pbar[0] + sInst[6 + 3 * mm + 1] * pbar[1] + sInst[6 + 3 * mm + 2] * pbar[2]);
                    pose[4 + mm] = park[(6 + mm) * kThreadsB];
                    const float y = park[(3 + mm) * kThreadsB], dg = park[mm * kThreadsB];
#pragma unroll
                    for (int k = 0; k < 3; ++k) pose[7 + 3 * mm + k] = y * pbar[k] + dg * vbar[k];
                }
                warp_sum16_to_smem(pose, lane, sRed + 16);
            }
            fence_before_sync();
        }
        if (wgrad_pending) { mbar_wait(mbar_w, parity_w); parity_w ^= 1; wgrad_pending = false; }
        // ---------------------------------------------------------------- segment epilogue: one partial row per (CTA, instance)
        fence_before_sync();
        __syncthreads();                                   // all groups passed their last commit wait: the accumulators are final
        fence_after_sync();
        float* out = partials + ((size_t)blockIdx.x + inst) * kGradStride;
        if (warp == 0) {
#pragma unroll
            for (int blk = 0; blk < 6; ++blk) {
                float d[16];
                tmem_ld16(s_tmem_base + kColDw + 16 * blk, d);
                wait_ld();
                if (lane <= 16) {
                    if (blk < 3) {                         // layer 0, coordinate blk: rows = PE features, row 16 = bias (c == 0 only)
                        if (lane < 16 || blk == 0) {
                            const int col = lane < 16 ? blk * 16 + lane : kEnc;
#pragma unroll
                            for (int o = 0; o < 16; ++o) out[kW0 + o * (kEnc + 1) + col] = d[o];
                        }
                    } else {                               // hidden layer l = blk - 2
                        const int l = blk - 2;
#pragma unroll
                        for (int o = 0; o < 16; ++o) out[kW1 + (l - 1) * kWStride + o * (kHid + 1) + lane] = d[o];
                    }
                }
            }
        } else if (warp == 1) {
            if (lane <= 16) out[kW4 + lane] = sRed[lane];
            else out[kNumW + lane - 17] = sRed[lane];
        }
        seg = seg_end;
    }
    if (lane == 0 && rays.cull_stats != nullptr && tiles_visited) {
        atomicAdd(rays.cull_stats, (unsigned long long)tiles_culled);
        atomicAdd(rays.cull_stats + 1, (unsigned long long)tiles_visited);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<512>(s_tmem_base);
}

// Sum the partial rows of every instance: the CTAs whose tile range intersects the instance's tiles are b0..b1
// (cta(T) = ((T + 1) * grid - 1) / all_tiles for tile T), their rows b + inst.
__global__ void reduce_segment_rows_umma_kernel(const float* __restrict__ partials, int grid, int tiles_per_inst,
                                                long long all_tiles, float* __restrict__ gloc, float* __restrict__ grot,
                                                float* __restrict__ gdim, float* __restrict__ gW) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int inst = blockIdx.y;
    if (f >= kNumW + kNumPose) return;
    const long long first = (long long)inst * tiles_per_inst, last = first + tiles_per_inst - 1;
    const int b0 = (int)(((first + 1) * grid - 1) / all_tiles);
    const int b1 = (int)(((last + 1) * grid - 1) / all_tiles);
    float s = 0.0f;
    for (int b = b0; b <= b1; ++b) s += partials[((size_t)b + inst) * kGradStride + f];
    if (f < kNumW) gW[(size_t)inst * kNumW + f] = s;
    else if (f < kNumW + 3) gloc[3 * inst + (f - kNumW)] = s;
    else if (f < kNumW + 6) gdim[3 * inst + (f - kNumW - 3)] = s;
    else grot[9 * inst + (f - kNumW - 6)] = s;
}

}  // namespace bu

static int g_bu_sms = 0;

int launch_field_backward_umma(const SceneDev& s, const RaysDev& r, const float* adjoint, float* partials,
                               float* gloc, float* grot, float* gdim, float* gW, cudaStream_t st) {
    if (!g_bu_sms) {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
            return fail("vsrd_b200: no CUDA device%s");
        if (cudaFuncSetAttribute(bu::field_backward_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)bu::kSmemBytes) != cudaSuccess)
            return fail("vsrd_b200: cannot reserve shared memory for field_backward_umma_kernel (built for sm_100a)%s");
        g_bu_sms = prop.multiProcessorCount;
    }
    const long long total = (long long)r.R * r.M;
    const int tiles_per_inst = (int)((total + fu::kTile - 1) / fu::kTile);
    const long long all_tiles = (long long)s.N * tiles_per_inst;
    const long long want = (all_tiles + bu::kGroups - 1) / bu::kGroups;
    const int grid = (int)(want < g_bu_sms ? want : g_bu_sms);
    bu::field_backward_umma_kernel<<<grid, bu::kThreadsB, bu::kSmemBytes, st>>>(s, r, (const float4*)adjoint, partials, tiles_per_inst);
    VSRD_CHECK_LAUNCH();
    const dim3 rgrid((kGradStride + 127) / 128, (unsigned)s.N);
    bu::reduce_segment_rows_umma_kernel<<<rgrid, 128, 0, st>>>(partials, grid, tiles_per_inst, all_tiles, gloc, grot, gdim, gW);
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // namespace vsrd
