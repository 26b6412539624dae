// Warp-level building blocks of the tensor-core field kernels (v4).
//
// One warp owns a tile of 32 (sample, instance) pairs of one instance: two 16-row MMA tiles.  All layer
// activations live in registers in the mma.sync accumulator ("C") layout and are chained into the next
// contraction as the A operand without leaving the register file:
//
//   lane = 4 g + t   (g = lane >> 2 in 0..7, t = lane & 3)
//   C fragment of an m16n8 tile:  c0 (row g, col 2t)  c1 (row g, col 2t+1)  c2 (row g+8, col 2t)  c3 (row g+8, col 2t+1)
//   A fragment of m16n8k8 (tf32): a0 (row g, k t)     a1 (row g+8, k t)     a2 (row g, k t+4)     a3 (row g+8, k t+4)
//   B fragment of m16n8k8 (tf32): b0 (k t, n g)       b1 (k t+4, n g)
//
// The contraction index k is only summed over, so it can be permuted freely as long as A and B agree.
// We let k-slot t of k-step j stand for channel 8j+2t and k-slot t+4 for channel 8j+2t+1: then the C
// fragment of n-tile j IS the A fragment of k-step j (a0=c0, a1=c2, a2=c1, a3=c3) and the weight
// fragments are staged once per CTA in that permuted order.
//
// Rows: a lane holds 4 rows of the 32-row tile, "slots" s = 0..3 -> row 8 s + g  (m-tile s >> 1, half s & 1).
// Activation registers are indexed x[mt][nt][q] (q = c0..c3); slot s holds x[s>>1][nt][2*(s&1) + {0,1}].
//
// Precision: every contraction runs as 3xTF32 (hi*hi + hi*lo + lo*hi with fp32 accumulation), which
// keeps ~21 mantissa bits: the silhouette parity bar (1e-4 abs) cannot be met by single-pass TF32
// (7.3e-5 from the hidden contractions alone).  profiles/r01_pipe_rates_b200.txt:
// the legacy tensor pipe sustains one m16n8k8 per 5-6 cycles per SM sub-partition and overlaps
// perfectly with >= 8 FP32 instructions per MMA, which is the ratio these kernels have.
#pragma once
#include "vsrd_common.cuh"

namespace vsrd {
namespace frag {

// ---- shared-memory image of one instance's weights -------------------------------------------
// float4 {b0_hi, b1_hi, b0_lo, b1_lo} per (fragment, lane); fragment order:
constexpr int kF0 = 0;              // layer 0 forward:   [ks 0..5][nt 0..1]     b0 = W0[8nt+g][8ks+2t], b1 = W0[8nt+g][8ks+2t+1]
constexpr int kF1 = kF0 + 12;       // hidden forward:    [l-1][ks 0..1][nt 0..1] b0 = Wl[8nt+g][8ks+2t]
constexpr int kR1 = kF1 + 12;       // hidden transposed: [l-1][ks 0..1][nt 0..1] b0 = Wl[8ks+2t][8nt+g], b1 = Wl[8ks+2t+1][8nt+g]
constexpr int kR0 = kR1 + 12;       // layer 0 transposed:[ks 0..1][nt 0..5]     b0 = W0[8ks+2t][8nt+g]
constexpr int kNumFrag = kR0 + 12;  // 48 fragments x 32 lanes x 16 B = 24 KB
constexpr int kFragFloat4 = kNumFrag * 32;
// followed by the biases in fp32: b0[16] b1[16] b2[16] b3[16] w4[16] b4
constexpr int kBias = 0;            // offsets into the float tail
constexpr int kTailW4 = 64;
constexpr int kTailB4 = 80;
constexpr int kTailFloats = 96;
constexpr size_t kWeightBytes = (size_t)kFragFloat4 * 16 + kTailFloats * 4;

__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;            // exact TF32 (truncated)
    lo = __float_as_uint(x - __uint_as_float(hi));    // remainder; the MMA reads its top 19 bits
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// d += A * B with A = ah + al, B = bh + bl (the lo*lo term is below fp32 rounding)
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const float4& b) {
    mma_tf32(d, al, __float_as_uint(b.x), __float_as_uint(b.y));
    mma_tf32(d, ah, __float_as_uint(b.z), __float_as_uint(b.w));
    mma_tf32(d, ah, __float_as_uint(b.x), __float_as_uint(b.y));
}

// A fragment (hi, lo) of k-step `nt` from an activation tile held in C layout.
__device__ __forceinline__ void a_from_c(const float (&c)[4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
    split(c[0], ah[0], al[0]);
    split(c[2], ah[1], al[1]);
    split(c[1], ah[2], al[2]);
    split(c[3], ah[3], al[3]);
}

// ---- sample-contracted products (weight gradients) ---------------------------------------------
// dW[o][i] = sum_samples hbar[s][o] g[s][i] contracts over the SAMPLE index, which the C layout keeps in
// the 8-valued lane coordinate g while both MMA operands need the contracted index in the 4-valued
// coordinate t.  movmatrix transposes an 8x8 tile of 16-bit pairs inside the register file, so the
// operands are carried as bf16 hi + bf16 lo (16 mantissa bits, products hi*hi + hi*lo + lo*hi): one
// m16n8k16 covers 16 samples.  Gradient tolerance is 1e-3; this keeps ~1e-5 per product.
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

// (x0, x1) -> packed bf16 pairs {lo half = x0, hi half = x1}: rounded value and rounded remainder.
__device__ __forceinline__ void pack_bf16_split(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}

// Pair (x0, x1) = C-layout elements (sample g, channels 2t, 2t+1) of an 8x8 block -> transposed operand
// register (samples 2t, 2t+1; channel g), hi and lo parts.
__device__ __forceinline__ void pack_transposed(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    uint32_t h, l;
    pack_bf16_split(x0, x1, h, l);
    hi = movmatrix_trans(h);
    lo = movmatrix_trans(l);
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A operand (16 outputs x 16 samples of one m-tile) of the weight-gradient product from the adjoint's
// two C tiles c[nt][q] (nt = output half): rows = outputs, k = samples.
__device__ __forceinline__ void wgrad_a_operand(const float (&c)[2][4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
    pack_transposed(c[0][0], c[0][1], ah[0], al[0]);   // outputs 0-7,  samples 0-7
    pack_transposed(c[1][0], c[1][1], ah[1], al[1]);   // outputs 8-15, samples 0-7
    pack_transposed(c[0][2], c[0][3], ah[2], al[2]);   // outputs 0-7,  samples 8-15
    pack_transposed(c[1][2], c[1][3], ah[3], al[3]);   // outputs 8-15, samples 8-15
}

// D (16 outputs x 8 inputs) += A (adjoint) x B, B = one C tile (16 samples x 8 input channels).
__device__ __forceinline__ void wgrad_tile(float (&D)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           float b0, float b1, float b2, float b3) {
    uint32_t bh0, bl0, bh1, bl1;
    pack_transposed(b0, b1, bh0, bl0);
    pack_transposed(b2, b3, bh1, bl1);
    mma_bf16(D, ah, bl0, bl1);
    mma_bf16(D, al, bh0, bh1);
    mma_bf16(D, ah, bh0, bh1);
}

// D += A x ones: every column holds the sum over the 16 samples (bias gradient).
__device__ __forceinline__ void wgrad_bias(float (&D)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4]) {
    constexpr uint32_t kOnes = 0x3f803f80u;
    mma_bf16(D, al, kOnes, kOnes);
    mma_bf16(D, ah, kOnes, kOnes);
}

// Reference weight layout (hyper_distance_field.py:57-73): layer l rows [out][fan_in + 1], bias last.
__device__ __forceinline__ void stage_weight_fragments(const float* __restrict__ W, float4* sF, float* sTail) {
    for (int f = threadIdx.x; f < kFragFloat4; f += blockDim.x) {
        const int lane = f & 31, frag = f >> 5;
        const int g = lane >> 2, t = lane & 3;
        int i0, i1;
        if (frag < kF1) {                       // layer 0 forward
            const int ks = frag >> 1, nt = frag & 1;
            i0 = (8 * nt + g) * (kEnc + 1) + 8 * ks + 2 * t;
            i1 = i0 + 1;
        } else if (frag < kR1) {                // hidden forward
            const int q = frag - kF1, l = q >> 2, ks = (q >> 1) & 1, nt = q & 1;
            i0 = kW1 + l * kWStride + (8 * nt + g) * (kHid + 1) + 8 * ks + 2 * t;
            i1 = i0 + 1;
        } else if (frag < kR0) {                // hidden transposed
            const int q = frag - kR1, l = q >> 2, ks = (q >> 1) & 1, nt = q & 1;
            i0 = kW1 + l * kWStride + (8 * ks + 2 * t) * (kHid + 1) + 8 * nt + g;
            i1 = i0 + (kHid + 1);
        } else {                                // layer 0 transposed
            const int q = frag - kR0, ks = q / 6, nt = q % 6;
            i0 = (8 * ks + 2 * t) * (kEnc + 1) + 8 * nt + g;
            i1 = i0 + (kEnc + 1);
        }
        uint32_t h0, l0, h1, l1;
        split(__ldg(W + i0), h0, l0);
        split(__ldg(W + i1), h1, l1);
        sF[f] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
    }
    for (int f = threadIdx.x; f < kTailFloats; f += blockDim.x) {
        float v = 0.0f;
        if (f < 16) v = __ldg(W + f * (kEnc + 1) + kEnc);
        else if (f < 64) v = __ldg(W + kW1 + ((f >> 4) - 1) * kWStride + (f & 15) * (kHid + 1) + kHid);
        else if (f <= kTailB4) v = __ldg(W + kW4 + (f - 64));
        sTail[f] = v;
    }
}

// ---- small helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(kFull, v, 1);
    v += __shfl_xor_sync(kFull, v, 2);
    return v;
}

// sin / cos with a three-term Cody-Waite reduction (|x| < ~1e4) and the single-precision minimax
// polynomials on [-pi/4, pi/4]; absolute error ~1e-7, no slow path, no local memory.
__host__ __device__ __forceinline__ void sincos_cw(float x, float& s, float& c) {
    const float kf = rintf(x * 0.63661977236758134308f);
    float r = fmaf(kf, -1.5707962513e+00f, x);
    r = fmaf(kf, -7.5497894159e-08f, r);
    r = fmaf(kf, -5.3903029534e-15f, r);
    const int q = (int)kf;
    const float r2 = r * r;
    float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, r2, -1.6666654611e-1f);
    sp = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, r2, 4.166664568298827e-2f);
    cp = fmaf(cp, r2, -0.5f);
    cp = fmaf(cp, r2, 1.0f);
    float ss = (q & 1) ? cp : sp;
    float cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

// Phi, phi of the exact-erf GELU (see vsrd_math.cuh::gelu_terms) with single-instruction reciprocal.
__device__ __forceinline__ void gelu_terms_fast(float z, float& Phi, float& phi) {
    const float az = fabsf(z);
    float E;                                                          // exp(-z^2/2)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E) : "f"(-0.72134752044448170368f * z * z));
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * kInvSqrt2, az, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float tail = 0.5f * poly * t * E;
    Phi = z >= 0.0f ? 1.0f - tail : tail;
    phi = kInvSqrt2Pi * E;
}

// Per-lane view of the positional encoding of its 4 rows: lane t owns the (cos, sin) pairs of
// frequencies k = t and k = t + 4 of each coordinate (channels 16 c + 2 k + {0, 1}), which are exactly
// the A-fragment elements of k-steps 2c and 2c+1.
struct Encoding {
    float cs[4][3][2];   // [slot][coordinate][k = t, t + 4]
    float sn[4][3][2];
};

__device__ __forceinline__ void encode(const float (&a)[4][3], int t, Encoding& e) {
    const float f = (float)(1 << t);
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincos_cw(f * a[s][c], sn, cs);        // f * a is exact: matches fl(2^k * fl(pi * u))
            e.cs[s][c][0] = cs; e.sn[s][c][0] = sn;
#pragma unroll
            for (int d = 0; d < 4; ++d) {          // four octaves up by the double-angle recurrence
                const float s2 = 2.0f * sn * cs;
                cs = (cs - sn) * (cs + sn);
                sn = s2;
            }
            e.cs[s][c][1] = cs; e.sn[s][c][1] = sn;
        }
}

// Row `row` of the 32-row tile lives in quad (row & 7) as slot (row >> 3).  Move one per-row value
// from the fragment layout (v[slot], identical in the 4 lanes of a quad) to lane == row.
__device__ __forceinline__ float rows_to_lanes(const float (&v)[4], int lane) {
    const int t = lane & 3;
    const float mine = t == 0 ? v[0] : (t == 1 ? v[1] : (t == 2 ? v[2] : v[3]));
    return __shfl_sync(kFull, mine, 4 * (lane & 7) + (lane >> 3));
}

// The inverse: lane == row holds x; returns x of the 4 rows this lane owns in the fragment layout.
__device__ __forceinline__ void lanes_to_rows(float x, int lane, float (&v)[4]) {
    const int g = lane >> 2;
#pragma unroll
    for (int s = 0; s < 4; ++s) v[s] = __shfl_sync(kFull, x, 8 * s + g);
}

}  // namespace frag
}  // namespace vsrd
