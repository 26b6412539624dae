// Warp-level building blocks of the tensor-core field BACKWARD kernel (vsrd_field_bwd_mma.cu).
//
// One warp owns a tile of 16 (sample, instance) pairs of one instance.  All layer activations live in registers in the
// mma.sync accumulator ("C") layout and are chained into the next contraction as the A operand without leaving the
// register file:
//
//   lane = 4 g + t   (g = lane >> 2 in 0..7, t = lane & 3)
//   C fragment of an m16n8 tile:       c0 (row g, col 2t)  c1 (row g, col 2t+1)  c2 (row g+8, col 2t)  c3 (row g+8, col 2t+1)
//   A fragment of m16n8k16 (bf16 x 2): a0 (row g, k 2t..2t+1)  a1 (row g+8, k 2t..2t+1)  a2 (row g, k 2t+8..)  a3 (row g+8, k 2t+8..)
//   B fragment of m16n8k16:            b0 (k 2t..2t+1, n g)    b1 (k 2t+8..2t+9, n g)
//
// so the C fragments of n-tiles (0, 1) of a 16-channel layer ARE the A fragment of the next layer's single k-step once
// each (c0, c1) / (c2, c3) pair is packed to bf16 x 2.  Rows: a lane holds rows g and g + 8 of the tile; activation
// registers are indexed x[mt][nt][q] with q = 0: (c0, c1) of row g, q = 1: (c2, c3) of row g + 8.
//
// Precision: operands are bf16 hi + bf16 lo (16 mantissa bits), products hi*hi + hi*lo + lo*hi with fp32 accumulation:
// ~1e-5 per contraction, two orders below the 1e-3 gradient tolerance (measured at the BASELINE shapes: the gradients
// differ from the fp32 reference's by 1e-5 .. 1.5e-4 and track its own error against fp64, tests/test_gpu_fullsize.py).
#pragma once
#include "vsrd_common.cuh"

namespace vsrd {
namespace frag {

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2) --------------------------------------
// sm_100 issues one two-wide fp32 FMA per warp instruction (fma.rn.f32x2); a scalar or an immediate
// broadcasts for free (`R.F32` operand form).  The field kernels are issue-bound on fp32 element-wise
// work between the MMAs (profiles/r01_v5_*: 61 % FADD/FMUL/FFMA), so every per-channel formula is
// written on the accumulator register pairs (c0, c1) / (c2, c3) of the C layout.
using f2 = float2;
__device__ __forceinline__ f2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __ffma2_rn(b, bc(-1.0f), a); }   // a - b, one rounding
__device__ __forceinline__ float hsum(f2 a) { return a.x + a.y; }

// ---- shared-memory image of one instance's weights -------------------------------------------
// Contractions run on mma.sync m16n8k16 with bf16 hi + bf16 lo operands (products hi*hi + hi*lo + lo*hi, fp32
// accumulation: ~2^-16 per operand, two orders below the 1e-3 gradient tolerance; the forward kernel, which has to meet
// the 1e-4 silhouette bar, keeps 3xTF32 on tcgen05).  One k-step covers a whole 16-channel layer, so a layer costs
// 3 MMAs per 8 outputs instead of the 6 of 3xTF32 on m16n8k8 -- measured on B200: both shapes issue at the same rate
// (profiles/r01_pipe_rates_b200.txt), and the tensor pipe's issue slots were 25 % of this kernel's stall samples.
// The C fragment of n-tiles (0, 1) IS the A fragment of the next layer's k-step (a0 = pack(c0, c1) of n-tile 0 ...), so
// activations chain in registers exactly as before.
// uint4 {b0_hi, b1_hi, b0_lo, b1_lo} per (fragment, lane), b0 = B[k = 2t, 2t+1][n = g], b1 = B[k = 2t+8, 2t+9][n = g]:
constexpr int kB0 = 0;              // layer 0 forward:    [c 0..2][nt 0..1]   B[k][n] = W0[8nt+n][16c+k]
constexpr int kB1 = kB0 + 6;        // hidden forward:     [l-1][nt 0..1]      B[k][n] = Wl[8nt+n][k]
constexpr int kBR1 = kB1 + 6;       // hidden transposed:  [l-1][nt 0..1]      B[k][n] = Wl[k][8nt+n]
constexpr int kBR0 = kBR1 + 6;      // layer 0 transposed: [nt 0..5]           B[k][n] = W0[k][8nt+n]
constexpr int kNumFrag = kBR0 + 6;  // 24 fragments x 32 lanes x 16 B = 12 KB
constexpr int kFragFloat4 = kNumFrag * 32;
// followed by the biases in fp32: b0[16] b1[16] b2[16] b3[16] w4[16] b4
constexpr int kBias = 0;            // offsets into the float tail
constexpr int kTailW4 = 64;
constexpr int kTailB4 = 80;
constexpr int kTailFloats = 96;
constexpr size_t kWeightBytes = (size_t)kFragFloat4 * 16 + kTailFloats * 4;

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
    f2 h;
    h.x = __uint_as_float(hi << 16);
    h.y = __uint_as_float(hi & 0xffff0000u);
    const f2 rem = sub2(make_float2(x0, x1), h);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rem.y), "f"(rem.x));
}
__device__ __forceinline__ void pack_bf16_split_scalar(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ void pack_transposed(f2 x, uint32_t& hi, uint32_t& lo);

// Pair (x0, x1) = C-layout elements (sample g, channels 2t, 2t+1) of an 8x8 block -> transposed operand
// register (samples 2t, 2t+1; channel g), hi and lo parts.
__device__ __forceinline__ void pack_transposed(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    uint32_t h, l;
    pack_bf16_split(x0, x1, h, l);
    hi = movmatrix_trans(h);
    lo = movmatrix_trans(l);
}

__device__ __forceinline__ void pack_transposed(f2 x, uint32_t& hi, uint32_t& lo) { pack_transposed(x.x, x.y, hi, lo); }

__device__ __forceinline__ void mma_bf16(f2 (&d)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0].x), "+f"(d[0].y), "+f"(d[1].x), "+f"(d[1].y)
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_bf16_zero(f2 (&d)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0].x), "=f"(d[0].y), "=f"(d[1].x), "=f"(d[1].y)
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}

// ---- layer contractions: one k-step (16 channels) of FOUR accumulator tiles ------------------------------------------
// x0, x1 share the A operand X (n-tiles w0, w1), y0, y1 share Y.  The three passes (lo*hi, hi*lo, hi*hi) are issued
// pass-major, so two MMAs into the same accumulator are always three independent MMAs apart: a back-to-back chain stalls
// on the tensor pipe's result latency.  kZeroX / kZeroY: the accumulators start from zero (the C operand is the zero
// register, no clearing is issued) instead of their current contents (bias).
template <bool kZeroX, bool kZeroY>
__device__ __forceinline__ void mma3b_quad(f2 (&x0)[2], f2 (&x1)[2], f2 (&y0)[2], f2 (&y1)[2],
                                           const uint32_t (&xh)[4], const uint32_t (&xl)[4],
                                           const uint32_t (&yh)[4], const uint32_t (&yl)[4],
                                           const float4& w0, const float4& w1) {
    const uint32_t w0h0 = __float_as_uint(w0.x), w0h1 = __float_as_uint(w0.y), w0l0 = __float_as_uint(w0.z), w0l1 = __float_as_uint(w0.w);
    const uint32_t w1h0 = __float_as_uint(w1.x), w1h1 = __float_as_uint(w1.y), w1l0 = __float_as_uint(w1.z), w1l1 = __float_as_uint(w1.w);
    if (kZeroX) { mma_bf16_zero(x0, xl, w0h0, w0h1); mma_bf16_zero(x1, xl, w1h0, w1h1); }
    else { mma_bf16(x0, xl, w0h0, w0h1); mma_bf16(x1, xl, w1h0, w1h1); }
    if (kZeroY) { mma_bf16_zero(y0, yl, w0h0, w0h1); mma_bf16_zero(y1, yl, w1h0, w1h1); }
    else { mma_bf16(y0, yl, w0h0, w0h1); mma_bf16(y1, yl, w1h0, w1h1); }
    mma_bf16(x0, xh, w0l0, w0l1); mma_bf16(x1, xh, w1l0, w1l1);
    mma_bf16(y0, yh, w0l0, w0l1); mma_bf16(y1, yh, w1l0, w1l1);
    mma_bf16(x0, xh, w0h0, w0h1); mma_bf16(x1, xh, w1h0, w1h1);
    mma_bf16(y0, yh, w0h0, w0h1); mma_bf16(y1, yh, w1h0, w1h1);
}

// A fragment (hi, lo) of the k-step from an activation held as two C tiles c[nt][q] (q: row g | row g + 8):
// a0 = channels (2t, 2t+1) of row g, a1 = of row g + 8, a2 / a3 = channels (2t+8, 2t+9).
__device__ __forceinline__ void a_bf16_from_c(const f2 (&c)[2][2], uint32_t (&ah)[4], uint32_t (&al)[4]) {
    pack_bf16_split(c[0][0].x, c[0][0].y, ah[0], al[0]);
    pack_bf16_split(c[0][1].x, c[0][1].y, ah[1], al[1]);
    pack_bf16_split(c[1][0].x, c[1][0].y, ah[2], al[2]);
    pack_bf16_split(c[1][1].x, c[1][1].y, ah[3], al[3]);
}

// A operand (16 outputs x 16 samples of one m-tile) of the weight-gradient product, rows = outputs, k = samples: from
// the packs a_bf16_from_c() made of the adjoint for its transposed contraction (the packing, 5 instructions per
// register, is shared): in place, (p0, p1, p2, p3) = (nt0 g | nt0 g+8 | nt1 g | nt1 g+8) ->
// (T p0, T p2, T p1, T p3).
__device__ __forceinline__ void wgrad_a_from_packs(uint32_t (&h)[4], uint32_t (&l)[4]) {
    const uint32_t h1 = h[1], l1 = l[1];
    h[0] = movmatrix_trans(h[0]); l[0] = movmatrix_trans(l[0]);
    h[1] = movmatrix_trans(h[2]); l[1] = movmatrix_trans(l[2]);
    h[2] = movmatrix_trans(h1);   l[2] = movmatrix_trans(l1);
    h[3] = movmatrix_trans(h[3]); l[3] = movmatrix_trans(l[3]);
}

// D (16 outputs x 8 inputs) += A (adjoint) x B, B = one C tile (16 samples x 8 input channels).
// lo = channels (2t, 2t+1) of the samples in rows g, hi = the same channels in rows g + 8.
__device__ __forceinline__ void wgrad_tile(f2 (&D)[2], const uint32_t (&ah)[4], const uint32_t (&al)[4], f2 lo_rows, f2 hi_rows) {
    uint32_t bh0, bl0, bh1, bl1;
    pack_transposed(lo_rows, bh0, bl0);
    pack_transposed(hi_rows, bh1, bl1);
    mma_bf16(D, ah, bl0, bl1);
    mma_bf16(D, al, bh0, bh1);
    mma_bf16(D, ah, bh0, bh1);
}

// The same for a B tile whose elements are four unrelated registers (b0, b1: channels 2t, 2t+1 of row g;
// b2, b3: of row g + 8).
__device__ __forceinline__ void wgrad_tile_scalar(f2 (&D)[2], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                                  float b0, float b1, float b2, float b3) {
    uint32_t h, l, bh0, bl0, bh1, bl1;
    pack_bf16_split_scalar(b0, b1, h, l);
    bh0 = movmatrix_trans(h); bl0 = movmatrix_trans(l);
    pack_bf16_split_scalar(b2, b3, h, l);
    bh1 = movmatrix_trans(h); bl1 = movmatrix_trans(l);
    mma_bf16(D, ah, bl0, bl1);
    mma_bf16(D, al, bh0, bh1);
    mma_bf16(D, ah, bh0, bh1);
}

// D += A x ones: every column holds the sum over the 16 samples (bias gradient).
__device__ __forceinline__ void wgrad_bias(f2 (&D)[2], const uint32_t (&ah)[4], const uint32_t (&al)[4]) {
    constexpr uint32_t kOnes = 0x3f803f80u;
    mma_bf16(D, al, kOnes, kOnes);
    mma_bf16(D, ah, kOnes, kOnes);
}

// Weight gradient of one hidden layer over the 16 samples of an m-tile: D[0], D[1] (inputs 0-7, 8-15) += adjoint x
// activation + tangent adjoint x tangent activation, D[2] += adjoint x ones (bias).  g / gd: [row half][channel pair].
// The MMAs are issued pass-major over the accumulators: two MMAs into the same accumulator are never back to back
// (three-pass chains per accumulator wait on the tensor pipe's result latency: 1.3 % of the kernel).  Measured and not
// kept: separate accumulators for the tangent products (five independent chains + 4 adds): slower.
__device__ __forceinline__ void wgrad_hidden(f2 (&D)[3][2], const uint32_t (&th)[4], const uint32_t (&tl)[4],
                                             const uint32_t (&tdh)[4], const uint32_t (&tdl)[4],
                                             const f2 (&g)[2][2], const f2 (&gd)[2][2]) {
    constexpr uint32_t kOnes = 0x3f803f80u;
    uint32_t bh[2][2], bl[2][2];                           // [n-tile][row half]
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) pack_transposed(g[hf][nt], bh[nt][hf], bl[nt][hf]);
    mma_bf16(D[0], th, bl[0][0], bl[0][1]); mma_bf16(D[1], th, bl[1][0], bl[1][1]); mma_bf16(D[2], tl, kOnes, kOnes);
    mma_bf16(D[0], tl, bh[0][0], bh[0][1]); mma_bf16(D[1], tl, bh[1][0], bh[1][1]); mma_bf16(D[2], th, kOnes, kOnes);
    mma_bf16(D[0], th, bh[0][0], bh[0][1]); mma_bf16(D[1], th, bh[1][0], bh[1][1]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) pack_transposed(gd[hf][nt], bh[nt][hf], bl[nt][hf]);
    mma_bf16(D[0], tdh, bl[0][0], bl[0][1]); mma_bf16(D[1], tdh, bl[1][0], bl[1][1]);
    mma_bf16(D[0], tdl, bh[0][0], bh[0][1]); mma_bf16(D[1], tdl, bh[1][0], bh[1][1]);
    mma_bf16(D[0], tdh, bh[0][0], bh[0][1]); mma_bf16(D[1], tdh, bh[1][0], bh[1][1]);
}

// Reference weight layout (hyper_distance_field.py:57-73): layer l rows [out][fan_in + 1], bias last.
//
// LayerNorm centring folded into the weights (as in vsrd_field_umma.cu::stage_weights_umma): the fragments hold the CENTRED
// weights P W_l and biases P b_l of layers 0..3 (P = I - 11^T / 16 over the output index), so every contraction delivers
// centred pre-activations (and tangents), the transposed contractions apply P to the adjoints for free, and the
// LayerNorm code skips its mean reductions.  The weight gradient the kernel accumulates is then the one w.r.t. P W; the
// row reduction maps it back (dW = P dW', reduce_segment_rows_kernel).
constexpr int kNumMeans = (kEnc + 1) + 3 * (kHid + 1);
__device__ __forceinline__ float centred_weight(const float* __restrict__ W, int i, const float* s_mean) {
    if (i < kW1) return __ldg(W + i) - s_mean[i % (kEnc + 1)];
    const int q = i - kW1, l = q / kWStride, j = (q - l * kWStride) % (kHid + 1);
    return __ldg(W + i) - s_mean[(kEnc + 1) + l * (kHid + 1) + j];
}
__device__ __forceinline__ void stage_weight_fragments(const float* __restrict__ W, float4* sF, float* sTail, float* s_mean) {
    for (int i = threadIdx.x; i < kNumMeans; i += blockDim.x) {                     // column means over the 16 outputs
        const bool first = i <= kEnc;
        const int l = first ? 0 : (i - (kEnc + 1)) / (kHid + 1), j = first ? i : (i - (kEnc + 1)) % (kHid + 1);
        const float* col = first ? W + j : W + kW1 + l * kWStride + j;
        const int stride = first ? kEnc + 1 : kHid + 1;
        float sum = 0.0f;
#pragma unroll
        for (int o = 0; o < kHid; ++o) sum += __ldg(col + o * stride);
        s_mean[i] = sum * (1.0f / kHid);
    }
    __syncthreads();
    for (int f = threadIdx.x; f < kFragFloat4; f += blockDim.x) {
        const int lane = f & 31, frag = f >> 5;
        const int g = lane >> 2, t = lane & 3;
        int idx[4];                             // the four elements k = 2t, 2t+1, 2t+8, 2t+9 of column n = g
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = 2 * t + (e & 1) + 8 * (e >> 1);
            if (frag < kB1) {                   // layer 0 forward
                const int c = frag >> 1, nt = frag & 1;
                idx[e] = (8 * nt + g) * (kEnc + 1) + 16 * c + k;
            } else if (frag < kBR1) {           // hidden forward
                const int q = frag - kB1, l = q >> 1, nt = q & 1;
                idx[e] = kW1 + l * kWStride + (8 * nt + g) * (kHid + 1) + k;
            } else if (frag < kBR0) {           // hidden transposed
                const int q = frag - kBR1, l = q >> 1, nt = q & 1;
                idx[e] = kW1 + l * kWStride + k * (kHid + 1) + 8 * nt + g;
            } else {                            // layer 0 transposed
                const int nt = frag - kBR0;
                idx[e] = k * (kEnc + 1) + 8 * nt + g;
            }
        }
        uint32_t h0, l0, h1, l1;
        pack_bf16_split_scalar(centred_weight(W, idx[0], s_mean), centred_weight(W, idx[1], s_mean), h0, l0);
        pack_bf16_split_scalar(centred_weight(W, idx[2], s_mean), centred_weight(W, idx[3], s_mean), h1, l1);
        sF[f] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
    }
    for (int f = threadIdx.x; f < kTailFloats; f += blockDim.x) {
        float v = 0.0f;
        if (f < 16) v = centred_weight(W, f * (kEnc + 1) + kEnc, s_mean);
        else if (f < 64) v = centred_weight(W, kW1 + ((f >> 4) - 1) * kWStride + (f & 15) * (kHid + 1) + kHid, s_mean);
        else if (f <= kTailB4) v = __ldg(W + kW4 + (f - 64));
        sTail[f] = v;
    }
}

// ---- small helpers ---------------------------------------------------------------------------
// two independent quad sums at once: (x, y) -> (sum over the quad of x, of y)
__device__ __forceinline__ f2 quad_sum2(f2 v) {
    f2 o;
    o.x = __shfl_xor_sync(kFull, v.x, 1); o.y = __shfl_xor_sync(kFull, v.y, 1);
    v = add2(v, o);
    o.x = __shfl_xor_sync(kFull, v.x, 2); o.y = __shfl_xor_sync(kFull, v.y, 2);
    return add2(v, o);
}

// Phi, phi of the exact-erf GELU (vsrd_math.cuh::gelu_terms; Abramowitz-Stegun 7.1.26 with ex2 / rcp approximations) for a
// channel pair.  zz = z^2 is returned for gelu'' = phi (2 - z^2).  The polynomial coefficients carry the factor 1/2 of
// the tail; Phi = 1/2 + copysign(1/2 - tail, z).
__device__ __forceinline__ void gelu_terms2(f2 z, f2& Phi, f2& phi, f2& zz) {
    zz = mul2(z, z);
    const f2 arg = mul2(zz, bc(-0.72134752044448170368f));
    f2 E, t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E.x) : "f"(arg.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E.y) : "f"(arg.y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(fmaf(0.3275911f * kInvSqrt2, fabsf(z.x), 1.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(fmaf(0.3275911f * kInvSqrt2, fabsf(z.y), 1.0f)));
    f2 poly = fma2(bc(0.5f * 1.061405429f), t, bc(0.5f * -1.453152027f));
    poly = fma2(poly, t, bc(0.5f * 1.421413741f));
    poly = fma2(poly, t, bc(0.5f * -0.284496736f));
    poly = fma2(poly, t, bc(0.5f * 0.254829592f));
    const f2 tail = mul2(mul2(poly, t), E);
    f2 q = fma2(tail, bc(-1.0f), bc(0.5f));                    // 1/2 - tail >= 0
    q.x = __uint_as_float(__float_as_uint(q.x) | (__float_as_uint(z.x) & 0x80000000u));
    q.y = __uint_as_float(__float_as_uint(q.y) | (__float_as_uint(z.y) & 0x80000000u));
    Phi = add2(q, bc(0.5f));
    phi = mul2(E, bc(kInvSqrt2Pi));
}

// sin / cos of a pair of arguments: three-term Cody-Waite reduction (|x| < ~1e4) and the single-precision minimax
// polynomials on [-pi/4, pi/4]; absolute error ~1e-7, no slow path, no local memory.
__device__ __forceinline__ void sincos_cw2(f2 x, f2& s, f2& c) {
    const f2 kx = mul2(x, bc(0.63661977236758134308f));
    f2 kf;
    kf.x = rintf(kx.x); kf.y = rintf(kx.y);
    f2 r = fma2(kf, bc(-1.5707962513e+00f), x);
    r = fma2(kf, bc(-7.5497894159e-08f), r);
    r = fma2(kf, bc(-5.3903029534e-15f), r);
    const int q0 = (int)kf.x, q1 = (int)kf.y;
    const f2 r2 = mul2(r, r);
    f2 sp = fma2(r2, bc(-1.9515295891e-4f), bc(8.3321608736e-3f));
    sp = fma2(sp, r2, bc(-1.6666654611e-1f));
    sp = fma2(mul2(sp, r2), r, r);
    f2 cp = fma2(r2, bc(2.443315711809948e-5f), bc(-1.388731625493765e-3f));
    cp = fma2(cp, r2, bc(4.166664568298827e-2f));
    cp = fma2(cp, r2, bc(-0.5f));
    cp = fma2(cp, r2, bc(1.0f));
    {
        const float ss = (q0 & 1) ? cp.x : sp.x, cc = (q0 & 1) ? sp.x : cp.x;
        s.x = (q0 & 2) ? -ss : ss;
        c.x = ((q0 + 1) & 2) ? -cc : cc;
    }
    {
        const float ss = (q1 & 1) ? cp.y : sp.y, cc = (q1 & 1) ? sp.y : cp.y;
        s.y = (q1 & 2) ? -ss : ss;
        c.y = ((q1 + 1) & 2) ? -cc : cc;
    }
}

// Positional encoding of the warp tile in A-operand form: pairs run over the two rows (g, g + 8) of an
// m-tile, so cs[mt][c][f] / sn[mt][c][f] ARE the register pairs (a0, a1) / (a2, a3) of k-step 2c + f.
// Lane t owns frequencies k = t (f = 0) and k = t + 4 (f = 1) of every coordinate c.
template <int MT>
struct EncodingT {
    f2 cs[MT][3][2];
    f2 sn[MT][3][2];
};

// a[mt][c] = PE argument fl(pi * u_c) of rows (g, g + 8) of m-tile mt.
template <int MT>
__device__ __forceinline__ void encode2(const f2 (&a)[MT][3], int t, EncodingT<MT>& e) {
    const float f = (float)(1 << t);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            f2 sn, cs;
            sincos_cw2(mul2(a[mt][c], bc(f)), sn, cs);     // f * a is exact: matches fl(2^k * fl(pi * u))
            e.cs[mt][c][0] = cs; e.sn[mt][c][0] = sn;
#pragma unroll
            for (int d = 0; d < 4; ++d) {                  // four octaves up by the double-angle recurrence
                const f2 s2 = mul2(add2(sn, sn), cs);
                cs = mul2(sub2(cs, sn), add2(cs, sn));
                sn = s2;
            }
            e.cs[mt][c][1] = cs; e.sn[mt][c][1] = sn;
        }
}

// Tiles of 16 * MT rows: row r (= lane, r < 16 MT) lives in quad (r & 7) as slot (r >> 3); pairs are
// (slot 2 mt, slot 2 mt + 1) = rows (g, g + 8) of m-tile mt.
template <int MT>
__device__ __forceinline__ void lanes_to_row_pairs(float x, int lane, f2 (&v)[MT], int lane_offset = 0) {
    const int g = lane >> 2;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        v[mt].x = __shfl_sync(kFull, x, lane_offset + 16 * mt + g);
        v[mt].y = __shfl_sync(kFull, x, lane_offset + 16 * mt + 8 + g);
    }
}
template <int MT>
__device__ __forceinline__ float row_slots_to_lanes(const float (&v)[2 * MT], int lane) {
    const int t = lane & 3;
    float mine = v[0];
#pragma unroll
    for (int s = 1; s < 2 * MT; ++s) mine = (t == s) ? v[s] : mine;
    return __shfl_sync(kFull, mine, 4 * (lane & 7) + ((lane >> 3) & (2 * MT - 1)));
}

}  // namespace frag
}  // namespace vsrd
