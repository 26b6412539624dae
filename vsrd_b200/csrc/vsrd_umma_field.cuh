// Per-thread math and tcgen05 helpers shared by the thread-per-sample field kernels (vsrd_field_umma.cu: forward,
// vsrd_field_bwd_umma.cu: backward).  One thread owns one sample; the 16 channels of a hidden layer live in 8 packed
// f32x2 registers; 16-wide contractions go through TMEM (A operand) x shared-memory B blocks in 3xTF32.
#pragma once
#include "vsrd_common.cuh"
#include "vsrd_umma.cuh"

namespace vsrd {
namespace fu {

using namespace umma;

constexpr int kGroupThreads = 128;           // one tile group = four warps = 128 TMEM lanes
constexpr int kTile = 128;                   // samples per tile
constexpr int kBlk = 256;                    // floats of one [16 x 16] K-major B block (stored hi then lo: 2 * kBlk)

__host__ __device__ constexpr int b_index(int n, int k) { return (k / 4) * 64 + (n / 8) * 32 + (n % 8) * 4 + (k % 4); }   // N = 16
constexpr uint32_t kLbo = 256, kSbo = 128;                // bytes: K-chunk stride, 8-row group stride

__device__ __forceinline__ void put_split(float* block, int n, int k, float v) {
    float hi, lo;
    split_tf32(v, hi, lo);
    block[b_index(n, k)] = hi;
    block[kBlk + b_index(n, k)] = lo;
}

// ---- per-thread math ---------------------------------------------------------------------------------------------
// sin / cos with a three-term Cody-Waite reduction and minimax polynomials on [-pi/4, pi/4] (abs. error ~1e-7)
__device__ __forceinline__ void sincos_cw(float x, float& s, float& c) {
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
    const float ss = (q & 1) ? cp : sp;
    const float cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

// e[2k] = cos(2^k a), e[2k+1] = sin(2^k a), k = 0..7: accurate anchors at k = 0 and k = 4 (2^k a is exact, so the
// argument equals fl(freq_k * u) of sinusoidal_encoder.py:16), three double-angle steps after each anchor
__device__ __forceinline__ void encode16(float a, float2 (&e)[8]) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float sn, cs;
        sincos_cw(half ? 16.0f * a : a, sn, cs);
        e[4 * half] = make_float2(cs, sn);
#pragma unroll
        for (int d = 1; d < 4; ++d) {
            const float s2 = 2.0f * sn * cs;
            cs = (cs - sn) * (cs + sn);
            sn = s2;
            e[4 * half + d] = make_float2(cs, sn);
        }
    }
}

// Two-wide fp32 arithmetic (fma.rn.f32x2 & co.: one issue slot for two channels)
using f2 = float2;
__device__ __forceinline__ f2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }

// Phi, phi of the exact-erf GELU for a channel pair (Abramowitz-Stegun 7.1.26 with ex2 / rcp approximations,
// |err| < 1.5e-7; the polynomial carries the factor 1/2 of the tail): Phi = 1/2 + copysign(1/2 - tail, z)
__device__ __forceinline__ void gelu_terms2(f2 z, f2& Phi, f2& phi) {
    const f2 arg = mul2(mul2(z, z), bc(-0.72134752044448170368f));
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

// h -> z = LayerNorm(h) (no affine, eps 1e-5), a = gelu(z), g = gelu'(z) / sigma; channel pairs (2i, 2i + 1).
// h arrives CENTRED (mean over the 16 channels == 0 up to rounding): the producing layer's weights and bias were
// centred when they were staged (vsrd_field_umma.cu::stage_weights_umma).
__device__ __forceinline__ void norm_gelu(const f2 (&h)[8], f2 (&z)[8], f2 (&a)[8], f2 (&g)[8]) {
    f2 var = bc(0.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) var = fma2(h[i], h[i], var);
    const float rs1 = rsqrtf((var.x + var.y) * (1.0f / 16.0f) + kLnEps);
    const f2 rs = bc(rs1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        z[i] = mul2(h[i], rs);
        f2 Phi, phi;
        gelu_terms2(z[i], Phi, phi);
        a[i] = mul2(z[i], Phi);
        g[i] = mul2(fma2(z[i], phi, Phi), rs);
    }
}

// adjoint of a = gelu(LayerNorm(h)) w.r.t. the CENTRED h:  zb = abar * gelu'(z) / sigma;  hbar = zb - z mean(z zb).
// (The full adjoint also subtracts mean(zb); that projection lives in the centred transposed weights the result is
// multiplied with next -- or, at layer 0, in the centred derivative weights behind g_c.)
__device__ __forceinline__ void norm_gelu_adjoint(const f2 (&abar)[8], const f2 (&z)[8], const f2 (&g)[8], f2 (&hbar)[8]) {
    f2 m2 = bc(0.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        hbar[i] = mul2(abar[i], g[i]);
        m2 = fma2(z[i], hbar[i], m2);
    }
    const f2 s2 = bc((m2.x + m2.y) * (-1.0f / 16.0f));
#pragma unroll
    for (int i = 0; i < 8; ++i) hbar[i] = fma2(z[i], s2, hbar[i]);
}

// v (8 channel pairs) -> hi / lo columns of the group's A operand
__device__ __forceinline__ void store_operand(uint32_t lane_base, int col_hi, int col_lo, const f2 (&v)[8]) {
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const f2 h = make_float2(__uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u));
        const f2 l = fma2(h, bc(-1.0f), v[i]);                 // exact: v - hi
        hi[2 * i] = h.x; hi[2 * i + 1] = h.y;
        lo[2 * i] = l.x; lo[2 * i + 1] = l.y;
    }
    tmem_st16(lane_base + col_hi, hi);
    tmem_st16(lane_base + col_lo, lo);
}
__device__ __forceinline__ void load_pairs(uint32_t taddr, f2 (&v)[8]) {
    float t[16];
    tmem_ld16(taddr, t);
    wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(t[2 * i], t[2 * i + 1]);
}

// The start address sits in the low 14 bits (in 16-byte units) of a shared-memory descriptor, so the descriptor of
// "base + byte offset" is one 64-bit add of a compile-time constant to the descriptor of the weight arena's base.
__device__ __forceinline__ uint64_t desc_at(uint64_t base_desc, int float_offset) { return base_desc + (uint64_t)(float_offset * 4 >> 4); }

// D[tmem_d, 16 columns] (+)= A(hi at a_hi, lo at a_lo; 16 columns = 2 k-steps) * B(block at float offset: hi, lo)^T, 3xTF32
__device__ __forceinline__ void mma3_16x16(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint64_t base_desc, int block, uint32_t idesc, bool accumulate) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const uint64_t bhi = desc_at(base_desc, block + ks * 128);                      // k-step = two 64-float K chunks
        const uint64_t blo = desc_at(base_desc, block + kBlk + ks * 128);
        mma_tf32_ts(tmem_d, a_hi + ks * 8, bhi, idesc, accumulate || ks > 0);
        mma_tf32_ts(tmem_d, a_lo + ks * 8, bhi, idesc, true);
        mma_tf32_ts(tmem_d, a_hi + ks * 8, blo, idesc, true);
    }
}
// accumulator columns + bias -> channel pairs
__device__ __forceinline__ void load_biased(uint32_t taddr, const float* bias, f2 (&v)[8]) {
    load_pairs(taddr, v);
    const float2* b = reinterpret_cast<const float2*>(bias);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = add2(v[i], b[i]);
}


}  // namespace fu
}  // namespace vsrd
