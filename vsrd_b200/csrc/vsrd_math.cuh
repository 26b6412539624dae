// Per-thread math of the VSRD silhouette renderer hot path (fp32).
//
// Everything here is `__host__ __device__` so the exact code the sm_100a kernels execute per
// (sample, instance) / per sample can also be compiled with g++ and checked against the CPU oracle
// without a GPU (tests/hostsim).  Warp-level glue (scans, reductions, staging) lives in
// vsrd_kernels.cu.
//
// Reference semantics (paths relative to the upstream repository):
//   box SDF ................ vsrd/rendering/sdfs.py:5-37
//   positional encoding .... vsrd/models/encoders/sinusoidal_encoder.py:9-19
//   residual MLP ........... vsrd/models/fields/hyper_distance_field.py:57-73
//   residual / union ....... scripts/main.py:433-492
//   opacity / compositing .. vsrd/rendering/renderers.py:212-263
// The spatial gradient the reference obtains with autograd (renderers.py:218-228) is computed
// analytically here (forward-mode tangents in the forward kernels, one tangent + a reverse sweep in
// the backward kernel: SURVEY.md App. D.6).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define VSRD_HD __host__ __device__ __forceinline__
#else
#define VSRD_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define VSRD_UNROLL _Pragma("unroll")
#define VSRD_NOUNROLL _Pragma("unroll 1")
#define VSRD_UNROLL2 _Pragma("unroll 2")
#else
#define VSRD_UNROLL
#define VSRD_NOUNROLL
#define VSRD_UNROLL2
#endif

namespace vsrd {

// ---- residual MLP geometry (configs/kitti_360/vsrd/*/config.json:142-162) -------------------
constexpr int kFreq = 8;               // SinusoidalEncoder(num_frequencies=8)
constexpr int kEnc = 3 * 2 * kFreq;    // 48 input channels
constexpr int kHid = 16;               // hidden width
constexpr int kW0 = 0;                 // layer 0: [16][48+1]
constexpr int kW1 = kHid * (kEnc + 1); // 784: layers 1..3: [16][16+1]
constexpr int kWStride = kHid * (kHid + 1);   // 272
constexpr int kW4 = kW1 + 3 * kWStride;       // 1600: layer 4: [1][16+1]
constexpr int kNumW = kW4 + kHid + 1;         // 1617
constexpr int kNumPose = 15;                  // grads of t(3), half extents(3), R(9)
constexpr int kGradStride = 1632;             // 1617 + 15

constexpr float kPiF = 3.14159274101257324f;  // float32(pi), as `frequencies` holds it
constexpr float kLnEps = 1e-5f;               // F.layer_norm default eps
constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;

// Offset of layer l (1..3) in the flat weight vector.
VSRD_HD constexpr int layer_offset(int l) { return l == 0 ? kW0 : (l == 4 ? kW4 : kW1 + (l - 1) * kWStride); }

// Kernel-side ("staged") weight layout: every layer transposed to [fan_in + 1][fan_out] so that the
// 16 outputs fed by one input are contiguous (4 x 128-bit shared-memory loads).  Same total size.
VSRD_HD constexpr int staged_index(int flat) {
    if (flat < kW1) { int o = flat / (kEnc + 1), j = flat % (kEnc + 1); return j * kHid + o; }
    if (flat < kW4) {
        int l = (flat - kW1) / kWStride, r = (flat - kW1) % kWStride;
        int o = r / (kHid + 1), i = r % (kHid + 1);
        return kW1 + l * kWStride + i * kHid + o;
    }
    return flat;
}

struct Instance {
    float t[3];     // location
    float R[9];     // rotation, row-major; local p = (x - t) @ R   (sdfs.py:25,34)
    float dim[3];   // half extents
};

VSRD_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
VSRD_HD float reluf_(float x) { return x > 0.0f ? x : 0.0f; }
VSRD_HD float signf_(float x) { return (x > 0.0f ? 1.0f : 0.0f) - (x < 0.0f ? 1.0f : 0.0f); }

// torch.lerp(a, b, w) for scalar weight (ATen/native/Lerp.h)
VSRD_HD float lerpf_(float a, float b, float w) {
    float diff = b - a;
    return (fabsf(w) < 0.5f) ? a + w * diff : b - diff * (1.0f - w);
}

// ---------------------------------------------------------------------------------------------
// Box SDF in the local frame, value and gradient.
// ---------------------------------------------------------------------------------------------
struct BoxEval {
    float y[3];     // x - t
    float p[3];     // local position
    float s[3];     // sign(p)
    float a[3];     // relu(|p| - dim)
    float q[3];     // |p| - dim
    float nrm;      // sqrt(sum a^2 + 1e-6)
    float ind[3];   // 1 at argmax(q) if the point is inside (max q < 0)
    float value;
    float gp[3];    // d box / d p
};

VSRD_HD void box_eval(const float x[3], const Instance& I, BoxEval& b) {
    VSRD_UNROLL for (int m = 0; m < 3; ++m) b.y[m] = x[m] - I.t[m];
    VSRD_UNROLL for (int k = 0; k < 3; ++k)
        b.p[k] = b.y[0] * I.R[k] + b.y[1] * I.R[3 + k] + b.y[2] * I.R[6 + k];
    float sumsq = 0.0f;
    VSRD_UNROLL for (int k = 0; k < 3; ++k) {
        b.s[k] = signf_(b.p[k]);
        b.q[k] = fabsf(b.p[k]) - I.dim[k];
        b.a[k] = reluf_(b.q[k]);
        sumsq += b.a[k] * b.a[k];
    }
    b.nrm = sqrtf(sumsq + 1e-6f);
    // torch.max returns the first maximal index
    int kmax = 0;
    float mx = b.q[0];
    if (b.q[1] > mx) { mx = b.q[1]; kmax = 1; }
    if (b.q[2] > mx) { mx = b.q[2]; kmax = 2; }
    const bool inside = mx < 0.0f;
    b.value = b.nrm - reluf_(-mx);
    VSRD_UNROLL for (int k = 0; k < 3; ++k) {
        b.ind[k] = (inside && k == kmax) ? 1.0f : 0.0f;
        b.gp[k] = b.s[k] * (b.a[k] / b.nrm + b.ind[k]);
    }
}

// ---------------------------------------------------------------------------------------------
// Residual MLP: value + NT tangents (forward mode).
//   a[c]       = float32(pi) * u_c,  u = (|p_x|, p_y, p_z) / scale   (main.py:437-442)
//   adot[t][c] = d a_c along tangent t.  kDiagonal: tangent t only moves coordinate t.
// ---------------------------------------------------------------------------------------------
struct MlpStash {
    float e[kEnc];          // encoding (cos, sin pairs)
    float z[4][kHid];       // LayerNorm outputs feeding layers 1..4
    float zd[4][kHid];      // their tangents
    float r[4];             // 1/sqrt(var + eps)
    float m[4];             // mean(z * centred tangent)
};

// Phi(z) = standard normal CDF, phi(z) = its density: gelu(z) = z Phi, gelu' = Phi + z phi,
// gelu'' = phi (2 - z^2).  The reference uses the exact erf GELU (F.gelu default).  Here erf comes from
// Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32 rounding level), which shares its single
// exponential with phi: ~15 instructions instead of ~40 for erff + expf.
#if defined(__CUDA_ARCH__)
#define VSRD_EXPF(x) __expf(x)
#define VSRD_RCPF(x) __frcp_rn(x)
#else
#define VSRD_EXPF(x) expf(x)
#define VSRD_RCPF(x) (1.0f / (x))
#endif
VSRD_HD void gelu_terms(float z, float& Phi, float& phi) {
    const float az = fabsf(z);
    const float E = VSRD_EXPF(-0.5f * z * z);
    const float t = VSRD_RCPF(1.0f + (0.3275911f * kInvSqrt2) * az);
    float poly = 1.061405429f;
    poly = poly * t - 1.453152027f;
    poly = poly * t + 1.421413741f;
    poly = poly * t - 0.284496736f;
    poly = poly * t + 0.254829592f;
    const float tail = 0.5f * poly * t * E;        // 1 - Phi(|z|)
    Phi = z >= 0.0f ? 1.0f - tail : tail;
    phi = kInvSqrt2Pi * E;
}

template <int NT, bool kDiagonal, bool kStash>
VSRD_HD void mlp_forward_dual(const float* __restrict__ Wt, const float a[3], const float (*adot)[3],
                              float& out, float* outd, MlpStash& st) {
    float h[kHid], hd[NT][kHid];
    VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
        h[o] = Wt[kEnc * kHid + o];   // bias row
        VSRD_UNROLL for (int t = 0; t < NT; ++t) hd[t][o] = 0.0f;
    }
    // ---- layer 0 fused with the positional encoding
    VSRD_UNROLL for (int c = 0; c < 3; ++c) {
        VSRD_UNROLL for (int k = 0; k < kFreq; ++k) {
            const float f = (float)(1 << k);
            float sn, cs;
            sincosf(f * a[c], &sn, &cs);   // f * a is exact (power of two), matches fl(freq_k * u)
            const int j = c * 2 * kFreq + 2 * k;
            if (kStash) { st.e[j] = cs; st.e[j + 1] = sn; }
            const float* w0 = Wt + j * kHid;
            const float* w1 = w0 + kHid;
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) h[o] += w0[o] * cs + w1[o] * sn;
            VSRD_UNROLL for (int t = 0; t < NT; ++t) {
                if (kDiagonal && t != c) continue;
                const float da = f * adot[t][c];
                const float de0 = -da * sn, de1 = da * cs;
                VSRD_UNROLL for (int o = 0; o < kHid; ++o) hd[t][o] += w0[o] * de0 + w1[o] * de1;
            }
        }
    }
    // ---- layers 1..4: LayerNorm -> GELU -> linear
    VSRD_UNROLL for (int l = 1; l <= 4; ++l) {
        float mean = 0.0f;
        VSRD_UNROLL for (int o = 0; o < kHid; ++o) mean += h[o];
        mean *= (1.0f / kHid);
        float var = 0.0f;
        VSRD_UNROLL for (int o = 0; o < kHid; ++o) { h[o] -= mean; var += h[o] * h[o]; }
        var *= (1.0f / kHid);
        const float r = 1.0f / sqrtf(var + kLnEps);
        VSRD_UNROLL for (int o = 0; o < kHid; ++o) h[o] *= r;   // h now holds z
        VSRD_UNROLL for (int t = 0; t < NT; ++t) {
            float mt = 0.0f;
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) mt += hd[t][o];
            mt *= (1.0f / kHid);
            float mz = 0.0f;
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) { hd[t][o] -= mt; mz += h[o] * hd[t][o]; }
            mz *= (1.0f / kHid);
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) hd[t][o] = r * (hd[t][o] - h[o] * mz);   // zd
            if (kStash && t == 0) st.m[l - 1] = mz;
        }
        if (kStash) {
            st.r[l - 1] = r;
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) { st.z[l - 1][o] = h[o]; st.zd[l - 1][o] = hd[0][o]; }
        }
        // GELU (exact, erf) and its derivative
        VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
            float Phi, phi;
            gelu_terms(h[o], Phi, phi);
            const float g1 = Phi + h[o] * phi;
            h[o] = h[o] * Phi;
            VSRD_UNROLL for (int t = 0; t < NT; ++t) hd[t][o] *= g1;
        }
        if (l < 4) {
            const float* W = Wt + kW1 + (l - 1) * kWStride;
            float hn[kHid], hdn[NT][kHid];
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
                hn[o] = W[kHid * kHid + o];
                VSRD_UNROLL for (int t = 0; t < NT; ++t) hdn[t][o] = 0.0f;
            }
            VSRD_UNROLL for (int i = 0; i < kHid; ++i) {
                const float* w = W + i * kHid;
                VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
                    hn[o] += w[o] * h[i];
                    VSRD_UNROLL for (int t = 0; t < NT; ++t) hdn[t][o] += w[o] * hd[t][i];
                }
            }
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
                h[o] = hn[o];
                VSRD_UNROLL for (int t = 0; t < NT; ++t) hd[t][o] = hdn[t][o];
            }
        } else {
            const float* w = Wt + kW4;
            float acc = w[kHid];
            float accd[NT];
            VSRD_UNROLL for (int t = 0; t < NT; ++t) accd[t] = 0.0f;
            VSRD_UNROLL for (int i = 0; i < kHid; ++i) {
                acc += w[i] * h[i];
                VSRD_UNROLL for (int t = 0; t < NT; ++t) accd[t] += w[i] * hd[t][i];
            }
            out = acc;
            VSRD_UNROLL for (int t = 0; t < NT; ++t) outd[t] = accd[t];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Forward: one (sample, instance).  d = box (+ sigmoid(mlp - 1)), G = d d / d x  (world frame).
// ---------------------------------------------------------------------------------------------
template <bool kResidual>
VSRD_HD void field_forward(const float x[3], const Instance& I, const float* __restrict__ Wt,
                           float scale, float& d, float G[3]) {
    BoxEval b;
    box_eval(x, I, b);
    float gp[3] = {b.gp[0], b.gp[1], b.gp[2]};
    d = b.value;
    if (kResidual) {
        float a[3], adot[3][3];
        const float sx[3] = {b.s[0], 1.0f, 1.0f};
        const float m[3] = {fabsf(b.p[0]), b.p[1], b.p[2]};
        VSRD_UNROLL for (int c = 0; c < 3; ++c) {
            a[c] = kPiF * (m[c] / scale);
            VSRD_UNROLL for (int t = 0; t < 3; ++t) adot[t][c] = (t == c) ? sx[c] * (kPiF / scale) : 0.0f;
        }
        float out, outd[3];
        MlpStash dummy;
        mlp_forward_dual<3, true, false>(Wt, a, adot, out, outd, dummy);
        const float res = sigmoidf_(out - 1.0f);
        const float sp = res * (1.0f - res);
        d += res;
        VSRD_UNROLL for (int c = 0; c < 3; ++c) gp[c] += sp * outd[c];
    }
    VSRD_UNROLL for (int m = 0; m < 3; ++m)
        G[m] = I.R[3 * m] * gp[0] + I.R[3 * m + 1] * gp[1] + I.R[3 * m + 2] * gp[2];
}

// ---------------------------------------------------------------------------------------------
// Forward, looped form used by the kernel (v3).  Same mathematics as field_forward above, organised
// for a small instruction footprint: the profile of the fully unrolled version showed every kernel
// stalled on instruction fetch (`stalled_no_instruction` 4 cycles per issue, profiles/r01_v2_*), so
// layers and input neurons are real loops and the layer inputs (gelu outputs and their 3 tangents)
// live in 64 lane-private scratch floats `act[row * stride]` (shared memory on the device).
// The positional encoding takes accurate sincosf at frequencies k = 0 and 4 and the double-angle
// recurrence for the three octaves in between (absolute error <= 8 * 2^-24, far below the 1e-4 bar).
// ---------------------------------------------------------------------------------------------
constexpr int kFwdActRows = 4 * kHid;   // g, gd_x, gd_y, gd_z

template <bool kResidual>
VSRD_HD void field_forward_looped(const float x[3], const Instance& I, const float* Wt, float scale,
                                  float* act, int stride, float& d, float G[3]) {
    BoxEval b;
    box_eval(x, I, b);
    float gp[3] = {b.gp[0], b.gp[1], b.gp[2]};
    d = b.value;
    if (kResidual) {
        float a[3], coef[3];
        {
            const float sx[3] = {b.s[0], 1.0f, 1.0f};
            const float m[3] = {fabsf(b.p[0]), b.p[1], b.p[2]};
            VSRD_UNROLL for (int c = 0; c < 3; ++c) { a[c] = kPiF * (m[c] / scale); coef[c] = sx[c] * (kPiF / scale); }
        }
        float hv[kHid], ht[3][kHid];
        VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
            hv[o] = Wt[kEnc * kHid + o];
            ht[0][o] = 0.0f; ht[1][o] = 0.0f; ht[2][o] = 0.0f;
        }
        // ---- layer 0 fused with the positional encoding: loop over octaves
        float cs[3], sn[3];
        VSRD_NOUNROLL for (int k = 0; k < kFreq; ++k) {
            const float f = (float)(1 << k);
            if ((k & 3) == 0) {
                VSRD_UNROLL for (int c = 0; c < 3; ++c) sincosf(f * a[c], &sn[c], &cs[c]);   // f * a exact
            } else {
                VSRD_UNROLL for (int c = 0; c < 3; ++c) {
                    const float s2 = 2.0f * sn[c] * cs[c];
                    cs[c] = (cs[c] - sn[c]) * (cs[c] + sn[c]);
                    sn[c] = s2;
                }
            }
            VSRD_UNROLL for (int c = 0; c < 3; ++c) {
                const float da = f * coef[c];
                const float de0 = -da * sn[c], de1 = da * cs[c];
                const float* w0 = Wt + (c * 2 * kFreq + 2 * k) * kHid;
                const float* w1 = w0 + kHid;
                VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
                    hv[o] += w0[o] * cs[c] + w1[o] * sn[c];
                    ht[c][o] += w0[o] * de0 + w1[o] * de1;
                }
            }
        }
        // ---- layers 1..4: LayerNorm -> GELU (registers) -> linear (loop over inputs, from scratch)
        float out = 0.0f, outd[3] = {0.0f, 0.0f, 0.0f};
        VSRD_NOUNROLL for (int l = 1; l <= 4; ++l) {
            float mean = 0.0f;
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) mean += hv[o];
            mean *= (1.0f / kHid);
            float var = 0.0f;
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) { hv[o] -= mean; var += hv[o] * hv[o]; }
            const float r = 1.0f / sqrtf(var * (1.0f / kHid) + kLnEps);
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) hv[o] *= r;
            VSRD_UNROLL for (int t = 0; t < 3; ++t) {
                float mt = 0.0f;
                VSRD_UNROLL for (int o = 0; o < kHid; ++o) mt += ht[t][o];
                mt *= (1.0f / kHid);
                float mz = 0.0f;
                VSRD_UNROLL for (int o = 0; o < kHid; ++o) { ht[t][o] -= mt; mz += hv[o] * ht[t][o]; }
                mz *= (1.0f / kHid);
                VSRD_UNROLL for (int o = 0; o < kHid; ++o) ht[t][o] = r * (ht[t][o] - hv[o] * mz);
            }
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
                float Phi, phi;
                gelu_terms(hv[o], Phi, phi);
                const float g1 = Phi + hv[o] * phi;
                act[o * stride] = hv[o] * Phi;
                act[(kHid + o) * stride] = g1 * ht[0][o];
                act[(2 * kHid + o) * stride] = g1 * ht[1][o];
                act[(3 * kHid + o) * stride] = g1 * ht[2][o];
            }
            if (l < 4) {
                const float* W = Wt + kW1 + (l - 1) * kWStride;
                VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
                    hv[o] = W[kHid * kHid + o];
                    ht[0][o] = 0.0f; ht[1][o] = 0.0f; ht[2][o] = 0.0f;
                }
                VSRD_UNROLL2 for (int i = 0; i < kHid; ++i) {
                    const float g = act[i * stride], t0 = act[(kHid + i) * stride];
                    const float t1 = act[(2 * kHid + i) * stride], t2 = act[(3 * kHid + i) * stride];
                    const float* w = W + i * kHid;
                    VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
                        const float wo = w[o];
                        hv[o] += wo * g; ht[0][o] += wo * t0; ht[1][o] += wo * t1; ht[2][o] += wo * t2;
                    }
                }
            } else {
                const float* w = Wt + kW4;
                out = w[kHid];
                VSRD_UNROLL2 for (int i = 0; i < kHid; ++i) {
                    const float wi = w[i];
                    out += wi * act[i * stride];
                    outd[0] += wi * act[(kHid + i) * stride];
                    outd[1] += wi * act[(2 * kHid + i) * stride];
                    outd[2] += wi * act[(3 * kHid + i) * stride];
                }
            }
        }
        const float res = sigmoidf_(out - 1.0f);
        const float sp = res * (1.0f - res);
        d += res;
        VSRD_UNROLL for (int c = 0; c < 3; ++c) gp[c] += sp * outd[c];
    }
    VSRD_UNROLL for (int m = 0; m < 3; ++m)
        G[m] = I.R[3 * m] * gp[0] + I.R[3 * m + 1] * gp[1] + I.R[3 * m + 2] * gp[2];
}

// ---------------------------------------------------------------------------------------------
// Backward: one (sample, instance).  Given adjoints (dd, dG) of (d, G) accumulate the gradient of
//   phi = dd * d(p; theta) + (R^T dG) . grad_p d(p; theta)
// w.r.t. t, half extents, R and the MLP weights.  "One tangent + one reverse sweep".
//
// The Sink receives per-sample contributions:
//   sink.layer0(hbar[16], hdbar[16], e[48], ed[48])      -> dW0[o][j] = hbar[o] e[j] + hdbar[o] ed[j]; bias hbar[o]
//   sink.hidden(l, hbar[16], hdbar[16], g[16], gd[16])   -> dWl[o][i] likewise (l = 1..3)
//   sink.last(obar, odbar, g[16], gd[16])                -> dW4[i] = obar g[i] + odbar gd[i]; bias obar
//   sink.pose(tbar[3], dimbar[3], Rbar[9])
// ---------------------------------------------------------------------------------------------
VSRD_HD void ln_gelu_reverse(const float z[kHid], const float zd[kHid], float r, float m,
                             const float gbar[kHid], const float gdbar[kHid],
                             const float g1[kHid], const float g2[kHid],
                             float hbar[kHid], float hdbar[kHid]) {
    float zbar[kHid], zdbar[kHid];
    float s_zb = 0.0f, s_zzb = 0.0f, s_zdb = 0.0f, s_zzdb = 0.0f, s_zdzdb = 0.0f;
    VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
        zbar[o] = gbar[o] * g1[o] + gdbar[o] * g2[o] * zd[o];
        zdbar[o] = gdbar[o] * g1[o];
        s_zb += zbar[o];
        s_zzb += z[o] * zbar[o];
        s_zdb += zdbar[o];
        s_zzdb += z[o] * zdbar[o];
        s_zdzdb += zd[o] * zdbar[o];
    }
    const float inv = 1.0f / kHid;
    s_zb *= inv; s_zzb *= inv; s_zdb *= inv; s_zzdb *= inv; s_zdzdb *= inv;
    VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
        hdbar[o] = r * (zdbar[o] - s_zdb - z[o] * s_zzdb);
        hbar[o] = r * (zbar[o] - s_zb - z[o] * s_zzb) - r * s_zdzdb * z[o] - r * m * hdbar[o] - r * s_zzdb * zd[o];
    }
}

template <class Sink>
VSRD_HD void mlp_reverse(const float* __restrict__ Wt, const MlpStash& st, const float adot[3],
                         float obar, float odbar, float abar[3], float adbar[3], Sink& sink) {
    float hbar[kHid], hdbar[kHid];
    // ---- layer 4 (16 -> 1)
    {
        float g[kHid], gd[kHid], g1[kHid], g2[kHid], gbar[kHid], gdbar[kHid];
        const float* w = Wt + kW4;
        VSRD_UNROLL for (int i = 0; i < kHid; ++i) {
            float Phi, phi;
            const float z = st.z[3][i];
            gelu_terms(z, Phi, phi);
            g[i] = z * Phi;
            g1[i] = Phi + z * phi;
            g2[i] = phi * (2.0f - z * z);
            gd[i] = g1[i] * st.zd[3][i];
            gbar[i] = w[i] * obar;
            gdbar[i] = w[i] * odbar;
        }
        sink.last(obar, odbar, g, gd);
        ln_gelu_reverse(st.z[3], st.zd[3], st.r[3], st.m[3], gbar, gdbar, g1, g2, hbar, hdbar);
    }
    // ---- layers 3..1 (16 -> 16)
    VSRD_UNROLL for (int l = 3; l >= 1; --l) {
        float g[kHid], gd[kHid], g1[kHid], g2[kHid], gbar[kHid], gdbar[kHid];
        const float* W = Wt + kW1 + (l - 1) * kWStride;
        VSRD_UNROLL for (int i = 0; i < kHid; ++i) {
            float Phi, phi;
            const float z = st.z[l - 1][i];
            gelu_terms(z, Phi, phi);
            g[i] = z * Phi;
            g1[i] = Phi + z * phi;
            g2[i] = phi * (2.0f - z * z);
            gd[i] = g1[i] * st.zd[l - 1][i];
            const float* w = W + i * kHid;
            float s0 = 0.0f, s1 = 0.0f;
            VSRD_UNROLL for (int o = 0; o < kHid; ++o) { s0 += w[o] * hbar[o]; s1 += w[o] * hdbar[o]; }
            gbar[i] = s0;
            gdbar[i] = s1;
        }
        sink.hidden(l, hbar, hdbar, g, gd);
        float hb2[kHid], hdb2[kHid];
        ln_gelu_reverse(st.z[l - 1], st.zd[l - 1], st.r[l - 1], st.m[l - 1], gbar, gdbar, g1, g2, hb2, hdb2);
        VSRD_UNROLL for (int o = 0; o < kHid; ++o) { hbar[o] = hb2[o]; hdbar[o] = hdb2[o]; }
    }
    // ---- layer 0 + positional encoding
    {
        float ed[kEnc];
        VSRD_UNROLL for (int c = 0; c < 3; ++c) {
            float ab = 0.0f, adb = 0.0f;
            VSRD_UNROLL for (int k = 0; k < kFreq; ++k) {
                const float f = (float)(1 << k);
                const int j = c * 2 * kFreq + 2 * k;
                const float cs = st.e[j], sn = st.e[j + 1];
                const float da = f * adot[c];
                ed[j] = -da * sn;
                ed[j + 1] = da * cs;
                const float* w0 = Wt + j * kHid;
                const float* w1 = w0 + kHid;
                float eb0 = 0.0f, eb1 = 0.0f, edb0 = 0.0f, edb1 = 0.0f;
                VSRD_UNROLL for (int o = 0; o < kHid; ++o) {
                    eb0 += w0[o] * hbar[o];
                    eb1 += w1[o] * hbar[o];
                    edb0 += w0[o] * hdbar[o];
                    edb1 += w1[o] * hdbar[o];
                }
                ab += f * (-eb0 * sn + eb1 * cs - da * (edb0 * cs + edb1 * sn));
                adb += f * (-edb0 * sn + edb1 * cs);
            }
            abar[c] = ab;
            adbar[c] = adb;
        }
        sink.layer0(hbar, hdbar, st.e, ed);
    }
}

template <bool kResidual, class Sink>
VSRD_HD void field_backward(const float x[3], const Instance& I, const float* __restrict__ Wt,
                            float scale, float dd, const float dG[3], Sink& sink) {
    BoxEval b;
    box_eval(x, I, b);
    // tangent direction in the local frame: v = R^T dG
    float v[3];
    VSRD_UNROLL for (int k = 0; k < 3; ++k) v[k] = I.R[k] * dG[0] + I.R[3 + k] * dG[1] + I.R[6 + k] * dG[2];

    float pbar[3], vbar[3], dimbar[3];
    // ---- box part: phi_box = dd * box + v . grad_p box
    float vs = 0.0f;
    VSRD_UNROLL for (int k = 0; k < 3; ++k) vs += v[k] * b.s[k] * b.a[k];
    const float inv_n = 1.0f / b.nrm;
    const float inv_n3 = inv_n * inv_n * inv_n;
    VSRD_UNROLL for (int k = 0; k < 3; ++k) {
        const float act = b.q[k] > 0.0f ? 1.0f : 0.0f;
        const float hess = act * v[k] * b.s[k] * inv_n - b.a[k] * vs * inv_n3;
        pbar[k] = dd * b.gp[k] + b.s[k] * hess;
        dimbar[k] = -(dd * (b.a[k] * inv_n + b.ind[k]) + hess);
        vbar[k] = b.gp[k];
    }
    if (kResidual) {
        float a[3], adot1[1][3], coef[3];
        const float sx[3] = {b.s[0], 1.0f, 1.0f};
        const float m[3] = {fabsf(b.p[0]), b.p[1], b.p[2]};
        VSRD_UNROLL for (int c = 0; c < 3; ++c) {
            a[c] = kPiF * (m[c] / scale);
            coef[c] = sx[c] * (kPiF / scale);
            adot1[0][c] = coef[c] * v[c];
        }
        MlpStash st;
        float out, outd[1];
        mlp_forward_dual<1, false, true>(Wt, a, adot1, out, outd, st);
        const float res = sigmoidf_(out - 1.0f);
        const float sp = res * (1.0f - res);
        const float spp = sp * (1.0f - 2.0f * res);
        const float obar = dd * sp + spp * outd[0];
        const float odbar = sp;
        float abar[3], adbar[3];
        mlp_reverse(Wt, st, adot1[0], obar, odbar, abar, adbar, sink);
        VSRD_UNROLL for (int c = 0; c < 3; ++c) {
            pbar[c] += abar[c] * coef[c];
            vbar[c] += adbar[c] * coef[c];
        }
    }
    // ---- chain to the instance pose: p = R^T (x - t), v = R^T dG
    float tbar[3], Rbar[9];
    VSRD_UNROLL for (int m = 0; m < 3; ++m) {
        tbar[m] = -(I.R[3 * m] * pbar[0] + I.R[3 * m + 1] * pbar[1] + I.R[3 * m + 2] * pbar[2]);
        VSRD_UNROLL for (int k = 0; k < 3; ++k) Rbar[3 * m + k] = b.y[m] * pbar[k] + dG[m] * vbar[k];
    }
    sink.pose(tbar, dimbar, Rbar);
}

// ---------------------------------------------------------------------------------------------
// Per-sample union over instances + SDF -> opacity.  `F` holds (d_i, G_i) as float4 with a stride of
// `stride` elements between instances.
// ---------------------------------------------------------------------------------------------
struct Vec4 { float x, y, z, w; };

struct UnionEval {
    float mneg;    // max_i(-d_i / T)
    float Z;       // sum_i exp(-d_i/T - mneg)
    float dbar;    // union SDF
    float g[3];    // its spatial gradient
};

template <class Load>
VSRD_HD void union_forward(Load load, int N, float T, UnionEval& u) {
    float mneg = -INFINITY;
    for (int i = 0; i < N; ++i) mneg = fmaxf(mneg, -(load(i).x / T));
    float Z = 0.0f, ds = 0.0f;
    for (int i = 0; i < N; ++i) {
        const float di = load(i).x;
        const float e = expf(-(di / T) - mneg);
        Z += e;
        ds += e * di;
    }
    const float dbar = ds / Z;
    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
    for (int i = 0; i < N; ++i) {
        const Vec4 f = load(i);
        const float w = expf(-(f.x / T) - mneg) / Z;
        const float c = w * (1.0f - (f.x - dbar) / T);
        g0 += c * f.y; g1 += c * f.z; g2 += c * f.w;
    }
    u.mneg = mneg; u.Z = Z; u.dbar = dbar;
    u.g[0] = g0; u.g[1] = g1; u.g[2] = g2;
}

// Register-resident variant for the compositing kernels (N <= 16): every instance's field value is loaded once
// (all loads in flight together), the IEEE division d_i / T — the one rounding that is amplified by exp — and
// exp(-d_i/T - mneg) are evaluated ONCE per instance and kept (q[], e[]); the remaining scalar divisions
// (1/Z, 1/T applied to the cancellation-free (d_i - dbar)) become reciprocal multiplications, 1 ulp apart from
// union_forward.  The 4-pass form above costs ~6 divisions + 4 exp per (sample, instance): 80 % of the
// compositing kernels' instructions (profiles/r01_v10_hotspots.txt).
template <int NMAX>
struct UnionRegs {
    Vec4 f[NMAX];     // (d_i, grad d_i)
    float e[NMAX];    // exp(-d_i/T - mneg), 0 for i >= N
    float invZ, invT;
};

template <int NMAX>
VSRD_HD void union_forward_regs(UnionRegs<NMAX>& R, int N, float T, UnionEval& u) {
    float q[NMAX];
    float mneg = -INFINITY;
    VSRD_UNROLL for (int i = 0; i < NMAX; ++i) {
        q[i] = -INFINITY;
        if (i < N) { q[i] = -(R.f[i].x / T); mneg = fmaxf(mneg, q[i]); }
    }
    float Z = 0.0f, ds = 0.0f;
    VSRD_UNROLL for (int i = 0; i < NMAX; ++i) {
        R.e[i] = 0.0f;
        if (i < N) {
            R.e[i] = expf(q[i] - mneg);
            Z += R.e[i];
            ds += R.e[i] * R.f[i].x;
        }
    }
    const float dbar = ds / Z;
    R.invZ = 1.0f / Z;
    R.invT = 1.0f / T;
    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
    VSRD_UNROLL for (int i = 0; i < NMAX; ++i) if (i < N) {
        const float w = R.e[i] * R.invZ;
        const float c = w * (1.0f - (R.f[i].x - dbar) * R.invT);
        g0 += c * R.f[i].y; g1 += c * R.f[i].z; g2 += c * R.f[i].w;
    }
    u.mneg = mneg; u.Z = Z; u.dbar = dbar;
    u.g[0] = g0; u.g[1] = g1; u.g[2] = g2;
}

struct OpacityEval {
    float gn, inv, cs;      // |g|, 1/max(|g|,1e-12), dir . n
    float n[3];
    float Pp, Pn;           // logistic CDF at the section ends (prev / next)
    float alpha;
};

VSRD_HD void opacity_forward(const UnionEval& u, const float dir[3], float delta, float sigma, float rho,
                             float eps, OpacityEval& o) {
    o.gn = sqrtf(u.g[0] * u.g[0] + u.g[1] * u.g[1] + u.g[2] * u.g[2]);
    o.inv = 1.0f / fmaxf(o.gn, 1e-12f);
    VSRD_UNROLL for (int k = 0; k < 3; ++k) o.n[k] = u.g[k] * o.inv;
    o.cs = dir[0] * o.n[0] + dir[1] * o.n[1] + dir[2] * o.n[2];
    const float chat = -lerpf_(reluf_(-o.cs * 0.5f + 0.5f), reluf_(-o.cs), rho);
    const float half = chat * delta / 2.0f;
    o.Pp = sigmoidf_((u.dbar - half) / sigma);
    o.Pn = sigmoidf_((u.dbar + half) / sigma);
    o.alpha = reluf_((o.Pp - o.Pn) / (o.Pp + eps));
}

// Adjoint of opacity_forward: given alpha_bar returns the adjoint of dbar and of g (cosine path only).
VSRD_HD void opacity_backward(const OpacityEval& o, const float dir[3], float delta, float sigma, float rho,
                              float eps, float alpha_bar, float& dbar_adj, float gbar[3]) {
    float Ppb = 0.0f, Pnb = 0.0f;
    if (o.alpha > 0.0f) {
        const float den = o.Pp + eps;
        Ppb = alpha_bar * (o.Pn + eps) / (den * den);
        Pnb = -alpha_bar / den;
    }
    const float spb = Ppb * o.Pp * (1.0f - o.Pp) / sigma;
    const float snb = Pnb * o.Pn * (1.0f - o.Pn) / sigma;
    dbar_adj = spb + snb;
    const float chat_adj = (snb - spb) * delta * 0.5f;
    // d chat / d cs, following torch.lerp's two branches (both have slope (1-rho), rho on the ends)
    const float dchat = (1.0f - rho) * 0.5f * ((-o.cs * 0.5f + 0.5f) > 0.0f ? 1.0f : 0.0f)
                      + rho * (-o.cs > 0.0f ? 1.0f : 0.0f);
    const float cs_adj = chat_adj * dchat;
    const float k = (o.gn > 1e-12f) ? cs_adj * o.inv : 0.0f;
    VSRD_UNROLL for (int c = 0; c < 3; ++c) gbar[c] = k * (dir[c] - o.cs * o.n[c]);
}

// Adjoint of union_forward for one sample: given the adjoint of dbar (`dbar_adj`), of g (`gbar`) and of
// the softmin weights (`wbar(i)`, label path), emit per-instance adjoints (dd_i, dG_i).
template <class Load, class WBar, class Store>
VSRD_HD void union_backward(Load load, WBar wbar, Store store, int N, float T, const UnionEval& u,
                            float dbar_adj, const float gbar[3]) {
    float S1 = 0.0f, S2 = 0.0f, S3 = 0.0f;
    for (int i = 0; i < N; ++i) {
        const Vec4 f = load(i);
        const float w = expf(-(f.x / T) - u.mneg) / u.Z;
        const float ui = 1.0f - (f.x - u.dbar) / T;
        const float gam = gbar[0] * f.y + gbar[1] * f.z + gbar[2] * f.w;
        S1 += wbar(i) * w;
        S2 += gam * ui * w;
        S3 += gam * w;
    }
    const float invT = 1.0f / T;
    for (int i = 0; i < N; ++i) {
        const Vec4 f = load(i);
        const float w = expf(-(f.x / T) - u.mneg) / u.Z;
        const float ui = 1.0f - (f.x - u.dbar) / T;
        const float c = w * ui;
        const float gam = gbar[0] * f.y + gbar[1] * f.z + gbar[2] * f.w;
        const float lam = gam * ui;
        Vec4 a;
        a.x = dbar_adj * c - w * invT * (wbar(i) - S1) - w * invT * (lam - S2) - w * gam * invT + c * S3 * invT;
        a.y = c * gbar[0]; a.z = c * gbar[1]; a.w = c * gbar[2];
        store(i, a);
    }
}

// union_backward on the register-resident values of union_forward_regs.
template <int NMAX, class WBar, class Store>
VSRD_HD void union_backward_regs(const UnionRegs<NMAX>& R, WBar wbar, Store store, int N, const UnionEval& u,
                                 float dbar_adj, const float gbar[3]) {
    float S1 = 0.0f, S2 = 0.0f, S3 = 0.0f;
    float w[NMAX], ui[NMAX], gam[NMAX];
    VSRD_UNROLL for (int i = 0; i < NMAX; ++i) if (i < N) {
        w[i] = R.e[i] * R.invZ;
        ui[i] = 1.0f - (R.f[i].x - u.dbar) * R.invT;
        gam[i] = gbar[0] * R.f[i].y + gbar[1] * R.f[i].z + gbar[2] * R.f[i].w;
        S1 += wbar(i) * w[i];
        S2 += gam[i] * ui[i] * w[i];
        S3 += gam[i] * w[i];
    }
    const float invT = R.invT;
    VSRD_UNROLL for (int i = 0; i < NMAX; ++i) if (i < N) {
        const float c = w[i] * ui[i];
        const float lam = gam[i] * ui[i];
        Vec4 a;
        a.x = dbar_adj * c - w[i] * invT * (wbar(i) - S1) - w[i] * invT * (lam - S2) - w[i] * gam[i] * invT + c * S3 * invT;
        a.y = c * gbar[0]; a.z = c * gbar[1]; a.w = c * gbar[2];
        store(i, a);
    }
}

}  // namespace vsrd
