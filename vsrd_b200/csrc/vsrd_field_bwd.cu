// sm_100a kernels of the VSRD silhouette-renderer hot path, part 3/3: per-(sample, instance) field
// backward (single-tangent forward + reverse sweep) and the parameter-gradient reduction.
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
namespace bwd2 {

constexpr int kRow = 36;             // 32 samples + 4 pad: conflict-free for lane-private and fragment access
constexpr int kRowsPerWarp = 160;
constexpr int kRowZ = 0;             // z_l  rows [16 l, 16 l + 16), l = 0..3 (input of layer l+1)
constexpr int kRowZd = 64;           // zd_l rows
constexpr int kRowA = 128;           // staged adjoints: hbar rows [128,144), hdbar rows [144,160)
constexpr int kWarpFloats = kRowsPerWarp * kRow;
constexpr size_t kSmemBytes = (size_t)(kWarps * kWarpFloats + (kNumW + 3) + kGradStride) * sizeof(float);

struct WgradAcc {
    float h[3][3][4];   // hidden layers 1..3: n-tiles {inputs 0-7, inputs 8-15, bias}
    float l0[7][4];     // layer 0: 6 input n-tiles + bias
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// D[nt] += A_v^T B_v + A_t^T B_t over this warp's 32 samples; D[NT] += bias column (sum of A_v).
// A rows: [0,16) value adjoints, [16,32) tangent adjoints.  Bv/Bt: 8*NT rows each.
template <int NT>
__device__ __forceinline__ void wgrad_mma(const float* A, const float* Bv,
                                          const float* Bt, float (&D)[NT + 1][4], int lane) {
    const int gid = lane >> 2, tig = lane & 3;
    const uint32_t one = (gid == 0) ? 0x3f800000u : 0u;   // B fragment of the all-ones bias column
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        const float* B = part ? Bt : Bv;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float* a = A + (part * 16) * kRow + 8 * ks;
            uint32_t ah[4], al[4];
            split_tf32(a[gid * kRow + tig], ah[0], al[0]);
            split_tf32(a[(gid + 8) * kRow + tig], ah[1], al[1]);
            split_tf32(a[gid * kRow + tig + 4], ah[2], al[2]);
            split_tf32(a[(gid + 8) * kRow + tig + 4], ah[3], al[3]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float* b = B + (8 * nt + gid) * kRow + 8 * ks;
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(b[tig], bh0, bl0);
                split_tf32(b[tig + 4], bh1, bl1);
                mma_tf32(D[nt], al, bh0, bh1);
                mma_tf32(D[nt], ah, bl0, bl1);
                mma_tf32(D[nt], ah, bh0, bh1);
            }
            if (part == 0) {
                mma_tf32(D[NT], al, one, one);
                mma_tf32(D[NT], ah, one, one);
            }
        }
    }
}

// Add the fragments of one layer into the CTA accumulator (reference layout [o][ld], bias last).
template <int NT>
__device__ __forceinline__ void flush_layer(float* acc, int ld, const float (&D)[NT + 1][4], int lane) {
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        atomicAdd(acc + gid * ld + 8 * nt + 2 * tig, D[nt][0]);
        atomicAdd(acc + gid * ld + 8 * nt + 2 * tig + 1, D[nt][1]);
        atomicAdd(acc + (gid + 8) * ld + 8 * nt + 2 * tig, D[nt][2]);
        atomicAdd(acc + (gid + 8) * ld + 8 * nt + 2 * tig + 1, D[nt][3]);
    }
    if (tig == 0) {
        atomicAdd(acc + gid * ld + 8 * NT, D[NT][0]);
        atomicAdd(acc + (gid + 8) * ld + 8 * NT, D[NT][2]);
    }
}

__device__ __forceinline__ void pe_recurrence(float a, float (&cs)[kFreq], float (&sn)[kFreq]) {
    sincosf(a, &sn[0], &cs[0]);
#pragma unroll
    for (int k = 1; k < kFreq; ++k) {
        sn[k] = 2.0f * sn[k - 1] * cs[k - 1];
        cs[k] = (cs[k - 1] - sn[k - 1]) * (cs[k - 1] + sn[k - 1]);
    }
}

__device__ __forceinline__ float dot16(const float* w, const float (&v)[kHid]) {
    float s = 0.0f;
#pragma unroll
    for (int o = 0; o < kHid; ++o) s += w[o] * v[o];
    return s;
}

// One (sample, instance): accumulates weight gradients into `acc` (tensor cores) and returns the pose
// gradients + last-layer weight gradients for the caller's butterfly.
__device__ __forceinline__ void sample_backward(
        const float x[3], const Instance& I, const float* Wt, float scale, float dd, const float dG[3],
        float* ws, int lane, WgradAcc& acc, float (&last)[kHid + 1], float (&pose)[kNumPose]) {
    BoxEval b;
    box_eval(x, I, b);
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = I.R[k] * dG[0] + I.R[3 + k] * dG[1] + I.R[6 + k] * dG[2];
    float pbar[3], vbar[3], dimbar[3];
    {
        float vs = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) vs += v[k] * b.s[k] * b.a[k];
        const float inv_n = 1.0f / b.nrm;
        const float inv_n3 = inv_n * inv_n * inv_n;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float act = b.q[k] > 0.0f ? 1.0f : 0.0f;
            const float hess = act * v[k] * b.s[k] * inv_n - b.a[k] * vs * inv_n3;
            pbar[k] = dd * b.gp[k] + b.s[k] * hess;
            dimbar[k] = -(dd * (b.a[k] * inv_n + b.ind[k]) + hess);
            vbar[k] = b.gp[k];
        }
    }
    float a[3], adot[3], coef[3];
    {
        const float sx[3] = {b.s[0], 1.0f, 1.0f};
        const float m[3] = {fabsf(b.p[0]), b.p[1], b.p[2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a[c] = kPiF * (m[c] / scale);
            coef[c] = sx[c] * (kPiF / scale);
            adot[c] = coef[c] * v[c];
        }
    }
    float* zrow = ws + kRowZ * kRow + lane;
    float* zdrow = ws + kRowZd * kRow + lane;
    float* arow = ws + kRowA * kRow + lane;

    // ------------------------------------------------------------------ phase 1: dual forward
    float lr[4], lm[4];     // LayerNorm 1/sigma and mean(z * centred tangent) per layer
    float out, outd;
    {
        float h[kHid], hd[kHid];
#pragma unroll
        for (int o = 0; o < kHid; ++o) { h[o] = Wt[kEnc * kHid + o]; hd[o] = 0.0f; }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float cs[kFreq], sn[kFreq];
            pe_recurrence(a[c], cs, sn);
#pragma unroll
            for (int k = 0; k < kFreq; ++k) {
                const float da = (float)(1 << k) * adot[c];
                const float de0 = -da * sn[k], de1 = da * cs[k];
                const float* w0 = Wt + (c * 2 * kFreq + 2 * k) * kHid;
                const float* w1 = w0 + kHid;
#pragma unroll
                for (int o = 0; o < kHid; ++o) {
                    h[o] += w0[o] * cs[k] + w1[o] * sn[k];
                    hd[o] += w0[o] * de0 + w1[o] * de1;
                }
            }
        }
#pragma unroll
        for (int l = 1; l <= 4; ++l) {
            float mean = 0.0f, mt = 0.0f;
#pragma unroll
            for (int o = 0; o < kHid; ++o) { mean += h[o]; mt += hd[o]; }
            mean *= (1.0f / kHid); mt *= (1.0f / kHid);
            float var = 0.0f;
#pragma unroll
            for (int o = 0; o < kHid; ++o) { h[o] -= mean; hd[o] -= mt; var += h[o] * h[o]; }
            const float r = 1.0f / sqrtf(var * (1.0f / kHid) + kLnEps);
            float mz = 0.0f;
#pragma unroll
            for (int o = 0; o < kHid; ++o) { h[o] *= r; mz += h[o] * hd[o]; }
            mz *= (1.0f / kHid);
            lr[l - 1] = r; lm[l - 1] = mz;
#pragma unroll
            for (int o = 0; o < kHid; ++o) {
                hd[o] = r * (hd[o] - h[o] * mz);
                zrow[(16 * (l - 1) + o) * kRow] = h[o];
                zdrow[(16 * (l - 1) + o) * kRow] = hd[o];
                float Phi, phi;
                gelu_terms(h[o], Phi, phi);
                hd[o] *= Phi + h[o] * phi;
                h[o] *= Phi;
            }
            if (l < 4) {
                const float* W = Wt + kW1 + (l - 1) * kWStride;
                float hn[kHid], hdn[kHid];
#pragma unroll
                for (int o = 0; o < kHid; ++o) { hn[o] = W[kHid * kHid + o]; hdn[o] = 0.0f; }
#pragma unroll
                for (int i = 0; i < kHid; ++i) {
                    const float* w = W + i * kHid;
#pragma unroll
                    for (int o = 0; o < kHid; ++o) { hn[o] += w[o] * h[i]; hdn[o] += w[o] * hd[i]; }
                }
#pragma unroll
                for (int o = 0; o < kHid; ++o) { h[o] = hn[o]; hd[o] = hdn[o]; }
            } else {
                const float* w = Wt + kW4;
                out = w[kHid]; outd = 0.0f;
#pragma unroll
                for (int i = 0; i < kHid; ++i) { out += w[i] * h[i]; outd += w[i] * hd[i]; }
            }
        }
    }
    const float res = sigmoidf_(out - 1.0f);
    const float sp = res * (1.0f - res);
    const float obar = dd * sp + sp * (1.0f - 2.0f * res) * outd;
    const float odbar = sp;

    // ------------------------------------------------------------------ phase 2: reverse sweep
    // The reverse sweep reads the same staged weights as the forward.  Without a compiler fence the
    // loads are CSE'd across the two phases, 1617 weights stay live and get spilled to local memory;
    // re-reading shared memory is far cheaper.
    float hbar[kHid], hdbar[kHid];
#pragma unroll
    for (int l = 4; l >= 1; --l) {
        asm volatile("" ::: "memory");
        const float* W = (l < 4) ? Wt + kW1 + (l - 1) * kWStride : Wt + kW4;
        if (l < 4) {
#pragma unroll
            for (int o = 0; o < kHid; ++o) { arow[o * kRow] = hbar[o]; arow[(16 + o) * kRow] = hdbar[o]; }
        }
        float z[kHid], zd[kHid], zbar[kHid], zdbar[kHid];
        float s_zb = 0.0f, s_zzb = 0.0f, s_zdb = 0.0f, s_zzdb = 0.0f, s_zdzdb = 0.0f;
#pragma unroll
        for (int i = 0; i < kHid; ++i) {
            z[i] = zrow[(16 * (l - 1) + i) * kRow];
            zd[i] = zdrow[(16 * (l - 1) + i) * kRow];
            float Phi, phi;
            gelu_terms(z[i], Phi, phi);
            const float g = z[i] * Phi;
            const float g1 = Phi + z[i] * phi;
            const float g2 = phi * (2.0f - z[i] * z[i]);
            const float gd = g1 * zd[i];
            float gbar, gdbar;
            if (l < 4) {
                zrow[(16 * (l - 1) + i) * kRow] = g;      // B operand of this layer's weight gradient
                zdrow[(16 * (l - 1) + i) * kRow] = gd;
                gbar = dot16(W + i * kHid, hbar);
                gdbar = dot16(W + i * kHid, hdbar);
            } else {
                last[i] = obar * g + odbar * gd;
                gbar = W[i] * obar;
                gdbar = W[i] * odbar;
            }
            zbar[i] = gbar * g1 + gdbar * g2 * zd[i];
            zdbar[i] = gdbar * g1;
            s_zb += zbar[i]; s_zzb += z[i] * zbar[i];
            s_zdb += zdbar[i]; s_zzdb += z[i] * zdbar[i]; s_zdzdb += zd[i] * zdbar[i];
        }
        if (l == 4) last[kHid] = obar;
        const float inv = 1.0f / kHid;
        s_zb *= inv; s_zzb *= inv; s_zdb *= inv; s_zzdb *= inv; s_zdzdb *= inv;
        const float r = lr[l - 1], m = lm[l - 1];
#pragma unroll
        for (int o = 0; o < kHid; ++o) {
            hdbar[o] = r * (zdbar[o] - s_zdb - z[o] * s_zzdb);
            hbar[o] = r * (zbar[o] - s_zb - z[o] * s_zzb) - r * s_zdzdb * z[o] - r * m * hdbar[o] - r * s_zzdb * zd[o];
        }
        if (l < 4) {
            __syncwarp();
            wgrad_mma<2>(ws + kRowA * kRow, ws + (kRowZ + 16 * (l - 1)) * kRow, ws + (kRowZd + 16 * (l - 1)) * kRow,
                         acc.h[l - 1], lane);
            __syncwarp();
        }
    }
    // ---- layer 0 + positional encoding
    float abar[3], adbar[3];
    {
        asm volatile("" ::: "memory");
#pragma unroll
        for (int o = 0; o < kHid; ++o) { arow[o * kRow] = hbar[o]; arow[(16 + o) * kRow] = hdbar[o]; }
        float* erow = ws + lane;                 // e rows [0,48), ed rows [48,96): over the consumed stash
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float cs[kFreq], sn[kFreq];
            pe_recurrence(a[c], cs, sn);
            float ab = 0.0f, adb = 0.0f;
#pragma unroll
            for (int k = 0; k < kFreq; ++k) {
                const float f = (float)(1 << k);
                const int j = c * 2 * kFreq + 2 * k;
                const float da = f * adot[c];
                erow[j * kRow] = cs[k];
                erow[(j + 1) * kRow] = sn[k];
                erow[(kEnc + j) * kRow] = -da * sn[k];
                erow[(kEnc + j + 1) * kRow] = da * cs[k];
                const float* w0 = Wt + j * kHid;
                const float eb0 = dot16(w0, hbar), eb1 = dot16(w0 + kHid, hbar);
                const float edb0 = dot16(w0, hdbar), edb1 = dot16(w0 + kHid, hdbar);
                ab += f * (-eb0 * sn[k] + eb1 * cs[k] - da * (edb0 * cs[k] + edb1 * sn[k]));
                adb += f * (-edb0 * sn[k] + edb1 * cs[k]);
            }
            abar[c] = ab;
            adbar[c] = adb;
        }
        __syncwarp();
        wgrad_mma<6>(ws + kRowA * kRow, ws, ws + kEnc * kRow, acc.l0, lane);
        __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        pbar[c] += abar[c] * coef[c];
        vbar[c] += adbar[c] * coef[c];
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        pose[m] = -(I.R[3 * m] * pbar[0] + I.R[3 * m + 1] * pbar[1] + I.R[3 * m + 2] * pbar[2]);
        pose[3 + m] = dimbar[m];
#pragma unroll
        for (int k = 0; k < 3; ++k) pose[6 + 3 * m + k] = b.y[m] * pbar[k] + dG[m] * vbar[k];
    }
}

}  // namespace bwd2

__global__ void __launch_bounds__(kThreads, 2) field_backward_mlp_kernel(
        SceneDev scene, RaysDev rays, const float4* __restrict__ adjoint, float* __restrict__ partials) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                                   // [kNumW] staged (transposed) weights
    float* sAcc = smem + kNumW + 3;                     // [kGradStride] CTA accumulator (16B aligned: 1620)
    float* sRows = sAcc + kGradStride;                  // [kWarps][160][36]
    const int inst = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    stage_weights(scene.W + (size_t)inst * kNumW, sW);
    for (int f = threadIdx.x; f < kGradStride; f += kThreads) sAcc[f] = 0.0f;
    __syncthreads();

    Instance I;
    load_instance(scene, inst, I);
    float* ws = sRows + (size_t)warp * bwd2::kWarpFloats;
    bwd2::WgradAcc acc;
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc.h[l][t][q] = 0.0f;
#pragma unroll
    for (int t = 0; t < 7; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc.l0[t][q] = 0.0f;

    const size_t total = (size_t)rays.R * rays.M;
    const size_t num_tiles = (total + kThreads - 1) / kThreads;
    for (size_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // The staged weights are loop-invariant; without this barrier the compiler hoists all 1617
        // shared-memory loads out of the tile loop and spills them to local memory (7 KB/thread).
        asm volatile("" ::: "memory");
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
        float last[kHid + 1], pose[kNumPose];
        bwd2::sample_backward(x, I, sW, scene.scale, dd, dG, ws, lane, acc, last, pose);
        // 17 last-layer + 15 pose values = exactly one 32-wide butterfly
        float v[32];
#pragma unroll
        for (int k = 0; k < kHid + 1; ++k) v[k] = last[k];
#pragma unroll
        for (int k = 0; k < kNumPose; ++k) v[kHid + 1 + k] = pose[k];
        WarpSink sink{nullptr, lane};
        sink.butterfly_step<16>(v);
        sink.butterfly_step<8>(v);
        sink.butterfly_step<4>(v);
        sink.butterfly_step<2>(v);
        sink.butterfly_step<1>(v);
        atomicAdd(sAcc + kW4 + lane, v[0]);                   // kW4 + 17 == kNumW: last layer then pose, contiguous
    }
    // ---- flush the persistent tensor-core accumulators
#pragma unroll
    for (int l = 0; l < 3; ++l) bwd2::flush_layer<2>(sAcc + kW1 + l * kWStride, kHid + 1, acc.h[l], lane);
    bwd2::flush_layer<6>(sAcc + kW0, kEnc + 1, acc.l0, lane);
    __syncthreads();
    float* out = partials + ((size_t)inst * gridDim.x + blockIdx.x) * kGradStride;
    for (int f = threadIdx.x; f < kGradStride; f += kThreads) out[f] = sAcc[f];
}

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
static int g_bwd_blocks_per_sm[2] = {0, 0};

// 0 = tensor-core kernel (default), 1 = SIMT kernel kept as an independent cross-check
// (VSRD_FIELD_IMPL=simt, read per call so a test can flip it inside one process)
static int backward_impl() {
    const char* impl = getenv("VSRD_FIELD_IMPL");
    return (impl && strcmp(impl, "simt") == 0) ? 1 : 0;
}

static int device_setup() {
    if (g_num_sms) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("vsrd_b200: no CUDA device%s");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail("vsrd_b200: cudaGetDeviceProperties failed%s");
    if (cudaFuncSetAttribute(field_backward_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bwd2::kSmemBytes) != cudaSuccess)
        return fail("vsrd_b200: cannot reserve %s of shared memory for field_backward_mlp_kernel", "105 KB");
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_bwd_blocks_per_sm[0], field_backward_box_kernel, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_bwd_blocks_per_sm[1], field_backward_mlp_kernel, kThreads,
                                                  bwd2::kSmemBytes);
    if (cudaGetLastError() != cudaSuccess || g_bwd_blocks_per_sm[0] < 1 || g_bwd_blocks_per_sm[1] < 1)
        return fail("vsrd_b200: kernels not loadable on this device (built for sm_100a)%s");
    g_num_sms = prop.multiProcessorCount;
    return 0;
}

static int backward_blocks(int N, int R, int M, bool residual) {
    const long long tiles = ((long long)R * M + kThreads - 1) / kThreads;
    long long resident = (long long)g_num_sms * g_bwd_blocks_per_sm[residual ? 1 : 0];
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

int vsrd_backward_blocks_per_instance(int num_instances, int num_rays, int num_intervals) {
    if (device_setup()) return -1;
    // the residual kernel has the lower occupancy; size for the larger grid so one buffer fits both
    const int a = backward_blocks(num_instances, num_rays, num_intervals, false);
    const int b = backward_blocks(num_instances, num_rays, num_intervals, true);
    // the tensor-core kernel writes one row per (CTA, instance) segment: at most #SMs + N rows
    const int rows = backward_mma_partial_rows(num_instances);
    if (rows < 0) return -1;
    const int n = num_instances > 0 ? num_instances : 1;
    const int c = (rows + n - 1) / n;
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
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
    if (total > 0 && s.W && backward_impl() == 0)
        return launch_field_backward_mma(s, r, adjoint, partials, grad_locations, grad_rotations, grad_half_extents,
                                         grad_mlp_weights, st);
    int G = 1;
    if (total == 0) {
        cudaMemsetAsync(partials, 0, (size_t)s.N * kGradStride * sizeof(float), st);
    } else {
        G = backward_blocks(s.N, r.R, r.M, s.W != nullptr);
        const dim3 grid((unsigned)G, (unsigned)s.N);
        if (s.W) field_backward_mlp_kernel<<<grid, kThreads, bwd2::kSmemBytes, st>>>(s, r, (const float4*)adjoint, partials);
        else field_backward_box_kernel<<<grid, kThreads, 0, st>>>(s, r, (const float4*)adjoint, partials);
        VSRD_CHECK_LAUNCH();
    }
    const dim3 rgrid((kGradStride + 127) / 128, (unsigned)s.N);
    reduce_partials_kernel<<<rgrid, 128, 0, st>>>(partials, G, grad_locations, grad_rotations, grad_half_extents,
                                                 s.W ? grad_mlp_weights : nullptr);
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
