// sm_100a kernels of the VSRD silhouette-renderer hot path, part 2/3: per-(sample, instance) field
// forward (box SDF + residual MLP, value and spatial gradient).
//
//   field_forward_mma_kernel   residual instances: warp tiles of 32 samples, contractions on the tensor
//                              cores (3xTF32 mma.sync, activations chained in registers), value by a
//                              forward sweep and d/dx by one reverse sweep (2F contraction flops per
//                              sample, the count SURVEY.md 8d credits)
//   field_forward_kernel<..>   box-only instances (warm-up steps) and the v3 SIMT residual path kept as
//                              an independent cross-check (VSRD_FIELD_IMPL=simt)
#include "vsrd_frag.cuh"

namespace vsrd {

// =============================================================================================
// v3 SIMT kernel: one thread per (sample, instance), forward-mode tangents.  grid = (ceil(R*M/128), N)
// =============================================================================================
template <bool kResidual>
__global__ void __launch_bounds__(kThreads) field_forward_kernel(SceneDev scene, RaysDev rays, float4* __restrict__ field) {
    __shared__ __align__(16) float sW[kResidual ? kNumW : 4];
    __shared__ float sAct[kResidual ? kThreads * kFwdActRows : 1];   // [warp][row][lane]: lane-private, conflict-free
    const int inst = blockIdx.y;
    if (kResidual) {
        stage_weights(scene.W + (size_t)inst * kNumW, sW);
        __syncthreads();
    }
    const size_t total = (size_t)rays.R * rays.M;
    const size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= total) return;
    const int r = (int)(idx / rays.M);
    const int j = (int)(idx - (size_t)r * rays.M);
    Instance I;
    load_instance(scene, inst, I);
    float x[3], d, G[3];
    sample_position(rays, r, j, x);
    float* act = sAct + (kResidual ? (threadIdx.x >> 5) * 32 * kFwdActRows + (threadIdx.x & 31) : 0);
    field_forward_looped<kResidual>(x, I, sW, scene.scale, act, 32, d, G);
    field[(size_t)inst * total + idx] = make_float4(d, G[0], G[1], G[2]);
}

// =============================================================================================
// v4 tensor-core kernel.  Persistent: gridDim.x CTAs of kFwdWarps warps split the N * tiles_per_inst
// warp tiles evenly; a CTA restages the weight fragments when its range crosses an instance boundary.
//
// Per warp tile (see vsrd_frag.cuh for the register layout):
//   1. lane == sample: sample position, box SDF, PE argument a_c = fl(pi * u_c)
//   2. fragment layout: positional encoding (each lane only the 12 (cos, sin) pairs per row it feeds
//      into the MMA), layer 0, then 4 x [LayerNorm -> GELU -> linear]; LayerNorm statistics are quad
//      reductions (the 16 channels of a row sit in the 4 lanes of a quad)
//   3. reverse sweep for d out / d a: the per-layer (z, gelu'(z) / sigma) pairs come back from a
//      lane-private shared-memory stash, contractions use the transposed weight fragments
//   4. lane == sample: residual = sigmoid(out - 1), chain rule through |p_x|, rotate the gradient to the
//      world frame, one coalesced float4 store
// =============================================================================================
constexpr int kFwdWarps = 12;
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kFwdStashFloats = 4 * 32 * 32;     // per warp: [layer][z 16 | g1/sigma 16][lane]
constexpr size_t kFwdSmemBytes = frag::kWeightBytes + (size_t)kFwdWarps * kFwdStashFloats * sizeof(float);

__global__ void __launch_bounds__(kFwdThreads, 1) field_forward_mma_kernel(
        SceneDev scene, RaysDev rays, float4* __restrict__ field, int tiles_per_inst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sF = reinterpret_cast<float4*>(smem_raw);
    float* sTail = reinterpret_cast<float*>(sF + frag::kFragFloat4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3;
    float* stash = sTail + frag::kTailFloats + (size_t)warp * kFwdStashFloats + lane;
    const float4* fragL = sF + lane;

    const int total = rays.R * rays.M;
    const long long all_tiles = (long long)scene.N * tiles_per_inst;
    const long long begin = all_tiles * blockIdx.x / gridDim.x;
    const long long end = all_tiles * (blockIdx.x + 1) / gridDim.x;

    for (long long seg = begin; seg < end;) {
        const int inst = (int)(seg / tiles_per_inst);
        const long long seg_end = min(end, (long long)(inst + 1) * tiles_per_inst);
        __syncthreads();                                   // previous instance's tiles are done
        frag::stage_weight_fragments(scene.W + (size_t)inst * kNumW, sF, sTail);
        __syncthreads();
        Instance I;
        load_instance(scene, inst, I);
        const float pi_scale = kPiF / scene.scale;

#pragma unroll 1
        for (long long tile = seg + warp; tile < seg_end; tile += kFwdWarps) {
            const int base = (int)(tile - (long long)inst * tiles_per_inst) * 32;
            // ------------------------------------------------------------ 1. lane == sample
            const int idx = min(base + lane, total - 1);
            const int r = idx / rays.M;
            const int j = idx - r * rays.M;
            float x[3];
            sample_position(rays, r, j, x);
            BoxEval b;
            box_eval(x, I, b);
            float arow[4][3];
            {
                const float m[3] = {fabsf(b.p[0]), b.p[1], b.p[2]};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float v[4];
                    frag::lanes_to_rows(kPiF * (m[c] / scene.scale), lane, v);
#pragma unroll
                    for (int s = 0; s < 4; ++s) arow[s][c] = v[s];
                }
            }
            // ------------------------------------------------------------ 2. forward sweep
            frag::Encoding e;
            frag::encode(arow, t, e);
            float h[2][2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const float b0 = sTail[8 * nt + 2 * t], b1 = sTail[8 * nt + 2 * t + 1];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) { h[mt][nt][0] = b0; h[mt][nt][1] = b1; h[mt][nt][2] = b0; h[mt][nt][3] = b1; }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    const int ks = 2 * c + f;
                    const float4 w0 = fragL[(frag::kF0 + 2 * ks) * 32], w1 = fragL[(frag::kF0 + 2 * ks + 1) * 32];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        uint32_t ah[4], al[4];
                        frag::split(e.cs[2 * mt][c][f], ah[0], al[0]);
                        frag::split(e.cs[2 * mt + 1][c][f], ah[1], al[1]);
                        frag::split(e.sn[2 * mt][c][f], ah[2], al[2]);
                        frag::split(e.sn[2 * mt + 1][c][f], ah[3], al[3]);
                        frag::mma3(h[mt][0], ah, al, w0);
                        frag::mma3(h[mt][1], ah, al, w1);
                    }
                }
            float out[4];
#pragma unroll 1
            for (int l = 1; l <= 4; ++l) {
                float* st = stash + (l - 1) * 32 * 32;
                // LayerNorm (no affine, eps 1e-5) + GELU per row slot
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int mt = s >> 1, q = 2 * (s & 1);
                    float v0 = h[mt][0][q], v1 = h[mt][0][q + 1], v2 = h[mt][1][q], v3 = h[mt][1][q + 1];
                    const float mean = frag::quad_sum((v0 + v1) + (v2 + v3)) * (1.0f / kHid);
                    v0 -= mean; v1 -= mean; v2 -= mean; v3 -= mean;
                    const float var = frag::quad_sum(fmaf(v0, v0, v1 * v1) + fmaf(v2, v2, v3 * v3)) * (1.0f / kHid);
                    const float rs = rsqrtf(var + kLnEps);
                    float z[4] = {v0 * rs, v1 * rs, v2 * rs, v3 * rs};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float Phi, phi;
                        frag::gelu_terms_fast(z[k], Phi, phi);
                        st[(4 * s + k) * 32] = z[k];
                        st[(16 + 4 * s + k) * 32] = fmaf(z[k], phi, Phi) * rs;      // gelu'(z) / sigma
                        z[k] *= Phi;
                    }
                    h[mt][0][q] = z[0]; h[mt][0][q + 1] = z[1]; h[mt][1][q] = z[2]; h[mt][1][q + 1] = z[3];
                }
                if (l < 4) {
                    float hn[2][2][4];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const float b0 = sTail[16 * l + 8 * nt + 2 * t], b1 = sTail[16 * l + 8 * nt + 2 * t + 1];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) { hn[mt][nt][0] = b0; hn[mt][nt][1] = b1; hn[mt][nt][2] = b0; hn[mt][nt][3] = b1; }
                    }
                    const float4* fl = fragL + (frag::kF1 + 4 * (l - 1)) * 32;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const float4 w0 = fl[(2 * ks) * 32], w1 = fl[(2 * ks + 1) * 32];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            uint32_t ah[4], al[4];
                            frag::a_from_c(h[mt][ks], ah, al);
                            frag::mma3(hn[mt][0], ah, al, w0);
                            frag::mma3(hn[mt][1], ah, al, w1);
                        }
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int q = 0; q < 4; ++q) h[mt][nt][q] = hn[mt][nt][q];
                } else {
                    const float w00 = sTail[frag::kTailW4 + 2 * t], w01 = sTail[frag::kTailW4 + 2 * t + 1];
                    const float w10 = sTail[frag::kTailW4 + 8 + 2 * t], w11 = sTail[frag::kTailW4 + 8 + 2 * t + 1];
                    const float b4 = sTail[frag::kTailB4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int mt = s >> 1, q = 2 * (s & 1);
                        out[s] = frag::quad_sum(fmaf(w00, h[mt][0][q], w01 * h[mt][0][q + 1])
                                                + fmaf(w10, h[mt][1][q], w11 * h[mt][1][q + 1])) + b4;
                    }
                }
            }
            // ------------------------------------------------------------ 3. reverse sweep: d out / d a
            // gbar: adjoint of the GELU outputs of layer l (C layout); starts as the last layer's weights
            float gb[2][2][4];
            {
                const float w00 = sTail[frag::kTailW4 + 2 * t], w01 = sTail[frag::kTailW4 + 2 * t + 1];
                const float w10 = sTail[frag::kTailW4 + 8 + 2 * t], w11 = sTail[frag::kTailW4 + 8 + 2 * t + 1];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    gb[mt][0][0] = w00; gb[mt][0][1] = w01; gb[mt][0][2] = w00; gb[mt][0][3] = w01;
                    gb[mt][1][0] = w10; gb[mt][1][1] = w11; gb[mt][1][2] = w10; gb[mt][1][3] = w11;
                }
            }
#pragma unroll 1
            for (int l = 4; l >= 1; --l) {
                const float* st = stash + (l - 1) * 32 * 32;
                // hbar = zb - mean(zb) - z mean(z zb),  zb = gbar * gelu'(z) / sigma
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int mt = s >> 1, q = 2 * (s & 1);
                    float z[4], zb[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { z[k] = st[(4 * s + k) * 32]; zb[k] = st[(16 + 4 * s + k) * 32]; }
                    zb[0] *= gb[mt][0][q]; zb[1] *= gb[mt][0][q + 1]; zb[2] *= gb[mt][1][q]; zb[3] *= gb[mt][1][q + 1];
                    const float m1 = frag::quad_sum((zb[0] + zb[1]) + (zb[2] + zb[3])) * (1.0f / kHid);
                    const float m2 = frag::quad_sum(fmaf(z[0], zb[0], z[1] * zb[1]) + fmaf(z[2], zb[2], z[3] * zb[3])) * (1.0f / kHid);
                    gb[mt][0][q] = zb[0] - m1 - z[0] * m2;
                    gb[mt][0][q + 1] = zb[1] - m1 - z[1] * m2;
                    gb[mt][1][q] = zb[2] - m1 - z[2] * m2;
                    gb[mt][1][q + 1] = zb[3] - m1 - z[3] * m2;
                }
                if (l > 1) {        // gbar_{l-1} = W_{l-1}^T hbar  (hidden layer l-1 maps gelu(z_{l-1}) to h_l)
                    float gn[2][2][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int q = 0; q < 4; ++q) gn[mt][nt][q] = 0.0f;
                    const float4* fl = fragL + (frag::kR1 + 4 * (l - 2)) * 32;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const float4 w0 = fl[(2 * ks) * 32], w1 = fl[(2 * ks + 1) * 32];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            uint32_t ah[4], al[4];
                            frag::a_from_c(gb[mt][ks], ah, al);
                            frag::mma3(gn[mt][0], ah, al, w0);
                            frag::mma3(gn[mt][1], ah, al, w1);
                        }
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int q = 0; q < 4; ++q) gb[mt][nt][q] = gn[mt][nt][q];
                }
            }
            // layer 0 transposed + positional-encoding adjoint: abar_c = sum_k 2^k (ebar_sin cos - ebar_cos sin)
            float abar[4][3];
            {
                uint32_t ah[2][2][4], al[2][2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) frag::a_from_c(gb[mt][ks], ah[mt][ks], al[mt][ks]);
                const float f0 = (float)(1 << t), f1 = 16.0f * f0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        const int nt = 2 * c + f;
                        const float4 w0 = fragL[(frag::kR0 + nt) * 32], w1 = fragL[(frag::kR0 + 6 + nt) * 32];
                        const float fk = f ? f1 : f0;
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            float eb[4] = {0.0f, 0.0f, 0.0f, 0.0f};      // (cos, sin) adjoints of rows g, g+8
                            frag::mma3(eb, ah[mt][0], al[mt][0], w0);
                            frag::mma3(eb, ah[mt][1], al[mt][1], w1);
                            acc[2 * mt] += fk * (eb[1] * e.cs[2 * mt][c][f] - eb[0] * e.sn[2 * mt][c][f]);
                            acc[2 * mt + 1] += fk * (eb[3] * e.cs[2 * mt + 1][c][f] - eb[2] * e.sn[2 * mt + 1][c][f]);
                        }
                    }
#pragma unroll
                    for (int s = 0; s < 4; ++s) abar[s][c] = frag::quad_sum(acc[s]);
                }
            }
            // ------------------------------------------------------------ 4. lane == sample
            const float o = frag::rows_to_lanes(out, lane);
            float ga[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v[4] = {abar[0][c], abar[1][c], abar[2][c], abar[3][c]};
                ga[c] = frag::rows_to_lanes(v, lane);
            }
            const float res = sigmoidf_(o - 1.0f);
            const float sp = res * (1.0f - res) * pi_scale;
            const float gp0 = b.gp[0] + sp * b.s[0] * ga[0];
            const float gp1 = b.gp[1] + sp * ga[1];
            const float gp2 = b.gp[2] + sp * ga[2];
            if (base + lane < total) {
                field[(size_t)inst * total + base + lane] = make_float4(
                    b.value + res,
                    I.R[0] * gp0 + I.R[1] * gp1 + I.R[2] * gp2,
                    I.R[3] * gp0 + I.R[4] * gp1 + I.R[5] * gp2,
                    I.R[6] * gp0 + I.R[7] * gp1 + I.R[8] * gp2);
            }
            __syncwarp();
        }
        seg = seg_end;
    }
}

static int g_fwd_sms = 0;

// 0 = tensor-core kernel (default), 1 = SIMT cross-check (VSRD_FIELD_IMPL=simt, read per call)
static int forward_impl() {
    const char* impl = getenv("VSRD_FIELD_IMPL");
    return (impl && strcmp(impl, "simt") == 0) ? 1 : 0;
}

static int forward_setup() {
    if (g_fwd_sms) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("vsrd_b200: no CUDA device%s");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail("vsrd_b200: cudaGetDeviceProperties failed%s");
    if (cudaFuncSetAttribute(field_forward_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kFwdSmemBytes) != cudaSuccess)
        return fail("vsrd_b200: cannot reserve %s of shared memory for field_forward_mma_kernel (built for sm_100a)", "217 KB");
    g_fwd_sms = prop.multiProcessorCount;
    return 0;
}

}  // namespace vsrd

using namespace vsrd;

extern "C" {

int vsrd_field_forward(const VsrdScene* scene, const VsrdRays* rays, float* field, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    VSRD_CHECK_ARG(field != nullptr, "field is NULL");
    const size_t total = (size_t)r.R * r.M;
    if (total == 0) return 0;
    VSRD_CHECK_ARG(total < (size_t)1 << 31, "R*M must be < 2^31");
    if (forward_setup()) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (s.W && forward_impl() == 0) {
        const int tiles_per_inst = (int)((total + 31) / 32);
        const long long all_tiles = (long long)s.N * tiles_per_inst;
        const long long want = (all_tiles + kFwdWarps - 1) / kFwdWarps;
        const int grid = (int)(want < g_fwd_sms ? want : g_fwd_sms);
        field_forward_mma_kernel<<<grid, kFwdThreads, kFwdSmemBytes, st>>>(s, r, (float4*)field, tiles_per_inst);
    } else {
        const dim3 grid((unsigned)((total + kThreads - 1) / kThreads), (unsigned)s.N);
        if (s.W) field_forward_kernel<true><<<grid, kThreads, 0, st>>>(s, r, (float4*)field);
        else field_forward_kernel<false><<<grid, kThreads, 0, st>>>(s, r, (float4*)field);
    }
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
