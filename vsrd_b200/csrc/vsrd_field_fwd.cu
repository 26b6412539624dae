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
// MT = m-tiles (16 rows) per warp tile; one CTA per SM.  Measured on B200 (profiles/r01_v6_*): 12 warps x 2
// m-tiles (168 registers) 0.321 ms, 16 warps x 1 m-tile (128 registers) 0.335 ms, 20 x 1 (96 registers,
// spills) 0.342 ms for the fine pass -- throughput follows the number of m-tiles in flight, which the
// register file bounds, not the warp count.  MT = 2 ships; MT = 1 stays selectable (VSRD_FWD_MT=1).
template <int MT>
struct FwdCfg {
    static constexpr int kWarps = MT == 2 ? 12 : 16;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kRows = 16 * MT;
    static constexpr int kStashPairs = 4 * 8 * MT * 32;     // per warp, float2: [layer][z pairs 4 MT | g1/sigma pairs 4 MT][lane]
    static constexpr size_t kSmemBytes = frag::kWeightBytes + (size_t)kWarps * kStashPairs * sizeof(float2);
};

template <int MT>
__global__ void __launch_bounds__(FwdCfg<MT>::kThreads, 1) field_forward_mma_kernel(
        SceneDev scene, RaysDev rays, float4* __restrict__ field, int tiles_per_inst) {
    using namespace frag;
    using Cfg = FwdCfg<MT>;
    constexpr int kRows = Cfg::kRows, kSlots = 2 * MT, kLayerPairs = 4 * MT;   // pair rows of z (then as many of g1/sigma) per layer
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_next_tile;
    float4* sF = reinterpret_cast<float4*>(smem_raw);
    float* sTail = reinterpret_cast<float*>(sF + kFragFloat4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3;
    float2* stash = reinterpret_cast<float2*>(sTail + kTailFloats) + (size_t)warp * Cfg::kStashPairs + lane;
    const float4* fragL = sF + lane;

    const int total = rays.R * rays.M;
    const long long all_tiles = (long long)scene.N * tiles_per_inst;
    const long long begin = all_tiles * blockIdx.x / gridDim.x;
    const long long end = all_tiles * (blockIdx.x + 1) / gridDim.x;
    const float cull_margin = kCullLogEps * scene_temperature(scene);
    unsigned tiles_visited = 0, tiles_culled = 0;          // per warp; two atomics per warp at the end

    for (long long seg = begin; seg < end;) {
        const int inst = (int)(seg / tiles_per_inst);
        const long long seg_end = min(end, (long long)(inst + 1) * tiles_per_inst);
        __syncthreads();                                   // previous instance's tiles are done
        stage_weight_fragments(scene.W + (size_t)inst * kNumW, sF, sTail);
        if (threadIdx.x == 0) s_next_tile = 0;
        __syncthreads();
        Instance I;
        load_instance(scene, inst, I);
        const float pi_scale = kPiF / scene.scale;
        const f2 w4p0 = make_float2(sTail[kTailW4 + 2 * t], sTail[kTailW4 + 2 * t + 1]);
        const f2 w4p1 = make_float2(sTail[kTailW4 + 8 + 2 * t], sTail[kTailW4 + 8 + 2 * t + 1]);
        const float b4 = sTail[kTailB4];

        // Warps take tiles from a CTA-wide counter (outputs are per sample, so the assignment is free to vary): with
        // instance culling the cost of a tile is bimodal, and a static stride leaves the slowest warp near the worst case.
#pragma unroll 1
        while (true) {
            int claimed = 0;
            if (lane == 0) claimed = atomicAdd(&s_next_tile, 1);
            const long long tile = seg + __shfl_sync(kFull, claimed, 0);
            if (tile >= seg_end) break;
            const int base = (int)(tile - (long long)inst * tiles_per_inst) * kRows;
            // ------------------------------------------------------------ 1. lane == sample (lanes < kRows)
            const int row = lane & (kRows - 1);
            const int idx = min(base + row, total - 1);
            const int r = idx / rays.M;
            const int j = idx - r * rays.M;
            float x[3];
            sample_position(rays, r, j, x);
            BoxEval b;
            box_eval(x, I, b);
            if (rays.bound != nullptr) {
                // instance culling (VsrdRays::union_bound): every sample of the tile is farther from this instance's box
                // than the nearest box + the residual's range + 30 T -> soft-min weight < 1e-13: box value suffices
                const bool in_range = lane < kRows && base + lane < total;
                const bool far = !in_range || b.value - (__ldg(rays.bound + idx) + 1.0f) > cull_margin;
                ++tiles_visited;
                if (__all_sync(kFull, far)) {
                    if (in_range)
                        field[(size_t)inst * total + base + lane] = make_float4(
                            b.value,
                            I.R[0] * b.gp[0] + I.R[1] * b.gp[1] + I.R[2] * b.gp[2],
                            I.R[3] * b.gp[0] + I.R[4] * b.gp[1] + I.R[5] * b.gp[2],
                            I.R[6] * b.gp[0] + I.R[7] * b.gp[1] + I.R[8] * b.gp[2]);
                    ++tiles_culled;
                    continue;
                }
            }
            f2 arow[MT][3];                                // PE arguments of rows (g, g + 8) of each m-tile
            {
                const float m[3] = {fabsf(b.p[0]), b.p[1], b.p[2]};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    f2 v[MT];
                    lanes_to_row_pairs<MT>(kPiF * (m[c] / scene.scale), lane, v);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) arow[mt][c] = v[mt];
                }
            }
            // ------------------------------------------------------------ 2. forward sweep
            EncodingT<MT> e;
            encode2<MT>(arow, t, e);
            f2 h[MT][2][2];                                // [m-tile][n-tile][row g | row g + 8]
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const f2 bias = make_float2(sTail[8 * nt + 2 * t], sTail[8 * nt + 2 * t + 1]);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) { h[mt][nt][0] = bias; h[mt][nt][1] = bias; }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    const int ks = 2 * c + f;
                    const float4 w0 = fragL[(kF0 + 2 * ks) * 32], w1 = fragL[(kF0 + 2 * ks + 1) * 32];
                    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) a_from_row_pairs(e.cs[mt][c][f], e.sn[mt][c][f], ah[mt], al[mt]);
                    mma3_step<false, MT>(h, ah, al, w0, w1);
                }
            float out[kSlots];
#pragma unroll 1
            for (int l = 1; l <= 4; ++l) {
                float2* st = stash + (l - 1) * 2 * kLayerPairs * 32;
                // LayerNorm (no affine, eps 1e-5) + GELU per row slot (slot s = 2 mt + half)
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    f2& p0 = h[s >> 1][0][s & 1];
                    f2& p1 = h[s >> 1][1][s & 1];
                    const float mean = quad_sum(hsum(add2(p0, p1))) * (1.0f / kHid);
                    p0 = add2(p0, bc(-mean)); p1 = add2(p1, bc(-mean));
                    const float var = quad_sum(hsum(fma2(p0, p0, mul2(p1, p1)))) * (1.0f / kHid);
                    const float rs = rsqrtf(var + kLnEps);
                    p0 = mul2(p0, bc(rs)); p1 = mul2(p1, bc(rs));
                    f2 Phi0, phi0, Phi1, phi1, zz;
                    gelu_terms2(p0, Phi0, phi0, zz);
                    gelu_terms2(p1, Phi1, phi1, zz);
                    st[(2 * s) * 32] = p0;
                    st[(2 * s + 1) * 32] = p1;
                    st[(kLayerPairs + 2 * s) * 32] = mul2(fma2(p0, phi0, Phi0), bc(rs));      // gelu'(z) / sigma
                    st[(kLayerPairs + 2 * s + 1) * 32] = mul2(fma2(p1, phi1, Phi1), bc(rs));
                    p0 = mul2(p0, Phi0); p1 = mul2(p1, Phi1);
                }
                if (l < 4) {
                    f2 hn[MT][2][2];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const f2 bias = make_float2(sTail[16 * l + 8 * nt + 2 * t], sTail[16 * l + 8 * nt + 2 * t + 1]);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) { hn[mt][nt][0] = bias; hn[mt][nt][1] = bias; }
                    }
                    const float4* fl = fragL + (kF1 + 4 * (l - 1)) * 32;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const float4 w0 = fl[(2 * ks) * 32], w1 = fl[(2 * ks + 1) * 32];
                        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) a_from_c(h[mt][ks], ah[mt], al[mt]);
                        mma3_step<false, MT>(hn, ah, al, w0, w1);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) { h[mt][nt][0] = hn[mt][nt][0]; h[mt][nt][1] = hn[mt][nt][1]; }
                } else {
#pragma unroll
                    for (int s = 0; s < kSlots; ++s)
                        out[s] = quad_sum(hsum(fma2(w4p0, h[s >> 1][0][s & 1], mul2(w4p1, h[s >> 1][1][s & 1])))) + b4;
                }
            }
            // ------------------------------------------------------------ 3. reverse sweep: d out / d a
            // gb: adjoint of the GELU outputs of layer l (C layout); starts as the last layer's weights
            f2 gb[MT][2][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                gb[mt][0][0] = w4p0; gb[mt][0][1] = w4p0;
                gb[mt][1][0] = w4p1; gb[mt][1][1] = w4p1;
            }
#pragma unroll 1
            for (int l = 4; l >= 1; --l) {
                const float2* st = stash + (l - 1) * 2 * kLayerPairs * 32;
                // hbar = zb - mean(zb) - z mean(z zb),  zb = gbar * gelu'(z) / sigma
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    f2& g0 = gb[s >> 1][0][s & 1];
                    f2& g1 = gb[s >> 1][1][s & 1];
                    const f2 z0 = st[(2 * s) * 32], z1 = st[(2 * s + 1) * 32];
                    const f2 zb0 = mul2(st[(kLayerPairs + 2 * s) * 32], g0), zb1 = mul2(st[(kLayerPairs + 2 * s + 1) * 32], g1);
                    const f2 m = quad_sum2(make_float2(hsum(add2(zb0, zb1)), hsum(fma2(z0, zb0, mul2(z1, zb1)))));
                    const float m1 = m.x * (-1.0f / kHid), m2 = m.y * (-1.0f / kHid);
                    g0 = fma2(z0, bc(m2), add2(zb0, bc(m1)));
                    g1 = fma2(z1, bc(m2), add2(zb1, bc(m1)));
                }
                if (l > 1) {        // gbar_{l-1} = W_{l-1}^T hbar  (hidden layer l-1 maps gelu(z_{l-1}) to h_l)
                    f2 gn[MT][2][2];
                    const float4* fl = fragL + (kR1 + 4 * (l - 2)) * 32;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const float4 w0 = fl[(2 * ks) * 32], w1 = fl[(2 * ks + 1) * 32];
                        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) a_from_c(gb[mt][ks], ah[mt], al[mt]);
                        if (ks == 0) mma3_step<true, MT>(gn, ah, al, w0, w1);
                        else mma3_step<false, MT>(gn, ah, al, w0, w1);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) { gb[mt][nt][0] = gn[mt][nt][0]; gb[mt][nt][1] = gn[mt][nt][1]; }
                }
            }
            // layer 0 transposed + positional-encoding adjoint: abar_c = sum_k 2^k (ebar_sin cos - ebar_cos sin)
            float abar[kSlots][3];
            {
                const float f0 = (float)(1 << t), f1 = 16.0f * f0;
                uint32_t ah[2][MT][4], al[2][MT][4];            // [k-step][m-tile]
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) a_from_c(gb[mt][ks], ah[ks][mt], al[ks][mt]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    f2 eb[MT][2][2];                             // [m-tile][octave half f]: (cos, sin) adjoints of rows g, g + 8
                    mma3_step<true, MT>(eb, ah[0], al[0], fragL[(kR0 + 2 * c) * 32], fragL[(kR0 + 2 * c + 1) * 32]);
                    mma3_step<false, MT>(eb, ah[1], al[1], fragL[(kR0 + 6 + 2 * c) * 32], fragL[(kR0 + 6 + 2 * c + 1) * 32]);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            const float fk = f ? f1 : f0;
                            const f2 cs = e.cs[mt][c][f], sn = e.sn[mt][c][f];
                            acc0 = fmaf(fk, fmaf(eb[mt][f][0].y, cs.x, -eb[mt][f][0].x * sn.x), acc0);
                            acc1 = fmaf(fk, fmaf(eb[mt][f][1].y, cs.y, -eb[mt][f][1].x * sn.y), acc1);
                        }
                        const f2 a2 = quad_sum2(make_float2(acc0, acc1));
                        abar[2 * mt][c] = a2.x;
                        abar[2 * mt + 1][c] = a2.y;
                    }
                }
            }
            // ------------------------------------------------------------ 4. lane == sample
            const float o = row_slots_to_lanes<MT>(out, lane);
            float ga[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v[kSlots];
#pragma unroll
                for (int s = 0; s < kSlots; ++s) v[s] = abar[s][c];
                ga[c] = row_slots_to_lanes<MT>(v, lane);
            }
            const float res = sigmoidf_(o - 1.0f);
            const float sp = res * (1.0f - res) * pi_scale;
            const float gp0 = b.gp[0] + sp * b.s[0] * ga[0];
            const float gp1 = b.gp[1] + sp * ga[1];
            const float gp2 = b.gp[2] + sp * ga[2];
            if (lane < kRows && base + lane < total) {
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
    if (lane == 0 && rays.cull_stats != nullptr && tiles_visited) {
        atomicAdd(rays.cull_stats, (unsigned long long)tiles_culled);
        atomicAdd(rays.cull_stats + 1, (unsigned long long)tiles_visited);
    }
}

// min over the instances of the BOX SDF at every sample (the culling bound, see VsrdRays::union_bound)
__global__ void union_bound_kernel(SceneDev scene, RaysDev rays, float* __restrict__ bound) {
    __shared__ Instance s_inst[VSRD_MAX_INSTANCES];       // 15 floats per instance, staged once per CTA
    for (int i = threadIdx.x; i < scene.N; i += blockDim.x) load_instance(scene, i, s_inst[i]);
    __syncthreads();
    const size_t total = (size_t)rays.R * rays.M;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int r = (int)(idx / rays.M);
    const int j = (int)(idx - (size_t)r * rays.M);
    float x[3];
    sample_position(rays, r, j, x);
    float lowest = INFINITY;
    for (int i = 0; i < scene.N; ++i) {
        BoxEval b;
        box_eval(x, s_inst[i], b);
        lowest = fminf(lowest, b.value);
    }
    bound[idx] = lowest;
}

static int g_fwd_sms = 0;
static int g_fwd_mt = 2;      // m-tiles per warp tile (VSRD_FWD_MT=1 selects the 16-row variant)

// 2 = tcgen05 kernel (default), 0 = mma.sync kernel (VSRD_FIELD_IMPL=mma), 1 = SIMT cross-check (VSRD_FIELD_IMPL=simt); read per call
static int forward_impl() {
    const char* impl = getenv("VSRD_FIELD_IMPL");
    if (impl && strcmp(impl, "simt") == 0) return 1;
    if (impl && strcmp(impl, "mma") == 0) return 0;
    return 2;
}

static int forward_setup() {
    if (g_fwd_sms) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("vsrd_b200: no CUDA device%s");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail("vsrd_b200: cudaGetDeviceProperties failed%s");
    if (cudaFuncSetAttribute(field_forward_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)FwdCfg<1>::kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(field_forward_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)FwdCfg<2>::kSmemBytes) != cudaSuccess)
        return fail("vsrd_b200: cannot reserve %s of shared memory for field_forward_mma_kernel (built for sm_100a)", "217 KB");
    const char* mt = getenv("VSRD_FWD_MT");
    if (mt && (mt[0] == '1' || mt[0] == '2')) g_fwd_mt = mt[0] - '0';
    g_fwd_sms = prop.multiProcessorCount;
    return 0;
}

template <int MT>
static void launch_forward_mma(const SceneDev& s, const RaysDev& r, float4* field, size_t total, cudaStream_t st) {
    using Cfg = FwdCfg<MT>;
    const int tiles_per_inst = (int)((total + Cfg::kRows - 1) / Cfg::kRows);
    const long long all_tiles = (long long)s.N * tiles_per_inst;
    const long long want = (all_tiles + Cfg::kWarps - 1) / Cfg::kWarps;
    const int grid = (int)(want < g_fwd_sms ? want : g_fwd_sms);
    field_forward_mma_kernel<MT><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(s, r, field, tiles_per_inst);
}

}  // namespace vsrd

using namespace vsrd;

extern "C" {

static int launch_field(const SceneDev& s, const RaysDev& r, float* field, void* stream) {
    const size_t total = (size_t)r.R * r.M;
    if (total == 0) return 0;
    VSRD_CHECK_ARG(field != nullptr, "field is NULL");
    VSRD_CHECK_ARG(total < (size_t)1 << 31, "R*M must be < 2^31");
    if (forward_setup()) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (s.W && forward_impl() == 2) {
        if (launch_field_forward_umma(s, r, field, total, st)) return 1;
    } else if (s.W && forward_impl() == 0) {
        if (g_fwd_mt == 2) launch_forward_mma<2>(s, r, (float4*)field, total, st);
        else launch_forward_mma<1>(s, r, (float4*)field, total, st);
    } else {
        const dim3 grid((unsigned)((total + kThreads - 1) / kThreads), (unsigned)s.N);
        if (s.W) field_forward_kernel<true><<<grid, kThreads, 0, st>>>(s, r, (float4*)field);
        else field_forward_kernel<false><<<grid, kThreads, 0, st>>>(s, r, (float4*)field);
    }
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_field_forward(const VsrdScene* scene, const VsrdRays* rays, float* field, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    return launch_field(s, r, field, stream);
}

int vsrd_union_bound(const VsrdScene* scene, const VsrdRays* rays, float* union_bound, void* stream) {
    SceneDev s; RaysDev r;
    if (check_scene(scene, s) || check_rays(rays, r)) return 1;
    const size_t total = (size_t)r.R * r.M;
    if (total == 0) return 0;
    VSRD_CHECK_ARG(union_bound != nullptr, "union_bound is NULL");
    VSRD_CHECK_ARG(total < (size_t)1 << 31, "R*M must be < 2^31");
    union_bound_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s, r, union_bound);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_field_points(const VsrdScene* scene, const float* points, int num_points, float* field, void* stream) {
    SceneDev s;
    if (check_scene(scene, s)) return 1;
    VSRD_CHECK_ARG(num_points >= 0, "num_points must be non-negative");
    VSRD_CHECK_ARG(num_points == 0 || points != nullptr, "points is NULL");
    const RaysDev r{num_points, 1, points, nullptr, nullptr, nullptr, nullptr, nullptr};     // points mode of sample_position()
    return launch_field(s, r, field, stream);
}

}  // extern "C"
