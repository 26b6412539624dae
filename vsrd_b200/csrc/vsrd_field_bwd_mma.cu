// sm_100a kernels of the VSRD silhouette-renderer hot path, part 3b: tensor-core field backward (v6).
//
// Replaces the autograd double-backward replay of the reference (renderers.py:218-228 create_graph=True,
// main.py:859) for the residual-MLP instances.  Given the adjoints (dd, dG) of every (sample, instance)
// field value and spatial gradient it accumulates the gradient of
//      phi = dd * d(p; theta) + (R^T dG) . grad_p d(p; theta)
// w.r.t. the 1617 MLP weights and the 15 pose parameters of the instance ("one tangent + one reverse
// sweep"; the scalar restatement is vsrd_math.cuh::field_backward).
//
// One warp owns a tile of 16 samples of one instance in the mma.sync fragment layout of vsrd_frag.cuh:
//   1. lane == sample: position, box SDF, tangent direction v = R^T dG, PE arguments
//   2. dual forward sweep (value + tangent along v): mma.sync m16n8k16 contractions on bf16 hi + lo operands chained in
//      registers, weights and biases CENTRED over the output index so that the LayerNorm code needs no mean reductions
//      (vsrd_frag.cuh::stage_weight_fragments); LayerNorm outputs (z, zd) of layers 1..3 and the (Phi, phi) of their
//      GELUs go to a lane-private stash in TENSOR MEMORY (tcgen05.st), those of layer 4 stay in registers
//   3. reverse sweep layer by layer: transposed contractions, LayerNorm/GELU second-order adjoints, and the weight
//      gradient dW_l += hbar^T g + hdbar^T gd as a sample-contracted m16n8k16 whose operands are transposed in registers
//      by movmatrix (the adjoints' bf16 packs are shared with the transposed contraction); the GELU factors come from
//      the stashed (Phi, phi), no MUFU
//   4. lane == sample: chain through |p_x|, the box SDF and the pose
// The weight-gradient accumulators live in tensor memory too (tcgen05.ld / add in registers by the MMAs / tcgen05.st,
// lane-private, no atomics, deterministic), the pose accumulators in per-warp shared memory; both are reduced once per
// (CTA, instance) segment into one partial row; reduce_segment_rows_kernel sums the rows of each instance and maps the
// gradient of the centred weights back.
//
// Tensor memory here is a 256 KB lane-private scratchpad, not an MMA operand space: the kernel has no tcgen05.mma (the
// tcgen05 backward that was built and measured is slower, vsrd_field_bwd_umma.cu), but every warp owns 164 TMEM columns
// of its lane quarter (BwdCfg) and moves 4 / 8 / 16 of them per tcgen05.ld / tcgen05.st.  Against the shared-memory
// stash this removes 16 STS.64 + 16 LDS.64 per layer and tile with their address arithmetic and bank conflicts, the
// recomputation of the GELU terms in the reverse sweep, and 42 LDS.128 / STS.128 of accumulator traffic per tile, and
// leaves the shared-memory pipe to the shuffles and movmatrix: 331 M -> 285 M warp instructions.
//
// Persistent: gridDim.x CTAs split the N * tiles_per_inst warp tiles into contiguous ranges -- evenly, or with instance
// culling by live work (backward_ranges_kernel) -- so any N fills the SMs; a CTA restages the weight fragments when its
// range crosses an instance boundary.  Segment (cta b, instance i) writes partial row b + i (strictly increasing along
// the tile order, hence unique).
//
// History of the contraction format (cfg2 fine pass, N = 8, R = 1000, M = 199): 3xTF32 on m16n8k8 0.741 ms; + centring
// folded into the weights 0.695 ms; bf16 hi + lo on m16n8k16 (half the tensor instructions: both shapes issue at the same
// rate on B200, and the legacy tensor pipe was a quarter of the stall samples) + pose accumulators in the shared memory
// the smaller weight image frees 0.613 ms; (Phi, phi) stash in tensor memory 0.590 ms; + (z, zd) stash 0.560 ms; +
// accumulators 0.541 ms; + the encoding parked in shared memory across the hidden layers (spills 240 -> 44 B), sample
// positions kept for phase 4, pass-major weight-gradient MMAs, unrolled reverse loop 0.533 ms.
#include "vsrd_frag.cuh"
#include "vsrd_umma.cuh"

namespace vsrd {
namespace bwd5 {

constexpr int kAccHidden = 0;                 // hidden layer l: fragments 3 (l-1) + {inputs 0-7, inputs 8-15, bias}
constexpr int kAccL0 = 9;                     // layer 0: input tiles 0..5, bias
constexpr int kAccLast = 16;                  // last layer: this lane's 4 channels
constexpr int kAccPose = 17;                  // (sum obar, pose 0..2) (pose 3..6) (pose 7..10) (pose 11..14), lane == sample

// MT = m-tiles (16 samples) per warp tile.  Shipped: MT = 1, 12 warps x 168 registers (3 warps per scheduler).
// Measured and not kept: MT = 2 (8 warps x 255 registers): slower; 16 warps x 128 registers with the stash in tensor
// memory (shared memory no longer limits the warp count): 0.556 ms, the same as 12 warps at that point -- the spills
// cost what the fourth warp per scheduler hides -- and 16 warps leave no TMEM columns for the accumulators; software
// prefetch of the next tile's adjoints into L1 / L2: +0.5 %.
template <int MT>
struct BwdCfg {
    static_assert(MT == 1, "the tensor-memory layout below is the one-m-tile kernel's");
    static constexpr int kWarps = 12;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kRows = 16 * MT;
    static constexpr bool kPoseInRegs = false;             // true: 16 more registers per thread instead of 24 KB of shared memory
    static constexpr int kAccFrags = kPoseInRegs ? 17 : 21;
    static constexpr int kAccFloat4 = kAccFrags * 32;
    static constexpr int kEncPairs = 12;                   // per lane: (cos, sin) x 3 coordinates x 2 octave halves, row pairs
    // shared memory per warp: accumulator slots (pose fragments live here; fragments 0..16 only at the flush), the
    // flush row, the sample positions and the parked encoding
    static constexpr size_t kSmemBytes = frag::kWeightBytes
        + (size_t)kWarps * (kAccFloat4 * sizeof(float4) + 32 * sizeof(float) + 96 * sizeof(float) + kEncPairs * 32 * sizeof(float2));
    // tensor memory: every warp owns 164 columns of its lane quarter (3 warps per quarter, 492 of the 512 columns):
    //   [0, 96)    layers 1..3 x { (z, zd) of the LayerNorm outputs: 16 | (Phi, phi) of their GELU: 16 }
    //   [96, 164)  weight-gradient accumulator fragments 0..16 (4 columns each)
    static constexpr uint32_t kTmemWarpCols = 164, kTmemAccCol = 96;
    static constexpr int kTmemCols = 512;
    static_assert(kWarps % 4 == 0 && (kWarps / 4) * kTmemWarpCols <= kTmemCols, "warps of one lane quarter share its 512 columns");
    static_assert(kTmemAccCol == 3 * 32 && kTmemAccCol + 4 * 17 == kTmemWarpCols, "stash of 3 layers + 17 accumulator fragments");
};

using frag::f2;
using frag::bc;
using frag::mul2;
using frag::add2;
using frag::fma2;
using frag::hsum;

// LayerNorm (no affine, eps 1e-5) of one row and of its tangent, on channel pairs.  In: p = h, d = hd
// (this lane's 2 x 2 of the 16 channels), both CENTRED already: the weights that produced them were centred when they
// were staged (frag::stage_weight_fragments).  Out: p = z, d = zd, rs = 1/sigma, mz = mean(z * centred tangent).
__device__ __forceinline__ void ln_dual2(f2& p0, f2& p1, f2& d0, f2& d1, float& rs, float& mz) {
    constexpr float inv = 1.0f / kHid;
    // (16 var, sum v d) in one packed reduction; z = v rs, so mean(z d) = rs * sum(v d) / 16
    const f2 q = frag::quad_sum2(make_float2(hsum(fma2(p0, p0, mul2(p1, p1))), hsum(fma2(p0, d0, mul2(p1, d1)))));
    rs = rsqrtf(fmaf(q.x, inv, kLnEps));
    mz = rs * q.y * inv;
    p0 = mul2(p0, bc(rs)); p1 = mul2(p1, bc(rs));
    d0 = mul2(fma2(p0, bc(-mz), d0), bc(rs));
    d1 = mul2(fma2(p1, bc(-mz), d1), bc(rs));
}

// GELU value / first / second derivative factors of one channel pair.
__device__ __forceinline__ void gelu_pair(f2 z, f2 zd, f2& g, f2& gd, f2& g1, f2& g2) {
    f2 Phi, phi, zz;
    frag::gelu_terms2(z, Phi, phi, zz);
    g = mul2(z, Phi);
    g1 = fma2(z, phi, Phi);
    g2 = mul2(phi, fma2(zz, bc(-1.0f), bc(2.0f)));
    gd = mul2(g1, zd);
}

// The same from the (Phi, phi) the forward sweep kept (TMEM stash): no MUFU, bit-identical to gelu_pair.
__device__ __forceinline__ void gelu_pair_kept(f2 z, f2 zd, f2 Phi, f2 phi, f2& g, f2& gd, f2& g1, f2& g2) {
    g = mul2(z, Phi);
    g1 = fma2(z, phi, Phi);
    g2 = mul2(phi, fma2(mul2(z, z), bc(-1.0f), bc(2.0f)));
    gd = mul2(g1, zd);
}

// Adjoint of (LayerNorm -> GELU) and of its tangent for one row (vsrd_math.cuh::ln_gelu_reverse) w.r.t. the CENTRED
// LayerNorm input: the terms - mean(zb), - mean(zdb) of the full adjoint are applied by the centred transposed weights
// the result is multiplied with next (P W)^T hb = W^T (P hb), and cancel in the gradient of the centred weights.
// gb / gdb: adjoints of gelu(z) and of its tangent; out hb / hdb: adjoints of the (centred) LayerNorm input and
// of its tangent.  [0] / [1]: the lane's two channel pairs.
__device__ __forceinline__ void ln_gelu_reverse_row2(const f2 (&z)[2], const f2 (&zd)[2], const f2 (&g1)[2],
                                                     const f2 (&g2)[2], float rs, float m, const f2 (&gb)[2],
                                                     const f2 (&gdb)[2], f2 (&hb)[2], f2 (&hdb)[2]) {
    constexpr float inv = 1.0f / kHid;
    f2 zb[2], zdb[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        zdb[k] = mul2(gdb[k], g1[k]);
        zb[k] = fma2(gb[k], g1[k], mul2(mul2(gdb[k], g2[k]), zd[k]));
    }
    // sum(z zb) + sum(zd zdb) and sum(z zdb) in one packed quad reduction
    const f2 s = frag::quad_sum2(make_float2(
        hsum(fma2(z[0], zb[0], fma2(z[1], zb[1], fma2(zd[0], zdb[0], mul2(zd[1], zdb[1]))))),
        hsum(fma2(z[0], zdb[0], mul2(z[1], zdb[1])))));
    const float n_zz = s.x * -inv, n_zzdb = s.y * -inv;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        hdb[k] = mul2(bc(rs), fma2(z[k], bc(n_zzdb), zdb[k]));
        hb[k] = mul2(bc(rs), fma2(zd[k], bc(n_zzdb), fma2(hdb[k], bc(-m), fma2(z[k], bc(n_zz), zb[k]))));
    }
}

// Box part of the backward at one sample (vsrd_math.cuh::field_backward, first half).
struct PoseTerms {
    BoxEval b;
    float v[3], pbar[3], vbar[3], dimbar[3], coef[3];
};

__device__ __forceinline__ void pose_terms(const float x[3], const Instance& I, float pi_scale, float dd,
                                           const float dG[3], PoseTerms& p) {
    box_eval(x, I, p.b);
    const BoxEval& b = p.b;
#pragma unroll
    for (int k = 0; k < 3; ++k) p.v[k] = I.R[k] * dG[0] + I.R[3 + k] * dG[1] + I.R[6 + k] * dG[2];
    float vs = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) vs += p.v[k] * b.s[k] * b.a[k];
    const float inv_n = 1.0f / b.nrm;
    const float inv_n3 = inv_n * inv_n * inv_n;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float act = b.q[k] > 0.0f ? 1.0f : 0.0f;
        const float hess = act * p.v[k] * b.s[k] * inv_n - b.a[k] * vs * inv_n3;
        p.pbar[k] = dd * b.gp[k] + b.s[k] * hess;
        p.dimbar[k] = -(dd * (b.a[k] * inv_n + b.ind[k]) + hess);
        p.vbar[k] = b.gp[k];
    }
    p.coef[0] = b.s[0] * pi_scale;
    p.coef[1] = pi_scale;
    p.coef[2] = pi_scale;
}

// range_starts[b], b = 0..grid: the first tile of CTA b when `grid` CTAs split the N * tiles_per_inst tiles (instance-
// major order) into contiguous ranges.  Without culling marks the kernels split the tiles evenly themselves.  With them a block of tiles
// weighs its live samples (the census vsrd_composite_backward keeps next to the marks, common.cuh::census_offset) plus
// one per tile (a dead tile still has to be scanned), so that the CTAs get equal shares of LIVE work: late in the
// schedule 3/4 of the tiles are dead, unevenly over the instances, and an even split of all tiles leaves the CTAs of the
// nearest instances with several times the work of the others (measured: 75 % of the tiles skipped, kernel only 1.9x
// faster).  One CTA: scans the block weights, places every range start by bisection + interpolation.  (Kept out of
// the big kernel: computing the ranges there cost it 100 B of extra register spills.  Counting the marks here instead
// of in the compositing kernel took 18 - 25 us on one SM.)
constexpr int kRangeThreads = 1024;
constexpr int kRangeMaxEntries = 4096;
__global__ void __launch_bounds__(kRangeThreads) backward_ranges_kernel(const unsigned char* __restrict__ live, int N, int tiles_per_inst,
                                                                       int group, int grid, long long* __restrict__ range_starts) {
    const long long all_tiles = (long long)N * tiles_per_inst;
    // entry k = `group` consecutive census blocks of one instance (group = 1 unless R is huge)
    __shared__ long long s_prefix[kRangeMaxEntries];       // entry weights, then their inclusive prefix sums
    __shared__ long long s_warp[kRangeThreads / 32];
    const int* census = reinterpret_cast<const int*>(live + census_offset(N, tiles_per_inst));
    const int nb = census_blocks(tiles_per_inst);
    const int entries_per_inst = (nb + group - 1) / group, entry_tiles = group * VSRD_CENSUS_BLOCK_TILES;
    const int num_entries = N * entries_per_inst;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (num_entries + kRangeThreads - 1) / kRangeThreads;
    const int k0 = min(threadIdx.x * per, num_entries), k1 = min(k0 + per, num_entries);
    long long mine = 0;
    for (int k = k0; k < k1; ++k) {
        const int inst = k / entries_per_inst, j = k - inst * entries_per_inst;
        const int len = min(entry_tiles, tiles_per_inst - j * entry_tiles);
        long long alive = 0;
        for (int q = j * group; q < min((j + 1) * group, nb); ++q) alive += __ldg(census + inst * nb + q);
        mine += alive + len;
        s_prefix[k] = mine;                                 // running sum inside the thread's run
    }
    long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long up = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += up; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long v = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long up = __shfl_up_sync(kFull, v, o); if (lane >= o) v += up; }
        s_warp[lane] = v;
    }
    __syncthreads();
    const long long before = incl - mine + (warp ? s_warp[warp - 1] : 0);
    for (int k = k0; k < k1; ++k) s_prefix[k] += before;
    __syncthreads();
    const long long total_weight = s_prefix[num_entries - 1];
    for (int b = threadIdx.x; b <= grid; b += blockDim.x) {
        if (b == grid) { range_starts[b] = all_tiles; continue; }
        const long long target = total_weight * b / grid;   // first entry whose inclusive prefix exceeds the target
        int lo = 0, hi = num_entries - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_prefix[mid] > target) hi = mid; else lo = mid + 1; }
        const int inst = lo / entries_per_inst, j = lo - inst * entries_per_inst;
        const int len = min(entry_tiles, tiles_per_inst - j * entry_tiles);
        const long long base = lo ? s_prefix[lo - 1] : 0, weight = s_prefix[lo] - base;
        range_starts[b] = (long long)inst * tiles_per_inst + (long long)j * entry_tiles + (target - base) * len / weight;
    }
}

// PAIR (MT = 1 only): the lane == sample phases (1 and 4) serve TWO 16-sample tiles at once, lanes 0-15 the first and
// lanes 16-31 the second tile of a pair taken from the live list; the fragment phases (2, 3) run once per half with the
// one-m-tile register footprint.  Without it half the lanes idle through phases 1 and 4 (~10 % of the instructions).
template <int MT, int PAIR>
__global__ void __launch_bounds__(BwdCfg<MT>::kThreads, 1) field_backward_mma_kernel(
        SceneDev scene, RaysDev rays, const float4* __restrict__ adjoint, float* __restrict__ partials,
        const long long* __restrict__ range_starts, int tiles_per_inst) {
    static_assert(PAIR == 0 || MT == 1, "tile pairs are a variant of the one-m-tile kernel");
    using Cfg = BwdCfg<MT>;
    constexpr int kWarpsB = Cfg::kWarps, kThreadsB = Cfg::kThreads, kRows = Cfg::kRows, kSlots = 2 * MT;
    constexpr int kAccFloat4 = Cfg::kAccFloat4, kAccFrags = Cfg::kAccFrags;
    constexpr uint32_t kTmemAccCol = Cfg::kTmemAccCol;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sF = reinterpret_cast<float4*>(smem_raw);
    float* sTail = reinterpret_cast<float*>(sF + frag::kFragFloat4);
    float4* sAcc = reinterpret_cast<float4*>(sTail + frag::kTailFloats);
    float* sRed = reinterpret_cast<float*>(sAcc + kWarpsB * kAccFloat4);   // [warp][32]: last layer (17) + pose (15)
    float* sPos = sRed + kWarpsB * 32;                      // [warp][3][32]: sample positions, phase 1 -> phase 4
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // [warp][12][32] pairs: the encoding waits here between layer 0 forward and layer 0 reverse (24 registers the
    // hidden-layer loops can use: register spills 240 -> 44 bytes)
    float2* sEnc = reinterpret_cast<float2*>(sPos + kWarpsB * 96) + (size_t)warp * Cfg::kEncPairs * 32 + lane;
    // Tensor memory as the warp's lane-private scratchpad (BwdCfg): this kernel issues no tcgen05.mma, it uses the
    // 256 KB of TMEM next to the tensor cores for what shared memory served before -- the (z, zd) stash, the GELU terms
    // (Phi, phi) the reverse sweep would otherwise recompute (2 MUFU + ~14 instructions per channel pair) and the
    // weight-gradient accumulators -- as x4 / x8 / x16 tcgen05.st / tcgen05.ld of the lane's own row: no bank
    // conflicts, no address arithmetic, and the shared-memory pipe is left to the shuffles and movmatrix.  Measured
    // (cfg2 fine pass): 0.613 -> 0.590 ms (Phi, phi) -> 0.560 ms (z, zd) -> 0.541 ms (accumulators).
    __shared__ uint32_t s_tmem;
    if (threadIdx.x < 32) umma::tmem_alloc<Cfg::kTmemCols>(&s_tmem);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    // a warp reaches the 32 lanes of quarter (warp % 4) only
    const uint32_t tmem_base = s_tmem + ((uint32_t)(warp & 3) << 21) + (uint32_t)(warp >> 2) * Cfg::kTmemWarpCols;

    const int t = lane & 3;
    const int quad_base = lane & ~3;
    float4* accL = sAcc + warp * kAccFloat4 + lane;         // accL[fragment * 32]
    const float4* fragL = sF + lane;

    const int total = rays.R * rays.M;
    // CTA ranges: balanced by live work (backward_ranges_kernel) with culling, an even split of all tiles without
    const long long all_tiles = (long long)scene.N * tiles_per_inst;
    const long long begin = range_starts ? range_starts[blockIdx.x] : all_tiles * blockIdx.x / gridDim.x;
    const long long end = range_starts ? range_starts[blockIdx.x + 1] : all_tiles * (blockIdx.x + 1) / gridDim.x;
    const float pi_scale = kPiF / scene.scale;
    unsigned tiles_visited = 0, tiles_culled = 0;          // per warp; two atomics per warp at the end
    constexpr int kChunkTiles = 2048;
    __shared__ unsigned short s_list[kChunkTiles];
    __shared__ float s_mean[frag::kNumMeans];
    __shared__ int s_live;

    for (long long seg = begin; seg < end;) {
        const int inst = (int)(seg / tiles_per_inst);
        const long long seg_end = min(end, (long long)(inst + 1) * tiles_per_inst);
        __syncthreads();                                    // previous segment fully flushed
        frag::stage_weight_fragments(scene.W + (size_t)inst * kNumW, sF, sTail, s_mean);
#pragma unroll
        for (int f = 0; f < kAccFrags; ++f) accL[f * 32] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        {
            const float zero4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int f = 0; f < kAccPose; ++f) umma::tmem_st4(tmem_base + kTmemAccCol + 4u * f, zero4);
        }
        __syncthreads();
        Instance I;
        load_instance(scene, inst, I);
        const float4* adj_inst = adjoint + (size_t)inst * total;
        const f2 w4p0 = make_float2(sTail[frag::kTailW4 + 2 * t], sTail[frag::kTailW4 + 2 * t + 1]);
        const f2 w4p1 = make_float2(sTail[frag::kTailW4 + 8 + 2 * t], sTail[frag::kTailW4 + 8 + 2 * t + 1]);
        const float b4 = sTail[frag::kTailB4];
        float pose_reg[Cfg::kPoseInRegs ? 16 : 1];          // (sum obar, pose 0..14) of this lane's samples
#pragma unroll
        for (int k = 0; k < (Cfg::kPoseInRegs ? 16 : 1); ++k) pose_reg[k] = 0.0f;

        // With culling the cost of a tile is bimodal; a static stride over all tiles leaves the slowest warp near the
        // worst case.  vsrd_composite_backward marked the tiles that received a non-zero adjoint (rays.live): per chunk
        // the CTA compacts the marked tile indices IN ORDER and the warps stride over that list: balanced, and the
        // accumulation order stays a function of the inputs only (deterministic).
        const unsigned char* live = rays.live != nullptr ? rays.live + (size_t)inst * tiles_per_inst : nullptr;
#pragma unroll 1
        for (long long chunk = seg; chunk < seg_end; chunk += kChunkTiles) {
        const int chunk_tiles = (int)min((long long)kChunkTiles, seg_end - chunk);
        int live_tiles = chunk_tiles;
        if (live != nullptr) {
            if (warp == 0) {
                const unsigned char* flags = live + (chunk - (long long)inst * tiles_per_inst);
                const int per = (chunk_tiles + 31) / 32, first = lane * per, last = min(first + per, chunk_tiles);
                static_assert(kChunkTiles <= 64 * 32, "one 64-bit mask of marks per lane");
                unsigned long long marks = 0;               // one pass over the lane's (at most 64) marks
#pragma unroll 8
                for (int q = first; q < last; ++q) marks |= (unsigned long long)(__ldg(flags + q) ? 1 : 0) << (q - first);
                const int mine = __popcll(marks);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += up; }
                int pos = incl - mine;
                for (; marks; marks &= marks - 1) s_list[pos++] = (unsigned short)(first + __ffsll((long long)marks) - 1);
                if (lane == 31) s_live = incl;
            }
            __syncthreads();
            live_tiles = s_live;
            if (lane == 0 && warp == 0) { tiles_visited += (unsigned)chunk_tiles; tiles_culled += (unsigned)(chunk_tiles - live_tiles); }
        }
        constexpr int kPerIter = PAIR ? 2 : 1;              // list entries one warp iteration consumes
#pragma unroll 1
        for (int k = warp * kPerIter; k < live_tiles; k += kWarpsB * kPerIter) {
            const int half = PAIR ? (lane >> 4) : 0;        // PAIR: lanes 16..31 own the second tile of the pair
            const int entry = min(k + half, live_tiles - 1);
            const bool has_tile = k + half < live_tiles;
            const long long tile = chunk + (live != nullptr ? (int)s_list[entry] : entry);
            const int base = (int)(tile - (long long)inst * tiles_per_inst) * kRows;
            const int row = lane & (kRows - 1);             // MT = 1 without PAIR: lanes 16..31 shadow rows 0..15 with zero adjoints
            const int idx = min(base + row, total - 1);
            const bool valid = (PAIR ? has_tile : lane < kRows) && base + row < total;
            const int r = idx / rays.M;
            const int j = idx - r * rays.M;
            // ------------------------------------------------------------ 1. lane == sample
            float4 adj = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // zero adjoints contribute exactly zero
            if (valid) adj = __ldg(adj_inst + idx);
            const unsigned nonzero = __ballot_sync(kFull, adj.x != 0.0f || adj.y != 0.0f || adj.z != 0.0f || adj.w != 0.0f);
            if (nonzero == 0u) continue;
            float pe_arg[3], pe_tan[3];                     // this lane's sample: PE arguments and their tangents along v
            {
                float x[3];
                sample_position(rays, r, j, x);
#pragma unroll
                for (int c = 0; c < 3; ++c) sPos[(warp * 3 + c) * 32 + lane] = x[c];      // phase 4 picks it up again
                BoxEval b;
                box_eval(x, I, b);
                const float m[3] = {fabsf(b.p[0]), b.p[1], b.p[2]};
                const float coef[3] = {b.s[0] * pi_scale, pi_scale, pi_scale};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float vc = I.R[c] * adj.y + I.R[3 + c] * adj.z + I.R[6 + c] * adj.w;
                    pe_arg[c] = kPiF * (m[c] / scene.scale);
                    pe_tan[c] = coef[c] * vc;
                }
            }
            float ga[3] = {0.0f, 0.0f, 0.0f}, gad[3] = {0.0f, 0.0f, 0.0f};     // encoding adjoints of this lane's sample
#pragma unroll 1
            for (int hh = 0; hh < kPerIter; ++hh) {
            const int loff = 16 * hh;                       // lanes that hold the samples of this half
            if (PAIR && ((nonzero >> loff) & 0xffffu) == 0u) continue;
            f2 arow[MT][3], adrow[MT][3];                  // pairs = rows (g, g + 8) of each m-tile
            float ddrow[kSlots];
            {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    f2 v[MT];
                    frag::lanes_to_row_pairs<MT>(pe_arg[c], lane, v, loff);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) arow[mt][c] = v[mt];
                    frag::lanes_to_row_pairs<MT>(pe_tan[c], lane, v, loff);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) adrow[mt][c] = v[mt];
                }
                f2 v[MT];
                frag::lanes_to_row_pairs<MT>(adj.x, lane, v, loff);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) { ddrow[2 * mt] = v[mt].x; ddrow[2 * mt + 1] = v[mt].y; }
            }
            // ------------------------------------------------------------ 2. dual forward sweep
            frag::EncodingT<MT> e;
            frag::encode2<MT>(arow, t, e);
            const float f0 = (float)(1 << t), f1 = 16.0f * f0;
            f2 h[MT][2][2], hd[MT][2][2];                  // [m-tile][n-tile][row g | row g + 8]
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const f2 bias = make_float2(sTail[8 * nt + 2 * t], sTail[8 * nt + 2 * t + 1]);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) { h[mt][nt][0] = bias; h[mt][nt][1] = bias; }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {       // one k-step per coordinate: frequencies t (a0, a1) and t + 4 (a2, a3), (cos, sin) pairs
                const float4 w0 = fragL[(frag::kB0 + 2 * c) * 32], w1 = fragL[(frag::kB0 + 2 * c + 1) * 32];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    uint32_t ah[4], al[4], adh[4], adl[4];
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        const f2 cs = e.cs[mt][c][f], sn = e.sn[mt][c][f];             // pairs over rows (g, g + 8)
                        const f2 da = mul2(adrow[mt][c], bc(f ? f1 : f0));
                        const f2 dcs = mul2(mul2(da, bc(-1.0f)), sn), dsn = mul2(da, cs);   // tangents of (cos, sin)
                        frag::pack_bf16_split_scalar(cs.x, sn.x, ah[2 * f], al[2 * f]);
                        frag::pack_bf16_split_scalar(cs.y, sn.y, ah[2 * f + 1], al[2 * f + 1]);
                        frag::pack_bf16_split_scalar(dcs.x, dsn.x, adh[2 * f], adl[2 * f]);
                        frag::pack_bf16_split_scalar(dcs.y, dsn.y, adh[2 * f + 1], adl[2 * f + 1]);
                    }
                    // the value accumulators start from the bias, the tangent ones from zero
                    if (c == 0) frag::mma3b_quad<false, true>(h[mt][0], h[mt][1], hd[mt][0], hd[mt][1], ah, al, adh, adl, w0, w1);
                    else frag::mma3b_quad<false, false>(h[mt][0], h[mt][1], hd[mt][0], hd[mt][1], ah, al, adh, adl, w0, w1);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)                      // park the encoding until layer 0's reverse step
#pragma unroll
                for (int f = 0; f < 2; ++f) { sEnc[(4 * c + 2 * f) * 32] = e.cs[0][c][f]; sEnc[(4 * c + 2 * f + 1) * 32] = e.sn[0][c][f]; }
            // layers 1..3: LayerNorm -> GELU -> linear; lane t keeps 1/sigma and mz of layer t + 1
            float rreg[kSlots], mreg[kSlots];
#pragma unroll
            for (int s = 0; s < kSlots; ++s) { rreg[s] = 0.0f; mreg[s] = 0.0f; }
#pragma unroll 1
            for (int l = 1; l <= 3; ++l) {
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    f2& p0 = h[s >> 1][0][s & 1];
                    f2& p1 = h[s >> 1][1][s & 1];
                    f2& d0 = hd[s >> 1][0][s & 1];
                    f2& d1 = hd[s >> 1][1][s & 1];
                    float rs, mz;
                    ln_dual2(p0, p1, d0, d1, rs, mz);
                    if (t == l - 1) { rreg[s] = rs; mreg[s] = mz; }
                    {                                       // (z, zd) of the row -> tensor memory
                        const float zs[8] = {p0.x, p0.y, p1.x, p1.y, d0.x, d0.y, d1.x, d1.y};
                        umma::tmem_st8(tmem_base + (uint32_t)((l - 1) * 32 + 8 * s), zs);
                    }
                    f2 Phi, phi, zz;
                    float keep[8];                          // (Phi, phi) of the row's two channel pairs -> tensor memory
                    frag::gelu_terms2(p0, Phi, phi, zz);
                    keep[0] = Phi.x; keep[1] = Phi.y; keep[2] = phi.x; keep[3] = phi.y;
                    d0 = mul2(d0, fma2(p0, phi, Phi));
                    p0 = mul2(p0, Phi);
                    frag::gelu_terms2(p1, Phi, phi, zz);
                    keep[4] = Phi.x; keep[5] = Phi.y; keep[6] = phi.x; keep[7] = phi.y;
                    d1 = mul2(d1, fma2(p1, phi, Phi));
                    p1 = mul2(p1, Phi);
                    umma::tmem_st8(tmem_base + (uint32_t)((l - 1) * 32 + 16 + 8 * s), keep);
                }
                f2 hn[MT][2][2], hdn[MT][2][2];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const f2 bias = make_float2(sTail[16 * l + 8 * nt + 2 * t], sTail[16 * l + 8 * nt + 2 * t + 1]);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) { hn[mt][nt][0] = bias; hn[mt][nt][1] = bias; }
                }
                const float4 w0 = fragL[(frag::kB1 + 2 * (l - 1)) * 32], w1 = fragL[(frag::kB1 + 2 * (l - 1) + 1) * 32];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    uint32_t ah[4], al[4], adh[4], adl[4];
                    frag::a_bf16_from_c(h[mt], ah, al);
                    frag::a_bf16_from_c(hd[mt], adh, adl);
                    frag::mma3b_quad<false, true>(hn[mt][0], hn[mt][1], hdn[mt][0], hdn[mt][1], ah, al, adh, adl, w0, w1);
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int q = 0; q < 2; ++q) { h[mt][nt][q] = hn[mt][nt][q]; hd[mt][nt][q] = hdn[mt][nt][q]; }
            }
            // ------------------------------------------------------------ 3. reverse sweep
            // layer 4 (16 -> 1), LayerNorm outputs stay in registers; hb / hdb: adjoints of the output of
            // linear layer 3 and of its tangent (C layout)
            f2 hb[MT][2][2], hdb[MT][2][2];
            umma::wait_st();                                // this sweep's stash and the previous tile's accumulators have landed
            {
                float last4[4];
                umma::tmem_ld4(tmem_base + kTmemAccCol + 4u * kAccLast, last4); umma::wait_ld();
                f2 last0 = make_float2(last4[0], last4[1]), last1 = make_float2(last4[2], last4[3]);
                float obsum = 0.0f;
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    f2 z[2] = {h[s >> 1][0][s & 1], h[s >> 1][1][s & 1]};
                    f2 zd[2] = {hd[s >> 1][0][s & 1], hd[s >> 1][1][s & 1]};
                    float rs, mz;
                    ln_dual2(z[0], z[1], zd[0], zd[1], rs, mz);
                    f2 g[2], gd[2], g1[2], g2[2];
                    gelu_pair(z[0], zd[0], g[0], gd[0], g1[0], g2[0]);
                    gelu_pair(z[1], zd[1], g[1], gd[1], g1[1], g2[1]);
                    const f2 oo = frag::quad_sum2(make_float2(hsum(fma2(w4p0, g[0], mul2(w4p1, g[1]))),
                                                              hsum(fma2(w4p0, gd[0], mul2(w4p1, gd[1])))));
                    const float out = oo.x + b4, outd = oo.y;
                    const float res = sigmoidf_(out - 1.0f);
                    const float sp = res * (1.0f - res);
                    const float obar = fmaf(ddrow[s], sp, sp * (1.0f - 2.0f * res) * outd);
                    const float odbar = sp;
                    obsum += obar;
                    last0 = fma2(bc(obar), g[0], fma2(bc(odbar), gd[0], last0));
                    last1 = fma2(bc(obar), g[1], fma2(bc(odbar), gd[1], last1));
                    const f2 gb[2] = {mul2(w4p0, bc(obar)), mul2(w4p1, bc(obar))};
                    const f2 gdb[2] = {mul2(w4p0, bc(odbar)), mul2(w4p1, bc(odbar))};
                    f2 hbv[2], hdbv[2];
                    ln_gelu_reverse_row2(z, zd, g1, g2, rs, mz, gb, gdb, hbv, hdbv);
                    hb[s >> 1][0][s & 1] = hbv[0]; hb[s >> 1][1][s & 1] = hbv[1];
                    hdb[s >> 1][0][s & 1] = hdbv[0]; hdb[s >> 1][1][s & 1] = hdbv[1];
                }
                {
                    const float v4[4] = {last0.x, last0.y, last1.x, last1.y};
                    umma::tmem_st4(tmem_base + kTmemAccCol + 4u * kAccLast, v4);
                }
                if constexpr (Cfg::kPoseInRegs) {
                    pose_reg[0] += obsum;
                } else {
                    float4 p0 = accL[kAccPose * 32];
                    p0.x += obsum;
                    accL[kAccPose * 32] = p0;
                }
            }
            // (unrolled: 1 % faster than a rolled loop since the stash left shared memory)
#pragma unroll
            for (int l = 3; l >= 1; --l) {
                // adjoints of gelu(z_l) and its tangent: W_l^T hb, W_l^T hdb
                f2 gb[MT][2][2], gdb[MT][2][2];
                uint32_t th[MT][4], tl[MT][4], tdh[MT][4], tdl[MT][4];     // hb / hdb as bf16 (hi, lo) packs: A operand here, transposed below
                {
                    const float4 w0 = fragL[(frag::kBR1 + 2 * (l - 1)) * 32], w1 = fragL[(frag::kBR1 + 2 * (l - 1) + 1) * 32];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        frag::a_bf16_from_c(hb[mt], th[mt], tl[mt]);
                        frag::a_bf16_from_c(hdb[mt], tdh[mt], tdl[mt]);
                        frag::mma3b_quad<true, true>(gb[mt][0], gb[mt][1], gdb[mt][0], gdb[mt][1], th[mt], tl[mt], tdh[mt], tdl[mt], w0, w1);
                        frag::wgrad_a_from_packs(th[mt], tl[mt]);         // the weight gradient's A operand: outputs x samples
                        frag::wgrad_a_from_packs(tdh[mt], tdl[mt]);
                    }
                }
                f2 D[3][2];                                 // this layer's weight-gradient accumulators (12 TMEM columns)
                const uint32_t tacc = tmem_base + kTmemAccCol + 4u * (kAccHidden + 3 * (l - 1));
                {
                    float v8[8], v4[4];
                    umma::tmem_ld8(tacc, v8); umma::tmem_ld4(tacc + 8u, v4); umma::wait_ld();
                    D[0][0] = make_float2(v8[0], v8[1]); D[0][1] = make_float2(v8[2], v8[3]);
                    D[1][0] = make_float2(v8[4], v8[5]); D[1][1] = make_float2(v8[6], v8[7]);
                    D[2][0] = make_float2(v4[0], v4[1]); D[2][1] = make_float2(v4[2], v4[3]);
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    f2 z[2][2], zd[2][2], g[2][2], gd[2][2], g1[2][2], g2[2][2];      // [row half][channel pair]
                    float keptz[16];                        // (z, zd) of the forward sweep: [row half][z pair 0, z pair 1, zd pair 0, zd pair 1]
                    float kept[16];                         // (Phi, phi): [row half][pair][Phi.xy phi.xy]
                    umma::tmem_ld16(tmem_base + (uint32_t)((l - 1) * 32), keptz);
                    umma::tmem_ld16(tmem_base + (uint32_t)((l - 1) * 32 + 16), kept);
                    umma::wait_ld();
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        z[hf][0] = make_float2(keptz[8 * hf], keptz[8 * hf + 1]); z[hf][1] = make_float2(keptz[8 * hf + 2], keptz[8 * hf + 3]);
                        zd[hf][0] = make_float2(keptz[8 * hf + 4], keptz[8 * hf + 5]); zd[hf][1] = make_float2(keptz[8 * hf + 6], keptz[8 * hf + 7]);
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const f2 Phi = make_float2(kept[8 * hf + 4 * q], kept[8 * hf + 4 * q + 1]);
                            const f2 phi = make_float2(kept[8 * hf + 4 * q + 2], kept[8 * hf + 4 * q + 3]);
                            gelu_pair_kept(z[hf][q], zd[hf][q], Phi, phi, g[hf][q], gd[hf][q], g1[hf][q], g2[hf][q]);
                        }
                    }
                    // weight gradient of linear layer l over the 16 samples of this m-tile
                    frag::wgrad_hidden(D, th[mt], tl[mt], tdh[mt], tdl[mt], g, gd);
                    // LayerNorm / GELU adjoint: hb, hdb <- adjoints of the output of linear layer l - 1
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int s = 2 * mt + hf;
                        const float rs = __shfl_sync(kFull, rreg[s], quad_base | (l - 1));
                        const float mz = __shfl_sync(kFull, mreg[s], quad_base | (l - 1));
                        const f2 gbv[2] = {gb[mt][0][hf], gb[mt][1][hf]};
                        const f2 gdbv[2] = {gdb[mt][0][hf], gdb[mt][1][hf]};
                        f2 hbv[2], hdbv[2];
                        ln_gelu_reverse_row2(z[hf], zd[hf], g1[hf], g2[hf], rs, mz, gbv, gdbv, hbv, hdbv);
                        hb[mt][0][hf] = hbv[0]; hb[mt][1][hf] = hbv[1];
                        hdb[mt][0][hf] = hdbv[0]; hdb[mt][1][hf] = hdbv[1];
                    }
                }
                {
                    const float v8[8] = {D[0][0].x, D[0][0].y, D[0][1].x, D[0][1].y, D[1][0].x, D[1][0].y, D[1][1].x, D[1][1].y};
                    const float v4[4] = {D[2][0].x, D[2][0].y, D[2][1].x, D[2][1].y};
                    umma::tmem_st8(tacc, v8); umma::tmem_st4(tacc + 8u, v4);
                }
            }
            // layer 0: weight gradient against the encoding and its tangent, then the encoding adjoint
            //   abar_c  = sum_k 2^k (ebar_sin cos - ebar_cos sin - da (edbar_cos cos + edbar_sin sin))
            //   adbar_c = sum_k 2^k (edbar_sin cos - edbar_cos sin)
            float abar[kSlots][3], adbar[kSlots][3];
#pragma unroll
            for (int c = 0; c < 3; ++c)                      // the encoding comes back from shared memory
#pragma unroll
                for (int f = 0; f < 2; ++f) { e.cs[0][c][f] = sEnc[(4 * c + 2 * f) * 32]; e.sn[0][c][f] = sEnc[(4 * c + 2 * f + 1) * 32]; }
            {
                f2 D0[7][2];
                const uint32_t tacc0 = tmem_base + kTmemAccCol + 4u * kAccL0;
                {
                    float v16[16], v8[8], v4[4];
                    umma::tmem_ld16(tacc0, v16); umma::tmem_ld8(tacc0 + 16u, v8); umma::tmem_ld4(tacc0 + 24u, v4); umma::wait_ld();
#pragma unroll
                    for (int n = 0; n < 4; ++n) { D0[n][0] = make_float2(v16[4 * n], v16[4 * n + 1]); D0[n][1] = make_float2(v16[4 * n + 2], v16[4 * n + 3]); }
#pragma unroll
                    for (int n = 0; n < 2; ++n) { D0[4 + n][0] = make_float2(v8[4 * n], v8[4 * n + 1]); D0[4 + n][1] = make_float2(v8[4 * n + 2], v8[4 * n + 3]); }
                    D0[6][0] = make_float2(v4[0], v4[1]); D0[6][1] = make_float2(v4[2], v4[3]);
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int s0 = 2 * mt, s1 = 2 * mt + 1;
                    uint32_t ah[4], al[4], adh[4], adl[4];
                    frag::a_bf16_from_c(hb[mt], ah, al);
                    frag::a_bf16_from_c(hdb[mt], adh, adl);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float acc0 = 0.0f, acc1 = 0.0f, accd0 = 0.0f, accd1 = 0.0f;
                        f2 eb[2][2], edb[2][2];                 // [octave half f]: (cos, sin) adjoints of rows g, g + 8
                        frag::mma3b_quad<true, true>(eb[0], eb[1], edb[0], edb[1], ah, al, adh, adl,
                                                     fragL[(frag::kBR0 + 2 * c) * 32], fragL[(frag::kBR0 + 2 * c + 1) * 32]);
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            const float fk = f ? f1 : f0;
                            const f2 cs = e.cs[mt][c][f], sn = e.sn[mt][c][f];
                            const f2 da = mul2(adrow[mt][c], bc(fk));
                            acc0 += fk * (eb[f][0].y * cs.x - eb[f][0].x * sn.x - da.x * (edb[f][0].x * cs.x + edb[f][0].y * sn.x));
                            accd0 += fk * (edb[f][0].y * cs.x - edb[f][0].x * sn.x);
                            acc1 += fk * (eb[f][1].y * cs.y - eb[f][1].x * sn.y - da.y * (edb[f][1].x * cs.y + edb[f][1].y * sn.y));
                            accd1 += fk * (edb[f][1].y * cs.y - edb[f][1].x * sn.y);
                        }
                        const f2 qa = frag::quad_sum2(make_float2(acc0, acc1));
                        const f2 qd = frag::quad_sum2(make_float2(accd0, accd1));
                        abar[s0][c] = qa.x; abar[s1][c] = qa.y;
                        adbar[s0][c] = qd.x; adbar[s1][c] = qd.y;
                    }
                    // weight gradient against the encoding and its tangent: the same packs, transposed
                    frag::wgrad_a_from_packs(ah, al);
                    frag::wgrad_a_from_packs(adh, adl);
                    frag::wgrad_bias(D0[6], ah, al);
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            const float fk = f ? f1 : f0;
                            const f2 cs = e.cs[mt][c][f], sn = e.sn[mt][c][f];
                            const f2 da = mul2(adrow[mt][c], bc(fk));
                            const f2 dcs = mul2(mul2(da, bc(-1.0f)), sn), dsn = mul2(da, cs);   // tangents of (cos, sin)
                            frag::wgrad_tile_scalar(D0[2 * c + f], ah, al, cs.x, sn.x, cs.y, sn.y);
                            frag::wgrad_tile_scalar(D0[2 * c + f], adh, adl, dcs.x, dsn.x, dcs.y, dsn.y);
                        }
                }
                {
                    float v16[16], v8[8];
#pragma unroll
                    for (int n = 0; n < 4; ++n) { v16[4 * n] = D0[n][0].x; v16[4 * n + 1] = D0[n][0].y; v16[4 * n + 2] = D0[n][1].x; v16[4 * n + 3] = D0[n][1].y; }
#pragma unroll
                    for (int n = 0; n < 2; ++n) { v8[4 * n] = D0[4 + n][0].x; v8[4 * n + 1] = D0[4 + n][0].y; v8[4 * n + 2] = D0[4 + n][1].x; v8[4 * n + 3] = D0[4 + n][1].y; }
                    const float v4[4] = {D0[6][0].x, D0[6][0].y, D0[6][1].x, D0[6][1].y};
                    umma::tmem_st16(tacc0, v16); umma::tmem_st8(tacc0 + 16u, v8); umma::tmem_st4(tacc0 + 24u, v4);
                }
            }
            // back to lane == sample: the lanes of this half pick up their rows' encoding adjoints
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float va[kSlots], vd[kSlots];
#pragma unroll
                for (int sl = 0; sl < kSlots; ++sl) { va[sl] = abar[sl][c]; vd[sl] = adbar[sl][c]; }
                const float a_mine = frag::row_slots_to_lanes<MT>(va, lane);
                const float d_mine = frag::row_slots_to_lanes<MT>(vd, lane);
                if (!PAIR || half == hh) { ga[c] = a_mine; gad[c] = d_mine; }
            }
            }   // halves
            // ------------------------------------------------------------ 4. lane == sample: pose
            {
                float x[3];                                 // stashed by phase 1 (reloading the ray costs an L2 round trip here)
#pragma unroll
                for (int c = 0; c < 3; ++c) x[c] = sPos[(warp * 3 + c) * 32 + lane];
                const float dG[3] = {adj.y, adj.z, adj.w};
                PoseTerms p;
                pose_terms(x, I, pi_scale, adj.x, dG, p);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    p.pbar[c] = fmaf(ga[c], p.coef[c], p.pbar[c]);
                    p.vbar[c] = fmaf(gad[c], p.coef[c], p.vbar[c]);
                }
                float pose[kNumPose];
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    pose[m] = -(I.R[3 * m] * p.pbar[0] + I.R[3 * m + 1] * p.pbar[1] + I.R[3 * m + 2] * p.pbar[2]);
                    pose[3 + m] = p.dimbar[m];
#pragma unroll
                    for (int k = 0; k < 3; ++k) pose[6 + 3 * m + k] = p.b.y[m] * p.pbar[k] + dG[m] * p.vbar[k];
                }
                if constexpr (Cfg::kPoseInRegs) {
                    if (PAIR || lane < kRows) {            // without PAIR the shadow lanes hold copies of rows 0..15: count each sample once
#pragma unroll
                        for (int m = 0; m < kNumPose; ++m) pose_reg[1 + m] += pose[m];
                    }
                } else {
                    float4 a0 = accL[kAccPose * 32], a1 = accL[(kAccPose + 1) * 32];
                    float4 a2 = accL[(kAccPose + 2) * 32], a3 = accL[(kAccPose + 3) * 32];
                    a0.y += pose[0]; a0.z += pose[1]; a0.w += pose[2];
                    a1.x += pose[3]; a1.y += pose[4]; a1.z += pose[5]; a1.w += pose[6];
                    a2.x += pose[7]; a2.y += pose[8]; a2.z += pose[9]; a2.w += pose[10];
                    a3.x += pose[11]; a3.y += pose[12]; a3.z += pose[13]; a3.w += pose[14];
                    accL[kAccPose * 32] = a0; accL[(kAccPose + 1) * 32] = a1;
                    accL[(kAccPose + 2) * 32] = a2; accL[(kAccPose + 3) * 32] = a3;
                }
            }
            __syncwarp();
        }
        if (live != nullptr) __syncthreads();               // the tile list is rebuilt for the next chunk
        }
        // ---------------------------------------------------------------- flush this segment
        {
            umma::wait_st();                                // tensor memory -> this warp's shared-memory slots (read across warps below)
#pragma unroll
            for (int f = 0; f < kAccPose; ++f) {
                float v4[4];
                umma::tmem_ld4(tmem_base + kTmemAccCol + 4u * f, v4); umma::wait_ld();
                accL[f * 32] = make_float4(v4[0], v4[1], v4[2], v4[3]);
            }
            // last layer: reduce over the 8 quads (lane bits 2..4); bias / pose: over all 32 lanes
            float4 last = accL[kAccLast * 32];
            float v[16];
            if constexpr (Cfg::kPoseInRegs) {
                v[0] = 0.25f * pose_reg[0];                // obar was accumulated by all 4 lanes of each quad
#pragma unroll
                for (int k = 1; k < 16; ++k) v[k] = pose_reg[k];
            } else {
                const float4 a0 = accL[kAccPose * 32], a1 = accL[(kAccPose + 1) * 32];
                const float4 a2 = accL[(kAccPose + 2) * 32], a3 = accL[(kAccPose + 3) * 32];
                v[0] = 0.25f * a0.x;                       // obar was accumulated by all 4 lanes of each quad
                v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
                v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                v[8] = a2.x; v[9] = a2.y; v[10] = a2.z; v[11] = a2.w;
                v[12] = a3.x; v[13] = a3.y; v[14] = a3.z; v[15] = a3.w;
            }
#pragma unroll
            for (int sh = 16; sh >= 1; sh >>= 1) {
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] += __shfl_xor_sync(kFull, v[k], sh);
                if (sh >= 4) {
                    last.x += __shfl_xor_sync(kFull, last.x, sh);
                    last.y += __shfl_xor_sync(kFull, last.y, sh);
                    last.z += __shfl_xor_sync(kFull, last.z, sh);
                    last.w += __shfl_xor_sync(kFull, last.w, sh);
                }
            }
            float* red = sRed + warp * 32;
            if (lane < 4) {
                red[2 * lane] = last.x; red[2 * lane + 1] = last.y;
                red[8 + 2 * lane] = last.z; red[8 + 2 * lane + 1] = last.w;
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 16; ++k) red[16 + k] = v[k];
            }
        }
        __syncthreads();
        {
            float* out = partials + ((size_t)blockIdx.x + inst) * kGradStride;
            for (int f = threadIdx.x; f < kGradStride; f += kThreadsB) {
                float s = 0.0f;
                if (f >= kW4) {
#pragma unroll
                    for (int w = 0; w < kWarpsB; ++w) s += sRed[w * 32 + f - kW4];
                } else {
                    int fragment, o, i, fan_in;
                    if (f < kW1) {
                        o = f / (kEnc + 1); i = f - o * (kEnc + 1); fan_in = kEnc;
                        fragment = kAccL0 + (i < kEnc ? (i >> 3) : 6);
                    } else {
                        int q = f - kW1;
                        const int l = q / kWStride;
                        q -= l * kWStride;
                        o = q / (kHid + 1); i = q - o * (kHid + 1); fan_in = kHid;
                        fragment = kAccHidden + 3 * l + (i < kHid ? (i >> 3) : 2);
                    }
                    const bool bias = i == fan_in;
                    const int tt = bias ? 0 : ((i & 7) >> 1);
                    const int comp = ((o >> 3) << 1) | (bias ? 0 : (i & 1));
                    const float* p = reinterpret_cast<const float*>(sAcc + fragment * 32 + 4 * (o & 7) + tt) + comp;
#pragma unroll
                    for (int w = 0; w < kWarpsB; ++w) s += p[(size_t)w * kAccFloat4 * 4];
                }
                out[f] = s;
            }
        }
        seg = seg_end;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_free<Cfg::kTmemCols>(s_tmem);
    if (lane == 0 && rays.cull_stats != nullptr && tiles_visited) {
        atomicAdd(rays.cull_stats, (unsigned long long)tiles_culled);
        atomicAdd(rays.cull_stats + 1, (unsigned long long)tiles_visited);
    }
}

// Sum the partial rows of every instance: CTA b wrote row b + inst for every instance whose tiles intersect its range
// [range_starts[b], range_starts[b + 1]).  The kernel accumulated the gradient w.r.t. the CENTRED weights P W_l, P b_l of
// layers 0..3 (frag::stage_weight_fragments); dW = P dW' subtracts, per input column, the mean over the 16 outputs.
// grid (5, N): blockIdx.x = layer 0 | hidden layer 1..3 | last layer + pose.
constexpr int kReduceThreads = 1024;
__global__ void __launch_bounds__(kReduceThreads) reduce_segment_rows_kernel(
        const float* __restrict__ partials, const long long* __restrict__ range_starts, int grid, int tiles_per_inst,
        float* __restrict__ gloc, float* __restrict__ grot, float* __restrict__ gdim, float* __restrict__ gW) {
    __shared__ int s_first, s_last;
    __shared__ float s_sum[kW1];                            // the largest section: layer 0, 16 x 49
    const int inst = blockIdx.y, section = blockIdx.x;
    const long long first = (long long)inst * tiles_per_inst, last = first + tiles_per_inst;     // [first, last)
    if (threadIdx.x == 0) { s_first = grid; s_last = -1; }
    __syncthreads();
    const long long all_tiles = (long long)gridDim.y * tiles_per_inst;
    for (int b = threadIdx.x; b < grid; b += blockDim.x) {
        const long long lo = range_starts ? range_starts[b] : all_tiles * b / grid;
        const long long hi = range_starts ? range_starts[b + 1] : all_tiles * (b + 1) / grid;
        if (lo < hi && lo < last && hi > first) { atomicMin(&s_first, b); atomicMax(&s_last, b); }
    }
    __syncthreads();
    const int base = section == 0 ? 0 : (section <= 3 ? kW1 + (section - 1) * kWStride : kW4);
    const int count = section == 0 ? kW1 : (section <= 3 ? kWStride : kGradStride - kW4);
    const int fan = section == 0 ? kEnc + 1 : kHid + 1;     // row length [inputs + bias] of the section's layer
    for (int k = threadIdx.x; k < count; k += blockDim.x) {
        const float* col = partials + (size_t)inst * kGradStride + base + k;
        float s = 0.0f;
        for (int b = s_first; b <= s_last; b += 8) {        // eight rows in flight: the loop is pure L2 latency otherwise
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = b + u <= s_last ? __ldg(col + (size_t)(b + u) * kGradStride) : 0.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        s_sum[k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < count; k += blockDim.x) {
        float s = s_sum[k];
        const int f = base + k;
        if (section <= 3) {                                 // P along the output index: same input column, all 16 outputs
            const int col = k % fan;
            float mean = 0.0f;
#pragma unroll
            for (int o = 0; o < kHid; ++o) mean += s_sum[o * fan + col];
            s -= mean * (1.0f / kHid);
        }
        if (f < kNumW) gW[(size_t)inst * kNumW + f] = s;
        else if (f < kNumW + 3) gloc[3 * inst + (f - kNumW)] = s;
        else if (f < kNumW + 6) gdim[3 * inst + (f - kNumW - 3)] = s;
        else grot[9 * inst + (f - kNumW - 6)] = s;
    }
}

static int g_sms = 0;
// Shipped configuration: one m-tile per warp pass, lane == sample phases shared by a PAIR of tiles, 12 warps x 168 registers:
// 0.738 ms at R=1000 S=100 N=8.  Measured and removed (round 1): the same without pairing 0.770 ms; two m-tiles per warp, 8 warps
// x 255 registers 0.810 ms.

static int setup() {
    if (g_sms) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("vsrd_b200: no CUDA device%s");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail("vsrd_b200: cudaGetDeviceProperties failed%s");
    if (cudaFuncSetAttribute(field_backward_mma_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)BwdCfg<1>::kSmemBytes) != cudaSuccess)
        return fail("vsrd_b200: cannot reserve %s of shared memory for field_backward_mma_kernel (built for sm_100a)", "166 KB");
    g_sms = prop.multiProcessorCount;
    return 0;
}

template <int MT, int PAIR>
static int launch(const SceneDev& s, const RaysDev& r, const float* adjoint, float* partials,
                  float* gloc, float* grot, float* gdim, float* gW, cudaStream_t st) {
    using Cfg = BwdCfg<MT>;
    const long long total = (long long)r.R * r.M;
    const int tiles_per_inst = (int)((total + Cfg::kRows - 1) / Cfg::kRows);
    const long long all_tiles = (long long)s.N * tiles_per_inst;
    const long long want = (all_tiles + Cfg::kWarps - 1) / Cfg::kWarps;
    const int grid = (int)(want < g_sms ? want : g_sms);
    // the CTA ranges live in one extra row behind the g_sms + N partial rows (backward_mma_partial_rows)
    long long* range_starts = reinterpret_cast<long long*>(partials + (size_t)(g_sms + s.N) * kGradStride);
    static_assert(kGradStride % 2 == 0, "the extra row must be 8-byte aligned");
    if (r.live != nullptr) {
        int group = 1;                                      // census blocks per scan entry, <= kRangeMaxEntries entries
        while ((long long)s.N * ((census_blocks(tiles_per_inst) + group - 1) / group) > kRangeMaxEntries) group *= 2;
        backward_ranges_kernel<<<1, kRangeThreads, 0, st>>>(r.live, s.N, tiles_per_inst, group, grid, range_starts);
    } else {
        range_starts = nullptr;                             // even split, computed by the kernels themselves
    }
    VSRD_CHECK_LAUNCH();
    field_backward_mma_kernel<MT, PAIR><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(
        s, r, (const float4*)adjoint, partials, range_starts, tiles_per_inst);
    VSRD_CHECK_LAUNCH();
    const dim3 rgrid(5, (unsigned)s.N);
    reduce_segment_rows_kernel<<<rgrid, kReduceThreads, 0, st>>>(partials, range_starts, grid, tiles_per_inst, gloc, grot, gdim, gW);
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // namespace bwd5

int backward_mma_tile_rows() {
    if (bwd5::setup()) return -1;
    return 16;
}

int backward_mma_partial_rows(int num_instances) {
    if (bwd5::setup()) return -1;
    return bwd5::g_sms + num_instances + 1;      // + one row for the CTA ranges (launch())
}

int launch_field_backward_mma(const SceneDev& s, const RaysDev& r, const float* adjoint, float* partials,
                              float* gloc, float* grot, float* gdim, float* gW, cudaStream_t st) {
    if (bwd5::setup()) return 1;
    return bwd5::launch<1, 1>(s, r, adjoint, partials, gloc, grot, gdim, gW, st);
}

}  // namespace vsrd

