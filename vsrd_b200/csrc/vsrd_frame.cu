// sm_100a kernels of the VSRD hot path, part 4: the per-step rows around the renderer.
//   a14/a15  projection_step_kernel   all-view box projection, Hungarian matching on the target view,
//                                     DIoU + smooth-L1 projection losses and their adjoint, one launch,
//                                     no host round trip (the reference: 136 Python calls, 137 host syncs
//                                     and a scipy call per step, scripts/main.py:339-415)
//   a2       ray_cdf_* / select_rays / gather_targets   weighted ray draw without replacement
//                                     (main.py:620-627) from a per-frame CDF, and the target gather (:656)
//   soft_masks_kernel                 SoftRasterizer soft masks for the synthetic frames
//                                     (transforms/geometric_transforms.py:267-309)
//   step_state_kernel                 device-resident annealing schedule (main.py:420-431) so a captured
//                                     CUDA graph can be replayed across optimisation steps
// All of this is latency-bound bookkeeping (KBs of data) except ray_cdf_build / soft_masks, which
// stream the [V,H,W,N] soft masks once per FRAME and are HBM-bound.
#include "vsrd_common.cuh"
#include "vsrd_frame_math.cuh"

namespace vsrd {

namespace {

constexpr int kProjThreads = 256;
constexpr double kInfD = 1.0e300;

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
    // deterministic: fixed shuffle tree, then a serial sum of the 8 warp totals
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kProjThreads / 32; ++w) t += s_red[w];
    return t;
}

// Warp-parallel shortest-augmenting-path assignment (Kuhn-Munkres with potentials), lanes = columns.
// cost is row-major [n][n] in shared memory; on return col_of_row[i] is the column matched to row i.
// Same optimum as scipy.optimize.linear_sum_assignment (main.py:383-386) whenever it is unique.
__device__ void warp_assignment(const float* cost, int n, double* u /*[n] shared*/, int* col_of_row /*[n] shared*/) {
    const int lane = threadIdx.x & 31;
    const bool col = lane < n;
    double v = 0.0;
    int p = -1;                                   // row matched to this column
    if (lane < n) u[lane] = 0.0;
    __syncwarp();
    for (int i = 0; i < n; ++i) {
        double minv = kInfD;
        int way = -1;
        bool used = false;
        int j0 = -1, i0 = i;
        while (true) {
            if (lane == j0) used = true;
            if (col && !used) {
                const double cur = (double)cost[i0 * n + lane] - u[i0] - v;
                if (cur < minv) { minv = cur; way = j0; }
            }
            double best = (col && !used) ? minv : kInfD;
            int j1 = (col && !used) ? lane : 64;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(kFull, best, o);
                const int oj = __shfl_xor_sync(kFull, j1, o);
                if (ob < best || (ob == best && oj < j1)) { best = ob; j1 = oj; }
            }
            const double delta = best;
            __syncwarp();
            if (col && used) { u[p] += delta; v -= delta; }
            else if (col) minv -= delta;
            if (lane == 0) u[i] += delta;         // the virtual column's row
            __syncwarp();
            j0 = j1;
            i0 = __shfl_sync(kFull, p, j0);
            if (i0 < 0) break;
        }
        // augment along the alternating path back to the virtual column
        while (j0 >= 0) {
            const int jprev = __shfl_sync(kFull, way, j0);
            const int pprev = jprev >= 0 ? __shfl_sync(kFull, p, jprev) : i;
            if (lane == j0) p = pprev;
            j0 = jprev;
        }
    }
    if (col) col_of_row[p] = lane;
    __syncwarp();
}

struct ProjArgs {
    int V, N, target_view;
    float height, width, eps;
    const float* extrinsics;      // [V,16]
    const float* intrinsics;      // [V,9]
    const float* world;           // [N,24]
    const float* gt_boxes;        // [V,N,4] or NULL (projection only)
    const uint8_t* visible;       // [V,N] or NULL (all visible)
    const int64_t* fixed_gt;      // [N] or NULL (run the assignment)
    float* boxes;                 // [V,N,4]
    int64_t* gt_indices;          // [N]
    float* losses;                // [2]
    float* grad_world;            // [2,N,24]
    float* scratch;               // [V*N*8] box gradients + [V*N*48] corner partials
};

__global__ void __launch_bounds__(kProjThreads) projection_step_kernel(ProjArgs a) {
    __shared__ float s_cost[VSRD_MAX_INSTANCES * VSRD_MAX_INSTANCES];
    __shared__ double s_u[VSRD_MAX_INSTANCES];
    __shared__ int s_match[VSRD_MAX_INSTANCES];
    __shared__ float s_red[kProjThreads / 32];
    const int V = a.V, N = a.N, pairs = V * N;

    // ---- a14: every (view, instance) pair
    for (int idx = threadIdx.x; idx < pairs; idx += blockDim.x) {
        const int v = idx / N, n = idx % N;
        BoxProjection bp;
        project_box(a.extrinsics + 16 * v, a.intrinsics + 9 * v, a.world + 24 * n, a.height, a.width, a.eps, bp);
#pragma unroll
        for (int k = 0; k < 4; ++k) a.boxes[4 * idx + k] = bp.box[k];
    }
    if (a.gt_boxes == nullptr) return;
    __syncthreads();

    // ---- a15: matching on the target view (cost = -DIoU, main.py:374-386)
    if (a.fixed_gt != nullptr) {
        for (int k = threadIdx.x; k < N; k += blockDim.x) s_match[k] = (int)a.fixed_gt[k];
    } else {
        const float* pd = a.boxes + 4 * (size_t)a.target_view * N;
        const float* gt = a.gt_boxes + 4 * (size_t)a.target_view * N;
        for (int e = threadIdx.x; e < N * N; e += blockDim.x)
            s_cost[e] = -diou_pair(pd + 4 * (e / N), gt + 4 * (e % N));
        __syncthreads();
        if (threadIdx.x < 32) warp_assignment(s_cost, N, s_u, s_match);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) a.gt_indices[k] = s_match[k];

    // ---- projection losses over the visible matched pairs of every view (main.py:391-415)
    float iou_sum = 0.0f, l1_sum = 0.0f, count = 0.0f;
    float* gbox = a.scratch;                              // [pairs][2][4]
    for (int idx = threadIdx.x; idx < pairs; idx += blockDim.x) {
        const int v = idx / N, k = idx % N, g = s_match[k];
        const bool vis = a.visible == nullptr || a.visible[v * N + g] != 0;
        float gi[4] = {0.0f, 0.0f, 0.0f, 0.0f}, gl[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (vis) {
            const float* pd = a.boxes + 4 * idx;
            const float* gt = a.gt_boxes + 4 * ((size_t)v * N + g);
            iou_sum += diou_loss_pair(pd, gt, gi);
            l1_sum += smooth_l1_pair(pd, gt, gl);
            count += 1.0f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { gbox[8 * idx + c] = gi[c]; gbox[8 * idx + 4 + c] = gl[c]; }
    }
    iou_sum = block_sum_256(iou_sum, s_red);
    l1_sum = block_sum_256(l1_sum, s_red);
    count = block_sum_256(count, s_red);
    if (threadIdx.x == 0) {
        a.losses[0] = iou_sum / count;                    // mean of an empty set is NaN, like torch.mean
        a.losses[1] = l1_sum / (4.0f * count);
    }
    if (a.grad_world == nullptr) return;

    // ---- adjoint: d loss / d world corners, per pair, then a fixed-order sum over the views
    float* part = a.scratch + 8 * (size_t)pairs;          // [pairs][2][24]
    const float inv_iou = 1.0f / count, inv_l1 = 1.0f / (4.0f * count);
    for (int idx = threadIdx.x; idx < pairs; idx += blockDim.x) {
        const int v = idx / N, n = idx % N;
        const float* E = a.extrinsics + 16 * v;
        const float* K = a.intrinsics + 9 * v;
        BoxProjection bp;
        project_box(E, K, a.world + 24 * n, a.height, a.width, a.eps, bp);
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            float g[4], gw[24];
            const float scale = which ? inv_l1 : inv_iou;
#pragma unroll
            for (int c = 0; c < 4; ++c) g[c] = gbox[8 * idx + 4 * which + c] * scale;
#pragma unroll
            for (int c = 0; c < 24; ++c) gw[c] = 0.0f;
            project_box_backward(E, K, bp, a.height, a.width, a.eps, g, gw);
#pragma unroll
            for (int c = 0; c < 24; ++c) part[(size_t)idx * 48 + 24 * which + c] = gw[c];
        }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * N * 24; o += blockDim.x) {
        const int which = o / (N * 24), n = (o / 24) % N, c = o % 24;
        float sum = 0.0f;
        for (int v = 0; v < V; ++v) sum += part[((size_t)v * N + n) * 48 + 24 * which + c];
        a.grad_world[o] = sum;
    }
}

// =============================================================================================
// a2: ray selection.  Per frame: weight = max over instances of the soft masks, inclusive CDF in double.
// =============================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;                       // pixels per thread
constexpr int kScanTile = kScanThreads * kScanItems; // 4096 pixels per CTA

__device__ __forceinline__ double block_scan_excl_256(double v, double* s_warp, double& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    double base = 0.0, tot = 0.0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) { if (w < warp) base += s_warp[w]; tot += s_warp[w]; }
    total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) ray_cdf_local_kernel(const float* __restrict__ masks, int64_t P, int N,
                                                                     double* __restrict__ cdf, double* __restrict__ tile_sums) {
    __shared__ double s_warp[kScanThreads / 32];
    __shared__ float s_w[kScanTile];
    const int64_t tile = (int64_t)blockIdx.x * kScanTile;
    // coalesced: consecutive lanes read consecutive pixels (N contiguous floats each)
    for (int k = threadIdx.x; k < kScanTile; k += kScanThreads) {
        const int64_t p = tile + k;
        float m = 0.0f;
        if (p < P) {
            m = __ldg(masks + p * N);
            for (int n = 1; n < N; ++n) m = fmaxf(m, __ldg(masks + p * N + n));
        }
        s_w[k] = m;
    }
    __syncthreads();
    const int first = threadIdx.x * kScanItems;
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) mine += (double)s_w[first + k];
    double total;
    double run = block_scan_excl_256(mine, s_warp, total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        run += (double)s_w[first + k];
        if (tile + first + k < P) cdf[tile + first + k] = run;
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) ray_cdf_tiles_kernel(double* tile_sums, int tiles) {
    // exclusive scan of the tile totals, one CTA, sequential chunks (tiles ~ 2200 for a KITTI-360 frame)
    __shared__ double s_chunk[1024];
    const int per = (tiles + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, tiles);
    double sum = 0.0;
    for (int t = lo; t < hi; ++t) sum += tile_sums[t];
    s_chunk[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = 0.0;
        for (int t = 0; t < 1024; ++t) { const double c = s_chunk[t]; s_chunk[t] = run; run += c; }
    }
    __syncthreads();
    double run = s_chunk[threadIdx.x];
    for (int t = lo; t < hi; ++t) { const double c = tile_sums[t]; tile_sums[t] = run; run += c; }
}

__global__ void __launch_bounds__(kScanThreads) ray_cdf_offset_kernel(double* __restrict__ cdf, int64_t P,
                                                                      const double* __restrict__ tile_sums) {
    const double off = tile_sums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    for (int k = threadIdx.x; k < kScanTile; k += kScanThreads)
        if (base + k < P) cdf[base + k] += off;
}

__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 ctr) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

constexpr int kSelectThreads = 1024;
constexpr int kSelectMaxRays = 8192;

// Sequential weighted sampling without replacement == i.i.d. draws from the full distribution with
// repeats rejected, in draw order.  One CTA: each round draws 1024 candidates by inverse-CDF search,
// rejects those already accepted or repeated earlier in the round, and appends the survivors in order.
__global__ void __launch_bounds__(kSelectThreads) select_rays_kernel(
        const double* __restrict__ cdf, int64_t P, const double* __restrict__ uniforms, int max_draws,
        uint64_t seed, const VsrdStepState* __restrict__ state, int R, int64_t* __restrict__ out, int32_t* __restrict__ status) {
    extern __shared__ int32_t s_sel[];                 // accepted [R] + candidates [1024]
    int32_t* s_acc = s_sel;
    int32_t* s_cand = s_sel + R;
    __shared__ int s_warp_cnt[kSelectThreads / 32];
    if (state != nullptr) seed = state->seed;
    const double total = cdf[P - 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int count = 0;
    if (!(total > 0.0)) {                               // no pixel has weight: nothing can be drawn
        for (int t = threadIdx.x; t < R; t += blockDim.x) out[t] = -1;
        if (threadIdx.x == 0 && status != nullptr) status[0] = R;
        return;
    }
    const int max_rounds = uniforms ? (max_draws + kSelectThreads - 1) / kSelectThreads : 4096;
    for (int round = 0; round < max_rounds && count < R; ++round) {
        const int draw = round * kSelectThreads + threadIdx.x;
        bool have = true;
        double u;
        if (uniforms) {
            have = draw < max_draws;
            u = have ? uniforms[draw] : 0.0;
        } else {
            const uint4 x = philox4x32_10(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                                          make_uint4((uint32_t)draw, 0u, 7u, 0u));
            u = ((double)(((uint64_t)x.x << 21) ^ (uint64_t)(x.y >> 11)) ) * (1.0 / 9007199254740992.0);   // 53 bits
        }
        int32_t cand = -1;
        if (have) {
            const double x = fmin(u * total, total * (1.0 - 1.1102230246251565e-16));
            int64_t lo = 0, hi = P - 1;                 // first index with cdf > x (clamped to the last pixel)
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (cdf[mid] > x) hi = mid; else lo = mid + 1; }
            cand = (int32_t)lo;
        }
        s_cand[threadIdx.x] = cand;
        __syncthreads();
        bool keep = have;
        if (keep) {
            for (int t = 0; t < count; ++t) if (s_acc[t] == cand) { keep = false; break; }
        }
        if (keep) {
            for (int t = 0; t < (int)threadIdx.x; ++t) if (s_cand[t] == cand) { keep = false; break; }
        }
        const unsigned ballot = __ballot_sync(kFull, keep);
        if (lane == 0) s_warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kSelectThreads / 32; ++w) { if (w < warp) before += s_warp_cnt[w]; all += s_warp_cnt[w]; }
        const int pos = count + before + __popc(ballot & ((1u << lane) - 1u));
        if (keep && pos < R) s_acc[pos] = cand;
        count = min(R, count + all);
        __syncthreads();
    }
    for (int t = threadIdx.x; t < R; t += blockDim.x) out[t] = t < count ? (int64_t)s_acc[t] : (int64_t)-1;
    if (threadIdx.x == 0 && status != nullptr) status[0] = R - count;    // 0 = success
}

__global__ void gather_targets_kernel(const float* __restrict__ masks, const int64_t* __restrict__ pix,
                                      const int64_t* __restrict__ gt_indices, int R, int N, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * N) return;
    const int r = idx / N, k = idx % N;
    const int64_t g = gt_indices ? gt_indices[k] : k;
    const int64_t p = pix[r];
    out[idx] = p >= 0 ? __ldg(masks + p * N + g) : 0.0f;
}

// =============================================================================================
// Soft masks of the synthetic frames: sigmoid(signed pixel distance to the instance polygon / temperature)
// out [V,H,W,N] (instance-minor, the layout main.py permutes the masks to at :300-315).
// =============================================================================================
__global__ void __launch_bounds__(256) soft_masks_kernel(const float* __restrict__ polygons, const int32_t* __restrict__ counts,
                                                         int V, int N, int PV, int H, int W, float temperature,
                                                         float* __restrict__ out) {
    extern __shared__ float s_poly[];                  // [N][PV][2] of this view
    __shared__ int s_cnt[VSRD_MAX_INSTANCES];
    const int v = blockIdx.y;
    for (int k = threadIdx.x; k < N * PV * 2; k += blockDim.x) s_poly[k] = polygons[(size_t)v * N * PV * 2 + k];
    for (int k = threadIdx.x; k < N; k += blockDim.x) s_cnt[k] = counts[v * N + k];
    __syncthreads();
    const int HW = H * W;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        const float px = (float)(pix % W), py = (float)(pix / W);
        float* o = out + ((size_t)v * HW + pix) * N;
        for (int n = 0; n < N; ++n) {
            const int cnt = s_cnt[n];
            float m = 0.0f;                            // instance absent from this view: zero mask (main.py:236-246)
            if (cnt >= 3) m = sigmoidf_(polygon_signed_distance(s_poly + (size_t)n * PV * 2, cnt, px, py) / temperature);
            o[n] = m;
        }
    }
}

// =============================================================================================
// Device-resident schedule (main.py:420-431): one thread advances the step and re-derives the scalars.
// =============================================================================================
__global__ void step_state_kernel(VsrdStepState* st, VsrdSchedule cfg, int64_t set_step) {
    const int64_t step = set_step >= 0 ? set_step : st->step + 1;
    const double x = (double)step / (double)cfg.num_steps;
    const double c = (cos(3.14159265358979323846 * x) + 1.0) / 2.0;
    st->step = step;
    st->temperature = (float)(c * ((double)cfg.max_temperature - (double)cfg.min_temperature) + (double)cfg.min_temperature);
    st->std_deviation = (float)(c * ((double)cfg.max_std_deviation - (double)cfg.min_std_deviation) + (double)cfg.min_std_deviation);
    st->cosine_ratio = (float)x;
    st->eikonal_weight = step >= cfg.warmup_steps ? cfg.eikonal_weight : 0.0f;
    // splitmix64 of (base seed, step): a fresh counter-based stream per step
    uint64_t z = cfg.seed + 0x9E3779B97F4A7C15ull * (uint64_t)(step + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    st->seed = z ^ (z >> 31);
}

}  // namespace

}  // namespace vsrd

using namespace vsrd;

// vsrd.operations.project_box_3d (geometric_operations.py:343-389) for a batch of camera-frame boxes: one thread per
// box; the backward recomputes the projection and applies its adjoint (project_box_backward with E = identity).
namespace vsrd {
__constant__ float kIdentity4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};

__global__ void project_box_3d_kernel(const float* __restrict__ boxes, const float* __restrict__ K, int num, float eps,
                                      float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num) return;
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(K + i);
    BoxProjection bp;
    project_box(kIdentity4, k, boxes + 24 * (size_t)b, 0.0f, 0.0f, eps, bp, false);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[4 * (size_t)b + i] = bp.box[i];
}

__global__ void project_box_3d_backward_kernel(const float* __restrict__ boxes, const float* __restrict__ K, int num, float eps,
                                               const float* __restrict__ grad_out, float* __restrict__ grad_boxes) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num) return;
    float k[9], g[4], gw[24];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(K + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = grad_out[4 * (size_t)b + i];
#pragma unroll
    for (int i = 0; i < 24; ++i) gw[i] = 0.0f;
    BoxProjection bp;
    project_box(kIdentity4, k, boxes + 24 * (size_t)b, 0.0f, 0.0f, eps, bp, false);
    project_box_backward(kIdentity4, k, bp, 0.0f, 0.0f, eps, g, gw, false);
#pragma unroll
    for (int i = 0; i < 24; ++i) grad_boxes[24 * (size_t)b + i] = gw[i];
}
}  // namespace vsrd

extern "C" {

size_t vsrd_projection_scratch_floats(int num_views, int num_instances) {
    return (size_t)num_views * (size_t)num_instances * 56;
}

int vsrd_projection_step(const VsrdViews* views, int num_instances, const float* world_boxes,
                         const float* gt_boxes_2d, const uint8_t* visible, const int64_t* fixed_gt_indices,
                         float* boxes_2d, int64_t* gt_indices, float* losses, float* grad_world_boxes,
                         float* scratch, void* stream) {
    VSRD_CHECK_ARG(views != nullptr, "views is NULL");
    VSRD_CHECK_ARG(views->num_views >= 1, "num_views must be positive");
    VSRD_CHECK_ARG(views->extrinsics && views->intrinsics, "view matrices must not be NULL");
    VSRD_CHECK_ARG(views->height > 0 && views->width > 0, "image size must be positive");
    VSRD_CHECK_ARG(num_instances >= 1 && num_instances <= VSRD_MAX_INSTANCES, "num_instances must be in [1, 32]");
    VSRD_CHECK_ARG(world_boxes && boxes_2d, "world_boxes / boxes_2d must not be NULL");
    if (gt_boxes_2d != nullptr) {
        VSRD_CHECK_ARG(views->target_view >= 0 && views->target_view < views->num_views, "target_view out of range");
        VSRD_CHECK_ARG(gt_indices && losses && scratch, "gt_indices / losses / scratch must not be NULL when gt boxes are given");
    }
    ProjArgs a{views->num_views, num_instances, views->target_view, (float)views->height, (float)views->width, 1e-6f,
               views->extrinsics, views->intrinsics, world_boxes, gt_boxes_2d, visible, fixed_gt_indices,
               boxes_2d, gt_indices, losses, grad_world_boxes, scratch};
    projection_step_kernel<<<1, kProjThreads, 0, (cudaStream_t)stream>>>(a);
    VSRD_CHECK_LAUNCH();
    return 0;
}

size_t vsrd_ray_cdf_scratch_doubles(int64_t num_pixels) {
    return (size_t)((num_pixels + kScanTile - 1) / kScanTile);
}

int vsrd_ray_cdf_build(const float* soft_masks, int64_t num_pixels, int num_instances, double* cdf, double* scratch,
                       void* stream) {
    VSRD_CHECK_ARG(soft_masks && cdf && scratch, "soft_masks / cdf / scratch must not be NULL");
    VSRD_CHECK_ARG(num_pixels >= 1 && num_pixels < ((int64_t)1 << 31), "num_pixels must be in [1, 2^31)");
    VSRD_CHECK_ARG(num_instances >= 1 && num_instances <= VSRD_MAX_INSTANCES, "num_instances must be in [1, 32]");
    const int tiles = (int)((num_pixels + kScanTile - 1) / kScanTile);
    cudaStream_t st = (cudaStream_t)stream;
    ray_cdf_local_kernel<<<tiles, kScanThreads, 0, st>>>(soft_masks, num_pixels, num_instances, cdf, scratch);
    VSRD_CHECK_LAUNCH();
    ray_cdf_tiles_kernel<<<1, 1024, 0, st>>>(scratch, tiles);
    VSRD_CHECK_LAUNCH();
    ray_cdf_offset_kernel<<<tiles, kScanThreads, 0, st>>>(cdf, num_pixels, scratch);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_select_rays(const double* cdf, int64_t num_pixels, const double* uniforms, int max_draws, uint64_t seed,
                     const VsrdStepState* step_state, int num_rays, int64_t* pixel_indices, int32_t* status, void* stream) {
    VSRD_CHECK_ARG(cdf && pixel_indices, "cdf / pixel_indices must not be NULL");
    VSRD_CHECK_ARG(num_pixels >= 1 && num_pixels < ((int64_t)1 << 31), "num_pixels must be in [1, 2^31)");
    VSRD_CHECK_ARG(num_rays >= 1 && num_rays <= kSelectMaxRays, "num_rays must be in [1, 8192]");
    VSRD_CHECK_ARG(uniforms == nullptr || max_draws >= num_rays, "max_draws must be >= num_rays when uniforms are injected");
    const size_t smem = sizeof(int32_t) * ((size_t)num_rays + kSelectThreads);
    select_rays_kernel<<<1, kSelectThreads, smem, (cudaStream_t)stream>>>(cdf, num_pixels, uniforms, max_draws, seed,
                                                                          step_state, num_rays, pixel_indices, status);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_gather_targets(const float* soft_masks, const int64_t* pixel_indices, const int64_t* gt_indices,
                        int num_rays, int num_instances, float* targets, void* stream) {
    VSRD_CHECK_ARG(soft_masks && pixel_indices && targets, "soft_masks / pixel_indices / targets must not be NULL");
    VSRD_CHECK_ARG(num_rays >= 0 && num_instances >= 1 && num_instances <= VSRD_MAX_INSTANCES, "bad sizes");
    const int total = num_rays * num_instances;
    if (total == 0) return 0;
    gather_targets_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(soft_masks, pixel_indices, gt_indices,
                                                                                 num_rays, num_instances, targets);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_soft_masks(const float* polygons, const int32_t* polygon_sizes, int num_views, int num_instances,
                    int max_vertices, int height, int width, float temperature, float* soft_masks, void* stream) {
    VSRD_CHECK_ARG(polygons && polygon_sizes && soft_masks, "polygons / polygon_sizes / soft_masks must not be NULL");
    VSRD_CHECK_ARG(num_views >= 1 && num_instances >= 1 && num_instances <= VSRD_MAX_INSTANCES, "bad view / instance count");
    VSRD_CHECK_ARG(max_vertices >= 3 && max_vertices <= 64, "max_vertices must be in [3, 64]");
    VSRD_CHECK_ARG(height > 0 && width > 0 && temperature > 0.0f, "bad image size / temperature");
    const size_t smem = sizeof(float) * (size_t)num_instances * max_vertices * 2;
    const int blocks = min((height * width + 255) / 256, 148 * 8);
    soft_masks_kernel<<<dim3(blocks, num_views), 256, smem, (cudaStream_t)stream>>>(
        polygons, polygon_sizes, num_views, num_instances, max_vertices, height, width, temperature, soft_masks);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_step_state_update(VsrdStepState* step_state, const VsrdSchedule* schedule, int64_t set_step, void* stream) {
    VSRD_CHECK_ARG(step_state && schedule, "step_state / schedule must not be NULL");
    VSRD_CHECK_ARG(schedule->num_steps >= 1, "num_steps must be positive");
    step_state_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_state, *schedule, set_step);
    VSRD_CHECK_LAUNCH();
    return 0;
}


int vsrd_project_box_3d(const float* boxes_3d, int num_boxes, const float* intrinsic_matrix, float epsilon,
                        float* boxes_2d, void* stream) {
    VSRD_CHECK_ARG(num_boxes >= 0, "num_boxes must be non-negative");
    if (num_boxes == 0) return 0;
    VSRD_CHECK_ARG(boxes_3d && intrinsic_matrix && boxes_2d, "NULL pointer");
    project_box_3d_kernel<<<(num_boxes + 63) / 64, 64, 0, (cudaStream_t)stream>>>(boxes_3d, intrinsic_matrix, num_boxes, epsilon, boxes_2d);
    VSRD_CHECK_LAUNCH();
    return 0;
}

int vsrd_project_box_3d_backward(const float* boxes_3d, int num_boxes, const float* intrinsic_matrix, float epsilon,
                                 const float* grad_boxes_2d, float* grad_boxes_3d, void* stream) {
    VSRD_CHECK_ARG(num_boxes >= 0, "num_boxes must be non-negative");
    if (num_boxes == 0) return 0;
    VSRD_CHECK_ARG(boxes_3d && intrinsic_matrix && grad_boxes_2d && grad_boxes_3d, "NULL pointer");
    project_box_3d_backward_kernel<<<(num_boxes + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
        boxes_3d, intrinsic_matrix, num_boxes, epsilon, grad_boxes_2d, grad_boxes_3d);
    VSRD_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
