// Standalone probe of the tcgen05 building blocks in vsrd_b200/csrc/vsrd_umma.cuh (sm_100a):
//   T1  TS-mode MMA: A [128 x 16] written to TMEM by its owning threads, B [16 x 16] K-major in shared memory,
//       exact small-integer operands -> D must equal the host product bit for bit (validates idesc, smem descriptor,
//       TMEM lane/column mapping, commit -> mbarrier, tcgen05.ld)
//   T2  3xTF32 split on random fp32 operands vs an fp64 host product (the accuracy the field kernels rely on)
//   T3  SS-mode MMA with MN-major operands: D[f][o] = sum_s act[s][f] * adj[s][o]  (the weight-gradient contraction:
//       A = act^T with M padded to 128, B = adj, K = 128 samples), single-pass TF32 with RNA rounding and 3xTF32
//   T4  latency of one (st -> fence -> barrier -> 6 MMAs -> commit -> wait -> ld) round trip for one 128-thread group,
//       and the throughput with 3 groups per CTA
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/umma_probe tools/umma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>
#include <cstring>

#include "../vsrd_b200/csrc/vsrd_umma.cuh"

using namespace vsrd::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// B operand, K-major canonical layout for N rows x K columns: 16-byte chunk (k / 4) of row n at
//   (k / 4) * lbo + (n / 8) * 128 + (n % 8) * 16 bytes, lbo = (N / 8) * 128
__host__ __device__ inline int b_kmajor_index(int n, int k, int N) { return (k / 4) * (N / 8) * 32 + (n / 8) * 32 + (n % 8) * 4 + (k % 4); }
// MN-major canonical layout for F (MN) x S (K) operands: element (f, s) at (s / 8) * lbo + (f / 4) * 128 + (s % 8) * 16 + (f % 4) * 4 bytes
__host__ __device__ inline int mn_index(int f, int s, int blocks) { return (s / 8) * blocks * 32 + (f / 4) * 32 + (s % 8) * 4 + (f % 4); }

struct Params {
    const float* A;      // [128][16]
    const float* B;      // [16][16]  (row n, col k)
    float* D;            // [128][16]
    int split;           // 0: plain TF32 (operands as given), 1: 3xTF32
};

__global__ void __launch_bounds__(128) ts_kernel(Params p) {
    __shared__ __align__(128) float sB[2][16 * 16];      // hi, lo
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc<64>(&tmem_base_slot);
    if (tid == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
    for (int i = tid; i < 256; i += 128) {
        const int n = i / 16, k = i % 16;
        float hi, lo;
        if (p.split) split_tf32(p.B[i], hi, lo); else { hi = p.B[i]; lo = 0.0f; }
        sB[0][b_kmajor_index(n, k, 16)] = hi;
        sB[1][b_kmajor_index(n, k, 16)] = lo;
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
    // A: columns [0,16) hi, [16,32) lo; D: columns [32,48)
    float a[16], ahi[16], alo[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        a[k] = p.A[tid * 16 + k];
        if (p.split) split_tf32(a[k], ahi[k], alo[k]); else { ahi[k] = a[k]; alo[k] = 0.0f; }
    }
    tmem_st16(lane_base + 0, ahi);
    tmem_st16(lane_base + 16, alo);
    wait_st();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc_tf32(128, 16);
        const uint32_t lbo = 2 * 128, sbo = 128;
        bool acc = false;
        for (int ks = 0; ks < 2; ++ks) {
            const uint64_t bhi = make_smem_desc(smem_u32(sB[0]) + ks * 2 * lbo, lbo, sbo);
            const uint64_t blo = make_smem_desc(smem_u32(sB[1]) + ks * 2 * lbo, lbo, sbo);
            mma_tf32_ts(tmem + 32, tmem + 0 + ks * 8, bhi, idesc, acc); acc = true;
            if (p.split) {
                mma_tf32_ts(tmem + 32, tmem + 16 + ks * 8, bhi, idesc, true);
                mma_tf32_ts(tmem + 32, tmem + 0 + ks * 8, blo, idesc, true);
            }
        }
        mma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after_sync();
    float d[16];
    tmem_ld16(lane_base + 32, d);
    wait_ld();
#pragma unroll
    for (int n = 0; n < 16; ++n) p.D[tid * 16 + n] = d[n];
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<64>(tmem);
}

// T3a: SS mode, D[128 x 16] = A[128 x 16] B[16 x 16]^T with A / B stored K-major or MN-major (same logical matrices)
__global__ void __launch_bounds__(128) ss_kernel(Params p, int a_mn, int b_mn) {
    __shared__ __align__(128) float sA[128 * 16];
    __shared__ __align__(128) float sB[16 * 16];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc<32>(&tmem_base_slot);
    if (tid == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
    for (int i = tid; i < 128 * 16; i += 128) {
        const int m = i / 16, k = i % 16;
        sA[a_mn ? mn_index(m, k, 32) : b_kmajor_index(m, k, 128)] = p.A[i];
    }
    for (int i = tid; i < 256; i += 128) {
        const int n = i / 16, k = i % 16;
        sB[b_mn ? mn_index(n, k, 4) : b_kmajor_index(n, k, 16)] = p.B[i];
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    const uint32_t tmem = tmem_base_slot;
    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc_tf32(128, 16, a_mn, b_mn);
        for (int ks = 0; ks < 2; ++ks) {
            // K-major: k-step = 2 chunks of 4 -> advance 2 * LBO;  MN-major: k-step = one group of 8 rows -> advance LBO
            const uint64_t da = a_mn ? make_smem_desc(smem_u32(sA) + ks * 32 * 128, 32 * 128, 128)
                                     : make_smem_desc(smem_u32(sA) + ks * 2 * 16 * 128, 16 * 128, 128);
            const uint64_t db = b_mn ? make_smem_desc(smem_u32(sB) + ks * 4 * 128, 4 * 128, 128)
                                     : make_smem_desc(smem_u32(sB) + ks * 2 * 2 * 128, 2 * 128, 128);
            mma_tf32_ss(tmem, da, db, idesc, ks > 0);
        }
        mma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after_sync();
    float d[16];
    tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16), d);
    wait_ld();
#pragma unroll
    for (int n = 0; n < 16; ++n) p.D[tid * 16 + n] = d[n];
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<32>(tmem);
}

// T3: D[f][o] = sum_s act[s][f] * adj[s][o], f < F (<= 48, padded rows ignored), o < 16, s < 128
struct WgParams {
    const float* act;    // [128][F]
    const float* adj;    // [128][16]
    float* D;            // [F][16]
    int F;
    int passes;          // 1: TF32 (rna), 3: 3xTF32
};

__global__ void __launch_bounds__(128) wgrad_kernel(WgParams p) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int fblocks = (p.F + 3) / 4;
    // A operand (act^T, MN-major): [16 k-groups][fblocks][8][4], hi then lo; the MMA reads 32 MN blocks per k-group, so
    // each k-group is padded to 32 blocks (the rows >= F of D are garbage-in-garbage-out and never read)
    float* sAhi = smem;
    float* sAlo = sAhi + 16 * 32 * 32;
    float* sBhi = sAlo + 16 * 32 * 32;                   // adj, MN-major with 4 blocks: [16][4][8][4]
    float* sBlo = sBhi + 16 * 4 * 32;
    if (warp == 0) tmem_alloc<32>(&tmem_base_slot);
    if (tid == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
    for (int i = tid; i < 2 * 16 * 32 * 32; i += 128) smem[i] = 0.0f;
    __syncthreads();
    {
        const int s = tid;
        for (int f = 0; f < p.F; ++f) {
            const float v = p.act[s * p.F + f];
            float hi, lo;
            if (p.passes == 3) split_tf32(v, hi, lo);
            else { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); hi = __uint_as_float(r); lo = 0.0f; }
            sAhi[mn_index(f, s, 32)] = hi;
            sAlo[mn_index(f, s, 32)] = lo;
        }
        for (int o = 0; o < 16; ++o) {
            const float v = p.adj[s * 16 + o];
            float hi, lo;
            if (p.passes == 3) split_tf32(v, hi, lo);
            else { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); hi = __uint_as_float(r); lo = 0.0f; }
            sBhi[mn_index(o, s, 4)] = hi;
            sBlo[mn_index(o, s, 4)] = lo;
        }
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    const uint32_t tmem = tmem_base_slot;
    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc_tf32(128, 16, true, true);
        bool acc = false;
        for (int kg = 0; kg < 16; ++kg) {
            const uint64_t ahi = make_smem_desc(smem_u32(sAhi) + kg * 32 * 128, 32 * 128, 128);
            const uint64_t alo = make_smem_desc(smem_u32(sAlo) + kg * 32 * 128, 32 * 128, 128);
            const uint64_t bhi = make_smem_desc(smem_u32(sBhi) + kg * 4 * 128, 4 * 128, 128);
            const uint64_t blo = make_smem_desc(smem_u32(sBlo) + kg * 4 * 128, 4 * 128, 128);
            mma_tf32_ss(tmem, ahi, bhi, idesc, acc); acc = true;
            if (p.passes == 3) {
                mma_tf32_ss(tmem, alo, bhi, idesc, true);
                mma_tf32_ss(tmem, ahi, blo, idesc, true);
            }
        }
        mma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after_sync();
    float d[16];
    tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16), d);
    wait_ld();
    if (tid < p.F)
        for (int o = 0; o < 16; ++o) p.D[tid * 16 + o] = d[o];
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<32>(tmem);
}

// T5: the weight-gradient contraction as shipped: bf16 operands staged K-major SWIZZLE_128B with the SAMPLES along K
//     (thread s writes element (row f, column s) of both operands), D[f][o] = sum_s act[s][f] * adj[s][o], fp32 accumulate.
//     A: F <= 24 rows (3 atoms of 8 rows per 64-sample k-block; the MMA reads 16 atoms = garbage rows beyond), B: 16 rows.
__device__ __forceinline__ uint32_t sw128_offset(int row, int sample) {      // bytes inside one k-block region of `atoms` atoms
    const int sp = sample & 63, chunk = sp >> 3, e = sp & 7;
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((chunk ^ (row & 7)) << 4)) + e * 2);
}
__global__ void __launch_bounds__(128) wgrad_bf16_kernel(WgParams p) {
    extern __shared__ __align__(1024) unsigned char smem_b[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // layout: A [2 k-blocks][3 atoms][1024], then B [2][2][1024], then 32 KB of slack for the A window
    unsigned char* sA = smem_b;
    unsigned char* sB = smem_b + 2 * 3 * 1024;
    if (warp == 0) tmem_alloc<32>(&tmem_base_slot);
    if (tid == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
    for (int i = tid; i < (2 * 3 * 1024 + 2 * 2 * 1024 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_b)[i] = 0x7fc07fc0u;  // bf16 NaNs
    __syncthreads();
    {
        const int s = tid, kb = s >> 6;
        for (int f = 0; f < p.F; ++f)
            *reinterpret_cast<__nv_bfloat16*>(sA + kb * 3 * 1024 + sw128_offset(f, s)) = __float2bfloat16_rn(p.act[s * p.F + f]);
        for (int o = 0; o < 16; ++o)
            *reinterpret_cast<__nv_bfloat16*>(sB + kb * 2 * 1024 + sw128_offset(o, s)) = __float2bfloat16_rn(p.adj[s * 16 + o]);
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    const uint32_t tmem = tmem_base_slot;
    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc_bf16(128, 16);
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(sA) + (ks >> 2) * 3 * 1024 + (ks & 3) * 32, 1024);
            const uint64_t db = make_smem_desc_sw128(smem_u32(sB) + (ks >> 2) * 2 * 1024 + (ks & 3) * 32, 1024);
            mma_bf16_ss(tmem, da, db, idesc, ks > 0);
        }
        mma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after_sync();
    float d[16];
    tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16), d);
    wait_ld();
    if (tid < p.F)
        for (int o = 0; o < 16; ++o) p.D[tid * 16 + o] = d[o];
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<32>(tmem);
}

// T4: G groups of 128 threads, each looping `iters` round trips of a 16 -> 16 layer (6 MMAs per round trip)
template <int G>
__global__ void __launch_bounds__(128 * G) latency_kernel(int iters, float* out, long long* cycles) {
    __shared__ __align__(128) float sB[2][256];
    __shared__ uint64_t mbar[G];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, group = warp >> 2, wq = warp & 3;
    if (warp == 0) tmem_alloc<256>(&tmem_base_slot);
    if (tid == 0) { for (int g = 0; g < G; ++g) mbar_init(&mbar[g], 1); mbar_fence_init(); }
    for (int i = tid; i < 256; i += 128 * G) {
        sB[0][b_kmajor_index(i / 16, i % 16, 16)] = (i / 16 == i % 16) ? 1.0f : 0.0f;     // identity
        sB[1][b_kmajor_index(i / 16, i % 16, 16)] = 0.0f;
    }
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_slot + group * 48;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * wq) << 16);
    const uint32_t idesc = make_idesc_tf32(128, 16);
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (float)(tid + k) * 0.001f;
    uint32_t parity = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float hi[16], lo[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) split_tf32(v[k] * 1.0001f, hi[k], lo[k]);
        tmem_st16(lane_base + 0, hi);
        tmem_st16(lane_base + 16, lo);
        wait_st();
        fence_before_sync();
        named_barrier(1 + group, 128);
        if (wq == 0 && (tid & 31) == 0) {
            fence_after_sync();
            for (int ks = 0; ks < 2; ++ks) {
                const uint64_t bhi = make_smem_desc(smem_u32(sB[0]) + ks * 512, 256, 128);
                const uint64_t blo = make_smem_desc(smem_u32(sB[1]) + ks * 512, 256, 128);
                mma_tf32_ts(tmem + 32, tmem + ks * 8, bhi, idesc, ks > 0);
                mma_tf32_ts(tmem + 32, tmem + 16 + ks * 8, bhi, idesc, true);
                mma_tf32_ts(tmem + 32, tmem + ks * 8, blo, idesc, true);
            }
            mma_commit(&mbar[group]);
        }
        mbar_wait(&mbar[group], parity);
        parity ^= 1;
        fence_after_sync();
        tmem_ld16(lane_base + 32, v);
        wait_ld();
    }
    const long long t1 = clock64();
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k];
    out[blockIdx.x * 128 * G + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<256>(tmem_base_slot);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, 128 * 48 * 4)); CK(cudaMalloc(&dB, 128 * 16 * 4)); CK(cudaMalloc(&dD, 128 * 48 * 4));
    std::vector<float> A(128 * 16), B(256), D(128 * 16);
    // ---- T1: exact operands
    for (int i = 0; i < 128 * 16; ++i) A[i] = (float)((i * 7 + 3) % 17 - 8) * 0.125f;
    for (int i = 0; i < 256; ++i) B[i] = (float)((i * 5 + 1) % 13 - 6) * 0.25f;
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    ts_kernel<<<1, 128>>>(Params{dA, dB, dD, 0});
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0.0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
            double ref = 0.0;
            for (int k = 0; k < 16; ++k) ref += (double)A[m * 16 + k] * B[n * 16 + k];
            worst = fmax(worst, fabs(ref - D[m * 16 + n]));
        }
    printf("T1 TS-mode exact operands: max |D - A B^T| = %.3e  %s\n", worst, worst == 0.0 ? "OK" : "MISMATCH");
    if (worst != 0.0) {
        printf("   D[0][0..7] ="); for (int n = 0; n < 8; ++n) printf(" %g", D[n]); printf("\n   want     =");
        for (int n = 0; n < 8; ++n) { double ref = 0; for (int k = 0; k < 16; ++k) ref += (double)A[k] * B[n * 16 + k]; printf(" %g", ref); }
        printf("\n");
    }
    // ---- T2: 3xTF32 accuracy
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2 - 1;
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int split = 0; split < 2; ++split) {
        ts_kernel<<<1, 128>>>(Params{dA, dB, dD, split});
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double e = 0.0, et = 0.0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
                double ref = 0.0, reft = 0.0;
                for (int k = 0; k < 16; ++k) { ref += (double)A[m * 16 + k] * B[n * 16 + k]; reft += (double)tf32_trunc(A[m * 16 + k]) * tf32_trunc(B[n * 16 + k]); }
                e = fmax(e, fabs(ref - D[m * 16 + n]));
                et = fmax(et, fabs(reft - D[m * 16 + n]));
            }
        printf("T2 %s: max abs error vs fp64 = %.3e (vs truncated-operand product %.3e)\n", split ? "3xTF32" : "1xTF32 (raw fp32 operands)", e, et);
    }
    // ---- T3a: SS mode, all four major combinations, exact operands
    for (int i = 0; i < 128 * 16; ++i) A[i] = (float)((i * 7 + 3) % 17 - 8) * 0.125f;
    for (int i = 0; i < 256; ++i) B[i] = (float)((i * 5 + 1) % 13 - 6) * 0.25f;
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int combo = 0; combo < 4; ++combo) {
        CK(cudaMemset(dD, 0xff, 128 * 16 * 4));
        ss_kernel<<<1, 128>>>(Params{dA, dB, dD, 0}, combo & 1, combo >> 1);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double w = 0.0; int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
                double ref = 0.0;
                for (int k = 0; k < 16; ++k) ref += (double)A[m * 16 + k] * B[n * 16 + k];
                const double err = fabs(ref - D[m * 16 + n]);
                if (!(err == 0.0)) ++bad;
                if (err > w) w = err;
            }
        printf("T3a SS-mode A %s-major, B %s-major: max err %.3e, %d / 2048 wrong; D[0][0..3] = %g %g %g %g, D[1][0] = %g\n", (combo & 1) ? "MN" : "K",
               (combo >> 1) ? "MN" : "K", w, bad, D[0], D[1], D[2], D[3], D[16]);
    }
    {
        double r0 = 0, r1 = 0, r16 = 0;
        for (int k = 0; k < 16; ++k) { r0 += (double)A[k] * B[k]; r1 += (double)A[k] * B[16 + k]; r16 += (double)A[16 + k] * B[k]; }
        printf("    want D[0][0] = %g, D[0][1] = %g, D[1][0] = %g\n", r0, r1, r16);
    }
    // ---- T3: weight-gradient contraction
    for (int F : {16, 17, 48}) {
        std::vector<float> act(128 * F), adj(128 * 16), W(F * 16);
        for (auto& v : act) v = (float)rand() / RAND_MAX * 2 - 1;
        for (auto& v : adj) v = (float)rand() / RAND_MAX * 2 - 1;
        CK(cudaMemcpy(dA, act.data(), act.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, adj.data(), adj.size() * 4, cudaMemcpyHostToDevice));
        const size_t smem = (2 * 16 * 32 * 32 + 2 * 16 * 4 * 32) * 4;
        CK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int passes : {1, 3}) {
            wgrad_kernel<<<1, 128, smem>>>(WgParams{dA, dB, dD, F, passes});
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(W.data(), dD, W.size() * 4, cudaMemcpyDeviceToHost));
            double e = 0.0, scale = 0.0;
            for (int f = 0; f < F; ++f)
                for (int o = 0; o < 16; ++o) {
                    double ref = 0.0;
                    for (int s = 0; s < 128; ++s) ref += (double)act[s * F + f] * adj[s * 16 + o];
                    e = fmax(e, fabs(ref - W[f * 16 + o]));
                    scale = fmax(scale, fabs(ref));
                }
            printf("T3 weight gradient F=%d, %dxTF32 (SS, MN-major): max abs error %.3e (max |ref| %.2f)\n", F, passes, e, scale);
        }
    }
    // ---- T5: bf16 SWIZZLE_128B K-major weight gradient
    for (int F : {16, 17, 24}) {
        std::vector<float> act(128 * F), adj(128 * 16), W(F * 16);
        for (auto& v : act) v = (float)rand() / RAND_MAX * 2 - 1;
        for (auto& v : adj) v = (float)rand() / RAND_MAX * 2 - 1;
        CK(cudaMemcpy(dA, act.data(), act.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, adj.data(), adj.size() * 4, cudaMemcpyHostToDevice));
        const size_t smem = 2 * 3 * 1024 + 2 * 2 * 1024 + 32768 + 1024;
        CK(cudaFuncSetAttribute(wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wgrad_bf16_kernel<<<1, 128, smem>>>(WgParams{dA, dB, dD, F, 1});
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(W.data(), dD, W.size() * 4, cudaMemcpyDeviceToHost));
        auto bf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u; memcpy(&x, &u, 4); return x; };
        double e = 0.0, eb = 0.0, scale = 0.0;
        for (int f = 0; f < F; ++f)
            for (int o = 0; o < 16; ++o) {
                double ref = 0.0, refb = 0.0;
                for (int s = 0; s < 128; ++s) { ref += (double)act[s * F + f] * adj[s * 16 + o]; refb += (double)bf(act[s * F + f]) * bf(adj[s * 16 + o]); }
                e = fmax(e, fabs(ref - W[f * 16 + o]));
                eb = fmax(eb, fabs(refb - W[f * 16 + o]));
                scale = fmax(scale, fabs(ref));
            }
        printf("T5 weight gradient F=%d, bf16 SW128 K-major SS: max abs error vs fp64 %.3e, vs bf16-rounded operands %.3e (max |ref| %.2f)\n", F, e, eb, scale);
    }
    // ---- T4: latency / throughput
    float* dout; long long* dcyc;
    CK(cudaMalloc(&dout, 148 * 512 * 4)); CK(cudaMalloc(&dcyc, 148 * 8));
    const int iters = 2000;
    long long cyc;
    latency_kernel<1><<<1, 128>>>(iters, dout, dcyc); CK(cudaDeviceSynchronize());
    latency_kernel<1><<<1, 128>>>(iters, dout, dcyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("T4 1 group  : %.1f cycles per round trip (st x2 + barrier + 6 MMA + commit + wait + ld)\n", (double)cyc / iters);
    latency_kernel<3><<<148, 384>>>(iters, dout, dcyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("T4 3 groups : %.1f cycles per round trip per group (%.1f per tile-layer per SM), 148 CTAs\n", (double)cyc / iters, (double)cyc / iters / 3);
    latency_kernel<4><<<148, 512>>>(iters, dout, dcyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("T4 4 groups : %.1f cycles per round trip per group (%.1f per tile-layer per SM), 148 CTAs\n", (double)cyc / iters, (double)cyc / iters / 4);
    return 0;
}
