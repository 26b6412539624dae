// Micro-benchmark of the sm_100a issue rates the field kernels are designed around:
// legacy mma.sync (TF32 m16n8k8, BF16 m16n8k16), movmatrix, shfl, FFMA, MUFU, and MMA+FFMA overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/pipe_rates tools/pipe_rates.cu
// Prints warp-instructions per cycle per SM for 4/8/16 resident warps per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MODE 0 tf32, 1 bf16, 2 movmatrix, 3 shfl, 4 ffma, 5 mufu ex2, 6 tf32 + 8 ffma per mma, 7 bf16 + 8 ffma per mma,
// 8 f16, 9 tf32 + 16 ffma per mma
template <int MODE>
__global__ void rate_kernel(float* out, long long* cycles, int seedi) {
    float acc[8][4];
    uint32_t a[4], b0, b1;
    for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;
    for (int q = 0; q < 4; ++q) a[q] = 0x3f800000u + threadIdx.x * 3 + q + seedi;
    b0 = 0x3f000000u + threadIdx.x + seedi; b1 = b0 + 7;
    float f[16];
    for (int i = 0; i < 16; ++i) f[i] = 1.0f + i * 0.001f + seedi;
    uint32_t m = threadIdx.x * 2654435761u + seedi;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0 || MODE == 6 || MODE == 9) {
#pragma unroll
            for (int i = 0; i < 8; ++i) mma_tf32(acc[i], a, b0, b1);
        }
        if (MODE == 1 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; ++i) mma_bf16(acc[i], a, b0, b1);
        }
        if (MODE == 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) mma_f16(acc[i], a, b0, b1);
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %0;" : "+r"(a[i & 3]));
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __shfl_xor_sync(0xffffffffu, f[i], 1 + (i & 1));
        }
        if (MODE == 4 || MODE == 6 || MODE == 7 || MODE == 9) {
            const int reps = (MODE == 4) ? 1 : (MODE == 9 ? 8 : 4);
#pragma unroll
            for (int r = 0; r < reps; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], 1.0001f, 0.5f);
        }
        if (MODE == 10 || MODE == 11) {       // packed fp32 pairs (Blackwell FFMA2): 16 independent chains
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, c1 = 0x3f8003473f800347ull, c2 = 0x3f0000003f000000ull;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(f[i]), "f"(f[i + 1]));
#pragma unroll
                for (int r = 0; r < 8; ++r) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(c1), "l"(c2));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f[i]), "=f"(f[i + 1]) : "l"(v));
            }
        }
        if (MODE == 11) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], 1.0001f, 0.5f);
        }
        if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        }
    }
    const long long t1 = clock64();
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) s += acc[i][q];
    for (int i = 0; i < 16; ++i) s += f[i];
    s += __uint_as_float(a[0] ^ a[1] ^ a[2] ^ a[3] ^ m);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double warp_instr_per_iter, float* out, long long* cyc, int sms) {
    for (int warps_per_sm : {4, 8, 16, 32}) {
        const int threads = 128;
        const int ctas_per_sm = warps_per_sm / 4;
        const int grid = sms * ctas_per_sm;
        rate_kernel<MODE><<<grid, threads>>>(out, cyc, 0);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        rate_kernel<MODE><<<grid, threads>>>(out, cyc, 1);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        long long h[4096];
        cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        double mean = 0.0;
        for (int i = 0; i < grid; ++i) mean += (double)h[i];
        mean /= grid;
        const double instr_per_sm = warp_instr_per_iter * ITERS * warps_per_sm;
        printf("%-28s warps/SM=%2d  cycles=%10.0f  warp-instr/cycle/SM=%7.3f  cycles/instr/SMSP=%7.3f  (%.3f ms)\n",
               name, warps_per_sm, mean, instr_per_sm / mean, mean / (instr_per_sm / 4.0), ms);
    }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 128);
    cudaMalloc(&cyc, sizeof(long long) * 4096);
    run<0>("mma.m16n8k8.tf32", 8, out, cyc, sms);
    run<1>("mma.m16n8k16.bf16", 8, out, cyc, sms);
    run<8>("mma.m16n8k16.f16", 8, out, cyc, sms);
    run<2>("movmatrix.trans.b16", 8, out, cyc, sms);
    run<3>("shfl.xor", 8, out, cyc, sms);
    run<4>("ffma", 16, out, cyc, sms);
    run<5>("mufu.ex2", 8, out, cyc, sms);
    run<10>("ffma2 (fma.rn.f32x2)", 64, out, cyc, sms);
    run<11>("64 ffma2 + 64 ffma (all)", 128, out, cyc, sms);
    run<6>("tf32 mma + 8 ffma/mma (mma)", 8, out, cyc, sms);
    run<9>("tf32 mma + 16 ffma/mma (mma)", 8, out, cyc, sms);
    run<7>("bf16 mma + 8 ffma/mma (mma)", 8, out, cyc, sms);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
