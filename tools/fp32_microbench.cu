// Development aid: issue throughput of the FP32 forms the encoder epilogue is made of -- scalar FFMA / FADD /
// FMUL against the packed f32x2 forms (FFMA2 / FADD2 / FMUL2) and MUFU.EX2 -- with the epilogue's occupancy
// (8 warps on one SM = 2 per SMSP) and with 16 warps.  Prints FP32 lane-operations per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fp32_microbench tools/fp32_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { T_FFMA, T_FADD, T_FMUL, T_FFMA2, T_FADD2, T_FMUL2, T_EX2, T_FFMA2_EX2, T_COUNT };
constexpr int ITERS = 512, ILP = 8;

__device__ __forceinline__ uint64_t pk(float a, float b) { return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32); }

template <int TEST>
__global__ void bench(float* out, long long* cycles, float seed) {
    float a[ILP * 2];
    uint64_t p[ILP];
#pragma unroll
    for (int i = 0; i < ILP * 2; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < ILP; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const float m = 1.0001f, c = 1e-4f;
    const uint64_t m2 = pk(m, m), c2 = pk(c, c);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (TEST == T_FFMA) { a[2 * i] = fmaf(a[2 * i], m, c); a[2 * i + 1] = fmaf(a[2 * i + 1], m, c); }
            if (TEST == T_FADD) { a[2 * i] = a[2 * i] + c; a[2 * i + 1] = a[2 * i + 1] + c; }
            if (TEST == T_FMUL) { a[2 * i] = a[2 * i] * m; a[2 * i + 1] = a[2 * i + 1] * m; }
            if (TEST == T_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
            if (TEST == T_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
            if (TEST == T_FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(m2));
            if (TEST == T_EX2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i + 1])); }
            if (TEST == T_FFMA2_EX2) {     // softmax-like mix: one packed FMA per two exponentials
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i + 1]));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP * 2; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(float)); cudaMalloc(&cyc, 8 * sizeof(long long));
    const char* names[T_COUNT] = {"FFMA", "FADD", "FMUL", "FFMA2 (fma.rn.f32x2)", "FADD2 (add.rn.f32x2)", "FMUL2 (mul.rn.f32x2)", "MUFU.EX2", "FFMA2 + 2 EX2"};
    for (int threads : {256, 512}) {
        printf("---- %d threads (%d warps per SMSP) on one SM; lane-ops per clock per SM\n", threads, threads / 128);
        for (int t = 0; t < T_COUNT; ++t) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (t) {
                    case T_FFMA: bench<T_FFMA><<<1, threads>>>(out, cyc, 1.f); break;
                    case T_FADD: bench<T_FADD><<<1, threads>>>(out, cyc, 1.f); break;
                    case T_FMUL: bench<T_FMUL><<<1, threads>>>(out, cyc, 1.f); break;
                    case T_FFMA2: bench<T_FFMA2><<<1, threads>>>(out, cyc, 1.f); break;
                    case T_FADD2: bench<T_FADD2><<<1, threads>>>(out, cyc, 1.f); break;
                    case T_FMUL2: bench<T_FMUL2><<<1, threads>>>(out, cyc, 1.f); break;
                    case T_EX2: bench<T_EX2><<<1, threads>>>(out, cyc, -1.f); break;
                    case T_FFMA2_EX2: bench<T_FFMA2_EX2><<<1, threads>>>(out, cyc, -1.f); break;
                }
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
            }
            long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            const double ops = (double)ITERS * ILP * 2 * threads * (t == T_FFMA2_EX2 ? 2 : 1);
            printf("%-24s %8lld cycles   %6.1f lane-ops/clk/SM\n", names[t], h, ops / (double)h);
        }
    }
    return 0;
}
