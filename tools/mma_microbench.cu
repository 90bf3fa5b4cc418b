// Development aid: cycles per tcgen05.mma for the shapes the bf16 encoder uses (one CTA, one issuer).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_microbench tools/mma_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_mask(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t m0) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, 0, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %5, %5, %5}, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(m0), "r"(0xFFFFFFFFu) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok);
}

__global__ void __launch_bounds__(128, 1) bench(long long* out, int extra_smem_traffic) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    const uint32_t sb = smem_u32(smem);
    const uint64_t dA = sw128_desc(sb), dB = sw128_desc(sb + 65536);
    if (warp == 0 && lane == 0) {
        uint32_t parity = 0;
        const int REP = 64;
        for (int test = 0; test < 8; ++test) {
            long long t0 = clock64();
            for (int i = 0; i < REP; ++i) {
                const uint64_t ko = (uint64_t)((i & 3) * 2) + (uint64_t)((i >> 2) & 3) * 1024;
                switch (test) {
                    case 0: mma_ss(tmem, dA + ko, dB + ko, idesc(128, 128), i > 0); break;                 // N=128 SS, same D
                    case 1: mma_ss(tmem + (i & 1) * 128, dA + ko, dB + ko, idesc(128, 128), i > 1); break;   // N=128 SS, alternate D
                    case 2: mma_ss(tmem, dA + ko, dB + ko, idesc(128, 256), i > 0); break;                  // N=256 SS
                    case 3: mma_ts(tmem, tmem + 448 + (i & 3) * 8, dB + ko, idesc(128, 128), i > 0); break;  // N=128 TS
                    case 4: mma_ss(tmem + (i & 7) * 16, dA + ko, dB + ko, idesc(128, 16), 0); break;         // N=16 SS
                    case 5: mma_ts(tmem + (i & 7) * 16, tmem + 448 + (i & 3) * 8, dB + ko, idesc(128, 16), 0); break;  // N=16 TS
                    case 6: mma_ss_mask(tmem + (i & 7) * 16, dA + ko, dB + ko, idesc(128, 16), 0xFFFF0000u); break;     // N=16 SS masked
                    case 7: mma_ss(tmem + (i & 3) * 32, dA + ko, dB + ko, idesc(128, 32), 0); break;         // N=32 SS
                }
            }
            long long t1 = clock64();
            commit(smem_u32(&bar));
            wait(smem_u32(&bar), parity); parity ^= 1;
            long long t2 = clock64();
            out[test * 2] = (t1 - t0); out[test * 2 + 1] = (t2 - t0);
        }
    } else if (extra_smem_traffic && warp >= 2) {
        // competing shared-memory traffic (what the epilogue warps generate)
        uint4* p = reinterpret_cast<uint4*>(smem + 131072);
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int it = 0; it < 20000; ++it) {
            uint4 v;
            const uint32_t a0 = smem_u32(p + ((it * 64 + threadIdx.x) & 2047)), a1 = smem_u32(p + ((it * 64 + threadIdx.x + 7) & 2047));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a0) : "memory");
            acc.x += v.x;
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(acc.x), "r"(acc.y), "r"(acc.z), "r"(acc.w) : "memory");
        }
        if (acc.x == 12345) out[31] = acc.x;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    long long* d; cudaMalloc(&d, 64 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
    const char* names[8] = {"N=128 SS same D", "N=128 SS alt D", "N=256 SS", "N=128 TS", "N=16 SS", "N=16 TS", "N=16 SS masked", "N=32 SS"};
    for (int traffic = 0; traffic < 2; ++traffic) {
        for (int rep = 0; rep < 2; ++rep) { bench<<<1, 128, 196608>>>(d, traffic); cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return 1; } }
        long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("--- competing smem traffic: %d\n", traffic);
        for (int t = 0; t < 8; ++t) printf("%-18s issue %6.1f cyc/MMA   issue+complete %6.1f cyc/MMA\n", names[t], h[2 * t] / 64.0, h[2 * t + 1] / 64.0);
    }
    return 0;
}
