// Round-2 groundwork: can the encoder pair two SMs on one weight stream?  (Run on B200 at the end of round 1: both
// checks pass, 256 / 256 rows each.)  A cluster of 2 CTAs issues
//   (1) one tcgen05.mma.cta_group::2 (M = 256 = 2 x 128 rows, N = 128, K = 16): every CTA supplies its own 128 A rows and
//       HALF of the B rows (64 of 128) -- i.e. each SM ingests half of a weight chunk;
//   (2) afterwards one tcgen05.mma.cta_group::1 per CTA into other TMEM columns (the attention MMAs need per-CTA B
//       operands): does mixing the two cta_group forms in one kernel work at run time?  (ptxas accepts it.)
// Expected print-out if both work: for CTA r, lane l: pair-MMA value (128 r + l + 1), solo-MMA value -(128 r + l + 1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/cg2_probe tools/cg2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity, int site) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) { printf("cg2_probe: barrier timeout at site %d (block %d)\n", site, (int)blockIdx.x); __trap(); }
    } while (!ok);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160) probe(float* out /*[2 ctas][128 lanes][2]*/, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* A = smem;                  // [128 rows x 64 k]
    uint8_t* Bh = smem + 16384;         // this CTA's HALF of B: [64 n x 64 k]
    uint8_t* Bfull = smem + 32768;      // a private full B [128 n x 64 k] for the cta_group::1 MMA
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmem_slot;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 64; i += blockDim.x) {
        const int r = i >> 6, k = i & 63;
        const uint32_t off = sw128_off(r, k >> 3) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(A + off) = __float2bfloat16((float)(128 * rank + r + 1));
        *reinterpret_cast<__nv_bfloat16*>(Bfull + off) = __float2bfloat16((r == k) ? -1.f : 0.f);
        if (r < 64) *reinterpret_cast<__nv_bfloat16*>(Bh + off) = __float2bfloat16((64 * (int)rank + r == k) ? 1.f : 0.f);
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {                    // the same logical warp in both CTAs, same smem slot address
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();                     // both CTAs' operand tiles and barriers are ready
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
    if (rank == 0 && warp == 4 && lane == 0) {
        // pair MMA: D[256 x 128] = [A_cta0; A_cta1] . [Bh_cta0; Bh_cta1]^T, issued by the leader CTA only
        const long long t0 = clock64();
        asm volatile("{.reg .pred p; setp.ne.b32 p, 0, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;}"
                     ::"r"(tmem), "l"(sw128_desc(smem_u32(A))), "l"(sw128_desc(smem_u32(Bh))), "r"(idesc(256, 128)) : "memory");
        // completion is multicast to the barrier at the same offset in both CTAs
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar0), "h"((uint16_t)3) : "memory");
        cyc[0] = clock64() - t0;
    }
    wait(bar0, 0, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 4 && lane == 0) {       // every CTA: a solo MMA into columns [128, 256) of its own TMEM
        asm volatile("{.reg .pred p; setp.ne.b32 p, 0, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                     ::"r"(tmem + 128), "l"(sw128_desc(smem_u32(A))), "l"(sw128_desc(smem_u32(Bfull))), "r"(idesc(128, 128)) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar1) : "memory");
    }
    wait(bar1, 0, 2);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        const uint32_t a = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t v0, v1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0) : "r"(a) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v1) : "r"(a + 128) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[(rank * 128 + warp * 32 + lane) * 2 + 0] = __uint_as_float(v0);
        out[(rank * 128 + warp * 32 + lane) * 2 + 1] = __uint_as_float(v1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 2 * 128 * 2 * sizeof(float)); cudaMalloc(&cyc, 4 * sizeof(long long));
    cudaMemset(out, 0, 2 * 128 * 2 * sizeof(float));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    probe<<<2, 160, 49152>>>(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    static float h[2 * 128 * 2];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    int ok_pair = 0, ok_solo = 0;
    for (int r = 0; r < 2; ++r)
        for (int l = 0; l < 128; ++l) {
            ok_pair += (h[(r * 128 + l) * 2] == (float)(128 * r + l + 1));
            ok_solo += (h[(r * 128 + l) * 2 + 1] == -(float)(128 * r + l + 1));
        }
    printf("cta_group::2 pair MMA (each CTA supplies half of B): %d / 256 rows correct\n", ok_pair);
    printf("cta_group::1 MMA issued afterwards in the same kernel: %d / 256 rows correct\n", ok_solo);
    for (int r = 0; r < 2; ++r) printf("cta %d lanes 0..3: pair %.0f %.0f %.0f %.0f   solo %.0f %.0f %.0f %.0f\n", r, h[(r * 128) * 2], h[(r * 128 + 1) * 2],
                                       h[(r * 128 + 2) * 2], h[(r * 128 + 3) * 2], h[(r * 128) * 2 + 1], h[(r * 128 + 1) * 2 + 1], h[(r * 128 + 2) * 2 + 1], h[(r * 128 + 3) * 2 + 1]);
    return 0;
}
