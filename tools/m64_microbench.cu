// Development aid for the round-2 encoder plan: can TWO M = 64 accumulators share the same TMEM
// columns (one in lanes 0-15 of every 32-lane quarter, the other in lanes 16-31 -- cute's TmemAllocMode::Interleaved),
// and what does an M = 64 MMA cost next to an M = 128 one?
//   * functional: A0 = rows filled with (row + 1), A1 = rows filled with -(row + 1), B = identity-like (B[n][k] = (n == k)),
//     D0 = A0 B^T at TMEM lane offset 0, D1 = A1 B^T at lane offset 16; all 128 lanes x 64 columns are read back and the host
//     prints which lane holds which row of which tile;
//   * timing: 64 back-to-back MMAs, M = 128 N = 128 vs M = 64 N = 128 (SS, K = 16), issue + completion cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/m64_microbench tools/m64_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) { printf("m64_microbench: barrier timeout\n"); __trap(); }
    } while (!ok);
}

__global__ void __launch_bounds__(160) probe(float* out /*[128 lanes][64 cols]*/, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* A0 = smem;                 // [128 rows x 64 k] chunk (only rows 0..63 used by the M = 64 MMAs)
    uint8_t* A1 = smem + 16384;
    uint8_t* B = smem + 32768;          // [128 n x 64 k]
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 64; i += blockDim.x) {
        const int r = i >> 6, k = i & 63;
        const uint32_t off = sw128_off(r, k >> 3) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(A0 + off) = __float2bfloat16((float)(r + 1));
        *reinterpret_cast<__nv_bfloat16*>(A1 + off) = __float2bfloat16(-(float)(r + 1));
        *reinterpret_cast<__nv_bfloat16*>(B + off) = __float2bfloat16((r == k) ? 1.f : 0.f);
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
    if (warp < 4) {       // zero the 64 probe columns of every lane first
        uint32_t z[32];
        for (int i = 0; i < 32; ++i) z[i] = 0u;
        const uint32_t a = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 64; c += 32)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                         ::"r"(a + c), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]), "r"(z[8]), "r"(z[9]), "r"(z[10]), "r"(z[11]),
                           "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]), "r"(z[16]), "r"(z[17]), "r"(z[18]), "r"(z[19]), "r"(z[20]), "r"(z[21]), "r"(z[22]), "r"(z[23]),
                           "r"(z[24]), "r"(z[25]), "r"(z[26]), "r"(z[27]), "r"(z[28]), "r"(z[29]), "r"(z[30]), "r"(z[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 4 && lane == 0) {
        const uint64_t dA0 = sw128_desc(smem_u32(A0)), dA1 = sw128_desc(smem_u32(A1)), dB = sw128_desc(smem_u32(B));
        // functional probe: N = 64 (columns 0..63), K = 16: D[r][n] = sum_k A[r][k] B[n][k] = A[r][n] for n < 16, else 0
        mma_ss(tmem, dA0, dB, idesc(64, 64), 0);                          // tile 0 at lane offset 0
        mma_ss(tmem + (16u << 16), dA1, dB, idesc(64, 64), 0);            // tile 1 at lane offset 16: the interleaved slot
        commit(bar0);
        wait(bar0, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // timing: columns 128.. so the probe columns stay intact
        for (int M : {128, 64}) {
            const long long t0 = clock64();
            for (int i = 0; i < 64; ++i) mma_ss(tmem + 128 + (i & 1) * 128, dA0 + (uint64_t)((i & 3) * 2), dB + (uint64_t)((i & 3) * 2), idesc(M, 128), 1);
            const long long t1 = clock64();
            commit(bar1);
            wait(bar1, M == 128 ? 0 : 1);
            const long long t2 = clock64();
            cyc[M == 128 ? 0 : 2] = t1 - t0; cyc[M == 128 ? 1 : 3] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        uint32_t v[32];
        const uint32_t a = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 64; c += 32) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                           "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(a + c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(v[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 128 * 64 * sizeof(float)); cudaMalloc(&cyc, 4 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    for (int rep = 0; rep < 2; ++rep) {
        probe<<<1, 160, 49152>>>(out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    static float h[128 * 64]; long long c[4];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("lane -> value in column 0 (tile 0 rows are +(row+1), tile 1 rows are -(row+1), 0 = untouched)\n");
    for (int l = 0; l < 128; ++l) printf("%s%4d:%6.0f", (l % 8 == 0) ? "\n" : "  ", l, h[l * 64 + 0]);
    printf("\ncolumn check on lane 0: ");
    for (int n = 0; n < 20; ++n) printf("%.0f ", h[n]);
    printf("\n64 MMAs N=128 K=16 SS:  M=128 issue %lld cyc, issue+complete %lld cyc (%.1f per MMA);  M=64 issue %lld, issue+complete %lld (%.1f per MMA)\n",
           c[0], c[1], c[1] / 64.0, c[2], c[3], c[3] / 64.0);
    return 0;
}
