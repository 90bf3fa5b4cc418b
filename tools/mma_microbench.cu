// Development aid: cycle costs of the tcgen05 / mbarrier primitives the bf16 encoder is built from
// (one CTA; the issuing warp stays converged and issues under elect.sync, like the kernel does).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_microbench tools/mma_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_mask(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t m0) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, 0, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %5, %5, %5}, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(m0), "r"(0xFFFFFFFFu) : "memory");
}
__device__ __forceinline__ void mma_ts_mask(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t m0) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, 0, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %5, %5, %5}, p;}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(m0), "r"(0xFFFFFFFFu) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr int NT_TESTS = 16;
constexpr int REP = 64;

// test ids
enum { T_SS128, T_SS256, T_TS128, T_SS16M, T_TS16M, T_SS32M, T_COMMIT, T_TRYWAIT, T_CHUNK4_COMMIT, T_RING_MMA, T_RING_EMPTY, T_PINGPONG, T_LDTM, T_LDTM_ST };

template <int TEST>
__device__ __forceinline__ void issue_loop(uint32_t tmem, uint64_t dA, uint64_t dB, bool leader, uint32_t bars) {
#pragma unroll 8
    for (int i = 0; i < REP; ++i) {
        const uint64_t ko = (uint64_t)((i & 3) * 2) + (uint64_t)((i >> 2) & 3) * 1024;
        if (TEST == T_SS128) { if (leader) mma_ss(tmem + (i & 1) * 128, dA + ko, dB + ko, idesc(128, 128), 1); }
        if (TEST == T_SS256) { if (leader) mma_ss(tmem, dA + ko, dB + ko, idesc(128, 256), 1); }
        if (TEST == T_TS128) { if (leader) mma_ts(tmem + (i & 1) * 128, tmem + 448 + (i & 3) * 8, dB + ko, idesc(128, 128), 1); }
        if (TEST == T_SS16M) { if (leader) mma_ss_mask(tmem + (i & 7) * 16, dA + ko, dB + ko, idesc(128, 16), 0xFFFF0000u); }
        if (TEST == T_TS16M) { if (leader) mma_ts_mask(tmem + (i & 7) * 16, tmem + 448 + (i & 7) * 8, dB + ko, idesc(128, 16), 0xFFFF0000u); }
        if (TEST == T_SS32M) { if (leader) mma_ss_mask(tmem + (i & 3) * 32, dA + ko, dB + ko, idesc(128, 32), 0xFFFF0000u); }
        if (TEST == T_COMMIT) { if (leader) commit(bars + 8 * (8 + (i & 7))); }
        if (TEST == T_CHUNK4_COMMIT) {
            if (leader) {
                mma_ss(tmem, dA + ko, dB + ko, idesc(128, 128), 1);
                if ((i & 3) == 3) commit(bars + 8 * (8 + ((i >> 2) & 7)));
            }
        }
    }
}

__global__ void __launch_bounds__(160, 1) bench(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[32];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    const uint32_t b0 = smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0 + 8 * i), "r"(i >= 8 && i < 16 ? 1 << 20 : 1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    fence_after();
    const uint32_t tmem = tslot;
    const uint32_t sb = smem_u32(smem);
    const uint64_t dA = sw128_desc(sb), dB = sw128_desc(sb + 65536);
    // barrier use: [0] = drain barrier (count 1); [8..15] = sinks (huge count, never complete);
    // [16..19] ring full, [20..23] ring empty (count 1); [24],[25] ping-pong
    uint32_t drain_parity = 0;
    auto drain = [&](bool leader) { if (leader) commit(b0); wait(b0, drain_parity); drain_parity ^= 1; fence_after(); };
    if (warp == 0) {
        const bool leader = elect_one();
        long long t0, t1, t2;
#define RUN(TEST)                                                     \
        drain(leader); __syncwarp(); t0 = clock64();                  \
        issue_loop<TEST>(tmem, dA, dB, leader, b0); __syncwarp();     \
        t1 = clock64(); drain(leader); t2 = clock64();                \
        if (lane == 0) { out[2 * TEST] = t1 - t0; out[2 * TEST + 1] = t2 - t0; }
        RUN(T_SS128) RUN(T_SS256) RUN(T_TS128) RUN(T_SS16M) RUN(T_TS16M) RUN(T_SS32M) RUN(T_COMMIT) RUN(T_CHUNK4_COMMIT)
        // try_wait on an already completed phase
        drain(leader); __syncwarp(); t0 = clock64();
        for (int i = 0; i < REP; ++i) wait(b0, drain_parity ^ 1);
        t1 = clock64();
        if (lane == 0) { out[2 * T_TRYWAIT] = t1 - t0; out[2 * T_TRYWAIT + 1] = t1 - t0; }
        // weight-ring emulation (4 stages): wait full -> [4 MMAs] -> commit empty ; producer = warp 1
        for (int with_mma = 1; with_mma >= 0; --with_mma) {
            drain(leader); __syncwarp(); t0 = clock64();
            uint32_t stage = 0, parity = (with_mma ? 0 : 0);
            static_assert(REP % 4 == 0, "");
            for (int c = 0; c < REP; ++c) {
                wait(b0 + 8 * (16 + stage), parity ^ (with_mma ? 0 : (REP / 4) & 1));
                fence_after();
                if (leader) {
                    if (with_mma) for (int k = 0; k < 4; ++k) mma_ss(tmem, dA + (uint64_t)(2 * k), dB + (uint64_t)(2 * k), idesc(128, 128), 1);
                    commit(b0 + 8 * (20 + stage));
                }
                if (++stage == 4) { stage = 0; parity ^= 1; }
            }
            __syncwarp(); t1 = clock64(); drain(leader); t2 = clock64();
            const int T = with_mma ? T_RING_MMA : T_RING_EMPTY;
            if (lane == 0) { out[2 * T] = t1 - t0; out[2 * T + 1] = t2 - t0; }
        }
        // ping-pong with warp 2: arrive [24] -> wait [25]
        __syncwarp(); t0 = clock64();
        for (int i = 0; i < REP; ++i) { if (lane == 0) arrive(b0 + 8 * 24); wait(b0 + 8 * 25, i & 1); }
        t1 = clock64();
        if (lane == 0) { out[2 * T_PINGPONG] = t1 - t0; out[2 * T_PINGPONG + 1] = t1 - t0; }
    } else if (warp == 1) {
        // ring producer: two passes of REP chunks (with and without MMAs)
        if (lane == 0) {
            uint32_t stage = 0, parity = 1;
            for (int c = 0; c < 2 * REP; ++c) {
                wait(b0 + 8 * (20 + stage), parity);
                arrive(b0 + 8 * (16 + stage));
                if (++stage == 4) { stage = 0; parity ^= 1; }
            }
        }
    } else if (warp == 2) {
        for (int i = 0; i < REP; ++i) { wait(b0 + 8 * 24, i & 1); if (lane == 0) arrive(b0 + 8 * 25); }
    }
    __syncthreads();
    // TMEM load throughput: 4 warps (one per lane quarter), x32 loads back to back
    if (warp < 4) {
        uint32_t v[32];
        const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16);
        long long t0 = clock64();
        uint32_t acc = 0;
        for (int i = 0; i < REP; ++i) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(base + (i & 7) * 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[31];
        }
        long long t1 = clock64();
        if (acc == 12345u) out[63] = acc;
        if (threadIdx.x == 0) { out[2 * T_LDTM] = t1 - t0; out[2 * T_LDTM + 1] = t1 - t0; }
        __syncwarp();
        t0 = clock64();
        for (int i = 0; i < REP; ++i) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                         ::"r"(base + (i & 7) * 32), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        t1 = clock64();
        if (threadIdx.x == 0) { out[2 * T_LDTM_ST] = t1 - t0; out[2 * T_LDTM_ST + 1] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- weight-ring test with a real TMA producer: producer lane issues 16 KiB cp.async.bulk copies from an
// L2-resident buffer into an NS-stage ring; the MMA warp waits, issues 4 MMAs (SS or TS), commits the stage.
template <int NS, bool TS>
__global__ void __launch_bounds__(96, 1) ring_bench(const uint8_t* __restrict__ wsrc, int n_src_chunks, long long* out, int slot) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[32];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    const uint32_t b0 = smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0 + 8 * i), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    fence_after();
    const uint32_t tmem = tslot;
    const uint32_t sb = smem_u32(smem);
    const uint64_t dA = sw128_desc(sb);
    constexpr int NCH = 256;
    if (warp == 0) {
        const bool leader = elect_one();
        uint32_t stage = 0, parity = 0;
        long long t0 = clock64();
        for (int c = 0; c < NCH; ++c) {
            wait(b0 + 8 * stage, parity);
            fence_after();
            const uint64_t dW = sw128_desc(sb + 65536 + stage * 16384);
            if (leader) {
                for (int k = 0; k < 4; ++k) {
                    if (TS) mma_ts(tmem, tmem + 448 + 8 * k, dW + (uint64_t)(2 * k), idesc(128, 128), 1);
                    else mma_ss(tmem, dA + (uint64_t)(2 * k), dW + (uint64_t)(2 * k), idesc(128, 128), 1);
                }
                commit(b0 + 8 * (8 + stage));
            }
            if (++stage == NS) { stage = 0; parity ^= 1; }
        }
        __syncwarp();
        long long t1 = clock64();
        if (leader) commit(b0 + 8 * 31);
        wait(b0 + 8 * 31, 0);
        long long t2 = clock64();
        if (lane == 0) { out[2 * slot] = (t1 - t0); out[2 * slot + 1] = (t2 - t0); }
    } else if (warp == 1 && lane == 0) {
        uint32_t stage = 0, parity = 1;
        for (int c = 0; c < NCH; ++c) {
            wait(b0 + 8 * (8 + stage), parity);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0 + 8 * stage), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sb + 65536 + stage * 16384), "l"(wsrc + (size_t)(c % n_src_chunks) * 16384), "r"(16384), "r"(b0 + 8 * stage) : "memory");
            if (++stage == NS) { stage = 0; parity ^= 1; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    long long* d; cudaMalloc(&d, 64 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
    const char* names[NT_TESTS] = {"MMA N=128 SS", "MMA N=256 SS", "MMA N=128 TS", "MMA N=16 SS masked", "MMA N=16 TS masked", "MMA N=32 SS masked",
                                   "commit alone", "try_wait (completed)", "4 MMA N=128 + commit", "ring: wait+4MMA+commit", "ring: wait+commit only",
                                   "arrive->wait ping-pong", "tcgen05.ld x32 + wait", "tcgen05.st x32 + wait"};
    for (int rep = 0; rep < 2; ++rep) { bench<<<1, 160, 196608>>>(d); cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return 1; } }
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int t = 0; t < 14; ++t) printf("%-26s issue %7.1f cyc/op   issue+complete %7.1f cyc/op\n", names[t], h[2 * t] / (double)REP, h[2 * t + 1] / (double)REP);
    // TMA-fed ring: one CTA, then 148 CTAs (all SMs streaming the same 1.5 MB from L2)
    uint8_t* w; cudaMalloc(&w, 96 * 16384); cudaMemset(w, 0x3c, 96 * 16384);
    auto run = [&](auto kern, int smem_bytes, int grid, const char* name) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        for (int rep = 0; rep < 2; ++rep) { kern<<<grid, 96, smem_bytes>>>(w, 96, d, 20); cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return; } }
        long long r[2]; cudaMemcpy(r, d + 40, sizeof(r), cudaMemcpyDeviceToHost);
        printf("%-44s grid %3d: %6.1f cyc/chunk (4 MMAs)  drained %6.1f\n", name, grid, r[0] / 256.0, r[1] / 256.0);
    };
    for (int grid : {1, 148}) {
        run(ring_bench<4, false>, 65536 + 4 * 16384, grid, "TMA ring 4 stages, SS (A,B from smem)");
        run(ring_bench<6, false>, 65536 + 6 * 16384, grid, "TMA ring 6 stages, SS");
        run(ring_bench<8, false>, 65536 + 8 * 16384, grid, "TMA ring 8 stages, SS");
        run(ring_bench<4, true>, 65536 + 4 * 16384, grid, "TMA ring 4 stages, TS (A from TMEM)");
        run(ring_bench<6, true>, 65536 + 6 * 16384, grid, "TMA ring 6 stages, TS");
        run(ring_bench<8, true>, 65536 + 8 * 16384, grid, "TMA ring 8 stages, TS");
    }
    return 0;
}
