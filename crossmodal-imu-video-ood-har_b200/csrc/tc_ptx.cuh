// tc_ptx.cuh -- inline-PTX wrappers shared by the sm_100a kernels: mbarrier, bulk async copy (TMA engine),
// tcgen05 (MMA / TMEM load-store / commit / fences), SWIZZLE_128B shared-memory descriptors.
#pragma once
#include "common.cuh"

namespace cmhar {
namespace tc {

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int site = 0) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 22)) {       // watchdog: a protocol bug must fail loudly, never hang the GPU
            if (spins == (1u << 22) + 1 && (threadIdx.x & 31) == 0 && blockIdx.x == 0)
                printf("cmhar bf16 kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u site %d)\n",
                       (int)blockIdx.x, (int)threadIdx.x, (bar & 0xffu) >> 3, parity, site);
            if (spins > (1u << 26)) __trap();
        }
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// L2 eviction policies: the streaming kernels read their input exactly once (evict_first), the weight images are
// re-read by every tile of every CTA (evict_last) -- without the hints a 67 MB feature-map stream per step pushes
// the 1.5 MB of encoder weights out of L2 and every ring stage becomes an HBM round trip.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// fire-and-forget prefetch of a contiguous range into L2 (no registers, no shared memory): the streaming kernels keep their loads in
// registers (64 KiB in flight per SM = 4.7 TB/s at ~2 us of HBM latency, Little's law); tiles a few iterations ahead are pulled into L2
// so that those loads see L2 latency instead.  p 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one lane of a converged warp (the same one every time): tcgen05.mma / commit are issued under this
// predicate while the whole warp stays converged, so descriptors live in uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int NTH>
__device__ __forceinline__ void epi_bar_n() { asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; both operands K-major SWIZZLE_128B
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T : A from tensor memory (lane = row, one 32-bit column = two consecutive
// bf16 k elements, a K=16 step reads 8 columns), B K-major SWIZZLE_128B in shared memory
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// same, but only the 16 TMEM lanes [16*win, 16*win+16) are written (disable-output-lane mask)
__device__ __forceinline__ void umma_rows16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int win) {
    const uint32_t hole = ~(0xFFFFu << ((win & 1) * 16));      // no array indexing: keeps the mask in registers
    const uint32_t m[4] = {(win >> 1) == 0 ? hole : 0xFFFFFFFFu, (win >> 1) == 1 ? hole : 0xFFFFFFFFu,
                           (win >> 1) == 2 ? hole : 0xFFFFFFFFu, (win >> 1) == 3 ? hole : 0xFFFFFFFFu};
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %5, %6, %7}, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]) : "memory");
}

// A operand from TMEM (lane = row, one 32-bit column = two consecutive k elements), B from smem
__device__ __forceinline__ void umma_ts_rows16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, int win) {
    const uint32_t hole = ~(0xFFFFu << ((win & 1) * 16));      // no array indexing: keeps the mask in registers
    const uint32_t m[4] = {(win >> 1) == 0 ? hole : 0xFFFFFFFFu, (win >> 1) == 1 ? hole : 0xFFFFFFFFu,
                           (win >> 1) == 2 ? hole : 0xFFFFFFFFu, (win >> 1) == 3 ? hole : 0xFFFFFFFFu};
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %5, %6, %7}, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 [0,14) | LBO=1 [16,30) | SBO=1024>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define TMEM_LD32(addr, v)                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                        \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25," \
                 "%26,%27,%28,%29,%30,%31}, [%32];"                                                               \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),       \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),     \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),     \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                          \
                 : "r"(addr)                                                                                      \
                 : "memory")

#define TMEM_LD16(addr, v)                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(addr)                                                                                      \
                 : "memory")

#define TMEM_ST32(addr, v)                                                                                        \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                  \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26," \
                 "%27,%28,%29,%30,%31,%32};"                                                                      \
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),        \
                   "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),   \
                   "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), \
                   "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), \
                   "r"(v[31])                                                                                     \
                 : "memory")

#define TMEM_ST16(addr, v)                                                                                        \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                  \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                                      \
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),        \
                   "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),   \
                   "r"(v[15])                                                                                     \
                 : "memory")

__device__ __forceinline__ float ex2_approx(float x) {       // 2^x, one MUFU, flush-to-zero (x <= 0 here)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// max(x, 0) fused into the conversion: one F2FP per pair instead of two FMNMX + one F2FP
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// byte offset of the 16-byte piece holding columns [8j, 8j+8) of row r inside a [128 x 64] SW128 chunk
__device__ __forceinline__ uint32_t sw128_off(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}

// store 32 consecutive columns (fp32 in v, already finished) of row r as bf16 into chunk `chunk_base`
// at 16-byte pieces j0..j0+3
__device__ __forceinline__ void store_bf16_32(uint8_t* chunk_base, int r, int j0, const float* v) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
        u.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
        u.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
        u.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(chunk_base + sw128_off(r, j0 + q)) = u;
    }
}

}  // namespace tc
}  // namespace cmhar
