// maha_score_tc.cu -- Mahalanobis OOD score from stored features as ONE streaming GEMM-plus-reduction kernel on
// the tensor cores (spec row A4, CMHAR_BF16 path of cmhar_maha_score).
//
// SPEC-DERIVED (no reference implementation): score(f) = min_c || f W - mu_c W ||^2 with W the
// whitening factor of the tied covariance (oracle/ood_spec.py).  HBM-bound by design: 512 B in + 4 B out per row
// against 40 960 flop (AI = 79).  Expanding the square in FEATURE space,
//     dist_c = |f W|^2 - 2 f . g_c + |mu_c W|^2 ,     g_c = W (mu_c W)^T     (a 128-vector per class, packed once),
// both products share the A operand f, so the whole stage is one GEMM  [Y | G] = F [W | g_0 .. g_31]  (N = 160)
// followed by a per-row reduction that needs no cross-lane traffic at all: the accumulator row of a feature row
// lives in one TMEM lane, i.e. in one epilogue thread.
//
// Split-bf16 operands as in head_tc.cu (x = hi + lo, D += A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulate):
// products exact to ~2^-17, so the score is fp32-grade although the tensor cores only take bf16.
//
// Persistent CTA per SM, three roles that overlap tile by tile:
//   * 2 x 8 staging warps: coalesced 512-byte row loads (16 rows in flight per warp), split into hi / lo bf16 and
//     stored as K-major SWIZZLE_128B A tiles (8-byte stores, conflict-free); the two groups work on alternate
//     tiles (= alternate A buffers), one converting while the other's loads are in flight;
//   * 4 epilogue warps: thread = row; |y|^2 over 128 columns, min over the 32 class columns, one 4-byte store;
//     the first of them also issues the MMAs, one tile ahead of the accumulator it drains: 24 tcgen05.mma (M=128,
//     N=160, K=16) per tile against the [W | G] image that stays RESIDENT in shared memory (80 KiB, loaded once
//     per CTA by cp.async.bulk) into one of two TMEM accumulators.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace mstc {

using namespace tc;

constexpr int NB = 160;                               // B rows: 128 whitened coordinates + 32 class cross terms
constexpr int BCH = NB * 128;                         // one [160 x 64] bf16 SW128 chunk = 20 480 B
constexpr int W_BYTES = 4 * BCH;                      // [k chunk 0: hi | lo][k chunk 1: hi | lo]
constexpr int ACH = 16384;                            // one [128 x 64] A chunk
constexpr int ABUF = 4 * ACH;                         // hi (2 chunks) | lo (2 chunks)
constexpr int OFF_W = 0, OFF_A = W_BYTES, OFF_BAR = OFF_A + 2 * ABUF, OFF_MUN = OFF_BAR + 128, SMEM_BYTES = OFF_MUN + 128;
static_assert(OFF_A % 1024 == 0 && SMEM_BYTES <= 232448, "shared memory map");
enum { B_W = 0, B_STAGED = 1, B_FREE = 3, B_ACCFULL = 5, B_ACCFREE = 7, B_COUNT = 9 };
// Role layout, templated for the variants compared in tools/bench_hbm_kernels.py (CMHAR_MAHA_VARIANT):
//   GROUPS  staging groups of 8 warps (group g stages this CTA's tiles g, g + GROUPS, ...)
//   MERGED  the MMAs are issued by epilogue warp 0 (one tile ahead of the accumulator it drains) instead of a
//           dedicated warp -- 20 warps = 5 per SM sub-partition leave 96 registers per thread, 21 leave 80
constexpr int ROWS_PER_WARP = 16;
template <int GROUPS, bool MERGED> struct Roles {
    static constexpr int STAGE_WARPS = 8 * GROUPS, EPI_WARP0 = STAGE_WARPS, MMA_WARP = MERGED ? EPI_WARP0 : EPI_WARP0 + 4;
    static constexpr int NT = (STAGE_WARPS + 4 + (MERGED ? 0 : 1)) * 32;
};
constexpr uint32_t ACC_STRIDE = 256;                  // TMEM columns between the two accumulators (160 used)
constexpr size_t SECTION_BYTES = W_BYTES + 32 * sizeof(float);      // image + |mu_c W|^2 (inf: empty / padding class)

// One thread per (B row n, k): n < 128 -> W[k][n]; n >= 128 -> g_c[k] = sum_j W[k][j] mu_w[c][j]  (fp64 accumulate).
__global__ void pack_kernel(const float* __restrict__ whiten, const float* __restrict__ mean_w, const float* __restrict__ valid,
                            int C, uint8_t* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NB * D) return;
    const int n = i / D, k = i % D;
    double v = 0.0;
    if (n < D) v = whiten[(size_t)k * D + n];
    else {
        const int c = n - D;
        if (c < C && valid[c] > 0.f)
            for (int j = 0; j < D; ++j) v += (double)whiten[(size_t)k * D + j] * (double)mean_w[(size_t)c * D + j];
    }
    const float vf = (float)v;
    const __nv_bfloat16 hi = __float2bfloat16_rn(vf);
    const __nv_bfloat16 lo = __float2bfloat16_rn((float)(v - (double)__bfloat162float(hi)));
    uint8_t* chunk = dst + (size_t)(k >> 6) * (2 * BCH);
    const uint32_t off = sw128_off(n, (k & 63) >> 3) + (uint32_t)((k & 7) * 2);
    *reinterpret_cast<__nv_bfloat16*>(chunk + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(chunk + BCH + off) = lo;
}

__global__ void mu_norm_kernel(const float* __restrict__ mean_w, const float* __restrict__ valid, int C, float* __restrict__ dst) {
    const int c = threadIdx.x;
    if (c >= 32) return;
    float s = INFINITY;
    if (c < C && valid[c] > 0.f) {
        double acc = 0.0;
        for (int k = 0; k < D; ++k) acc += (double)mean_w[c * D + k] * (double)mean_w[c * D + k];
        s = (float)acc;
    }
    dst[c] = s;
}

__device__ __forceinline__ float4 ld_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// convert this warp's rows of one tile (registers t) into the hi / lo A tiles of buffer `abuf`.
// The staging warps are bound by the number of instructions they issue (two warps per scheduler; measured on the fit kernel, which
// stages the same way: 940 instructions per thread and tile, half of them address arithmetic and predicates), so: row0 is a multiple
// of 8, which makes the SW128 offset a per-thread constant + a compile-time function of j + one XOR; the hi halves are unpacked with
// a shift and a mask; full tiles load with immediate offsets and no predicates.
template <int ROWS>
__device__ __forceinline__ void stage_rows(const float4 (&t)[ROWS], uint8_t* abuf, int row0, int piece, int sub) {
    uint8_t* base = abuf + (row0 >> 3) * 1024 + sub;
    const uint32_t p4 = (uint32_t)piece << 4;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const float4 x = t[j];
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
        const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - __uint_as_float(u01 << 16), x.y - __uint_as_float(u01 & 0xffff0000u));
        const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - __uint_as_float(u23 << 16), x.w - __uint_as_float(u23 & 0xffff0000u));
        uint8_t* dst = base + ((j >> 3) * 1024 + (j & 7) * 128) + (p4 ^ (uint32_t)((j & 7) << 4));
        *reinterpret_cast<uint2*>(dst) = make_uint2(u01, u23);
        *reinterpret_cast<uint2*>(dst + 2 * ACH) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
}
// this warp's rows of a tile -> registers: src = address of (first row, this lane's 4 elements); rows past
// `valid` read as zero
template <int ROWS>
__device__ __forceinline__ void load_rows(float4 (&t)[ROWS], const float* __restrict__ src, int valid) {
    if (valid >= ROWS) {
#pragma unroll
        for (int j = 0; j < ROWS; ++j) t[j] = ld_stream(src + j * D);
    } else {
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            t[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < valid) t[j] = ld_stream(src + j * D);
        }
    }
}

template <int GROUPS, bool MERGED, bool CFENCE>
__global__ void __launch_bounds__(Roles<GROUPS, MERGED>::NT, 1) maha_score_tc_kernel(const uint8_t* __restrict__ section,
                                                                                     const float* __restrict__ feat, long long n,
                                                                                     float* __restrict__ score, int pf_on) {
    using R = Roles<GROUPS, MERGED>;
    constexpr int N_STAGE_WARPS = R::STAGE_WARPS, EPI_WARP0 = R::EPI_WARP0, MMA_WARP = R::MMA_WARP;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    const long long tiles = (n + 127) / 128;

    if (tid == 0) {
        mbar_init(BAR(B_W), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(B_STAGED + b), 8);
            mbar_init(BAR(B_FREE + b), 1);
            mbar_init(BAR(B_ACCFULL + b), 1);
            mbar_init(BAR(B_ACCFREE + b), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < N_STAGE_WARPS) {
        // ================================================================= staging: fp32 rows -> split-bf16 A tiles
        // Two groups of 8 warps, out of phase: group g stages this CTA's tiles g, g+2, ... into A buffer g; inside a
        // group warp w owns rows 16 w .. 16 w + 15 and lane l holds elements [4 l, 4 l + 4) of a row.  A step is:
        // convert + store the rows held in registers, proxy fence, arrive, THEN issue the loads of the group's next
        // tile -- the fence is a MEMBAR that would otherwise wait for every load still in flight -- so one group
        // converts while the other group's 64 KiB are on their way from HBM.
        const int g = warp >> 3, gw = warp & 7;
        const int kc = lane >> 4, piece = (lane & 15) >> 1, sub = (lane & 1) * 8;
        const int row0 = gw * ROWS_PER_WARP;
        float4 t[ROWS_PER_WARP];
        // 32-bit tile bookkeeping (tiles < 2^31 checked by the launcher) keeps this role inside its register budget
        const int ntiles = (int)tiles, step = GROUPS * (int)gridDim.x;
        const float* lane_src = feat + (size_t)row0 * D + 4 * lane;
        auto fetch = [&](int tile) __attribute__((always_inline)) {
            const long long left = n - ((long long)tile * 128 + row0);
            load_rows<ROWS_PER_WARP>(t, lane_src + (size_t)tile * (128 * D), left < ROWS_PER_WARP ? (int)left : ROWS_PER_WARP);
        };
        int tile = (int)blockIdx.x + g * (int)gridDim.x;
        // L2 prefetch of the tiles PF_AHEAD .. steps ahead (one lane per CTA; see tc_ptx.cuh l2_prefetch_bulk)
        constexpr int PF_AHEAD = 3;
        auto prefetch_tile = [&](int tl) __attribute__((always_inline)) {
            if (tl < ntiles) {
                const long long rows = n - (long long)tl * 128;
                l2_prefetch_bulk(feat + (size_t)tl * (128 * D), (uint32_t)((rows < 128 ? rows : 128) * D * 4));
            }
        };
        if (pf_on > 0 && warp == 0 && lane == 0)
            for (int a = 1; a < PF_AHEAD; ++a) prefetch_tile(tile + a * step);
        if (tile < ntiles) fetch(tile);
        for (int it = g; tile < ntiles; tile += step, it += GROUPS) {        // it = index within this CTA's tile sequence
            const int b = it & 1;                                    // A buffer == accumulator index of this tile
            if (pf_on > 0 && warp == 0 && lane == 0) prefetch_tile(tile + PF_AHEAD * step);
            mbar_wait(BAR(B_FREE + b), (uint32_t)(((it >> 1) & 1) ^ 1), 83);       // the MMAs that read this buffer are complete
            if (CFENCE) {
                // consumer-side proxy fence: this warp never executes the MEMBAR, so the refill loads issued row by row
                // right after each row is consumed stay in flight across the hand-off (a full tile of prefetch)
                // (rows are refilled half a tile at a time, right after the half is converted and stored)
                const bool more = tile + step < ntiles;
                const long long left = more ? n - ((long long)(tile + step) * 128 + row0) : 0;
                const int valid = left < ROWS_PER_WARP ? (int)left : ROWS_PER_WARP;
                const float* nsrc = lane_src + (size_t)(tile + step) * (128 * D);
                uint8_t* abuf = smem + OFF_A + b * ABUF + kc * ACH;
                constexpr int HR = ROWS_PER_WARP / 2;
                float4 (&ta)[HR] = *reinterpret_cast<float4 (*)[HR]>(&t[0]);
                float4 (&tb)[HR] = *reinterpret_cast<float4 (*)[HR]>(&t[HR]);
                stage_rows<HR>(ta, abuf, row0, piece, sub);
                load_rows<HR>(ta, nsrc, valid);
                stage_rows<HR>(tb, abuf, row0 + HR, piece, sub);
                load_rows<HR>(tb, nsrc + HR * D, valid - HR);
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_STAGED + b));       // release: the warp's shared-memory stores are ordered before it
            } else {
                stage_rows<ROWS_PER_WARP>(t, smem + OFF_A + b * ABUF + kc * ACH, row0, piece, sub);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_STAGED + b));
                if (tile + step < ntiles) fetch(tile + step);
            }
        }
    } else {
        // ================================================================= epilogue (thread = row) and MMA issue
        const bool is_epi = warp < EPI_WARP0 + 4;
        const bool issuer = (warp == MMA_WARP);
        const int q = warp & 3;                                      // the TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        float* mun = reinterpret_cast<float*>(smem + OFF_MUN);       // |mu_c W|^2, read back as warp-uniform broadcasts
        if (is_epi) {
            if (q == 0) mun[lane] = __ldg(reinterpret_cast<const float*>(section + W_BYTES) + lane);
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        // ---- MMA issue (one converged warp, one elected lane): 24 MMAs of the CTA's tile `it` into accumulator it & 1
        const bool leader = issuer && elect_one();
        constexpr uint32_t ID160 = idesc_bf16(128, NB);
        uint32_t staged_par = 0u, accfree_par = 3u;                  // one parity bit per buffer; fresh barrier: parity 1 passes
        auto issue_tile = [&](const int b) __attribute__((always_inline)) {
            mbar_wait(BAR(B_ACCFREE + b), (accfree_par >> b) & 1u, 81);
            accfree_par ^= 1u << b;
            mbar_wait(BAR(B_STAGED + b), (staged_par >> b) & 1u, 82);
            staged_par ^= 1u << b;
            if (CFENCE) fence_async_smem();      // generic-proxy stores (acquired through the barrier) -> async-proxy reads of the MMAs
            tc_fence_after();
            const uint32_t abase = sbase + OFF_A + b * ABUF;
            const uint32_t d = tmem + ACC_STRIDE * (uint32_t)b;
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
                const uint64_t a_hi = sw128_desc(abase + kc * ACH), a_lo = sw128_desc(abase + 2 * ACH + kc * ACH);
                const uint64_t b_hi = sw128_desc(sbase + OFF_W + kc * 2 * BCH), b_lo = sw128_desc(sbase + OFF_W + kc * 2 * BCH + BCH);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t o = (uint64_t)(2 * k);
                    if (leader) {
                        umma(d, a_hi + o, b_hi + o, ID160, (kc == 0 && k == 0) ? 0u : 1u);
                        umma(d, a_lo + o, b_hi + o, ID160, 1u);
                        umma(d, a_hi + o, b_lo + o, ID160, 1u);
                    }
                }
            }
            if (leader) { tc_commit(BAR(B_FREE + b)); tc_commit(BAR(B_ACCFULL + b)); }
        };
        if (issuer) {
            if (leader) {                                            // the [W | G] image: resident for the life of the CTA
                mbar_expect_tx(BAR(B_W), W_BYTES);
                for (int c = 0; c < 4; ++c) bulk_g2s(sbase + OFF_W + c * BCH, section + (size_t)c * BCH, BCH, BAR(B_W));
            }
            __syncwarp();
            mbar_wait(BAR(B_W), 0, 80);
            if (!MERGED) {                                           // dedicated warp: just run ahead of the epilogue
                long long it = 0;
                for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) issue_tile((int)(it & 1));
            } else if ((long long)blockIdx.x < tiles) {
                issue_tile(0);
            }
        }
        if (is_epi) {
            uint32_t full_par = 0u;
            long long it = 0;
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
                const int b = (int)(it & 1);
                if (MERGED && issuer && tile + gridDim.x < tiles) issue_tile(b ^ 1);      // keep the tensor pipe one tile ahead of the drain
                mbar_wait(BAR(B_ACCFULL + b), (full_par >> b) & 1u, 84);
                full_par ^= 1u << b;
                tc_fence_after();
                const uint32_t acc = lane_base + ACC_STRIDE * (uint32_t)b;
                uint32_t v[32];
                float yy0 = 0.f, yy1 = 0.f, yy2 = 0.f, yy3 = 0.f;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    TMEM_LD32(acc + 32 * h, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        yy0 = fmaf(__uint_as_float(v[i]), __uint_as_float(v[i]), yy0);
                        yy1 = fmaf(__uint_as_float(v[i + 1]), __uint_as_float(v[i + 1]), yy1);
                        yy2 = fmaf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 2]), yy2);
                        yy3 = fmaf(__uint_as_float(v[i + 3]), __uint_as_float(v[i + 3]), yy3);
                    }
                }
                TMEM_LD32(acc + 128, v);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_ACCFREE + b));      // accumulator b may be overwritten
                float best = INFINITY;
#pragma unroll
                for (int c = 0; c < 32; ++c) best = fminf(best, fmaf(-2.f, __uint_as_float(v[c]), mun[c]));
                const long long r = tile * 128 + row;
                if (r < n) score[r] = fmaxf((yy0 + yy1) + (yy2 + yy3) + best, 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

}  // namespace mstc

size_t maha_score_tc_bytes() { return mstc::SECTION_BYTES; }

int pack_maha_score_tc(const float* f32, const MahaLayout& ml, uint8_t* dst, cudaStream_t st) {
    mstc::pack_kernel<<<(mstc::NB * D + 255) / 256, 256, 0, st>>>(f32 + ml.whiten(), f32 + ml.mean_w(), f32 + ml.valid(), ml.C, dst);
    CMHAR_LAUNCH_CHECK();
    mstc::mu_norm_kernel<<<1, 32, 0, st>>>(f32 + ml.mean_w(), f32 + ml.valid(), ml.C, reinterpret_cast<float*>(dst + mstc::W_BYTES));
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

// section: device pointer to the [W | G] image + class norms inside the maha blob; feat (n,128) fp32, 16-byte aligned
template <int GROUPS, bool MERGED, bool CFENCE>
static int launch_variant(const uint8_t* section, const float* feat, long long n, float* score, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    auto kern = mstc::maha_score_tc_kernel<GROUPS, MERGED, CFENCE>;
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, mstc::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    const long long tiles = (n + 127) / 128;
    CMHAR_REQUIRE(tiles < 0x7fffffffLL, "too many rows");
    const int grid = (int)(tiles < (long long)sm_count() ? tiles : (long long)sm_count());
    static int pf = -1;
    if (pf < 0) { const char* e = dev_getenv("CMHAR_L2_PREFETCH"); pf = e ? atoi(e) : 1; }      // development switch (default on)
    kern<<<grid, mstc::Roles<GROUPS, MERGED>::NT, mstc::SMEM_BYTES, st>>>(section, feat, n, score, pf);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int launch_maha_score_tc(const uint8_t* section, const float* feat, long long n, float* score, cudaStream_t st) {
    static int variant = -1;
    if (variant < 0) { const char* e = dev_getenv("CMHAR_MAHA_VARIANT"); variant = e ? atoi(e) : 0; }      // development switch
    // Measured on B200, 2 M rows (tools/bench_maha_variants.py): one staging group + dedicated MMA warp 213 us (74 % of the
    // copy bandwidth) with the canonical writer-side proxy fence, 210 us with the consumer-side fence that keeps a full
    // tile of loads in flight -- i.e. not latency-bound any more: a tile moves 64 KiB of staging stores + 24 x 9 KiB of
    // operand reads through the SM's 128 B/clk shared-memory port (2 240 cycles of the ~3 800 per tile).  Two staging
    // groups (21 warps -> 80 registers, spills) 292 us; MMAs issued from an epilogue warp (descriptors leave the uniform
    // datapath) 583 us -- both dropped.
    if (variant == 3) return launch_variant<1, false, true>(section, feat, n, score, st);
    return launch_variant<1, false, false>(section, feat, n, score, st);
}

}  // namespace cmhar
