// maha_fit_tc.cu -- Mahalanobis sufficient statistics on the tensor cores (spec row A3, CMHAR_BF16 path).
//
// SPEC-DERIVED (no reference implementation): count_c, sum_c = sum_{label = c} f and the second
// moment M = sum f f^T over feature rows f in R^128; the tied covariance is finalised on the host in fp64 after
// the NCCL all-reduce of these three buffers (ood.py).
//
// The CUDA-core kernel (ood.cu) is fp32-FMA bound at 4 % of the HBM roofline the stage should sit on (AI = 64
// flop/B).  Here both reductions are GEMMs over the ROW dimension, K = rows:
//     M    (128 x 128) += F^T F          A = B = F^T tile
//     sum  (128 x 64)  += F^T O          A = F^T tile, B = one-hot(labels)^T tile (class slots x rows)
// with F split as hi + lo bf16: hi^T hi and X = hi^T lo on the tensor core (lo^T hi = X^T is added at the flush), the diagonal of
// lo^T lo -- the only systematic bias of dropping that term -- on the CUDA cores of the staging warps; both products for the
// class sums (one-hot entries are exact in bf16); fp32 accumulation in TMEM
// across all the tiles of a CTA, flushed with fp64 atomics every 256 tiles.
// F^T is never materialised: the feature tile is stored exactly as the score kernel stores it -- [128 rows x 64
// features] SWIZZLE_128B chunks, coalesced 512-byte row loads and conflict-free 8-byte stores -- and handed to the
// tensor core as an MN-MAJOR operand (instruction-descriptor bits 15/16; canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units, cute/atom/mma_traits_sm100.hpp: LBO = 16 KiB between the two 64-feature groups, SBO = 1 KiB
// between 8-row groups, a K = 16 step advances the start address by 2 KiB).  The first version scattered 2-byte
// elements into a transposed K-major tile (192 stores per thread per tile) and was staging-bound at 29 % of HBM.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace fittc {

using namespace tc;

constexpr int CHUNK = 16384;                       // [128 rows x 64 features] bf16, SWIZZLE_128B
constexpr int TILE = 2 * CHUNK;                    // 128 rows x 128 features
constexpr int OFF_THI = 0, OFF_TLO = TILE, OFF_OH = 2 * TILE, BUF = 2 * TILE + CHUNK;     // one staging buffer = 80 KiB
constexpr int OFF_CNT = 2 * BUF;                   // int counters [128]
constexpr int OFF_BAR = OFF_CNT + 512;
constexpr int SMEM_BYTES = OFF_BAR + 128;
enum { B_STAGED = 0, B_FREE = 2, B_ACC = 4, B_DRAINED = 5, B_COUNT = 6 };
constexpr int NSW = 8;                             // staging warps (8 or 16; 16 warps of 8 rows each: 226 / 125 us against 218 / 130 us for 2 M / 1 M rows,
                                                   // 96 registers and spills).
                                                   // 9 warps = 3 on one scheduler: 16 384 / 96 = 168 registers per thread is the ceiling
constexpr int RPW = 128 / NSW;                     // rows per staging warp
constexpr int LPR = 32 / RPW;                      // lanes per row of the one-hot tile, 8 / LPR pieces each
constexpr int FG = NSW / 4, CPG = 128 / FG;        // flush: FG warp groups, CPG accumulator columns each
constexpr int NT = NSW * 32 + 32;                  // staging warps + the MMA warp
constexpr int FLUSH_TILES = 256;                   // 32 768 rows of fp32 accumulation between fp64 flushes
constexpr int NCLS = 64;                           // class slots of the one-hot tile (classes <= 64)
constexpr uint32_t TM_M = 0, TM_S = 128, TM_X = 192;      // second moment (hi hi + lo lo) | class sums | cross term X = hi^T lo
constexpr int TM_COLS = 512;                                // 320 columns used; allocations are powers of two
constexpr int XLD = 129;                                    // padded row stride (floats) of the X scratch the flush transposes through

// MN-major SWIZZLE_128B descriptor of a [128 rows(K) x 128 features(MN)] tile made of two 64-feature chunks
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(CHUNK >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
constexpr uint32_t MN_BOTH = (1u << 15) | (1u << 16);      // instruction descriptor: A and B are MN-major

__device__ __forceinline__ float4 ld_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(NT, 1) maha_fit_tc_kernel(const float* __restrict__ feat, const long long* __restrict__ labels,
                                                            long long n, int C, double* __restrict__ count,
                                                            double* __restrict__ sum, double* __restrict__ second, int pf_on, int pf_lanes, int cfence) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    int* cnt = reinterpret_cast<int*>(smem + OFF_CNT);
    const long long tiles = (n + 127) / 128;
    constexpr int MMA_WARP = NSW;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(BAR(B_STAGED + b), NSW); mbar_init(BAR(B_FREE + b), 1); }
        mbar_init(BAR(B_ACC), 1);
        mbar_init(BAR(B_DRAINED), NSW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 128) cnt[tid] = 0;
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == MMA_WARP) {
        const bool leader = elect_one();
        uint32_t staged_parity[2] = {0, 0}, drained_parity = 0;
        long long it = 0;
        int since_flush = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            mbar_wait(BAR(B_STAGED + b), staged_parity[b], 90);
            staged_parity[b] ^= 1;
            // consumer-side proxy fence: the staging warps' generic-proxy stores (acquired through the barrier) -> the MMAs' async-proxy
            // reads.  fence.proxy.async compiles to MEMBAR.ALL.CTA, which also waits for every global load the executing thread has in
            // flight: executed by the staging warps it would serialise the half-tile of loads they keep ahead of the staging.
            if (cfence) fence_async_smem();
            tc_fence_after();
            const uint32_t base = sbase + b * BUF;
            const uint64_t dHi = mn_desc(base + OFF_THI), dLo = mn_desc(base + OFF_TLO), dOh = mn_desc(base + OFF_OH);
            constexpr uint32_t IDM = idesc_bf16(128, 128) | MN_BOTH, IDS = idesc_bf16(128, NCLS) | MN_BOTH;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                       // K = 16 rows per step: two 8-row groups = 2 KiB
                const uint64_t o = (uint64_t)(ks * (2048 >> 4));
                const uint32_t first = (since_flush == 0 && ks == 0) ? 0u : 1u;
                if (leader) {
                    // F^T F = hi^T hi + lo^T lo + X + X^T with X = hi^T lo: the two cross products are each other's transposes, so only
                    // one is issued (into its own accumulator) and the flush adds X + X^T -- three MMAs per step instead of four: the
                    // shared-memory port (8 KiB of operand reads per MMA at 128 B/clk) was the bound, 2 560 + 640 cycles of it per tile
                    // against 2 930 cycles of HBM time
                    // The lo^T lo term matters on the diagonal only (squares: the one systematic bias, ~1e-6 relative); off the diagonal
                    // it is a zero-mean sum at 2^-18 of the entry.  The staging warps accumulate the diagonal on the CUDA cores, so the
                    // tensor core issues two 128 x 128 MMAs per step, not four.
                    umma(tmem + TM_M, dHi + o, dHi + o, IDM, first);
                    umma(tmem + TM_X, dHi + o, dLo + o, IDM, first);
                    umma(tmem + TM_S, dHi + o, dOh + o, IDS, first);       // sum^T[feature][class] += F^T one-hot
                    umma(tmem + TM_S, dLo + o, dOh + o, IDS, 1u);
                }
            }
            if (leader) tc_commit(BAR(B_FREE + b));
            ++since_flush;
            const bool last = tile + gridDim.x >= tiles;
            if (since_flush == FLUSH_TILES || last) {
                if (leader) tc_commit(BAR(B_ACC));
                mbar_wait(BAR(B_DRAINED), drained_parity, 91);      // the staging warps have read the accumulators
                drained_parity ^= 1;
                tc_fence_after();
                since_flush = 0;
            }
        }
    } else {
        const int grp = warp >> 2;                                  // CPG accumulator columns per warp group when flushing
        const int row = (warp & 3) * 32 + lane;                     // accumulator row (feature) when flushing
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t free_parity[2] = {1, 1}, acc_parity = 0;
        long long it = 0;
        int since_flush = 0;
        float lo2[4] = {0.f, 0.f, 0.f, 0.f};                         // sum of lo^2 of features 4 lane .. 4 lane + 3 over this warp's rows
        // staging role: warp w owns rows RPW w .. RPW w + RPW - 1 of the tile; lane l holds features [4 l, 4 l + 4) of a row
        const int kc = lane >> 4, piece = (lane & 15) >> 1, sub = (lane & 1) * 8;
        const int oh_row = warp * RPW + lane / LPR, oh_q = lane % LPR;       // one-hot tile: LPR lanes per row, 8 / LPR pieces each
        // Loads run half a tile ahead of the staging: the first RPW / 2 rows of the next tile are requested as soon as those of this
        // one are staged, the second half likewise, so loads stay in flight while the warp converts and stores.
        // The staging warps are bound by the NUMBER of instructions they issue (ncu: 940 per thread and tile, half of them address
        // arithmetic, bounds predicates and selects; two warps per scheduler issue in 49 % of the cycles): full tiles whose rows all
        // carry a valid label take a path with immediate load / store offsets and no predicates or selects.
        float4 t[RPW];
        auto load_label = [&](long long tl) -> long long {
            const long long r = tl * 128 + oh_row;
            long long lab = (tl < tiles && r < n) ? __ldg(labels + r) : -1;
            return (lab < 0 || lab >= C) ? -1 : lab;                 // rows with a label outside [0, C) are skipped entirely
        };
#define FIT_ISSUE_HALF(TL, J0)                                                                           \
        {                                                                                                \
            const long long row0 = (TL) * 128 + warp * RPW;                                              \
            const float* src = feat + row0 * D + 4 * lane;                                               \
            if ((TL) < tiles && row0 + RPW <= n) {                                                       \
                _Pragma("unroll") for (int j = (J0); j < (J0) + RPW / 2; ++j) t[j] = ld_stream(src + j * D);      \
            } else {                                                                                     \
                _Pragma("unroll") for (int j = (J0); j < (J0) + RPW / 2; ++j) {                        \
                    t[j] = make_float4(0.f, 0.f, 0.f, 0.f);                                              \
                    if ((TL) < tiles && row0 + j < n) t[j] = ld_stream(src + j * D);                     \
                }                                                                                        \
            }                                                                                            \
        }
        // store offsets inside a [128 x 64] SW128 chunk: row r = RPW warp + j -> (r >> 3) * 1024 + (r & 7) * 128 + ((piece ^ (r & 7)) << 4);
        // RPW is a multiple of 8, so everything but the XOR term is a per-thread constant plus a compile-time function of j
        const uint32_t st_base = (uint32_t)(kc * CHUNK) + (uint32_t)(warp * (RPW / 8) * 1024) + (uint32_t)sub;
        uint32_t pxor[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) pxor[c] = (uint32_t)((piece ^ c) << 4);
#define FIT_STAGE_ROW(J, X)                                                                              \
        {                                                                                                \
            const __nv_bfloat162 h01 = __floats2bfloat162_rn((X).x, (X).y), h23 = __floats2bfloat162_rn((X).z, (X).w);      \
            const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);  \
            const float e0 = (X).x - __uint_as_float(u01 << 16), e1 = (X).y - __uint_as_float(u01 & 0xffff0000u);           \
            const float e2 = (X).z - __uint_as_float(u23 << 16), e3 = (X).w - __uint_as_float(u23 & 0xffff0000u);           \
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(e0, e1), l23 = __floats2bfloat162_rn(e2, e3);                  \
            lo2[0] = fmaf(e0, e0, lo2[0]); lo2[1] = fmaf(e1, e1, lo2[1]);                                                   \
            lo2[2] = fmaf(e2, e2, lo2[2]); lo2[3] = fmaf(e3, e3, lo2[3]);                                                   \
            uint8_t* dst = tile_base + (((J) >> 3) * 1024 + ((J) & 7) * 128) + pxor[(J) & 7];                               \
            *reinterpret_cast<uint2*>(dst + OFF_THI) = make_uint2(u01, u23);                                                \
            *reinterpret_cast<uint2*>(dst + OFF_TLO) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23)); \
        }
        long long lab_next = load_label((long long)blockIdx.x);
        FIT_ISSUE_HALF((long long)blockIdx.x, 0)
        FIT_ISSUE_HALF((long long)blockIdx.x, RPW / 2)
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            const long long next_tile = tile + gridDim.x;
            if (pf_on && warp == 0 && lane < pf_lanes) {            // L2 prefetch of the tile pf_on steps ahead (tc_ptx.cuh l2_prefetch_bulk),
                const long long tl = tile + (long long)pf_on * gridDim.x;      // split over pf_lanes lanes
                if (tl < tiles) {
                    const long long rows = n - tl * 128;
                    const int nrow = (int)(rows < 128 ? rows : 128), per = (128 + pf_lanes - 1) / pf_lanes;
                    const int r0 = lane * per, r1 = (r0 + per < nrow) ? r0 + per : nrow;
                    if (r1 > r0) tc::l2_prefetch_bulk(feat + ((size_t)tl * 128 + r0) * D, (uint32_t)((r1 - r0) * D * 4));
                }
            }
            const long long lab_oh = lab_next;
            mbar_wait(BAR(B_FREE + b), free_parity[b], 92);          // the MMAs that read this buffer are complete
            free_parity[b] ^= 1;
            uint8_t* buf = smem + b * BUF;
            uint8_t* tile_base = buf + st_base;
            const unsigned keep_mask = __ballot_sync(0xffffffffu, lab_oh >= 0);      // row RPW w + j <-> lanes LPR j ..
            // (the lo^2 diagonal term uses the exact fp32 remainder x - hi, not its bf16 rounding: a 2^-9 relative change of a 1e-6 term)
            if (keep_mask == 0xffffffffu) {
#pragma unroll
                for (int j = 0; j < RPW / 2; ++j) FIT_STAGE_ROW(j, t[j])
                FIT_ISSUE_HALF(next_tile, 0)
#pragma unroll
                for (int j = RPW / 2; j < RPW; ++j) FIT_STAGE_ROW(j, t[j])
                FIT_ISSUE_HALF(next_tile, RPW / 2)
            } else {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                    for (int j = (RPW / 2) * hf; j < (RPW / 2) * hf + RPW / 2; ++j) {
                        const float4 x = ((keep_mask >> (LPR * j)) & 1u) ? t[j] : make_float4(0.f, 0.f, 0.f, 0.f);     // unlabeled rows contribute nothing
                        FIT_STAGE_ROW(j, x)
                    }
                    if (hf == 0) { FIT_ISSUE_HALF(next_tile, 0) } else { FIT_ISSUE_HALF(next_tile, RPW / 2) }
                }
            }
            lab_next = load_label(next_tile);
            {   // one-hot row: 64 class slots = 8 pieces of 16 bytes, this lane writes pieces (8 / LPR) oh_q ..
                const int lab = (int)lab_oh;
                const uint32_t one = 0x3F80u << (16 * (lab & 1));    // bf16 1.0 in the low or high half of its word
                const int wsel = (lab & 7) >> 1, psel = lab >> 3;    // word within the piece, piece within the row (negative label: no match)
#pragma unroll
                for (int q = 0; q < 8 / LPR; ++q) {
                    const int p = (8 / LPR) * oh_q + q;
                    const bool hit = lab >= 0 && psel == p;
                    *reinterpret_cast<uint4*>(buf + OFF_OH + sw128_off(oh_row, p)) =
                        make_uint4(hit && wsel == 0 ? one : 0u, hit && wsel == 1 ? one : 0u, hit && wsel == 2 ? one : 0u, hit && wsel == 3 ? one : 0u);
                }
                if (oh_q == 0 && lab >= 0) atomicAdd(cnt + lab, 1);
            }
            if (!cfence) fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_STAGED + b));
            ++since_flush;
            const bool last = tile + gridDim.x >= tiles;
            if (since_flush == FLUSH_TILES || last) {
                // ---- flush: fp32 accumulators -> fp64 atomics (accumulator row = lane row, this thread's 64 columns)
                mbar_wait(BAR(B_ACC), acc_parity, 93);
                acc_parity ^= 1;
                tc_fence_after();
                uint32_t v[32];
                // X (this thread's part of its row) goes through a padded shared-memory scratch so that X^T can be read back: both staging
                // buffers are idle here (every MMA that read them has completed: B_ACC), the scratch lies over buffer 0
                float* xs = reinterpret_cast<float*>(smem);
#pragma unroll 1
                for (int ch = 0; ch < CPG / 32; ++ch) {
                    const int c0 = CPG * grp + 32 * ch;
                    TMEM_LD32(lane_base + TM_X + c0, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) xs[row * XLD + c0 + j] = __uint_as_float(v[j]);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NSW * 32) : "memory");
#pragma unroll 1
                for (int ch = 0; ch < CPG / 32; ++ch) {
                    const int c0 = CPG * grp + 32 * ch;
                    TMEM_LD32(lane_base + TM_M + c0, v);
                    tc_wait_ld();
#pragma unroll 8
                    for (int j = 0; j < 32; ++j)
                        atomicAdd(second + (size_t)row * D + c0 + j,
                                  (double)__uint_as_float(v[j]) + ((double)xs[row * XLD + c0 + j] + (double)xs[(c0 + j) * XLD + row]));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {                         // diagonal of lo^T lo (see the MMA loop)
                    atomicAdd(second + (size_t)(4 * lane + q) * (D + 1), (double)lo2[q]);
                    lo2[q] = 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NSW * 32) : "memory");      // the scratch is staging buffer 0 again from here on
                if constexpr (FG == 2) { TMEM_LD32(lane_base + TM_S + 32 * grp, v); }       // sum^T: lane = feature, column = class slot
                else { TMEM_LD16(lane_base + TM_S + 16 * grp, v); }
                tc_wait_ld();
#pragma unroll 8
                for (int j = 0; j < NCLS / FG; ++j) {
                    const int c = (NCLS / FG) * grp + j;
                    if (c < C) atomicAdd(sum + (size_t)c * D + row, (double)__uint_as_float(v[j]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_DRAINED));
                since_flush = 0;
            }
        }
        // class counts: exact integers
        asm volatile("bar.sync 1, %0;" ::"n"(NSW * 32) : "memory");
        if (tid < C && cnt[tid] != 0) atomicAdd(count + tid, (double)cnt[tid]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS));
    }
}

#undef FIT_ISSUE_HALF
#undef FIT_STAGE_ROW
}  // namespace fittc

int launch_maha_fit_tc(const float* feat, const long long* labels, long long n, int C, double* count, double* sum, double* second,
                       cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(fittc::maha_fit_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fittc::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    const long long tiles = (n + 127) / 128;
    const int grid = (int)(tiles < (long long)sm_count() ? tiles : (long long)sm_count());
    static int pf = -1, pfl = -1;
    if (pf < 0) { const char* e = dev_getenv("CMHAR_L2_PREFETCH"); pf = e ? atoi(e) : 3; }       // development switch: tiles ahead (0 = off)
    if (pfl < 0) { const char* e = dev_getenv("CMHAR_L2_PF_LANES"); pfl = e ? atoi(e) : 1; if (pfl < 1 || pfl > 32) pfl = 1; }
    static int cf = -1;
    if (cf < 0) { const char* e = dev_getenv("CMHAR_FIT_CFENCE"); cf = e ? atoi(e) : 1; }        // development switch: 0 = writer-side proxy fence
    fittc::maha_fit_tc_kernel<<<grid, fittc::NT, fittc::SMEM_BYTES, st>>>(feat, labels, n, C, count, sum, second, pf, pfl, cf);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
