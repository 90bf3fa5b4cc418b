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
// with F split as hi + lo bf16 (all four hi/lo products for M -- the lo*lo term removes the only systematic bias,
// on the diagonal -- and both for the class sums; one-hot entries are exact in bf16), fp32 accumulation in TMEM
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
constexpr int NT = 8 * 32 + 32;                    // 8 staging warps + the MMA warp
constexpr int FLUSH_TILES = 256;                   // 32 768 rows of fp32 accumulation between fp64 flushes
constexpr int NCLS = 64;                           // class slots of the one-hot tile (classes <= 64)
constexpr uint32_t TM_M = 0, TM_S = 128;

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
                                                            double* __restrict__ sum, double* __restrict__ second, int pf_on) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    int* cnt = reinterpret_cast<int*>(smem + OFF_CNT);
    const long long tiles = (n + 127) / 128;
    constexpr int MMA_WARP = 8;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(BAR(B_STAGED + b), 8); mbar_init(BAR(B_FREE + b), 1); }
        mbar_init(BAR(B_ACC), 1);
        mbar_init(BAR(B_DRAINED), 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 128) cnt[tid] = 0;
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(256));
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
            tc_fence_after();
            const uint32_t base = sbase + b * BUF;
            const uint64_t dHi = mn_desc(base + OFF_THI), dLo = mn_desc(base + OFF_TLO), dOh = mn_desc(base + OFF_OH);
            constexpr uint32_t IDM = idesc_bf16(128, 128) | MN_BOTH, IDS = idesc_bf16(128, NCLS) | MN_BOTH;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                       // K = 16 rows per step: two 8-row groups = 2 KiB
                const uint64_t o = (uint64_t)(ks * (2048 >> 4));
                const uint32_t first = (since_flush == 0 && ks == 0) ? 0u : 1u;
                if (leader) {
                    umma(tmem + TM_M, dHi + o, dHi + o, IDM, first);
                    umma(tmem + TM_M, dLo + o, dHi + o, IDM, 1u);
                    umma(tmem + TM_M, dHi + o, dLo + o, IDM, 1u);
                    umma(tmem + TM_M, dLo + o, dLo + o, IDM, 1u);
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
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;                     // accumulator row (feature) when flushing
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t free_parity[2] = {1, 1}, acc_parity = 0;
        long long it = 0;
        int since_flush = 0;
        // staging role: warp w owns rows 16 w .. 16 w + 15 of the tile; lane l holds features [4 l, 4 l + 4) of a row
        const int kc = lane >> 4, piece = (lane & 15) >> 1, sub = (lane & 1) * 8;
        const int oh_row = warp * 16 + (lane >> 1), oh_half = lane & 1;      // one-hot tile: two lanes per row, 4 pieces each
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            if (pf_on && warp == 0 && lane == 0) {                  // L2 prefetch of the tile 3 steps ahead (tc_ptx.cuh l2_prefetch_bulk)
                const long long tl = tile + 3 * (long long)gridDim.x;
                if (tl < tiles) {
                    const long long rows = n - tl * 128;
                    tc::l2_prefetch_bulk(feat + (size_t)tl * (128 * D), (uint32_t)((rows < 128 ? rows : 128) * D * 4));
                }
            }
            const long long r_oh = tile * 128 + oh_row;
            long long lab_oh = (r_oh < n) ? __ldg(labels + r_oh) : -1;
            if (lab_oh < 0 || lab_oh >= C) lab_oh = -1;              // rows with a label outside [0, C) are skipped entirely
            float4 t[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {                           // issued together with the label load above, not after it
                const long long r = tile * 128 + warp * 16 + j;
                t[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < n) t[j] = ld_stream(feat + r * D + 4 * lane);
            }
            mbar_wait(BAR(B_FREE + b), free_parity[b], 92);          // the MMAs that read this buffer are complete
            free_parity[b] ^= 1;
            uint8_t* buf = smem + b * BUF;
            const unsigned keep_mask = __ballot_sync(0xffffffffu, lab_oh >= 0);      // row 16 w + j <-> lanes 2j, 2j+1
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 x = ((keep_mask >> (2 * j)) & 1u) ? t[j] : make_float4(0.f, 0.f, 0.f, 0.f);     // unlabeled rows contribute nothing
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
                const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2bfloat162_rn(x.z - f23.x, x.w - f23.y);
                const uint32_t off = (uint32_t)(kc * CHUNK) + sw128_off(warp * 16 + j, piece) + sub;
                *reinterpret_cast<uint2*>(buf + OFF_THI + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                *reinterpret_cast<uint2*>(buf + OFF_TLO + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            }
            {   // one-hot row: 64 class slots = 8 pieces of 16 bytes, this lane writes pieces 4 oh_half .. 4 oh_half + 3
                const int lab = (int)lab_oh;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int p = 4 * oh_half + q;
                    uint32_t w[4] = {0u, 0u, 0u, 0u};
                    if (lab >= 0 && (lab >> 3) == p) w[(lab & 7) >> 1] = 0x3F80u << (16 * (lab & 1));      // bf16 1.0
                    *reinterpret_cast<uint4*>(buf + OFF_OH + sw128_off(oh_row, p)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                if (oh_half == 0 && lab >= 0) atomicAdd(cnt + lab, 1);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_STAGED + b));
            ++since_flush;
            const bool last = tile + gridDim.x >= tiles;
            if (since_flush == FLUSH_TILES || last) {
                // ---- flush: fp32 accumulators -> fp64 atomics (accumulator row = lane row, this thread's 64 columns)
                mbar_wait(BAR(B_ACC), acc_parity, 93);
                acc_parity ^= 1;
                tc_fence_after();
                uint32_t v[64];
                TMEM_LD32(lane_base + TM_M + 64 * half, v);
                TMEM_LD32(lane_base + TM_M + 64 * half + 32, (v + 32));
                tc_wait_ld();
#pragma unroll 8
                for (int j = 0; j < 64; ++j) atomicAdd(second + (size_t)row * D + 64 * half + j, (double)__uint_as_float(v[j]));
                TMEM_LD32(lane_base + TM_S + 32 * half, v);           // sum^T: lane = feature, column = class slot
                tc_wait_ld();
#pragma unroll 8
                for (int j = 0; j < 32; ++j) {
                    const int c = 32 * half + j;
                    if (c < C) atomicAdd(sum + (size_t)c * D + row, (double)__uint_as_float(v[j]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_DRAINED));
                since_flush = 0;
            }
        }
        // class counts: exact integers
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < C && cnt[tid] != 0) atomicAdd(count + tid, (double)cnt[tid]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
    }
}

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
    static int pf = -1;
    if (pf < 0) { const char* e = dev_getenv("CMHAR_L2_PREFETCH"); pf = e ? atoi(e) : 1; }      // development switch (default on)
    fittc::maha_fit_tc_kernel<<<grid, fittc::NT, fittc::SMEM_BYTES, st>>>(feat, labels, n, C, count, sum, second, pf);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
