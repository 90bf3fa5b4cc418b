// maha_fit_tc.cu -- Mahalanobis sufficient statistics on the tensor cores (spec row A3, CMHAR_BF16 path).
//
// SPEC-DERIVED (no reference implementation): count_c, sum_c = sum_{label = c} f and the second
// moment M = sum f f^T over feature rows f in R^128; the tied covariance is finalised on the host in fp64 after
// the NCCL all-reduce of these three buffers (ood.py).
//
// The CUDA-core kernel (ood.cu) is fp32-FMA bound at 4 % of the HBM roofline the stage should sit on (AI = 64
// flop/B).  Here both reductions are GEMMs over the ROW dimension, K = rows:
//     M       (128 x 128) += F^T F          A = B = F^T tile
//     sum     (128 x 128) += O^T F          A = one-hot(labels)^T tile (class slots x rows), B = F^T tile
// with F split as hi + lo bf16 (all four hi/lo products for M -- the lo*lo term removes the only systematic bias,
// on the diagonal -- and both for the class sums; one-hot entries are exact in bf16), fp32 accumulation in TMEM
// across all the tiles of a CTA, flushed with fp64 atomics every 256 tiles.
// F^T has rows contiguous along K, i.e. it is the TRANSPOSE of the feature tile in memory: the 8 staging warps
// read feature rows (coalesced float4) and scatter 2-byte elements into K-major SWIZZLE_128B tiles -- a warp's 32
// lanes hold consecutive rows of the same feature, so each store instruction covers 64 contiguous bytes.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace fittc {

using namespace tc;

constexpr int CHUNK = 16384;                       // [128 m-rows x 64 k] bf16
constexpr int TILE = 2 * CHUNK;                    // 128 rows of K
constexpr int OFF_THI = 0, OFF_TLO = TILE, OFF_OH = 2 * TILE, BUF = 3 * TILE;     // one staging buffer = 96 KiB
constexpr int OFF_CNT = 2 * BUF;                   // int counters [128]
constexpr int OFF_BAR = OFF_CNT + 512;
constexpr int SMEM_BYTES = OFF_BAR + 128;
enum { B_STAGED = 0, B_FREE = 2, B_ACC = 4, B_DRAINED = 5, B_COUNT = 6 };
constexpr int NT = 8 * 32 + 32;                    // 8 staging warps + the MMA warp
constexpr int FLUSH_TILES = 256;                   // 32 768 rows of fp32 accumulation between fp64 flushes
constexpr uint32_t TM_M = 0, TM_S = 128;

// byte offset of element (m, k) inside a [128 x 128] K-major SW128 tile made of two [128 x 64] chunks
__device__ __forceinline__ uint32_t t_off(int m, int k) { return (uint32_t)((k >> 6) * CHUNK) + sw128_off(m, (k & 63) >> 3) + (uint32_t)((k & 7) * 2); }

__global__ void __launch_bounds__(NT, 1) maha_fit_tc_kernel(const float* __restrict__ feat, const long long* __restrict__ labels,
                                                            long long n, int C, double* __restrict__ count,
                                                            double* __restrict__ sum, double* __restrict__ second) {
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
        constexpr uint32_t ID128 = idesc_bf16(128, 128);
        uint32_t staged_parity[2] = {0, 0}, drained_parity = 0;
        long long it = 0;
        int since_flush = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            mbar_wait(BAR(B_STAGED + b), staged_parity[b], 90);
            staged_parity[b] ^= 1;
            tc_fence_after();
            const uint32_t base = sbase + b * BUF;
#pragma unroll 1
            for (int kc = 0; kc < 2; ++kc) {
                const uint64_t dHi = sw128_desc(base + OFF_THI + kc * CHUNK), dLo = sw128_desc(base + OFF_TLO + kc * CHUNK),
                               dOh = sw128_desc(base + OFF_OH + kc * CHUNK);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t o = (uint64_t)(2 * k);
                    const uint32_t first = (since_flush == 0 && kc == 0 && k == 0) ? 0u : 1u;
                    if (leader) {
                        umma(tmem + TM_M, dHi + o, dHi + o, ID128, first);
                        umma(tmem + TM_M, dLo + o, dHi + o, ID128, 1u);
                        umma(tmem + TM_M, dHi + o, dLo + o, ID128, 1u);
                        umma(tmem + TM_M, dLo + o, dLo + o, ID128, 1u);
                        umma(tmem + TM_S, dOh + o, dHi + o, ID128, first);
                        umma(tmem + TM_S, dOh + o, dLo + o, ID128, 1u);
                    }
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
        const int row = (warp & 3) * 32 + lane;                     // row inside the tile == accumulator row when flushing
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t free_parity[2] = {1, 1}, acc_parity = 0;
        long long it = 0;
        int since_flush = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            const long long r = tile * 128 + row;
            const bool ok = r < n;
            float4 t[16];
            const float4* src = reinterpret_cast<const float4*>(feat + r * D + 64 * half);
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = ok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            long long lab = ok ? __ldg(labels + r) : -1;
            if (lab < 0 || lab >= C) lab = -1;                       // rows with a label outside [0, C) are skipped entirely
            if (lab < 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(BAR(B_FREE + b), free_parity[b], 92);          // the MMAs that read this buffer are complete
            free_parity[b] ^= 1;
            uint8_t* buf = smem + b * BUF;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float x4[4] = {t[i].x, t[i].y, t[i].z, t[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int f = 64 * half + 4 * i + e;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(x4[e]);
                    const __nv_bfloat16 lo = __float2bfloat16_rn(x4[e] - __bfloat162float(hi));
                    const uint32_t off = t_off(f, row);
                    *reinterpret_cast<__nv_bfloat16*>(buf + OFF_THI + off) = hi;
                    *reinterpret_cast<__nv_bfloat16*>(buf + OFF_TLO + off) = lo;
                }
            }
            const unsigned short one = 0x3F80, zero = 0;             // bf16 1.0
#pragma unroll 16
            for (int c = 64 * half; c < 64 * half + 64; ++c)
                *reinterpret_cast<unsigned short*>(buf + OFF_OH + t_off(c, row)) = (c == (int)lab) ? one : zero;
            if (half == 0 && lab >= 0) atomicAdd(cnt + (int)lab, 1);
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
                TMEM_LD32(lane_base + TM_S + 64 * half, v);
                TMEM_LD32(lane_base + TM_S + 64 * half + 32, (v + 32));
                tc_wait_ld();
                if (row < C) {
#pragma unroll 8
                    for (int j = 0; j < 64; ++j) atomicAdd(sum + (size_t)row * D + 64 * half + j, (double)__uint_as_float(v[j]));
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
    fittc::maha_fit_tc_kernel<<<grid, fittc::NT, fittc::SMEM_BYTES, st>>>(feat, labels, n, C, count, sum, second);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
