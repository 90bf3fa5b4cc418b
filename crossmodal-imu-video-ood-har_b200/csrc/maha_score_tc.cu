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
//   * 8 staging warps: coalesced 512-byte row loads (16 rows in flight per warp), split into hi / lo bf16 and
//     stored as K-major SWIZZLE_128B A tiles (8-byte stores, conflict-free), double buffered; the loads of tile
//     i+1 are issued while tile i is being converted;
//   * 1 MMA warp: 24 tcgen05.mma (M=128, N=160, K=16) per tile against the [W | G] image that stays RESIDENT in
//     shared memory (80 KiB, loaded once per CTA by cp.async.bulk) into one of two TMEM accumulators;
//   * 4 epilogue warps: thread = row; |y|^2 over 128 columns, min over the 32 class columns, one 4-byte store.
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
constexpr int N_STAGE_WARPS = 8, EPI_WARP0 = 8, MMA_WARP = 12, NT = 13 * 32;
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

__global__ void __launch_bounds__(NT, 1) maha_score_tc_kernel(const uint8_t* __restrict__ section, const float* __restrict__ feat,
                                                              long long n, float* __restrict__ score) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    const long long tiles = (n + 127) / 128;

    if (tid == 0) {
        mbar_init(BAR(B_W), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(B_STAGED + b), N_STAGE_WARPS);
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

    if (warp == MMA_WARP) {
        // ================================================================= MMA issuer
        const bool leader = elect_one();
        if (leader) {                                   // the [W | G] image: resident for the life of the CTA
            mbar_expect_tx(BAR(B_W), W_BYTES);
            for (int q = 0; q < 4; ++q) bulk_g2s(sbase + OFF_W + q * BCH, section + (size_t)q * BCH, BCH, BAR(B_W));
        }
        __syncwarp();
        mbar_wait(BAR(B_W), 0, 80);
        constexpr uint32_t ID160 = idesc_bf16(128, NB);
        uint32_t staged_par = 0u, accfree_par = 3u;      // one parity bit per buffer; fresh barrier: waiting on parity 1 passes
        long long it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            mbar_wait(BAR(B_ACCFREE + b), (accfree_par >> b) & 1u, 81);
            accfree_par ^= 1u << b;
            mbar_wait(BAR(B_STAGED + b), (staged_par >> b) & 1u, 82);
            staged_par ^= 1u << b;
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
        }
    } else if (warp < N_STAGE_WARPS) {
        // ================================================================= staging: fp32 rows -> split-bf16 A tiles
        // warp w owns rows 16 w .. 16 w + 15 of the tile; lane l holds elements [4 l, 4 l + 4) of a row
        const int kc = lane >> 4, piece = (lane & 15) >> 1, sub = (lane & 1) * 8;
        float4 t[16];
        auto load_tile = [&](long long tile) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const long long r = tile * 128 + warp * 16 + j;
                t[j] = (r < n) ? ld_stream(feat + r * D + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        uint32_t free_par = 3u;
        long long it = 0;
        if ((long long)blockIdx.x < tiles) load_tile(blockIdx.x);
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            mbar_wait(BAR(B_FREE + b), (free_par >> b) & 1u, 83);    // the MMAs that read this buffer are complete
            free_par ^= 1u << b;
            uint8_t* abuf = smem + OFF_A + b * ABUF + kc * ACH;
            const long long next = tile + gridDim.x;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int row = warp * 16 + j;
                const float4 x = t[j];
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
                const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2bfloat162_rn(x.z - f23.x, x.w - f23.y);
                const uint32_t off = sw128_off(row, piece) + sub;
                *reinterpret_cast<uint2*>(abuf + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                *reinterpret_cast<uint2*>(abuf + 2 * ACH + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                if (next < tiles) {                                  // refill this register with the next tile's row right away
                    const long long r = next * 128 + row;
                    t[j] = (r < n) ? ld_stream(feat + r * D + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_STAGED + b));
        }
    } else {
        // ================================================================= epilogue: thread = row
        const int q = warp - EPI_WARP0;                              // == warp & 3: the TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        float* mun = reinterpret_cast<float*>(smem + OFF_MUN);       // |mu_c W|^2, read back as warp-uniform broadcasts
        if (q == 0) mun[lane] = __ldg(reinterpret_cast<const float*>(section + W_BYTES) + lane);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t full_par = 0u;
        long long it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int b = (int)(it & 1);
            mbar_wait(BAR(B_ACCFULL + b), (full_par >> b) & 1u, 84);
            full_par ^= 1u << b;
            tc_fence_after();
            const uint32_t acc = lane_base + ACC_STRIDE * (uint32_t)b;
            uint32_t v[64];
            float yy0 = 0.f, yy1 = 0.f, yy2 = 0.f, yy3 = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                TMEM_LD32(acc + 64 * h, v);
                TMEM_LD32(acc + 64 * h + 32, (v + 32));
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 64; i += 4) {
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
            if (lane == 0) mbar_arrive(BAR(B_ACCFREE + b));          // accumulator b may be overwritten
            float best = INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c) best = fminf(best, fmaf(-2.f, __uint_as_float(v[c]), mun[c]));
            const long long r = tile * 128 + row;
            if (r < n) score[r] = fmaxf((yy0 + yy1) + (yy2 + yy3) + best, 0.f);
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
int launch_maha_score_tc(const uint8_t* section, const float* feat, long long n, float* score, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(mstc::maha_score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mstc::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    const long long tiles = (n + 127) / 128;
    const int grid = (int)(tiles < (long long)sm_count() ? tiles : (long long)sm_count());
    mstc::maha_score_tc_kernel<<<grid, mstc::NT, mstc::SMEM_BYTES, st>>>(section, feat, n, score);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
