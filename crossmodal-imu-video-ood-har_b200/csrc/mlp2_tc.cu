// mlp2_tc.cu -- one launch per projection head (CMHAR_BF16): Linear -> BatchNorm -> ReLU -> Linear -> L2 normalise,
// reference src/models/models.py:226-234 (ProjectionHead.forward) followed by :288-289 (F.normalize(dim=1)).
//
// The step used to run this as four launches per modality (two linear_tc tiles grids, l2_normalize, to_chunk_images
// inside cmhar_similarity) -- a dependency chain of ~40 us for 256 rows.  Here one CTA owns 128 rows and runs the whole
// chain with every intermediate in tensor memory:
//   x (bf16 operand image, K1/64 chunks)  --A ring (3 x 16 KiB, cp.async.bulk)-->
//   H  = x W0'^T            N = 512: four 128-column fp32 accumulators = all 512 TMEM columns (BatchNorm folded into W0', b0')
//   h  = relu(H + b0') as bf16, packed IN PLACE into columns [0,256) (block c -> [64c, 64c+64)), the A operand of
//   Y  = h W1^T             N = 256: two accumulators in columns [256,512), A from TMEM, W1 from the same weight ring
//   y  = (Y + b1) / max(|Y + b1|_2, 1e-12)  (row = TMEM lane = two threads, one partial-sum exchange through smem)
//   -> fp32 rows (optional) and the bf16 SWIZZLE_128B operand image the similarity kernel streams (optional).
// Weights: the linear blobs' bf16 chunk images ([n tile][k chunk], 16 KiB each) through an 8-stage ring.  Per CTA
// (K1 + 4 K1 + 16) chunks = 1.0 MB for the video head (K1 = 768), 0.4 MB for the IMU head: ingest-bound at ~48 B/clk.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace mlp2 {

using namespace tc;

constexpr int CHUNK = 16384;
constexpr int HID = 512, OUT = 256;
constexpr int NA = 3, NW = 8;
constexpr int OFF_A = 0, OFF_W = NA * CHUNK, OFF_MISC = OFF_W + NW * CHUNK, OFF_BAR = OFF_MISC + 1024, SMEM_BYTES = OFF_BAR + 256;
enum { B_AFULL = 0, B_AEMPTY = NA, B_WFULL = 2 * NA, B_WEMPTY = 2 * NA + NW, B_ACC1 = 2 * NA + 2 * NW, B_HRDY, B_ACC2, B_COUNT };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier area too small");
constexpr int NT = 8 * 32 + 64;

struct Args {
    const uint8_t* x_img;      // [row tile][kc1] chunk images
    int kc1;                   // K1 / 64
    const uint8_t* w0_img;     // [4 n tiles][kc1]
    const float* b0;           // (512) folded
    const uint8_t* w1_img;     // [2 n tiles][8]
    const float* b1;           // (256)
    long long n;
    int l2norm;
    float* y;                  // (n, 256) or null
    uint8_t* y_img;            // [row tile][4] chunk images or null
};

// hidden element k (0..511) as a bf16 TMEM A operand: block k/128 -> columns [64 block, +64), half (k%128)/64 -> +32
__host__ __device__ constexpr uint32_t hid_col(int k) { return (uint32_t)(64 * (k / 128) + 32 * ((k % 128) / 64) + (k % 64) / 2); }

__global__ void __launch_bounds__(NT, 1) mlp2_tc_kernel(const Args p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long trace_t0 = (tid == 0) ? trace_begin() : 0ull;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    const long long tile = blockIdx.x;
    const int kc1 = p.kc1;
    constexpr int MMA_WARP = 8, LOAD_WARP = 9;

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(BAR(B_AFULL + s), 1); mbar_init(BAR(B_AEMPTY + s), 1); }
        for (int s = 0; s < NW; ++s) { mbar_init(BAR(B_WFULL + s), 1); mbar_init(BAR(B_WEMPTY + s), 1); }
        mbar_init(BAR(B_ACC1), 1); mbar_init(BAR(B_HRDY), 8); mbar_init(BAR(B_ACC2), 1);
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

    if (warp == LOAD_WARP) {
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last();       // weights: re-read by every row tile
            uint32_t sa = 0, pa = 1, sw = 0, pw = 1;
            auto push_w = [&](const uint8_t* src) {
                mbar_wait(BAR(B_WEMPTY + sw), pw, 90);
                mbar_expect_tx(BAR(B_WFULL + sw), CHUNK);
                bulk_g2s_hint(sbase + OFF_W + sw * CHUNK, src, CHUNK, BAR(B_WFULL + sw), keep);
                if (++sw == NW) { sw = 0; pw ^= 1; }
            };
            for (int kc = 0; kc < kc1; ++kc) {
                mbar_wait(BAR(B_AEMPTY + sa), pa, 91);
                mbar_expect_tx(BAR(B_AFULL + sa), CHUNK);
                bulk_g2s(sbase + OFF_A + sa * CHUNK, p.x_img + ((size_t)tile * kc1 + kc) * CHUNK, CHUNK, BAR(B_AFULL + sa));
                if (++sa == NA) { sa = 0; pa ^= 1; }
                for (int nb = 0; nb < 4; ++nb) push_w(p.w0_img + ((size_t)nb * kc1 + kc) * CHUNK);
            }
            for (int kc = 0; kc < HID / 64; ++kc)
                for (int nb = 0; nb < 2; ++nb) push_w(p.w1_img + ((size_t)nb * (HID / 64) + kc) * CHUNK);
        }
    } else if (warp == MMA_WARP) {
        const bool leader = elect_one();
        constexpr uint32_t ID128 = idesc_bf16(128, 128);
        uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
        for (int kc = 0; kc < kc1; ++kc) {
            mbar_wait(BAR(B_AFULL + sa), pa, 92);
            const uint64_t dA = sw128_desc(sbase + OFF_A + sa * CHUNK);
            for (int nb = 0; nb < 4; ++nb) {
                mbar_wait(BAR(B_WFULL + sw), pw, 93);
                tc_fence_after();
                const uint64_t dW = sw128_desc(sbase + OFF_W + sw * CHUNK);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (leader) umma(tmem + 128 * nb, dA + (uint64_t)(2 * k), dW + (uint64_t)(2 * k), ID128, (kc > 0 || k > 0) ? 1u : 0u);
                if (leader) tc_commit(BAR(B_WEMPTY + sw));
                if (++sw == NW) { sw = 0; pw ^= 1; }
            }
            if (leader) tc_commit(BAR(B_AEMPTY + sa));
            if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        if (leader) tc_commit(BAR(B_ACC1));
        mbar_wait(BAR(B_HRDY), 0, 94);                      // hidden activation is back in TMEM as bf16
        tc_fence_after();
        for (int kc = 0; kc < HID / 64; ++kc) {
            for (int nb = 0; nb < 2; ++nb) {
                mbar_wait(BAR(B_WFULL + sw), pw, 95);
                tc_fence_after();
                const uint64_t dW = sw128_desc(sbase + OFF_W + sw * CHUNK);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (leader) umma_ts(tmem + 256 + 128 * nb, tmem + hid_col(64 * kc + 16 * k), dW + (uint64_t)(2 * k), ID128, (kc > 0 || k > 0) ? 1u : 0u);
                if (leader) tc_commit(BAR(B_WEMPTY + sw));
                if (++sw == NW) { sw = 0; pw ^= 1; }
            }
        }
        if (leader) tc_commit(BAR(B_ACC2));
    } else {
        // ------------------------------------------------------------- epilogue: thread = (row, column half)
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const long long r = tile * 128 + row;
        const bool ok = r < p.n;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        auto quarter_bar = [&] { asm volatile("bar.sync %0, 64;" ::"r"(2 + (warp & 3)) : "memory"); };   // the two warps sharing these rows
        mbar_wait(BAR(B_ACC1), 0, 96);
        tc_fence_after();
        {
            uint32_t v[64];
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                TMEM_LD32(lane_base + 128 * c + 64 * half, v);
                TMEM_LD32(lane_base + 128 * c + 64 * half + 32, (v + 32));
                tc_wait_ld();
                // block c's bf16 image lands in [64c, 64c+64), which overlaps fp32 columns of block c/2 owned by the OTHER
                // half's thread: both threads of a row must have finished loading this block before either stores
                quarter_bar();
                const float* b = p.b0 + 128 * c + 64 * half;
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 64; i += 4) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + i));
                    pk[i / 2] = pack_bf16_relu(__uint_as_float(v[i]) + bb.x, __uint_as_float(v[i + 1]) + bb.y);
                    pk[i / 2 + 1] = pack_bf16_relu(__uint_as_float(v[i + 2]) + bb.z, __uint_as_float(v[i + 3]) + bb.w);
                }
                TMEM_ST32(lane_base + 64 * c + 32 * half, pk);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_HRDY));
        }
        mbar_wait(BAR(B_ACC2), 0, 97);
        tc_fence_after();
        {
            uint32_t o[128];
#pragma unroll
            for (int q = 0; q < 4; ++q) TMEM_LD32(lane_base + 256 + 128 * half + 32 * q, (o + 32 * q));
            tc_wait_ld();
            float f[128];
            float ss = 0.f;
            const float* b = p.b1 + 128 * half;
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b + i));
                f[i] = __uint_as_float(o[i]) + bb.x; f[i + 1] = __uint_as_float(o[i + 1]) + bb.y;
                f[i + 2] = __uint_as_float(o[i + 2]) + bb.z; f[i + 3] = __uint_as_float(o[i + 3]) + bb.w;
                ss = fmaf(f[i], f[i], ss); ss = fmaf(f[i + 1], f[i + 1], ss); ss = fmaf(f[i + 2], f[i + 2], ss); ss = fmaf(f[i + 3], f[i + 3], ss);
            }
            if (p.l2norm) {
                float* part = reinterpret_cast<float*>(smem + OFF_MISC);
                part[half * 128 + row] = ss;
                quarter_bar();
                const float tot = part[row] + part[128 + row];                      // fixed order: identical in both threads
                const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);                  // F.normalize: x / max(|x|_2, eps)
#pragma unroll
                for (int i = 0; i < 128; ++i) f[i] *= inv;
            }
            if (p.y && ok) {
                float4* dst = reinterpret_cast<float4*>(p.y + r * OUT + 128 * half);
#pragma unroll
                for (int i = 0; i < 32; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            }
            if (p.y_img) {
                if (!ok) {
#pragma unroll
                    for (int i = 0; i < 128; ++i) f[i] = 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint8_t* chunk = p.y_img + ((size_t)tile * (OUT / 64) + 2 * half + (q >> 1)) * CHUNK;
                    store_bf16_32(chunk, row, (q & 1) * 4, f + 32 * q);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
    if (tid == 0) trace_end(TRACE_LINEAR, trace_t0);
}

}  // namespace mlp2

int launch_mlp2_tc(const uint8_t* x_img, int K1, const uint8_t* w0_img, const float* b0, const uint8_t* w1_img, const float* b1,
                   long long n, int l2norm, float* y, uint8_t* y_img, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(mlp2::mlp2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mlp2::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    mlp2::Args p{x_img, K1 / 64, w0_img, b0, w1_img, b1, n, l2norm, y, y_img};
    const long long tiles = (n + 127) / 128;
    CMHAR_REQUIRE(tiles <= 0x7fffffffLL, "too many rows");
    mlp2::mlp2_tc_kernel<<<(unsigned)tiles, mlp2::NT, mlp2::SMEM_BYTES, st>>>(p);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
