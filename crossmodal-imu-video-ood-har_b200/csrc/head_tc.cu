// head_tc.cu -- classifier head + MSP / energy / Mahalanobis scores on the tensor cores (tcgen05), the
// north-star "OOD scoring stage as one fused GEMM-plus-reduction kernel".
//
// Same algebra as head.cu (reference src/models/models.py:312-326,338 with BatchNorm folded; spec rows A1, A2,
// A4 of oracle/ood_spec.py).  On CUDA cores this stage is fp32-FMA bound at ~20 TFLOP/s -- 1-3 % of the HBM
// roofline it should sit on (AI = 214 flop/B for the head, 79 for the Mahalanobis score).  Here
// every layer is a chain of tcgen05.mma with SPLIT-bf16 operands: x = hi + lo (hi = bf16(x), lo = bf16(x - hi))
// on both sides and D += A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation in TMEM, i.e. products are
// exact to ~2^-17 relative -- fp32-grade logits, so arg-max does not depend on bf16 rounding.
//
// One CTA owns 128 feature rows:
//   F (fp32, global) --split--> smem A tiles (hi | lo)
//   Y  = F  Wh            N=128   (whitening, if a Mahalanobis state is attached)     TMEM [256,384)
//   H1 = relu(F  W0 + b0) N=256                                                      TMEM [0,256)
//   H2 = relu(H1 W1 + b1) N=128   A operand = H1 from TMEM                           TMEM [384,512)
//   Z  = H2 W2 + b2       N=32    A operand = H2 from TMEM                           TMEM [0,32)
//   G  = Y  Mu^T          N=32    A operand = Y  from TMEM                           TMEM [32,64)
//   dist_c = |y|^2 - 2 G_c + |mu_c|^2 ; scores from Z and dist in registers.
// Activations never touch shared memory after the first layer: the epilogue writes (hi | lo) bf16 pairs back
// into the TMEM columns of the accumulator slice it just read (a thread's 64 fp32 columns become 32 columns of
// hi and 32 of lo) and the next GEMM reads them as TMEM A operands.  Weights are pre-split bf16 SWIZZLE_128B
// images in the blobs ([hi chunk | lo chunk] per 64-wide k block) streamed by cp.async.bulk through a 4 x 32 KiB
// mbarrier ring: 352 KiB per tile, the binding resource (L2 -> SM ingest ~48 B/clk) next to the 9.8 K cycles
// of MMA issue.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace headtc {

using namespace tc;

constexpr int H1 = 256, H2 = 128, CP = 32;           // the reference head (classes padded to 32)
constexpr int STAGE = 32768;                          // ring slot: [hi chunk 128x64 | lo chunk 128x64]
constexpr int NSLOT = 4;
constexpr int OFF_FHI = 0, OFF_FLO = 32768;           // F as SS A operand, 2 chunks each
constexpr int OFF_RING = 65536;
constexpr int OFF_MISC = OFF_RING + NSLOT * STAGE;    // ynorm partials [2][128] floats
constexpr int OFF_BAR = OFF_MISC + 1024;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB");
enum { B_FULL = 0, B_EMPTY = B_FULL + NSLOT, B_F = B_EMPTY + NSLOT, B_ACC_Y, B_ACC_H1A, B_ACC_H1B, B_ACC_H2, B_ACC_OUT,
       B_Y_RDY, B_H1_RDY, B_H2_RDY, B_ACC_PRE, B_COUNT };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier area too small");
constexpr uint32_t TM_H1 = 0, TM_Y = 256, TM_H2 = 384, TM_Z = 0, TM_G = 32;
constexpr int NT = 8 * 32 + 64;

// image sizes (bytes): a "stage" is [hi | lo] of one (n block, k chunk)
constexpr int IMG_W0 = 4 * STAGE, IMG_W1 = 4 * STAGE, IMG_W2 = 2 * 8192;           // head section
constexpr int IMG_WH = 2 * STAGE, IMG_MU = 2 * 8192;                               // maha section
constexpr size_t HEAD_TC_BYTES = IMG_W0 + IMG_W1 + IMG_W2;
constexpr size_t MAHA_TC_BYTES = IMG_WH + IMG_MU + CP * sizeof(float);             // + |mu_c|^2 (inf for empty classes)

// bf16 A operand written back over a 64-column fp32 slice: element e (0..63) of the slice -> hi at column e/2,
// lo at column 32 + e/2.  Element k of a 128-wide block lives in slice k/64.
__host__ __device__ constexpr uint32_t col_hi(int k) { return (uint32_t)(128 * (k / 128) + 64 * ((k % 128) / 64) + (k % 64) / 2); }
__host__ __device__ constexpr uint32_t col_lo(int k) { return col_hi(k) + 32; }

// src: (K,N) row-major when !src_is_nk (element (n,k) = src[k*ld + n]), (N,K) row-major otherwise.
// dst stage = [hi chunk rows_img x 64 | lo chunk], SWIZZLE_128B K-major; rows >= n_valid are zero.
__global__ void pack_split_stage_kernel(const float* __restrict__ src, int ld, int src_is_nk, int n0, int n_valid, int rows_img,
                                        int k0, uint8_t* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per (row, 16-byte piece)
    if (i >= rows_img * 8) return;
    const int r = i >> 3, j = i & 7;
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int n = n0 + r, k = k0 + j * 8 + e;
        const float v = (n < n_valid) ? (src_is_nk ? src[(size_t)n * ld + k] : src[(size_t)k * ld + n]) : 0.f;
        hi[e] = __bfloat162float(__float2bfloat16_rn(v));
        lo[e] = v - hi[e];
    }
    uint4 u, w;
    u.x = pack_bf16(hi[0], hi[1]); u.y = pack_bf16(hi[2], hi[3]); u.z = pack_bf16(hi[4], hi[5]); u.w = pack_bf16(hi[6], hi[7]);
    w.x = pack_bf16(lo[0], lo[1]); w.y = pack_bf16(lo[2], lo[3]); w.z = pack_bf16(lo[4], lo[5]); w.w = pack_bf16(lo[6], lo[7]);
    *reinterpret_cast<uint4*>(dst + sw128_off(r, j)) = u;
    *reinterpret_cast<uint4*>(dst + rows_img * 128 + sw128_off(r, j)) = w;
}

__global__ void mu_norm_kernel(const float* __restrict__ mean_w, const float* __restrict__ valid, int C, float* __restrict__ dst) {
    const int c = threadIdx.x;
    if (c >= CP) return;
    float s = INFINITY;
    if (c < C && valid[c] > 0.f) {
        s = 0.f;
        for (int k = 0; k < D; ++k) s = fmaf(mean_w[c * D + k], mean_w[c * D + k], s);
    }
    dst[c] = s;
}

struct Args {
    FwdArgs f;
    const uint8_t* head_tc;      // W0 | W1 | W2 images, or null
    const float *b0, *b1, *b2;   // folded fp32 biases (b2: Cp() entries)
    const uint8_t* maha_tc;      // Wh | Mu images | mu norms, or null
    int classes;
    // optional late-fusion pre-layer (spec row A6): F = relu([x1 | x2] Wf'^T + bf') computed here from bf16 operand images
    // (plain bf16 MMAs, fp32 accumulate) instead of being read from global memory; f.x is unused then
    const uint8_t* pre_w;        // [kc] chunk images of Wf' (128 output rows), or null
    const float* pre_bias;       // (128) folded
    const uint8_t* pre_x1;       // [row tile][pre_kc1] chunk images
    const uint8_t* pre_x2;       // [row tile][pre_kc2] chunk images
    int pre_kc1, pre_kc2;
    float* fused_out;            // (n,128) fp32 rows of F, or null
};

__global__ void __launch_bounds__(NT, 1) head_tc_kernel(const Args p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long trace_t0 = (tid == 0) ? trace_begin() : 0ull;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    const FwdArgs& a = p.f;
    const bool do_head = p.head_tc != nullptr, do_maha = p.maha_tc != nullptr;
    const long long tiles = (a.n + 127) / 128;
    constexpr int MMA_WARP = 8, LOAD_WARP = 9;

    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
        mbar_init(BAR(B_F), 8);
        for (int i = B_ACC_Y; i <= B_ACC_OUT; ++i) mbar_init(BAR(i), 1);
        mbar_init(BAR(B_Y_RDY), 8); mbar_init(BAR(B_H1_RDY), 8); mbar_init(BAR(B_H2_RDY), 8);
        mbar_init(BAR(B_ACC_PRE), 1);
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
        // ================================================================= weight producer
        if (lane == 0) {
            uint32_t slot = 0, parity = 1;
            auto push = [&](const uint8_t* src, uint32_t bytes) {
                mbar_wait(BAR(B_EMPTY + slot), parity, 60);
                mbar_expect_tx(BAR(B_FULL + slot), bytes);
                bulk_g2s(sbase + OFF_RING + slot * STAGE, src, bytes, BAR(B_FULL + slot));
                if (++slot == NSLOT) { slot = 0; parity ^= 1; }
            };
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                if (p.pre_w) {          // pre-layer: one slot per 64-wide k chunk = [activation chunk | weight chunk]
                    const int kcn = p.pre_kc1 + p.pre_kc2;
                    for (int kc = 0; kc < kcn; ++kc) {
                        const uint8_t* xa = (kc < p.pre_kc1) ? p.pre_x1 + ((size_t)tile * p.pre_kc1 + kc) * 16384
                                                             : p.pre_x2 + ((size_t)tile * p.pre_kc2 + (kc - p.pre_kc1)) * 16384;
                        mbar_wait(BAR(B_EMPTY + slot), parity, 72);
                        mbar_expect_tx(BAR(B_FULL + slot), STAGE);
                        bulk_g2s(sbase + OFF_RING + slot * STAGE, xa, 16384, BAR(B_FULL + slot));
                        bulk_g2s(sbase + OFF_RING + slot * STAGE + 16384, p.pre_w + (size_t)kc * 16384, 16384, BAR(B_FULL + slot));
                        if (++slot == NSLOT) { slot = 0; parity ^= 1; }
                    }
                }
                if (do_maha) for (int s = 0; s < 2; ++s) push(p.maha_tc + s * STAGE, STAGE);
                if (do_head) {
                    for (int s = 0; s < 4; ++s) push(p.head_tc + s * STAGE, STAGE);                      // W0: (h0,k0) (h0,k1) (h1,k0) (h1,k1)
                    for (int s = 0; s < 4; ++s) push(p.head_tc + IMG_W0 + s * STAGE, STAGE);             // W1: k0..k3
                    for (int s = 0; s < 2; ++s) push(p.head_tc + IMG_W0 + IMG_W1 + s * 8192, 8192);      // W2: k0, k1 (32 rows)
                }
                if (do_maha) for (int s = 0; s < 2; ++s) push(p.maha_tc + IMG_WH + s * 8192, 8192);      // Mu: k0, k1 (32 rows)
            }
        }
    } else if (warp == MMA_WARP) {
        // ================================================================= MMA issuer
        const bool leader = elect_one();
        constexpr uint32_t ID128 = idesc_bf16(128, 128), ID32 = idesc_bf16(128, 32);
        uint32_t slot = 0, parity = 0;
        Phase ph;
        // one ring stage = 64 k columns = 4 k-steps x 3 split terms.  A from smem (hi/lo descriptors) ...
        auto stage_ss = [&](uint32_t d, uint64_t a_hi, uint64_t a_lo, uint32_t idesc, int rows_img, bool first) {
            mbar_wait(BAR(B_FULL + slot), parity, 61);
            tc_fence_after();
            const uint64_t b_hi = sw128_desc(sbase + OFF_RING + slot * STAGE), b_lo = sw128_desc(sbase + OFF_RING + slot * STAGE + rows_img * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (leader) {
                    umma(d, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
                    umma(d, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1u);
                    umma(d, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, 1u);
                }
            }
            if (leader) tc_commit(BAR(B_EMPTY + slot));
            if (++slot == NSLOT) { slot = 0; parity ^= 1; }
        };
        // ... or from TMEM: the stage covers elements [k_first, k_first + 64) of the activation row held at a_base
        auto stage_ts = [&](uint32_t d, uint32_t a_base, int k_first, uint32_t idesc, int rows_img, bool first) {
            mbar_wait(BAR(B_FULL + slot), parity, 62);
            tc_fence_after();
            const uint64_t b_hi = sw128_desc(sbase + OFF_RING + slot * STAGE), b_lo = sw128_desc(sbase + OFF_RING + slot * STAGE + rows_img * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (leader) {
                    const uint32_t ahi = a_base + col_hi(k_first + 16 * k), alo = a_base + col_lo(k_first + 16 * k);
                    umma_ts(d, ahi, b_hi + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
                    umma_ts(d, alo, b_hi + (uint64_t)(2 * k), idesc, 1u);
                    umma_ts(d, ahi, b_lo + (uint64_t)(2 * k), idesc, 1u);
                }
            }
            if (leader) tc_commit(BAR(B_EMPTY + slot));
            if (++slot == NSLOT) { slot = 0; parity ^= 1; }
        };
        const uint64_t dFhi = sw128_desc(sbase + OFF_FHI), dFlo = sw128_desc(sbase + OFF_FLO);
        const uint64_t CH = 16384 >> 4;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            if (p.pre_w) {
                // accumulates in the H2 columns: free until this tile's H1 is done, and the previous tile's Z / G MMAs that read
                // H2 / Y from TMEM precede these in the (in-order) tensor pipe
                const int kcn = p.pre_kc1 + p.pre_kc2;
                for (int kc = 0; kc < kcn; ++kc) {
                    mbar_wait(BAR(B_FULL + slot), parity, 73);
                    tc_fence_after();
                    const uint64_t dA = sw128_desc(sbase + OFF_RING + slot * STAGE), dB = sw128_desc(sbase + OFF_RING + slot * STAGE + 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (leader) umma(tmem + TM_H2, dA + (uint64_t)(2 * k), dB + (uint64_t)(2 * k), ID128, (kc > 0 || k > 0) ? 1u : 0u);
                    if (leader) tc_commit(BAR(B_EMPTY + slot));
                    if (++slot == NSLOT) { slot = 0; parity ^= 1; }
                }
                if (leader) tc_commit(BAR(B_ACC_PRE));
            }
            mbar_wait(BAR(B_F), ph.next(B_F), 63);                 // F staged (and the previous tile's TMEM fully consumed)
            tc_fence_after();
            if (do_maha) {
                for (int kc = 0; kc < 2; ++kc) stage_ss(tmem + TM_Y, dFhi + kc * CH, dFlo + kc * CH, ID128, 128, kc == 0);
                if (leader) tc_commit(BAR(B_ACC_Y));
            }
            if (do_head) {
                for (int h = 0; h < 2; ++h) {
                    for (int kc = 0; kc < 2; ++kc) stage_ss(tmem + TM_H1 + 128 * h, dFhi + kc * CH, dFlo + kc * CH, ID128, 128, kc == 0);
                    if (leader) tc_commit(BAR(h == 0 ? B_ACC_H1A : B_ACC_H1B));
                }
                mbar_wait(BAR(B_H1_RDY), ph.next(B_H1_RDY), 64);
                tc_fence_after();
                for (int kc = 0; kc < 4; ++kc) stage_ts(tmem + TM_H2, tmem + TM_H1, 64 * kc, ID128, 128, kc == 0);
                if (leader) tc_commit(BAR(B_ACC_H2));
                mbar_wait(BAR(B_H2_RDY), ph.next(B_H2_RDY), 65);
                tc_fence_after();
                for (int kc = 0; kc < 2; ++kc) stage_ts(tmem + TM_Z, tmem + TM_H2, 64 * kc, ID32, 32, kc == 0);
            }
            if (do_maha) {
                mbar_wait(BAR(B_Y_RDY), ph.next(B_Y_RDY), 66);
                tc_fence_after();
                for (int kc = 0; kc < 2; ++kc) stage_ts(tmem + TM_G, tmem + TM_Y, 64 * kc, ID32, 32, kc == 0);
            }
            if (leader) tc_commit(BAR(B_ACC_OUT));
        }
    } else {
        // ================================================================= epilogue: thread = (row, 64-column half)
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int c0 = 64 * half;
        float* ynorm = reinterpret_cast<float*>(smem + OFF_MISC);          // [2][128]
        Phase ph;
        uint32_t v[64];
        auto load64 = [&](uint32_t taddr) {
            TMEM_LD32(taddr, v);
            TMEM_LD32(taddr + 32, (v + 32));
            tc_wait_ld();
        };
        // 64 finished fp32 values -> (hi | lo) bf16 pairs over the slice they came from
        auto store_split = [&](uint32_t taddr, const float* y) {
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(y[2 * i]), h1 = __float2bfloat16_rn(y[2 * i + 1]);
                hi[i] = pack_bf16(__bfloat162float(h0), __bfloat162float(h1));
                lo[i] = pack_bf16(y[2 * i] - __bfloat162float(h0), y[2 * i + 1] - __bfloat162float(h1));
            }
            TMEM_ST32(taddr, hi);
            TMEM_ST32(taddr + 32, lo);
        };
        auto arrive_tmem = [&](int bar_idx) {
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(bar_idx));
        };
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const long long r = tile * 128 + row;
            const bool ok = r < a.n;
            // ---- stage F: this thread's 64 features -> hi / lo bf16 rows of the SS A tiles
            {
                float4 t[16];
                if (p.pre_w) {      // F = relu(pre-layer accumulator + bias): this thread's 64 columns of the H2 buffer
                    mbar_wait(BAR(B_ACC_PRE), ph.next(B_ACC_PRE), 74);
                    tc_fence_after();
                    load64(lane_base + TM_H2 + c0);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.pre_bias + c0) + i);
                        t[i] = make_float4(fmaxf(__uint_as_float(v[4 * i]) + b.x, 0.f), fmaxf(__uint_as_float(v[4 * i + 1]) + b.y, 0.f),
                                           fmaxf(__uint_as_float(v[4 * i + 2]) + b.z, 0.f), fmaxf(__uint_as_float(v[4 * i + 3]) + b.w, 0.f));
                    }
                    if (ok && p.fused_out) {
                        float4* dst = reinterpret_cast<float4*>(p.fused_out + r * D + c0);
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[i] = t[i];
                    }
                } else {
                    const float4* src = reinterpret_cast<const float4*>(a.x + r * a.xstride + c0);
#pragma unroll
                    for (int i = 0; i < 16; ++i) t[i] = ok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                uint8_t* chunk_hi = smem + OFF_FHI + half * 16384;
                uint8_t* chunk_lo = smem + OFF_FLO + half * 16384;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float hi[32], lo[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float x4[4] = {t[8 * q + i].x, t[8 * q + i].y, t[8 * q + i].z, t[8 * q + i].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            hi[4 * i + e] = __bfloat162float(__float2bfloat16_rn(x4[e]));
                            lo[4 * i + e] = x4[e] - hi[4 * i + e];
                        }
                    }
                    store_bf16_32(chunk_hi, row, 4 * q, hi);
                    store_bf16_32(chunk_lo, row, 4 * q, lo);
                }
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_F));
            }
            float yn = 0.f;
            if (do_maha) {       // ---- Y: |y|^2 partial, then (hi | lo) in place
                mbar_wait(BAR(B_ACC_Y), ph.next(B_ACC_Y), 67);
                tc_fence_after();
                load64(lane_base + TM_Y + c0);
                float y[64];
#pragma unroll
                for (int i = 0; i < 64; ++i) { y[i] = __uint_as_float(v[i]); yn = fmaf(y[i], y[i], yn); }
                store_split(lane_base + TM_Y + c0, y);
                ynorm[half * 128 + row] = yn;
                arrive_tmem(B_Y_RDY);
            }
            if (do_head) {
                const float *b0 = p.b0, *b1 = p.b1, *b2 = p.b2;
                // ---- H1 halves: bias + ReLU -> (hi | lo) in place
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(BAR(h == 0 ? B_ACC_H1A : B_ACC_H1B), ph.next(h == 0 ? B_ACC_H1A : B_ACC_H1B), 68);
                    tc_fence_after();
                    load64(lane_base + TM_H1 + 128 * h + c0);
                    float y[64];
#pragma unroll
                    for (int i = 0; i < 64; i += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(b0 + 128 * h + c0 + i));
                        y[i] = fmaxf(__uint_as_float(v[i]) + b.x, 0.f); y[i + 1] = fmaxf(__uint_as_float(v[i + 1]) + b.y, 0.f);
                        y[i + 2] = fmaxf(__uint_as_float(v[i + 2]) + b.z, 0.f); y[i + 3] = fmaxf(__uint_as_float(v[i + 3]) + b.w, 0.f);
                    }
                    store_split(lane_base + TM_H1 + 128 * h + c0, y);
                }
                arrive_tmem(B_H1_RDY);
                // ---- H2
                mbar_wait(BAR(B_ACC_H2), ph.next(B_ACC_H2), 69);
                tc_fence_after();
                load64(lane_base + TM_H2 + c0);
                {
                    float y[64];
#pragma unroll
                    for (int i = 0; i < 64; i += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + c0 + i));
                        y[i] = fmaxf(__uint_as_float(v[i]) + b.x, 0.f); y[i + 1] = fmaxf(__uint_as_float(v[i + 1]) + b.y, 0.f);
                        y[i + 2] = fmaxf(__uint_as_float(v[i + 2]) + b.z, 0.f); y[i + 3] = fmaxf(__uint_as_float(v[i + 3]) + b.w, 0.f);
                    }
                    store_split(lane_base + TM_H2 + c0, y);
                }
                arrive_tmem(B_H2_RDY);
                // ---- logits (half 0 threads) and Mahalanobis distances (half 1 threads)
                mbar_wait(BAR(B_ACC_OUT), ph.next(B_ACC_OUT), 70);
                tc_fence_after();
                if (half == 0) {
                    TMEM_LD32(lane_base + TM_Z, v);
                    tc_wait_ld();
                    if (ok) {
                        const int C = p.classes;
                        float z[CP];
                        float m = -INFINITY;
                        int idx = 0;
#pragma unroll
                        for (int c = 0; c < CP; ++c) {
                            z[c] = (c < C) ? __uint_as_float(v[c]) + __ldg(b2 + c) : -INFINITY;
                            if (z[c] > m) { m = z[c]; idx = c; }                     // first maximal index (torch max(1))
                        }
                        float se = 0.f;
#pragma unroll
                        for (int c = 0; c < CP; ++c) se += (c < C) ? expf(z[c] - m) : 0.f;
                        if (a.logits_out) {
                            float* dst = a.logits_out + r * C;
                            if ((C & 3) == 0) {
#pragma unroll
                                for (int c = 0; c < CP; c += 4) if (c < C) *reinterpret_cast<float4*>(dst + c) = make_float4(z[c], z[c + 1], z[c + 2], z[c + 3]);
                            } else {
#pragma unroll
                                for (int c = 0; c < CP; ++c) if (c < C) dst[c] = z[c];
                            }
                        }
                        if (a.pred_out) a.pred_out[r] = idx;
                        if (a.msp_out) a.msp_out[r] = -1.f / se;
                        if (a.energy_out) a.energy_out[r] = -(m + logf(se));
                    }
                }
            } else {
                mbar_wait(BAR(B_ACC_OUT), ph.next(B_ACC_OUT), 71);
                tc_fence_after();
            }
            if (do_maha && half == 1) {
                TMEM_LD32(lane_base + TM_G, v);
                tc_wait_ld();
                // both |y|^2 partials were stored before their warps arrived on B_Y_RDY, which the MMA warp observed
                // before committing B_ACC_OUT: visible here by release/acquire cumulativity
                if (ok && a.maha_out) {
                    const float* mun = reinterpret_cast<const float*>(p.maha_tc + IMG_WH + IMG_MU);
                    const float yy = ynorm[row] + ynorm[128 + row];
                    float best = INFINITY;
#pragma unroll
                    for (int c = 0; c < CP; ++c) best = fminf(best, yy - 2.f * __uint_as_float(v[c]) + __ldg(mun + c));
                    a.maha_out[r] = fmaxf(best, 0.f);
                }
            }
            // B_F of the next tile collects one arrival per epilogue warp AFTER it finished this tile, and every thread
            // has passed B_ACC_OUT (committed after all of this tile's MMAs): F tiles and TMEM may be overwritten.
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
    if (tid == 0) trace_end(TRACE_HEAD, trace_t0);
}

}  // namespace headtc

// ---- packing (called from cmhar_head_pack / cmhar_maha_pack): images of the folded fp32 matrices -----------
bool head_tc_eligible(int h1, int h2, int C) { return h1 == headtc::H1 && h2 == headtc::H2 && C >= 2 && C <= headtc::CP; }
size_t head_tc_bytes() { return headtc::HEAD_TC_BYTES; }
size_t maha_tc_bytes() { return headtc::MAHA_TC_BYTES; }

static int pack_stage(const float* src, int ld, int src_is_nk, int n0, int n_valid, int rows_img, int k0, uint8_t* dst, cudaStream_t st) {
    headtc::pack_split_stage_kernel<<<(rows_img * 8 + 255) / 256, 256, 0, st>>>(src, ld, src_is_nk, n0, n_valid, rows_img, k0, dst);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int pack_head_tc(const float* f32, const HeadLayout& hl, uint8_t* dst, cudaStream_t st) {
    using namespace headtc;
    for (int s = 0; s < 4; ++s) { const int rc = pack_stage(f32 + hl.w0(), hl.h1, 0, 128 * (s >> 1), hl.h1, 128, 64 * (s & 1), dst + (size_t)s * STAGE, st); if (rc) return rc; }
    for (int s = 0; s < 4; ++s) { const int rc = pack_stage(f32 + hl.w1(), hl.h2, 0, 0, hl.h2, 128, 64 * s, dst + IMG_W0 + (size_t)s * STAGE, st); if (rc) return rc; }
    for (int s = 0; s < 2; ++s) { const int rc = pack_stage(f32 + hl.w2(), hl.Cp(), 0, 0, hl.C, 32, 64 * s, dst + IMG_W0 + IMG_W1 + (size_t)s * 8192, st); if (rc) return rc; }
    return CMHAR_OK;
}

int pack_maha_tc(const float* f32, const MahaLayout& ml, uint8_t* dst, cudaStream_t st) {
    using namespace headtc;
    for (int s = 0; s < 2; ++s) { const int rc = pack_stage(f32 + ml.whiten(), D, 0, 0, D, 128, 64 * s, dst + (size_t)s * STAGE, st); if (rc) return rc; }
    for (int s = 0; s < 2; ++s) { const int rc = pack_stage(f32 + ml.mean_w(), D, 1, 0, ml.C, 32, 64 * s, dst + IMG_WH + (size_t)s * 8192, st); if (rc) return rc; }
    mu_norm_kernel<<<1, 32, 0, st>>>(f32 + ml.mean_w(), f32 + ml.valid(), ml.C, reinterpret_cast<float*>(dst + IMG_WH + IMG_MU));
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

// head_tc / maha_tc: device pointers to the tensor-core sections of the blobs (null = stage absent)
struct HeadPreLayer {          // optional late-fusion layer in front of the head (see headtc::Args)
    const uint8_t* w; const float* bias; const uint8_t* x1; const uint8_t* x2; int kc1, kc2; float* fused_out;
};

int launch_head_forward_tc(const FwdArgs& a, const uint8_t* head_tc, const float* head_f32, const HeadLayout& hl,
                           const uint8_t* maha_tc, cudaStream_t stream, const HeadPreLayer* pre) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(headtc::head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, headtc::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    headtc::Args p{};
    p.f = a;
    p.head_tc = head_tc;
    if (head_tc) { p.b0 = head_f32 + hl.b0(); p.b1 = head_f32 + hl.b1(); p.b2 = head_f32 + hl.b2(); p.classes = hl.C; }
    p.maha_tc = maha_tc;
    if (pre) { p.pre_w = pre->w; p.pre_bias = pre->bias; p.pre_x1 = pre->x1; p.pre_x2 = pre->x2; p.pre_kc1 = pre->kc1; p.pre_kc2 = pre->kc2; p.fused_out = pre->fused_out; }
    const long long tiles = (a.n + 127) / 128;
    const int grid = (int)(tiles < (long long)sm_count() ? tiles : (long long)sm_count());
    headtc::head_tc_kernel<<<grid, headtc::NT, headtc::SMEM_BYTES, stream>>>(p);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
