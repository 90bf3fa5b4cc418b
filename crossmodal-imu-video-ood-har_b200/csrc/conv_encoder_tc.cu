// conv_encoder_tc.cu -- the 1-D temporal conv / BatchNorm / ReLU IMU encoder on tensor cores (north-star item 1; bf16 path).
//
// SPEC-DEFINED -- NOT IN THE REFERENCE (see conv_encoder.cu / conv_encoder.py / oracle/fusion_spec.py):
//     Conv1d(6 -> 32, k5, s1, p2) BN ReLU -> Conv1d(32 -> 64, k5, s2, p2) BN ReLU -> Conv1d(64 -> 128, k5, s2, p2) BN ReLU -> time mean
// 8.2 MFLOP per 6 000-byte window (AI 1 370 flop/B): a compute-bound stack, so it belongs on the tensor cores -- the CUDA-core
// kernel of conv_encoder.cu runs it at 16 TFLOP/s of fp32 FMA (1.97 M windows/s).
//
// Every layer is an implicit GEMM over im2col tiles that live ONLY in shared memory; the layer's epilogue writes the NEXT layer's
// im2col tile directly (each activation row goes to the 2-3 (row, tap) places that read it), so no activation ever leaves the SM:
//   L1   A1 [256 positions x 64]   = [hi(x) taps (30) 0 0 | lo(x) taps (30) 0 0]    B = W1 image [32 x 64] = [W | W]   (split-precision input)
//        D1 [128 lanes x 32 cols] x 2 tiles                                         tcgen05.mma M128 N32 K16 x 4 per tile
//   L2   A2 [128 positions x 160]  tap-major (tap k = columns 32k .. 32k+31)        B = W2 image [64 x 160]
//        D2 [128 x 64]                                                              M128 N64 K16 x 10
//   L3   TRANSPOSED: A = W3 image [128 channels x 320] (tap k = 64-wide chunk k)    B = A3 [64 positions x 320]
//        D3 [128 lanes = channels x 64 cols = positions]                            M128 N64 K16 x 20
//        -> BN/ReLU per lane and the time mean is a per-thread sum over the columns: no cross-lane reduction.
// BatchNorm (eval) scale is folded into the bf16 weight images, the shift into fp32 biases (pack time, fp64).
// One persistent CTA per SM, two window slots in flight: 4 builder/epilogue warps per slot (thread = TMEM lane = tile row) and
// one MMA-issuing warp that alternates between the slots, so one slot's epilogue runs under the other slot's MMAs.
// The three im2col tiles of a slot overlay each other (each is dead when the next is written); weights stay resident (108 KiB).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace convtc {

using namespace tc;

constexpr int C0 = 6, C1 = 32, C2 = 64, C3 = 128, KW = 5, MAX_L = 256;
// weight images (bytes): W1 [32 rows x 64] 4 KiB | W2 3 chunks of [64 rows x 64] 8 KiB | W3 5 chunks of [128 rows x 64] 16 KiB | biases
constexpr int W1_BYTES = 4096, W2_BYTES = 3 * 8192, W3_BYTES = 5 * 16384;
constexpr int WT_BYTES = W1_BYTES + W2_BYTES + W3_BYTES;                  // 110 592
constexpr int BIAS_FLOATS = C1 + C2 + C3;                                  // 224
constexpr int IMG_BYTES = WT_BYTES + BIAS_FLOATS * 4;
// shared memory map
constexpr int S_W = 0;                                   // resident weight images + biases
constexpr int S_SLOT = (IMG_BYTES + 1023) / 1024 * 1024; // 2 slots
constexpr int SLOT_TILE = 3 * 16384;                     // A1 (2 x 16 KiB) / A2 (3 x 16 KiB) / A3 (5 x 8 KiB) overlaid
constexpr int SLOT_X = C0 * (MAX_L + 4) * 4;             // staged window with a 2-sample zero halo
constexpr int SLOT_BYTES = SLOT_TILE + (SLOT_X + 1023) / 1024 * 1024;
constexpr int S_BAR = S_SLOT + 2 * SLOT_BYTES;
constexpr int SMEM_BYTES = S_BAR + 64;
static_assert(SMEM_BYTES <= 232448, "conv tensor-core kernel exceeds the shared memory of a CTA");
enum { B_W = 0, B_RDY = 1 /*[2]*/, B_ACC = 3 /*[2]*/, B_X = 5 /*[2]: the slot's next window has landed in shared memory*/ };
constexpr int NT = 288;                                  // 8 builder / epilogue warps + 1 MMA warp
// TMEM columns per slot: D1 tile 0 [0,32) tile 1 [32,64) | D2 [64,128) | D3 [128,192)
constexpr uint32_t T_SLOT = 192, T_D1 = 0, T_D2 = 64, T_D3 = 128;

__host__ __device__ constexpr int out_len(int L, int stride) { return (L + 4 - KW) / stride + 1; }

__global__ void __launch_bounds__(NT, 1) conv_encoder_tc_kernel(const uint8_t* __restrict__ img, const float* __restrict__ x, long long n, int L,
                                                                long long xstride, float* __restrict__ feat) {
    extern __shared__ __align__(1024) uint8_t smem_conv[];
    uint8_t* const smem = smem_conv;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L2n = out_len(L, 2), L3n = out_len(L2n, 2);
    const uint32_t sbase = smem_u32(smem);
    auto BAR = [&](int i) { return sbase + S_BAR + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S_BAR + 48);
    // windows of this CTA: pair p = blockIdx.x + it * gridDim.x holds windows 2p (slot 0) and 2p + 1 (slot 1)
    const long long pairs = (n + 1) / 2;

    if (tid == 0) {
        mbar_init(BAR(B_W), 1);
        for (int g = 0; g < 2; ++g) { mbar_init(BAR(B_RDY + g), 4); mbar_init(BAR(B_ACC + g), 1); mbar_init(BAR(B_X + g), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(BAR(B_W), IMG_BYTES);
        bulk_g2s(sbase + S_W, img, IMG_BYTES, BAR(B_W));
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // ================================================================= MMA issuer (one elected lane)
        const bool leader = elect_one();
        constexpr uint32_t ID32 = idesc_bf16(128, 32), ID64 = idesc_bf16(128, 64);
        const uint64_t dW1 = sw128_desc(sbase + S_W), dW2 = sw128_desc(sbase + S_W + W1_BYTES), dW3 = sw128_desc(sbase + S_W + W1_BYTES + W2_BYTES);
        mbar_wait(BAR(B_W), 0, 1);
        uint32_t rdy_bits = 0;
        for (long long p = blockIdx.x; p < pairs; p += gridDim.x) {
            const bool has1 = (2 * p + 1 < n);
            for (int layer = 0; layer < 3; ++layer) {
                for (int g = 0; g < 2; ++g) {
                    if (g == 1 && !has1) continue;
                    mbar_wait(BAR(B_RDY + g), (rdy_bits >> g) & 1u, 2 + layer);
                    rdy_bits ^= (1u << g);
                    tc_fence_after();
                    const uint32_t T = tmem + T_SLOT * (uint32_t)g;
                    const uint64_t dA = sw128_desc(sbase + S_SLOT + g * SLOT_BYTES);
                    if (layer == 0) {
                        for (int j = 0; j < 2; ++j)
                            for (int k = 0; k < 4; ++k)
                                if (leader) umma(T + T_D1 + 32 * j, dA + (uint64_t)j * 1024 + (uint64_t)(2 * k), dW1 + (uint64_t)(2 * k), ID32, k > 0 ? 1u : 0u);
                    } else if (layer == 1) {
                        for (int s = 0; s < 10; ++s)
                            if (leader) umma(T + T_D2, dA + (uint64_t)(s >> 2) * 1024 + (uint64_t)((s & 3) * 2), dW2 + (uint64_t)(s >> 2) * 512 + (uint64_t)((s & 3) * 2),
                                             ID64, s > 0 ? 1u : 0u);
                    } else {
                        for (int s = 0; s < 20; ++s)      // transposed: weights are the A operand, the im2col tile the B operand
                            if (leader) umma(T + T_D3, dW3 + (uint64_t)(s >> 2) * 1024 + (uint64_t)((s & 3) * 2), dA + (uint64_t)(s >> 2) * 512 + (uint64_t)((s & 3) * 2),
                                             ID64, s > 0 ? 1u : 0u);
                    }
                    if (leader) tc_commit(BAR(B_ACC + g));
                }
            }
        }
    } else {
        // ================================================================= builder / epilogue: 4 warps per slot, thread = tile row
        const int g = warp >> 2, r = (warp & 3) * 32 + lane;            // slot, row of the 128-row tiles (== TMEM lane)
        const int gt = tid & 127;                                       // thread index inside the slot's group
        uint8_t* const tile = smem + S_SLOT + g * SLOT_BYTES;
        float* const xs = reinterpret_cast<float*>(tile + SLOT_TILE);   // [6][L]: the staged window
        const float* bias = reinterpret_cast<const float*>(smem + S_W + WT_BYTES);
        const uint32_t T = tmem + ((uint32_t)((warp & 3) * 32) << 16) + T_SLOT * (uint32_t)g;
        const bool bulk_x = ((((uintptr_t)x) | ((uintptr_t)xstride * 4) | (uintptr_t)(C0 * L * 4)) & 15) == 0;
        uint32_t acc_par = 0, x_par = 0;
        uint32_t v[32];
        float f[32];
        auto group_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };
        auto publish = [&] {
            tc_wait_st();
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_RDY + g));
        };
        auto wait_acc = [&](int site) { mbar_wait(BAR(B_ACC + g), acc_par, site); acc_par ^= 1; tc_fence_after(); };
        // one activation row (CH channels, already ReLU'd or zero) of position p -> the (row, tap) places of the next layer's
        // im2col tile that read it: rows t = (p + 2 - k) / 2 for the taps k of p's parity, 0 <= t < rows_next
        mbar_wait(BAR(B_W), 0, 1);
        for (long long p = blockIdx.x; p < pairs; p += gridDim.x) {
            const long long w = 2 * p + g;
            if (w >= n) break;
            // ---- the window (6 x L fp32, contiguous): bulk-copied into shared memory one window ahead (issued right after the
            // previous window's im2col build, so the copy runs under that window's three layers); plain loads when misaligned
            const float* src = x + w * xstride;
            if (bulk_x) {
                if (p == (long long)blockIdx.x && gt == 0) {          // first window of the slot: nothing was prefetched
                    mbar_expect_tx(BAR(B_X + g), (uint32_t)(C0 * L * 4));
                    bulk_g2s(smem_u32(xs), src, (uint32_t)(C0 * L * 4), BAR(B_X + g));
                }
                mbar_wait(BAR(B_X + g), x_par, 9);
                x_par ^= 1;
            } else {
                for (int e = gt; e < C0 * L; e += 128) xs[e] = __ldg(src + e);
                group_sync();
            }
            // ---- A1: two tiles of 128 positions; row = [hi(x) 30 taps, 0, 0 | lo(x) 30 taps, 0, 0]
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                const int pos = 128 * j + r;
                float hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) { hi[c] = 0.f; lo[c] = 0.f; }
                if (pos < L) {
#pragma unroll
                    for (int k = 0; k < KW; ++k) {
                        const int q = pos + k - 2;
                        const bool in = (q >= 0 && q < L);
#pragma unroll
                        for (int ci = 0; ci < C0; ++ci) {
                            const float xv = in ? xs[ci * L + q] : 0.f;
                            const float h = __bfloat162float(__float2bfloat16_rn(xv));
                            hi[k * C0 + ci] = h; lo[k * C0 + ci] = xv - h;
                        }
                    }
                }
                store_bf16_32(tile + j * 16384, r, 0, hi);
                store_bf16_32(tile + j * 16384, r, 4, lo);
            }
            group_sync();                 // every thread of the slot has read its samples: the staging buffer is free
            if (bulk_x && gt == 0) {
                const long long wn = 2 * (p + gridDim.x) + g;
                if (wn < n) {
                    fence_async_smem();   // generic-proxy reads of xs are ordered before the async-proxy write that reuses it
                    mbar_expect_tx(BAR(B_X + g), (uint32_t)(C0 * L * 4));
                    bulk_g2s(smem_u32(xs), x + wn * xstride, (uint32_t)(C0 * L * 4), BAR(B_X + g));
                }
            }
            publish();
            // ---- L1 epilogue: relu(D1 + b1) (zero beyond the window) -> A2, tap-major 32-channel groups
            wait_acc(11);                 // (all four warps published before this commit: nobody is still writing A1, which A2 overlays)
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                const int pos = 128 * j + r;
                TMEM_LD32(T + T_D1 + 32 * j, v);
                tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(bias + c);
                    f[c] = fmaxf(__uint_as_float(v[c]) + b.x, 0.f); f[c + 1] = fmaxf(__uint_as_float(v[c + 1]) + b.y, 0.f);
                    f[c + 2] = fmaxf(__uint_as_float(v[c + 2]) + b.z, 0.f); f[c + 3] = fmaxf(__uint_as_float(v[c + 3]) + b.w, 0.f);
                }
                if (pos >= L) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] = 0.f;
                }
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                    const int tt = pos + 2 - k;
                    if (tt >= 0 && !(tt & 1) && (tt >> 1) < L2n) store_bf16_32(tile + (k >> 1) * 16384, tt >> 1, (k & 1) * 4, f);
                }
                if (pos < 2 || pos == 255) {
                    // the conv's zero padding: positions -2 and -1 are read by row 0, taps 0 and 1; position 256 (windows of 255 /
                    // 256 samples) by row 127, tap 4 -- nobody else writes those places
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] = 0.f;
                    if (pos < 2) store_bf16_32(tile, 0, pos * 4, f);
                    else store_bf16_32(tile + 2 * 16384, 127, 0, f);
                }
            }
            publish();
            // ---- L2 epilogue: relu(D2 + b2) (zero beyond L2n) -> A3: tap k = chunk k of [64 positions x 64 channels]
            wait_acc(12);
            {
                float f2[64];
#pragma unroll
                for (int cc = 0; cc < 64; cc += 32) {
                    TMEM_LD32(T + T_D2 + cc, v);
                    tc_wait_ld();
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(bias + C1 + cc + c);
                        f2[cc + c] = fmaxf(__uint_as_float(v[c]) + b.x, 0.f); f2[cc + c + 1] = fmaxf(__uint_as_float(v[c + 1]) + b.y, 0.f);
                        f2[cc + c + 2] = fmaxf(__uint_as_float(v[c + 2]) + b.z, 0.f); f2[cc + c + 3] = fmaxf(__uint_as_float(v[c + 3]) + b.w, 0.f);
                    }
                }
                if (r >= L2n) {
#pragma unroll
                    for (int c = 0; c < 64; ++c) f2[c] = 0.f;
                }
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                    const int tt = r + 2 - k;
                    if (tt >= 0 && !(tt & 1) && (tt >> 1) < 64) {
                        store_bf16_32(tile + k * 8192, tt >> 1, 0, f2);
                        store_bf16_32(tile + k * 8192, tt >> 1, 4, f2 + 32);
                    }
                }
                if (r < 2 || r == 127) {  // zero padding of layer 3: positions -2, -1 (row 0, taps 0 and 1) and 128 (row 63, tap 4)
#pragma unroll
                    for (int c = 0; c < 64; ++c) f2[c] = 0.f;
                    uint8_t* dst = (r < 2) ? tile + r * 8192 : tile + 4 * 8192;
                    const int row = (r < 2) ? 0 : 63;
                    store_bf16_32(dst, row, 0, f2);
                    store_bf16_32(dst, row, 4, f2 + 32);
                }
            }
            publish();
            // ---- L3 epilogue (transposed accumulator: lane = channel, column = position): time mean of relu(D3 + b3)
            wait_acc(13);
            {
                const float b3 = bias[C1 + C2 + r];
                float sum = 0.f;
#pragma unroll
                for (int cc = 0; cc < 64; cc += 32) {
                    TMEM_LD32(T + T_D3 + cc, v);
                    tc_wait_ld();
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (cc + c < L3n) sum += fmaxf(__uint_as_float(v[c]) + b3, 0.f);
                }
                feat[w * C3 + r] = sum / (float)L3n;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ---- weight images: BatchNorm scale folded (fp64), bf16, SWIZZLE_128B K-major
__device__ __forceinline__ double bn_scale(const float* bn_w, const float* bn_var, int co) {
    return bn_w ? (double)bn_w[co] / sqrt((double)bn_var[co] + (double)BN_EPS) : 1.0;
}
__global__ void pack_conv_tc_kernel(const cmhar_conv_encoder_params p, uint8_t* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    __nv_bfloat16* const w1 = reinterpret_cast<__nv_bfloat16*>(img);
    __nv_bfloat16* const w2 = reinterpret_cast<__nv_bfloat16*>(img + W1_BYTES);
    __nv_bfloat16* const w3 = reinterpret_cast<__nv_bfloat16*>(img + W1_BYTES + W2_BYTES);
    float* const bias = reinterpret_cast<float*>(img + WT_BYTES);
    auto put = [](__nv_bfloat16* chunk, int row, int col, double val) {
        chunk[(sw128_off(row, col >> 3) >> 1) + (col & 7)] = __float2bfloat16_rn((float)val);
    };
    if (i < C1 * 64) {                                   // W1 [32 x 64] = [W | W], column = tap * 6 + ci
        const int co = i >> 6, c = i & 63, cc = c & 31;
        double val = 0.0;
        if (cc < KW * C0) val = (double)p.layer[0].weight[((size_t)co * C0 + cc % C0) * KW + cc / C0] * bn_scale(p.layer[0].bn_weight, p.layer[0].bn_var, co);
        put(w1, co, c, val);
    }
    if (i < 3 * C2 * 64) {                               // W2: 3 chunks of [64 x 64], column kk = 64 q + c = tap * 32 + ci
        const int q = i / (C2 * 64), co = (i >> 6) % C2, c = i & 63, kk = 64 * q + c;
        double val = 0.0;
        if (kk < KW * C1) val = (double)p.layer[1].weight[((size_t)co * C1 + kk % C1) * KW + kk / C1] * bn_scale(p.layer[1].bn_weight, p.layer[1].bn_var, co);
        put(w2 + q * 4096, co, c, val);
    }
    if (i < KW * C3 * 64) {                              // W3: chunk = tap, [128 x 64] = [co][ci]
        const int k = i / (C3 * 64), co = (i >> 6) % C3, ci = i & 63;
        put(w3 + k * 8192, co, ci, (double)p.layer[2].weight[((size_t)co * C2 + ci) * KW + k] * bn_scale(p.layer[2].bn_weight, p.layer[2].bn_var, co));
    }
    if (i < BIAS_FLOATS) {
        const int l = i < C1 ? 0 : (i < C1 + C2 ? 1 : 2), co = i - (l == 0 ? 0 : (l == 1 ? C1 : C1 + C2));
        const cmhar_conv_layer_params& q = p.layer[l];
        const double g = bn_scale(q.bn_weight, q.bn_var, co), b = q.bias ? (double)q.bias[co] : 0.0;
        bias[i] = (float)(q.bn_weight ? (b - (double)q.bn_mean[co]) * g + (double)q.bn_bias[co] : b);
    }
}

}  // namespace convtc

size_t conv_tc_image_bytes() { return (size_t)convtc::IMG_BYTES; }

int pack_conv_tc(const cmhar_conv_encoder_params* p, void* img, cudaStream_t st) {
    const int total = convtc::KW * convtc::C3 * 64;
    convtc::pack_conv_tc_kernel<<<(total + 255) / 256, 256, 0, st>>>(*p, reinterpret_cast<uint8_t*>(img));
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int launch_conv_encoder_tc(const void* img, const float* x, long long n, int L, long long xstride, float* feat, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(convtc::conv_encoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, convtc::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    const long long pairs = (n + 1) / 2;
    const int grid = (int)(pairs < (long long)sm_count() ? pairs : (long long)sm_count());
    convtc::conv_encoder_tc_kernel<<<grid, convtc::NT, convtc::SMEM_BYTES, st>>>(reinterpret_cast<const uint8_t*>(img), x, n, L, xstride, feat);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
