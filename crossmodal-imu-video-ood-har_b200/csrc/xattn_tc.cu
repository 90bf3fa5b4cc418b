// xattn_tc.cu -- the cross-attention fusion block as ONE tcgen05 launch (spec row A6, BASELINE configs[2]; bf16 path).
//
// SPEC-DEFINED -- NOT IN THE REFERENCE (fusion.py / oracle/fusion_spec.py): the S IMU tokens of a window attend over its T frame
// tokens (8 heads of 16), y = LayerNorm(tokens + out_proj(attn)), fused = mean over the tokens.  With the video projection folded
// into the kv projection (fusion.py, _packed_kv_folded) the block is
//     q = tokens Wq^T + bq                [S x 128]           [k | v] = pooled_frames Wkv'^T (+ b)      [T x 256], K = F (512)
//     o_h = softmax(q_h k_h^T / 4) v_h    per head            y = LN(tokens + o Wo^T + bo')             fused = mean_s y
// (the k bias cancels in the softmax; the v bias goes through the attention average unchanged, so it is folded into bo' at pack time).
// Tile = 8 windows = 128 token rows (row = 16 window + token) against the 128 frame rows of the same windows (T = 16), which is
// exactly the shape of one encoder layer's attention (imu_encoder_bf16.cu) with K / V coming from another source, so the same
// machinery is used: Q, P and O live as bf16 pairs in tensor memory over their own accumulators (TMEM A operands); K and V^T are
// SWIZZLE_128B tiles in shared memory; scores and P V are one 128x16x16 MMA per (head, window) restricted to the window's 16 lanes
// by the disable-output-lane mask; the residual (tokens + bo') is preloaded into the out-projection accumulator; V^T comes out of
// its GEMM transposed (weights as the A operand).  The frame rows arrive as the bf16 operand image the pooling kernel wrote
// (cmhar_video_pool_frames_img): plain cp.async.bulk copies, one ring item per 64-wide k chunk = [frames | Wk | Wv] (48 KiB).
// Roles: warps 0-7 epilogue (thread = (row, column half)), warp 8 MMA issuer, warp 9 producer.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace xattn {

using namespace tc;

constexpr int NSTAGE = 3, STAGE = 3 * CHUNK;                 // [F_c | Wk_c | Wv_c]; Wq / Wo (2 chunks each) use the first 32 KiB of a stage
constexpr int OFF_TK = 0;                                     // tokens (bf16, A of the Q GEMM), then K (B of the score MMAs)
constexpr int OFF_VT = 32768;
constexpr int OFF_RING = 65536;
constexpr int OFF_PAR = OFF_RING + NSTAGE * STAGE;            // bq' | bo' | gamma | beta (fp32)
constexpr int OFF_BAR = OFF_PAR + 4 * 128 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "xattn kernel exceeds the shared memory of a CTA");
enum { B_FULL = 0, B_EMPTY = NSTAGE, B_ACC = 2 * NSTAGE /*[4] A,B,C,R*/, B_TOK = B_ACC + 4, B_QK, B_V, B_P, B_O, B_PAR, B_COUNT };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier area too small");
constexpr uint32_t TM_A = 0, TM_B = 128, TM_C = 256, TM_R = 384;
constexpr int NT = 320;
constexpr uint32_t XATTN_MAGIC = 0x434d4836u;

struct Args {
    const uint8_t* blob;        // header | chunks [Wq0 Wq1 Wo0 Wo1 (Wk_c Wv_c) x kc] | params
    const float* tokens;        // (n, S, 128) fp32
    const uint8_t* frame_img;   // operand image of n * 16 rows x F
    long long n;
    int S, kc;
    float eps;
    float* fused;               // (n, 128)
};

struct Phase {
    uint32_t bits = 0;
    __device__ __forceinline__ uint32_t next(int i) { const uint32_t p = (bits >> i) & 1u; bits ^= (1u << i); return p; }
};

__global__ void __maxnreg__(168) xattn_tc_kernel(const Args a) {
    constexpr int CW = 64;
    extern __shared__ __align__(1024) uint8_t smem_x[];
    uint8_t* const smem = smem_x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem);
    auto BAR = [&](int i) { return sbase + OFF_BAR + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    const uint8_t* wch = a.blob + 1024;
    const float* gpar = reinterpret_cast<const float*>(wch + (size_t)(4 + 2 * a.kc) * CHUNK);
    const long long tiles = (a.n + 7) / 8;
    const int S = a.S, kc = a.kc;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
        for (int i = 0; i < 4; ++i) mbar_init(BAR(B_ACC + i), 1);
        for (int i = B_TOK; i <= B_O; ++i) mbar_init(BAR(i), 8);
        mbar_init(BAR(B_PAR), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 9) {
        // ================================================================= producer: per tile [Wq] [F_c | Wk_c | Wv_c] x kc [Wo]
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last(), once = l2_policy_evict_first();
            mbar_expect_tx(BAR(B_PAR), 4 * 128 * 4);
            bulk_g2s(sbase + OFF_PAR, gpar, 4 * 128 * 4, BAR(B_PAR));
            uint32_t st = 0, par = 1;
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                for (int it = 0; it < kc + 2; ++it) {
                    mbar_wait(BAR(B_EMPTY + st), par, 1);
                    const uint32_t dst = sbase + OFF_RING + st * STAGE;
                    if (it == 0 || it == kc + 1) {
                        mbar_expect_tx(BAR(B_FULL + st), 2 * CHUNK);
                        bulk_g2s_hint(dst, wch + (size_t)(it == 0 ? 0 : 2) * CHUNK, 2 * CHUNK, BAR(B_FULL + st), keep);
                    } else {
                        const int c = it - 1;
                        mbar_expect_tx(BAR(B_FULL + st), 3 * CHUNK);
                        bulk_g2s_hint(dst, a.frame_img + ((size_t)tile * kc + c) * CHUNK, CHUNK, BAR(B_FULL + st), once);
                        bulk_g2s_hint(dst + CHUNK, wch + (size_t)(4 + 2 * c) * CHUNK, 2 * CHUNK, BAR(B_FULL + st), keep);
                    }
                    if (++st == NSTAGE) { st = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 8) {
        // ================================================================= MMA issuer (warp converged, one elected lane)
        const bool leader = elect_one();
        Phase ph;
        uint32_t st = 0, par = 0;
        constexpr uint32_t ID128 = idesc_bf16(128, 128), ID16 = idesc_bf16(128, 16);
        const uint64_t CH = CHUNK >> 4;
        const uint64_t dTK = sw128_desc(sbase + OFF_TK), dVT = sw128_desc(sbase + OFF_VT);
        auto ring = [&]() -> uint64_t {
            mbar_wait(BAR(B_FULL + st), par, 2);
            tc_fence_after();
            return sw128_desc(sbase + OFF_RING + st * STAGE);
        };
        auto ring_done = [&]() { if (leader) tc_commit(BAR(B_EMPTY + st)); if (++st == NSTAGE) { st = 0; par ^= 1; } };
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            // ---- Q = tokens Wq^T
            mbar_wait(BAR(B_TOK), ph.next(B_TOK), 3);
            tc_fence_after();
            {
                const uint64_t dW = ring();
                for (int c = 0; c < 2; ++c)
                    for (int k = 0; k < 4; ++k)
                        if (leader) umma(tmem + TM_A, dTK + (uint64_t)c * CH + (uint64_t)(2 * k), dW + (uint64_t)c * CH + (uint64_t)(2 * k), ID128, (c > 0 || k > 0) ? 1u : 0u);
                ring_done();
                if (leader) tc_commit(BAR(B_ACC + 0));
            }
            // ---- K = frames Wk'^T (frames as A) and V^T = Wv' frames^T (weights as A), one k chunk per ring item
            for (int c = 0; c < kc; ++c) {
                const uint64_t dF = ring(), dWk = dF + CH, dWv = dF + 2 * CH;
                for (int k = 0; k < 4; ++k)
                    if (leader) umma(tmem + TM_B, dF + (uint64_t)(2 * k), dWk + (uint64_t)(2 * k), ID128, (c > 0 || k > 0) ? 1u : 0u);
                for (int k = 0; k < 4; ++k)
                    if (leader) umma(tmem + TM_C, dWv + (uint64_t)(2 * k), dF + (uint64_t)(2 * k), ID128, (c > 0 || k > 0) ? 1u : 0u);
                ring_done();
            }
            if (leader) { tc_commit(BAR(B_ACC + 1)); tc_commit(BAR(B_ACC + 2)); }
            // ---- compact scores S[r][16h + t] = q_h(r) . k_h(frame t of r's window)
            mbar_wait(BAR(B_QK), ph.next(B_QK), 5);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const uint64_t off = (uint64_t)(h >> 2) * CH + (uint64_t)((h & 3) * 2);
                    if (leader) umma_ts_rows16(tmem + TM_B + 16 * h, tmem + TM_A + tm_bf16_col<CW>(16 * h), dTK + off + (uint64_t)(j * 128), ID16, j);
                }
            }
            if (leader) tc_commit(BAR(B_ACC + 1));
            // ---- O = P V
            mbar_wait(BAR(B_V), ph.next(B_V), 7);
            mbar_wait(BAR(B_P), ph.next(B_P), 6);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint64_t koff = (uint64_t)(j >> 2) * CH + (uint64_t)((j & 3) * 2);
#pragma unroll
                for (int h = 0; h < H; ++h)
                    if (leader) umma_ts_rows16(tmem + TM_C + 16 * h, tmem + TM_B + tm_bf16_col<CW>(16 * h), dVT + koff + (uint64_t)(h * 128), ID16, j);
            }
            if (leader) tc_commit(BAR(B_ACC + 2));
            // ---- out-projection on top of the residual: R(tokens + bo') += O Wo^T
            mbar_wait(BAR(B_O), ph.next(B_O), 8);
            tc_fence_after();
            {
                const uint64_t dW = ring();
                for (int c = 0; c < 2; ++c)
                    for (int k = 0; k < 4; ++k)
                        if (leader) umma_ts(tmem + TM_R, tmem + TM_C + tm_bf16_col<CW>(64 * c + 16 * k), dW + (uint64_t)c * CH + (uint64_t)(2 * k), ID128, 1u);
                ring_done();
                if (leader) tc_commit(BAR(B_ACC + 3));
            }
        }
    } else {
        // ================================================================= epilogue (warps 0-7): thread = (row, 64-column half)
        const int wq = warp >> 2, row = (warp & 3) * 32 + lane, win = row >> 4, tok = row & 15, c0 = wq * CW;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float* PAR = reinterpret_cast<const float*>(smem + OFF_PAR);      // bq' | bo' | gamma | beta
        auto chunk_of = [](int c) { return (c >> 6) * CHUNK; };
        auto piece_of = [](int c) { return (c & 63) >> 3; };
        auto ld4 = [](const float* p) { return *reinterpret_cast<const float4*>(p); };
        Phase ph;
        uint32_t v[32];
        float f[32];
        auto publish = [&](int bar) { tc_wait_st(); fence_async_smem(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(BAR(bar)); };
        auto publish_tmem = [&](int bar) { tc_wait_st(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(BAR(bar)); };
        auto store_tmem_bf16 = [&](uint32_t taddr, const float* y32) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(y32[2 * i], y32[2 * i + 1]);
            TMEM_ST16(taddr, pk);
        };
        mbar_wait(BAR(B_PAR), 0, 20);
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const long long w = tile * 8 + win;
            const bool live = (w < a.n) && (tok < S);
            // ---- tokens -> bf16 A tile, tokens + bo' -> R
            {
                const float* src = a.tokens + ((size_t)w * S + tok) * D + c0;
#pragma unroll
                for (int cc = 0; cc < CW; cc += 32) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 t = live ? __ldg(reinterpret_cast<const float4*>(src + cc + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        f[i] = t.x; f[i + 1] = t.y; f[i + 2] = t.z; f[i + 3] = t.w;
                    }
                    store_bf16_32(smem + OFF_TK + chunk_of(c0 + cc), row, piece_of(c0 + cc), f);
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b = ld4(PAR + 128 + c0 + cc + i);
                        v[i] = __float_as_uint(f[i] + b.x); v[i + 1] = __float_as_uint(f[i + 1] + b.y);
                        v[i + 2] = __float_as_uint(f[i + 2] + b.z); v[i + 3] = __float_as_uint(f[i + 3] + b.w);
                    }
                    TMEM_ST32(lane_base + TM_R + c0 + cc, v);
                }
            }
            publish(B_TOK);
            // ---- Q (+ pre-scaled bias) -> bf16 pairs over its accumulator
            mbar_wait(BAR(B_ACC + 0), ph.next(B_ACC + 0), 11);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < CW; cc += 32) {
                TMEM_LD32(lane_base + TM_A + c0 + cc, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b = ld4(PAR + c0 + cc + i);
                    f[i] = __uint_as_float(v[i]) + b.x; f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
                    f[i + 2] = __uint_as_float(v[i + 2]) + b.z; f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
                }
                store_tmem_bf16(lane_base + TM_A + c0 + (cc >> 1), f);
            }
            // ---- K -> shared memory (over the token tile: the Q GEMM is complete), V^T -> shared memory
#pragma unroll 1
            for (int m = 1; m < 3; ++m) {
                mbar_wait(BAR(B_ACC + m), ph.next(B_ACC + m), 12);
                tc_fence_after();
                uint8_t* dst = smem + (m == 1 ? OFF_TK : OFF_VT);
#pragma unroll
                for (int cc = 0; cc < CW; cc += 32) {
                    TMEM_LD32(lane_base + 128 * m + c0 + cc, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                    store_bf16_32(dst + chunk_of(c0 + cc), row, piece_of(c0 + cc), f);
                }
                publish(m == 1 ? B_QK : B_V);
            }
            // ---- softmax over the 16 frame keys of the row's window (scores pre-multiplied by log2(e)/4 through Wq')
            mbar_wait(BAR(B_ACC + 1), ph.next(B_ACC + 1), 13);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < CW; cc += 32) {
                TMEM_LD32(lane_base + TM_B + c0 + cc, v);
                tc_wait_ld();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 16; ++i) { f[hh * 16 + i] = __uint_as_float(v[hh * 16 + i]); mx = fmaxf(mx, f[hh * 16 + i]); }
                    float den = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) { f[hh * 16 + i] = ex2_approx(f[hh * 16 + i] - mx); den += f[hh * 16 + i]; }
                    const float inv = rcp_approx(den);
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[hh * 16 + i] *= inv;
                }
                store_tmem_bf16(lane_base + TM_B + c0 + (cc >> 1), f);
            }
            publish_tmem(B_P);
            // ---- O -> bf16 pairs over its accumulator
            mbar_wait(BAR(B_ACC + 2), ph.next(B_ACC + 2), 15);
            tc_fence_after();
            {
                uint32_t vv[CW];
#pragma unroll
                for (int cc = 0; cc < CW; cc += 32) TMEM_LD32(lane_base + TM_C + c0 + cc, (vv + cc));
                tc_wait_ld();
#pragma unroll
                for (int cc = 0; cc < CW; cc += 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(vv[cc + i]);
                    store_tmem_bf16(lane_base + TM_C + c0 + (cc >> 1), f);
                }
            }
            publish_tmem(B_O);
            // ---- y = LN(tokens + attn Wo^T + bo'), fused = mean over the window's S tokens
            mbar_wait(BAR(B_ACC + 3), ph.next(B_ACC + 3), 16);
            tc_fence_after();
            {
                float x[CW];
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int cc = 0; cc < CW; cc += 32) {
                    TMEM_LD32(lane_base + TM_R + c0 + cc, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) { x[cc + i] = __uint_as_float(v[i]); s1 += x[cc + i]; s2 = fmaf(x[cc + i], x[cc + i], s2); }
                }
                // the row's other 64 columns live in the thread of warp (warp ^ 4), same lane: exchange through shared memory
                float2* st = reinterpret_cast<float2*>(smem + OFF_VT);           // V^T is dead (the P V MMAs are complete)
                st[wq * 128 + row] = make_float2(s1, s2);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
                const float2 o0 = st[row], o1 = st[128 + row];
                const float mean = (o0.x + o1.x) * (1.f / D);
                const float rstd = rsqrtf(fmaxf((o0.y + o1.y) * (1.f / D) - mean * mean, 0.f) + a.eps);
                const float scale = live ? 1.f / (float)S : 0.f;                  // padding rows contribute nothing to the token mean
#pragma unroll
                for (int i = 0; i < CW; i += 4) {
                    const float4 g = ld4(PAR + 256 + c0 + i), b = ld4(PAR + 384 + c0 + i);
                    x[i] = ((x[i] - mean) * rstd * g.x + b.x) * scale; x[i + 1] = ((x[i + 1] - mean) * rstd * g.y + b.y) * scale;
                    x[i + 2] = ((x[i + 2] - mean) * rstd * g.z + b.z) * scale; x[i + 3] = ((x[i + 3] - mean) * rstd * g.w + b.w) * scale;
                }
                // sum over the 16 lanes of the window (tokens), butterfly inside the half warp
#pragma unroll
                for (int i = 0; i < CW; ++i) {
                    float t = x[i];
                    t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 4);
                    t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 1);
                    x[i] = t;
                }
                if (tok == 0 && w < a.n) {
                    float4* dst = reinterpret_cast<float4*>(a.fused + (size_t)w * D + c0);
#pragma unroll
                    for (int i = 0; i < CW / 4; ++i) dst[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
                }
                // the statistics buffer (= V^T tile) must not be overwritten by the next tile's V^T drain before both threads read it
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
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

__global__ void pack_xattn_params_kernel(const float* bq, const float* bo, const float* g, const float* b, float qscale, float* dst) {
    const int i = threadIdx.x;
    if (i < 128) { dst[i] = bq[i] * qscale; dst[128 + i] = bo[i]; dst[256 + i] = g[i]; dst[384 + i] = b[i]; }
}

}  // namespace xattn
}  // namespace cmhar

using namespace cmhar;

extern "C" {

size_t cmhar_xattn_blob_bytes(int32_t frame_dim) {
    if (frame_dim < 64 || frame_dim % 64 || frame_dim > 2048) return 0;
    return 1024 + (size_t)(4 + 2 * (frame_dim / 64)) * tc::CHUNK + 4 * 128 * sizeof(float);
}

int cmhar_xattn_pack(const float* wq, const float* bq, const float* wk, const float* wv, const float* wo, const float* bo_folded,
                     const float* ln_weight, const float* ln_bias, int32_t frame_dim, void* blob, cmhar_stream_t s) {
    CMHAR_REQUIRE(wq && bq && wk && wv && wo && bo_folded && ln_weight && ln_bias && blob, "cmhar_xattn_pack: null argument");
    CMHAR_REQUIRE(cmhar_xattn_blob_bytes(frame_dim) != 0, "cmhar_xattn_pack: frame_dim %d must be a multiple of 64 in [64, 2048]", frame_dim);
    cudaStream_t st = (cudaStream_t)s;
    uint8_t* ch = reinterpret_cast<uint8_t*>(blob) + 1024;
    const int kc = frame_dim / 64;
    const float qscale = 0.25f * tc::LOG2E;              // 1/sqrt(head_dim) and the base-2 softmax folded into Wq, bq
    int c = 0;
    auto put = [&](const float* src, int ld, int col0, float scale) -> int {
        pack_chunk_kernel<<<4, 256, 0, st>>>(src, ld, 0, col0, 64, scale, ch + (size_t)c * tc::CHUNK);
        ++c;
        CMHAR_LAUNCH_CHECK();
        return CMHAR_OK;
    };
    for (int k = 0; k < 2; ++k) { const int rc = put(wq, D, 64 * k, qscale); if (rc != CMHAR_OK) return rc; }
    for (int k = 0; k < 2; ++k) { const int rc = put(wo, D, 64 * k, 1.f); if (rc != CMHAR_OK) return rc; }
    for (int k = 0; k < kc; ++k) {
        int rc = put(wk, frame_dim, 64 * k, 1.f); if (rc != CMHAR_OK) return rc;
        rc = put(wv, frame_dim, 64 * k, 1.f); if (rc != CMHAR_OK) return rc;
    }
    xattn::pack_xattn_params_kernel<<<1, 128, 0, st>>>(bq, bo_folded, ln_weight, ln_bias, qscale, reinterpret_cast<float*>(ch + (size_t)c * tc::CHUNK));
    CMHAR_LAUNCH_CHECK();
    BlobHeader h{};
    h.magic = xattn::XATTN_MAGIC; h.a = frame_dim; h.has_bf16 = 1;
    write_header_kernel<<<1, 1, 0, st>>>(reinterpret_cast<BlobHeader*>(blob), h);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_xattn_forward(const void* blob, const float* tokens, const void* frame_img, int64_t n, int32_t seq, int32_t frames, int32_t frame_dim,
                        float ln_eps, float* fused_out, cmhar_stream_t s) {
    if (n <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(blob && tokens && frame_img && fused_out, "cmhar_xattn_forward: null argument");
    if (frames != 16 || seq < 1 || seq > 16 || cmhar_xattn_blob_bytes(frame_dim) == 0) {
        set_error("cmhar_xattn_forward: the fused kernel serves 16 frames, 1..16 tokens and frame_dim %% 64 == 0 (got %d frames, %d tokens, dim %d)", frames, seq, frame_dim);
        return CMHAR_ERR_UNSUPPORTED;
    }
    CMHAR_REQUIRE((((uintptr_t)frame_img) & 1023) == 0 && (((uintptr_t)tokens | (uintptr_t)fused_out) & 15) == 0 && (((uintptr_t)blob) & 1023) == 0,
                  "cmhar_xattn_forward: blob / frame image must be 1024-byte aligned, rows 16-byte aligned");
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(xattn::xattn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, xattn::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    xattn::Args a{reinterpret_cast<const uint8_t*>(blob), tokens, reinterpret_cast<const uint8_t*>(frame_img), n, seq, frame_dim / 64, ln_eps, fused_out};
    const long long tiles = (n + 7) / 8;
    const int grid = (int)(tiles < (long long)sm_count() ? tiles : (long long)sm_count());
    xattn::xattn_tc_kernel<<<grid, xattn::NT, xattn::SMEM_BYTES, (cudaStream_t)s>>>(a);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // extern "C"
