// imu_encoder_bf16_pair.cu -- TWO tiles in flight per SM: the throughput form of the bf16 tcgen05 IMU encoder
// (same algebra and the same 2e-2 contract as imu_encoder_bf16.cu; reference src/models/models.py:30-50,100-132,
// torch/nn/modules/transformer.py:946-990).
//
// Why: with ONE 128-row tile per SM (imu_encoder_bf16.cu) a layer is a serial chain of ~20 MMA <-> epilogue hand-offs --
// the tensor pipe idles while the epilogue drains an accumulator and vice versa (ncu: tensor pipe active 32 % of the
// elapsed cycles).  A second tile would hide that, but one tile already takes all 512 TMEM columns and 96 KiB of operand
// tiles.  Here the two tiles ("sides") of an SM run HALF A LAYER APART:
//
//      slot s:     side (s & 1)      attention half of layer l   : K, V^T, Q GEMMs, scores, softmax, P V, out-projection, LN1
//                  the other side    feed-forward half of layer l': 8 x [FFN1 64-wide chunk -> ReLU -> FFN2 k-chunk], LN2
//
//   so the big resources are needed by one side at a time and are SHARED:
//     * TMEM (512 columns): residual R0 [0,128) | residual R1 [128,256) | attention scratch AS [256,448) | FFN scratch FS [448,512)
//         AS: K / V^T / Q accumulators in [0,128); Q (bf16 pairs) parked in [128,192); compact scores in [0,128); P (bf16) over the
//         scores it came from ([0,32) u [64,96)); O accumulators in the columns that are free by then ([32,64) u [96,128) u [128,192));
//         O (bf16) over P.  FS: one 64-wide FFN1 accumulator, ReLU'd bf16 written back over it (A operand of the FFN2 k-chunk).
//     * shared memory: h of side 0 | h of side 1 | K | V^T (32 KiB each; K / V^T belong to the side in its attention half)
//         | weight ring A (attention half, 2 x 16 KiB) | weight ring F (feed-forward half, 2 x 16 KiB) | parameter blocks.
//   The ONE MMA-issuing thread serves both sides and takes whichever side's next step is ready (mbarrier test), so the tensor
//   pipe works on one tile while the epilogue warps of the other drain; 8 epilogue warps per side (16 in all) follow their own
//   tile through a two-barrier handshake (RDY[side]: epilogue -> MMA, ACC[side]: tcgen05.commit -> epilogue).  tcgen05 MMAs
//   execute in issue order, which is what makes the shared scratch safe: a side's first attention MMA is issued after the other
//   side's last one, and it is the COMMIT of that MMA the epilogue waits for before it touches AS / K / V^T.
//   One producer lane feeds both rings (their chunk sequences are fixed by the slot order) without blocking on either.
//
// Weight images: a second section of the encoder blob in THIS kernel's consumption order -- per layer [Wk | Wv | Wq | Wo]
// (8 chunks, ring A) then [W1 rows 64j..64j+63 as two 64-row SW128 tiles | W2 k-chunk j] for j = 0..7 (16 chunks, ring F).
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace tcp {

using namespace tc;

constexpr int PAIR_CHUNKS_PER_LAYER = 24;
// shared memory map (bytes)
constexpr int P_H0 = 0, P_K = 65536, P_VT = 98304;
constexpr int P_WA = 131072, P_WF = 163840;              // 2 stages each
constexpr int P_SB = 196608;                             // static: final LN g,b (256 floats) | bo_fold[0] (128)
constexpr int P_SB_FLOATS = 384;
constexpr int P_PB = P_SB + P_SB_FLOATS * 4;             // [side][2][PB_FLOATS]
constexpr int P_ST = P_PB + 4 * PB_FLOATS * 4;           // LN statistics [side][flip][wq][128] float2
constexpr int P_BAR = P_ST + 2 * 2 * 2 * 128 * 8;
constexpr int P_SMEM = P_BAR + 256;
static_assert(P_SMEM <= 232448, "pair kernel exceeds the 227 KiB of shared memory a CTA may use");
static_assert(P_PB % 16 == 0 && P_ST % 16 == 0 && P_BAR % 8 == 0, "alignment");

enum {
    PB_WA_FULL = 0, PB_WA_EMPTY = 2, PB_WF_FULL = 4, PB_WF_EMPTY = 6,
    PB_ACC_A = 8,          // [side] attention-half MMA step complete (tcgen05.commit of the A-role MMA warp)
    PB_ACC_F = 10,         // [side] feed-forward-half MMA step complete (F-role MMA warp)
    PB_RDY_A = 12,         // [side] epilogue pass complete, next step is an attention-half step (8 warp arrivals)
    PB_RDY_F = 14,         // [side] ... next step is a feed-forward-half step
    PB_PBFULL = 16,        // [side][2]
    PB_PBEMPTY = 20,       // [side][2]
    PB_STATIC = 24,
    PB_COUNT = 25
};
static_assert(PB_COUNT * 8 + 8 <= 256, "barrier area too small");

constexpr uint32_t TP_AS = 256, TP_FS = 448;             // TMEM columns: R(side) = 128 * side
constexpr int NT_PAIR = 640;                             // 16 epilogue warps + 2 MMA warps (one per role) + 2 producer warps (one per ring)

struct Slot {
    bool a_valid, f_valid;
    int a_side, f_side, a_layer, f_layer, a_tile, f_tile;
};
// slot s of a CTA that owns m tiles (tile i runs on side i & 1; a tile is 2L half-steps long and side 1 starts one half later)
__device__ __forceinline__ Slot slot_info(int s, int m, int L) {
    Slot r;
    const int u = s & 1, v = u ^ 1, twoL = 2 * L;
    const int pa = (s - u) / twoL;
    r.a_side = u; r.a_tile = 2 * pa + u; r.a_valid = r.a_tile < m; r.a_layer = ((s - u) - pa * twoL) >> 1;
    const int q = s - v;
    r.f_side = v; r.f_valid = false; r.f_layer = 0; r.f_tile = 0;
    if (q >= 1) {
        const int pf = q / twoL;
        r.f_tile = 2 * pf + v; r.f_valid = r.f_tile < m; r.f_layer = ((q - pf * twoL) - 1) >> 1;
    }
    return r;
}
__device__ __forceinline__ int slot_count(int m, int L) { return m <= 0 ? 0 : ((m - 1) / 2) * 2 * L + ((m - 1) & 1) + 2 * L; }

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// mbarrier wait with an optional nanosleep back-off between polls (a warp that spins without sleeping takes issue slots from
// the MMA-issuing warp of the same SM sub-partition)
__device__ __forceinline__ void mbar_wait_bo(uint32_t bar, uint32_t parity, int site, unsigned sleep_ns) {
    uint32_t spins = 0;
    while (!mbar_test(bar, parity)) {
        if (sleep_ns) __nanosleep(sleep_ns);
        if (++spins > (1u << 24)) {
            if ((threadIdx.x & 31) == 0) printf("cmhar pair kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u site %d)\n",
                                                (int)blockIdx.x, (int)threadIdx.x, (bar & 0xffu) >> 3, parity, site);
            __trap();
        }
    }
}

// ======================================================================================== kernel
__global__ void __maxnreg__(96) imu_forward_bf16_pair_kernel(const Bf16Args args) {
    constexpr int CW = 64;
    extern __shared__ __align__(1024) uint8_t smem_tc[];
    uint8_t* const smem = smem_tc;
    const FwdArgs& a = args.f;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long trace_t0 = (tid == 0) ? trace_begin() : 0ull;
    const BlobHeader* eh = reinterpret_cast<const BlobHeader*>(a.enc_blob);
    const int S = eh->a, n_layers = eh->b;
    const size_t fp32_bytes = (EncLayout::fp32_floats(n_layers) * sizeof(float) + 1023) / 1024 * 1024;
    const uint8_t* wchunks1 = reinterpret_cast<const uint8_t*>(a.enc_blob) + 1024 + fp32_bytes;      // single-tile kernel's images
    const float* gparams = reinterpret_cast<const float*>(wchunks1 + (size_t)(1 + n_layers * CHUNKS_PER_LAYER) * CHUNK);
    const uint8_t* wchunks = wchunks1 + encoder_bf16_single_bytes(n_layers);                             // this kernel's images
    const long long tiles = (a.n + 7) / 8;
    const int m = (int)((tiles - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x);   // tiles of this CTA
    const int n_slots = slot_count(m, n_layers);
    int tlog_n = 0;          // development: (code, clock64) timeline of block 0, lane 0 of every warp (args.tlog, tools/timeline_pair.py)
#define TL(code) do { if (args.tlog && blockIdx.x == 0 && lane == 0 && tlog_n < TLOG_CAP) { \
        args.tlog[(warp * TLOG_CAP + tlog_n) * 2] = (code); args.tlog[(warp * TLOG_CAP + tlog_n) * 2 + 1] = clock64(); ++tlog_n; } } while (0)

    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + P_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P_BAR + 8 * PB_COUNT);

    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(BAR(PB_WA_FULL + i), 1);          // ring A / F full + empty
        for (int g = 0; g < 2; ++g) {
            mbar_init(BAR(PB_ACC_A + g), 1); mbar_init(BAR(PB_ACC_F + g), 1);
            mbar_init(BAR(PB_RDY_A + g), 8); mbar_init(BAR(PB_RDY_F + g), 8);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(BAR(PB_PBFULL + i), 1); mbar_init(BAR(PB_PBEMPTY + i), 8); }
        mbar_init(BAR(PB_STATIC), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // mbarrier polling is expensive on this machine (test_wait ~150 cycles, try_wait ~90 on a completed phase, ~60 to wake a
    // sleeping warp), so nobody polls two barriers: every role has ONE next thing to wait for and sleeps on it.
    if (warp == 18) {
        // ================================================================= producer of ring A (+ parameter blocks)
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last();
            mbar_expect_tx(BAR(PB_STATIC), P_SB_FLOATS * 4);
            bulk_g2s(sbase + P_SB, gparams + SB_FLN, P_SB_FLOATS * 4, BAR(PB_STATIC));
            uint32_t st = 0, par = 1;                       // fresh barriers: waiting on parity 1 passes
            uint32_t pb_bits = 0xFu;
            for (int s = 0; s < n_slots; ++s) {
                const Slot sl = slot_info(s, m, n_layers);
                if (!sl.a_valid) continue;
                const int bi = sl.a_side * 2 + (sl.a_layer & 1);
                mbar_wait(BAR(PB_PBEMPTY + bi), (pb_bits >> bi) & 1u, 19);
                pb_bits ^= (1u << bi);
                mbar_expect_tx(BAR(PB_PBFULL + bi), PB_FLOATS * 4);
                bulk_g2s(sbase + P_PB + bi * PB_FLOATS * 4, gparams + SB_FLOATS + (size_t)sl.a_layer * PB_FLOATS, PB_FLOATS * 4, BAR(PB_PBFULL + bi));
                const int first = (sl.a_layer == 0) ? 1 : 0;              // list = [patch chunk] + 8 chunks of the layer
                for (int j = 0; j < first + 8; ++j) {
                    const int c = (first && j == 0) ? 0 : 1 + sl.a_layer * PAIR_CHUNKS_PER_LAYER + (j - first);
                    mbar_wait(BAR(PB_WA_EMPTY + st), par, 1);
                    if (args.ablate & ABL_NO_TMA) mbar_arrive(BAR(PB_WA_FULL + st));
                    else {
                        mbar_expect_tx(BAR(PB_WA_FULL + st), CHUNK);
                        bulk_g2s_hint(sbase + P_WA + st * CHUNK, wchunks + (size_t)c * CHUNK, CHUNK, BAR(PB_WA_FULL + st), keep);
                    }
                    TL(c);
                    if (++st == 2) { st = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 19) {
        // ================================================================= producer of ring F
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last();
            uint32_t st = 0, par = 1;
            for (int s = 0; s < n_slots; ++s) {
                const Slot sl = slot_info(s, m, n_layers);
                if (!sl.f_valid) continue;
                for (int j = 0; j < 16; ++j) {
                    const int c = 1 + sl.f_layer * PAIR_CHUNKS_PER_LAYER + 8 + j;
                    mbar_wait(BAR(PB_WF_EMPTY + st), par, 1);
                    if (args.ablate & ABL_NO_TMA) mbar_arrive(BAR(PB_WF_FULL + st));
                    else {
                        mbar_expect_tx(BAR(PB_WF_FULL + st), CHUNK);
                        bulk_g2s_hint(sbase + P_WF + st * CHUNK, wchunks + (size_t)c * CHUNK, CHUNK, BAR(PB_WF_FULL + st), keep);
                    }
                    TL(c);
                    if (++st == 2) { st = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 16) {
        // ================================================================= MMA issuer of the attention halves (warp converged)
        const bool leader = elect_one();
        const bool no_dense = (args.ablate & ABL_NO_DENSE_MMA) != 0, no_attn = (args.ablate & ABL_NO_ATTN_MMA) != 0;     // timing ablations (development)
        constexpr uint32_t ID128 = idesc_bf16(128, 128), ID16 = idesc_bf16(128, 16);
        const uint64_t CH = CHUNK >> 4;
        const uint64_t dK = sw128_desc(sbase + P_K), dVT = sw128_desc(sbase + P_VT);
        uint32_t stA = 0, parA = 0;
        uint32_t rdy_bits = 0;                                        // parity per side of RDY_A
        auto ringA = [&]() -> uint64_t {
            mbar_wait(BAR(PB_WA_FULL + stA), parA, 2);
            tc_fence_after();
            return sw128_desc(sbase + P_WA + stA * CHUNK);
        };
        auto ringA_done = [&]() { if (leader) tc_commit(BAR(PB_WA_EMPTY + stA)); if (++stA == 2) { stA = 0; parA ^= 1; } };
        for (int s = 0; s < n_slots; ++s) {
            const Slot sl = slot_info(s, m, n_layers);
            if (!sl.a_valid) continue;
            const int g = sl.a_side;
            const uint32_t R = tmem + 128u * (uint32_t)g, AS = tmem + TP_AS;
            const uint64_t dH = sw128_desc(sbase + P_H0 + g * 32768);
            // steps: -1 patch embedding (layer 0 only), 0 K, 1 V^T, 2 Q, 3 scores, 4 P V, 5 out-projection
            for (int step = (sl.a_layer == 0 ? -1 : 0); step < 6; ++step) {
                mbar_wait(BAR(PB_RDY_A + g), (rdy_bits >> g) & 1u, 4);
                rdy_bits ^= (1u << g);
                tc_fence_after();
                TL(1000 * g + 10 + step);
                if (step == -1) {
                    const uint64_t dW = ringA();
                    for (int k = 0; k < 3; ++k)
                        if (leader && !no_dense) umma(R, dH + (uint64_t)(2 * k), dW + (uint64_t)(2 * k), ID128, 1u);
                    ringA_done();
                } else if (step <= 2) {
                    for (int c = 0; c < 2; ++c) {
                        const uint64_t dW = ringA();
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t wa = dW + (uint64_t)(2 * k), oa = dH + (uint64_t)c * CH + (uint64_t)(2 * k);
                            if (leader && !no_dense) umma(AS, step == 1 ? wa : oa, step == 1 ? oa : wa, ID128, (c > 0 || k > 0) ? 1u : 0u);
                        }
                        ringA_done();
                    }
                } else if (step == 3) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            const uint64_t off = (uint64_t)(h >> 2) * CH + (uint64_t)((h & 3) * 2);
                            if (leader && !no_attn) umma_ts_rows16(AS + 16 * h, AS + 128 + 8 * h, dK + off + (uint64_t)(j * 128), ID16, j);
                        }
                    }
                } else if (step == 4) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint64_t koff = (uint64_t)(j >> 2) * CH + (uint64_t)((j & 3) * 2);
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            const uint32_t ocol = (h < 2) ? 32 + 16 * h : (h < 4) ? 96 + 16 * (h - 2) : 128 + 16 * (h - 4);
                            if (leader && !no_attn) umma_ts_rows16(AS + ocol, AS + tm_bf16_col<CW>(16 * h), dVT + koff + (uint64_t)(h * 128), ID16, j);
                        }
                    }
                } else {
                    for (int kc = 0; kc < 2; ++kc) {
                        const uint64_t dW = ringA();
                        for (int k = 0; k < 4; ++k)
                            if (leader && !no_dense) umma_ts(R, AS + tm_bf16_col<CW>(64 * kc + 16 * k), dW + (uint64_t)(2 * k), ID128, 1u);
                        ringA_done();
                    }
                }
                if (leader) tc_commit(BAR(PB_ACC_A + g));
                TL(1000 * g + 510 + step);
            }
        }
    } else if (warp == 17) {
        // ================================================================= MMA issuer of the feed-forward halves (warp converged)
        const bool leader = elect_one();
        const bool no_dense = (args.ablate & ABL_NO_DENSE_MMA) != 0;
        constexpr uint32_t ID128 = idesc_bf16(128, 128), ID64 = idesc_bf16(128, 64);
        const uint64_t CH = CHUNK >> 4;
        uint32_t stF = 0, parF = 0;
        uint32_t rdy_bits = 0;                                        // parity per side of RDY_F
        auto ringF = [&]() -> uint64_t {
            mbar_wait(BAR(PB_WF_FULL + stF), parF, 3);
            tc_fence_after();
            return sw128_desc(sbase + P_WF + stF * CHUNK);
        };
        auto ringF_done = [&]() { if (leader) tc_commit(BAR(PB_WF_EMPTY + stF)); if (++stF == 2) { stF = 0; parF ^= 1; } };
        for (int s = 0; s < n_slots; ++s) {
            const Slot sl = slot_info(s, m, n_layers);
            if (!sl.f_valid) continue;
            const int g = sl.f_side;
            const uint32_t R = tmem + 128u * (uint32_t)g, FS = tmem + TP_FS;
            const uint64_t dH = sw128_desc(sbase + P_H0 + g * 32768);
            // steps: 0 = FFN1 chunk 0; 1..7 = FFN2 k-chunk (step-1) then FFN1 chunk step; 8 = FFN2 k-chunk 7
            for (int step = 0; step < 9; ++step) {
                mbar_wait(BAR(PB_RDY_F + g), (rdy_bits >> g) & 1u, 5);
                rdy_bits ^= (1u << g);
                tc_fence_after();
                TL(1000 * g + 100 + step);
                if (step >= 1) {
                    const uint64_t dW = ringF();
                    for (int k = 0; k < 4; ++k)
                        if (leader && !no_dense) umma_ts(R, FS + tm_bf16_col<32>(16 * k), dW + (uint64_t)(2 * k), ID128, 1u);
                    ringF_done();
                }
                if (step <= 7) {
                    const uint64_t dW = ringF();
                    for (int kh = 0; kh < 2; ++kh)
                        for (int k = 0; k < 4; ++k)
                            if (leader && !no_dense) umma(FS, dH + (uint64_t)kh * CH + (uint64_t)(2 * k), dW + (uint64_t)kh * (8192 >> 4) + (uint64_t)(2 * k), ID64,
                                                          (kh > 0 || k > 0) ? 1u : 0u);
                    ringF_done();
                }
                if (leader) tc_commit(BAR(PB_ACC_F + g));
                TL(1000 * g + 600 + step);
            }
        }
    } else {
        // ================================================================= epilogue: 8 warps per side follow their own tiles
        const int g = warp >> 3, wq = (warp >> 2) & 1, quarter = warp & 3;
        const int row = quarter * 32 + lane;                 // token row == TMEM lane
        const int win = row >> 4, tok = row & 15;
        const int c0 = wq * CW;
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const uint32_t tR = lane_base + 128u * (uint32_t)g, tAS = lane_base + TP_AS, tFS = lane_base + TP_FS;
        uint8_t* const Hs = smem + P_H0 + g * 32768;
        float2* const stats = reinterpret_cast<float2*>(smem + P_ST) + g * (2 * 2 * 128);
        const float* SB = reinterpret_cast<const float*>(smem + P_SB);        // [final LN g | b | bo_fold(0)]
        const float* tok_bias = gparams + SB_TOK;
        auto chunk_of = [](int c) { return (c >> 6) * CHUNK; };
        auto piece_of = [](int c) { return (c & 63) >> 3; };
        auto ld4 = [](const float* p) { return *reinterpret_cast<const float4*>(p); };
        uint32_t acc_par = 0;
        const bool no_epi = (args.ablate & 64) != 0;          // timing ablation (development): hand-shakes only, no epilogue work
        uint32_t v[32];
        float f[32];
        uint32_t accf_par = 0;
        auto wait_acc = [&](int site) { TL(site); mbar_wait(BAR(PB_ACC_A + g), acc_par, site); acc_par ^= 1; tc_fence_after(); TL(100 + site); };
        auto wait_acc_f = [&](int site) { TL(site); mbar_wait(BAR(PB_ACC_F + g), accf_par, site); accf_par ^= 1; tc_fence_after(); TL(100 + site); };
        // `to_f`: the step this pass unblocks is issued by the feed-forward-half MMA warp (else the attention-half one)
        auto publish = [&](bool to_f = false) {   // generic-proxy smem writes + TMEM accesses visible to the MMA thread, then one arrival per warp
            tc_wait_st();
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR((to_f ? PB_RDY_F : PB_RDY_A) + g));
            TL(200);
        };
        auto publish_tmem = [&](bool to_f = false) {
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR((to_f ? PB_RDY_F : PB_RDY_A) + g));
            TL(201);
        };
        // y (32 finished fp32 columns c0+cc..) -> h (bf16, A/B operand) and y + next_bias -> R
        auto write_h = [&](const float* y32, int cc, const float* next_bias) {
            store_bf16_32(Hs + chunk_of(c0 + cc), row, piece_of(c0 + cc), y32);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 nb = ld4(next_bias + c0 + cc + i);
                v[i] = __float_as_uint(y32[i] + nb.x); v[i + 1] = __float_as_uint(y32[i + 1] + nb.y);
                v[i + 2] = __float_as_uint(y32[i + 2] + nb.z); v[i + 3] = __float_as_uint(y32[i + 3] + nb.w);
            }
            TMEM_ST32(tR + c0 + cc, v);
        };
        uint32_t ln_flip = 0;
        auto quarter_bar = [&] { asm volatile("bar.sync %0, 64;" ::"r"(1 + 4 * g + quarter) : "memory"); };
        // mean / rstd of this thread's residual row (its two column halves live in two threads of the same lane quarter)
        auto row_stats = [&](const float s1, const float s2, float& mean, float& rstd) {
            float2* st = stats + ln_flip * (2 * 128);
            ln_flip ^= 1;
            st[wq * 128 + row] = make_float2(s1, s2);
            quarter_bar();
            const float2 o0 = st[row], o1 = st[128 + row];
            mean = (o0.x + o1.x) * (1.f / D);
            rstd = rsqrtf(fmaxf((o0.y + o1.y) * (1.f / D) - mean * mean, 0.f) + LN_EPS);
        };
        // LayerNorm of the residual row in R: statistics in one pass over the slice, normalisation in a second (re-read from TMEM)
        // `out_stats` non-null (last layer): the normalised row goes back into R as fp32 and its (sum, sum of squares) are returned
        auto layer_norm_R = [&](const float* gb, const float* next_bias, float* out_stats) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int cc = 0; cc < CW; cc += 32) {
                TMEM_LD32(tR + c0 + cc, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) { const float xv = __uint_as_float(v[i]); s1 += xv; s2 = fmaf(xv, xv, s2); }
            }
            float mean, rstd;
            row_stats(s1, s2, mean, rstd);
            float o1 = 0.f, o2 = 0.f;
#pragma unroll
            for (int cc = 0; cc < CW; cc += 32) {
                TMEM_LD32(tR + c0 + cc, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 gg = ld4(gb + c0 + cc + i);
                    const float4 bb = ld4(gb + D + c0 + cc + i);
                    f[i] = (__uint_as_float(v[i]) - mean) * rstd * gg.x + bb.x;
                    f[i + 1] = (__uint_as_float(v[i + 1]) - mean) * rstd * gg.y + bb.y;
                    f[i + 2] = (__uint_as_float(v[i + 2]) - mean) * rstd * gg.z + bb.z;
                    f[i + 3] = (__uint_as_float(v[i + 3]) - mean) * rstd * gg.w + bb.w;
                }
                if (out_stats) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) { o1 += f[i]; o2 = fmaf(f[i], f[i], o2); v[i] = __float_as_uint(f[i]); }
                    TMEM_ST32(tR + c0 + cc, v);
                } else {
                    write_h(f, cc, next_bias);
                }
            }
            if (out_stats) { out_stats[0] = o1; out_stats[1] = o2; }
        };
        auto store_tmem_bf16 = [&](uint32_t taddr, const float* y32) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(y32[2 * i], y32[2 * i + 1]);
            TMEM_ST16(taddr, pk);
        };
        auto dump_R = [&](long long tile_idx) {
            float* dst = args.dbg_out + ((size_t)tile_idx * 128 + row) * D + c0;
#pragma unroll
            for (int cc = 0; cc < CW; cc += 32) {
                TMEM_LD32(tR + c0 + cc, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) dst[cc + i] = __uint_as_float(v[i]);
            }
        };

        mbar_wait(BAR(PB_STATIC), 0, 20);
        uint32_t pb_bits = 0;
        for (int ti = g; ti < m; ti += 2) {
            const long long tile = (long long)blockIdx.x + (long long)ti * (long long)gridDim.x;
            const long long w0 = tile * 8;
            // ---- stage patches (bf16, K = 48 split precision -> pieces 0..5 of chunk 0 of h) and preload R = tok_bias
            if (wq == 0) {
                float p16[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) p16[i] = 0.f;
                if (tok > 0 && tok < S && w0 + win < a.n) {
                    const float* src = a.x + (w0 + win) * a.xstride + (tok - 1) * P;
#pragma unroll
                    for (int i = 0; i < 16; ++i) p16[i] = __ldg(src + i);
                }
                float lo16[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float hi = __bfloat162float(__float2bfloat16_rn(p16[i]));
                    lo16[i] = p16[i] - hi;
                }
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const float* srcv = (q == 2 || q == 3) ? lo16 : p16;
                    const int o = (q & 1) * 8;
                    uint4 u;
                    u.x = pack_bf16(srcv[o + 0], srcv[o + 1]); u.y = pack_bf16(srcv[o + 2], srcv[o + 3]);
                    u.z = pack_bf16(srcv[o + 4], srcv[o + 5]); u.w = pack_bf16(srcv[o + 6], srcv[o + 7]);
                    *reinterpret_cast<uint4*>(Hs + sw128_off(row, q)) = u;
                }
            }
            {
                const float* tb = tok_bias + (tok < S ? tok : 0) * D + c0;
#pragma unroll
                for (int cc = 0; cc < CW; cc += 32) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(tb + cc + i));
                        v[i] = __float_as_uint(t.x); v[i + 1] = __float_as_uint(t.y);
                        v[i + 2] = __float_as_uint(t.z); v[i + 3] = __float_as_uint(t.w);
                    }
                    TMEM_ST32(tR + c0 + cc, v);
                }
            }
            publish();
            // ---- h0 = R ; h = bf16(h0) ; R = h0 + b_o(layer 0)
            wait_acc(11);
            if (!no_epi) {
#pragma unroll
            for (int cc = 0; cc < CW; cc += 32) {
                TMEM_LD32(tR + c0 + cc, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                write_h(f, cc, SB + 256);
            }
            }
            if (args.dbg_stage == 0) { tc_wait_st(); dump_R(tile); }
            publish();

            for (int l = 0; l < n_layers; ++l) {
                const bool last = (l + 1 == n_layers);
                const int pbi = g * 2 + (l & 1);
                const float* PB = reinterpret_cast<const float*>(smem + P_PB + pbi * PB_FLOATS * 4);
                // ---- K and V^T accumulators (AS[0,128)) -> bf16 B-operand tiles in shared memory
#pragma unroll 1
                for (int mm = 0; mm < 2; ++mm) {
                    wait_acc(12);
                    uint8_t* dst = smem + (mm == 0 ? P_K : P_VT);
                    if (!no_epi)
#pragma unroll
                    for (int cc = 0; cc < CW; cc += 32) {
                        TMEM_LD32(tAS + c0 + cc, v);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                        store_bf16_32(dst + chunk_of(c0 + cc), row, piece_of(c0 + cc), f);
                    }
                    publish();
                }
                // ---- Q (+ pre-scaled bias) -> bf16 pairs parked in AS[128,192) (A operand of the score MMAs)
                wait_acc(13);
                mbar_wait(BAR(PB_PBFULL + pbi), (pb_bits >> (l & 1)) & 1u, 21);
                pb_bits ^= (1u << (l & 1));
                if (!no_epi)
#pragma unroll
                for (int cc = 0; cc < CW; cc += 32) {
                    TMEM_LD32(tAS + c0 + cc, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b = ld4(PB + PB_BQ + c0 + cc + i);
                        f[i] = __uint_as_float(v[i]) + b.x; f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
                        f[i + 2] = __uint_as_float(v[i + 2]) + b.z; f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
                    }
                    store_tmem_bf16(tAS + 128 + ((c0 + cc) >> 1), f);
                }
                publish_tmem();
                // ---- softmax over the compact scores: this thread owns heads 4 wq .. 4 wq + 3 of its row
                wait_acc(14);
                auto softmax_rows = [&](auto full_tag) {
#pragma unroll
                    for (int cc = 0; cc < CW; cc += 32) {
                        TMEM_LD32(tAS + c0 + cc, v);
                        tc_wait_ld();
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            float mx = -INFINITY;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float sv = (decltype(full_tag)::value || i < S) ? __uint_as_float(v[hh * 16 + i]) : -INFINITY;
                                f[hh * 16 + i] = sv;
                                mx = fmaxf(mx, sv);
                            }
                            float den = 0.f;
#pragma unroll
                            for (int i = 0; i < 16; ++i) { f[hh * 16 + i] = ex2_approx(f[hh * 16 + i] - mx); den += f[hh * 16 + i]; }
                            const float inv = rcp_approx(den);
#pragma unroll
                            for (int i = 0; i < 16; ++i) f[hh * 16 + i] *= inv;
                        }
                        store_tmem_bf16(tAS + c0 + (cc >> 1), f);         // P over the scores it came from
                    }
                };
                if (no_epi) {}
                else if (S == CMHAR_MAX_SEQ) softmax_rows(std::true_type{});
                else softmax_rows(std::false_type{});
                publish_tmem();
                // ---- O accumulators (scattered over the free columns of AS) -> bf16 pairs over P (A operand of the out-projection)
                wait_acc(15);
                if (!no_epi)
#pragma unroll
                for (int cc = 0; cc < CW; cc += 32) {
                    const uint32_t src = wq ? (uint32_t)(128 + cc) : (uint32_t)(32 + 2 * cc);      // heads 0,1 | 2,3 | 4..7
                    TMEM_LD32(tAS + src, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                    store_tmem_bf16(tAS + c0 + (cc >> 1), f);
                }
                publish_tmem();
                // ---- LN1: h1 = LN(R) ; h = bf16(h1) ; R = h1 + b_2
                wait_acc(16);
                if (!no_epi) layer_norm_R(PB + PB_LN1, PB + PB_B2, nullptr);
                if (args.dbg_stage == 1 && l == 0) { tc_wait_st(); dump_R(tile); }
                publish(true);
                // ---- 8 FFN1 chunks of 64 hidden units: relu(acc + b_1) -> bf16 pairs over the accumulator (A operand of FFN2)
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    wait_acc_f(17);
                    if (!no_epi) {
                    TMEM_LD32(tFS + 32 * wq, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b = ld4(PB + PB_B1 + j * 64 + 32 * wq + i);
                        f[i] = __uint_as_float(v[i]) + b.x; f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
                        f[i + 2] = __uint_as_float(v[i + 2]) + b.z; f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
                    }
                    {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16_relu(f[2 * i], f[2 * i + 1]);
                        TMEM_ST16(tFS + 32 * wq, pk);
                    }
                    }
                    publish_tmem(true);
                }
                // ---- LN2: h2 = LN(R) ; h = bf16(h2) ; R = h2 + b_o(next layer)
                wait_acc_f(18);
                if (!last) {
                    if (!no_epi) layer_norm_R(PB + PB_LN2, PB + PB_BON, nullptr);
                    if (args.dbg_stage == 2 && l == 0) { tc_wait_st(); dump_R(tile); }
                    publish();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(PB_PBEMPTY + pbi));
                } else {
                    // last layer: LN2 (back into R), then the encoder's final LayerNorm (models.py:127), then the outputs,
                    // 32 columns at a time (registers are capped at 112: no 64-float row copy)
                    float os[2];
                    layer_norm_R(PB + PB_LN2, nullptr, os);
                    float mean, rstd;
                    row_stats(os[0], os[1], mean, rstd);
                    tc_wait_st();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(PB_PBEMPTY + pbi));
                    const bool valid = (w0 + win < a.n);
                    const long long w = w0 + win;
#pragma unroll
                    for (int cc = 0; cc < CW; cc += 32) {
                        TMEM_LD32(tR + c0 + cc, v);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = (__uint_as_float(v[i]) - mean) * rstd * SB[c0 + cc + i] + SB[D + c0 + cc + i];
                        if (a.tokens_out && valid && tok < S) {
                            float4* dst = reinterpret_cast<float4*>(a.tokens_out + ((w0 + win) * S + tok) * D + c0 + cc);
#pragma unroll
                            for (int i = 0; i < 8; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                        }
                        if (tok == 0) {
                            if (a.cls_out && valid) {
                                float4* dst = reinterpret_cast<float4*>(a.cls_out + (w0 + win) * D + c0 + cc);
#pragma unroll
                                for (int i = 0; i < 8; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                            }
                            if (a.cls_img) {
                                // the CLS row as row (w % 128) of the bf16 operand image the projection-head / fusion kernels
                                // stream as their A operand; windows past n (last tile) are written as zero rows
                                if (!valid) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) f[i] = 0.f;
                                }
                                uint8_t* img = reinterpret_cast<uint8_t*>(a.cls_img) + (size_t)(w >> 7) * (2 * CHUNK);
                                store_bf16_32(img + chunk_of(c0 + cc), (int)(w & 127), piece_of(c0 + cc), f);
                            }
                        }
                    }
                }
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
    if (tid == 0) trace_end(TRACE_ENCODER, trace_t0);
#undef TL
}

// ================================================================================ weight packing (pair order)
// W1 rows row0 .. row0+63 as TWO [64 rows x 64 k] SWIZZLE_128B K-major tiles (k 0..63 | k 64..127), 8 KiB each
__global__ void pack_w1_half_chunk_kernel(const float* __restrict__ w1 /*(512,128)*/, int row0, uint8_t* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (tile, row, 16-byte piece)
    if (i >= 2 * 64 * 8) return;
    const int t = i >> 9, r = (i >> 3) & 63, j = i & 7;
    float vals[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) vals[e] = w1[(size_t)(row0 + r) * D + 64 * t + j * 8 + e];
    uint4 u;
    u.x = pack_bf16(vals[0], vals[1]); u.y = pack_bf16(vals[2], vals[3]);
    u.z = pack_bf16(vals[4], vals[5]); u.w = pack_bf16(vals[6], vals[7]);
    *reinterpret_cast<uint4*>(dst + 8192 * t + sw128_off(r, j)) = u;
}

}  // namespace tcp

size_t encoder_bf16_pair_bytes(int layers) { return (size_t)(1 + layers * tcp::PAIR_CHUNKS_PER_LAYER) * tc::CHUNK; }

// `dst` = the pair section (behind the single-tile kernel's images and the parameter blocks)
int pack_encoder_bf16_pair(const cmhar_imu_encoder_params* p, void* dst_section, cudaStream_t st) {
    using namespace tc;
    uint8_t* dst = reinterpret_cast<uint8_t*>(dst_section);
    int c = 0;
    auto put = [&](const float* src, int ld, int row0, int col0, float scale) -> int {
        pack_chunk_kernel<<<4, 256, 0, st>>>(src, ld, row0, col0, 64, scale, dst + (size_t)c * CHUNK);
        ++c;
        CMHAR_LAUNCH_CHECK();
        return CMHAR_OK;
    };
#define PUTP(...) do { int _rc = put(__VA_ARGS__); if (_rc != CMHAR_OK) return _rc; } while (0)
    pack_patch_chunk_kernel<<<4, 256, 0, st>>>(p->patch_weight, dst);
    ++c;
    CMHAR_LAUNCH_CHECK();
    for (int l = 0; l < p->layers; ++l) {
        const cmhar_encoder_layer_params& q = p->layer[l];
        for (int k = 0; k < 2; ++k) PUTP(q.in_proj_weight, D, 1 * D, k * 64, 1.f);                    // Wk
        for (int k = 0; k < 2; ++k) PUTP(q.in_proj_weight, D, 2 * D, k * 64, 1.f);                    // Wv
        for (int k = 0; k < 2; ++k) PUTP(q.in_proj_weight, D, 0, k * 64, 0.25f * LOG2E);              // Wq (softmax uses 2^x)
        for (int k = 0; k < 2; ++k) PUTP(q.out_proj_weight, D, 0, k * 64, 1.f);                       // Wo
        for (int j = 0; j < 8; ++j) {
            tcp::pack_w1_half_chunk_kernel<<<4, 256, 0, st>>>(q.linear1_weight, 64 * j, dst + (size_t)c * CHUNK);
            ++c;
            CMHAR_LAUNCH_CHECK();
            PUTP(q.linear2_weight, FF, 0, 64 * j, 1.f);                                                  // W2 k-chunk j
        }
    }
#undef PUTP
    return CMHAR_OK;
}

int launch_imu_forward_bf16_pair(const Bf16Args& args, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(tcp::imu_forward_bf16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::P_SMEM));
        configured[dev & 63] = true;
    }
    const long long tiles = (args.f.n + 7) / 8;
    const long long pairs = (tiles + 1) / 2;
    const int grid = (int)((pairs < (long long)sm_count()) ? pairs : (long long)sm_count());
    tcp::imu_forward_bf16_pair_kernel<<<grid, tcp::NT_PAIR, tcp::P_SMEM, stream>>>(args);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
