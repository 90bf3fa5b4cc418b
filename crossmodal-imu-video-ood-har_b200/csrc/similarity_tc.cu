// similarity_tc.cu -- bf16 tcgen05 path of the contrastive similarity matrix (CMHAR_BF16, dim % 64 == 0).
//
// Same contract as similarity.cu (reference src/models/losses.py:25-54,67-87): S = A B^T is
// reduced tile by tile -- softplus(-(s*t+b)) sum, per-row (max, sum-exp) partials, diagonal, optional
// materialised S -- and never written to HBM unless asked.  Here the 128 x 128 tiles are tcgen05 MMAs:
//   * a pre-pass converts both embedding matrices once into bf16 SWIZZLE_128B chunk images
//     ([128 rows x 64 k] = 16 KiB each, zero padded), so every operand tile is a plain cp.async.bulk copy;
//   * one CTA keeps its 128-row A tile resident in shared memory and streams its share of the B tiles
//     through a 6-stage mbarrier ring (weight-stationary: 64 KiB of ingest per 16 MMAs);
//   * fp32 accumulators are double buffered in TMEM (2 x 128 columns): the 8 epilogue warps reduce tile t
//     (MUFU-bound: ex2 + lg2 per element) while the tensor pipe computes tile t+1;
//   * column statistics need a reduction across TMEM lanes; instead the caller runs a second pass with the
//     operands swapped (S^T tiles, MMA time is ~1/3 of the epilogue) and takes row statistics again.
// B = 4096, dim = 256: 8.6 GFLOP, 1024 tiles; the fp32 CUDA-core kernel needs ~0.35 ms for it.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace simtc {

using namespace tc;

constexpr int CHUNK = 16384;
constexpr int NSTAGE = 6;
constexpr int MAX_KC = 4;                                   // dim <= 256
constexpr int OFF_A = 0;                                    // resident A tile: MAX_KC chunks
constexpr int OFF_B = MAX_KC * CHUNK;                       // B ring
constexpr int OFF_BAR = OFF_B + NSTAGE * CHUNK;
constexpr int SMEM_BYTES = OFF_BAR + 256;
enum { B_AFULL = 0, B_FULL = 1, B_EMPTY = B_FULL + NSTAGE, B_ACCFULL = B_EMPTY + NSTAGE, B_ACCEMPTY = B_ACCFULL + 2, B_COUNT = B_ACCEMPTY + 2 };
static_assert(B_COUNT * 8 + 8 <= 160, "barrier area too small (the per-warp loss partials sit at +160)");
constexpr int NT = 8 * 32 + 64;                             // 8 epilogue warps, MMA warp, producer warp

struct Args {
    const uint8_t* a_img;      // [row tile][kc] chunk images
    const uint8_t* b_img;
    long long na, nb;
    int kc;                    // dim / 64
    long long diag_offset;
    float* sim_out;            // (na, nb) or null; `transposed`: written as sim_out[c * ld + r]
    int transposed;
    float zscale, zbias;       // z * log2(e) = s * zscale + zbias
    double* sigmoid_sum;
    float lse_scale;           // v * log2(e)... see kernel
    float2* row_part;          // [2 * n_col_tiles][na] or null
    float* diag_out;
    // B operand partitioned over several buffers (ranks' shards, possibly peer-GPU memory mapped over NVLink):
    // column tile ct lives in b_parts[ct / tiles_per_part] at local tile ct % tiles_per_part.  tiles_per_part == 0: b_img.
    const uint8_t* b_parts[CMHAR_MAX_PEERS];
    long long tiles_per_part;
    // deterministic loss reduction without a zeroed accumulator: every CTA stores its partial into cta_part[cta], the
    // last CTA to finish (ticket) adds them in index order, scales, stores the result to every sum_dst and re-arms the ticket
    double* cta_part;
    unsigned int* ticket;
    double* sum_dst[CMHAR_MAX_PEERS];
    int n_dst;
    double out_scale;
};

// fp32 (n, dim) row-major -> bf16 chunk images [ceil(n/128)][dim/64][128 x 64 SW128]
__global__ void to_chunk_images_kernel(const float* __restrict__ src, long long n, int dim, uint8_t* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per (row, 16-byte piece)
    const int pieces = dim / 8;
    const long long n_pad = (n + 127) / 128 * 128;
    if (i >= n_pad * pieces) return;
    const long long r = i / pieces;
    const int p = (int)(i % pieces), kc = p >> 3, j = p & 7;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (r < n) {
        lo = __ldg(reinterpret_cast<const float4*>(src + r * dim + p * 8));
        hi = __ldg(reinterpret_cast<const float4*>(src + r * dim + p * 8 + 4));
    }
    uint4 u;
    u.x = pack_bf16(lo.x, lo.y); u.y = pack_bf16(lo.z, lo.w); u.z = pack_bf16(hi.x, hi.y); u.w = pack_bf16(hi.z, hi.w);
    uint8_t* chunk = dst + ((size_t)(r >> 7) * (dim / 64) + kc) * CHUNK;
    *reinterpret_cast<uint4*>(chunk + sw128_off((int)(r & 127), j)) = u;
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(NT, 1) similarity_tc_kernel(const Args p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long trace_t0 = (tid == 0) ? trace_begin() : 0ull;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    const long long n_ct = (p.nb + 127) / 128;
    const long long rt = blockIdx.y;
    const int G = gridDim.x, g = blockIdx.x;
    const int my_tiles = (int)((n_ct - g + G - 1) / G);      // column tiles g, g+G, ...
    constexpr int MMA_WARP = 8, LOAD_WARP = 9;

    if (tid == 0) {
        mbar_init(BAR(B_AFULL), 1);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(BAR(B_ACCFULL + b), 1); mbar_init(BAR(B_ACCEMPTY + b), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == LOAD_WARP) {
        if (lane == 0) {
            mbar_expect_tx(BAR(B_AFULL), (uint32_t)(p.kc * CHUNK));
            for (int kc = 0; kc < p.kc; ++kc)
                bulk_g2s(sbase + OFF_A + kc * CHUNK, p.a_img + ((size_t)rt * p.kc + kc) * CHUNK, CHUNK, BAR(B_AFULL));
            uint32_t stage = 0, parity = 1;
            for (int t = 0; t < my_tiles; ++t) {
                const long long ct = g + (long long)t * G;
                for (int kc = 0; kc < p.kc; ++kc) {
                    mbar_wait(BAR(B_EMPTY + stage), parity, 50);
                    mbar_expect_tx(BAR(B_FULL + stage), CHUNK);
                    const uint8_t* src = p.tiles_per_part
                        ? p.b_parts[ct / p.tiles_per_part] + ((size_t)(ct % p.tiles_per_part) * p.kc + kc) * CHUNK
                        : p.b_img + ((size_t)ct * p.kc + kc) * CHUNK;
                    bulk_g2s(sbase + OFF_B + stage * CHUNK, src, CHUNK, BAR(B_FULL + stage));
                    if (++stage == NSTAGE) { stage = 0; parity ^= 1; }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        const bool leader = elect_one();
        constexpr uint32_t ID128 = idesc_bf16(128, 128);
        uint32_t stage = 0, parity = 0, acc_parity[2] = {1, 1};
        mbar_wait(BAR(B_AFULL), 0, 51);
        tc_fence_after();
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t & 1;
            mbar_wait(BAR(B_ACCEMPTY + buf), acc_parity[buf], 52);     // epilogue has drained this accumulator
            acc_parity[buf] ^= 1;
            tc_fence_after();
            for (int kc = 0; kc < p.kc; ++kc) {
                mbar_wait(BAR(B_FULL + stage), parity, 53);
                tc_fence_after();
                const uint64_t dA = sw128_desc(sbase + OFF_A + kc * CHUNK), dB = sw128_desc(sbase + OFF_B + stage * CHUNK);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (leader) umma(tmem + 128 * buf, dA + (uint64_t)(2 * k), dB + (uint64_t)(2 * k), ID128, (kc > 0 || k > 0) ? 1u : 0u);
                if (leader) tc_commit(BAR(B_EMPTY + stage));
                if (++stage == NSTAGE) { stage = 0; parity ^= 1; }
            }
            if (leader) tc_commit(BAR(B_ACCFULL + buf));
        }
    } else {
        // ------------------------------------------------------------- epilogue: thread = (row, 64-column half)
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const long long r = rt * 128 + row;
        const bool row_ok = r < p.na;
        uint32_t full_parity[2] = {0, 0};
        float sp = 0.f;
        uint32_t v[64];
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t & 1;
            const long long ct = g + (long long)t * G;
            const long long c_base = ct * 128 + half * 64;
            mbar_wait(BAR(B_ACCFULL + buf), full_parity[buf], 54);
            full_parity[buf] ^= 1;
            tc_fence_after();
            TMEM_LD32(lane_base + 128 * buf + half * 64, v);
            TMEM_LD32(lane_base + 128 * buf + half * 64 + 32, (v + 32));
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_ACCEMPTY + buf));          // values are in registers: release the buffer
            if (!row_ok) continue;
            const bool interior = (c_base + 64 <= p.nb);
            if (p.sim_out) {
                if (!p.transposed) {
                    float* dst = p.sim_out + r * p.nb + c_base;
                    if (interior && ((uintptr_t)dst & 15) == 0) {
#pragma unroll
                        for (int i = 0; i < 64; i += 4)
                            *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                    } else {
                        for (int i = 0; i < 64; ++i) if (c_base + i < p.nb) dst[i] = __uint_as_float(v[i]);
                    }
                } else {
                    for (int i = 0; i < 64; ++i) if (c_base + i < p.nb) p.sim_out[(c_base + i) * p.na + r] = __uint_as_float(v[i]);
                }
            }
            if (p.sigmoid_sum || p.cta_part) {        // softplus(-z) = max(-z, 0) + ln2 * log2(1 + 2^(-|z| log2 e)); z log2 e = s * zscale + zbias
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const float zl = fmaf(__uint_as_float(v[i]), p.zscale, p.zbias);
                    const float e = lg2_approx(1.f + ex2_approx(-fabsf(zl)));
                    const float term = fmaxf(-zl, 0.f) + e;              // in units of log2
                    acc += (interior || c_base + i < p.nb) ? term : 0.f;
                }
                sp += acc;
            }
            if (p.row_part) {
                float m = -INFINITY;
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (interior || c_base + i < p.nb) m = fmaxf(m, __uint_as_float(v[i]) * p.lse_scale);
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (interior || c_base + i < p.nb) s += ex2_approx((__uint_as_float(v[i]) * p.lse_scale - m) * LOG2E);
                p.row_part[(ct * 2 + half) * p.na + r] = make_float2(m, s);
            }
            if (p.diag_out) {
                const long long dc = r + p.diag_offset;
                if (dc >= c_base && dc < c_base + 64 && dc < p.nb) {
                    float d = 0.f;
#pragma unroll
                    for (int i = 0; i < 64; ++i) d = (c_base + i == dc) ? __uint_as_float(v[i]) : d;
                    p.diag_out[r] = d * p.lse_scale;
                }
            }
        }
        if (p.sigmoid_sum || p.cta_part) {
            double d = (double)sp * 0.6931471805599453;                  // log2 units -> natural log
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            if (lane == 0) {
                if (p.cta_part) reinterpret_cast<double*>(smem + OFF_BAR + 160)[warp] = d;
                else atomicAdd(p.sigmoid_sum, d);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0 && p.cta_part) {
        const double* wp = reinterpret_cast<const double*>(smem + OFF_BAR + 160);
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += wp[w];                              // fixed order
        const unsigned int n_cta = gridDim.x * gridDim.y;
        p.cta_part[blockIdx.y * gridDim.x + blockIdx.x] = s;
        __threadfence();
        if (atomicAdd(p.ticket, 1u) == n_cta - 1) {                          // last CTA of the launch
            __threadfence();
            double tot = 0.0;
            for (unsigned int i = 0; i < n_cta; ++i) tot += *reinterpret_cast<volatile double*>(p.cta_part + i);
            tot *= p.out_scale;
            for (int i = 0; i < p.n_dst; ++i) *reinterpret_cast<volatile double*>(p.sum_dst[i]) = tot;
            *p.ticket = 0u;                                                  // re-armed for the next launch (graph replays)
            __threadfence_system();                                          // a destination may be peer-GPU memory
        }
    }
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
    }
    if (tid == 0) trace_end(TRACE_SIM, trace_t0);
}

}  // namespace simtc

// Workspace layout behind the fp32 path's partials: [a images][b images][row partials (2*ct, na)][col partials (2*rt, nb)]
size_t similarity_tc_work_bytes(long long na, long long nb, int dim) {
    const long long rt = (na + 127) / 128, ct = (nb + 127) / 128;
    return 1024 + (size_t)(rt + ct) * 128 * dim * 2 + (size_t)(2 * ct * na + 2 * rt * nb) * sizeof(float2);
}

int launch_lse_merge(const float2* part, long long n, int parts, float* lse, cudaStream_t st);     // similarity.cu

int launch_similarity_tc(const float* a, const float* b, long long na, long long nb, int dim, long long diag_offset,
                         float* sim_out, float sig_scale, float sig_bias, double* sigmoid_sum_out, float lse_scale,
                         float* row_lse_out, float* col_lse_out, float* diag_out, void* work, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(simtc::similarity_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, simtc::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    const long long rt = (na + 127) / 128, ct = (nb + 127) / 128;
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)work + 1023) & ~(uintptr_t)1023);
    uint8_t* a_img = base;
    uint8_t* b_img = a_img + (size_t)rt * 128 * dim * 2;
    float2* row_part = reinterpret_cast<float2*>(b_img + (size_t)ct * 128 * dim * 2);
    float2* col_part = row_part + 2 * ct * na;
    const int pieces = dim / 8;
    simtc::to_chunk_images_kernel<<<(unsigned)((rt * 128 * pieces + 255) / 256), 256, 0, st>>>(a, na, dim, a_img);
    CMHAR_LAUNCH_CHECK();
    simtc::to_chunk_images_kernel<<<(unsigned)((ct * 128 * pieces + 255) / 256), 256, 0, st>>>(b, nb, dim, b_img);
    CMHAR_LAUNCH_CHECK();
    const float LOG2E_F = 1.4426950408889634f;
    auto pass = [&](bool transposed) -> int {
        simtc::Args p{};
        p.a_img = transposed ? b_img : a_img; p.b_img = transposed ? a_img : b_img;
        p.na = transposed ? nb : na; p.nb = transposed ? na : nb;
        p.kc = dim / 64;
        p.transposed = transposed ? 1 : 0;
        p.lse_scale = lse_scale;
        if (!transposed) {
            p.diag_offset = diag_offset; p.sim_out = sim_out; p.sigmoid_sum = sigmoid_sum_out; p.diag_out = diag_out;
            p.zscale = sig_scale * LOG2E_F; p.zbias = sig_bias * LOG2E_F;
            p.row_part = row_lse_out ? row_part : nullptr;
        } else {
            p.row_part = col_part;
        }
        const long long r_tiles = (p.na + 127) / 128, c_tiles = (p.nb + 127) / 128;
        long long G = (2LL * sm_count() + r_tiles - 1) / r_tiles;       // ~2 waves of CTAs, each streaming >= 1 column tile
        if (G > c_tiles) G = c_tiles;
        if (G < 1) G = 1;
        CMHAR_REQUIRE(r_tiles <= 65535, "too many row tiles");
        simtc::similarity_tc_kernel<<<dim3((unsigned)G, (unsigned)r_tiles), simtc::NT, simtc::SMEM_BYTES, st>>>(p);
        CMHAR_LAUNCH_CHECK();
        return CMHAR_OK;
    };
    if (sim_out || sigmoid_sum_out || row_lse_out || diag_out) { const int rc = pass(false); if (rc) return rc; }
    if (row_lse_out) { const int rc = launch_lse_merge(row_part, na, (int)(2 * ct), row_lse_out, st); if (rc) return rc; }
    if (col_lse_out) {
        int rc = pass(true);
        if (rc) return rc;
        rc = launch_lse_merge(col_part, nb, (int)(2 * rt), col_lse_out, st);
        if (rc) return rc;
    }
    return CMHAR_OK;
}

// Operands already as bf16 chunk images (what the projection-head kernel writes): no conversion pre-pass.  The B operand
// may be partitioned over `n_parts` buffers of `rows_per_part` rows (the ranks' shards; peer pointers are read over NVLink
// by the same cp.async.bulk ring that feeds the MMAs, so the transfer overlaps the math tile by tile).
size_t similarity_img_work_bytes(long long na, long long nb) {
    const long long rt = (na + 127) / 128, ct = (nb + 127) / 128;
    long long G = (2LL * sm_count() + rt - 1) / rt;
    if (G > ct) G = ct;
    if (G < 1) G = 1;
    return 64 + (size_t)(G * rt) * sizeof(double);
}

int launch_similarity_img(const uint8_t* a_img, long long na, const uint8_t* const* b_parts, int n_parts, long long rows_per_part,
                          long long nb, int dim, long long diag_offset, float sig_scale, float sig_bias, double out_scale,
                          double* const* sum_dst, int n_dst, void* work, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(simtc::similarity_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, simtc::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    simtc::Args p{};
    p.a_img = a_img;
    p.na = na; p.nb = nb; p.kc = dim / 64; p.diag_offset = diag_offset;
    const float LOG2E_F = 1.4426950408889634f;
    p.zscale = sig_scale * LOG2E_F; p.zbias = sig_bias * LOG2E_F;
    p.lse_scale = 1.f;
    if (n_parts == 1) { p.b_img = b_parts[0]; p.tiles_per_part = 0; }
    else { for (int i = 0; i < n_parts; ++i) p.b_parts[i] = b_parts[i]; p.tiles_per_part = rows_per_part / 128; }
    p.ticket = reinterpret_cast<unsigned int*>(work);
    p.cta_part = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(work) + 64);
    p.n_dst = n_dst;
    for (int i = 0; i < n_dst; ++i) p.sum_dst[i] = sum_dst[i];
    p.out_scale = out_scale;
    const long long r_tiles = (na + 127) / 128, c_tiles = (nb + 127) / 128;
    long long G = (2LL * sm_count() + r_tiles - 1) / r_tiles;
    if (G > c_tiles) G = c_tiles;
    if (G < 1) G = 1;
    CMHAR_REQUIRE(r_tiles <= 65535, "too many row tiles");
    simtc::similarity_tc_kernel<<<dim3((unsigned)G, (unsigned)r_tiles), simtc::NT, simtc::SMEM_BYTES, st>>>(p);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
