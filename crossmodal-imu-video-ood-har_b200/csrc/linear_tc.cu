// linear_tc.cu -- bf16 tcgen05 path of the small dense layers (CMHAR_BF16): video projection, projection heads,
// late-fusion layer (reference src/models/models.py:213,226-234; spec row A6 for the concat variant).
//
// At the benchmark's batch (M = 256) these six layers were the second largest share of the step: the fp32
// CUDA-core kernel needs a k-split over every SM plus a reduce launch to be latency-tolerable (~1 800 SM-us per
// layer).  Here one CTA computes a 128 x 128 output tile with 4 MMAs per 64-wide k chunk:
//   * the activation tile is read as fp32 (optionally as the concatenation [x1 | x2]), converted to bf16 and
//     stored as a K-major SWIZZLE_128B A tile by the 8 staging warps, 4-stage ring;
//   * the weights were packed once into bf16 chunk images [n tile][k chunk] (BN folded), streamed by
//     cp.async.bulk through a second 4-stage ring;
//   * fp32 accumulation in TMEM; the same 8 warps then apply bias (+ ReLU) and write fp32 rows.
// A layer occupies ceil(M/128) * ceil(N/128) SMs for a few microseconds (12 CTAs for the 512 -> 768 video
// projection at M = 256) instead of all of them.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace lintc {

using namespace tc;

constexpr int CHUNK = 16384;
constexpr int NS = 4;                                  // barrier slots (maximum ring depth); the depth in use is a launch parameter
constexpr int SMEM_BYTES = 2 * NS * CHUNK + 256;
constexpr int smem_bytes(int ns) { return 2 * ns * CHUNK + 256; }
enum { B_AFULL = 0, B_AEMPTY = B_AFULL + NS, B_BFULL = B_AEMPTY + NS, B_BEMPTY = B_BFULL + NS, B_ACC = B_BEMPTY + NS, B_COUNT };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier area too small");
constexpr int NT = 8 * 32 + 64;

struct Args {
    const uint8_t* w_img;      // [n tile][k chunk] 16 KiB chunk images
    const float* bias;         // (N)
    const float* x1;           // (n, K1)
    const float* x2;           // (n, K - K1) or null
    int K1, K, N;
    long long n;
    int relu;
    float* y;                  // (n, N)
    int ns;                    // ring depth (2..NS): 2 stages = 64 KiB of shared memory, three CTAs of different layers / batches per SM
    const uint8_t* a_img;      // activations already as bf16 chunk images [row tile][k chunk] (the previous layer's y_img), or null
    uint8_t* y_img;            // also / instead write the output as bf16 chunk images [row tile][N/64] for the next layer, or null
};

// folded fp32 W^T (K, N) row-major -> bf16 chunk images [ceil(N/128)][K/64][128 x 64 SW128], zero padded rows
__global__ void pack_linear_chunks_kernel(const float* __restrict__ wt, int K, int N, uint8_t* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per (n row, 16-byte piece)
    const int pieces = K / 8;
    const int n_pad = (N + 127) / 128 * 128;
    if (i >= (long long)n_pad * pieces) return;
    const int n = (int)(i / pieces), p = (int)(i % pieces), kc = p >> 3, j = p & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (n < N) ? wt[(size_t)(p * 8 + e) * N + n] : 0.f;
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    uint8_t* chunk = dst + ((size_t)(n >> 7) * (K / 64) + kc) * CHUNK;
    *reinterpret_cast<uint4*>(chunk + sw128_off(n & 127, j)) = u;
}

__global__ void __launch_bounds__(NT, 2) linear_tc_kernel(const Args p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long trace_t0 = (tid == 0) ? trace_begin() : 0ull;
    const int ns = p.ns;
    const int OFF_A = 0, OFF_B = ns * CHUNK, OFF_BAR = 2 * ns * CHUNK;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
    const int kc_n = p.K / 64;
    const long long row0 = (long long)blockIdx.x * 128;
    const int nt = blockIdx.y;
    constexpr int MMA_WARP = 8, LOAD_WARP = 9;

    if (tid == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(BAR(B_AFULL + s), p.a_img ? 1 : 8); mbar_init(BAR(B_AEMPTY + s), 1);
            mbar_init(BAR(B_BFULL + s), 1); mbar_init(BAR(B_BEMPTY + s), 1);
        }
        mbar_init(BAR(B_ACC), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == LOAD_WARP) {
        if (lane == 0) {
            uint32_t stage = 0, parity = 1;
            for (int kc = 0; kc < kc_n; ++kc) {
                mbar_wait(BAR(B_BEMPTY + stage), parity, 80);
                mbar_expect_tx(BAR(B_BFULL + stage), CHUNK);
                bulk_g2s(sbase + OFF_B + stage * CHUNK, p.w_img + ((size_t)nt * kc_n + kc) * CHUNK, CHUNK, BAR(B_BFULL + stage));
                if (p.a_img) {      // the activation tile is a ready-made operand image: a plain bulk copy, no staging warps
                    mbar_wait(BAR(B_AEMPTY + stage), parity, 85);
                    mbar_expect_tx(BAR(B_AFULL + stage), CHUNK);
                    bulk_g2s(sbase + OFF_A + stage * CHUNK, p.a_img + ((size_t)blockIdx.x * kc_n + kc) * CHUNK, CHUNK, BAR(B_AFULL + stage));
                }
                if (++stage == ns) { stage = 0; parity ^= 1; }
            }
        }
    } else if (warp == MMA_WARP) {
        const bool leader = elect_one();
        constexpr uint32_t ID128 = idesc_bf16(128, 128);
        uint32_t stage = 0, parity = 0;
        for (int kc = 0; kc < kc_n; ++kc) {
            mbar_wait(BAR(B_AFULL + stage), parity, 81);
            mbar_wait(BAR(B_BFULL + stage), parity, 82);
            tc_fence_after();
            const uint64_t dA = sw128_desc(sbase + OFF_A + stage * CHUNK), dB = sw128_desc(sbase + OFF_B + stage * CHUNK);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (leader) umma(tmem, dA + (uint64_t)(2 * k), dB + (uint64_t)(2 * k), ID128, (kc > 0 || k > 0) ? 1u : 0u);
            if (leader) { tc_commit(BAR(B_AEMPTY + stage)); tc_commit(BAR(B_BEMPTY + stage)); }
            if (++stage == ns) { stage = 0; parity ^= 1; }
        }
        if (leader) tc_commit(BAR(B_ACC));
    } else {
        // ---- staging + epilogue warps: thread = (row, 32-column half of a k chunk) / (row, 64 output columns)
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const long long r = row0 + row;
        const bool ok = r < p.n;
        uint32_t stage = 0, parity = 1;
        // the global loads of chunk kc+1 are issued before chunk kc is converted and stored: the L2 round trip of the
        // activation tile is off the per-chunk critical path
        auto fetch = [&](int kc, float4 (&t)[8]) {
            const int k0 = kc * 64 + half * 32;
            const float* src = (k0 < p.K1) ? p.x1 + r * p.K1 + k0 : p.x2 + r * (p.K - p.K1) + (k0 - p.K1);
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = ok ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        float4 cur[8], nxt[8];
        if (!p.a_img) fetch(0, cur);
        for (int kc = 0; kc < (p.a_img ? 0 : kc_n); ++kc) {
            if (kc + 1 < kc_n) fetch(kc + 1, nxt);
            float f[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) { f[4 * i] = cur[i].x; f[4 * i + 1] = cur[i].y; f[4 * i + 2] = cur[i].z; f[4 * i + 3] = cur[i].w; }
            mbar_wait(BAR(B_AEMPTY + stage), parity, 83);
            store_bf16_32(smem + OFF_A + stage * CHUNK, row, half * 4, f);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_AFULL + stage));
            if (++stage == ns) { stage = 0; parity ^= 1; }
#pragma unroll
            for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
        }
        mbar_wait(BAR(B_ACC), 0, 84);
        tc_fence_after();
        uint32_t v[64];
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        TMEM_LD32(lane_base + half * 64, v);
        TMEM_LD32(lane_base + half * 64 + 32, (v + 32));
        tc_wait_ld();
        const int c0 = nt * 128 + half * 64;
        if (ok && p.y) {
            float* dst = p.y + r * p.N + c0;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                if (c0 + i < p.N) {           // N % 4 == 0
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i));
                    float4 o = make_float4(__uint_as_float(v[i]) + b.x, __uint_as_float(v[i + 1]) + b.y,
                                           __uint_as_float(v[i + 2]) + b.z, __uint_as_float(v[i + 3]) + b.w);
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(dst + i) = o;
                }
            }
        }
        if (p.y_img && c0 < p.N) {
            // the same 64 columns as one row of chunk (nt*2 + half) of this row tile's operand image (N % 64 == 0): what the
            // next layer's producer copies straight into its A ring.  Rows past n are written as zeros.
            uint8_t* chunk = p.y_img + ((size_t)blockIdx.x * (p.N / 64) + (c0 >> 6)) * CHUNK;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    o[e] = ok ? __uint_as_float(v[8 * j + e]) + __ldg(p.bias + c0 + 8 * j + e) : 0.f;
                    if (p.relu) o[e] = fmaxf(o[e], 0.f);
                }
                uint4 u;
                u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
                *reinterpret_cast<uint4*>(chunk + sw128_off(row, j)) = u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
    }
    if (tid == 0) trace_end(TRACE_LINEAR, trace_t0);
}

}  // namespace lintc

bool linear_tc_eligible(int in_dim, int out_dim) { return in_dim >= 64 && in_dim % 64 == 0 && out_dim % 4 == 0; }
size_t linear_tc_bytes(int in_dim, int out_dim) { return (size_t)((out_dim + 127) / 128) * (in_dim / 64) * lintc::CHUNK; }

int pack_linear_tc(const float* wt_f32, int in_dim, int out_dim, uint8_t* dst, cudaStream_t st) {
    const long long total = (long long)((out_dim + 127) / 128 * 128) * (in_dim / 8);
    lintc::pack_linear_chunks_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(wt_f32, in_dim, out_dim, dst);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int launch_linear_tc(const uint8_t* w_img, const float* bias, const float* x1, const float* x2, int K1, long long n, int K, int N,
                     int relu, float* y, cudaStream_t st, const uint8_t* a_img = nullptr, uint8_t* y_img = nullptr) {
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(lintc::linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lintc::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    static int ns = 0;
    if (ns == 0) { const char* e = dev_getenv("CMHAR_LINEAR_NS"); ns = e ? atoi(e) : 2; if (ns < 2 || ns > lintc::NS) ns = 2; }      // development switch
    // operand-image input: both rings are plain bulk copies, so a 4-deep ring keeps 128 KiB in flight per CTA and the k loop
    // stops being a chain of serial L2 round trips (K = 512: 16 us -> ~4 us per tile); fp32-row input keeps the 2-deep ring
    // (its staging warps are the bound) so that two CTAs of different layers / batches fit one SM
    const int depth = a_img ? lintc::NS : ns;
    lintc::Args p{w_img, bias, x1, x2, x2 ? K1 : K, K, N, n, relu, y, depth, a_img, y_img};
    const long long mt = (n + 127) / 128;
    CMHAR_REQUIRE(mt <= 0x7fffffffLL, "too many rows");
    lintc::linear_tc_kernel<<<dim3((unsigned)mt, (unsigned)((N + 127) / 128)), lintc::NT, lintc::smem_bytes(depth), st>>>(p);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
