// ood.cu -- OOD scoring stage.  SPEC-DERIVED: the reference has no OOD code (SURVEY.md F2, rows
// A1-A5); definitions are the literature ones restated in oracle/ood_spec.py.
//   A1 MSP      score = -max softmax(logits)
//   A2 energy   score = -T logsumexp(logits / T)
//   A3 fit      per-class counts / sums and sum f f^T (double) -> all-reduced, finalised on host
//   A4 score    min_c || f W - mu_c W ||^2    (W = whitening factor of the tied covariance)
//   A5 AUROC/FPR95 from order-preserving score histograms
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {

// ---- A1/A2 from stored logits, C = 32 (the reference head): TMA-fed ring, one thread per row -----------------
// HBM-bound (128 B in, 16 B out per row).  The group-of-lanes kernel below spends ~480 lane-instructions per row
// on shuffles and redundant bookkeeping and sits at 60 % of the copy bandwidth; here a producer lane streams
// 16 KiB tiles (128 rows) into a 4-stage shared-memory ring with cp.async.bulk (no registers, no per-thread
// address arithmetic for the loads) and each consumer thread owns one row of a tile: 8 conflict-free 16-byte
// shared loads (piece order rotated by row & 7), max / arg-max / sum-exp entirely in registers, three coalesced
// stores.  ~4x fewer instructions per row, and the bytes in flight no longer depend on warp scheduling.
namespace lsring {
using namespace tc;
constexpr int ROWS = 128, CLS = 32, TILE_BYTES = ROWS * CLS * 4, NSTAGE = 4;
constexpr int OFF_BAR = NSTAGE * TILE_BYTES, SMEM_BYTES = OFF_BAR + 64;
constexpr int NT = ROWS + 32;

__global__ void __launch_bounds__(NT) logit_scores_ring_kernel(const float* __restrict__ logits, long long n, float invT, float T,
                                                               long long* __restrict__ pred, float* __restrict__ msp,
                                                               float* __restrict__ energy) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    auto FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(NSTAGE + s); };
    const long long tiles = (n + ROWS - 1) / ROWS;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), ROWS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == ROWS / 32) {
        if (lane == 0) {                                   // producer
            uint32_t stage = 0, parity = 1;
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(EMPTY(stage), parity, 95);
                const long long r0 = tile * ROWS;
                const uint32_t bytes = (uint32_t)((n - r0 < ROWS ? n - r0 : ROWS) * (CLS * 4));
                mbar_expect_tx(FULL(stage), bytes);
                bulk_g2s(sbase + stage * TILE_BYTES, logits + r0 * CLS, bytes, FULL(stage));
                if (++stage == NSTAGE) { stage = 0; parity ^= 1; }
            }
        }
        return;
    }
    uint32_t stage = 0, parity = 0;
    const float4* myrow_base = reinterpret_cast<const float4*>(smem) + tid * (CLS / 4);
    const int rot = tid & 7;
    const float kInvT2 = invT * LOG2E;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(FULL(stage), parity, 96);
        const long long r = tile * ROWS + tid;
        float4 v[8];
        const float4* src = myrow_base + stage * (TILE_BYTES / 16);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = src[j ^ rot];            // v[j] holds classes 4 (j ^ rot) ..
        __syncwarp();
        if (lane == 0) mbar_arrive(EMPTY(stage));                   // the row is in registers: the slot may be refilled
        if (++stage == NSTAGE) { stage = 0; parity ^= 1; }
        if (r >= n) continue;
        float m = -INFINITY;
        int idx = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                               // first maximal index (torch max(1)): ties -> lowest class
            const int c0 = 4 * (j ^ rot);
            const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (x[e] > m || (x[e] == m && c0 + e < idx)) { m = x[e]; idx = c0 + e; }
        }
        float s1 = 0.f, sT = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                               // (v - m) first: exact at the maximum, no cancellation later
            v[j].x -= m; v[j].y -= m; v[j].z -= m; v[j].w -= m;
            s1 += ex2_approx(v[j].x * LOG2E) + ex2_approx(v[j].y * LOG2E) + ex2_approx(v[j].z * LOG2E) + ex2_approx(v[j].w * LOG2E);
        }
        if (invT == 1.f) sT = s1;
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                sT += ex2_approx(v[j].x * kInvT2) + ex2_approx(v[j].y * kInvT2) + ex2_approx(v[j].z * kInvT2) + ex2_approx(v[j].w * kInvT2);
        }
        if (pred) pred[r] = idx;
        if (msp) msp[r] = -1.f / s1;
        if (energy) energy[r] = -(m + T * __logf(sT));
    }
}
}  // namespace lsring

// ---- A1/A2 from stored logits (HBM-bound: 128 B in, 16 B out per row at C = 32) ---------------------------
// Fast path (C % 4 == 0, C <= 128): a group of G = C/4 lanes owns a row, every lane loads one float4, so a warp
// reads 32/G consecutive rows as ONE contiguous 512-byte span; the row reductions are log2(G) xor-shuffles.
template <int G>
__global__ void __launch_bounds__(256) logit_scores_vec_kernel(const float* __restrict__ logits, long long n, float invT,
                                                               float T, long long* __restrict__ pred, float* __restrict__ msp,
                                                               float* __restrict__ energy) {
    constexpr int RPW = 32 / G;                                   // rows per warp per load
    constexpr int U = 4;                                          // independent 512-byte loads in flight per warp
    const int lane = threadIdx.x & 31, sub = lane % G;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp_global * (RPW * U); base < n; base += n_warps * (RPW * U)) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = base + u * RPW + lane / G;
            v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (row < n) {
                const float4* p = reinterpret_cast<const float4*>(logits + row * (4 * G)) + sub;
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = base + u * RPW + lane / G;
            float m = v[u].x;
            int idx = 4 * sub;
            if (v[u].y > m) { m = v[u].y; idx = 4 * sub + 1; }
            if (v[u].z > m) { m = v[u].z; idx = 4 * sub + 2; }
            if (v[u].w > m) { m = v[u].w; idx = 4 * sub + 3; }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {                 // first maximal index (torch max(1))
                const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
                const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
                if (m2 > m || (m2 == m && i2 < idx)) { m = m2; idx = i2; }
            }
            float s1 = __expf(v[u].x - m) + __expf(v[u].y - m) + __expf(v[u].z - m) + __expf(v[u].w - m);
            float sT = (invT == 1.f) ? s1
                                     : __expf((v[u].x - m) * invT) + __expf((v[u].y - m) * invT) + __expf((v[u].z - m) * invT) + __expf((v[u].w - m) * invT);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                sT += __shfl_xor_sync(0xffffffffu, sT, o);
            }
            if (row < n && sub == 0) {
                if (pred) pred[row] = idx;
                if (msp) msp[row] = -1.f / s1;
                if (energy) energy[row] = -(m + T * __logf(sT));
            }
        }
    }
}

// generic path: one warp per row
__global__ void __launch_bounds__(256) logit_scores_kernel(const float* __restrict__ logits, long long n, int C,
                                                           float invT, float T, long long* __restrict__ pred,
                                                           float* __restrict__ msp, float* __restrict__ energy) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* z = logits + row * C;
    float m = -INFINITY;
    int idx = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float v = z[c];
        if (v > m) { m = v; idx = c; }
    }
    const float gm = warp_max(m);
    idx = (m == gm) ? idx : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) idx = min(idx, __shfl_xor_sync(0xffffffffu, idx, o));
    float s1 = 0.f, sT = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float v = z[c];
        s1 += expf(v - gm);
        sT += expf((v - gm) * invT);
    }
    s1 = warp_sum(s1);
    sT = warp_sum(sT);
    if (lane == 0) {
        if (pred) pred[row] = idx;
        if (msp) msp[row] = -1.f / s1;
        if (energy) energy[row] = -(gm + T * logf(sT));
    }
}

// ---- A4 from stored features -----------------------------------------------------------------
// 64 rows per tile: Y = F W (64x128x128 register-tiled fp32 GEMM), then squared distances to the
// whitened class means.
__global__ void __launch_bounds__(256) maha_score_kernel(const char* __restrict__ blob, const float* __restrict__ feat,
                                                         long long n, float* __restrict__ score) {
    constexpr int LDY = D + 4;
    extern __shared__ __align__(16) float ms_smem[];
    float* F = ms_smem;                 // [64][128]
    float* Y = ms_smem + 64 * D;        // [64][132]
    const BlobHeader* mh = reinterpret_cast<const BlobHeader*>(blob);
    const MahaLayout ml{mh->a};
    const float* maha = reinterpret_cast<const float*>(blob + sizeof(BlobHeader));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tiles = (n + 63) / 64;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long r0 = tile * 64;
        __syncthreads();
        for (int e = tid; e < 64 * (D / 4); e += 256) {
            const int r = e / (D / 4), c4 = e % (D / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < n) v = __ldg(reinterpret_cast<const float4*>(feat + (r0 + r) * D) + c4);
            reinterpret_cast<float4*>(F + r * D)[c4] = v;
        }
        __syncthreads();
        {
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            const float* a0 = F + warp * 8 * D;
            const float* w0 = maha + ml.whiten() + lane * 4;
#pragma unroll 2
            for (int k = 0; k < D; k += 4) {
                float4 w[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) w[kk] = __ldg(reinterpret_cast<const float4*>(w0 + (size_t)(k + kk) * D));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 a = *reinterpret_cast<const float4*>(a0 + i * D + k);
                    acc[i][0] = fmaf(a.x, w[0].x, acc[i][0]); acc[i][1] = fmaf(a.x, w[0].y, acc[i][1]);
                    acc[i][2] = fmaf(a.x, w[0].z, acc[i][2]); acc[i][3] = fmaf(a.x, w[0].w, acc[i][3]);
                    acc[i][0] = fmaf(a.y, w[1].x, acc[i][0]); acc[i][1] = fmaf(a.y, w[1].y, acc[i][1]);
                    acc[i][2] = fmaf(a.y, w[1].z, acc[i][2]); acc[i][3] = fmaf(a.y, w[1].w, acc[i][3]);
                    acc[i][0] = fmaf(a.z, w[2].x, acc[i][0]); acc[i][1] = fmaf(a.z, w[2].y, acc[i][1]);
                    acc[i][2] = fmaf(a.z, w[2].z, acc[i][2]); acc[i][3] = fmaf(a.z, w[2].w, acc[i][3]);
                    acc[i][0] = fmaf(a.w, w[3].x, acc[i][0]); acc[i][1] = fmaf(a.w, w[3].y, acc[i][1]);
                    acc[i][2] = fmaf(a.w, w[3].z, acc[i][2]); acc[i][3] = fmaf(a.w, w[3].w, acc[i][3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(Y + (warp * 8 + i) * LDY + lane * 4) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        {   // thread (r = tid/4, q = tid%4) scans classes q, q+4, ...
            const int r = tid >> 2, q = tid & 3;
            float best = INFINITY;
            for (int c = q; c < ml.C; c += 4) {
                const float* mu = maha + ml.mean_w() + (size_t)c * D;
                float d = 0.f;
#pragma unroll 8
                for (int j = 0; j < D; j += 4) {
                    const float4 y = *reinterpret_cast<const float4*>(Y + r * LDY + j);
                    const float4 m4 = __ldg(reinterpret_cast<const float4*>(mu + j));
                    const float dx = y.x - m4.x, dy = y.y - m4.y, dz = y.z - m4.z, dw = y.w - m4.w;
                    d = fmaf(dx, dx, d); d = fmaf(dy, dy, d); d = fmaf(dz, dz, d); d = fmaf(dw, dw, d);
                }
                if (__ldg(maha + ml.valid() + c) > 0.f) best = fminf(best, d);
            }
            best = fminf(best, __shfl_xor_sync(0xffffffffu, best, 1));
            best = fminf(best, __shfl_xor_sync(0xffffffffu, best, 2));
            if (q == 0 && r0 + r < n) score[r0 + r] = best;
        }
    }
}

// ---- A3 sufficient statistics --------------------------------------------------------------
// Persistent CTAs; each keeps a private double-precision 128x128 second-moment accumulator and
// per-class sums in shared memory, feeds them from 32-row fp32 register-tiled partial products,
// and adds them to the global double buffers once at the end (one atomic per element per CTA).
constexpr int MA_ROWS = 32;
__global__ void __launch_bounds__(256, 1) maha_accumulate_kernel(const float* __restrict__ feat,
                                                                 const long long* __restrict__ labels, long long n,
                                                                 int C, double* __restrict__ count,
                                                                 double* __restrict__ sum, double* __restrict__ second) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    double* s2 = reinterpret_cast<double*>(sm_raw);                 // [128][128]
    double* ssum = s2 + D * D;                                      // [C][128]
    double* scnt = ssum + (size_t)C * D;                            // [C]
    float* F = reinterpret_cast<float*>(scnt + ((C + 1) & ~1));     // [MA_ROWS][128], 16-byte aligned
    int* lab = reinterpret_cast<int*>(F + MA_ROWS * D);             // [MA_ROWS]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // thread owns rows ty*8.., cols tx*8..
    for (int e = tid; e < D * D + C * D + C; e += 256) s2[e] = 0.0;
    const long long chunks = (n + MA_ROWS - 1) / MA_ROWS;
    float acc[8][8];
    int pending = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const long long r0 = ch * MA_ROWS;
        __syncthreads();
        for (int e = tid; e < MA_ROWS * (D / 4); e += 256) {
            const int r = e / (D / 4), c4 = e % (D / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < n) {
                const long long y = labels[r0 + r];
                if (y >= 0 && y < C) v = __ldg(reinterpret_cast<const float4*>(feat + (r0 + r) * D) + c4);
            }
            reinterpret_cast<float4*>(F + r * D)[c4] = v;
        }
        if (tid < MA_ROWS) {
            long long y = (r0 + tid < n) ? labels[r0 + tid] : -1;
            lab[tid] = (y >= 0 && y < C) ? (int)y : -1;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < MA_ROWS; ++r) {
            const float4 a0 = *reinterpret_cast<const float4*>(F + r * D + ty * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(F + r * D + ty * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(F + r * D + tx * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(F + r * D + tx * 8 + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        // class sums / counts: thread (half = tid/128, c = tid%128) walks every other row
        {
            const int c = tid & 127;
            for (int r = tid >> 7; r < MA_ROWS; r += 2) {
                const int y = lab[r];
                if (y >= 0) {
                    atomicAdd(&ssum[(size_t)y * D + c], (double)F[r * D + c]);
                    if (c == 0) atomicAdd(&scnt[y], 1.0);
                }
            }
        }
        if (++pending == 16) {        // flush fp32 partials (512 rows) into the double accumulator
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) { s2[(ty * 8 + i) * D + tx * 8 + j] += (double)acc[i][j]; acc[i][j] = 0.f; }
            pending = 0;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s2[(ty * 8 + i) * D + tx * 8 + j] += (double)acc[i][j];
    __syncthreads();
    for (int e = tid; e < D * D; e += 256) if (s2[e] != 0.0) atomicAdd(&second[e], s2[e]);
    for (int e = tid; e < C * D; e += 256) if (ssum[e] != 0.0) atomicAdd(&sum[e], ssum[e]);
    for (int e = tid; e < C; e += 256) if (scnt[e] != 0.0) atomicAdd(&count[e], scnt[e]);
}

// ---- A5 histograms ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t score_key(float v) {      // order-preserving float -> uint32
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256) score_key_range_kernel(const float* __restrict__ s, long long n,
                                                              uint32_t* __restrict__ mm) {
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = s[i];
        if (v != v) continue;                                  // NaN scores are ignored
        const uint32_t k = score_key(v);
        lo = min(lo, k); hi = max(hi, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}

__global__ void __launch_bounds__(256) score_histogram_kernel(const float* __restrict__ s, long long n, uint32_t key_lo,
                                                              int shift, int bins, unsigned long long* __restrict__ hist) {
    // Warp-aggregated: scores concentrate in few bins (MSP lives in [-1, 0], a held-out class in one mode of the Mahalanobis score), and one
    // global atomic per score on a handful of addresses ran at 1 % of the HBM roofline.  The lanes of a warp that hit the same bin are
    // matched (`match.any`) and their leader adds the group's population in ONE atomic.  Every lane takes part in every round (the loop
    // bound is warp-uniform); lanes past the end or holding a NaN carry a private negative key and add nothing.
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long base = first - lane; base < n; base += stride) {
        const long long i = base + lane;
        long long b = -1 - lane;
        if (i < n) {
            const float v = s[i];
            if (v == v) {
                const uint32_t k = score_key(v);
                b = (k < key_lo) ? 0 : (long long)((k - key_lo) >> shift);
                if (b >= bins) b = bins - 1;
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        if (b >= 0 && lane == __ffs(peers) - 1) atomicAdd(&hist[b], (unsigned long long)__popc(peers));
    }
}


// ---- near-tie rows (the "bf16_refined" mode of the classifiers): rows whose top-2 logit margin is below rel_tau * max|logit| are
// the only ones whose arg-max a bounded logit error can change; they are re-run on the fp32 path by the caller.
// work[0] = max |logit| as float bits (atomicMax on the non-negative pattern), work[1] = number of selected rows.
__global__ void near_tie_absmax_kernel(const float* __restrict__ logits, long long total, unsigned int* __restrict__ work) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(logits[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m == m) atomicMax(work, __float_as_uint(m));
}
__global__ void near_tie_select_kernel(const float* __restrict__ logits, long long n, int C, float rel_tau, unsigned int* __restrict__ work,
                                       long long* __restrict__ idx) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float tau = rel_tau * __uint_as_float(work[0]);
    const float* z = logits + r * C;
    float a = -INFINITY, b = -INFINITY;
    bool bad = false;
    for (int c = 0; c < C; ++c) {
        const float v = z[c];
        bad |= !(v == v);
        if (v > a) { b = a; a = v; } else if (v > b) b = v;
    }
    if (bad || a - b < tau) idx[atomicAdd(work + 1, 1u)] = r;
}

}  // namespace cmhar

namespace cmhar {
int launch_head_forward(const FwdArgs& a, int precision, cudaStream_t stream);                                  // head.cu
int launch_maha_fit_tc(const float* feat, const long long* labels, long long n, int C, double* count, double* sum, double* second,
                       cudaStream_t st);                                                                         // maha_fit_tc.cu
int launch_maha_score_tc(const uint8_t* section, const float* feat, long long n, float* score, cudaStream_t st);   // maha_score_tc.cu
}

using namespace cmhar;

extern "C" {

int cmhar_logit_scores(const float* logits, int64_t n, int32_t classes, float temperature, int64_t* pred_out,
                       float* msp_out, float* energy_out, cmhar_stream_t s) {
    if (n <= 0) return CMHAR_OK;                 // empty shard
    CMHAR_REQUIRE(logits && classes >= 1 && temperature > 0.f, "cmhar_logit_scores: bad argument");
    const float invT = 1.f / temperature;
    long long* pred = reinterpret_cast<long long*>(pred_out);
    cudaStream_t st = (cudaStream_t)s;
    const int G = classes / 4;
    const bool vec = (classes % 4 == 0) && G >= 1 && G <= 32 && (G & (G - 1)) == 0 && ((uintptr_t)logits & 15) == 0;
    if (classes == lsring::CLS && ((uintptr_t)logits & 15) == 0) {
        static bool configured[64] = {};
        int dev = 0;
        CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
        if (!configured[dev & 63]) {
            CMHAR_CHECK_CUDA(cudaFuncSetAttribute(lsring::logit_scores_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lsring::SMEM_BYTES));
            configured[dev & 63] = true;
        }
        const long long tiles = (n + lsring::ROWS - 1) / lsring::ROWS;
        const unsigned grid = (unsigned)(tiles < 3LL * sm_count() ? tiles : 3LL * sm_count());      // 3 x 64 KiB rings per SM
        lsring::logit_scores_ring_kernel<<<grid, lsring::NT, lsring::SMEM_BYTES, st>>>(logits, n, invT, temperature, pred, msp_out, energy_out);
    } else if (vec) {
        const long long rows_per_block = 8LL * (32 / G) * 4;
        const long long want = (n + rows_per_block - 1) / rows_per_block;
        const unsigned grid = (unsigned)(want < 16LL * sm_count() ? want : 16LL * sm_count());
        switch (G) {
            case 1: logit_scores_vec_kernel<1><<<grid, 256, 0, st>>>(logits, n, invT, temperature, pred, msp_out, energy_out); break;
            case 2: logit_scores_vec_kernel<2><<<grid, 256, 0, st>>>(logits, n, invT, temperature, pred, msp_out, energy_out); break;
            case 4: logit_scores_vec_kernel<4><<<grid, 256, 0, st>>>(logits, n, invT, temperature, pred, msp_out, energy_out); break;
            case 8: logit_scores_vec_kernel<8><<<grid, 256, 0, st>>>(logits, n, invT, temperature, pred, msp_out, energy_out); break;
            case 16: logit_scores_vec_kernel<16><<<grid, 256, 0, st>>>(logits, n, invT, temperature, pred, msp_out, energy_out); break;
            default: logit_scores_vec_kernel<32><<<grid, 256, 0, st>>>(logits, n, invT, temperature, pred, msp_out, energy_out); break;
        }
    } else {
        logit_scores_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(logits, n, classes, invT, temperature, pred, msp_out, energy_out);
    }
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_maha_score(const void* maha_blob, const float* feat, int64_t n, float* score, int32_t precision,
                     cmhar_stream_t s) {
    if (n <= 0) return CMHAR_OK;                 // empty shard
    CMHAR_REQUIRE(maha_blob && feat && score, "cmhar_maha_score: null argument");
    CMHAR_REQUIRE(precision == CMHAR_FP32 || precision == CMHAR_BF16, "bad precision %d", precision);
    if (precision == CMHAR_BF16) {       // tensor-core kernels (whitening + class-mean products as split-bf16 MMAs) when eligible
        BlobInfo mi{};
        if (((uintptr_t)feat & 15) == 0 && lookup_blob(maha_blob, &mi) && mi.magic == MAHA_MAGIC && mi.has_tc) {
            // streaming kernel: one N=160 GEMM against the resident [W | G] image + per-row reduction
            const uint8_t* sec = reinterpret_cast<const uint8_t*>(maha_blob) + tc_section_offset(MahaLayout{mi.a}.total()) + maha_score_section_offset();
            return launch_maha_score_tc(sec, feat, n, score, (cudaStream_t)s);
        }
        FwdArgs a{};
        a.maha_blob = reinterpret_cast<const char*>(maha_blob);
        a.x = feat; a.n = n; a.xstride = D; a.maha_out = score;
        return launch_head_forward(a, precision, (cudaStream_t)s);
    }
    const long long tiles = (n + 63) / 64;
    const int grid = (int)((tiles < 3LL * sm_count()) ? tiles : 3LL * sm_count());
    const size_t smem = sizeof(float) * (64 * D + 64 * (D + 4));
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(maha_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63] = true;
    }
    maha_score_kernel<<<grid, 256, smem, (cudaStream_t)s>>>(reinterpret_cast<const char*>(maha_blob), feat, n, score);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_maha_accumulate(const float* feat, const int64_t* labels, int64_t n, int32_t classes, double* count,
                          double* sum, double* second, int32_t precision, cmhar_stream_t s) {
    if (n <= 0) return CMHAR_OK;                 // empty shard: the statistics stay as they are
    CMHAR_REQUIRE(feat && labels && count && sum && second, "cmhar_maha_accumulate: null argument");
    CMHAR_REQUIRE(classes >= 1 && classes <= 64, "classes=%d outside [1,64]", classes);
    CMHAR_REQUIRE(precision == CMHAR_FP32 || precision == CMHAR_BF16, "bad precision %d", precision);
    if (precision == CMHAR_BF16 && ((uintptr_t)feat & 15) == 0)       // both reductions as split-bf16 tcgen05 GEMMs over the rows
        return launch_maha_fit_tc(feat, reinterpret_cast<const long long*>(labels), n, classes, count, sum, second, (cudaStream_t)s);
    const size_t smem = sizeof(double) * ((size_t)D * D + (size_t)classes * D + ((classes + 1) & ~1)) + sizeof(float) * MA_ROWS * D + sizeof(int) * MA_ROWS;
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(maha_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured[dev & 63] = true;
    }
    const long long chunks = (n + MA_ROWS - 1) / MA_ROWS;
    const int grid = (int)((chunks < (long long)sm_count()) ? chunks : (long long)sm_count());
    maha_accumulate_kernel<<<grid, 256, smem, (cudaStream_t)s>>>(feat, reinterpret_cast<const long long*>(labels), n, classes, count, sum, second);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_near_tie_rows(const float* logits, int64_t n, int32_t classes, float rel_tau, uint32_t* work2, int64_t* idx_out, cmhar_stream_t s) {
    if (n <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(logits && work2 && idx_out && classes >= 2 && rel_tau >= 0.f, "cmhar_near_tie_rows: bad argument");
    cudaStream_t st = (cudaStream_t)s;
    CMHAR_CHECK_CUDA(cudaMemsetAsync(work2, 0, 2 * sizeof(uint32_t), st));
    const long long total = (long long)n * classes;
    const long long blocks = (total + 255) / 256;
    near_tie_absmax_kernel<<<(unsigned)(blocks < 8LL * sm_count() ? blocks : 8LL * sm_count()), 256, 0, st>>>(logits, total, work2);
    CMHAR_LAUNCH_CHECK();
    near_tie_select_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logits, n, classes, rel_tau, work2, reinterpret_cast<long long*>(idx_out));
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_score_key_range(const float* scores, int64_t n, uint32_t* key_min_max, cmhar_stream_t s) {
    if (n <= 0) return CMHAR_OK;                 // an empty shard (a rank that holds no row of one population) is legal
    CMHAR_REQUIRE(scores && key_min_max, "cmhar_score_key_range: null argument");
    const long long blocks = (n + 255) / 256;
    const int grid = (int)((blocks < 8LL * sm_count()) ? blocks : 8LL * sm_count());
    score_key_range_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(scores, n, key_min_max);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_score_histogram(const float* scores, int64_t n, uint32_t key_lo, int32_t shift, int32_t bins,
                          unsigned long long* hist, cmhar_stream_t s) {
    if (n <= 0) return CMHAR_OK;                 // empty shard: nothing to add
    CMHAR_REQUIRE(scores && hist && bins >= 1 && shift >= 0 && shift < 32, "cmhar_score_histogram: bad argument");
    const long long blocks = (n + 255) / 256;
    const int grid = (int)((blocks < 8LL * sm_count()) ? blocks : 8LL * sm_count());
    score_histogram_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(scores, n, key_lo, shift, bins, hist);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // extern "C"
