// head.cu -- classifier head + logit / Mahalanobis OOD scores from stored CLS features.
//
// Replaces IMUClassifier.classifier (reference src/models/models.py:312-326,338: [Linear, BatchNorm1d,
// ReLU, Dropout] x 2, Linear; BN folded with its running statistics at pack time) and the arg-max of
// Evaluator.predict (src/eval/evaluator.py:45).  MSP / energy / Mahalanobis are spec rows A1, A2, A4
// (no reference implementation -- parity unpinned, oracle/ood_spec.py).
//
// One CTA owns a tile of R feature rows and runs the whole chain with the activations resident in
// shared memory.  The (K,N) transposed weight matrices stream L2 -> smem as 32-row slabs through a
// 2-deep cp.async.bulk (TMA engine) ring that runs across layer and tile boundaries, so the next
// slab is always in flight while the current one is multiplied:
//   * R = 32 (large batches): 4 row groups x 64 column threads, 8 x {4,2,1} register tiles -- FMA bound;
//   * R = 8  (small batches: more CTAs, each one latency-bound on its 344 KB of weights): 1 row group x
//     256 column threads.
// fp32 FMA arithmetic throughout (the head is 0.5 % of the path's FLOPs; arg-max must not depend on
// tensor-core rounding).  Algorithmic bytes per row: 512 in + 128 logits + 20 (pred, msp, energy, maha).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace headk {

constexpr int NT = 256;
constexpr int SLAB_K = 32;                    // k rows per weight slab
constexpr int SLAB_FLOATS = SLAB_K * 256;     // widest layer: 256 columns
constexpr int MAX_PHASES = 4;                 // 3 head layers + whitening

struct Phase {
    const float* wt;      // (K,N) row-major
    int K, N;
};

struct Plan {
    Phase ph[MAX_PHASES];
    int n_phases;
    int slabs_per_tile;
};

// shared memory (floats): feat [R][128] | hid1 [R][256] | hid2 [R][256] | logit [R][64] | slabs [2][32][256] | barriers
template <int R>
struct Smem {
    static constexpr int feat = 0;
    static constexpr int hid1 = feat + R * D;
    static constexpr int hid2 = hid1 + R * 256;
    static constexpr int logit = hid2 + R * 256;
    static constexpr int slab = logit + R * 64;
    static constexpr int bar = slab + 2 * SLAB_FLOATS;
    static constexpr size_t bytes = (size_t)bar * sizeof(float) + 64;
};

// the slab stream: (phase, slab-in-phase) cursor advancing cyclically over the plan
struct Cursor {
    int phase = 0, slab = 0;
    __device__ __forceinline__ void advance(const Plan& pl) {
        if (++slab * SLAB_K >= pl.ph[phase].K) { slab = 0; if (++phase == pl.n_phases) phase = 0; }
    }
};

template <int RG, int RPT, int CPT>
__device__ __forceinline__ void gemm_phase(const Plan& pl, int phase_idx, const float* __restrict__ in_s, int ldin,
                                           const float* __restrict__ bias, bool relu, float* __restrict__ out_s, int ldout,
                                           float* slabs, uint32_t bar0, long long& consumed, long long total_slabs,
                                           Cursor& issue) {
    constexpr int CT = NT / RG;
    const int tid = threadIdx.x, tx = tid % CT, rg = tid / CT;
    const Phase ph = pl.ph[phase_idx];
    float acc[RPT][CPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[r][j] = 0.f;
    const float* arow = in_s + (size_t)(rg * RPT) * ldin;
    for (int k0 = 0; k0 < ph.K; k0 += SLAB_K) {
        const int buf = (int)(consumed & 1);
        tc::mbar_wait(bar0 + 8u * buf, (uint32_t)((consumed >> 1) & 1), 40);
        const float* ws = slabs + buf * SLAB_FLOATS;
        const int rows = min(SLAB_K, ph.K - k0);
        for (int kk = 0; kk < rows; kk += 4) {
            float4 a4[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) a4[r] = *reinterpret_cast<const float4*>(arow + r * ldin + k0 + kk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float w[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) w[j] = (tx + CT * j < ph.N) ? ws[(kk + q) * ph.N + tx + CT * j] : 0.f;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const float av = q == 0 ? a4[r].x : q == 1 ? a4[r].y : q == 2 ? a4[r].z : a4[r].w;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) acc[r][j] = fmaf(av, w[j], acc[r][j]);
                }
            }
        }
        ++consumed;
        __syncthreads();                                   // everyone is done with this buffer
        if (tid == 0 && consumed + 1 < total_slabs) {      // refill it with the slab two ahead
            const Phase nx = pl.ph[issue.phase];
            const int nrows = min(SLAB_K, nx.K - issue.slab * SLAB_K);
            const uint32_t bytes = (uint32_t)(nrows * nx.N * sizeof(float));
            tc::mbar_expect_tx(bar0 + 8u * buf, bytes);
            tc::bulk_g2s(tc::smem_u32(slabs + buf * SLAB_FLOATS), nx.wt + (size_t)issue.slab * SLAB_K * nx.N, bytes, bar0 + 8u * buf);
            issue.advance(pl);
        }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int col = tx + CT * j;
        if (col < ph.N) {
            const float b = bias ? __ldg(bias + col) : 0.f;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const float v = acc[r][j] + b;
                out_s[(size_t)(rg * RPT + r) * ldout + col] = relu ? fmaxf(v, 0.f) : v;
            }
        }
    }
    __syncthreads();
}

template <int RG, int RPT>
__device__ __forceinline__ void gemm_dispatch(const Plan& pl, int phase_idx, const float* in_s, int ldin, const float* bias,
                                              bool relu, float* out_s, int ldout, float* slabs, uint32_t bar0,
                                              long long& consumed, long long total_slabs, Cursor& issue) {
    constexpr int CT = NT / RG;
    const int cpt = (pl.ph[phase_idx].N + CT - 1) / CT;
    if (cpt <= 1) gemm_phase<RG, RPT, 1>(pl, phase_idx, in_s, ldin, bias, relu, out_s, ldout, slabs, bar0, consumed, total_slabs, issue);
    else if (cpt == 2) gemm_phase<RG, RPT, 2>(pl, phase_idx, in_s, ldin, bias, relu, out_s, ldout, slabs, bar0, consumed, total_slabs, issue);
    else gemm_phase<RG, RPT, 4>(pl, phase_idx, in_s, ldin, bias, relu, out_s, ldout, slabs, bar0, consumed, total_slabs, issue);
}

template <int RG, int RPT>
__global__ void __launch_bounds__(NT) head_scores_kernel(const FwdArgs a) {
    constexpr int R = RG * RPT;
    using L = Smem<R>;
    extern __shared__ __align__(128) float hs[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* head = nullptr;
    const float* maha = nullptr;
    HeadLayout hl{0, 0, 0};
    MahaLayout ml{0};
    Plan pl{};
    if (a.head_blob) {
        const BlobHeader* hh = reinterpret_cast<const BlobHeader*>(a.head_blob);
        hl = HeadLayout{hh->a, hh->b, hh->c};
        head = reinterpret_cast<const float*>(a.head_blob + sizeof(BlobHeader));
        pl.ph[0] = Phase{head + hl.w0(), D, hl.h1};
        pl.ph[1] = Phase{head + hl.w1(), hl.h1, hl.h2};
        pl.ph[2] = Phase{head + hl.w2(), hl.h2, hl.Cp()};
        pl.n_phases = 3;
    }
    const bool do_maha = a.maha_blob && a.maha_out;
    if (do_maha) {
        const BlobHeader* mh = reinterpret_cast<const BlobHeader*>(a.maha_blob);
        ml = MahaLayout{mh->a};
        maha = reinterpret_cast<const float*>(a.maha_blob + sizeof(BlobHeader));
        pl.ph[pl.n_phases++] = Phase{maha + ml.whiten(), D, D};
    }
    for (int p = 0; p < pl.n_phases; ++p) pl.slabs_per_tile += (pl.ph[p].K + SLAB_K - 1) / SLAB_K;

    float* feat = hs + L::feat;
    float* hid1 = hs + L::hid1;
    float* hid2 = hs + L::hid2;
    float* logit = hs + L::logit;
    float* slabs = hs + L::slab;
    const uint32_t bar0 = tc::smem_u32(hs + L::bar);
    const long long tiles = (a.n + R - 1) / R;
    const long long my_tiles = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long total_slabs = my_tiles * pl.slabs_per_tile;
    if (tid == 0) {
        tc::mbar_init(bar0, 1);
        tc::mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Cursor issue;
    if (tid == 0) {                                        // prime the ring: slabs 0 and 1
        for (int b = 0; b < 2 && b < total_slabs; ++b) {
            const Phase nx = pl.ph[issue.phase];
            const int nrows = min(SLAB_K, nx.K - issue.slab * SLAB_K);
            const uint32_t bytes = (uint32_t)(nrows * nx.N * sizeof(float));
            tc::mbar_expect_tx(bar0 + 8u * b, bytes);
            tc::bulk_g2s(tc::smem_u32(slabs + b * SLAB_FLOATS), nx.wt + (size_t)issue.slab * SLAB_K * nx.N, bytes, bar0 + 8u * b);
            issue.advance(pl);
        }
    }
    long long consumed = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long r0 = tile * R;
        for (int e = tid; e < R * (D / 4); e += NT) {
            const long long r = r0 + e / (D / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < a.n) v = __ldg(reinterpret_cast<const float4*>(a.x + r * a.xstride) + (e % (D / 4)));
            reinterpret_cast<float4*>(feat)[e] = v;
        }
        __syncthreads();
        if (head) {
            gemm_dispatch<RG, RPT>(pl, 0, feat, D, head + hl.b0(), true, hid1, 256, slabs, bar0, consumed, total_slabs, issue);
            gemm_dispatch<RG, RPT>(pl, 1, hid1, 256, head + hl.b1(), true, hid2, 256, slabs, bar0, consumed, total_slabs, issue);
            gemm_dispatch<RG, RPT>(pl, 2, hid2, 256, head + hl.b2(), false, logit, 64, slabs, bar0, consumed, total_slabs, issue);
            const int C = hl.C;
            for (int r = warp; r < R; r += NT / 32) {      // one warp scores one row
                const long long gw = r0 + r;
                if (gw >= a.n) break;
                const float z0 = (lane < C) ? logit[r * 64 + lane] : -INFINITY;
                const float z1 = (lane + 32 < C) ? logit[r * 64 + lane + 32] : -INFINITY;
                if (a.logits_out) {
                    if (lane < C) a.logits_out[gw * C + lane] = z0;
                    if (lane + 32 < C) a.logits_out[gw * C + lane + 32] = z1;
                }
                const float m = warp_max(fmaxf(z0, z1));
                int idx = (z0 == m) ? lane : ((z1 == m) ? lane + 32 : 0x7fffffff);      // first arg-max (torch max(1))
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) idx = min(idx, __shfl_xor_sync(0xffffffffu, idx, o));
                const float se = warp_sum(((lane < C) ? expf(z0 - m) : 0.f) + ((lane + 32 < C) ? expf(z1 - m) : 0.f));
                if (lane == 0) {
                    if (a.pred_out) a.pred_out[gw] = idx;
                    if (a.msp_out) a.msp_out[gw] = -1.f / se;
                    if (a.energy_out) a.energy_out[gw] = -(m + logf(se));
                }
            }
        }
        if (do_maha) {
            float* white = hid1;                            // y = feat @ whiten
            gemm_dispatch<RG, RPT>(pl, pl.n_phases - 1, feat, D, nullptr, false, white, 256, slabs, bar0, consumed, total_slabs, issue);
            for (int r = warp; r < R; r += NT / 32) {
                const long long gw = r0 + r;
                if (gw >= a.n) break;
                const float4 y = *reinterpret_cast<const float4*>(white + r * 256 + lane * 4);
                float best = INFINITY;
                for (int c = 0; c < ml.C; ++c) {
                    const float4 mu = __ldg(reinterpret_cast<const float4*>(maha + ml.mean_w() + (size_t)c * D + lane * 4));
                    const float dx = y.x - mu.x, dy = y.y - mu.y, dz = y.z - mu.z, dw = y.w - mu.w;
                    const float dist = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw);
                    if (__ldg(maha + ml.valid() + c) > 0.f) best = fminf(best, dist);
                }
                if (lane == 0) a.maha_out[gw] = best;
            }
        }
        __syncthreads();                                    // feat / logit / white are rewritten by the next tile
    }
}

}  // namespace headk

struct HeadPreLayer;
int launch_head_forward_tc(const FwdArgs& a, const uint8_t* head_tc, const float* head_f32, const HeadLayout& hl,
                           const uint8_t* maha_tc, cudaStream_t stream, const HeadPreLayer* pre);  // head_tc.cu

// x = features (n, xstride >= 128) fp32.
// CMHAR_BF16 and blobs that carry the tensor-core section: the tcgen05 split-bf16 kernel (head_tc.cu).
// Otherwise fp32 CUDA cores -- large batches: 32-row tiles, one persistent CTA per SM; small batches: 8-row tiles
// so that more SMs take part.
int launch_head_forward(const FwdArgs& a, int precision, cudaStream_t stream) {
    using namespace headk;
    if (precision == CMHAR_BF16 && (a.xstride & 3) == 0 && ((uintptr_t)a.x & 15) == 0) {
        BlobInfo hi{}, mi{};
        const bool want_maha = a.maha_blob && a.maha_out;
        const bool head_ok = !a.head_blob || (lookup_blob(a.head_blob, &hi) && hi.magic == HEAD_MAGIC && hi.has_tc);
        const bool maha_ok = !want_maha || (lookup_blob(a.maha_blob, &mi) && mi.magic == MAHA_MAGIC && mi.has_tc);
        if (head_ok && maha_ok && (a.head_blob || want_maha)) {
            HeadLayout hl{hi.a, hi.b, hi.c};
            const uint8_t* htc = a.head_blob ? reinterpret_cast<const uint8_t*>(a.head_blob) + tc_section_offset(hl.total()) : nullptr;
            const float* hf32 = a.head_blob ? reinterpret_cast<const float*>(a.head_blob + sizeof(BlobHeader)) : nullptr;
            const uint8_t* mtc = want_maha ? reinterpret_cast<const uint8_t*>(a.maha_blob) + tc_section_offset(MahaLayout{mi.a}.total()) : nullptr;
            return launch_head_forward_tc(a, htc, hf32, hl, mtc, stream, nullptr);
        }
    }
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(head_scores_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<32>::bytes));
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(head_scores_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<8>::bytes));
        configured[dev & 63] = true;
    }
    const int sms = sm_count();
    if (a.n >= 24LL * sms) {
        const long long tiles = (a.n + 31) / 32;
        const int grid = (int)(tiles < sms ? tiles : sms);
        head_scores_kernel<4, 8><<<grid, headk::NT, Smem<32>::bytes, stream>>>(a);
    } else {
        const long long tiles = (a.n + 7) / 8;
        const int grid = (int)(tiles < 2LL * sms ? tiles : 2LL * sms);
        head_scores_kernel<1, 8><<<grid, headk::NT, Smem<8>::bytes, stream>>>(a);
    }
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // namespace cmhar
