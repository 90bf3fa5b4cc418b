// conv_encoder.cu -- 1-D temporal conv / BatchNorm / ReLU IMU encoder (north-star item 1).
//
// SPEC-DEFINED -- NOT IN THE REFERENCE: the reference's IMU encoder is the PatchTST transformer
// (src/models/models.py:53-132); there is no conv stack to be faithful to.  The block is defined
// in conv_encoder.py / oracle/fusion_spec.py:
//     Conv1d(6 -> 32, k5, s1, p2) BN ReLU -> Conv1d(32 -> 64, k5, s2, p2) BN ReLU -> Conv1d(64 -> 128, k5, s2, p2) BN ReLU
//     -> mean over time -> (B, 128)
// BatchNorm (eval) is folded into the conv weights at pack time.  8.2 MFLOP per 6 000-byte window: fp32-FMA
// bound (AI 1 370 flop/B), NOT HBM-bound as the north star assumes -- no conv stack on a 6 x 250 window is.
//
// One CTA (128 threads) per window, persistent over windows.  The window is staged once in shared memory with a
// 2-sample zero halo on both sides (vectorised, coalesced float4 loads of the contiguous 6 000 bytes), every
// layer reads its input from a halo-padded shared-memory tile and writes the next one; layer 3 is reduced to
// the time mean in registers and never stored.  Inner loop: for one input channel a thread loads the
// (stride*16 + 4)-sample input strip of its 16 outputs into registers once and reuses it for the 5 taps
// (80 FMAs per 36 shared-memory loads); weights are stored [c_in][tap][c_out] so a warp's weight load is one
// contiguous line.
#include "common.cuh"

namespace cmhar {
size_t conv_tc_image_bytes();                                                                     // conv_encoder_tc.cu
int pack_conv_tc(const cmhar_conv_encoder_params* p, void* img, cudaStream_t st);
int launch_conv_encoder_tc(const void* img, const float* x, long long n, int L, long long xstride, float* feat, cudaStream_t st);
namespace convenc {

constexpr int NT = 128;
constexpr int C0 = 6, C1 = 32, C2 = 64, C3 = 128, KW = 5, PAD = 2;
constexpr int TB = 16;                    // outputs per register strip
constexpr int MAX_L = 256;                // window length supported by the shared-memory tiles

// blob (floats after the header): w1 [6][5][32] b1 [32] | w2 [32][5][64] b2 [64] | w3 [64][5][128] b3 [128]
struct Layout {
    static constexpr size_t w1 = 0, b1 = w1 + C0 * KW * C1, w2 = b1 + C1, b2 = w2 + C1 * KW * C2, w3 = b2 + C2,
                            b3 = w3 + C2 * KW * C3, total = b3 + C3;
};
constexpr uint32_t CONV_MAGIC = 0x434d4835u;

__host__ __device__ constexpr int out_len(int L, int stride) { return (L + 2 * PAD - KW) / stride + 1; }

// fold BN into the conv: dst[ci][k][co] = w[co][ci][k] * g[co], bias[co] = (b[co] - mean[co]) * g[co] + beta[co]
__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ bn_w,
                                 const float* __restrict__ bn_b, const float* __restrict__ bn_mean, const float* __restrict__ bn_var,
                                 int cin, int cout, float* __restrict__ dst_w, float* __restrict__ dst_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cin * KW * cout) {
        const int co = i % cout, k = (i / cout) % KW, ci = i / (cout * KW);
        const double g = bn_w ? (double)bn_w[co] / sqrt((double)bn_var[co] + (double)BN_EPS) : 1.0;
        dst_w[i] = (float)((double)w[((size_t)co * cin + ci) * KW + k] * g);
    }
    if (i < cout) {
        const double g = bn_w ? (double)bn_w[i] / sqrt((double)bn_var[i] + (double)BN_EPS) : 1.0;
        const double v = b ? (double)b[i] : 0.0;
        dst_b[i] = (float)(bn_w ? (v - (double)bn_mean[i]) * g + (double)bn_b[i] : v);
    }
}

// in: [CIN][lin_pad] (sample t at index t + PAD, zero halo); out: [COUT][lout_pad] the same way, or (POOL) the
// time mean of relu(.) accumulated into pooled[COUT].
template <int CIN, int COUT, int STRIDE, bool POOL>
__device__ __forceinline__ void conv_layer(const float* __restrict__ in, int lin_pad, int lout, const float* __restrict__ w,
                                           const float* __restrict__ bias, float* __restrict__ out, int lout_pad,
                                           float* __restrict__ pooled) {
    constexpr int GROUPS = NT / COUT;                 // time strips handled in parallel
    constexpr int SPAN = STRIDE * (TB - 1) + KW;      // input samples behind TB outputs
    const int co = threadIdx.x % COUT, grp = threadIdx.x / COUT;
    const float bv = __ldg(bias + co);
    float pool_acc = 0.f;
    for (int t0 = grp * TB; t0 < lout; t0 += GROUPS * TB) {
        float acc[TB];
#pragma unroll
        for (int i = 0; i < TB; ++i) acc[i] = bv;
#pragma unroll 2
        for (int ci = 0; ci < CIN; ++ci) {
            float x[SPAN];
            const float* src = in + ci * lin_pad + t0 * STRIDE;          // index (t*STRIDE - PAD) + PAD
#pragma unroll
            for (int i = 0; i < SPAN; ++i) x[i] = src[i];                // warp-uniform address: shared-memory broadcast
            float wk[KW];
#pragma unroll
            for (int k = 0; k < KW; ++k) wk[k] = __ldg(w + ((size_t)ci * KW + k) * COUT + co);
#pragma unroll
            for (int k = 0; k < KW; ++k)
#pragma unroll
                for (int i = 0; i < TB; ++i) acc[i] = fmaf(x[i * STRIDE + k], wk[k], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < TB; ++i) {
            const float v = fmaxf(acc[i], 0.f);
            if (t0 + i < lout) {
                if (POOL) pool_acc += v;
                else out[co * lout_pad + t0 + i + PAD] = v;
            }
        }
    }
    if (POOL) atomicAdd(pooled + co, pool_acc);       // shared memory, <= GROUPS adds per address
}

__global__ void __launch_bounds__(NT) conv_encoder_kernel(const char* __restrict__ blob, const float* __restrict__ x, long long n,
                                                          int L, long long xstride, float* __restrict__ feat) {
    extern __shared__ __align__(16) float cs[];
    const int L1 = out_len(L, 1), L2 = out_len(L1, 2), L3 = out_len(L2, 2);
    // tiles sized for SPAN overreads: a strip may read up to STRIDE*TB + KW samples past its last valid output
    const int p0 = L + 2 * PAD + 2 * TB + 8, p1 = L1 + 2 * PAD + 2 * TB + 8, p2 = L2 + 2 * PAD + 2 * TB + 8;
    float* xs = cs;                               // [6][p0]
    float* a1 = xs + C0 * p0;                     // [32][p1]
    float* a2 = a1 + C1 * p1;                     // [64][p2]
    float* pooled = a2 + C2 * p2;                 // [128]
    const float* wf = reinterpret_cast<const float*>(blob + sizeof(BlobHeader));
    const int tid = threadIdx.x;
    // halos and tails stay zero for the whole kernel: only the interior is rewritten per window
    for (int i = tid; i < C0 * p0 + C1 * p1 + C2 * p2; i += NT) cs[i] = 0.f;
    __syncthreads();
    for (long long wdx = blockIdx.x; wdx < n; wdx += gridDim.x) {
        const float* src = x + wdx * xstride;
        if (((C0 * L) & 3) == 0 && (xstride & 3) == 0 && ((uintptr_t)x & 15) == 0) {
            // the window's 6*L floats are contiguous: coalesced 16-byte loads, scattered into the padded channel rows
            for (int e = tid; e < C0 * L / 4; e += NT) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src) + e);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = 4 * e + j;
                    xs[(idx / L) * p0 + (idx % L) + PAD] = vv[j];
                }
            }
        } else {
            for (int e = tid; e < C0 * L; e += NT) xs[(e / L) * p0 + (e % L) + PAD] = __ldg(src + e);
        }
        if (tid < C3) pooled[tid] = 0.f;
        __syncthreads();
        conv_layer<C0, C1, 1, false>(xs, p0, L1, wf + Layout::w1, wf + Layout::b1, a1, p1, nullptr);
        __syncthreads();
        conv_layer<C1, C2, 2, false>(a1, p1, L2, wf + Layout::w2, wf + Layout::b2, a2, p2, nullptr);
        __syncthreads();
        conv_layer<C2, C3, 2, true>(a2, p2, L3, wf + Layout::w3, wf + Layout::b3, nullptr, 0, pooled);
        __syncthreads();
        if (tid < C3) feat[wdx * C3 + tid] = pooled[tid] / (float)L3;
        __syncthreads();
    }
}

static size_t smem_bytes(int L) {
    const int L1 = out_len(L, 1), L2 = out_len(L1, 2);
    const int p0 = L + 2 * PAD + 2 * TB + 8, p1 = L1 + 2 * PAD + 2 * TB + 8, p2 = L2 + 2 * PAD + 2 * TB + 8;
    return sizeof(float) * ((size_t)C0 * p0 + (size_t)C1 * p1 + (size_t)C2 * p2 + C3);
}

}  // namespace convenc
}  // namespace cmhar

using namespace cmhar;

extern "C" {

size_t cmhar_conv_encoder_blob_bytes(void) { return tc_section_offset(convenc::Layout::total) + conv_tc_image_bytes(); }

int cmhar_conv_encoder_pack(const cmhar_conv_encoder_params* p, void* blob, cmhar_stream_t s) {
    using namespace convenc;
    CMHAR_REQUIRE(p && blob, "cmhar_conv_encoder_pack: null argument");
    cudaStream_t st = (cudaStream_t)s;
    float* f = reinterpret_cast<float*>(reinterpret_cast<char*>(blob) + sizeof(BlobHeader));
    const int cin[3] = {C0, C1, C2}, cout[3] = {C1, C2, C3};
    const size_t wo[3] = {Layout::w1, Layout::w2, Layout::w3}, bo[3] = {Layout::b1, Layout::b2, Layout::b3};
    for (int l = 0; l < 3; ++l) {
        const cmhar_conv_layer_params& q = p->layer[l];
        CMHAR_REQUIRE(q.weight, "conv layer %d: null weight", l);
        CMHAR_REQUIRE(!q.bn_weight || (q.bn_bias && q.bn_mean && q.bn_var), "conv layer %d: BatchNorm needs weight, bias, mean and var", l);
        const int tot = cin[l] * KW * cout[l];
        pack_conv_kernel<<<(tot + 255) / 256, 256, 0, st>>>(q.weight, q.bias, q.bn_weight, q.bn_bias, q.bn_mean, q.bn_var, cin[l], cout[l],
                                                            f + wo[l], f + bo[l]);
        CMHAR_LAUNCH_CHECK();
    }
    // tensor-core section (bf16 path): weight images with the BatchNorm scale folded in, at the next 1 KiB boundary
    {
        const int rc = pack_conv_tc(p, reinterpret_cast<char*>(blob) + tc_section_offset(Layout::total), st);
        if (rc != CMHAR_OK) return rc;
    }
    BlobHeader h{};
    h.magic = CONV_MAGIC;
    h.has_bf16 = 1;
    write_header_kernel<<<1, 1, 0, st>>>(reinterpret_cast<BlobHeader*>(blob), h);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_conv_encoder_forward(const void* blob, const float* x, int64_t n, int32_t window, int64_t x_window_stride, float* feat_out,
                               cmhar_stream_t s) {
    return cmhar_conv_encoder_forward_ex(blob, x, n, window, x_window_stride, feat_out, CMHAR_FP32, s);
}

int cmhar_conv_encoder_forward_ex(const void* blob, const float* x, int64_t n, int32_t window, int64_t x_window_stride, float* feat_out,
                                  int32_t precision, cmhar_stream_t s) {
    using namespace convenc;
    if (n <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(blob && x && feat_out, "cmhar_conv_encoder_forward: null argument");
    CMHAR_REQUIRE(window >= 16 && window <= MAX_L, "window length %d outside [16, %d]", window, MAX_L);
    CMHAR_REQUIRE(x_window_stride >= (int64_t)C0 * window, "x_window_stride %lld shorter than 6 x window", (long long)x_window_stride);
    CMHAR_REQUIRE(precision == CMHAR_FP32 || precision == CMHAR_BF16, "bad precision %d", precision);
    if (precision == CMHAR_BF16)        // implicit-GEMM tcgen05 kernel (conv_encoder_tc.cu)
        return launch_conv_encoder_tc(reinterpret_cast<const char*>(blob) + tc_section_offset(Layout::total), x, n, window, x_window_stride,
                                      feat_out, (cudaStream_t)s);
    const size_t smem = smem_bytes(window);
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(conv_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_L)));
        configured[dev & 63] = true;
    }
    const long long cap = 3LL * sm_count();
    conv_encoder_kernel<<<(unsigned)(n < cap ? n : cap), convenc::NT, smem, (cudaStream_t)s>>>(reinterpret_cast<const char*>(blob), x, n, window,
                                                                                     x_window_stride, feat_out);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // extern "C"
