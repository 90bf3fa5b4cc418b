// imu_encoder_fp32.cu -- fp32 (CUDA-core) fused IMU path: patch embed -> 4x post-LN transformer
// layer -> final LN -> classifier head -> arg-max / MSP / energy / Mahalanobis, one launch.
//
// Restates, for eval mode, reference src/models/models.py:30-50 (PatchEmbedding), :100-132
// (IMUEncoder.forward incl. the positional-encoding truncation, F4), :328-339 (IMUClassifier) and
// the nn.TransformerEncoderLayer post-norm block (torch/nn/modules/transformer.py:946-990).
// This is the 1e-3 ("fp32") contract path; the bf16 tcgen05 path lives in imu_encoder_bf16.cu.
//
// Work decomposition: one CTA owns `wpb` whole windows (wpb*S <= 64 token rows).  The residual
// stream h (64x128) and a 64x512 scratch (qkv | attention output, later the FFN hidden) stay in
// shared memory for all layers; the transposed weights (K-major rows, 3.2 MB fp32 for 4 layers)
// are streamed from L2 with coalesced float4 loads shared by the 8 warps through L1.
#include "common.cuh"

namespace cmhar {

constexpr int ROWS = 64;
constexpr int NT = 256;
constexpr int LDB = 512;

struct FwdArgs {
    const char* enc_blob;      // BlobHeader + 1 KiB-aligned fp32 section (dims are read on the device)
    const char* head_blob;     // nullable
    const char* maha_blob;     // nullable
    const float* x;
    long long n, xstride;
    float* cls_out;
    float* tokens_out;
    float* logits_out;
    long long* pred_out;
    float* msp_out;
    float* energy_out;
    float* maha_out;
    void* cls_img;             // nullable: CLS features also as a bf16 operand image [ceil(n/128)][2][128 x 64] (bf16 path only)
};

// acc[i][j] = sum_k A[warp*8+i][k] * Wt[k][lane*4+j]
template <int K>
__device__ __forceinline__ void gemm_64x128(const float* __restrict__ A, int lda,
                                            const float* __restrict__ Wt, int ldw,
                                            float (&acc)[8][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* a0 = A + (warp * 8) * lda;
    const float* w0 = Wt + lane * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 w[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            w[kk] = __ldg(reinterpret_cast<const float4*>(w0 + (size_t)(k + kk) * ldw));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(a0 + i * lda + k);
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
}

// in-place LayerNorm of the 64 rows of h (warp w owns rows 8w..8w+7, lane owns 4 columns)
__device__ __forceinline__ void layer_norm_rows(float* h, const float* __restrict__ gb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gb + lane * 4));
    const float4 b = __ldg(reinterpret_cast<const float4*>(gb + D + lane * 4));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4* p = reinterpret_cast<float4*>(h + (warp * 8 + i) * D + lane * 4);
        float4 v = *p;
        const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / D);
        v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
        const float var = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) * (1.f / D);
        const float r = 1.f / sqrtf(var + LN_EPS);
        v.x = v.x * r * g.x + b.x; v.y = v.y * r * g.y + b.y;
        v.z = v.z * r * g.z + b.z; v.w = v.w * r * g.w + b.w;
        *p = v;
    }
}

__global__ void __launch_bounds__(NT, 1) imu_forward_fp32_kernel(const FwdArgs a) {
    extern __shared__ __align__(16) float smem_f32[];
    float* const smem = smem_f32;
    float* h = smem;                    // [64][128]
    float* buf = smem + ROWS * D;       // [64][512]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const BlobHeader* eh = reinterpret_cast<const BlobHeader*>(a.enc_blob);
    if (eh->magic != ENC_MAGIC) {
        if (tid == 0 && blockIdx.x == 0) printf("cmhar: encoder blob not packed (bad magic)\n");
        return;
    }
    const int S = eh->a, n_layers = eh->b;
    const int wpb = (ROWS / S < 8) ? ROWS / S : 8;
    const int nrows = wpb * S;
    const long long tiles = (a.n + wpb - 1) / wpb;
    const float* enc = reinterpret_cast<const float*>(a.enc_blob + 1024);

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long w0 = tile * wpb;
        __syncthreads();
        // ---- stage the live samples: buf[r][0..15] = patch (s-1) of window w, zeros for CLS/pad rows
        for (int e = tid; e < ROWS * 4; e += NT) {
            const int r = e >> 2, q = e & 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) {
                const int w = r / S, s = r - w * S;
                if (s > 0 && w0 + w < a.n) {
                    const float* src = a.x + (w0 + w) * a.xstride + (s - 1) * P + q * 4;
                    v.x = __ldg(src); v.y = __ldg(src + 1); v.z = __ldg(src + 2); v.w = __ldg(src + 3);
                }
            }
            *reinterpret_cast<float4*>(buf + r * LDB + q * 4) = v;
        }
        __syncthreads();
        {   // patch embedding + CLS + positional encoding (models.py:37-50,118-123)
            float acc[8][4];
            gemm_64x128<P>(buf, LDB, enc + EncLayout::patch_wt, D, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = warp * 8 + i;
                const int s = (r < nrows) ? (r % S) : 0;
                const float4 tb = __ldg(reinterpret_cast<const float4*>(enc + EncLayout::tok_bias + s * D + lane * 4));
                float4 o = make_float4(acc[i][0] + tb.x, acc[i][1] + tb.y, acc[i][2] + tb.z, acc[i][3] + tb.w);
                if (r >= nrows) o = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(h + r * D + lane * 4) = o;
            }
        }
        __syncthreads();

        for (int l = 0; l < n_layers; ++l) {
            const float* L = enc + EncLayout::layers0 + (size_t)l * EncLayout::layer_floats;
            // ---- qkv = h W_in^T + b_in  (q rows pre-scaled by 1/sqrt(16) at pack time)
#pragma unroll 1
            for (int c0 = 0; c0 < 3 * D; c0 += 128) {
                float acc[8][4];
                gemm_64x128<D>(h, D, L + EncLayout::l_w_in + c0, 3 * D, acc);
                const float4 bb = __ldg(reinterpret_cast<const float4*>(L + EncLayout::l_b_in + c0 + lane * 4));
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4*>(buf + (warp * 8 + i) * LDB + c0 + lane * 4) =
                        make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
            }
            __syncthreads();
            // ---- attention: one thread per (window, head, query row); S x S scores in registers
            for (int it = tid; it < wpb * H * S; it += NT) {
                const int i = it % S, hh = (it / S) % H, w = it / (S * H);
                const int r = w * S + i;
                float q[HD];
#pragma unroll
                for (int d4 = 0; d4 < HD / 4; ++d4) {
                    const float4 t = *reinterpret_cast<const float4*>(buf + r * LDB + hh * HD + d4 * 4);
                    q[d4 * 4] = t.x; q[d4 * 4 + 1] = t.y; q[d4 * 4 + 2] = t.z; q[d4 * 4 + 3] = t.w;
                }
                float sc[CMHAR_MAX_SEQ];
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < CMHAR_MAX_SEQ; ++j) {
                    if (j < S) {
                        const float* kr = buf + (w * S + j) * LDB + D + hh * HD;
                        float d = 0.f;
#pragma unroll
                        for (int d4 = 0; d4 < HD / 4; ++d4) {
                            const float4 t = *reinterpret_cast<const float4*>(kr + d4 * 4);
                            d = fmaf(q[d4 * 4], t.x, d); d = fmaf(q[d4 * 4 + 1], t.y, d);
                            d = fmaf(q[d4 * 4 + 2], t.z, d); d = fmaf(q[d4 * 4 + 3], t.w, d);
                        }
                        sc[j] = d;
                        m = fmaxf(m, d);
                    } else {
                        sc[j] = -INFINITY;
                    }
                }
                float den = 0.f;
#pragma unroll
                for (int j = 0; j < CMHAR_MAX_SEQ; ++j) {
                    sc[j] = (j < S) ? expf(sc[j] - m) : 0.f;
                    den += sc[j];
                }
                const float inv = 1.f / den;
                float o[HD];
#pragma unroll
                for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
                for (int j = 0; j < CMHAR_MAX_SEQ; ++j) {
                    if (j < S) {
                        const float pj = sc[j] * inv;
                        const float* vr = buf + (w * S + j) * LDB + 2 * D + hh * HD;
#pragma unroll
                        for (int d4 = 0; d4 < HD / 4; ++d4) {
                            const float4 t = *reinterpret_cast<const float4*>(vr + d4 * 4);
                            o[d4 * 4] = fmaf(pj, t.x, o[d4 * 4]); o[d4 * 4 + 1] = fmaf(pj, t.y, o[d4 * 4 + 1]);
                            o[d4 * 4 + 2] = fmaf(pj, t.z, o[d4 * 4 + 2]); o[d4 * 4 + 3] = fmaf(pj, t.w, o[d4 * 4 + 3]);
                        }
                    }
                }
#pragma unroll
                for (int d4 = 0; d4 < HD / 4; ++d4)
                    *reinterpret_cast<float4*>(buf + r * LDB + 3 * D + hh * HD + d4 * 4) =
                        make_float4(o[d4 * 4], o[d4 * 4 + 1], o[d4 * 4 + 2], o[d4 * 4 + 3]);
            }
            __syncthreads();
            {   // ---- h = LN1(h + attn W_o^T + b_o)
                float acc[8][4];
                gemm_64x128<D>(buf + 3 * D, LDB, L + EncLayout::l_w_o, D, acc);
                const float4 bb = __ldg(reinterpret_cast<const float4*>(L + EncLayout::l_b_o + lane * 4));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4* p = reinterpret_cast<float4*>(h + (warp * 8 + i) * D + lane * 4);
                    float4 v = *p;
                    v.x += acc[i][0] + bb.x; v.y += acc[i][1] + bb.y; v.z += acc[i][2] + bb.z; v.w += acc[i][3] + bb.w;
                    *p = v;
                }
                __syncwarp();
                layer_norm_rows(h, L + EncLayout::l_ln1);     // each warp normalises the rows it wrote
            }
            __syncthreads();
            // ---- hidden = relu(h W_1^T + b_1)
#pragma unroll 1
            for (int c0 = 0; c0 < FF; c0 += 128) {
                float acc[8][4];
                gemm_64x128<D>(h, D, L + EncLayout::l_w1 + c0, FF, acc);
                const float4 bb = __ldg(reinterpret_cast<const float4*>(L + EncLayout::l_b1 + c0 + lane * 4));
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4*>(buf + (warp * 8 + i) * LDB + c0 + lane * 4) =
                        make_float4(fmaxf(acc[i][0] + bb.x, 0.f), fmaxf(acc[i][1] + bb.y, 0.f),
                                    fmaxf(acc[i][2] + bb.z, 0.f), fmaxf(acc[i][3] + bb.w, 0.f));
            }
            __syncthreads();
            {   // ---- h = LN2(h + hidden W_2^T + b_2)
                float acc[8][4];
                gemm_64x128<FF>(buf, LDB, L + EncLayout::l_w2, D, acc);
                const float4 bb = __ldg(reinterpret_cast<const float4*>(L + EncLayout::l_b2 + lane * 4));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4* p = reinterpret_cast<float4*>(h + (warp * 8 + i) * D + lane * 4);
                    float4 v = *p;
                    v.x += acc[i][0] + bb.x; v.y += acc[i][1] + bb.y; v.z += acc[i][2] + bb.z; v.w += acc[i][3] + bb.w;
                    *p = v;
                }
                __syncwarp();
                layer_norm_rows(h, L + EncLayout::l_ln2);
            }
            __syncthreads();
        }
        // ---- final LayerNorm (models.py:127) and outputs
        layer_norm_rows(h, enc + EncLayout::final_ln);
        __syncthreads();
        if (a.tokens_out) {
            for (int e = tid; e < nrows * (D / 4); e += NT) {
                const int r = e / (D / 4), c4 = e % (D / 4);
                const int w = r / S;
                if (w0 + w < a.n)
                    reinterpret_cast<float4*>(a.tokens_out + ((w0 + w) * S + (r - w * S)) * D)[c4] =
                        reinterpret_cast<const float4*>(h + r * D)[c4];
            }
        }
        if (a.cls_out) {
            for (int e = tid; e < wpb * (D / 4); e += NT) {
                const int w = e / (D / 4), c4 = e % (D / 4);
                if (w0 + w < a.n)
                    reinterpret_cast<float4*>(a.cls_out + (w0 + w) * D)[c4] =
                        reinterpret_cast<const float4*>(h + (w * S) * D)[c4];
            }
        }
        __syncthreads();       // h is rewritten by the next tile
    }
}

int launch_head_forward(const FwdArgs& a, int precision, cudaStream_t stream);     // head.cu

// The classifier head and the OOD scores run as a second launch on the CLS features the encoder just
// wrote (head.cu): on the 8 CLS rows of one tile the head is latency-bound on its 344 KB of weights and
// used to cost 30 % of the fused kernel; batched over 32-row tiles it is 6 %.
int launch_head_after_encoder(const FwdArgs& a, int precision, cudaStream_t stream) {
    if (!a.head_blob && !(a.maha_blob && a.maha_out)) return CMHAR_OK;
    FwdArgs h = a;
    h.x = a.cls_out; h.xstride = D;
    h.cls_out = nullptr; h.tokens_out = nullptr;
    return launch_head_forward(h, precision, stream);
}

int launch_imu_forward_fp32(const FwdArgs& a, cudaStream_t stream) {
    static bool configured[64] = {};
    const size_t smem = (size_t)(ROWS * D + ROWS * LDB) * sizeof(float);
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(imu_forward_fp32_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63] = true;
    }
    const long long tiles = (a.n + 3) / 4;     // >= real tile count (wpb >= 4); surplus CTAs exit
    const int grid = (int)((tiles < (long long)sm_count()) ? tiles : (long long)sm_count());
    imu_forward_fp32_kernel<<<grid, NT, smem, stream>>>(a);
    CMHAR_LAUNCH_CHECK();
    return launch_head_after_encoder(a, CMHAR_FP32, stream);
}

}  // namespace cmhar
