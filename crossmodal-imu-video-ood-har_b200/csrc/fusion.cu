// fusion.cu -- cross-attention fusion block (spec row A6: "cross-attention or concat-MLP" of the north star;
// the reference has NO fusion block -- SURVEY.md F3 -- so this is defined in-repo (fusion.py) and checked
// against the plain-PyTorch oracle oracle/fusion_spec.py: self-consistency, not reference parity).
//
//   q  = imu_tokens W_q^T + b_q                (n*S, 128)   cmhar_linear_forward
//   kv = frame_feat [W_k; W_v]^T + [b_k; b_v]  (n*T, 256)   cmhar_linear_forward
//   o  = softmax(q_h k_h^T / sqrt(16)) v_h      8 heads      cross_attention_kernel        (this file)
//   y  = LayerNorm(imu_tokens + o W_o^T + b_o)               cmhar_linear_forward + residual_ln_pool_kernel
//   fused = mean_s y                            (n, 128)     residual_ln_pool_kernel       (this file)
//
// Both kernels are memory/latency-bound (65 K MAC per window); one CTA per window, K/V staged in shared memory.
#include "common.cuh"

namespace cmhar {

constexpr int XA_MAX_T = 32;

// one CTA (128 threads) per window: thread = (head, query)
__global__ void __launch_bounds__(128) cross_attention_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                              long long n, int S, int T, float scale_log2e,
                                                              float* __restrict__ out) {
    __shared__ __align__(16) float kv_s[XA_MAX_T * 2 * D];       // [t][K(128) | V(128)]
    const int tid = threadIdx.x;
    for (long long w = blockIdx.x; w < n; w += gridDim.x) {
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(kv + (size_t)w * T * 2 * D);
        for (int e = tid; e < T * 2 * D / 4; e += 128) reinterpret_cast<float4*>(kv_s)[e] = __ldg(src + e);
        __syncthreads();
        const int h = tid >> 4, s = tid & 15;
        if (s < S) {
            float qv[HD];
            const float4* qp = reinterpret_cast<const float4*>(q + ((size_t)w * S + s) * D + h * HD);
#pragma unroll
            for (int i = 0; i < HD / 4; ++i) {
                const float4 t = __ldg(qp + i);
                qv[4 * i] = t.x * scale_log2e; qv[4 * i + 1] = t.y * scale_log2e;
                qv[4 * i + 2] = t.z * scale_log2e; qv[4 * i + 3] = t.w * scale_log2e;
            }
            float sc[XA_MAX_T];
            float m = -INFINITY;
#pragma unroll
            for (int t = 0; t < XA_MAX_T; ++t) {
                if (t < T) {
                    const float* kr = kv_s + t * 2 * D + h * HD;
                    float d = 0.f;
#pragma unroll
                    for (int i = 0; i < HD; ++i) d = fmaf(qv[i], kr[i], d);
                    sc[t] = d;
                    m = fmaxf(m, d);
                } else {
                    sc[t] = -INFINITY;
                }
            }
            float den = 0.f;
#pragma unroll
            for (int t = 0; t < XA_MAX_T; ++t) { sc[t] = (t < T) ? exp2f(sc[t] - m) : 0.f; den += sc[t]; }
            const float inv = 1.f / den;
            float o[HD];
#pragma unroll
            for (int i = 0; i < HD; ++i) o[i] = 0.f;
#pragma unroll
            for (int t = 0; t < XA_MAX_T; ++t) {
                if (t < T) {
                    const float p = sc[t] * inv;
                    const float* vr = kv_s + t * 2 * D + D + h * HD;
#pragma unroll
                    for (int i = 0; i < HD; ++i) o[i] = fmaf(p, vr[i], o[i]);
                }
            }
            float4* op = reinterpret_cast<float4*>(out + ((size_t)w * S + s) * D + h * HD);
#pragma unroll
            for (int i = 0; i < HD / 4; ++i) op[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
    }
}

// pooled[w] = mean_s LayerNorm(x[w,s] + a[w,s]); one CTA (4 warps) per window, a warp per token row
__global__ void __launch_bounds__(128) residual_ln_pool_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               long long n, int S, float eps, float* __restrict__ pooled) {
    __shared__ __align__(16) float part[4][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane);
    for (long long w = blockIdx.x; w < n; w += gridDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = warp; s < S; s += 4) {
            const size_t row = (size_t)w * S + s;
            float4 v = __ldg(reinterpret_cast<const float4*>(x + row * D) + lane);
            const float4 r = __ldg(reinterpret_cast<const float4*>(a + row * D) + lane);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / D);
            v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
            const float var = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) * (1.f / D);
            const float rs = 1.f / sqrtf(var + eps);
            acc.x += v.x * rs * g.x + b.x; acc.y += v.y * rs * g.y + b.y;
            acc.z += v.z * rs * g.z + b.z; acc.w += v.w * rs * g.w + b.w;
        }
        __syncthreads();
        reinterpret_cast<float4*>(part[warp])[lane] = acc;
        __syncthreads();
        if (warp == 0) {
            float4 t = reinterpret_cast<const float4*>(part[0])[lane];
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                const float4 u = reinterpret_cast<const float4*>(part[k])[lane];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            const float invS = 1.f / (float)S;
            reinterpret_cast<float4*>(pooled + (size_t)w * D)[lane] = make_float4(t.x * invS, t.y * invS, t.z * invS, t.w * invS);
        }
    }
}

}  // namespace cmhar

using namespace cmhar;

extern "C" {

int cmhar_cross_attention(const float* q, const float* kv, int64_t n, int32_t s_len, int32_t t_len, float* out,
                          cmhar_stream_t s) {
    CMHAR_REQUIRE(q && kv && out, "cmhar_cross_attention: null argument");
    CMHAR_REQUIRE(s_len >= 1 && s_len <= 16 && t_len >= 1 && t_len <= XA_MAX_T, "cmhar_cross_attention: need 1..16 queries and 1..%d keys per window (got %d, %d)",
                  XA_MAX_T, s_len, t_len);
    if (n <= 0) return CMHAR_OK;
    const unsigned grid = (unsigned)(n < 16LL * sm_count() ? n : 16LL * sm_count());
    cross_attention_kernel<<<grid, 128, 0, (cudaStream_t)s>>>(q, kv, n, s_len, t_len, 0.25f * 1.4426950408889634f, out);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_residual_ln_pool(const float* x, const float* a, const float* gamma, const float* beta, int64_t n,
                           int32_t s_len, float eps, float* pooled, cmhar_stream_t s) {
    CMHAR_REQUIRE(x && a && gamma && beta && pooled && s_len >= 1, "cmhar_residual_ln_pool: bad argument");
    if (n <= 0) return CMHAR_OK;
    const unsigned grid = (unsigned)(n < 16LL * sm_count() ? n : 16LL * sm_count());
    residual_ln_pool_kernel<<<grid, 128, 0, (cudaStream_t)s>>>(x, a, gamma, beta, n, s_len, eps, pooled);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // extern "C"
