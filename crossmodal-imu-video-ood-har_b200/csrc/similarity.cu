// similarity.cu -- contrastive IMU<->video similarity matrix with fused loss epilogues (fp32 path).
//
// Replaces reference src/models/losses.py:25-54 (SigmoidContrastiveLoss.forward: S = I V^T,
// z = S*t + b, BCE == mean softplus(-z) over ALL pairs) and :67-87
// (InfoNCELoss.forward: symmetric cross-entropy with diagonal targets == row/column logsumexp
// minus the diagonal).  The B x B matrix is never written unless sim_out is given: each 64x64
// tile is reduced in registers to (a) a softplus partial sum, (b) per-row and per-column
// (max, sum-exp) partials that a second tiny kernel merges.
#include "common.cuh"

namespace cmhar {

constexpr int ST = 64;     // tile edge
constexpr int SK = 32;     // k chunk

struct SimArgs {
    const float* a;
    const float* b;
    long long na, nb;
    int dim;
    long long diag_offset;
    float* sim_out;
    float sig_scale, sig_bias;
    double* sigmoid_sum;
    float lse_scale;
    float2* row_part;      // [n_col_tiles][na]  or null
    float2* col_part;      // [n_row_tiles][nb]  or null
    float* diag_out;
};

__device__ __forceinline__ float softplus_neg(float z) {      // log(1 + exp(-z)), stable
    return fmaxf(-z, 0.f) + log1pf(expf(-fabsf(z)));
}

__global__ void __launch_bounds__(256) similarity_fp32_kernel(const SimArgs p) {
    __shared__ __align__(16) float As[SK][ST + 4];
    __shared__ __align__(16) float Bs[SK][ST + 4];
    __shared__ float2 colred[16][ST];
    __shared__ double sred[8];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.y * ST, col0 = (long long)blockIdx.x * ST;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.dim; k0 += SK) {
        __syncthreads();
        for (int e = tid; e < ST * (SK / 4); e += 256) {       // 64 rows x 8 float4
            const int r = e / (SK / 4), k4 = (e % (SK / 4)) * 4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (k0 + k4 < p.dim) {
                if (row0 + r < p.na) va = __ldg(reinterpret_cast<const float4*>(p.a + (row0 + r) * p.dim + k0 + k4));
                if (col0 + r < p.nb) vb = __ldg(reinterpret_cast<const float4*>(p.b + (col0 + r) * p.dim + k0 + k4));
            }
            As[k4][r] = va.x; As[k4 + 1][r] = va.y; As[k4 + 2][r] = va.z; As[k4 + 3][r] = va.w;
            Bs[k4][r] = vb.x; Bs[k4 + 1][r] = vb.y; Bs[k4 + 2][r] = vb.z; Bs[k4 + 3][r] = vb.w;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < SK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
    }

    // ---------------- epilogue on the register tile
    float sp = 0.f;
    float rm[4], rs[4], cm[4], cs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { rm[i] = -INFINITY; rs[i] = 0.f; cm[i] = -INFINITY; cs[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = row0 + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long c = col0 + tx * 4 + j;
            const bool ok = (r < p.na) && (c < p.nb);
            const float s = acc[i][j];
            if (ok) {
                if (p.sim_out) p.sim_out[r * p.nb + c] = s;
                if (p.sigmoid_sum) sp += softplus_neg(fmaf(s, p.sig_scale, p.sig_bias));
                if (p.diag_out && c == r + p.diag_offset) p.diag_out[r] = s * p.lse_scale;
            }
            const float v = ok ? s * p.lse_scale : -INFINITY;
            rm[i] = fmaxf(rm[i], v);
            cm[j] = fmaxf(cm[j], v);
        }
    }
    if (p.row_part || p.col_part) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long r = row0 + ty * 4 + i, c = col0 + tx * 4 + j;
                if (r < p.na && c < p.nb) {
                    const float v = acc[i][j] * p.lse_scale;
                    rs[i] += expf(v - rm[i]);
                    cs[j] += expf(v - cm[j]);
                }
            }
    }
    if (p.row_part) {        // merge across the 16 tx lanes that share a row (contiguous half-warp)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float m = rm[i], s = rs[i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
                const float mn = fmaxf(m, m2);
                s = (mn == -INFINITY) ? 0.f : s * expf(m - mn) + s2 * expf(m2 - mn);
                m = mn;
            }
            const long long r = row0 + ty * 4 + i;
            if (tx == 0 && r < p.na) p.row_part[(long long)blockIdx.x * p.na + r] = make_float2(m, s);
        }
    }
    if (p.col_part) {        // merge across the 16 ty groups through shared memory
#pragma unroll
        for (int j = 0; j < 4; ++j) colred[ty][tx * 4 + j] = make_float2(cm[j], cs[j]);
        __syncthreads();
        if (tid < ST) {
            float m = -INFINITY, s = 0.f;
            for (int g = 0; g < 16; ++g) {
                const float2 q = colred[g][tid];
                const float mn = fmaxf(m, q.x);
                s = (mn == -INFINITY) ? 0.f : s * expf(m - mn) + q.y * expf(q.x - mn);
                m = mn;
            }
            const long long c = col0 + tid;
            if (c < p.nb) p.col_part[(long long)blockIdx.y * p.nb + c] = make_float2(m, s);
        }
    }
    if (p.sigmoid_sum) {
        double d = (double)sp;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if ((tid & 31) == 0) sred[tid >> 5] = d;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sred[w];
            atomicAdd(p.sigmoid_sum, t);
        }
    }
}

// lse[i] = log sum over `parts` partial (max, sumexp) pairs
__global__ void lse_merge_kernel(const float2* __restrict__ part, long long n, int parts, float* __restrict__ lse) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m = -INFINITY, s = 0.f;
    for (int t = 0; t < parts; ++t) {
        const float2 q = part[(long long)t * n + i];
        const float mn = fmaxf(m, q.x);
        s = (mn == -INFINITY) ? 0.f : s * expf(m - mn) + q.y * expf(q.x - mn);
        m = mn;
    }
    lse[i] = m + logf(s);
}

int launch_lse_merge(const float2* part, long long n, int parts, float* lse, cudaStream_t st) {
    lse_merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, n, parts, lse);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

size_t similarity_tc_work_bytes(long long na, long long nb, int dim);                                   // similarity_tc.cu
int launch_similarity_tc(const float* a, const float* b, long long na, long long nb, int dim, long long diag_offset,
                         float* sim_out, float sig_scale, float sig_bias, double* sigmoid_sum_out, float lse_scale,
                         float* row_lse_out, float* col_lse_out, float* diag_out, void* work, cudaStream_t st);

static bool tc_eligible(int64_t na, int64_t nb, int32_t dim) { return dim % 64 == 0 && dim <= 256 && na >= 1 && nb >= 1; }

}  // namespace cmhar

using namespace cmhar;

extern "C" {

size_t cmhar_similarity_work_bytes(int64_t na, int64_t nb, int32_t dim) {
    const int64_t rt = (na + ST - 1) / ST, ct = (nb + ST - 1) / ST;
    const size_t fp32_path = (size_t)(ct * na + rt * nb) * sizeof(float2) + 256;
    const size_t tc_path = tc_eligible(na, nb, dim) ? similarity_tc_work_bytes(na, nb, dim) : 0;
    return fp32_path > tc_path ? fp32_path : tc_path;
}

int cmhar_similarity(const float* a, const float* b, int64_t na, int64_t nb, int32_t dim, int64_t diag_offset,
                     float* sim_out, float sig_scale, float sig_bias, double* sigmoid_sum_out, float lse_scale,
                     float* row_lse_out, float* col_lse_out, float* diag_out, void* work, int32_t precision,
                     cmhar_stream_t s) {
    CMHAR_REQUIRE(a && b && dim > 0 && (dim & 3) == 0, "cmhar_similarity: bad argument (dim %% 4 == 0 required)");
    CMHAR_REQUIRE(precision == CMHAR_FP32 || precision == CMHAR_BF16, "bad precision %d", precision);
    CMHAR_REQUIRE(!(row_lse_out || col_lse_out) || work, "row/col logsumexp outputs need the work buffer");
    if (na <= 0 || nb <= 0) return CMHAR_OK;
    cudaStream_t st = (cudaStream_t)s;
    // bf16 tensor-core path (needs the workspace for the bf16 operand images)
    if (precision == CMHAR_BF16 && work && tc_eligible(na, nb, dim))
        return launch_similarity_tc(a, b, na, nb, dim, diag_offset, sim_out, sig_scale, sig_bias, sigmoid_sum_out, lse_scale,
                                    row_lse_out, col_lse_out, diag_out, work, st);
    const long long rt = (na + ST - 1) / ST, ct = (nb + ST - 1) / ST;
    CMHAR_REQUIRE(rt <= 65535, "too many row tiles");
    SimArgs p{};
    p.a = a; p.b = b; p.na = na; p.nb = nb; p.dim = dim; p.diag_offset = diag_offset;
    p.sim_out = sim_out; p.sig_scale = sig_scale; p.sig_bias = sig_bias; p.sigmoid_sum = sigmoid_sum_out;
    p.lse_scale = lse_scale; p.diag_out = diag_out;
    float2* w = reinterpret_cast<float2*>(work);
    p.row_part = row_lse_out ? w : nullptr;
    p.col_part = col_lse_out ? w + ct * na : nullptr;
    similarity_fp32_kernel<<<dim3((unsigned)ct, (unsigned)rt), 256, 0, st>>>(p);
    CMHAR_LAUNCH_CHECK();
    if (row_lse_out) {
        lse_merge_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(p.row_part, na, (int)ct, row_lse_out);
        CMHAR_LAUNCH_CHECK();
    }
    if (col_lse_out) {
        lse_merge_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(p.col_part, nb, (int)rt, col_lse_out);
        CMHAR_LAUNCH_CHECK();
    }
    return CMHAR_OK;
}

}  // extern "C"
