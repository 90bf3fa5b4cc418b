// maha_finalize.cu -- Mahalanobis fit, finalisation ON THE DEVICE (row A3; BASELINE configs[3]).
// The (all-reduced) sufficient statistics [count C | sum C x 128 | second moment 128 x 128] (fp64) become the scorer state without
// leaving the stream: class means, tied covariance Sigma = (sum f f^T - sum_c n_c mu_c mu_c^T) / N (+ ridge I), Cholesky
// Sigma = G G^T, whitening factor W = G^-T (so that d_c(f) = |f W - mu_c W|^2) and the whitened means mu_c W -- the algebra of
// ood.finalize_mahalanobis / oracle/ood_spec.py (host fp64), which took 0.8 of the 1.13 ms of a configs[3] step (D2H of the statistics,
// numpy Cholesky + LAPACK triangular inverse, H2D of the factors, two stream synchronisations).  Here it is one fp64 CTA, ~0.1 ms,
// followed on the same stream by cmhar_maha_pack's kernels; the host reads back one status word.
// No reference implementation exists for this stage: parity unpinned, spec oracle = oracle/ood_spec.py.
#include "common.cuh"

namespace cmhar {
namespace mfin {

constexpr int NT = 1024;
constexpr int LD = D + 1;                      // padded row stride (doubles): column accesses spread over the banks
constexpr int MAX_C = 64;                      // cmhar_maha_accumulate's limit

// dynamic shared memory: A[128][LD] (covariance -> strictly lower G, upper incl. diagonal W = G^-T) | mean[C][128] | diag[128]
__host__ __device__ inline size_t smem_bytes(int C) { return sizeof(double) * ((size_t)D * LD + (size_t)C * D + D) + 16; }

// fit64 layout (doubles): mean (C,128) | cov (128,128) | whiten (128,128) | mean_whitened (C,128)
__global__ void __launch_bounds__(NT, 1) maha_finalize_kernel(const double* __restrict__ stats, int C, double ridge, double* __restrict__ fit64,
                                                              float* __restrict__ whiten_f32, float* __restrict__ mean_w_f32,
                                                              float* __restrict__ count_f32, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* A = reinterpret_cast<double*>(smem_raw);
    double* mean = A + (size_t)D * LD;
    double* diag = mean + (size_t)C * D;
    int* fail = reinterpret_cast<int*>(diag + D);
    const int tid = threadIdx.x;
    const double* count = stats;
    const double* ssum = stats + C;
    const double* second = stats + C + (size_t)C * D;
    if (tid == 0) *fail = 0;
    double total = 0.0;
    for (int c = 0; c < C; ++c) total += count[c];            // every thread: C <= 64 loads from L2
    for (int e = tid; e < C * D; e += NT) {
        const int c = e / D;
        const double m = ssum[e] / fmax(count[c], 1.0);
        mean[e] = m;
        fit64[e] = m;
    }
    if (tid < C) count_f32[tid] = (float)count[tid];
    __syncthreads();
    if (!(total > 0.0)) {                                       // no labelled rows: nothing to factor
        if (tid == 0) *info = -1;
        return;
    }
    // covariance, symmetrised, lower triangle incl. diagonal into A; the full matrix into fit64
    double* cov_out = fit64 + (size_t)C * D;
    for (int e = tid; e < D * D; e += NT) {
        const int i = e / D, j = e - i * D;
        if (j > i) continue;
        double s = 0.5 * (second[(size_t)i * D + j] + second[(size_t)j * D + i]);
        double corr = 0.0;
        for (int c = 0; c < C; ++c) corr += count[c] * mean[c * D + i] * mean[c * D + j];
        double v = (s - corr) / total;
        if (i == j) v += ridge;
        A[i * LD + j] = v;
        cov_out[(size_t)i * D + j] = v;
        cov_out[(size_t)j * D + i] = v;
    }
    __syncthreads();
    // right-looking Cholesky on the lower triangle: G below the diagonal, its diagonal in diag[]
    for (int k = 0; k < D; ++k) {
        if (tid == 0) {
            const double d = A[k * LD + k];
            if (!(d > 0.0)) { if (*fail == 0) *fail = k + 1; diag[k] = 1.0; }       // not positive definite: reported, loop finishes harmlessly
            else diag[k] = sqrt(d);
        }
        __syncthreads();
        const double dk = diag[k];
        for (int i = k + 1 + tid; i < D; i += NT) A[i * LD + k] /= dk;
        __syncthreads();
        const int m = D - 1 - k;                                // trailing block rows / columns k+1 .. 127
        for (int e = tid; e < m * m; e += NT) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) A[i * LD + j] -= A[i * LD + k] * A[j * LD + k];
        }
        __syncthreads();
    }
    // W = G^-T, upper triangle incl. diagonal of A: thread j solves G x = e_j by forward substitution, x_i = W[j][i] (row j of the upper
    // triangle is this thread's alone; the strictly lower triangle and diag[] are read-only now)
    if (tid < D) {
        const int j = tid;
        A[j * LD + j] = 1.0 / diag[j];
        for (int i = j + 1; i < D; ++i) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = j;
            for (; k + 3 < i; k += 4) {
                s0 += A[i * LD + k] * A[j * LD + k];
                s1 += A[i * LD + k + 1] * A[j * LD + k + 1];
                s2 += A[i * LD + k + 2] * A[j * LD + k + 2];
                s3 += A[i * LD + k + 3] * A[j * LD + k + 3];
            }
            for (; k < i; ++k) s0 += A[i * LD + k] * A[j * LD + k];
            // careful: A[i][k] with k < i is G (lower); A[j][k] with k >= j is this thread's x_k (upper, row j)
            A[j * LD + i] = -((s0 + s1) + (s2 + s3)) / diag[i];
        }
    }
    __syncthreads();
    // outputs: whiten[k][j] = W[k][j] = x^{(j)}_k ... stored above as A[j][i] = x^{(j)}_i, i.e. A[j][i] = (G^-1)[i][j] = W[j][i]
    double* whiten_out = fit64 + (size_t)C * D + (size_t)D * D;
    for (int e = tid; e < D * D; e += NT) {
        const int r = e / D, c2 = e - r * D;
        const double w = (c2 >= r) ? A[r * LD + c2] : 0.0;     // W is upper triangular
        whiten_out[e] = w;
        whiten_f32[e] = (float)w;
    }
    double* mw_out = whiten_out + (size_t)D * D;
    for (int e = tid; e < C * D; e += NT) {
        const int c = e / D, j = e - c * D;
        double s = 0.0;
        for (int k = 0; k <= j; ++k) s += mean[c * D + k] * A[k * LD + j];
        mw_out[e] = s;
        mean_w_f32[e] = (float)s;
    }
    __syncthreads();
    if (tid == 0) *info = *fail;
}

}  // namespace mfin
}  // namespace cmhar

using namespace cmhar;

extern "C" {

size_t cmhar_maha_fit64_doubles(int32_t C) { return (C < 1 || C > mfin::MAX_C) ? 0 : 2 * (size_t)C * D + 2 * (size_t)D * D; }

int cmhar_maha_finalize(const double* stats, int32_t C, double ridge, double* fit64, float* whiten_f32, float* mean_w_f32,
                        float* count_f32, int32_t* info, cmhar_stream_t s) {
    CMHAR_REQUIRE(stats && fit64 && whiten_f32 && mean_w_f32 && count_f32 && info, "cmhar_maha_finalize: null argument");
    CMHAR_REQUIRE(C >= 1 && C <= mfin::MAX_C, "cmhar_maha_finalize: classes=%d outside [1,%d]", C, mfin::MAX_C);
    CMHAR_REQUIRE(ridge >= 0.0, "cmhar_maha_finalize: negative ridge");
    const size_t smem = mfin::smem_bytes(C);
    static size_t configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured[dev & 63] < smem) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(mfin::maha_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mfin::smem_bytes(mfin::MAX_C)));
        configured[dev & 63] = mfin::smem_bytes(mfin::MAX_C);
    }
    mfin::maha_finalize_kernel<<<1, mfin::NT, smem, (cudaStream_t)s>>>(stats, C, ridge, fit64, whiten_f32, mean_w_f32, count_f32, info);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // extern "C"
