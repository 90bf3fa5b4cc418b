// maha_finalize.cu -- Mahalanobis fit, finalisation ON THE DEVICE (row A3; BASELINE configs[3]).
// The (all-reduced) sufficient statistics [count C | sum C x 128 | second moment 128 x 128] (fp64) become the scorer state without
// leaving the stream: class means, tied covariance Sigma = (sum f f^T - sum_c n_c mu_c mu_c^T) / N (+ ridge I), Cholesky
// Sigma = G G^T, whitening factor W = G^-T (so that d_c(f) = |f W - mu_c W|^2) and the whitened means mu_c W -- the algebra of
// ood.finalize_mahalanobis / oracle/ood_spec.py (host fp64), which took 0.8 of the 1.13 ms of a configs[3] step (D2H of the statistics,
// numpy Cholesky + LAPACK triangular inverse, H2D of the factors, two stream synchronisations).  Here it is one fp64 CTA,
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
    __shared__ double cnt[MAX_C];
    __shared__ double total_s;
    const int tid = threadIdx.x, row = tid >> 3, l8 = tid & 7;      // covariance phase: 8 lanes per matrix row
    const double* ssum = stats + C;
    const double* second = stats + C + (size_t)C * D;
    if (tid == 0) *fail = 0;
    if (tid < C) { cnt[tid] = stats[tid]; count_f32[tid] = (float)stats[tid]; }
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int c = 0; c < C; ++c) t += cnt[c]; total_s = t; }
    for (int e = tid; e < C * D; e += NT) {
        const double m = ssum[e] / fmax(cnt[e / D], 1.0);
        mean[e] = m;
        fit64[e] = m;
    }
    __syncthreads();
    const double total = total_s;
    if (!(total > 0.0)) {                                       // no labelled rows: nothing to factor
        if (tid == 0) *info = -1;
        return;
    }
    // covariance, symmetrised: lower triangle incl. diagonal into A, the full matrix into fit64.  Row `row`, columns l8, l8 + 8, ...
    double* cov_out = fit64 + (size_t)C * D;
    for (int j = l8; j <= row; j += 8) {
        const int i = row;
        double corr = 0.0;
        for (int c = 0; c < C; ++c) corr += cnt[c] * mean[c * D + i] * mean[c * D + j];
        double v = (0.5 * (second[(size_t)i * D + j] + second[(size_t)j * D + i]) - corr) / total;
        if (i == j) v += ridge;
        A[i * LD + j] = v;
        cov_out[(size_t)i * D + j] = v;
        cov_out[(size_t)j * D + i] = v;
    }
    __syncthreads();
    // Cholesky-Crout, column by column: G[i][k] = (A[i][k] - sum_{p<k} G[i][p] G[k][p]) / G[k][k] for all rows i >= k at once.
    // G's strictly lower triangle stays in A, 1 / G[k][k] in diag[] (reciprocal square root: no fp64 division on the critical path).
    // This phase is bound by the NUMBER of instructions one SM issues (ncu: 0.36 IPC per scheduler, fp64 pipe 12 % busy), so it runs on
    // 256 threads -- two lanes per row, one shuffle per column -- with exact trip counts: 8 lanes per row and predicated fixed-length
    // dot products cost 313 K warp instructions, three times as many.  A warp whose 16 rows lie above column k leaves the loop; the two
    // barriers per column are named barriers counted over the warps still in it.
    if (tid < 256) {
        const int r2 = tid >> 1, h = tid & 1, w8 = tid >> 5;
        const double* rowp = A + r2 * LD;
        for (int k = 0; k < D; ++k) {
            if ((k >> 4) > w8) break;
            const unsigned live = 32u * (unsigned)(8 - (k >> 4));
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (r2 >= k) {
                const double* rk = A + k * LD;
                int p = h;
                for (; p + 6 < k; p += 8) {
                    a0 = fma(rowp[p], rk[p], a0); a1 = fma(rowp[p + 2], rk[p + 2], a1);
                    a2 = fma(rowp[p + 4], rk[p + 4], a2); a3 = fma(rowp[p + 6], rk[p + 6], a3);
                }
                for (; p < k; p += 2) a0 = fma(rowp[p], rk[p], a0);
            }
            double part = (a0 + a1) + (a2 + a3);
            part += __shfl_xor_sync(0xffffffffu, part, 1);      // every lane of the warp takes part (rows above k add zeros)
            double v = 0.0;
            if (r2 >= k) {
                v = rowp[k] - part;
                if (r2 == k && h == 0) {
                    if (!(v > 0.0)) { if (*fail == 0) *fail = k + 1; diag[k] = 1.0; }      // not positive definite: reported, the loop finishes harmlessly
                    else diag[k] = rsqrt(v);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(live) : "memory");
            if (r2 > k && h == 0) A[r2 * LD + k] = v * diag[k];
            asm volatile("bar.sync 1, %0;" ::"r"(live) : "memory");
        }
    }
    __syncthreads();
    // X = G^-1 by recursive doubling, stored transposed (W = X^T = G^-T) in the upper triangle incl. diagonal of A: X[i][j] lives at
    // A[j][i].  With X11 = inv(G11), X22 = inv(G22) of two adjacent b x b diagonal blocks, the block below them is
    // X21 = -X22 (G21 X11): two small products per level, 7 levels -- a dependent chain of ~100 fp64 operations, where forward substitution
    // column by column has 127 steps of ~13 (dependent fp64 operations cost ~50 cycles each here: that route measured 170 K cycles).
    if (tid < D) A[tid * LD + tid] = diag[tid];
    __syncthreads();
    for (int lb = 0; lb < 7; ++lb) {
        const int bsz = 1 << lb, outs = 64 << lb;               // (64 / b) pairs x b^2 elements
        // T = G21 X11 into X21's place: T[r][c] = sum_{q=c}^{b-1} G[o+b+r][o+q] X[o+q][o+c]
        double t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * NT;
            t[u] = 0.0;
            if (e < outs) {
                const int pair = e >> (2 * lb), rem = e & (bsz * bsz - 1), r = rem >> lb, c = rem & (bsz - 1), o = pair << (lb + 1);
                const double* g = A + (o + bsz + r) * LD + o;   // G row, contiguous in q
                const double* x = A + (o + c) * LD + o;         // X column c of X11 = row o + c of the upper triangle, contiguous in q
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int q = c;
                for (; q + 3 < bsz; q += 4) { a0 = fma(g[q], x[q], a0); a1 = fma(g[q + 1], x[q + 1], a1); a2 = fma(g[q + 2], x[q + 2], a2); a3 = fma(g[q + 3], x[q + 3], a3); }
                for (; q < bsz; ++q) a0 = fma(g[q], x[q], a0);
                t[u] = (a0 + a1) + (a2 + a3);
            }
        }
        __syncthreads();                                        // (nothing read above is written below, but keep the phases apart)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * NT;
            if (e < outs) {
                const int pair = e >> (2 * lb), rem = e & (bsz * bsz - 1), r = rem >> lb, c = rem & (bsz - 1), o = pair << (lb + 1);
                A[(o + c) * LD + (o + bsz + r)] = t[u];
            }
        }
        __syncthreads();
        // X21[r][c] = -sum_{q=0}^{r} X[o+b+r][o+b+q] T[q][c]
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * NT;
            t[u] = 0.0;
            if (e < outs) {
                const int pair = e >> (2 * lb), rem = e & (bsz * bsz - 1), r = rem >> lb, c = rem & (bsz - 1), o = pair << (lb + 1);
                const double* x22 = A + (o + bsz) * LD + (o + bsz + r);      // X[o+b+r][o+b+q] = A[(o+b+q)][o+b+r]: stride LD in q
                const double* tt = A + (o + c) * LD + (o + bsz);             // T[q][c], contiguous in q
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int q = 0;
                for (; q + 3 <= r; q += 4) {
                    a0 = fma(x22[q * LD], tt[q], a0); a1 = fma(x22[(q + 1) * LD], tt[q + 1], a1);
                    a2 = fma(x22[(q + 2) * LD], tt[q + 2], a2); a3 = fma(x22[(q + 3) * LD], tt[q + 3], a3);
                }
                for (; q <= r; ++q) a0 = fma(x22[q * LD], tt[q], a0);
                t[u] = -((a0 + a1) + (a2 + a3));
            }
        }
        __syncthreads();                                        // every T has been read before any X21 overwrites it
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * NT;
            if (e < outs) {
                const int pair = e >> (2 * lb), rem = e & (bsz * bsz - 1), r = rem >> lb, c = rem & (bsz - 1), o = pair << (lb + 1);
                A[(o + c) * LD + (o + bsz + r)] = t[u];
            }
        }
        __syncthreads();
    }
    // outputs: whiten[r][c] = W[r][c] = A[r][c] for c >= r (upper triangular), 0 below
    double* whiten_out = fit64 + (size_t)C * D + (size_t)D * D;
    for (int e = tid; e < D * D; e += NT) {
        const int r = e / D, c2 = e - r * D;
        const double w = (c2 >= r) ? A[r * LD + c2] : 0.0;
        whiten_out[e] = w;
        whiten_f32[e] = (float)w;
    }
    double* mw_out = whiten_out + (size_t)D * D;
    for (int e = tid; e < C * D; e += NT) {
        const int c = e / D, j = e - c * D;
        double acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.0;
        const double* mrow = mean + c * D;
        const double* wcol = A + j;
        int k = 0;
        for (; k + 7 <= j; k += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = fma(mrow[k + u], wcol[(k + u) * LD], acc[u]);
        }
        for (; k <= j; ++k) acc[0] = fma(mrow[k], wcol[k * LD], acc[0]);
        const double sum = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
        mw_out[e] = sum;
        mean_w_f32[e] = (float)sum;
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
