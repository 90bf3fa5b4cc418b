// dense.cu -- fp32 dense layers (Linear with folded BatchNorm + optional ReLU), row L2
// normalisation and the video-tail pooling reduction.
//
// Replaces reference src/models/models.py:213 (VideoEncoder.projection), :226-234
// (ProjectionHead: Linear-BN-ReLU-Linear), :288-289 (F.normalize) and :210-215 (spatial average
// pool + temporal mean; both are linear so pool-then-project == project-then-pool).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {

constexpr int LT_ROWS = 32, LT_COLS = 64, LT_K = 32, LT_SPLIT_MAX = 8;

// y (n,N) = act(x (n,K) @ Wt (K,N) + b);  K % 4 == 0, N % 4 == 0.
// 32 x 64 output tile per CTA (128 threads, 4 x 4 register micro-tiles).  The callers are the
// projection heads at batch 256 (M = 256): the problem is latency-bound, not FLOP-bound, so
//   * tiles are small (a 64 x 128 tiling would leave 140 of 148 SMs idle),
//   * the k loop is split over gridDim.z CTAs (each k slice lands in its own partial buffer and a
//     fixed-order reduction adds them: deterministic, no atomics),
//   * the next k slab is prefetched into registers while the current one is multiplied.
// X2 != nullptr: the input row is the concatenation [X (K1 columns) | X2 (K - K1 columns)] (late-fusion
// concat-MLP: the concatenated feature is never materialised).
__global__ void __launch_bounds__(128) linear_fp32_kernel(const float* __restrict__ Wt, const float* __restrict__ bias,
                                                          const float* __restrict__ X, const float* __restrict__ X2, int K1,
                                                          long long n, int K, int N,
                                                          int relu, float* __restrict__ Y, float* __restrict__ partial) {
    __shared__ __align__(16) float As[LT_K][LT_ROWS + 4];     // transposed: As[k][row]
    __shared__ __align__(16) float Ws[LT_K][LT_COLS];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;        // cols tx*4.., rows ty*4..
    const long long row0 = (long long)blockIdx.x * LT_ROWS;
    const int col0 = blockIdx.y * LT_COLS;
    const int slabs = (K + LT_K - 1) / LT_K;
    const int per = (slabs + gridDim.z - 1) / gridDim.z;
    const int s_begin = blockIdx.z * per, s_end = min(slabs, s_begin + per);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // per-thread prefetch registers: 2 float4 of X (32 rows x 8 float4 / 128 thr) and 4 float4 of W
    float4 xa[2], wa[4];
    auto fetch = [&](int slab) {
        const int k0 = slab * LT_K;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = tid + q * 128, r = e >> 3, k4 = (e & 7) * 4;
            xa[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = k0 + k4;
            if (row0 + r < n && k < K)
                xa[q] = (k < K1) ? __ldg(reinterpret_cast<const float4*>(X + (row0 + r) * K1 + k))
                                 : __ldg(reinterpret_cast<const float4*>(X2 + (row0 + r) * (K - K1) + (k - K1)));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 128, k = e >> 4, c4 = (e & 15) * 4;
            wa[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + k < K && col0 + c4 < N) wa[q] = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + k) * N + col0 + c4));
        }
    };
    if (s_begin < s_end) fetch(s_begin);
    for (int slab = s_begin; slab < s_end; ++slab) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = tid + q * 128, r = e >> 3, k4 = (e & 7) * 4;
            As[k4][r] = xa[q].x; As[k4 + 1][r] = xa[q].y; As[k4 + 2][r] = xa[q].z; As[k4 + 3][r] = xa[q].w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 128, k = e >> 4, c4 = (e & 15) * 4;
            *reinterpret_cast<float4*>(&Ws[k][c4]) = wa[q];
        }
        __syncthreads();
        if (slab + 1 < s_end) fetch(slab + 1);               // in flight while this slab is multiplied
#pragma unroll 8
        for (int k = 0; k < LT_K; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], w4[j], acc[i][j]);
        }
    }
    const int col = col0 + tx * 4;
    if (col >= N) return;
    if (gridDim.z > 1) {                                      // k-split: raw partial sums, finished by linear_reduce_kernel
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = row0 + ty * 4 + i;
            if (r < n)
                *reinterpret_cast<float4*>(partial + ((size_t)blockIdx.z * n + r) * N + col) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        return;
    }
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = row0 + ty * 4 + i;
        if (r >= n) continue;
        float4 o = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(Y + r * N + col) = o;
    }
}

// y = act(sum_z partial[z] + b), fixed summation order
__global__ void __launch_bounds__(256) linear_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                                                            long long n, int N, int splits, int relu, float* __restrict__ Y) {
    const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total4 = n * N / 4;
    if (i4 >= total4) return;
    float4 s = __ldg(reinterpret_cast<const float4*>(bias + (i4 * 4) % N));
    for (int z = 0; z < splits; ++z) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(partial + (size_t)z * n * N) + i4);
        s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    if (relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
    reinterpret_cast<float4*>(Y)[i4] = s;
}

// one warp per row: y = x / max(||x||, 1e-12)
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ x, long long n, int dim,
                                                           float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* xr = x + row * dim;
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) { const float v = xr[c]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < dim; c += 32) y[row * dim + c] = xr[c] * inv;
}

// ---- video tail: pooled[b][c] = mean_{t,p} fmap[b][t][c][p] ---------------------------------------
// One thread per (b, c); adjacent threads read adjacent hw-blocks, so a warp's request is one
// contiguous 32*hw*sizeof(T) span per frame (1 KiB for bf16 4x4).  All `frames` loads of a
// thread are independent and unrolled to keep >= 256 B in flight per thread (HBM-bound stage).
// 128-thread CTAs (4 warps x 40 registers = 5 K registers, no shared memory): small enough to become resident on
// an SM whose tensor-core CTA (the encoder: 320 threads x 168 registers, 224 KiB of shared memory) leaves only
// ~9 K registers and 2.8 KiB free, so the HBM-bound pooling of one batch overlaps the encoder of another.
constexpr int POOL_NT = 128;
constexpr int POOL_UNROLL = 4;                // frames per explicit load batch (x up to two 16-byte loads each)
template <typename T, int VEC /* elements per 16-byte vector, 0 = scalar path */, int U = POOL_UNROLL /* frames per load batch */>
__global__ void __launch_bounds__(POOL_NT) video_pool_kernel(const T* __restrict__ fmap, long long n, int frames,
                                                         int channels, int hw, float* __restrict__ pooled, int batched, uint8_t* __restrict__ img,
                                                         uint8_t* __restrict__ fimg = nullptr) {
    // fimg (optional): the per-FRAME spatial means as a bf16 operand image of n * frames rows x channels (row = b * frames + t) --
    // the cross-attention block's frame tokens, from the same single pass over the feature maps as the clip mean
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels) return;
    const unsigned long long trace_t0 = (threadIdx.x == 0) ? trace_begin() : 0ull;
    // read-once stream: evict_first in L2, so the feature maps do not displace the weight images other kernels re-read
    uint64_t stream_policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream_policy));
    const size_t fstride = (size_t)channels * hw;
    for (long long b = blockIdx.y; b < n; b += gridDim.y) {
    const T* base = fmap + ((size_t)b * frames * channels + c) * hw;
    float acc = 0.f;
    if constexpr (VEC > 0) {
        const int nv = hw / VEC;
        auto ld16 = [&](const uint4* q) {
            uint4 u;
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(q), "l"(stream_policy));
            return u;
        };
        auto sum16 = [&](const uint4& u) -> float {
            if (sizeof(T) == 2) {
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
                float sacc = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h2[q]); sacc += f.x + f.y; }
                return sacc;
            } else {
                const float* f = reinterpret_cast<const float*>(&u);
                return (f[0] + f[1]) + (f[2] + f[3]);
            }
        };
        auto put_frame = [&](int t, float fsum) {
            const long long row = b * frames + t;
            uint8_t* chunk = fimg + ((size_t)(row >> 7) * (channels >> 6) + (c >> 6)) * 16384;
            *reinterpret_cast<__nv_bfloat16*>(chunk + tc::sw128_off((int)(row & 127), (c & 63) >> 3) + (c & 7) * 2) = __float2bfloat16_rn(fsum / (float)hw);
        };
        if (nv <= 2 && batched) {
            // explicit batches: the 8 loads of 4 frames are ALL issued before the first is consumed (the compiler otherwise
            // pairs every volatile load with its use and keeps ~2 in flight per thread)
            for (int t0 = 0; t0 < frames; t0 += U) {
                uint4 u[U][2];
#pragma unroll
                for (int i = 0; i < U; ++i) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        u[i][v] = make_uint4(0u, 0u, 0u, 0u);
                        if (t0 + i < frames && v < nv) u[i][v] = ld16(reinterpret_cast<const uint4*>(base + (t0 + i) * fstride) + v);
                    }
                }
                // fence for the scheduler: an empty volatile asm that "rewrites" every loaded register keeps all the adds behind
                // all the loads
#pragma unroll
                for (int i = 0; i < U; ++i) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) asm volatile("" : "+r"(u[i][v].x), "+r"(u[i][v].y), "+r"(u[i][v].z), "+r"(u[i][v].w));
                }
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    float fs = 0.f;
#pragma unroll
                    for (int v = 0; v < 2; ++v) fs += sum16(u[i][v]);
                    acc += fs;
                    if (fimg && t0 + i < frames) put_frame(t0 + i, fs);
                }
            }
        } else {
            for (int t = 0; t < frames; ++t) {
                const uint4* p = reinterpret_cast<const uint4*>(base + t * fstride);
                float fs = 0.f;
                for (int v = 0; v < nv; ++v) fs += sum16(ld16(p + v));
                acc += fs;
                if (fimg) put_frame(t, fs);
            }
        }
    } else {
        for (int t = 0; t < frames; ++t) {
            float fs = 0.f;
            for (int p = 0; p < hw; ++p) {
                if (sizeof(T) == 2) fs += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base + t * fstride)[p]);
                else fs += reinterpret_cast<const float*>(base + t * fstride)[p];
            }
            acc += fs;
            if (fimg) {
                const long long row = b * frames + t;
                uint8_t* chunk = fimg + ((size_t)(row >> 7) * (channels >> 6) + (c >> 6)) * 16384;
                *reinterpret_cast<__nv_bfloat16*>(chunk + tc::sw128_off((int)(row & 127), (c & 63) >> 3) + (c & 7) * 2) = __float2bfloat16_rn(fs / (float)hw);
            }
        }
    }
    const float mean = acc / (float)(frames * hw);
    if (pooled) pooled[b * channels + c] = mean;
    if (img) {      // the same value as one bf16 element of the operand image [ceil(n/128)][channels/64][128 x 64 SW128] the
                    // projection layer's producer copies straight into its A ring (a warp writes 64 contiguous bytes)
        uint8_t* chunk = img + ((size_t)(b >> 7) * (channels >> 6) + (c >> 6)) * 16384;
        *reinterpret_cast<__nv_bfloat16*>(chunk + tc::sw128_off((int)(b & 127), (c & 63) >> 3) + (c & 7) * 2) = __float2bfloat16_rn(mean);
    }
    }
    if (threadIdx.x == 0) trace_end(TRACE_POOL, trace_t0);
}


// ---- video tail, co-resident variant ----------------------------------------------------------------------
// The kernel above needs ~16 resident CTAs per SM to saturate HBM, and a CTA trace of the pipelined step
// (tools/cta_trace.py) shows what that costs: ~1 000 pooling CTAs of several in-flight batches sit on ~70 of
// the 148 SMs waiting for HBM, and the encoder's whole-SM CTAs cannot start there -- the two stages time-share
// the chip and their times ADD.  This variant takes its bytes in flight from 32 KiB of shared memory instead of
// thousands of registers: ONE 128-thread CTA per SM (<= 40 registers per thread, 33 KiB of shared memory) streams
// [128 channels x hw] slabs (4 KiB for bf16 4x4 maps) through an 8-stage cp.async.bulk ring and fits NEXT to an
// encoder CTA (whose weight ring was cut from 6 to 4 stages to make the room: no loss, tools/enc_sweep.py), so
// the HBM-bound stage of one batch overlaps the tensor-bound stage of the others on the same SMs.
namespace poolring {
using namespace tc;
constexpr int CB = 128;                       // channels per slab == threads per CTA
constexpr int RING_BYTES = 32768;
constexpr int MAX_STAGES = 16;
constexpr int STAGE_TARGET = 16384;          // bytes per ring stage (whole frames)
constexpr int OFF_BAR = RING_BYTES, SMEM_BYTES = RING_BYTES + 2 * MAX_STAGES * 8;

template <typename T>
__global__ void __launch_bounds__(CB, 9) video_pool_ring_kernel(const T* __restrict__ fmap, long long n, int frames, int channels,
                                                                 int hw, float* __restrict__ pooled, int stage_target, int cpt,
                                                                 uint8_t* __restrict__ img, int pf_ahead) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned long long trace_t0 = (tid == 0) ? trace_begin() : 0ull;
    const uint32_t sbase = smem_u32(smem), bar0 = sbase + OFF_BAR;
    const int chan_bytes = hw * (int)sizeof(T);               // multiple of 16, <= 64
    const int CBC = CB * cpt;                                 // channels per CTA: thread tid owns channels tid + 128 j, j < cpt <= 4
    const int slab_bytes = CBC * chan_bytes;                  // one frame of this CTA's channels (contiguous in HBM: larger bulk
                                                              // copies stream better, 16 KiB when the CTA takes all 512 channels)
    // a ring stage holds G frames (16 KiB for bf16 4x4 maps): every stage costs one barrier round trip between the
    // consumers and the issuing thread (~0.7 us), so single-frame 4 KiB stages capped a CTA at 6 GB/s
    int G = stage_target / slab_bytes; G = G < 1 ? 1 : (G > frames ? frames : G);
    const int stage_bytes = G * slab_bytes;
    const int nst = RING_BYTES / stage_bytes < MAX_STAGES ? RING_BYTES / stage_bytes : MAX_STAGES;
    const int spu = (frames + G - 1) / G;                     // stages per unit
    auto FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(MAX_STAGES + s); };
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), CB / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int cblocks = channels / CBC;
    const int units = (int)n * cblocks;                        // unit = (clip, channel block) (launcher: fits int)
    const int my_units = ((int)blockIdx.x < units) ? (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total = my_units * spu;
    const uint64_t stream_policy = l2_policy_evict_first();   // read-once: do not displace the weight images in L2
    // producer state (thread 0 only): stage-step p of this CTA's sequence goes to ring stage p % nst
    int p_unit = blockIdx.x, p_f = 0, p_s = 0, p_n = 0;
    uint32_t p_par = 1u;                                      // parity to wait for on EMPTY (fresh barrier: 1 passes)
    auto issue = [&]() {
        mbar_wait(EMPTY(p_s), p_par, 97);
        const int g = (frames - p_f < G) ? frames - p_f : G;
        mbar_expect_tx(FULL(p_s), (uint32_t)(g * slab_bytes));
        const int b = p_unit / cblocks, cb = p_unit - b * cblocks;
        const T* src = fmap + (((size_t)b * frames + p_f) * channels + (size_t)cb * CBC) * hw;
        for (int f = 0; f < g; ++f)
            bulk_g2s_hint(sbase + p_s * stage_bytes + f * slab_bytes, src + (size_t)f * channels * hw, (uint32_t)slab_bytes, FULL(p_s), stream_policy);
        p_f += g;
        if (p_f == frames) { p_f = 0; p_unit += gridDim.x; }
        if (++p_s == nst) { p_s = 0; p_par ^= 1u; }
        ++p_n;
    };
    // L2 prefetch cursor (thread 0 only), pf_ahead stage-steps in front of the producer: the ring's 32 KiB in flight cover ~1 us of L2
    // latency, not the ~3 us of HBM under load -- cp.async.bulk.prefetch.L2 costs neither registers nor shared memory, so the bytes
    // in flight towards HBM are no longer limited by what fits next to an encoder CTA
    int q_unit = blockIdx.x, q_f = 0, q_n = 0;
    auto prefetch = [&](bool fire) {                          // fire = false: advance the cursor only (steps the ring itself holds)
        const int g = (frames - q_f < G) ? frames - q_f : G;
        if (fire) {
            const int b = q_unit / cblocks, cb = q_unit - b * cblocks;
            const T* src = fmap + (((size_t)b * frames + q_f) * channels + (size_t)cb * CBC) * hw;
            if (cblocks == 1) l2_prefetch_bulk(src, (uint32_t)(g * slab_bytes));            // whole frames: contiguous
            else for (int f = 0; f < g; ++f) l2_prefetch_bulk(src + (size_t)f * channels * hw, (uint32_t)slab_bytes);
        }
        q_f += g;
        if (q_f == frames) { q_f = 0; q_unit += gridDim.x; }
        ++q_n;
    };
    if (tid == 0) {
        while (p_n < nst && p_n < total) issue();
        if (pf_ahead > 0)
            while (q_n < nst + pf_ahead && q_n < total) prefetch(q_n >= nst);
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int f0 = 0, s = 0, unit = blockIdx.x;
    uint32_t par = 0u;
    const float inv = 1.f / (float)(frames * hw);
    for (int it = 0; it < total; ++it) {
        mbar_wait(FULL(s), par, 98);
        const int g = (frames - f0 < G) ? frames - f0 : G;
        const uint8_t* base = smem + s * stage_bytes + tid * chan_bytes;
        for (int f = 0; f < g; ++f) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < cpt) {
                    const uint4* src = reinterpret_cast<const uint4*>(base + f * slab_bytes + j * (CB * chan_bytes));
                    for (int v = 0; v < chan_bytes / 16; ++v) {
                        const uint4 u = src[v];
                        if (sizeof(T) == 2) {
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                            for (int q = 0; q < 4; ++q) { const float2 x = __bfloat1622float2(h2[q]); acc[j] += x.x + x.y; }
                        } else {
                            const float* x = reinterpret_cast<const float*>(&u);
                            acc[j] += (x[0] + x[1]) + (x[2] + x[3]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(EMPTY(s));                 // this warp's channels of the stage are in registers
        if (++s == nst) { s = 0; par ^= 1u; }
        if (tid == 0 && p_n < total) {
            if (pf_ahead > 0 && q_n < total) prefetch(true);
            issue();
        }
        f0 += g;
        if (f0 == frames) {
            const int b = unit / cblocks, cb = unit - b * cblocks;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < cpt) {
                    const int c = cb * CBC + j * CB + tid;
                    const float mean = acc[j] * inv;
                    if (pooled) pooled[(size_t)b * channels + c] = mean;
                    if (img) {                                      // bf16 operand image, as video_pool_kernel writes it
                        uint8_t* chunk = img + ((size_t)(b >> 7) * (channels >> 6) + (c >> 6)) * 16384;
                        *reinterpret_cast<__nv_bfloat16*>(chunk + tc::sw128_off((int)(b & 127), (c & 63) >> 3) + (c & 7) * 2) = __float2bfloat16_rn(mean);
                    }
                }
                acc[j] = 0.f;
            }
            f0 = 0; unit += gridDim.x;
        }
    }
    if (tid == 0) trace_end(TRACE_POOL, trace_t0);
}
}  // namespace poolring

}  // namespace cmhar

namespace cmhar {
bool linear_tc_eligible(int in_dim, int out_dim);                                                        // linear_tc.cu
size_t linear_tc_bytes(int in_dim, int out_dim);
int pack_linear_tc(const float* wt_f32, int in_dim, int out_dim, uint8_t* dst, cudaStream_t st);
int launch_linear_tc(const uint8_t* w_img, const float* bias, const float* x1, const float* x2, int K1, long long n, int K, int N,
                     int relu, float* y, cudaStream_t st, const uint8_t* a_img, uint8_t* y_img);
static size_t linear_fp32_floats(int in_dim, int out_dim) { return (size_t)in_dim * out_dim + out_dim; }
}  // namespace cmhar

using namespace cmhar;

extern "C" {

// [header][fp32 W'^T (in,out) | bias] and, when in_dim % 64 == 0, at the next 1 KiB the bf16 chunk images of
// the tensor-core path.
size_t cmhar_linear_blob_bytes(int32_t in_dim, int32_t out_dim) {
    if (in_dim < 4 || out_dim < 4 || (in_dim & 3) || (out_dim & 3)) return 0;
    if (!linear_tc_eligible(in_dim, out_dim)) return sizeof(BlobHeader) + linear_fp32_floats(in_dim, out_dim) * sizeof(float);
    return tc_section_offset(linear_fp32_floats(in_dim, out_dim)) + linear_tc_bytes(in_dim, out_dim);
}

int cmhar_linear_pack(const float* weight, const float* bias, const float* bn_weight, const float* bn_bias,
                      const float* bn_mean, const float* bn_var, int32_t in_dim, int32_t out_dim, void* blob,
                      cmhar_stream_t s) {
    CMHAR_REQUIRE(weight && blob, "cmhar_linear_pack: null argument");
    CMHAR_REQUIRE(cmhar_linear_blob_bytes(in_dim, out_dim) != 0, "linear dims (%d,%d) must be multiples of 4", in_dim, out_dim);
    CMHAR_REQUIRE(!bn_weight || (bn_bias && bn_mean && bn_var), "BatchNorm needs weight, bias, mean and var");
    cudaStream_t st = (cudaStream_t)s;
    float* f = reinterpret_cast<float*>(reinterpret_cast<char*>(blob) + sizeof(BlobHeader));
    const long long tot = (long long)in_dim * out_dim;
    transpose_fold_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(weight, out_dim, in_dim, f, out_dim, 0, 1.f, bn_weight, bn_var);
    CMHAR_LAUNCH_CHECK();
    bias_fold_kernel<<<(out_dim + 255) / 256, 256, 0, st>>>(bias, out_dim, f + tot, 0, 1.f, bn_weight, bn_bias, bn_mean, bn_var);
    CMHAR_LAUNCH_CHECK();
    BlobHeader h{};
    h.magic = LIN_MAGIC; h.a = in_dim; h.b = out_dim;
    if (linear_tc_eligible(in_dim, out_dim)) {
        h.has_bf16 = 1;
        const int rc = pack_linear_tc(f, in_dim, out_dim, reinterpret_cast<uint8_t*>(blob) + tc_section_offset(linear_fp32_floats(in_dim, out_dim)), st);
        if (rc) return rc;
    }
    register_blob(blob, BlobInfo{LIN_MAGIC, in_dim, out_dim, 0, h.has_bf16});
    write_header_kernel<<<1, 1, 0, st>>>(reinterpret_cast<BlobHeader*>(blob), h);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

size_t cmhar_linear_work_bytes(int64_t n, int32_t out_dim) {
    return (size_t)LT_SPLIT_MAX * (size_t)(n > 0 ? n : 0) * (size_t)out_dim * sizeof(float);
}

static int linear_forward_impl(const void* blob, const float* x, const float* x2, int32_t in_dim1, int64_t n, int32_t in_dim,
                               int32_t out_dim, int32_t relu, float* y, void* work, size_t work_bytes, int32_t precision,
                               cmhar_stream_t s) {
    CMHAR_REQUIRE(blob && x && y, "cmhar_linear_forward: null argument");
    CMHAR_REQUIRE(cmhar_linear_blob_bytes(in_dim, out_dim) != 0, "linear dims (%d,%d) must be multiples of 4", in_dim, out_dim);
    CMHAR_REQUIRE(precision == CMHAR_FP32 || precision == CMHAR_BF16, "bad precision %d", precision);
    if (n <= 0) return CMHAR_OK;
    const float* f = reinterpret_cast<const float*>(reinterpret_cast<const char*>(blob) + sizeof(BlobHeader));
    if (precision == CMHAR_BF16) {       // tensor-core tiles when the blob carries the bf16 images
        BlobInfo bi{};
        const int k1 = x2 ? in_dim1 : in_dim;
        if (lookup_blob(blob, &bi) && bi.magic == LIN_MAGIC && bi.has_tc && bi.a == in_dim && bi.b == out_dim && (k1 % 32) == 0 &&
            ((uintptr_t)x & 15) == 0 && (!x2 || ((uintptr_t)x2 & 15) == 0) && ((uintptr_t)y & 15) == 0)
            return launch_linear_tc(reinterpret_cast<const uint8_t*>(blob) + tc_section_offset(linear_fp32_floats(in_dim, out_dim)),
                                    f + (size_t)in_dim * out_dim, x, x2, k1, n, in_dim, out_dim, relu, y, (cudaStream_t)s, nullptr, nullptr);
    }
    const unsigned gx = (unsigned)((n + LT_ROWS - 1) / LT_ROWS), gy = (unsigned)((out_dim + LT_COLS - 1) / LT_COLS);
    // k-split only when the plain grid cannot fill the machine and the caller gave a workspace
    int splits = 1;
    if (work && (long long)gx * gy < sm_count()) {
        const int slabs = (in_dim + LT_K - 1) / LT_K;
        splits = (int)((2LL * sm_count() + (long long)gx * gy - 1) / ((long long)gx * gy));
        if (splits > LT_SPLIT_MAX) splits = LT_SPLIT_MAX;
        if (splits > slabs) splits = slabs;
        if ((size_t)splits * n * out_dim * sizeof(float) > work_bytes) splits = 1;
    }
    linear_fp32_kernel<<<dim3(gx, gy, splits), 128, 0, (cudaStream_t)s>>>(f, f + (size_t)in_dim * out_dim, x, x2, x2 ? in_dim1 : in_dim,
                                                                          n, in_dim, out_dim, relu, y, reinterpret_cast<float*>(work));
    CMHAR_LAUNCH_CHECK();
    if (splits > 1) {
        const long long total4 = n * out_dim / 4;
        linear_reduce_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)s>>>(
            reinterpret_cast<const float*>(work), f + (size_t)in_dim * out_dim, n, out_dim, splits, relu, y);
        CMHAR_LAUNCH_CHECK();
    }
    return CMHAR_OK;
}

int cmhar_linear_forward(const void* blob, const float* x, int64_t n, int32_t in_dim, int32_t out_dim, int32_t relu,
                         float* y, void* work, size_t work_bytes, int32_t precision, cmhar_stream_t s) {
    return linear_forward_impl(blob, x, nullptr, in_dim, n, in_dim, out_dim, relu, y, work, work_bytes, precision, s);
}

int cmhar_concat_linear_forward(const void* blob, const float* x1, int32_t in_dim1, const float* x2, int32_t in_dim2, int64_t n,
                                int32_t out_dim, int32_t relu, float* y, void* work, size_t work_bytes, int32_t precision,
                                cmhar_stream_t s) {
    CMHAR_REQUIRE(x2 && in_dim1 >= 4 && in_dim2 >= 4 && !(in_dim1 & 3) && !(in_dim2 & 3),
                  "cmhar_concat_linear_forward: both inputs need a multiple-of-4 width (%d, %d)", in_dim1, in_dim2);
    return linear_forward_impl(blob, x1, x2, in_dim1, n, in_dim1 + in_dim2, out_dim, relu, y, work, work_bytes, precision, s);
}

size_t cmhar_operand_image_bytes(int64_t n, int32_t dim) {
    if (n <= 0 || dim < 64 || dim % 64) return 0;
    return (size_t)((n + 127) / 128) * (size_t)(dim / 64) * 16384;
}

int cmhar_linear_forward_img(const void* blob, const float* x, const void* x_img, int64_t n, int32_t in_dim, int32_t out_dim,
                             int32_t relu, float* y, void* y_img, cmhar_stream_t s) {
    CMHAR_REQUIRE(blob && (x || x_img) && (y || y_img), "cmhar_linear_forward_img: needs an input and an output");
    CMHAR_REQUIRE(!y_img || out_dim % 64 == 0, "cmhar_linear_forward_img: an output image needs out_dim %% 64 == 0 (got %d)", out_dim);
    if (n <= 0) return CMHAR_OK;
    BlobInfo bi{};
    CMHAR_REQUIRE(lookup_blob(blob, &bi) && bi.magic == LIN_MAGIC && bi.has_tc && bi.a == in_dim && bi.b == out_dim,
                  "cmhar_linear_forward_img: the blob has no tensor-core section for (%d,%d)", in_dim, out_dim);
    CMHAR_REQUIRE(x_img || (in_dim % 32 == 0 && ((uintptr_t)x & 15) == 0), "cmhar_linear_forward_img: misaligned fp32 input");
    CMHAR_REQUIRE((!y || ((uintptr_t)y & 15) == 0) && (!x_img || ((uintptr_t)x_img & 15) == 0) && (!y_img || ((uintptr_t)y_img & 15) == 0),
                  "cmhar_linear_forward_img: misaligned buffer");
    const float* f = reinterpret_cast<const float*>(reinterpret_cast<const char*>(blob) + sizeof(BlobHeader));
    return launch_linear_tc(reinterpret_cast<const uint8_t*>(blob) + tc_section_offset(linear_fp32_floats(in_dim, out_dim)),
                            f + (size_t)in_dim * out_dim, x, nullptr, in_dim, n, in_dim, out_dim, relu, y, (cudaStream_t)s,
                            reinterpret_cast<const uint8_t*>(x_img), reinterpret_cast<uint8_t*>(y_img));
}

int cmhar_l2_normalize(const float* x, int64_t n, int32_t dim, float* y, cmhar_stream_t s) {
    CMHAR_REQUIRE(x && y && dim > 0, "cmhar_l2_normalize: bad argument");
    if (n <= 0) return CMHAR_OK;
    l2_normalize_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)s>>>(x, n, dim, y);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

static bool pool_ring_eligible(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw) {
    const int esize = is_bf16 ? 2 : 4;
    return ((uintptr_t)fmap & 15) == 0 && channels % poolring::CB == 0 && (hw * esize) % 16 == 0 && hw * esize <= 64 &&
           n * (channels / poolring::CB) * frames < 0x7fffffffLL;
}

static int launch_pool_ring(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw,
                            float* pooled, cudaStream_t st, uint8_t* img = nullptr) {
    const int esize = is_bf16 ? 2 : 4;
    static bool configured[64] = {};
    int dev = 0;
    CMHAR_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(poolring::video_pool_ring_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, poolring::SMEM_BYTES));
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(poolring::video_pool_ring_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, poolring::SMEM_BYTES));
        // an SM configured by this kernel alone must still be able to take an encoder CTA: ask for the largest carveout
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(poolring::video_pool_ring_kernel<__nv_bfloat16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CMHAR_CHECK_CUDA(cudaFuncSetAttribute(poolring::video_pool_ring_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured[dev & 63] = true;
    }
    static int cpt_cap = -1;
    if (cpt_cap < 0) { const char* e = dev_getenv("CMHAR_POOL_CPT"); cpt_cap = e ? atoi(e) : 1; }     // 4 channels per thread (16 KiB slabs, one unit per clip): +12 % alone, but 256 units on 148 CTAs is 2 uneven rounds
    int cpt = 1;
    for (int c = 4; c >= 1; --c) if (c <= cpt_cap && channels % (poolring::CB * c) == 0 && poolring::CB * c * hw * esize <= 16384) { cpt = c; break; }
    const long long units = n * (channels / (poolring::CB * cpt));
    static int per_sm = -1;
    if (per_sm < 0) { const char* e = dev_getenv("CMHAR_POOL_CTAS_PER_SM"); per_sm = e ? atoi(e) : 1; }
    const long long cap = (long long)per_sm * sm_count();
    const unsigned grid = (unsigned)(units < cap ? units : cap);
    static int stage_target = -1;
    if (stage_target < 0) { const char* e = dev_getenv("CMHAR_POOL_STAGE"); stage_target = e ? atoi(e) : poolring::STAGE_TARGET; }
    static int pf_ahead = -1;
    if (pf_ahead < 0) { const char* e = dev_getenv("CMHAR_POOL_PF"); pf_ahead = e ? atoi(e) : 0; }       // ring stages prefetched into L2 ahead of the producer (measured: 1.25 -> 1.09 TB/s, off)
    if (is_bf16) poolring::video_pool_ring_kernel<__nv_bfloat16><<<grid, poolring::CB, poolring::SMEM_BYTES, st>>>((const __nv_bfloat16*)fmap, n, frames, channels, hw, pooled, stage_target, cpt, img, pf_ahead);
    else poolring::video_pool_ring_kernel<float><<<grid, poolring::CB, poolring::SMEM_BYTES, st>>>((const float*)fmap, n, frames, channels, hw, pooled, stage_target, cpt, img, pf_ahead);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_video_pool_coresident(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw,
                                float* pooled, cmhar_stream_t s) {
    CMHAR_REQUIRE(fmap && pooled && frames > 0 && channels > 0 && hw > 0, "cmhar_video_pool_coresident: bad argument");
    if (n <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(pool_ring_eligible(fmap, is_bf16, n, frames, channels, hw),
                  "cmhar_video_pool_coresident: needs a 16-byte aligned map, channels %% 128 == 0 and 16..64 bytes per channel (got channels=%d hw=%d)", channels, hw);
    return launch_pool_ring(fmap, is_bf16, n, frames, channels, hw, pooled, (cudaStream_t)s);
}

static int video_pool_impl(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw,
                           float* pooled, uint8_t* img, cmhar_stream_t s, uint8_t* fimg = nullptr) {
    if (n <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(fmap && (pooled || img || fimg) && frames > 0 && channels > 0 && hw > 0, "cmhar_video_pool: bad argument");
    CMHAR_REQUIRE(!(img || fimg) || channels % 64 == 0, "cmhar_video_pool_img: an operand image needs channels %% 64 == 0 (got %d)", channels);
    if (n <= 0) return CMHAR_OK;
    cudaStream_t st = (cudaStream_t)s;
    // Default: the flooding kernel below (101 % of the measured copy bandwidth at 2 048 clips; 60 % at 256).  The
    // co-resident ring kernel (cmhar_video_pool_coresident, or CMHAR_POOL_MODE=2 here) does overlap the encoder --
    // encoder || pooling 28.5 -> 22.5-27 us per 256-window step -- but one 32 KiB ring per SM sustains only ~1.4 TB/s
    // (HBM latency under load is ~3 us), which stretches every lane's dependency chain: the full 16-lane step measured
    // 29.9 us on one B200 box and 34.3 us on another, against a steady 32.5 us for the flooding kernel.
    static int mode = -1;
    if (mode < 0) { const char* e = dev_getenv("CMHAR_POOL_MODE"); mode = e ? atoi(e) : 1; }        // 1 = flood (default), 2 = ring
    if (mode == 2 && !fimg && pool_ring_eligible(fmap, is_bf16, n, frames, channels, hw))
        return launch_pool_ring(fmap, is_bf16, n, frames, channels, hw, pooled, st, img);
    static int batched = -1;
    if (batched < 0) { const char* e = dev_getenv("CMHAR_POOL_BATCH"); batched = e ? atoi(e) : 1; }      // development switch
    static int pad_smem = -1;
    if (pad_smem < 0) { const char* e = dev_getenv("CMHAR_POOL_SMEM"); pad_smem = e ? atoi(e) : 0; }      // development switch: unused dynamic
                                                                                                      // smem that keeps the CTAs off SMs holding an encoder CTA
    static int cap_y = -1;
    if (cap_y < 0) { const char* e = dev_getenv("CMHAR_POOL_GRIDY"); cap_y = e ? atoi(e) : 0; }      // development switch
    long long gy = n < 32768 ? n : 32768;
    if (cap_y > 0 && gy > cap_y) gy = cap_y;
    dim3 grid((channels + POOL_NT - 1) / POOL_NT, (unsigned)gy);
    const bool aligned = ((uintptr_t)fmap & 15) == 0;
    if (is_bf16) {
        if (aligned && hw % 8 == 0)
        {
            if (batched == 8) video_pool_kernel<__nv_bfloat16, 8, 8><<<grid, POOL_NT, pad_smem, st>>>((const __nv_bfloat16*)fmap, n, frames, channels, hw, pooled, batched, img, fimg);
            else video_pool_kernel<__nv_bfloat16, 8><<<grid, POOL_NT, pad_smem, st>>>((const __nv_bfloat16*)fmap, n, frames, channels, hw, pooled, batched, img, fimg);
        }
        else
            video_pool_kernel<__nv_bfloat16, 0><<<grid, POOL_NT, pad_smem, st>>>((const __nv_bfloat16*)fmap, n, frames, channels, hw, pooled, batched, img, fimg);
    } else {
        if (aligned && hw % 4 == 0)
            video_pool_kernel<float, 4><<<grid, POOL_NT, pad_smem, st>>>((const float*)fmap, n, frames, channels, hw, pooled, batched, img, fimg);
        else
            video_pool_kernel<float, 0><<<grid, POOL_NT, pad_smem, st>>>((const float*)fmap, n, frames, channels, hw, pooled, batched, img, fimg);
    }
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_video_pool(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw,
                     float* pooled, cmhar_stream_t s) {
    CMHAR_REQUIRE(pooled, "cmhar_video_pool: null output");
    return video_pool_impl(fmap, is_bf16, n, frames, channels, hw, pooled, nullptr, s);
}

int cmhar_video_pool_img(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw,
                         float* pooled, void* pooled_img, cmhar_stream_t s) {
    return video_pool_impl(fmap, is_bf16, n, frames, channels, hw, pooled, reinterpret_cast<uint8_t*>(pooled_img), s);
}

int cmhar_video_pool_frames_img(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw,
                                float* pooled, void* pooled_img, void* frame_img, cmhar_stream_t s) {
    CMHAR_REQUIRE(n <= 0 || frame_img, "cmhar_video_pool_frames_img: null frame image");
    return video_pool_impl(fmap, is_bf16, n, frames, channels, hw, pooled, reinterpret_cast<uint8_t*>(pooled_img), s,
                           reinterpret_cast<uint8_t*>(frame_img));
}

}  // extern "C"
