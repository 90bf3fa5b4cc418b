// video_frames.cu -- the two hand-written kernels either side of the video trunk when the trunk runs on the device in
// channels-last bf16 under a CUDA graph (host side: video_trunk.py).  The trunk itself (torchvision resnet18 /
// mobilenet_v2, reference src/models/models.py:163-173) is third-party library code (cuDNN) and stays that.
//   frames_normalize_kernel   decoded uint8 HWC frames -> normalised bf16 channels-last pixels: the reference's per-frame CPU transform
//                             ToTensor + Normalize (src/data/datasets.py:52-58) moved behind the H2D copy, so a clip crosses PCIe as
//                             602 KB of bytes instead of 2.4 MB of floats
//   video_pool_nhwc_kernel    the video tail's reduction (models.py:210-211,215) straight from the trunk's channels-last output
//                             (physical layout (B*T, h*w, F)): no NCHW copy of the feature map exists
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cmhar {
namespace vframes {

// one thread = 4 pixels = 12 input bytes (three aligned 32-bit loads); CPAD output channels per pixel, channels >= 3 zero
// (CPAD = 8: one 16-byte store per pixel and a first convolution whose input channels are a multiple of 8).
// Arithmetic order is torch's: (x / 255 - mean) / std in fp32, IEEE divisions, no contraction; one rounding to bf16.
template <int CPAD>
__global__ void __launch_bounds__(256) frames_normalize_kernel(const uint8_t* __restrict__ in, long long n_px, float3 mean, float3 sd,
                                                               __nv_bfloat16* __restrict__ out) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 pixels
    const long long p0 = q * 4;
    if (p0 >= n_px) return;
    const float mu[3] = {mean.x, mean.y, mean.z}, sg[3] = {sd.x, sd.y, sd.z};
    uint8_t px[12];
    const int cnt = (n_px - p0 >= 4) ? 4 : (int)(n_px - p0);
    if (cnt == 4 && ((uintptr_t)in & 3) == 0) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(in + p0 * 3);
        uint32_t u[3] = {__ldg(w), __ldg(w + 1), __ldg(w + 2)};
#pragma unroll
        for (int i = 0; i < 12; ++i) px[i] = (uint8_t)(u[i >> 2] >> (8 * (i & 3)));
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) px[i] = (i < cnt * 3) ? in[p0 * 3 + i] : (uint8_t)0;
    }
    __align__(16) __nv_bfloat16 v[4][CPAD];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
#pragma unroll
        for (int c = 0; c < CPAD; ++c)
            v[p][c] = (c < 3) ? __float2bfloat16_rn(__fdiv_rn(__fsub_rn(__fdiv_rn((float)px[p * 3 + (c < 3 ? c : 0)], 255.f), mu[c < 3 ? c : 0]), sg[c < 3 ? c : 0]))
                              : __float2bfloat16_rn(0.f);
    }
    __nv_bfloat16* dst = out + p0 * CPAD;
    if (cnt == 4 && ((uintptr_t)out & 15) == 0 && (CPAD * 8) % 16 == 0) {     // CPAD 4 / 8: 16-byte stores
        const uint4* s = reinterpret_cast<const uint4*>(&v[0][0]);
#pragma unroll
        for (int i = 0; i < CPAD * 8 / 16; ++i) reinterpret_cast<uint4*>(dst)[i] = s[i];
    } else if (cnt == 4 && ((uintptr_t)out & 7) == 0) {                        // CPAD 3: 24 bytes per thread as three 8-byte stores
        const uint2* s = reinterpret_cast<const uint2*>(&v[0][0]);
#pragma unroll
        for (int i = 0; i < CPAD; ++i) reinterpret_cast<uint2*>(dst)[i] = s[i];
    } else {
        for (int p = 0; p < cnt; ++p)
#pragma unroll
            for (int c = 0; c < CPAD; ++c) dst[p * CPAD + c] = v[p][c];
    }
}

// fmap physical (n * frames, hw, channels); CTA = (64 channel groups of 16 bytes) x (GY frame slices: frames t = y mod GY) of one clip.
// A warp reads 512 contiguous bytes per pixel; a thread keeps up to 8 16-byte loads in flight; the slices meet in shared memory.
// pooled (n, channels) fp32 / img: clip means as fp32 rows / bf16 operand image; fimg: per-frame spatial means (row = clip * frames + t).
constexpr int GX = 64, GY = 8, LD_BATCH = 8;
template <typename T>
__global__ void __launch_bounds__(GX * GY) video_pool_nhwc_kernel(const T* __restrict__ fmap, long long n, int frames, int channels, int hw,
                                                                  float* __restrict__ pooled, uint8_t* __restrict__ img,
                                                                  uint8_t* __restrict__ fimg) {
    constexpr int V = 16 / (int)sizeof(T);                      // channels per 16-byte load
    __shared__ float red[GY][GX][V + 1];
    const int groups = channels / V;
    const int gx = threadIdx.x, y = threadIdx.y;
    const int g = blockIdx.x * GX + gx;
    const bool live = g < groups;
    uint64_t stream_policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream_policy));     // read-once stream
    auto ld16 = [&](const uint4* q) {
        uint4 u;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(q), "l"(stream_policy));
        return u;
    };
    auto add16 = [&](float (&a)[V], const uint4& u) {
        if constexpr (sizeof(T) == 2) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); a[2 * k] += f.x; a[2 * k + 1] += f.y; }
        } else {
            const float* f = reinterpret_cast<const float*>(&u);
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] += f[k];
        }
    };
    auto put_img = [&](uint8_t* image, long long row, const float (&m)[V]) {      // V consecutive channels of one row
        const int c = g * V;
        uint8_t* chunk = image + ((size_t)(row >> 7) * (channels >> 6) + (c >> 6)) * 16384;
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(chunk + tc::sw128_off((int)(row & 127), (c & 63) >> 3) + (c & 7) * 2);
        if constexpr (V == 8) {
            uint4 u;
            u.x = tc::pack_bf16(m[0], m[1]); u.y = tc::pack_bf16(m[2], m[3]); u.z = tc::pack_bf16(m[4], m[5]); u.w = tc::pack_bf16(m[6], m[7]);
            *reinterpret_cast<uint4*>(d) = u;
        } else {
            uint2 u;
            u.x = tc::pack_bf16(m[0], m[1]); u.y = tc::pack_bf16(m[2], m[3]);
            *reinterpret_cast<uint2*>(d) = u;
        }
    };
    const float inv_hw = 1.f / (float)hw, inv_all = 1.f / (float)((long long)frames * hw);
    for (long long b = blockIdx.y; b < n; b += gridDim.y) {
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        if (live) {
            for (int t = y; t < frames; t += GY) {
                const uint4* base = reinterpret_cast<const uint4*>(fmap + ((size_t)(b * frames + t) * hw) * channels) + g;
                float fs[V];
#pragma unroll
                for (int k = 0; k < V; ++k) fs[k] = 0.f;
                for (int p0 = 0; p0 < hw; p0 += LD_BATCH) {
                    uint4 u[LD_BATCH];
#pragma unroll
                    for (int i = 0; i < LD_BATCH; ++i) {
                        u[i] = make_uint4(0u, 0u, 0u, 0u);
                        if (p0 + i < hw) u[i] = ld16(base + (size_t)(p0 + i) * groups);
                    }
#pragma unroll
                    for (int i = 0; i < LD_BATCH; ++i) asm volatile("" : "+r"(u[i].x), "+r"(u[i].y), "+r"(u[i].z), "+r"(u[i].w));      // all loads before the first add
#pragma unroll
                    for (int i = 0; i < LD_BATCH; ++i) add16(fs, u[i]);
                }
#pragma unroll
                for (int k = 0; k < V; ++k) acc[k] += fs[k];
                if (fimg) {
#pragma unroll
                    for (int k = 0; k < V; ++k) fs[k] *= inv_hw;
                    put_img(fimg, b * frames + t, fs);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) red[y][gx][k] = acc[k];
        __syncthreads();
        if (y == 0 && live) {
            float m[V];
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float sum = 0.f;
#pragma unroll
                for (int yy = 0; yy < GY; ++yy) sum += red[yy][gx][k];
                m[k] = sum * inv_all;
            }
            if (pooled) {
#pragma unroll
                for (int k = 0; k < V; ++k) pooled[(size_t)b * channels + g * V + k] = m[k];
            }
            if (img) put_img(img, b, m);
        }
        __syncthreads();
    }
}

}  // namespace vframes
}  // namespace cmhar

using namespace cmhar;

extern "C" {

int cmhar_frames_normalize(const uint8_t* frames_u8, int64_t n_pixels, const float* mean3, const float* std3, int32_t cpad,
                           void* out_bf16, cmhar_stream_t s) {
    CMHAR_REQUIRE(n_pixels <= 0 || (frames_u8 && out_bf16 && mean3 && std3), "cmhar_frames_normalize: null argument");
    CMHAR_REQUIRE(cpad == 3 || cpad == 4 || cpad == 8, "cmhar_frames_normalize: cpad must be 3, 4 or 8 (got %d)", cpad);
    if (n_pixels <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "cmhar_frames_normalize: zero std");
    const float3 mu = make_float3(mean3[0], mean3[1], mean3[2]), sg = make_float3(std3[0], std3[1], std3[2]);
    const long long quads = (n_pixels + 3) / 4;
    CMHAR_REQUIRE((quads + 255) / 256 < 0x7fffffffLL, "cmhar_frames_normalize: too many pixels");
    const unsigned grid = (unsigned)((quads + 255) / 256);
    cudaStream_t st = (cudaStream_t)s;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    if (cpad == 8) vframes::frames_normalize_kernel<8><<<grid, 256, 0, st>>>(frames_u8, n_pixels, mu, sg, out);
    else if (cpad == 4) vframes::frames_normalize_kernel<4><<<grid, 256, 0, st>>>(frames_u8, n_pixels, mu, sg, out);
    else vframes::frames_normalize_kernel<3><<<grid, 256, 0, st>>>(frames_u8, n_pixels, mu, sg, out);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

int cmhar_video_pool_nhwc(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw,
                          float* pooled, void* pooled_img, void* frame_img, cmhar_stream_t s) {
    if (n <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(fmap && (pooled || pooled_img || frame_img) && frames > 0 && channels > 0 && hw > 0, "cmhar_video_pool_nhwc: bad argument");
    const int V = is_bf16 ? 8 : 4;
    CMHAR_REQUIRE(((uintptr_t)fmap & 15) == 0 && channels % V == 0,
                  "cmhar_video_pool_nhwc: needs a 16-byte aligned map and channels %% %d == 0 (got %d)", V, channels);
    CMHAR_REQUIRE(!(pooled_img || frame_img) || channels % 64 == 0, "cmhar_video_pool_nhwc: an operand image needs channels %% 64 == 0 (got %d)", channels);
    const int groups = channels / V;
    dim3 grid((groups + vframes::GX - 1) / vframes::GX, (unsigned)(n < 65535 ? n : 65535)), block(vframes::GX, vframes::GY);
    cudaStream_t st = (cudaStream_t)s;
    if (is_bf16)
        vframes::video_pool_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)fmap, n, frames, channels, hw, pooled,
                                                                                reinterpret_cast<uint8_t*>(pooled_img), reinterpret_cast<uint8_t*>(frame_img));
    else
        vframes::video_pool_nhwc_kernel<float><<<grid, block, 0, st>>>((const float*)fmap, n, frames, channels, hw, pooled,
                                                                        reinterpret_cast<uint8_t*>(pooled_img), reinterpret_cast<uint8_t*>(frame_img));
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // extern "C"
