// common.cuh -- shared host/device helpers for the cmhar_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/cmhar_b200.h"

namespace cmhar {

// ---- error reporting -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CMHAR_CHECK_CUDA(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::cmhar::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                               __FILE__, __LINE__);                                         \
            return CMHAR_ERR_CUDA;                                                          \
        }                                                                                   \
    } while (0)

#define CMHAR_REQUIRE(cond, ...)                                                            \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            ::cmhar::set_error(__VA_ARGS__);                                                \
            return CMHAR_ERR_INVALID;                                                       \
        }                                                                                   \
    } while (0)

#define CMHAR_LAUNCH_CHECK()                                                                \
    do {                                                                                    \
        ::cmhar::count_launch();                                                            \
        CMHAR_CHECK_CUDA(cudaGetLastError());                                               \
    } while (0)

// ---- sizes fixed by the reference configuration -------------------------------------------
constexpr int D = CMHAR_D_MODEL;      // 128
constexpr int H = CMHAR_NHEAD;        // 8
constexpr int HD = D / H;             // 16
constexpr int FF = CMHAR_FFN;         // 512
constexpr int P = CMHAR_PATCH;        // 16
constexpr float LN_EPS = 1e-5f;
constexpr float BN_EPS = 1e-5f;

// ---- fp32 encoder blob layout (float offsets) ----------------------------------------------
// [patch_wt (16,128)] [tok_bias (16,128)] [final ln w,b (2,128)] then per layer:
// [w_in_t (128,384)] [b_in 384] [w_o_t (128,128)] [b_o 128] [w1_t (128,512)] [b1 512]
// [w2_t (512,128)] [b2 128] [ln1 w,b 256] [ln2 w,b 256]
struct EncLayout {
    static constexpr size_t patch_wt = 0;
    static constexpr size_t tok_bias = patch_wt + P * D;
    static constexpr size_t final_ln = tok_bias + CMHAR_MAX_SEQ * D;
    static constexpr size_t layers0 = final_ln + 2 * D;
    static constexpr size_t l_w_in = 0;
    static constexpr size_t l_b_in = l_w_in + D * 3 * D;
    static constexpr size_t l_w_o = l_b_in + 3 * D;
    static constexpr size_t l_b_o = l_w_o + D * D;
    static constexpr size_t l_w1 = l_b_o + D;
    static constexpr size_t l_b1 = l_w1 + D * FF;
    static constexpr size_t l_w2 = l_b1 + FF;
    static constexpr size_t l_b2 = l_w2 + FF * D;
    static constexpr size_t l_ln1 = l_b2 + D;
    static constexpr size_t l_ln2 = l_ln1 + 2 * D;
    static constexpr size_t layer_floats = l_ln2 + 2 * D;
    static constexpr size_t header_ints = 8;      // [magic, seq, layers, ...] stored as int32 in front
    __host__ __device__ static constexpr size_t fp32_floats(int layers) { return layers0 + (size_t)layers * layer_floats; }
};
constexpr uint32_t ENC_MAGIC = 0x434d4831u;   // "CMH1"
constexpr uint32_t HEAD_MAGIC = 0x434d4832u;
constexpr uint32_t MAHA_MAGIC = 0x434d4833u;
constexpr uint32_t LIN_MAGIC = 0x434d4834u;

struct BlobHeader {          // 64 bytes in front of every blob
    uint32_t magic;
    int32_t a, b, c, d;      // meaning depends on the blob kind
    int32_t has_bf16;
    int32_t pad[10];
};
static_assert(sizeof(BlobHeader) == 64, "header must stay 64 bytes");

// head blob (floats after header): w0_t (128,h1) b0 h1 | w1_t (h1,h2) b1 h2 | w2_t (h2,Cp) b2 Cp
// h1, h2 multiples of 4; the class dimension is padded to Cp = 4*ceil(C/4) (zero columns) so every
// matrix row starts 16-byte aligned (the head kernel streams them with cp.async.bulk).
struct HeadLayout {
    int h1, h2, C;
    __host__ __device__ int Cp() const { return (C + 3) & ~3; }
    __host__ __device__ size_t w0() const { return 0; }
    __host__ __device__ size_t b0() const { return w0() + (size_t)D * h1; }
    __host__ __device__ size_t w1() const { return b0() + h1; }
    __host__ __device__ size_t b1() const { return w1() + (size_t)h1 * h2; }
    __host__ __device__ size_t w2() const { return b1() + h2; }
    __host__ __device__ size_t b2() const { return w2() + (size_t)h2 * Cp(); }
    __host__ __device__ size_t total() const { return b2() + Cp(); }
};

// maha blob (floats after header): whiten (128,128) row-major [k][j], mean_w (C,128), mnorm (C)
// mnorm_c = ||mean_w_c||^2 or +inf for empty classes.
struct MahaLayout {
    int C;
    __host__ __device__ size_t whiten() const { return 0; }
    __host__ __device__ size_t mean_w() const { return (size_t)D * D; }
    __host__ __device__ size_t valid() const { return mean_w() + (size_t)C * D; }
    __host__ __device__ size_t total() const { return valid() + C; }
};

// The tensor-core section of a head / maha blob starts at the next 1 KiB boundary behind the fp32 section.
__host__ __device__ inline size_t tc_section_offset(size_t fp32_floats) {
    return (sizeof(BlobHeader) + fp32_floats * sizeof(float) + 1023) / 1024 * 1024;
}
struct BlobInfo { uint32_t magic; int a, b, c, has_tc; };
void register_blob(const void* blob, const BlobInfo& info);
bool lookup_blob(const void* blob, BlobInfo* info);
void release_blob(const void* blob);

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- CTA trace (development aid, cmhar_debug_cta_trace): when a buffer is installed, thread 0 of every CTA of the
// instrumented kernels appends {kernel id, SM id, start, end} (globaltimer ns) -- the only way to see how the CTAs
// of concurrently running kernels share the SMs without nsys.  One global load per CTA when disabled.
// (defined here: the library is a single translation unit, see cmhar_b200.cu)
__device__ unsigned long long* g_cta_trace = nullptr;        // [0] = record counter, [1] = capacity, then 4 words per record
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long trace_begin() { return g_cta_trace ? trace_now() : 0ull; }
__device__ __forceinline__ void trace_end(int kernel_id, unsigned long long t0) {
    unsigned long long* buf = g_cta_trace;
    if (!buf || t0 == 0ull) return;
    const unsigned long long slot = atomicAdd(buf, 1ull);
    if (slot >= buf[1]) return;
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* r = buf + 2 + 4 * slot;
    r[0] = (unsigned long long)kernel_id; r[1] = smid; r[2] = t0; r[3] = trace_now();
}
enum { TRACE_ENCODER = 1, TRACE_POOL = 2, TRACE_HEAD = 3, TRACE_LINEAR = 4, TRACE_SIM = 5 };

// Development switches (A/B measurements on one box) are read from the environment ONLY when the process enabled them through
// cmhar_debug_set_option("dev_env", 1): a stray CMHAR_* variable in a production environment changes nothing.
extern std::atomic<int> g_dev_env;
inline const char* dev_getenv(const char* name) { return g_dev_env.load(std::memory_order_relaxed) ? getenv(name) : nullptr; }

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace cmhar
