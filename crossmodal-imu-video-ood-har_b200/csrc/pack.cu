// pack.cu -- one-time weight preparation (BatchNorm folding, transposes, scale folding) and the
// C-ABI entry points of the IMU path.  Packing runs on the device so the ABI needs nothing but
// the state_dict tensors' device pointers (SURVEY.md Appendix B names them).
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace cmhar {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- host-side registry of packed blobs: the launchers need a blob's dimensions (and whether it carries the
// tensor-core section) without a device->host read.  Keyed by the blob pointer; an unknown pointer (a blob the
// caller copied elsewhere) simply takes the CUDA-core path.
std::atomic<int> g_dev_env{0};
static std::mutex g_blob_mu;
static std::unordered_map<const void*, BlobInfo> g_blobs;
void register_blob(const void* blob, const BlobInfo& info) {
    std::lock_guard<std::mutex> lk(g_blob_mu);
    g_blobs[blob] = info;
}
void release_blob(const void* blob) {
    std::lock_guard<std::mutex> lk(g_blob_mu);
    g_blobs.erase(blob);
}
bool lookup_blob(const void* blob, BlobInfo* info) {
    std::lock_guard<std::mutex> lk(g_blob_mu);
    auto it = g_blobs.find(blob);
    if (it == g_blobs.end()) return false;
    *info = it->second;
    return true;
}

bool head_tc_eligible(int h1, int h2, int C);                                                   // head_tc.cu
size_t head_tc_bytes();
size_t maha_tc_bytes();
size_t maha_score_tc_bytes();
int pack_maha_score_tc(const float* f32, const MahaLayout& ml, uint8_t* dst, cudaStream_t st);              // maha_score_tc.cu
// the streaming score kernel's [W | G] image sits behind the head kernel's maha section, 1 KiB aligned
inline size_t maha_score_section_offset() { return (maha_tc_bytes() + 1023) / 1024 * 1024; }
int pack_head_tc(const float* f32, const HeadLayout& hl, uint8_t* dst, cudaStream_t st);
int pack_maha_tc(const float* f32, const MahaLayout& ml, uint8_t* dst, cudaStream_t st);

// dst[k*ldd + n] = src[n*K + k] * (n < scaled_rows ? scale : 1) * (col_scale ? col_scale(n) : 1)
// col scale = bn_w[n] / sqrt(bn_var[n] + eps) when bn_w != nullptr (computed in double)
__global__ void transpose_fold_kernel(const float* __restrict__ src, int N, int K, float* __restrict__ dst,
                                      int ldd, int scaled_rows, float scale,
                                      const float* __restrict__ bn_w, const float* __restrict__ bn_var) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * K) return;
    const int n = (int)(i % N), k = (int)(i / N);          // consecutive threads -> consecutive n (coalesced store)
    double v = src[(size_t)n * K + k];
    if (n < scaled_rows) v *= scale;
    if (bn_w) v *= (double)bn_w[n] / sqrt((double)bn_var[n] + (double)BN_EPS);
    dst[(size_t)k * ldd + n] = (float)v;
}

// dst[n] = (b[n] - mean[n]) * bn_w[n]/sqrt(var[n]+eps) + bn_b[n]   (or b[n]*scale for n<scaled_rows)
__global__ void bias_fold_kernel(const float* __restrict__ b, int N, float* __restrict__ dst, int scaled_rows,
                                 float scale, const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                 const float* __restrict__ bn_mean, const float* __restrict__ bn_var) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double v = b ? (double)b[n] : 0.0;
    if (n < scaled_rows) v *= scale;
    if (bn_w) v = (v - (double)bn_mean[n]) * ((double)bn_w[n] / sqrt((double)bn_var[n] + (double)BN_EPS)) + (double)bn_b[n];
    dst[n] = (float)v;
}

__global__ void tok_bias_kernel(const float* __restrict__ cls, const float* __restrict__ pos,
                                const float* __restrict__ pbias, int seq, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CMHAR_MAX_SEQ * D) return;
    const int s = i / D, c = i % D;
    float v = 0.f;
    if (s < seq) v = (s == 0 ? cls[c] : pbias[c]) + pos[s * D + c];
    dst[i] = v;
}

__global__ void copy2_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = a[i];
    else if (i < 2 * n) dst[i] = b[i - n];
}

__global__ void write_header_kernel(BlobHeader* dst, BlobHeader hdr) { *dst = hdr; }

__global__ void maha_valid_kernel(const float* __restrict__ count, int C, float* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) dst[c] = count ? (count[c] > 0.f ? 1.f : 0.f) : 1.f;
}

static int transpose_fold(const float* src, int N, int K, float* dst, int ldd, int scaled_rows, float scale,
                          const float* bn_w, const float* bn_var, cudaStream_t st) {
    const long long tot = (long long)N * K;
    transpose_fold_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, N, K, dst, ldd, scaled_rows, scale, bn_w, bn_var);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}
static int bias_fold(const float* b, int N, float* dst, int scaled_rows, float scale, const float* bn_w,
                     const float* bn_b, const float* bn_mean, const float* bn_var, cudaStream_t st) {
    bias_fold_kernel<<<(N + 255) / 256, 256, 0, st>>>(b, N, dst, scaled_rows, scale, bn_w, bn_b, bn_mean, bn_var);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}
static int copy2(const float* a, const float* b, int n, float* dst, cudaStream_t st) {
    copy2_kernel<<<(2 * n + 255) / 256, 256, 0, st>>>(a, b, n, dst);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}
static int write_header(void* blob, const BlobHeader& h, cudaStream_t st) {
    write_header_kernel<<<1, 1, 0, st>>>(reinterpret_cast<BlobHeader*>(blob), h);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

// implemented in imu_encoder_bf16.cu: bf16 operand images for the tcgen05 path
size_t encoder_bf16_bytes(int layers);
int pack_encoder_bf16(const cmhar_imu_encoder_params* p, const float* fp32_section, void* bf16_section, cudaStream_t st);

}  // namespace cmhar

using namespace cmhar;

#define TRY(expr)                      \
    do {                               \
        int _rc = (expr);              \
        if (_rc != CMHAR_OK) return _rc; \
    } while (0)

extern "C" {

int cmhar_abi_version(void) { return CMHAR_ABI_VERSION; }
const char* cmhar_last_error(void) { return g_err; }
int64_t cmhar_launch_count(void) { return (int64_t)g_launches.load(); }

size_t cmhar_imu_encoder_blob_bytes(int32_t seq, int32_t layers) {
    if (seq < 2 || seq > CMHAR_MAX_SEQ || layers < 1 || layers > CMHAR_MAX_LAYERS) return 0;
    size_t fp32 = EncLayout::fp32_floats(layers) * sizeof(float);
    fp32 = (fp32 + 1023) / 1024 * 1024;
    return sizeof(BlobHeader) + 960 /* keep sections 1 KiB aligned */ + fp32 + encoder_bf16_bytes(layers);
}

int cmhar_imu_encoder_pack(const cmhar_imu_encoder_params* p, void* blob, cmhar_stream_t s) {
    CMHAR_REQUIRE(p && blob, "cmhar_imu_encoder_pack: null argument");
    CMHAR_REQUIRE(p->seq >= 2 && p->seq <= CMHAR_MAX_SEQ, "seq=%d outside [2,%d]", p->seq, CMHAR_MAX_SEQ);
    CMHAR_REQUIRE(p->layers >= 1 && p->layers <= CMHAR_MAX_LAYERS, "layers=%d outside [1,%d]", p->layers, CMHAR_MAX_LAYERS);
    CMHAR_REQUIRE(((uintptr_t)blob & 1023) == 0, "blob must be 1024-byte aligned");
    cudaStream_t st = (cudaStream_t)s;
    float* f = reinterpret_cast<float*>(reinterpret_cast<char*>(blob) + 1024);
    TRY(transpose_fold(p->patch_weight, D, P, f + EncLayout::patch_wt, D, 0, 1.f, nullptr, nullptr, st));
    tok_bias_kernel<<<(CMHAR_MAX_SEQ * D + 255) / 256, 256, 0, st>>>(p->cls_token, p->pos_encoding, p->patch_bias, p->seq, f + EncLayout::tok_bias);
    CMHAR_LAUNCH_CHECK();
    TRY(copy2(p->norm_weight, p->norm_bias, D, f + EncLayout::final_ln, st));
    for (int l = 0; l < p->layers; ++l) {
        const cmhar_encoder_layer_params& q = p->layer[l];
        float* L = f + EncLayout::layers0 + (size_t)l * EncLayout::layer_floats;
        // 1/sqrt(head_dim) = 0.25 folded into the q rows: exact (power of two)
        TRY(transpose_fold(q.in_proj_weight, 3 * D, D, L + EncLayout::l_w_in, 3 * D, D, 0.25f, nullptr, nullptr, st));
        TRY(bias_fold(q.in_proj_bias, 3 * D, L + EncLayout::l_b_in, D, 0.25f, nullptr, nullptr, nullptr, nullptr, st));
        TRY(transpose_fold(q.out_proj_weight, D, D, L + EncLayout::l_w_o, D, 0, 1.f, nullptr, nullptr, st));
        TRY(bias_fold(q.out_proj_bias, D, L + EncLayout::l_b_o, 0, 1.f, nullptr, nullptr, nullptr, nullptr, st));
        TRY(transpose_fold(q.linear1_weight, FF, D, L + EncLayout::l_w1, FF, 0, 1.f, nullptr, nullptr, st));
        TRY(bias_fold(q.linear1_bias, FF, L + EncLayout::l_b1, 0, 1.f, nullptr, nullptr, nullptr, nullptr, st));
        TRY(transpose_fold(q.linear2_weight, D, FF, L + EncLayout::l_w2, D, 0, 1.f, nullptr, nullptr, st));
        TRY(bias_fold(q.linear2_bias, D, L + EncLayout::l_b2, 0, 1.f, nullptr, nullptr, nullptr, nullptr, st));
        TRY(copy2(q.norm1_weight, q.norm1_bias, D, L + EncLayout::l_ln1, st));
        TRY(copy2(q.norm2_weight, q.norm2_bias, D, L + EncLayout::l_ln2, st));
    }
    size_t fp32 = (EncLayout::fp32_floats(p->layers) * sizeof(float) + 1023) / 1024 * 1024;
    TRY(pack_encoder_bf16(p, f, reinterpret_cast<char*>(blob) + 1024 + fp32, st));
    BlobHeader h{};
    h.magic = ENC_MAGIC; h.a = p->seq; h.b = p->layers; h.has_bf16 = 1;
    return write_header(blob, h, st);
}

size_t cmhar_head_blob_bytes(int32_t h1, int32_t h2, int32_t C) {
    if (h1 < 4 || h1 > 256 || (h1 & 3) || h2 < 4 || h2 > 256 || (h2 & 3) || C < 2 || C > CMHAR_MAX_CLASSES) return 0;
    HeadLayout hl{h1, h2, C};
    return tc_section_offset(hl.total()) + (head_tc_eligible(h1, h2, C) ? head_tc_bytes() : 0);
}

int cmhar_head_pack(const cmhar_head_params* p, void* blob, cmhar_stream_t s) {
    CMHAR_REQUIRE(p && blob, "cmhar_head_pack: null argument");
    CMHAR_REQUIRE(cmhar_head_blob_bytes(p->hidden1, p->hidden2, p->classes) != 0,
                  "unsupported head dims (%d,%d,%d): hidden <= 256 and a multiple of 4, classes <= %d", p->hidden1, p->hidden2, p->classes, CMHAR_MAX_CLASSES);
    cudaStream_t st = (cudaStream_t)s;
    HeadLayout hl{p->hidden1, p->hidden2, p->classes};
    float* f = reinterpret_cast<float*>(reinterpret_cast<char*>(blob) + sizeof(BlobHeader));
    TRY(transpose_fold(p->w0, hl.h1, D, f + hl.w0(), hl.h1, 0, 1.f, p->bn0_weight, p->bn0_var, st));
    TRY(bias_fold(p->b0, hl.h1, f + hl.b0(), 0, 1.f, p->bn0_weight, p->bn0_bias, p->bn0_mean, p->bn0_var, st));
    TRY(transpose_fold(p->w1, hl.h2, hl.h1, f + hl.w1(), hl.h2, 0, 1.f, p->bn1_weight, p->bn1_var, st));
    TRY(bias_fold(p->b1, hl.h2, f + hl.b1(), 0, 1.f, p->bn1_weight, p->bn1_bias, p->bn1_mean, p->bn1_var, st));
    CMHAR_CHECK_CUDA(cudaMemsetAsync(f + hl.w2(), 0, ((size_t)hl.h2 * hl.Cp() + hl.Cp()) * sizeof(float), st));   // zero pad columns
    TRY(transpose_fold(p->w2, hl.C, hl.h2, f + hl.w2(), hl.Cp(), 0, 1.f, nullptr, nullptr, st));
    TRY(bias_fold(p->b2, hl.C, f + hl.b2(), 0, 1.f, nullptr, nullptr, nullptr, nullptr, st));
    BlobHeader h{};
    h.magic = HEAD_MAGIC; h.a = hl.h1; h.b = hl.h2; h.c = hl.C;
    if (head_tc_eligible(hl.h1, hl.h2, hl.C)) {      // split-bf16 operand images for the tensor-core head (head_tc.cu)
        h.has_bf16 = 1;
        TRY(pack_head_tc(f, hl, reinterpret_cast<uint8_t*>(blob) + tc_section_offset(hl.total()), st));
    }
    register_blob(blob, BlobInfo{HEAD_MAGIC, hl.h1, hl.h2, hl.C, h.has_bf16});
    return write_header(blob, h, st);
}

size_t cmhar_maha_blob_bytes(int32_t C) {
    if (C < 1 || C > 1024) return 0;
    MahaLayout ml{C};
    return tc_section_offset(ml.total()) + (C <= 32 ? maha_score_section_offset() + maha_score_tc_bytes() : 0);
}

int cmhar_maha_pack(const float* whiten, const float* mean_whitened, const float* class_count, int32_t C,
                    void* blob, cmhar_stream_t s) {
    CMHAR_REQUIRE(whiten && mean_whitened && blob, "cmhar_maha_pack: null argument");
    CMHAR_REQUIRE(cmhar_maha_blob_bytes(C) != 0 && C <= 64, "classes=%d outside [1,64] (the limit of cmhar_maha_accumulate, which produces the state packed here)", C);
    cudaStream_t st = (cudaStream_t)s;
    MahaLayout ml{C};
    float* f = reinterpret_cast<float*>(reinterpret_cast<char*>(blob) + sizeof(BlobHeader));
    CMHAR_CHECK_CUDA(cudaMemcpyAsync(f + ml.whiten(), whiten, sizeof(float) * D * D, cudaMemcpyDeviceToDevice, st));
    CMHAR_CHECK_CUDA(cudaMemcpyAsync(f + ml.mean_w(), mean_whitened, sizeof(float) * C * D, cudaMemcpyDeviceToDevice, st));
    maha_valid_kernel<<<(C + 127) / 128, 128, 0, st>>>(class_count, C, f + ml.valid());
    CMHAR_LAUNCH_CHECK();
    BlobHeader h{};
    h.magic = MAHA_MAGIC; h.a = C;
    if (C <= 32) {
        h.has_bf16 = 1;
        TRY(pack_maha_tc(f, ml, reinterpret_cast<uint8_t*>(blob) + tc_section_offset(ml.total()), st));
        TRY(pack_maha_score_tc(f, ml, reinterpret_cast<uint8_t*>(blob) + tc_section_offset(ml.total()) + maha_score_section_offset(), st));
    }
    register_blob(blob, BlobInfo{MAHA_MAGIC, C, 0, 0, h.has_bf16});
    return write_header(blob, h, st);
}

}  // extern "C"
