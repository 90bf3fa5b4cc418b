// api.cu -- C-ABI entry points of the fused IMU path (see include/cmhar_b200.h).
#include "common.cuh"

namespace cmhar {
int launch_imu_forward_fp32(const FwdArgs& a, cudaStream_t stream);
int launch_imu_forward_bf16(const FwdArgs& a, cudaStream_t stream);
int launch_head_forward(const FwdArgs& a, int precision, cudaStream_t stream);
int launch_imu_forward_bf16_debug(const FwdArgs& a, int stage, float* dump, int* progress, cudaStream_t stream);
int launch_mlp2_tc(const uint8_t* x_img, int K1, const uint8_t* w0_img, const float* b0, const uint8_t* w1_img, const float* b1,
                   long long n, int l2norm, float* y, uint8_t* y_img, cudaStream_t st);                       // mlp2_tc.cu
size_t similarity_img_work_bytes(long long na, long long nb);                                                  // similarity_tc.cu
int launch_similarity_img(const uint8_t* a_img, long long na, const uint8_t* const* b_parts, int n_parts, long long rows_per_part,
                          long long nb, int dim, long long diag_offset, float sig_scale, float sig_bias, double out_scale,
                          double* const* sum_dst, int n_dst, void* work, cudaStream_t st);

// tensor-core section of a packed linear blob (null when the blob is unknown or has none for these dims)
static const uint8_t* linear_tc_section(const void* blob, int in_dim, int out_dim, const float** bias) {
    BlobInfo bi{};
    if (!blob || !lookup_blob(blob, &bi) || bi.magic != LIN_MAGIC || !bi.has_tc || bi.a != in_dim || bi.b != out_dim) return nullptr;
    const float* f = reinterpret_cast<const float*>(reinterpret_cast<const char*>(blob) + sizeof(BlobHeader));
    if (bias) *bias = f + (size_t)in_dim * out_dim;
    return reinterpret_cast<const uint8_t*>(blob) + tc_section_offset(linear_fp32_floats(in_dim, out_dim));
}
}  // namespace cmhar

using namespace cmhar;

extern "C" {

static int imu_forward_impl(const void* encoder_blob, const void* head_blob, const void* maha_blob, const float* x,
                            int64_t n_windows, int64_t x_window_stride, const cmhar_imu_outputs& o, int32_t precision,
                            cmhar_stream_t s) {
    if (n_windows <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(encoder_blob && x, "cmhar_imu_forward: null encoder blob or input");
    CMHAR_REQUIRE(((uintptr_t)encoder_blob & 1023) == 0, "encoder blob must be 1024-byte aligned");
    CMHAR_REQUIRE(head_blob || !(o.logits || o.pred || o.msp || o.energy),
                  "logits/pred/msp/energy outputs need a head blob");
    CMHAR_REQUIRE(maha_blob || !o.maha, "maha_out needs a maha blob");
    CMHAR_REQUIRE(o.cls || !(head_blob || o.maha),
                  "cls_out is required with a head / maha blob: the head runs as a second launch on the stored CLS features");
    CMHAR_REQUIRE(x_window_stride >= 16, "x_window_stride=%lld too small", (long long)x_window_stride);
    CMHAR_REQUIRE(precision == CMHAR_FP32 || precision == CMHAR_BF16, "bad precision %d", precision);
    if (o.cls_img && precision != CMHAR_BF16) {
        set_error("cmhar_imu_forward_ex: the CLS operand image is written by the bf16 tensor-core kernel only");
        return CMHAR_ERR_UNSUPPORTED;
    }
    CMHAR_REQUIRE(!o.cls_img || ((uintptr_t)o.cls_img & 1023) == 0, "cls_img must be 1024-byte aligned");
    FwdArgs a{};
    a.enc_blob = reinterpret_cast<const char*>(encoder_blob);
    a.head_blob = reinterpret_cast<const char*>(head_blob);
    a.maha_blob = reinterpret_cast<const char*>(maha_blob);
    a.x = x; a.n = n_windows; a.xstride = x_window_stride;
    a.cls_out = o.cls; a.tokens_out = o.tokens; a.logits_out = o.logits;
    a.pred_out = reinterpret_cast<long long*>(o.pred);
    a.msp_out = o.msp; a.energy_out = o.energy; a.maha_out = o.maha; a.cls_img = o.cls_img;
    if (precision == CMHAR_BF16) return launch_imu_forward_bf16(a, (cudaStream_t)s);
    return launch_imu_forward_fp32(a, (cudaStream_t)s);
}

int cmhar_imu_forward(const void* encoder_blob, const void* head_blob, const void* maha_blob, const float* x,
                      int64_t n_windows, int64_t x_window_stride, float* cls_out, float* tokens_out,
                      float* logits_out, int64_t* pred_out, float* msp_out, float* energy_out, float* maha_out,
                      int32_t precision, cmhar_stream_t s) {
    cmhar_imu_outputs o{cls_out, nullptr, tokens_out, logits_out, pred_out, msp_out, energy_out, maha_out};
    return imu_forward_impl(encoder_blob, head_blob, maha_blob, x, n_windows, x_window_stride, o, precision, s);
}

int cmhar_imu_forward_ex(const void* encoder_blob, const void* head_blob, const void* maha_blob, const float* x,
                         int64_t n_windows, int64_t x_window_stride, const cmhar_imu_outputs* out, int32_t precision,
                         cmhar_stream_t s) {
    CMHAR_REQUIRE(out, "cmhar_imu_forward_ex: null output struct");
    return imu_forward_impl(encoder_blob, head_blob, maha_blob, x, n_windows, x_window_stride, *out, precision, s);
}

int cmhar_debug_cta_trace(uint64_t* device_buffer, int64_t capacity_records) {
    unsigned long long* ptr = reinterpret_cast<unsigned long long*>(device_buffer);
    if (ptr) {
        CMHAR_REQUIRE(capacity_records > 0, "cmhar_debug_cta_trace: capacity must be positive");
        const unsigned long long head[2] = {0ull, (unsigned long long)capacity_records};
        CMHAR_CHECK_CUDA(cudaMemcpy(ptr, head, sizeof(head), cudaMemcpyHostToDevice));
    }
    CMHAR_CHECK_CUDA(cudaMemcpyToSymbol(cmhar::g_cta_trace, &ptr, sizeof(ptr)));
    return CMHAR_OK;
}

int cmhar_blob_release(const void* blob) {
    if (blob) cmhar::release_blob(blob);
    return CMHAR_OK;
}

int cmhar_debug_set_option(const char* key, int32_t value) {
    CMHAR_REQUIRE(key, "cmhar_debug_set_option: null key");
    if (strcmp(key, "enc_kernel") == 0) {
        CMHAR_REQUIRE(value >= 0 && value <= 2, "cmhar_debug_set_option: enc_kernel must be 0, 1 or 2");
        cmhar::g_enc_kernel.store(value, std::memory_order_relaxed);
        return CMHAR_OK;
    }
    if (strcmp(key, "dev_env") == 0) {
        cmhar::g_dev_env.store(value != 0, std::memory_order_relaxed);
        return CMHAR_OK;
    }
    set_error("cmhar_debug_set_option: unknown key '%s'", key);
    return CMHAR_ERR_INVALID;
}

int cmhar_debug_imu_bf16(const void* encoder_blob, const float* x, int64_t n_windows, int64_t x_window_stride,
                         int32_t stage, float* residual_dump, float* cls_out, int32_t* progress_host_mapped,
                         cmhar_stream_t s) {
    CMHAR_REQUIRE(encoder_blob && x && residual_dump, "cmhar_debug_imu_bf16: null argument");
    if (n_windows <= 0) return CMHAR_OK;
    FwdArgs a{};
    a.enc_blob = reinterpret_cast<const char*>(encoder_blob);
    a.x = x; a.n = n_windows; a.xstride = x_window_stride; a.cls_out = cls_out;
    return launch_imu_forward_bf16_debug(a, stage, residual_dump, progress_host_mapped, (cudaStream_t)s);
}

int cmhar_head_forward(const void* head_blob, const void* maha_blob, const float* feat, int64_t n, float* logits_out,
                       int64_t* pred_out, float* msp_out, float* energy_out, float* maha_out, int32_t precision,
                       cmhar_stream_t s) {
    CMHAR_REQUIRE(precision == CMHAR_FP32 || precision == CMHAR_BF16, "bad precision %d", precision);
    CMHAR_REQUIRE(feat && (head_blob || maha_blob), "cmhar_head_forward: null argument");
    CMHAR_REQUIRE(head_blob || !(logits_out || pred_out || msp_out || energy_out),
                  "logits/pred/msp/energy outputs need a head blob");
    CMHAR_REQUIRE(maha_blob || !maha_out, "maha_out needs a maha blob");
    if (n <= 0) return CMHAR_OK;
    FwdArgs a{};
    a.head_blob = reinterpret_cast<const char*>(head_blob);
    a.maha_blob = reinterpret_cast<const char*>(maha_blob);
    a.x = feat; a.n = n; a.xstride = D;
    a.logits_out = logits_out; a.pred_out = reinterpret_cast<long long*>(pred_out);
    a.msp_out = msp_out; a.energy_out = energy_out; a.maha_out = maha_out;
    return launch_head_forward(a, precision, (cudaStream_t)s);
}

int cmhar_head_kernel_kind(const void* head_blob, const void* maha_blob, int32_t precision) {
    if (precision != CMHAR_BF16 || !(head_blob || maha_blob)) return 0;
    BlobInfo hi{}, mi{};
    const bool head_ok = !head_blob || (lookup_blob(head_blob, &hi) && hi.magic == HEAD_MAGIC && hi.has_tc);
    const bool maha_ok = !maha_blob || (lookup_blob(maha_blob, &mi) && mi.magic == MAHA_MAGIC && mi.has_tc);
    return (head_ok && maha_ok) ? 1 : 0;
}

int cmhar_mlp2_forward_img(const void* blob0, const void* blob1, const void* x_img, int64_t n, int32_t in_dim, int32_t hidden,
                           int32_t out_dim, int32_t l2_normalize, float* y, void* y_img, cmhar_stream_t s) {
    CMHAR_REQUIRE(blob0 && blob1 && x_img && (y || y_img), "cmhar_mlp2_forward_img: null argument");
    CMHAR_REQUIRE(((uintptr_t)x_img & 1023) == 0 && (!y_img || ((uintptr_t)y_img & 1023) == 0) && (!y || ((uintptr_t)y & 15) == 0),
                  "cmhar_mlp2_forward_img: operand images must be 1024-byte aligned, rows 16-byte aligned");
    if (n <= 0) return CMHAR_OK;
    const float *b0 = nullptr, *b1 = nullptr;
    const uint8_t* w0 = linear_tc_section(blob0, in_dim, hidden, &b0);
    const uint8_t* w1 = linear_tc_section(blob1, hidden, out_dim, &b1);
    if (hidden != 512 || out_dim != 256 || in_dim < 64 || in_dim % 64 || !w0 || !w1) {
        set_error("cmhar_mlp2_forward_img: fused projection head needs (in %% 64 == 0) -> 512 -> 256 with packed tensor-core sections; got %d -> %d -> %d",
                  in_dim, hidden, out_dim);
        return CMHAR_ERR_UNSUPPORTED;
    }
    return launch_mlp2_tc(reinterpret_cast<const uint8_t*>(x_img), in_dim, w0, b0, w1, b1, n, l2_normalize, y,
                          reinterpret_cast<uint8_t*>(y_img), (cudaStream_t)s);
}

int cmhar_fused_head_forward(const void* fusion_blob, const void* x1_img, int32_t in_dim1, const void* x2_img, int32_t in_dim2,
                             int64_t n, const void* head_blob, const void* maha_blob, float* fused_out, float* logits_out,
                             int64_t* pred_out, float* msp_out, float* energy_out, float* maha_out, cmhar_stream_t s) {
    CMHAR_REQUIRE(fusion_blob && x1_img && x2_img && head_blob, "cmhar_fused_head_forward: null argument");
    CMHAR_REQUIRE(maha_blob || !maha_out, "maha_out needs a maha blob");
    CMHAR_REQUIRE(((uintptr_t)x1_img & 1023) == 0 && ((uintptr_t)x2_img & 1023) == 0, "operand images must be 1024-byte aligned");
    if (n <= 0) return CMHAR_OK;
    const float* fb = nullptr;
    const uint8_t* fw = linear_tc_section(fusion_blob, in_dim1 + in_dim2, D, &fb);
    BlobInfo hi{}, mi{};
    const bool want_maha = maha_blob && maha_out;
    const bool head_ok = lookup_blob(head_blob, &hi) && hi.magic == HEAD_MAGIC && hi.has_tc;
    const bool maha_ok = !want_maha || (lookup_blob(maha_blob, &mi) && mi.magic == MAHA_MAGIC && mi.has_tc);
    if (!fw || in_dim1 % 64 || in_dim2 % 64 || in_dim1 < 64 || in_dim2 < 64 || !head_ok || !maha_ok) {
        set_error("cmhar_fused_head_forward: needs a packed Linear(%d + %d -> 128) fusion blob, the reference head layout and (optionally) a <= 32-class Mahalanobis blob",
                  in_dim1, in_dim2);
        return CMHAR_ERR_UNSUPPORTED;
    }
    FwdArgs a{};
    a.head_blob = reinterpret_cast<const char*>(head_blob);
    a.maha_blob = reinterpret_cast<const char*>(maha_blob);
    a.n = n; a.xstride = D;
    a.logits_out = logits_out; a.pred_out = reinterpret_cast<long long*>(pred_out);
    a.msp_out = msp_out; a.energy_out = energy_out; a.maha_out = maha_out;
    HeadLayout hl{hi.a, hi.b, hi.c};
    const uint8_t* htc = reinterpret_cast<const uint8_t*>(head_blob) + tc_section_offset(hl.total());
    const float* hf32 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(head_blob) + sizeof(BlobHeader));
    const uint8_t* mtc = want_maha ? reinterpret_cast<const uint8_t*>(maha_blob) + tc_section_offset(MahaLayout{mi.a}.total()) : nullptr;
    HeadPreLayer pre{fw, fb, reinterpret_cast<const uint8_t*>(x1_img), reinterpret_cast<const uint8_t*>(x2_img), in_dim1 / 64, in_dim2 / 64, fused_out};
    return launch_head_forward_tc(a, htc, hf32, hl, mtc, (cudaStream_t)s, &pre);
}

size_t cmhar_similarity_img_work_bytes(int64_t na, int64_t nb) {
    if (na <= 0 || nb <= 0) return 64;
    return similarity_img_work_bytes(na, nb);
}

int cmhar_similarity_img(const void* a_img, int64_t na, const void* const* b_imgs, int32_t n_parts, int64_t rows_per_part,
                         int64_t nb, int32_t dim, float sig_scale, float sig_bias, double out_scale, double* const* sum_dst,
                         int32_t n_dst, void* work, cmhar_stream_t s) {
    CMHAR_REQUIRE(a_img && b_imgs && sum_dst && work, "cmhar_similarity_img: null argument");
    CMHAR_REQUIRE(n_parts >= 1 && n_parts <= CMHAR_MAX_PEERS && n_dst >= 1 && n_dst <= CMHAR_MAX_PEERS,
                  "cmhar_similarity_img: n_parts / n_dst outside [1,%d]", CMHAR_MAX_PEERS);
    CMHAR_REQUIRE(dim >= 64 && dim % 64 == 0 && dim <= 256, "cmhar_similarity_img: dim %% 64 == 0 and dim <= 256 required (got %d)", dim);
    CMHAR_REQUIRE(n_parts == 1 || (rows_per_part > 0 && rows_per_part % 128 == 0 && (int64_t)n_parts * rows_per_part >= nb),
                  "cmhar_similarity_img: partitioned B needs rows_per_part %% 128 == 0 covering nb");
    CMHAR_REQUIRE(((uintptr_t)a_img & 1023) == 0 && ((uintptr_t)work & 7) == 0, "cmhar_similarity_img: misaligned buffer");
    for (int i = 0; i < n_parts; ++i) CMHAR_REQUIRE(b_imgs[i] && ((uintptr_t)b_imgs[i] & 1023) == 0, "cmhar_similarity_img: bad B part %d", i);
    for (int i = 0; i < n_dst; ++i) CMHAR_REQUIRE(sum_dst[i] && ((uintptr_t)sum_dst[i] & 7) == 0, "cmhar_similarity_img: bad destination %d", i);
    if (na <= 0 || nb <= 0) return CMHAR_OK;
    return launch_similarity_img(reinterpret_cast<const uint8_t*>(a_img), na, reinterpret_cast<const uint8_t* const*>(b_imgs), n_parts,
                                 rows_per_part, nb, dim, 0, sig_scale, sig_bias, out_scale, sum_dst, n_dst, work, (cudaStream_t)s);
}

}  // extern "C"
