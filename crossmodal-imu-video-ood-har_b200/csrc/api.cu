// api.cu -- C-ABI entry points of the fused IMU path (see include/cmhar_b200.h).
#include "common.cuh"

namespace cmhar {
int launch_imu_forward_fp32(const FwdArgs& a, cudaStream_t stream);
int launch_imu_forward_bf16(const FwdArgs& a, cudaStream_t stream);
int launch_head_forward(const FwdArgs& a, int precision, cudaStream_t stream);
int launch_imu_forward_bf16_debug(const FwdArgs& a, int stage, float* dump, int* progress, cudaStream_t stream);
}  // namespace cmhar

using namespace cmhar;

extern "C" {

int cmhar_imu_forward(const void* encoder_blob, const void* head_blob, const void* maha_blob, const float* x,
                      int64_t n_windows, int64_t x_window_stride, float* cls_out, float* tokens_out,
                      float* logits_out, int64_t* pred_out, float* msp_out, float* energy_out, float* maha_out,
                      int32_t precision, cmhar_stream_t s) {
    if (n_windows <= 0) return CMHAR_OK;
    CMHAR_REQUIRE(encoder_blob && x, "cmhar_imu_forward: null encoder blob or input");
    CMHAR_REQUIRE(((uintptr_t)encoder_blob & 1023) == 0, "encoder blob must be 1024-byte aligned");
    CMHAR_REQUIRE(head_blob || !(logits_out || pred_out || msp_out || energy_out),
                  "logits/pred/msp/energy outputs need a head blob");
    CMHAR_REQUIRE(maha_blob || !maha_out, "maha_out needs a maha blob");
    CMHAR_REQUIRE(cls_out || !(head_blob || maha_out),
                  "cls_out is required with a head / maha blob: the head runs as a second launch on the stored CLS features");
    CMHAR_REQUIRE(x_window_stride >= 16, "x_window_stride=%lld too small", (long long)x_window_stride);
    CMHAR_REQUIRE(precision == CMHAR_FP32 || precision == CMHAR_BF16, "bad precision %d", precision);
    if (n_windows <= 0) return CMHAR_OK;
    FwdArgs a{};
    a.enc_blob = reinterpret_cast<const char*>(encoder_blob);
    a.head_blob = reinterpret_cast<const char*>(head_blob);
    a.maha_blob = reinterpret_cast<const char*>(maha_blob);
    a.x = x; a.n = n_windows; a.xstride = x_window_stride;
    a.cls_out = cls_out; a.tokens_out = tokens_out; a.logits_out = logits_out;
    a.pred_out = reinterpret_cast<long long*>(pred_out);
    a.msp_out = msp_out; a.energy_out = energy_out; a.maha_out = maha_out;
    if (precision == CMHAR_BF16) return launch_imu_forward_bf16(a, (cudaStream_t)s);
    return launch_imu_forward_fp32(a, (cudaStream_t)s);
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

}  // extern "C"
