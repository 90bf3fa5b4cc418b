// imu_encoder_bf16.cu -- bf16 tcgen05/TMEM path of the fused IMU encoder (placeholder until the
// tensor-core megakernel lands; the call fails loudly instead of silently using another path).
#include "common.cuh"

namespace cmhar {

size_t encoder_bf16_bytes(int layers) { (void)layers; return 0; }

int pack_encoder_bf16(const cmhar_imu_encoder_params* p, const float* fp32_section, void* bf16_section, cudaStream_t st) {
    (void)p; (void)fp32_section; (void)bf16_section; (void)st;
    return CMHAR_OK;
}

int launch_imu_forward_bf16(const FwdArgs& a, cudaStream_t stream) {
    (void)a; (void)stream;
    set_error("cmhar_imu_forward: CMHAR_BF16 path not built in this version");
    return CMHAR_ERR_UNSUPPORTED;
}

}  // namespace cmhar
