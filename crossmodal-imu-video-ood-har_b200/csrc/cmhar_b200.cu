// cmhar_b200.cu -- single translation unit of libcmhar_b200.so (unity build: kernels defined in
// one file are launched from another without relocatable device code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC cmhar_b200.cu
#include "common.cuh"
namespace cmhar { struct FwdArgs; }
#include "pack.cu"
#include "imu_encoder_fp32.cu"
#include "imu_encoder_bf16.cu"
#include "imu_encoder_bf16_pair.cu"
#include "head.cu"
#include "head_tc.cu"
#include "dense.cu"
#include "linear_tc.cu"
#include "mlp2_tc.cu"
#include "similarity.cu"
#include "similarity_tc.cu"
#include "fusion.cu"
#include "conv_encoder.cu"
#include "conv_encoder_tc.cu"
#include "maha_score_tc.cu"
#include "ood.cu"
#include "maha_fit_tc.cu"
#include "peer.cu"
#include "api.cu"
