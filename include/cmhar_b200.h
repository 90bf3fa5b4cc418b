/*
 * cmhar_b200.h -- C ABI of the B200-native (sm_100a) cross-modal inference / OOD-scoring hot path.
 *
 * Drop-in boundary for YOUNESELBOUKNIFY/CrossModal-IMU-Video-OOD-HAR.  The reference has no FFI
 * (it is pure PyTorch); its "operator API" for this path is the eval-mode
 * forward of the nn.Modules in src/models/models.py, src/models/losses.py and the loop in
 * src/eval/evaluator.py.  Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are dense row-major unless a stride argument says otherwise;
 *   - no function allocates, frees or retains caller memory, none synchronises the device, all
 *     work is enqueued on the given stream (so calls can be captured into a CUDA graph);
 *   - return value: 0 = ok, negative = error; cmhar_last_error() returns a thread-local message;
 *   - weights are handed over once through a *_pack call that folds BatchNorm, transposes and
 *     (for the bf16 path) converts into a caller-owned blob; the blob is immutable afterwards
 *     and can be shared by concurrent streams (the reference's nn.DataParallel use,
 *     main.py:89-95, calls forward from several threads);
 *   - "precision": CMHAR_FP32 = fp32 CUDA-core arithmetic (1e-3 contract),
 *                  CMHAR_BF16 = bf16 tcgen05 tensor-core GEMMs with fp32 accumulation in TMEM
 *                               (2e-2 contract).
 *   - OOD scores: larger = more out-of-distribution.
 */
#ifndef CMHAR_B200_H
#define CMHAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMHAR_ABI_VERSION 3

#define CMHAR_OK               0
#define CMHAR_ERR_INVALID     -1   /* bad argument / unsupported shape                       */
#define CMHAR_ERR_CUDA        -2   /* a CUDA runtime call failed (message has the detail)    */
#define CMHAR_ERR_UNSUPPORTED -3   /* valid request this build cannot serve                  */

#define CMHAR_FP32 0
#define CMHAR_BF16 1

#define CMHAR_MAX_LAYERS 8
#define CMHAR_D_MODEL    128       /* configs/config.py:79  (kernels are specialised to it)  */
#define CMHAR_NHEAD      8         /* configs/config.py:80                                    */
#define CMHAR_FFN        512       /* src/models/models.py:88                                 */
#define CMHAR_PATCH      16        /* configs/config.py:77-78 (patch == stride)               */
#define CMHAR_MAX_SEQ    16        /* 1 + (250-16)/16 + 1 tokens survive, models.py:122-123   */
#define CMHAR_MAX_CLASSES 64
#define CMHAR_MAX_PEERS   8         /* ranks of one NVSwitch box        */
#define CMHAR_PEER_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t)                              */

typedef void* cmhar_stream_t;      /* cudaStream_t */

int         cmhar_abi_version(void);
const char* cmhar_last_error(void);
/* number of kernel launches this library has enqueued since load (all threads) */
int64_t     cmhar_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * IMU encoder  (replaces PatchEmbedding.forward + IMUEncoder.forward,
 *               reference src/models/models.py:30-50,100-132, and the
 *               nn.TransformerEncoderLayer stack it calls, torch/nn/modules/transformer.py:946-990)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const float *in_proj_weight;   /* (384,128) */
    const float *in_proj_bias;     /* (384)     */
    const float *out_proj_weight;  /* (128,128) */
    const float *out_proj_bias;    /* (128)     */
    const float *linear1_weight;   /* (512,128) */
    const float *linear1_bias;     /* (512)     */
    const float *linear2_weight;   /* (128,512) */
    const float *linear2_bias;     /* (128)     */
    const float *norm1_weight, *norm1_bias;   /* (128) */
    const float *norm2_weight, *norm2_bias;   /* (128) */
} cmhar_encoder_layer_params;

typedef struct {
    int32_t seq;                   /* tokens kept = min(1+C*N, N+1), 2..16                    */
    int32_t layers;                /* 1..CMHAR_MAX_LAYERS                                     */
    const float *cls_token;        /* (128)       models.py:75                                */
    const float *pos_encoding;     /* (>=seq,128) models.py:82                                */
    const float *patch_weight;     /* (128,16) of channel 0 -- the only live channel (F4)     */
    const float *patch_bias;       /* (128)                                                   */
    const float *norm_weight, *norm_bias;     /* final LayerNorm, models.py:98                */
    cmhar_encoder_layer_params layer[CMHAR_MAX_LAYERS];
} cmhar_imu_encoder_params;

/* Classifier head (replaces IMUClassifier.classifier, models.py:312-326: [Linear,BN,ReLU,Dropout]x2,
 * Linear).  BatchNorm is folded with its running statistics (eval mode). */
typedef struct {
    int32_t hidden1, hidden2, classes;        /* 256, 128, 32; each <= 256, classes <= 64     */
    const float *w0, *b0, *bn0_weight, *bn0_bias, *bn0_mean, *bn0_var;   /* (h1,128)          */
    const float *w1, *b1, *bn1_weight, *bn1_bias, *bn1_mean, *bn1_var;   /* (h2,h1)           */
    const float *w2, *b2;                                                 /* (classes,h2)      */
} cmhar_head_params;

size_t cmhar_imu_encoder_blob_bytes(int32_t seq, int32_t layers);
int    cmhar_imu_encoder_pack(const cmhar_imu_encoder_params* p, void* blob, cmhar_stream_t s);
size_t cmhar_head_blob_bytes(int32_t hidden1, int32_t hidden2, int32_t classes);
int    cmhar_head_pack(const cmhar_head_params* p, void* blob, cmhar_stream_t s);

/* Mahalanobis scorer state: whitening matrix and whitened class means (see cmhar_maha_*) */
size_t cmhar_maha_blob_bytes(int32_t classes);
int    cmhar_maha_pack(const float* whiten /*(128,128): dist=||f@whiten - mu_w||^2*/,
                       const float* mean_whitened /*(classes,128)*/,
                       const float* class_count /*(classes) or NULL; count<=0 => class skipped*/,
                       int32_t classes, void* blob, cmhar_stream_t s);

/* Windows -> encoder -> [head -> logits, arg-max, MSP, energy] [-> Mahalanobis]: the fused encoder launch,
 * followed (when a head / maha blob is given) by the head + scores launch on the CLS features it wrote.
 * Replaces IMUEncoder.forward / IMUClassifier.forward (models.py:100-132,328-339) and the per-batch
 * body of Evaluator.predict (src/eval/evaluator.py:44-45).
 *   x            channel-0 samples of window 0; window w starts at x + w*x_window_stride (floats);
 *                at least 16*(seq-1) contiguous samples are read per window.  Pass the (B,6,L)
 *                tensor's data pointer with x_window_stride = 6*L: channels 1..5 and the trailing
 *                samples are dead inputs of the reference and are never touched.
 *   head_blob    NULL => encoder only (logits/pred/msp/energy must be NULL)
 *   maha_blob    NULL => no Mahalanobis score (maha_out must be NULL)
 *   any output pointer may be NULL.  pred = first arg-max index (torch ``logits.max(1)``).
 */
int cmhar_imu_forward(const void* encoder_blob, const void* head_blob, const void* maha_blob,
                      const float* x, int64_t n_windows, int64_t x_window_stride,
                      float* cls_out        /* (n,128); required with head_blob / maha_out */,
                      float* tokens_out     /* (n,seq,128)  */,
                      float* logits_out     /* (n,classes)  */,
                      int64_t* pred_out     /* (n)          */,
                      float* msp_out        /* (n)  = -max softmax            */,
                      float* energy_out     /* (n)  = -logsumexp(logits)      */,
                      float* maha_out       /* (n)  = min_c Mahalanobis^2     */,
                      int32_t precision, cmhar_stream_t s);

/* Same launch(es) with the outputs in a struct, plus `cls_img`: the CLS features ALSO as a bf16 SWIZZLE_128B operand image
 * [ceil(n/128)][2][128 x 64] (cmhar_operand_image_bytes(n, 128)) -- the A operand cmhar_mlp2_forward_img and
 * cmhar_fused_head_forward stream with plain bulk copies, so the step's next kernels never read fp32 rows.
 * cls_img needs CMHAR_BF16 (CMHAR_ERR_UNSUPPORTED otherwise).  Rows past n_windows inside the last 8-window encoder tile
 * are written as zeros; rows past that are left untouched (allocate the image zeroed). */
typedef struct {
    float*   cls;      void* cls_img;
    float*   tokens;   float* logits;
    int64_t* pred;     float* msp;
    float*   energy;   float* maha;
} cmhar_imu_outputs;
int cmhar_imu_forward_ex(const void* encoder_blob, const void* head_blob, const void* maha_blob,
                         const float* x, int64_t n_windows, int64_t x_window_stride,
                         const cmhar_imu_outputs* out, int32_t precision, cmhar_stream_t s);

/* Diagnostic hook of the bf16 tcgen05 path (used by tests/tools only): runs the encoder on
 * n_windows and dumps the fp32 residual stream held in TMEM, (ceil(n/8)*128, 128) floats, right
 * after `stage`: 0 = patch embedding, 1 = layer-0 attention + LayerNorm1, 2 = layer-0 output. */
int cmhar_debug_imu_bf16(const void* encoder_blob, const float* x, int64_t n_windows,
                         int64_t x_window_stride, int32_t stage, float* residual_dump,
                         float* cls_out, int32_t* progress_host_mapped /* NULL or pinned [grid][16] */,
                         cmhar_stream_t s);

/* Diagnostic hook (tools only): installs (or, with NULL, removes) a device buffer of 2 + 4 * capacity_records
 * uint64 words; while installed, thread 0 of every CTA of the encoder / pooling / head / dense / similarity kernels
 * appends {kernel id, SM id, start ns, end ns} (globaltimer).  Word 0 counts the records.  Synchronous. */
int cmhar_debug_cta_trace(uint64_t* device_buffer, int64_t capacity_records);

/* Packed blobs are caller-owned memory; the library keeps a host-side record of every blob it packed (dimensions, whether the
 * tensor-core section exists) keyed by the blob pointer.  Call this BEFORE freeing or reusing a blob's memory so that a later
 * allocation at the same address is not mistaken for it.  Unknown pointers are ignored. */
int cmhar_blob_release(const void* blob);

/* Development switch (tools / A-B measurements only; never read from the environment, so a stray variable cannot change
 * results): key "enc_kernel" = 0 / 1 the single-tile tcgen05 kernel (default), 2 the two-tiles-in-flight kernel
 * imu_forward_bf16_pair_kernel (bit-identical results); key "dev_env" = 1 lets the launchers read their CMHAR_*
 * development variables (CMHAR_ABLATE, CMHAR_POOL_*, CMHAR_EPI_WARPS, ...) from the environment -- they are ignored
 * otherwise.  Process wide.
 * Returns CMHAR_ERR_INVALID for an unknown key / value. */
int cmhar_debug_set_option(const char* key, int32_t value);

/* Same head + scores from stored features (row-major (n,128) fp32).
 * CMHAR_FP32: fp32 FMA arithmetic.  CMHAR_BF16: the tensor-core kernel -- every layer, the whitening and the
 * class-mean products are tcgen05 MMAs on split-bf16 operands (x = hi + lo, three products per term, fp32
 * accumulation in TMEM: ~2^-17 relative, fp32-grade logits) -- for the reference head layout (256, 128, <= 32
 * classes); other layouts, or blobs copied after packing, run the fp32 CUDA-core kernel (more accurate, ~10x slower):
 * cmhar_head_kernel_kind tells which one a call will launch, so the choice is never hidden from the caller. */
int cmhar_head_forward(const void* head_blob, const void* maha_blob, const float* feat, int64_t n,
                       float* logits_out, int64_t* pred_out, float* msp_out, float* energy_out,
                       float* maha_out, int32_t precision, cmhar_stream_t s);

/* Which kernel cmhar_head_forward / the head stage of cmhar_imu_forward launches for these blobs at this precision:
 * 1 = tcgen05 split-bf16 kernel (head_tc_kernel), 0 = fp32 CUDA-core kernel (head_scores_kernel). */
int cmhar_head_kernel_kind(const void* head_blob, const void* maha_blob, int32_t precision);

/* Late-fusion classifier in ONE launch (spec row A6: fusion.py LateFusionClassifier; no reference implementation):
 *   fused = relu(BN([x1 | x2] Wf^T + bf))  -- bf16 tcgen05 MMAs on the two inputs' operand images, never concatenated --
 *   then the classifier head and the scores exactly as cmhar_head_forward(CMHAR_BF16) computes them (split-bf16, fp32-grade)
 *   on `fused`, which stays on chip (fused_out, optional, receives its fp32 rows).
 * fusion_blob: cmhar_linear_pack of Linear(in_dim1 + in_dim2 -> 128) + BatchNorm; in_dim1, in_dim2 multiples of 64;
 * head_blob of the reference layout (256, 128, <= 32 classes), maha_blob optional.  CMHAR_ERR_UNSUPPORTED otherwise. */
int cmhar_fused_head_forward(const void* fusion_blob, const void* x1_img, int32_t in_dim1, const void* x2_img, int32_t in_dim2,
                             int64_t n, const void* head_blob, const void* maha_blob, float* fused_out, float* logits_out,
                             int64_t* pred_out, float* msp_out, float* energy_out, float* maha_out, cmhar_stream_t s);

/* MSP / energy from stored logits (spec rows A1, A2; no reference implementation). */
int cmhar_logit_scores(const float* logits, int64_t n, int32_t classes, float temperature,
                       int64_t* pred_out, float* msp_out, float* energy_out, cmhar_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Dense layers (replace nn.Linear / ProjectionHead.forward / F.normalize,
 *               reference src/models/models.py:213,226-234,288-289)
 * ------------------------------------------------------------------------------------------ */
/* blob for y = act(bn(x W^T + b)): folds BN (pass NULL bn_* for none) and transposes. */
size_t cmhar_linear_blob_bytes(int32_t in_dim, int32_t out_dim);
int    cmhar_linear_pack(const float* weight /*(out,in)*/, const float* bias /*(out) or NULL*/,
                         const float* bn_weight, const float* bn_bias, const float* bn_mean,
                         const float* bn_var, int32_t in_dim, int32_t out_dim, void* blob,
                         cmhar_stream_t s);
/* y (n,out) = relu?(x (n,in) @ W'^T + b');  in_dim % 4 == 0, out_dim % 4 == 0.
 * `work` (optional, cmhar_linear_work_bytes) lets small batches split the k loop over more CTAs;
 * partial sums are added in a fixed order, so results do not depend on scheduling.
 * CMHAR_BF16 and in_dim % 64 == 0: 128 x 128 output tiles on the tensor cores (tcgen05, bf16 operands -- the
 * activations converted on the fly, the weights from the bf16 images in the blob -- fp32 accumulation in TMEM). */
size_t cmhar_linear_work_bytes(int64_t n, int32_t out_dim);
int    cmhar_linear_forward(const void* blob, const float* x, int64_t n, int32_t in_dim,
                            int32_t out_dim, int32_t relu, float* y, void* work, size_t work_bytes,
                            int32_t precision, cmhar_stream_t s);
/* Late-fusion concat-MLP first layer (spec row A6, no reference implementation):
 * y = relu?([x1 | x2] @ W'^T + b') with W' (out, in_dim1 + in_dim2); the concatenation is never materialised. */
int    cmhar_concat_linear_forward(const void* blob, const float* x1, int32_t in_dim1, const float* x2,
                                   int32_t in_dim2, int64_t n, int32_t out_dim, int32_t relu, float* y,
                                   void* work, size_t work_bytes, int32_t precision, cmhar_stream_t s);
/* Tensor-core dense layer (CMHAR_BF16 tiles) with operand images on either side: a chain of layers hands its
 * activations over as bf16 SWIZZLE_128B chunk images [ceil(n/128)][dim/64][128 x 64] (16 KiB each, the layout
 * cmhar_similarity builds for its operands) instead of fp32 rows, so the consumer's A operand is a plain
 * cp.async.bulk copy -- no staging warps, no per-chunk L2 round trip (reference ProjectionHead.forward,
 * src/models/models.py:226-234: Linear -> BN -> ReLU -> Linear).  x / x_img: exactly one input form is used (x_img
 * wins); y / y_img: either or both outputs.  Bit-identical to cmhar_linear_forward(CMHAR_BF16): both round the
 * activation to bf16 (RN) before the MMA. */
size_t cmhar_operand_image_bytes(int64_t n, int32_t dim);
int    cmhar_linear_forward_img(const void* blob, const float* x, const void* x_img, int64_t n, int32_t in_dim,
                                int32_t out_dim, int32_t relu, float* y, void* y_img, cmhar_stream_t s);

/* A whole projection head in ONE launch (reference ProjectionHead.forward, src/models/models.py:226-234, and the
 * F.normalize(dim=1) that follows it in CrossModalModel.forward, :288-289):
 *   y = normalize?(relu(BN(x W0^T + b0)) W1^T + b1), x as a bf16 operand image, the 512-wide hidden activation kept in
 *   tensor memory, y as fp32 rows (n, 256) and / or as the bf16 operand image cmhar_similarity_img streams.
 * blob0 / blob1: cmhar_linear_pack blobs of Linear(in_dim -> 512)+BN and Linear(512 -> 256); in_dim % 64 == 0.
 * Other dimensions: CMHAR_ERR_UNSUPPORTED (chain cmhar_linear_forward_img + cmhar_l2_normalize instead). */
int    cmhar_mlp2_forward_img(const void* blob0, const void* blob1, const void* x_img, int64_t n, int32_t in_dim,
                              int32_t hidden, int32_t out_dim, int32_t l2_normalize, float* y, void* y_img, cmhar_stream_t s);

/* rows x / max(||x||_2, 1e-12)   (F.normalize(dim=1), models.py:288-289); in place allowed */
int    cmhar_l2_normalize(const float* x, int64_t n, int32_t dim, float* y, cmhar_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Cross-attention fusion block (spec row A6 -- the reference has no fusion block; defined in
 * fusion.py, oracle oracle/fusion_spec.py: self-consistency, not reference parity)
 * ------------------------------------------------------------------------------------------ */
/* The whole cross-attention fusion block as ONE tensor-core launch (CMHAR_BF16; spec row A6 -- not in the reference):
 *   q = tokens Wq^T + bq, [k | v] = frame tokens [Wk | Wv]^T, 8-head attention of the `seq` IMU tokens of a window over its 16 frame
 *   tokens, y = LayerNorm(tokens + attn Wo^T + bo), fused = mean over the tokens -> (n, 128).
 * Weights are packed once into a caller-owned blob (1 KiB aligned): wq (128,128), bq (128), wk / wv (128, frame_dim) -- the caller may
 * have folded a preceding Linear into them --, wo (128,128), bo_folded = bo + Wo bv (the value bias commutes with the attention
 * average; the key bias cancels in the softmax), LayerNorm weight / bias; all fp32 device pointers.  tokens: (n, seq, 128) fp32;
 * frame_img: bf16 operand image of n * 16 rows x frame_dim (cmhar_video_pool_frames_img).  Returns CMHAR_ERR_UNSUPPORTED for other
 * shapes (frames != 16, seq > 16, frame_dim % 64 != 0): run the chained route (cmhar_linear_forward, cmhar_cross_attention,
 * cmhar_residual_ln_pool) then. */
size_t cmhar_xattn_blob_bytes(int32_t frame_dim);
int cmhar_xattn_pack(const float* wq, const float* bq, const float* wk, const float* wv, const float* wo, const float* bo_folded,
                     const float* ln_weight, const float* ln_bias, int32_t frame_dim, void* blob, cmhar_stream_t s);
int cmhar_xattn_forward(const void* blob, const float* tokens, const void* frame_img, int64_t n, int32_t seq, int32_t frames,
                        int32_t frame_dim, float ln_eps, float* fused_out, cmhar_stream_t s);

/* q (n*s_len,128) = projected IMU tokens, kv (n*t_len,256) = [K | V] projected frame tokens ->
 * out (n*s_len,128) = concat_h softmax(q_h k_h^T / 4) v_h, 8 heads of 16.  s_len <= 16, t_len <= 32. */
int cmhar_cross_attention(const float* q, const float* kv, int64_t n, int32_t s_len, int32_t t_len,
                          float* out, cmhar_stream_t s);
/* pooled (n,128) = mean_s LayerNorm(x[n,s,:] + a[n,s,:]; gamma, beta, eps) */
int cmhar_residual_ln_pool(const float* x, const float* a, const float* gamma, const float* beta,
                           int64_t n, int32_t s_len, float eps, float* pooled, cmhar_stream_t s);

/* ------------------------------------------------------------------------------------------
 * 1-D conv / BatchNorm / ReLU IMU encoder (north-star item 1; spec row A6 -- NOT in the reference, whose encoder is
 * the transformer above; defined in conv_encoder.py, oracle oracle/fusion_spec.py: self-consistency only)
 *   Conv1d(6->32,k5,s1,p2) BN ReLU, Conv1d(32->64,k5,s2,p2) BN ReLU, Conv1d(64->128,k5,s2,p2) BN ReLU, time mean
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const float *weight;            /* (c_out, c_in, 5) */
    const float *bias;              /* (c_out) or NULL  */
    const float *bn_weight, *bn_bias, *bn_mean, *bn_var;   /* (c_out) each, or all NULL */
} cmhar_conv_layer_params;
typedef struct { cmhar_conv_layer_params layer[3]; } cmhar_conv_encoder_params;
size_t cmhar_conv_encoder_blob_bytes(void);
int    cmhar_conv_encoder_pack(const cmhar_conv_encoder_params* p, void* blob, cmhar_stream_t s);
/* x: window w at x + w*x_window_stride, (6, window) row-major; feat_out (n,128) */
int    cmhar_conv_encoder_forward(const void* blob, const float* x, int64_t n, int32_t window,
                                  int64_t x_window_stride, float* feat_out, cmhar_stream_t s);
/* Same with an explicit precision: CMHAR_FP32 = the CUDA-core kernel above (fp32 FMA); CMHAR_BF16 = the tensor-core kernel
 * (conv_encoder_tc.cu): every layer an implicit GEMM (tcgen05.mma, bf16 operands, fp32 accumulation in TMEM) over im2col tiles
 * that exist only in shared memory, BatchNorm scale folded into the weight images, split-precision (hi + lo) input samples. */
int    cmhar_conv_encoder_forward_ex(const void* blob, const float* x, int64_t n, int32_t window,
                                     int64_t x_window_stride, float* feat_out, int32_t precision, cmhar_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Video tail (replaces VideoEncoder.forward after the trunk, models.py:210-216:
 *             adaptive_avg_pool2d -> per-frame Linear -> temporal mean == Linear(mean_{t,h,w}))
 * ------------------------------------------------------------------------------------------ */
/* fmap (n*frames, channels, hw) contiguous, dtype bf16 (fmap_is_bf16=1) or fp32 ->
 * pooled (n, channels) fp32 = mean over frames and hw.  The projection is cmhar_linear_forward. */
int cmhar_video_pool(const void* fmap, int32_t fmap_is_bf16, int64_t n, int32_t frames,
                     int32_t channels, int32_t hw, float* pooled, cmhar_stream_t s);

/* Same reduction, with the result (also) written as a bf16 operand image [ceil(n/128)][channels/64][128 x 64] for
 * cmhar_linear_forward_img (the projection layer then needs no fp32 staging); pooled may be NULL. channels % 64 == 0. */
int cmhar_video_pool_img(const void* fmap, int32_t fmap_is_bf16, int64_t n, int32_t frames, int32_t channels,
                         int32_t hw, float* pooled, void* pooled_img, cmhar_stream_t s);
/* Same single pass over the feature maps, additionally emitting the per-FRAME spatial means (reference
 * src/models/models.py:210-211 before the temporal mean of :215) as a bf16 operand image of n * frames rows x channels
 * (row = clip * frames + frame): the frame tokens of the cross-attention fusion block (fusion.py).  pooled / pooled_img
 * (the clip means) may be NULL. */
int cmhar_video_pool_frames_img(const void* fmap, int32_t fmap_is_bf16, int64_t n, int32_t frames, int32_t channels,
                                int32_t hw, float* pooled, void* pooled_img, void* frame_img, cmhar_stream_t s);

/* Same reduction by the co-resident kernel: ONE 128-thread CTA per SM streams [128 channels x hw] slabs through a
 * 32 KiB cp.async.bulk ring (<= 40 registers per thread), small enough to be resident next to an encoder CTA, so
 * the HBM-bound pooling of one batch overlaps the tensor-bound encoder of others on the same SMs.  Sustains about a
 * quarter of the HBM bandwidth on its own -- use cmhar_video_pool when the stage has the GPU to itself.
 * Requires a 16-byte aligned map, channels % 128 == 0 and 16..64 bytes per channel (CMHAR_ERR_INVALID otherwise). */
int cmhar_video_pool_coresident(const void* fmap, int32_t is_bf16, int64_t n, int32_t frames, int32_t channels,
                                int32_t hw, float* pooled, cmhar_stream_t s);

/* ---- video trunk on the device: the two kernels either side of the channels-last bf16 trunk -------------- */
/* ToTensor + Normalize of the reference's per-frame transform (src/data/datasets.py:52-58: x/255, (x - mean) / std) on the
 * device: frames_u8 = n_pixels interleaved RGB pixels (decoded HWC frames, any number of frames back to back) ->
 * out_bf16 = n_pixels x cpad bf16 (channels-last pixels; cpad 3, 4 or 8, channels >= 3 are zero so that the first
 * convolution can run with padded input channels).  mean3 / std3 are HOST pointers to three floats.  Same fp32
 * operation order as torch, one rounding to bf16. */
int cmhar_frames_normalize(const uint8_t* frames_u8, int64_t n_pixels, const float* mean3, const float* std3, int32_t cpad,
                           void* out_bf16, cmhar_stream_t s);
/* cmhar_video_pool_frames_img for a CHANNELS-LAST feature map -- physical layout (n*frames, hw, channels), what a
 * channels-last trunk writes (reference models.py:209 output, permuted) -- so that no NCHW copy of the map is made:
 * pooled (n, channels) fp32, pooled_img / frame_img bf16 operand images; any of the three may be NULL (not all).
 * Needs a 16-byte aligned map and channels % 8 == 0 (bf16) / % 4 == 0 (fp32); images need channels % 64 == 0. */
int cmhar_video_pool_nhwc(const void* fmap, int32_t fmap_is_bf16, int64_t n, int32_t frames, int32_t channels, int32_t hw,
                          float* pooled, void* pooled_img, void* frame_img, cmhar_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Contrastive similarity (replaces SigmoidContrastiveLoss.forward / InfoNCELoss.forward,
 *                         reference src/models/losses.py:25-54,67-87)
 * ------------------------------------------------------------------------------------------ */
/* a (na,dim), b (nb,dim) fp32.  sim_out (na,nb) = a b^T, optional (NULL = never materialised).
 * sigmoid_sum_out[0] += sum_ij softplus(-(sim*scale + bias))   (double; caller zeroes it)
 * row_lse_out (na) / col_lse_out (nb): logsumexp over the row / column of sim*lse_scale; the
 * column reduction is accumulated as (max, sumexp) pairs in col_work (2*nb floats per row-tile
 * group, see cmhar_similarity_work_bytes) and finished by the same call.
 * diag_out (min(na,nb)) = sim_ii * lse_scale.   Any of the outputs may be NULL.
 * precision CMHAR_BF16 with a workspace and dim % 64 == 0, dim <= 256: the 128x128 tiles run on the tensor
 * cores (tcgen05, bf16 operands converted once into the workspace, fp32 accumulation in TMEM; column
 * statistics come from a second pass over the transposed tiles).  Otherwise fp32 CUDA-core tiles. */
size_t cmhar_similarity_work_bytes(int64_t na, int64_t nb, int32_t dim);
int cmhar_similarity(const float* a, const float* b, int64_t na, int64_t nb, int32_t dim,
                     int64_t diag_offset /* column index of row 0's positive (sharded rows) */,
                     float* sim_out, float sig_scale, float sig_bias, double* sigmoid_sum_out,
                     float lse_scale, float* row_lse_out, float* col_lse_out, float* diag_out,
                     void* work, int32_t precision, cmhar_stream_t s);

/* Sigmoid contrastive loss (losses.py:37-52) from operands that ALREADY are bf16 operand images (what
 * cmhar_mlp2_forward_img writes), with the B operand optionally partitioned over `n_parts` buffers of `rows_per_part` rows
 * each (rows_per_part % 128 == 0 when n_parts > 1): the ranks' shards of the video embeddings.  Pointers of other ranks
 * (cmhar_peer_open) are read over NVLink by the same cp.async.bulk ring that feeds the MMAs -- the all-gather of
 * SURVEY.md section 8e row 3 happens inside the GEMM, tile by tile, and no gathered copy is ever written.
 *   result = out_scale * sum_ij softplus(-(a_i . b_j * sig_scale + sig_bias))   over this call's (na x nb) block,
 * reduced deterministically (per-CTA partials added in index order by the last CTA; no zero-initialised accumulator, no
 * atomics on the value) and stored to every pointer of sum_dst (n_dst >= 1; a rank's slot in every rank's slot array).
 * b_imgs / sum_dst are HOST arrays of device pointers.  work: cmhar_similarity_img_work_bytes, zeroed once by the caller.
 * dim % 64 == 0, dim <= 256. */
size_t cmhar_similarity_img_work_bytes(int64_t na, int64_t nb);
int cmhar_similarity_img(const void* a_img, int64_t na, const void* const* b_imgs, int32_t n_parts, int64_t rows_per_part,
                         int64_t nb, int32_t dim, float sig_scale, float sig_bias, double out_scale,
                         double* const* sum_dst, int32_t n_dst, void* work, cmhar_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Peer-GPU memory (one process per GPU).  The ONLY entry points that allocate: a buffer other ranks
 * read over NVLink must come from cudaMalloc (CUDA IPC), not from a caching allocator's sub-allocation.
 * ------------------------------------------------------------------------------------------ */
int cmhar_peer_alloc(size_t bytes, void** ptr_out);                 /* zero-filled; synchronous                         */
int cmhar_peer_free(void* ptr);
int cmhar_peer_export(const void* ptr, void* handle_out /* CMHAR_PEER_HANDLE_BYTES */);
int cmhar_peer_open(const void* handle, void** ptr_out);            /* maps another rank's buffer into this process      */
int cmhar_peer_close(void* ptr);
/* Barrier across the ranks of one box, enqueued on the stream: flag_blocks (HOST array of `world` device pointers, [rank] = the
 * local block, the others peer-mapped; CMHAR_MAX_PEERS uint64 each, zeroed at allocation), local_epoch = one uint64 of local
 * device memory (zeroed).  Optionally finishes a reduction after the barrier: *sum_out = scale * sum(slots[0..n_slots)). */
int cmhar_peer_barrier(void* const* flag_blocks, int32_t rank, int32_t world, void* local_epoch, const double* slots,
                       int32_t n_slots, double scale, double* sum_out, cmhar_stream_t s);

/* ------------------------------------------------------------------------------------------
 * Mahalanobis OOD (spec rows A3/A4 -- no reference implementation)
 * ------------------------------------------------------------------------------------------ */
/* Accumulates (+=) sufficient statistics of feat (n,128) fp32 with int64 labels:
 * count (classes) double, sum (classes,128) double, second (128,128) double = sum f f^T.
 * Rows whose label is outside [0,classes) are skipped.  These three buffers are what the
 * multi-GPU fit all-reduces over NCCL.  CMHAR_BF16: both reductions run as split-bf16 tcgen05 GEMMs over the row
 * dimension (all four hi/lo products for the second moment, fp32 accumulation in TMEM flushed in fp64);
 * CMHAR_FP32: fp32 FMA register tiles flushed into fp64. */
int cmhar_maha_accumulate(const float* feat, const int64_t* labels, int64_t n, int32_t classes,
                          double* count, double* sum, double* second, int32_t precision, cmhar_stream_t s);
/* Finalisation of the fit ON THE DEVICE, stream-ordered behind the accumulate calls and the NCCL all-reduce (no host round trip):
 * stats = [count (classes) | sum (classes,128) | second (128,128)] doubles, contiguous, as accumulated above ->
 *   fit64 (cmhar_maha_fit64_doubles(classes) doubles) = [mean (classes,128) | cov (128,128) | whiten (128,128) | mean_whitened (classes,128)]
 *   with cov = (second - sum_c n_c mu_c mu_c^T) / N, symmetrised, + ridge I; cov = G G^T (Cholesky); whiten = G^-T
 *   whiten_f32 / mean_w_f32 / count_f32: the fp32 inputs of cmhar_maha_pack, which may follow on the same stream
 *   info (device int32): 0 = ok, k > 0 = pivot k of the Cholesky factorisation is not positive, -1 = no labelled rows.
 * One fp64 CTA (~0.1 ms); classes <= 64.  Same algebra as the host finalisation it replaces (ood.finalize_mahalanobis, spec oracle
 * oracle/ood_spec.py); no reference implementation exists. */
size_t cmhar_maha_fit64_doubles(int32_t classes);
int cmhar_maha_finalize(const double* stats, int32_t classes, double ridge, double* fit64, float* whiten_f32,
                        float* mean_w_f32, float* count_f32, int32_t* info, cmhar_stream_t s);
/* score (n) = min_c || feat@whiten - mean_whitened_c ||^2.  CMHAR_BF16: whitening and class-mean products as
 * split-bf16 tcgen05 MMAs (fp32-grade, see cmhar_head_forward) when classes <= 32; CMHAR_FP32: fp32 FMA. */
int cmhar_maha_score(const void* maha_blob, const float* feat, int64_t n, float* score, int32_t precision,
                     cmhar_stream_t s);

/* ------------------------------------------------------------------------------------------
 * AUROC / FPR95 (spec row A5): order-preserving histograms of float scores
 * ------------------------------------------------------------------------------------------ */
/* Rows whose arg-max a bounded logit error could change: top-2 margin < rel_tau * max |logit| (or a non-finite logit).
 * Writes their indices (any order) to idx_out (capacity n) and their number to work2[1] (work2[0] = max |logit| as float
 * bits); work2 = 2 uint32 of device memory.  The classifiers' "bf16_refined" precision re-runs exactly those windows on the
 * fp32 path, which restores the reference's predicted labels (src/eval/evaluator.py:44-45) at bf16 throughput. */
int cmhar_near_tie_rows(const float* logits, int64_t n, int32_t classes, float rel_tau, uint32_t* work2, int64_t* idx_out,
                        cmhar_stream_t s);

/* key(score) = monotone uint32 image of the float; bin = (key - key_lo) >> shift, clamped to
 * [0,bins).  hist (bins) uint64 is accumulated (+=).  min/max keys via cmhar_score_key_range. */
int cmhar_score_key_range(const float* scores, int64_t n, uint32_t* key_min_max /*[2], caller
                          initialises to {0xffffffff,0}*/, cmhar_stream_t s);
int cmhar_score_histogram(const float* scores, int64_t n, uint32_t key_lo, int32_t shift,
                          int32_t bins, unsigned long long* hist, cmhar_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* CMHAR_B200_H */
