/* qavit_b200 -- C ABI of the B200-native QA-ViT / HQA-ViT hot path (libqavit_b200.so).
 *
 * The reference has no native code and no FFI: its boundary is the Python nn.Module tree
 * (SURVEY.md section 8b).  Each entry point below replaces the ATen op sequence behind one reference
 * module's forward (and the autograd graph behind its backward); the reference file:line it stands in
 * for is cited per function (H = HQAViT_CIFAR100.py).  qa-vit_b200/modules.py binds them with ctypes
 * from torch.autograd.Functions; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless said otherwise;
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*), nothing synchronises,
 *     nothing allocates: activations to keep for backward live in `saved`, temporaries in `scratch`,
 *     both sized by qavit_block_workspace();
 *   - return 0 on success, non-zero on error with the message in qavit_last_error();
 *   - dtype 0 = fp32 run (fp32 SIMT GEMMs; the 1e-4 parity mode), 1 = bf16 run (tcgen05 GEMMs with bf16
 *     operands / fp32 accumulation, bf16 activations, fp32 residual stream, norms, softmax and gradients
 *     of parameters -- the reference's autocast(bfloat16) recipe).
 */
#ifndef QAVIT_B200_H
#define QAVIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QAVIT_ABI_VERSION 3

/* Index of every parameter tensor a (TokenLearner-wrapped) quad block reads.  qavit_block_param_name(i)
 * returns the reference state_dict suffix (relative to the wrapper prefix "stageS_blocks.I." for HQAViT,
 * to "blocks.I." for QAViT, to "global_bank." for the bank entries). */
enum {
  QP_NORM1_W, QP_NORM1_B,
  QP_SWA_QKV_W, QP_SWA_QKV_B, QP_SWA_EK, QP_SWA_EV, QP_SWA_PROJ_W, QP_SWA_PROJ_B, QP_SWA_NORM_W, QP_SWA_NORM_B,
  QP_MSDA_QKV_W, QP_MSDA_QKV_B, QP_MSDA_EK, QP_MSDA_EV, QP_MSDA_PROJ_W, QP_MSDA_PROJ_B, QP_MSDA_NORM_W, QP_MSDA_NORM_B,
  QP_CGA_Q_W, QP_CGA_Q_B, QP_CGA_K_W, QP_CGA_K_B, QP_CGA_V_W, QP_CGA_V_B, QP_CGA_BK_W, QP_CGA_BK_B, QP_CGA_BV_W,
  QP_CGA_BV_B, QP_CGA_PROJ_W, QP_CGA_PROJ_B, QP_CGA_NORM_W, QP_CGA_NORM_B,
  QP_CROSS_Q_W, QP_CROSS_Q_B, QP_CROSS_K_W, QP_CROSS_K_B, QP_CROSS_V_W, QP_CROSS_V_B, QP_CROSS_PROJ_W, QP_CROSS_PROJ_B,
  QP_NSWA_W, QP_NSWA_B, QP_NMSDA_W, QP_NMSDA_B, QP_NCGA_W, QP_NCGA_B, QP_NCROSS_W, QP_NCROSS_B,
  QP_CSWA_W, QP_CSWA_B, QP_CMSDA_W, QP_CMSDA_B, QP_CCGA_W, QP_CCGA_B, QP_CCROSS_W, QP_CCROSS_B,
  QP_FUSION_W,
  QP_BMLP_FC1_W, QP_BMLP_FC1_B, QP_BMLP_FC2_W, QP_BMLP_FC2_B,
  QP_NORM2_W, QP_NORM2_B,
  QP_FFN_GAMMA, QP_FFN_FC1_W, QP_FFN_FC1_B, QP_FFN_DWN_W, QP_FFN_DWN_B, QP_FFN_SCALE, QP_FFN_DW_W, QP_FFN_DW_B,
  QP_FFN_PDN_W, QP_FFN_PDN_B, QP_FFN_FC2_W, QP_FFN_FC2_B,
  QP_TL_LN_W, QP_TL_LN_B, QP_TL_FC_W, QP_TL_FC_B, QP_UP_FC_W, QP_UP_FC_B, QP_UP_LN_W, QP_UP_LN_B,
  QP_BANK_K, QP_BANK_V, QP_BANK_WN_W, QP_BANK_WN_B, QP_BANK_WC_W, QP_BANK_WC_B, QP_BANK_WG_W, QP_BANK_WG_B,
  QP_COUNT
};

typedef struct qavit_block_cfg {
  int32_t batch;             /* images in this call                                                       */
  int32_t tokens;            /* tokens the quad block sees (learned tokens M when token_learner, else N)   */
  int32_t tokens_full;       /* N of the surrounding stream (== tokens when token_learner == 0)            */
  int32_t token_learner;     /* 1: TokenLearner -> block -> TokenUpMix wrapper (H:1104-1123)               */
  int32_t dim, heads, bank_size, groups, window, linformer_k, msda_seq_len;
  int32_t n_dilations, dilations[4], pool_stride;
  int32_t compress_dim, bottleneck_hidden, ffn_hidden;
  int32_t ffn_v1;            /* 1: QAViT.py CCFFFN (no norms / scale / gamma, QAViT.py:553-582)            */
  int32_t dwconv_bias;       /* depthwise conv carries a bias (QAViT.py, QAViTv2.py)                       */
  int32_t bank_v1;           /* 1: QAViT.py bank constants, no update counter (QAViT.py:203-224)           */
  int32_t train;             /* 1: bank writes happen (module.training)                                    */
  int32_t dtype;             /* 0 fp32, 1 bf16                                                             */
  float dropout;             /* config.dropout: SDPA dropout_p + the nn.Dropout sites of the block (train only; H:416-465, 648-656, 697-710) */
  float drop_path;           /* this block's DropPath rate (H:1066-1069, 1187), applied per image in train mode */
  int32_t tokens_out;        /* tokens TokenUpMix emits (its Linear's out_features, fixed at construction: H:1099-1103); 0 = tokens_full.
                                Differs from tokens_full when a model built for 32 x 32 is fed larger images (STL-10 recipe,
                                HQAViT_Tiny_stl10.py:250-282: the first block maps 576 -> 16 -> 64 tokens)                  */
} qavit_block_cfg;

const char* qavit_last_error(void);
int qavit_abi_version(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py reports the per-step delta) */
long long qavit_launch_count(void);
/* state_dict suffix of parameter `index` (NULL when out of range); *scope: 0 = quad block, 1 = wrapper, 2 = bank */
const char* qavit_block_param_name(int index, int* scope);

/* Bytes of `saved` (forward -> backward) and `scratch` (max of forward / backward temporaries). */
int qavit_block_workspace(const qavit_block_cfg* cfg, size_t* saved_bytes, size_t* scratch_bytes);

/* QuadAttentionBlock.forward (H:1071-1085) incl. the four branches (H:403-626), GlobalTokenBank.write
 * (H:296-321, mutates params[QP_BANK_K/V] and *update_count in train mode), HybridFusion, BottleneckMLP,
 * CCFFFN (H:632-712), and -- when cfg->token_learner -- TokenLearner / TokenUpMix (H:971-1031).
 *   params[QP_COUNT] : fp32 parameter tensors (reference layouts); entries not used by the config may be NULL
 *   x   [batch, tokens_full, dim] fp32      out [batch, tokens_out (or tokens_full), dim] fp32 */
int qavit_block_forward(const qavit_block_cfg* cfg, const void* const* params, long long* update_count,
                        unsigned long long* rng, const float* x, float* out, void* saved, void* scratch, void* stream);

/* rng: device {seed, offset} Philox state, required when train != 0 and (dropout > 0 or drop_path > 0), else may be
 * NULL.  Forward snapshots it into `saved` and advances the offset ON THE DEVICE (CUDA-graph replays draw fresh masks);
 * backward regenerates every mask from the snapshot -- no mask is stored. */
/* Backward of the above.  grads[QP_COUNT]: fp32 buffers the parameter gradients are ACCUMULATED into (zero them
 * first; NULL for the write_* / branch .norm parameters, which the reference never trains, H:315).
 *   dout [batch, tokens_out (or tokens_full), dim] fp32     dx [batch, tokens_full, dim] fp32 (overwritten) */
int qavit_block_backward(const qavit_block_cfg* cfg, const void* const* params, float* const* grads, const float* x,
                         const float* dout, float* dx, const void* saved, void* scratch, void* stream);

/* PatchEmbed.forward + pos_embed (H:1136-1138, 1250): im2col-free conv-as-GEMM -> LayerNorm -> + pos.
 *   img [B, Cin, S, S] fp32; W [d, Cin, p, p]; pre [B*N, d] and stats [B*N, 2] are kept for backward.
 *   dtype 1 (bf16 run) with `scratch` of qavit_patch_embed_scratch_bytes(): stride == kernel, so the patch rows are one
 *   gather of the image into a bf16 [B*N, Cin*p*p] matrix and the convolution / its weight gradient are tcgen05 GEMMs;
 *   dtype 0 (or scratch NULL): fp32 SIMT kernels (the 1e-4 parity mode). */
size_t qavit_patch_embed_scratch_bytes(int B, int Cin, int S, int p, int d);
int qavit_patch_embed_forward(const float* img, int B, int Cin, int S, int p, int d, const float* W, const float* bias,
                              const float* ln_w, const float* ln_b, const float* pos, float* pre, float* stats,
                              float* out, int dtype, void* scratch, void* stream);
int qavit_patch_embed_backward(const float* img, const float* dout, int B, int Cin, int S, int p, int d,
                               const float* pre, const float* stats, const float* ln_w, float* dpre_scratch, float* dW,
                               float* dbias, float* dln_w, float* dln_b, float* dpos, int dtype, void* scratch,
                               void* stream);

/* Final norm -> token mean -> head (H:1273-1275).  x [B, N, d] fp32 -> logits [B, classes]. */
int qavit_head_forward(const float* x, int B, int N, int d, const float* ln_w, const float* ln_b, const float* W,
                       const float* bias, int classes, float* stats, float* pooled, float* logits, void* stream);
int qavit_head_backward(const float* x, const float* dlogits, int B, int N, int d, const float* ln_w, const float* stats,
                        const float* pooled, const float* W, int classes, float* dpooled_scratch, float* dx,
                        float* dln_w, float* dln_b, float* dW, float* dbias, void* stream);

/* nn.Dropout on a contiguous fp32 tensor of n elements (n % 8 == 0; pos_drop, H:1155 / 1251): y = x * keep / (1 - p).
 * Forward writes the rng snapshot it used to snap[2] (device) and advances rng; backward applies the same mask to dy. */
int qavit_dropout_forward(const float* x, float* y, long long n, float p, unsigned long long* rng, unsigned long long* snap,
                          void* stream);
int qavit_dropout_backward(const float* dy, float* dx, long long n, float p, const unsigned long long* snap, void* stream);

/* CrossEntropyLoss(label_smoothing) with optional two-target mixup form (H:1373, 1404-1408).
 *   ya / yb: int64 class ids (yb may be NULL); loss: 1 float; dlogits may be NULL (forward only).
 *   lam_dev: optional DEVICE scalar overriding lam (the mixup weight of a CUDA-graph-replayed step).
 *   row_loss: B floats of scratch -- rows are summed in a fixed order, the loss is bitwise reproducible.
 *   err_flag: optional device int; bit 0 is raised when a label lies outside [0, classes) (torch asserts device-side
 *   there); such labels are clamped, no out-of-row read happens. */
int qavit_cross_entropy(const float* logits, const long long* ya, const long long* yb, float lam, const float* lam_dev, int B,
                        int classes, float label_smoothing, float* loss, float* dlogits, float* row_loss, int* err_flag,
                        void* stream);
/* y[i] = x[i] * (*scalar_dev) (chain-rule factor of a scalar loss); cudaMemsetAsync(p, 0, bytes) on the caller's stream. */
int qavit_scale_by_scalar(const float* x, const float* scalar_dev, long long n, float* y, void* stream);
int qavit_memset_zero(void* p, size_t bytes, void* stream);

/* Gradient clipping + AdamW on flat fp32 buffers (H:1413-1439; torch.optim.AdamW single-tensor rule).
 * The n_seg segments [seg_off[i], seg_off[i+1]) are the parameter tensors: seg_flags bit0 = has a gradient this
 * step (others are skipped entirely, also by weight decay), bit1 = per-parameter clip to per_param_max first
 * (names containing cnn_stem / dwconv).  Segments start on 4-float boundaries (padding belongs to the preceding
 * segment and stays zero); total_elems = seg_off[n_seg].  norms: n_seg + 2 floats; norms[n_seg] returns the global
 * norm before the global clip, norms[n_seg + 1] the global clip coefficient. */
int qavit_clip_grads(float* grads, const long long* seg_off, const int* seg_flags, int n_seg, float per_param_max,
                     float max_norm, float* norms, long long total_elems, void* stream);
/* Same with a factor `prescale` every gradient carries before clipping (1 / world_size when the buffer holds the SUM over
 * data-parallel ranks): norms and clip coefficients are those of prescale * g and the factor is folded into the one
 * scaling pass (the data-parallel mean costs no extra kernel). */
int qavit_clip_grads_scaled(float* grads, const long long* seg_off, const int* seg_flags, int n_seg, float per_param_max,
                            float max_norm, float prescale, float* norms, long long total_elems, void* stream);
/* transforms.ToTensor() + transforms.Normalize(mean, std) of a whole batch on the device (H:1300; HQAViT_C100_Finetune.py:100-116):
 * out[b, c, y, x] = (in / 255 - mean[c]) / std[c], one rounding per torch op (bit-identical to the torchvision pipeline);
 * hflip != 0 mirrors every image (RandomHorizontalFlip(p = 1), the second TTA view).  in_kind 0: uint8 [B, H, W, C], 1: uint8
 * [B, C, H, W], 2: float [B, C, H, W] already in [0, 1].  out: float [B, C, H, W], must not alias in. */
int qavit_normalize_images(const void* in, int in_kind, int B, int C, int H, int W, const float* mean, const float* stdv, int hflip,
                           float* out, void* stream);
/* y[i] = scale * x[i] (bank state averaged over data-parallel ranks: tail of the gradient buffer -> global_k / global_v). */
int qavit_scaled_copy(const float* x, float scale, long long n, float* y, void* stream);
int qavit_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const long long* seg_off,
                     const int* seg_flags, int n_seg, const float* hyper /* device: lr, beta1, beta2, eps, wd, bc1, bc2 */,
                     long long total_elems, void* stream);
/* Same step with ModelEMA.update (H:139-149) folded in: ema = d * ema + (1 - d) * p_new for every element of the flat
 * buffer (also segments AdamW skips); d = hyper[7], 1 - d = hyper[8] (rounded from double like torch's alpha).  SURVEY 8(f)-2. */
int qavit_adamw_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema,
                         const long long* seg_off, const int* seg_flags, int n_seg,
                         const float* hyper /* device: lr, beta1, beta2, eps, wd, bc1, bc2, ema_decay, 1 - ema_decay */,
                         long long total_elems, void* stream);
/* norms[i] = L2 norm of segment i of a flat buffer (0 for segments whose flag bit0 is clear) -- the per-tensor gradient /
 * parameter norms GradientMonitor.log_gradients collects (H:198-242) without its ~3 k host syncs. */
int qavit_segment_norms(const float* buf, const long long* seg_off, const int* seg_flags, int n_seg, float* norms,
                        void* stream);
/* On-device CutMix (mode 1) / MixUp (mode 2: lam * img + lam_b * img[perm], lam_b = 1 - lam rounded from double like the
 * torch expression) of img [B, C, H, W] fp32 against img[perm] (H:1379-1399); box = columns [x1, x2) x rows [y1, y2).
 * out must not alias img.  SURVEY 8(f)-3. */
int qavit_batch_mix(const float* img, float* out, const long long* perm, int B, int C, int H, int W, int mode, int x1, int y1,
                    int x2, int y2, float lam, float lam_b, void* stream);

/* nn.LayerNorm (eps configurable, C <= 256) for the modules around the blocks -- SplitFusion / LMFAdapter / RRCV /
 * ConvNeXtBlock norms (H:723, 816, 875, 922-939).  x fp32 or bf16 (x_bf16), y / dy fp32, dx in x's dtype,
 * stats [rows, 2] kept for backward, dgamma / dbeta accumulated. */
int qavit_layer_norm_forward(const void* x, int x_bf16, long long rows, int C, const float* w, const float* b, float eps,
                             float* y, float* stats, void* stream);
int qavit_layer_norm_backward(const void* x, int x_bf16, const float* dy, long long rows, int C, const float* w,
                              const float* stats, void* dx, float* dgamma, float* dbeta, void* stream);

/* Depthwise k x k convolution (k = 3 / 5 / 7, stride 1, same padding, bias) on channels-last maps [B, H, W, C]:
 * ConvNeXtBlock.dwconv H:722, LMFAdapter.dwconv_3x3 / dwconv_5x5 H:811-812 (lateral path, scope row f-1).
 * w [C, 1, k, k] fp32, x / y / dy / dx fp32 or bf16 (is_bf16); dw / dbias fp32, accumulated. dx may be NULL. */
int qavit_dwconv_forward(const void* x, int is_bf16, int B, int H, int W, int C, int K, const float* w, const float* bias,
                         void* y, void* stream);
int qavit_dwconv_backward(const void* x, const void* dy, int is_bf16, int B, int H, int W, int C, int K, const float* w,
                          void* dx, float* dw, float* dbias, void* stream);

/* nn.Linear / 1x1 convolution on row-major activations: y[M, N] = x[M, K] W^T + b and its backward (dW, db
 * accumulated; dx may be NULL).  x / y / dy / dx fp32 or bf16 (is_bf16); *_scratch: N*K bf16 for the converted weight. */
int qavit_linear_forward(const void* x, int is_bf16, long long M, int K, const float* W, const float* bias, int N, void* y,
                         void* wb_scratch, void* stream);
int qavit_linear_backward(const void* x, const void* dy, int is_bf16, long long M, int K, int N, const float* W, void* dx,
                          float* dW, float* db, void* wbt_scratch, void* stream);

/* ---- HQAViT lateral CNN path (scope row f-1): CNNStemModel H:742-793 -> LMFAdapter x3 H:799-849 -> RRCV x3 H:855-907.
 * One call per direction for the whole path: image in, the refined token maps R2 / R3 / R4 out.  Activations are
 * channels-last [B * H * W, C] in the run's activation type (dtype 0: fp32, 1: bf16); R2..R4 / dR2..dR4 are fp32
 * [B, grid * grid, dim] (LayerNorm outputs stay fp32 under autocast).
 * params[i] / grads[i] follow qavit_lateral_param_name(cfg, i) (state_dict names relative to the HQAViT module;
 * BatchNorm running_mean / running_var / num_batches_tracked entries are buffers: updated in train mode, grads NULL).
 * train = module.training (BatchNorm batch statistics).  The bilinear resize of H:824 is not implemented: the token
 * grid must equal img_size / 4 (all reference configurations). */
typedef struct qavit_lateral_cfg {
  int32_t batch, img_size, in_channels;
  int32_t c_stem, c2, c3, c4;        /* 32, cnn_c2, cnn_c3, cnn_c4 */
  int32_t rrcv_channels, rrcv_blocks;
  int32_t dim, grid;                 /* embed_dim, tokens per side */
  int32_t train, dtype;
  float bn_eps, bn_momentum;
  /* stem_kind 0: HQAViT_CIFAR100.py's stem (two 3x3 stride-2 conv + BatchNorm + GELU units, H:742-793).
   * stem_kind 1: HQAViTv2_CIFAR100.py's stem (V:753-833): 4x4 stride-4 patchify conv + LayerNorm([c2, g, g]), 2 / 3 / 2 ConvNeXt
   *   blocks with LayerScale at c2 / c3 / c4 channels, LayerNorm([C, g, g]) + 1x1 conv between the stages; c_stem is unused.
   *   stem_drop_path[j]: DropPath rate of ConvNeXt block j (train mode; the reference hard-codes 0, 0, 0, .1, .1, .1, .1); when any
   *   is > 0, rng = device {seed, offset} Philox state (snapshotted and advanced by forward, like qavit_splitfusion_forward). */
  int32_t stem_kind;
  float stem_drop_path[7];
  unsigned long long* rng;
} qavit_lateral_cfg;
int qavit_lateral_param_count(const qavit_lateral_cfg* cfg);
const char* qavit_lateral_param_name(const qavit_lateral_cfg* cfg, int index);
int qavit_lateral_workspace(const qavit_lateral_cfg* cfg, size_t* saved_bytes, size_t* scratch_bytes);
int qavit_lateral_forward(const qavit_lateral_cfg* cfg, const void* const* params, const float* img, float* R2, float* R3,
                          float* R4, void* saved, void* scratch, void* stream);
/* The same path in phases, so that a caller can overlap it with the token path stage by stage (R_i is first needed by fuse_i;
 * dR_i is known right after fuse_i's backward): `parts` selects the stem (feature maps kept in `saved`; its backward consumes the
 * three feature-map gradients, which the adapters' backward leaves in `saved`) and / or the LMFAdapter + RRCV chain of stage 2 / 3 / 4.
 * Forward: the stem part must run first (it also makes the bf16 weight copies of every part).  Backward: every adapter part must
 * have run (with dR_i == NULL if that stage took no part in the loss) before the stem part.  All parts of one forward / backward
 * share `saved`; calls that may run concurrently need distinct `scratch`. */
#define QAVIT_LATERAL_STEM 1u
#define QAVIT_LATERAL_ADAPTER2 2u
#define QAVIT_LATERAL_ADAPTER3 4u
#define QAVIT_LATERAL_ADAPTER4 8u
#define QAVIT_LATERAL_ALL 15u
int qavit_lateral_forward_parts(const qavit_lateral_cfg* cfg, const void* const* params, const float* img, float* R2, float* R3,
                                float* R4, void* saved, void* scratch, void* stream, unsigned parts);
int qavit_lateral_backward_parts(const qavit_lateral_cfg* cfg, const void* const* params, float* const* grads, const float* img,
                                 const float* dR2, const float* dR3, const float* dR4, const void* saved, void* scratch, void* stream,
                                 unsigned parts);
/* dR2..dR4 may be NULL (stage not fused); parameter gradients are accumulated; the image receives no gradient. */
int qavit_lateral_backward(const qavit_lateral_cfg* cfg, const void* const* params, float* const* grads, const float* img,
                           const float* dR2, const float* dR3, const float* dR4, const void* saved, void* scratch, void* stream);

/* ---- SplitFusion.forward H:945-963: T_in, R, out, dT, dR fp32 [rows, dim] (the residual stream).
 * params / grads follow qavit_splitfusion_param_name(i).  drop_p is cat_mlp[3].p (applied when train != 0; the mask is
 * regenerated in backward from the Philox state `rng` = {seed, offset} (device), which forward snapshots and advances). */
typedef struct qavit_splitfusion_cfg {
  long long rows;
  int32_t dim, dtype, train;
  float drop_p;
} qavit_splitfusion_cfg;
const char* qavit_splitfusion_param_name(int index);
int qavit_splitfusion_workspace(const qavit_splitfusion_cfg* cfg, size_t* saved_bytes, size_t* scratch_bytes);
int qavit_splitfusion_forward(const qavit_splitfusion_cfg* cfg, const void* const* params, unsigned long long* rng,
                              const float* T_in, const float* R, float* out, void* saved, void* scratch, void* stream);
int qavit_splitfusion_backward(const qavit_splitfusion_cfg* cfg, const void* const* params, float* const* grads, const float* T_in,
                               const float* R, const float* dout, float* dT, float* dR, const void* saved, void* scratch,
                               void* stream);

/* ---- Data-parallel gradient all-reduce (scope row D1; new: the reference is single-GPU).  One process per GPU; `comm` is the
 * caller's ncclComm_t (in a PyTorch process: ProcessGroupNCCL._comm_ptr()), `stream` the stream the reduction is ordered on.
 * The library resolves NCCL from the libnccl.so.2 already loaded in the process (it does not link against it).
 *   qavit_dp_available     : 1 when NCCL could be resolved
 *   qavit_dp_comm_info     : size / rank of the communicator
 *   qavit_dp_allreduce_sum : in-place SUM all-reduce of n_buckets fp32 device ranges as one NCCL group -- the buckets are contiguous
 *                            ranges of the flat gradient buffer (+ the GlobalTokenBank state riding in its tail); the 1 / world of the
 *                            mean is applied by qavit_clip_grads_scaled's grad_prescale, so no extra pass touches the gradients. */
int qavit_dp_available(void);
int qavit_dp_comm_info(void* comm, int* world, int* rank);
int qavit_dp_allreduce_sum(void* comm, float* const* bufs, const size_t* counts, int n_buckets, void* stream);

/* Test hook for the fused TokenLearner / TokenUpMix kernels of bf16 runs (H:971-1031; 16 learned tokens, <= 64 stream tokens,
 * 192 channels).  op 0 / 1 = TokenLearner forward / backward, 2 / 3 = TokenUpMix + LayerNorm forward / backward; `in` / `out` are
 * arrays of fp32 device pointers in the order documented next to the definition (csrc/block.cu). */
int qavit_test_tokens_fused(int op, int B, int N, int C, const float* const* in, float* const* out, void* stream);
/* Test hook for the fused per-branch LayerNorm + compress Linear + fusion scale kernels (H:1074-1079; d = 192, compress_dim = 48):
 * op 0 forward, 1 backward; pointer order documented next to the definition (csrc/block.cu). */
int qavit_test_cmp_fused(int op, long long R, const void* const* in, void* const* out, void* stream);

/* Test hook for the fused CCF-FFN mid-section of bf16 runs (H:704-709: GELU -> dwconv_norm -> depthwise 3x3 * scale ->
 * post_dwconv_norm on side x side token maps; side 4: C a multiple of 32 up to 128, side 8: C a multiple of 8 up to 128): op 0 forward, 1 backward; pointer order documented next to
 * the definition (csrc/block.cu). */
int qavit_test_ffn_mid(int op, int B, int side, int C, const void* const* in, void* const* out, void* stream);

/* Unit-test hooks for the GEMM flavours (bf16 tcgen05 and fp32 SIMT) behind the block. */
int qavit_test_gemm_nt(int use_tc, const void* A, int lda, int M, int N, int K, const float* W, const void* Wb,
                       const float* bias, void* C, int c_f32, void* stream);
/* tcgen05 NT GEMM with the fused epilogues: mode 0: C = A W^T + b | 1: C = pre-activation, C2 = gelu(pre) |
 * 2: C2 = aux + pre (aux: bf16 residual) | 3: C = pre * gelu'(aux) (aux: bf16 stored pre-activation).  All bf16 [M, N]. */
int qavit_test_gemm_epi(const void* A, int lda, int M, int N, int K, const void* Wb, const float* bias, void* C, void* C2,
                        int mode, const void* aux, void* stream);
int qavit_test_gemm_tn(int use_tc, const void* dY, int ldy, const void* X, int ldx, int M, int N, int K, float* dW,
                       float* db, void* stream);
int qavit_convert_weight(const float* w, int N, int K, void* wb, void* wbt, void* stream);

#ifdef __cplusplus
}
#endif
#endif
