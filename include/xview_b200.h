/* xview_b200 - C ABI of the B200-native inference + fusion + score() hot path of
 * ethz-asl/modular_semantic_segmentation.
 *
 * This is the drop-in boundary: the reference reaches its compute through TensorFlow ops
 * called from xview/models/*.py; a maintainer replaces those call sites by ctypes bindings to
 * the functions below.  Each entry point cites the reference statement it
 * replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; xv_last_error() returns a
 *     thread-local message.  No C++ exception crosses this boundary.
 *   - pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - activations are NHWC float32 at the boundary (the reference layout/dtype); kernels are
 *     HWIO ([kh,kw,Cin,Cout]); transposed-conv kernels are [kh,kw,Cout,Cin].
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     stream-ordered and asynchronous unless stated otherwise.
 *   - label tensors are int64 (`label_bytes` = 8, the reference's tf.argmax dtype) or uint8
 *     (`label_bytes` = 1, the compact internal form).
 *   - one handle is used from one host thread at a time.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef XVIEW_B200_H_
#define XVIEW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XV_ABI_VERSION 1

/* ---- library / device ------------------------------------------------------------- */
int xv_abi_version(void);
const char* xv_last_error(void);
/* Binds the calling thread to `device`, queries SM count / shared-memory limits and resolves
 * the driver entry point used to build TMA descriptors.  (tf.Session creation,
 * xview/models/base_model.py:166-172) */
int xv_init(int device);
int xv_device_sm_count(int* out);

/* ---- plain memory / stream helpers (session feed/fetch, base_model.py:263-313) ----- */
int xv_malloc(void** out, size_t bytes);
int xv_free(void* p);
int xv_malloc_host(void** out_host, size_t bytes);       /* pinned */
int xv_free_host(void* p_host);
int xv_memcpy_h2d(void* dst, const void* src_host, size_t bytes, void* stream);
int xv_memcpy_d2h(void* dst_host, const void* src, size_t bytes, void* stream);
int xv_memset(void* dst, int value, size_t bytes, void* stream);
int xv_stream_sync(void* stream);

/* ---- input pipeline step (xview/datasets/data_baseclass.py:64-79: stack + astype('float32')) -- */
/* Raw sensor values -> float32 on the device, after the host->device copy.
 * src_dtype: 0 = uint8 (rgb), 1 = uint16 (depth), 2 = int16, 3 = int32. */
int xv_convert_to_f32(const void* src, int src_dtype, int64_t n, float* dst, void* stream);

/* ---- FCN expert (xview/models/simple_fcn.py:137-170 `fcn`, :10-87 encoder, :90-134 decoder) */
typedef struct xv_fcn xv_fcn;

#define XV_PRECISION_BF16 0 /* tcgen05 bf16 implicit-GEMM convolutions, fp32 accumulate   */
#define XV_PRECISION_FP32 1 /* validation mode: fp32 CUDA-core kernels, reference op order */

/* dropout sites, simple_fcn.py:51-53,61-63,72-78,124-126 */
#define XV_DROP_POOL3 1u    /* also enables pool4 dropout: reference quirk simple_fcn.py:61 */
#define XV_DROP_CONV4_3 2u
#define XV_DROP_CONV5_3 4u
#define XV_DROP_FEATURES 8u

typedef struct xv_dropout_cfg {
  float rate;                 /* tf.layers.dropout rate; survivors are scaled by 1/(1-rate)  */
  uint32_t sites;             /* OR of XV_DROP_*                                             */
  int32_t num_samples;        /* T >= 1 Monte-Carlo samples sharing one weight load          */
  uint64_t seed;              /* Philox key of the fused dropout masks                       */
  /* optional external keep-masks (uint8, one byte per element of the [T*N,h,w,c] activation):
   * order pool3, pool4, conv4_3, conv5_3, features.  NULL = fused Philox.                   */
  const uint8_t* ext_mask[5];
  /* XV_DROP_FLAG_*.  KEEP_FIRST: the dropout-free network rides along as an extra leading
   * sample that shares the trunk and every weight load (variance_mix.py:68-69 takes the
   * probabilities from a separate no-dropout pass): score / prob / label then describe that
   * dropout-free pass ([N,...]) and the moments are taken over the num_samples dropout passes;
   * external masks still cover the num_samples dropout passes only.                            */
  uint32_t flags;
} xv_dropout_cfg;
#define XV_DROP_FLAG_KEEP_FIRST 1u

typedef struct xv_fcn_outputs {
  /* T == 1 (or no dropout): per-image results; T > 1: per-sample results, batch = T*N        */
  float* score;               /* [B,H,W,C] pre-softmax class scores ('score', simple_fcn.py:133) */
  float* prob;                /* [B,H,W,C] softmax (basic_fusion_model.py:21)                 */
  int64_t* label_i64;         /* [B,H,W]   argmax (basic_fusion_model.py:22)                  */
  uint8_t* label_u8;          /* [B,H,W]   same, compact                                      */
  /* T > 1 only: moments over the T samples (variance_mix.py:62-66), no sample tensor is ever
   * written                                                                                 */
  float* mean_prob;           /* [N,H,W,C]                                                    */
  float* var_prob;            /* [N,H,W,C] population variance                                */
  float* mean_var;            /* [N,H,W]   mean over classes of var_prob                      */
} xv_fcn_outputs;

int xv_fcn_create(xv_fcn** out, int cin, int num_units, int num_classes, int batchnorm,
                  int precision);
/* Mid-level fusion network (fusion_fcn.py:11-40 over vgg16.py:7-51): one encoder-only handle per
 * modality (role XV_ROLE_ENCODER) plus one head handle (role XV_ROLE_HEAD, head_cin = 512 * number
 * of towers) that owns score_conv4/5 on the channel-concatenated conv4_3 / conv5_3, the bilinear
 * transposed convolutions and the final score layer. */
/* `batchnorm` of xv_fcn_create / xv_fcn_create_ex: 0 = none, 1 = every layer (simple_fcn.py
 * batchnorm=True), XV_BN_DECODER = only the decoder's "upscore" and "score" layers: fusion_fcn.py:39
 * calls decoder() with its default batchnorm=True (simple_fcn.py:91) although towers and heads
 * are built without batch norm. */
#define XV_BN_DECODER 2
#define XV_ROLE_EXPERT 0
#define XV_ROLE_ENCODER 1
#define XV_ROLE_HEAD 2
int xv_fcn_create_ex(xv_fcn** out, int cin, int num_units, int num_classes, int batchnorm,
                     int precision, int role, int head_cin);
/* Adapnet expert (xview/models/adapnet.py:99-173: ResNet-50 style blocks with atrous stage-2
 * convolutions, two batch-normalised transposed convolutions), test-time graph with batch norm on
 * its moving statistics - what the fusion models build with expert_model='adapnet'
 * (basic_fusion_model.py:13-16).  Returns the same handle type: parameters are set with
 * xv_fcn_set_param_host under the names below the prefix ("block_layer_4/stage_2/kernel",
 * ".../gamma", "first_deconvolution_upconv/kernel" [4,4,num_units,2048], ...), then
 * xv_fcn_finalize / xv_fcn_forward (drop = NULL; score, prob and label outputs) /
 * xv_fcn_get_layer_host ("block_0_1" .. "block_16", "shortcut", "merge", "score") /
 * xv_fcn_destroy.  cin <= 3 in bf16 precision. */
int xv_adapnet_create(xv_fcn** out, int cin, int num_units, int num_classes, int precision);
int xv_fcn_forward_encoder(xv_fcn* net, const float* x, int n, int h, int w, void* stream);
int xv_fcn_forward_head(xv_fcn* head, xv_fcn* const* towers_host, int num_towers,
                        const xv_fcn_outputs* outputs, void* stream);
int xv_fcn_destroy(xv_fcn* net);
/* `name` is the variable name below the expert prefix, e.g. "conv1_1/kernel", "score/bias",
 * "upscore/kernel", "conv3_2/moving_variance" (layout: SURVEY.md Appendix B;
 * base_model.py:395-451 import_weights assigns the same arrays to tf variables). */
int xv_fcn_set_param_host(xv_fcn* net, const char* name, const float* data_host,
                          const int64_t* shape_host, int ndim);
/* Packs the weights for the device (folds test-time batch norm, converts to bf16, detects the
 * channel-diagonal bilinear transposed convolutions).  Must be called after the last
 * xv_fcn_set_param_host and before xv_fcn_forward. */
int xv_fcn_finalize(xv_fcn* net);
/* x: [N,H,W,cin] float32.  H and W must be multiples of 16 (xview/datasets/augmentation.py
 * crop_multiple).  drop may be NULL (deterministic test-time network, simple_fcn.py:218-224). */
int xv_fcn_forward(xv_fcn* net, const float* x, int n, int h, int w, const xv_dropout_cfg* drop,
                   const xv_fcn_outputs* outputs, void* stream);
/* Diagnostics: copies the activation of a named layer of the LAST forward call to the host as
 * float32 NHWC (the reference returns every layer in the dict of simple_fcn.py:37-87).
 * shape_out_host receives [B,H,W,C].  Synchronous. */
int xv_fcn_get_layer_host(xv_fcn* net, const char* layer, float* out_host, size_t capacity_floats,
                          int64_t* shape_out_host, void* stream);

/* ---- training of the expert (SimpleFCN.fit: base_model.py:179-261, loss simple_fcn.py:205-215 +
 * utils.py:43-53, tf.train.AdamOptimizer base_model.py:153-162).  bf16 path, no batch norm. ---- */
/* Moves fp32 master copies of the trainable variables (conv kernels/biases, score_conv4/5, score;
 * the bilinear transposed convs are not trainable, simple_fcn.py:81,119-121) plus Adam moments
 * to the device.  num_params_out = length of the flat parameter / gradient vectors. */
int xv_fcn_train_begin(xv_fcn* net, int64_t* num_params_out);
/* Position of a variable ("conv3_2/kernel", "score/bias", ...) in the flat vectors. */
int xv_fcn_param_span(xv_fcn* net, const char* name, int64_t* offset_out, int64_t* size_out);
/* Forward + backward of one batch: x [N,H,W,cin] float32, labels [N,H,W] int32 (labels outside
 * [0,C) are ignored).  grads (device float32 [num_params]) is OVERWRITTEN with the gradient of
 * the summed cross-entropy, divided by the number of valid pixels if normalize != 0; loss_out
 * (device float64[2], may be NULL) receives {sum of -log p[label], number of valid pixels}. */
int xv_fcn_train_gradients(xv_fcn* net, const float* x, const int32_t* labels, int n, int h, int w,
                           int train_encoder, int normalize, float* grads, double* loss_out,
                           void* stream);
/* Same, for data-parallel training with the gradient all-reduce overlapping the backward pass:
 * the flat gradient is cut into xv_fcn_grad_buckets() contiguous buckets; bucket_events_host[i]
 * (a cudaEvent_t, may be NULL) is recorded on `stream` as soon as the i-th bucket IN COMPLETION
 * ORDER is final - event i covers the flat range [offsets[nb-1-i], offsets[nb-i]) - so the caller
 * can start that bucket's all-reduce on another stream while the rest of the backward pass runs.
 * normalize must be 0 (scale with xv_scale_by_count after the all-reduce). */
int xv_fcn_train_gradients_ex(xv_fcn* net, const float* x, const int32_t* labels, int n, int h,
                              int w, int train_encoder, int normalize, float* grads,
                              double* loss_out, void* const* bucket_events_host, int num_events,
                              void* stream);
/* offsets_out[0..nb] (ascending, offsets_out[nb] = num_params) of the nb gradient buckets. */
int xv_fcn_grad_buckets(xv_fcn* net, int64_t* offsets_out, int capacity, int* num_buckets_out);
/* grads *= 1 / (1e-20 + loss[1]) (after summing un-normalised gradients and counts over ranks). */
int xv_scale_by_count(float* grads, int64_t n, const double* loss, void* stream);
/* One Adam step (lr_t = lr sqrt(1-beta2^t)/(1-beta1^t)) on the master parameters followed by a
 * refresh of the bf16 operand copies used by the forward and data-gradient kernels. */
int xv_fcn_adam_step(xv_fcn* net, const float* grads, float learning_rate, float beta1,
                     float beta2, float epsilon, void* stream);
/* One step of the trainer named by config['trainer'] (base_model.py:157-162) with the TensorFlow
 * 1.x defaults: XV_OPT_ADAM (beta1 .9, beta2 .999, eps 1e-8), XV_OPT_ADAGRAD (accumulator starts
 * at 0.1), XV_OPT_RMSPROP (decay .9, momentum 0, eps 1e-10, mean square starts at 1); then the
 * same refresh of the bf16 operand copies.  Switching kinds re-initialises the slots. */
#define XV_OPT_ADAM 0
#define XV_OPT_ADAGRAD 1
#define XV_OPT_RMSPROP 2
int xv_fcn_optimizer_step(xv_fcn* net, const float* grads, int kind, float learning_rate,
                          void* stream);
/* Current master parameters (flat) to the host; synchronous. */
int xv_fcn_get_params_host(xv_fcn* net, float* out_host, int64_t capacity, void* stream);
int xv_fcn_train_end(xv_fcn* net);

/* ---- single layers (xview/models/custom_layers.py:124-139 conv2d, :71-121 deconv2d) ---- */
/* x [N,H,W,cin] -> out [N,H,W,cout]; k in {1,3}; stride 1 'same'.  precision BF16 needs
 * cin % 64 == 0 (or k == 3 with cin <= 3, the conv1_1 operand-packing path). */
int xv_conv2d(const float* x, const float* w_hwio_host, const float* bias_host, int n, int h,
              int w, int cin, int cout, int k, int relu, int precision, float* out,
              void* stream);
/* Gradient of the loss wrt the 3x3 kernel of such a layer, as the minimize() of
 * base_model.py:157-162 differentiates custom_layers.py:131: x [N,H,W,cin] layer input, dy
 * [N,H,W,cout] gradient wrt the pre-activation output (device float32, rounded to bf16 like the
 * activations fit() keeps) -> dw [3,3,cin,cout] (device float32, OVERWRITTEN), fp32
 * accumulation.  use_tensor_cores != 0: the tcgen05 kernel of fit() (cin % 64 == 0), else the
 * CUDA-core reference kernel (cin % 32 == 0); cout % 64 == 0.  The data gradient of the layer
 * is xv_conv2d itself on the flipped, transposed kernel. */
int xv_conv2d_weight_gradient(const float* x, const float* dy, int n, int h, int w, int cin,
                              int cout, int use_tensor_cores, float* dw, void* stream);
/* conv2d_transpose 'same', no bias: x [N,h,w,cin] -> out [N,h*stride,w*stride,cout] */
int xv_deconv2d(const float* x, const float* w_khkwoi_host, int n, int h, int w, int cin,
                int cout, int k, int stride, int relu, float* out, void* stream);
int xv_maxpool2x2(const float* x, int n, int h, int w, int c, float* out, void* stream);
/* tf.layers.batch_normalization(training=True) as custom_layers.py:116,132-134 applies it after a
 * convolution: x [N,H,W,c] float32 -> y = [relu](gamma * (x - mean) / sqrt(var + 1e-3) + beta) with
 * the statistics of the batch; mean_out / var_out (device float32 [c], biased variance) may be
 * NULL; moving_mean / moving_var (device float32 [c], may be NULL) are updated in place with
 * momentum 0.99 (Bessel-corrected variance). */
int xv_batchnorm_train(const float* x, const float* gamma, const float* beta, int n, int h, int w,
                       int c, int relu, float* y, float* mean_out, float* var_out,
                       float* moving_mean, float* moving_var, void* stream);
/* Its backward pass: dy = gradient wrt y, y = the forward output (ReLU mask if relu != 0) ->
 * dx [N,H,W,c], dgamma [c], dbeta [c] (all device float32, OVERWRITTEN). */
int xv_batchnorm_train_backward(const float* x, const float* y, const float* dy,
                                const float* gamma, int n, int h, int w, int c, int relu,
                                float* dx, float* dgamma, float* dbeta, void* stream);

/* ---- per-pixel fusion stage ---------------------------------------------------------- */
/* tf.nn.softmax + tf.argmax, basic_fusion_model.py:21-22.  prob / label may be NULL. */
int xv_softmax_argmax(const float* score, int64_t npix, int num_classes, float* prob,
                      void* label, int label_bytes, void* stream);
/* Bayes fusion through the decision table of bayes_mix.py:61-112: out = lut[l_0][l_1]...;
 * labels_host: host array of `num_experts` device pointers; lut: device int32 [C^M]. */
int xv_bayes_fuse_lut(const void* const* labels_host, int num_experts, int label_bytes,
                      const int32_t* lut, int num_classes, int64_t npix, void* out, void* stream);
/* One step of BayesFusion.score() behind the experts (basic_fusion_model.py:21-22 argmax,
 * bayes_mix.py:61-112 decision table, base_model.py:140-151 + :306-311 confusion matrix) as ONE
 * kernel: takes the low-resolution class scores left by the LAST xv_fcn_forward call of each of
 * the `num_experts` handles (bilinear decoder fast path), decodes each expert's label per pixel,
 * looks the fused label up in lut (device int32 [C^M]) and ACCUMULATES the confusion matrix
 * against gt_labels (device int32 [N,H,W], negative = ignore) into cm (device int64 [C,C]).
 * fused_out (device uint8 [N,H,W]) may be NULL.  No label map touches HBM otherwise. */
int xv_bayes_decode_score(xv_fcn* const* experts_host, int num_experts, const int32_t* lut,
                          int num_classes, const int32_t* gt_labels, int64_t* cm,
                          uint8_t* fused_out, void* stream);
/* Literal Bayes fusion, bayes_mix.py:12-58: score = sum_m log_cond[m][l_m][:] + log_prior;
 * log_cond: device [M,C,C] = log(1e-20 + conditional), log_prior: device [C].
 * score (float32 [npix,C]) and label may be NULL. */
int xv_bayes_fuse_score(const void* const* labels_host, int num_experts, int label_bytes,
                        const float* log_cond, const float* log_prior, int num_classes,
                        int64_t npix, float* score, void* label, void* stream);
/* Dirichlet fusion, dirichlet_mix.py:14-36 + :100-113.  alpha_m1: device [M,C_out,C_gt] =
 * sigma*alpha - 1, log_norm: device [M,C_gt] = lbeta(sigma*alpha[:,c]), log_prior: device [C]
 * = log(1e-20 + prior). */
int xv_dirichlet_fuse(const float* const* probs_host, int num_experts, const float* alpha_m1,
                      const float* log_norm, const float* log_prior, int num_classes,
                      int64_t npix, float* score, void* label, int label_bytes, void* stream);
/* Same rule with a bit-exact argmax: label (and score, if requested) equal the float32 statement
 * of dirichlet_mix.py:100-113 with the fixed operation order documented in DESIGN.md 4.3
 * (sequential sums, IEEE division, correctly rounded logarithm, separately rounded products and
 * sums; oracle.dirichlet_fusion_f32).  Pixels whose two best fast scores differ by more than a
 * rigorous bound on |fast - exact| keep the fast result (same argmax); the others are re-evaluated
 * in the exact arithmetic; with score != NULL every pixel is.  abs_alpha_m1_max =
 * max_c sum_{m,k} |alpha_m1[m][k][c]|, abs_norm_max = max_c sum_m |log_norm[m][c]| + max_c
 * |log_prior[c]| (host-side table magnitudes that scale the bound).  num_exact (device int64, may
 * be NULL) is incremented by the number of re-evaluated pixels. */
int xv_dirichlet_fuse_exact(const float* const* probs_host, int num_experts, const float* alpha_m1,
                            const float* log_norm, const float* log_prior, int num_classes,
                            int64_t npix, float abs_alpha_m1_max, float abs_norm_max, float* score,
                            void* label, int label_bytes, int64_t* num_exact, void* stream);
/* One step of DirichletFusion behind its two experts as ONE kernel: the decoder tail of both
 * (x8 upsampling + bias + softmax, simple_fcn.py:129-133 / basic_fusion_model.py:21) from the
 * low-resolution class scores left by the LAST xv_fcn_forward call of each handle, Dirichlet fusion
 * (dirichlet_mix.py:14-36,100-113; abs_alpha_m1_max >= 0: bit-exact argmax as in
 * xv_dirichlet_fuse_exact, < 0: fast arithmetic), then the fused label (label, may be NULL) and / or
 * the confusion matrix ACCUMULATED against gt_labels into cm (base_model.py:140-151; both NULL or
 * both set).  Labels are identical to xv_fcn_forward(prob) + xv_dirichlet_fuse_exact. */
int xv_dirichlet_decode_score(xv_fcn* const* experts_host, int num_experts, const float* alpha_m1,
                              const float* log_norm, const float* log_prior, int num_classes,
                              float abs_alpha_m1_max, float abs_norm_max, const int32_t* gt_labels,
                              int64_t* cm, void* label, int label_bytes, int64_t* num_exact,
                              void* stream);
/* average_mix.py:18-21; every sum and quotient individually rounded (bit-exact vs float32 numpy) */
int xv_average_fuse(const float* const* probs_host, int num_experts, int num_classes,
                    int64_t npix, float* score, void* label, int label_bytes, void* stream);
/* variance_mix.py:7-15; vars_host[m]: device float32 [npix] */
int xv_variance_fuse(const float* const* probs_host, const float* const* vars_host,
                     int num_experts, int num_classes, int64_t npix, float* score, void* label,
                     int label_bytes, void* stream);
/* MC-dropout moments over materialised samples [T,npix,C] (variance_mix.py:62-66,
 * bayesian_fcn.py:48-57).  Every output may be NULL. */
int xv_mc_moments(const float* samples, int num_samples, int64_t npix, int num_classes,
                  float* mean, float* var, float* mean_var, float* entropy, float* cond_entropy,
                  float* sum_var, void* stream);
/* Per-pixel maximum-likelihood Dirichlet fit over MC-dropout samples [T,npix,C]: moment
 * initialisation + Minka fixed point (dirichlet_fastfit.py:188-204,376-395, the batched form of
 * what dirichlet_mix.py:237-242 calls per class).  alpha: float32 [npix,C]; iterations
 * (int32 [npix], fixed-point steps taken) may be NULL. */
int xv_dirichlet_fit_samples(const float* samples, int num_samples, int64_t npix, int num_classes,
                             float tol, int maxiter, float* alpha, int32_t* iterations,
                             void* stream);
/* Uncertainty-mixed Dirichlet fusion, uncertainty_dirichlet_mix.py:18-52: per pixel and expert the
 * parameters are cond*(1-mix) + mix*(I+1) with mix = mean_c var / max(var); vars_host[m]: device
 * float32 [npix,C] MC-dropout variances, cond_params: device [M,C_out,C_gt], log_prior: device
 * [C] = log(1e-20 + prior). */
int xv_dirichlet_uncertainty_fuse(const float* const* probs_host, const float* const* vars_host,
                                  int num_experts, const float* cond_params,
                                  const float* log_prior, int num_classes, int64_t npix,
                                  float* score, void* label, int label_bytes, void* stream);
/* Dirichlet sufficient statistics, dirichlet_mix.py:142-163: ACCUMULATES into
 * stats (device float64 [C,C]) and counts (device int64 [C]). */
int xv_dirichlet_suffstats(const float* prob, const int32_t* labels, int64_t npix,
                           int num_classes, double* stats, int64_t* counts, void* stream);
/* Confusion matrix, base_model.py:140-151: ACCUMULATES into cm (device int64 [C,C], rows =
 * ground-truth label, cols = prediction); negative labels are ignored. */
int xv_confusion_accumulate(const void* pred, int pred_bytes, const int32_t* labels, int64_t npix,
                            int num_classes, int64_t* cm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XVIEW_B200_H_ */
