// Internal C++ interface between the C-ABI layer (abi.cu) and the kernel files.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xv {

// ------------------------------------------------------------- conv_igemm_sm100.cu
struct ConvIgemmParams {
  CUtensorMap tmap_in;    // bf16 [N,H,W,Cin] viewed as (Cin, W, H, N), box {64, tw, th, 1}, SW128
  CUtensorMap tmap_w;     // bf16 [CoutPad, taps*Cin], box {64, BLOCK_N}, SW128
  CUtensorMap tmap_out;   // bf16 [N,H,W,Cout] box {64, tw, th, 1}, SW128 (bf16 epilogue only)
  const float* bias;      // [CoutPad] fp32
  float* out_f32;         // fp32 epilogue: [N,H,W,cout]
  int N, H, W;
  int cin;                // multiple of 64
  int cout;               // real output channels (<= CoutPad)
  int th, tw;             // spatial tile, th*tw == 128
  int tiles_x, tiles_y;
  int n_blocks;           // CoutPad / BLOCK_N
  int relu;
  int debug_flags;        // timing experiments only (xv_bench_conv_igemm); 0 in production
  const float* x_raw;     // conv1_1 mode: raw fp32 input [N,H,W,cin_raw] (operand packed on the fly)
  // generic filter geometry (taps template argument 0): tap t reads the input shifted by
  // ((t / kw) * dil - pad, (t % kw) * dil - pad); covers dilated 3x3 (adapnet.py:83-86) and the
  // 4x4 form of the stride-2 7x7 convolution on the space-to-depth input (adapnet.py:121)
  int taps, kw, dil, pad;
  // bf16 epilogue of the pixel-major kernel (BLOCK_N = 256): out = relu(act(conv + bias) +
  // residual), the block output of adapnet.py:49,96; tmap_res describes the residual tensor
  // exactly like tmap_out describes the output
  int has_residual;
  CUtensorMap tmap_res;
  // stride-2 filters on the transposed-role kernel (7x7 of adapnet.py:121): the input is read
  // through four descriptors that each sample every second pixel, one per (row, column) parity;
  // tap offset e = t * dil - pad selects parity e & 1 and coordinate shift e >> 1
  int stride2;
  CUtensorMap tmap_in_par[4];
  int patch_w;            // conv_c1_sm100.cu: floats per row of the input patch
  int halo;               // conv_igemm_t: halo variant, tmap_in box {64, 16, 18, 1}
  int w_rows;             // conv_igemm_t: rows of the weight box, 64 (Cout <= 64) or 128
  // conv_igemm_2cta: fused 2x2 max pool in the epilogue, 1 = pooled output only, 2 = full + pooled;
  // tmap_pool = pooled tensor [N, H/2, W/2, Cout], box {64, tw/2, th/2, 1}
  int pool_mode;
  CUtensorMap tmap_pool;
};
// conv_igemm_rowpair_sm100.cu: 3x3 conv (Cin = 64, Cout <= 64) + bias + ReLU + 2x2 max pool, both
// halves of the accumulator lanes in use; tmap_in_par[0/1] = even / odd input rows, box
// {64, 16, 17, 1}; tmap_w box {64, 64}; tmap_out = pooled output, box {64, 8, 16, 1}
// pool == false: the full activation instead (tmap_out box {64, 16, 8, 1} on [N, H, W, 64]).
int launch_conv_igemm_rowpair(const ConvIgemmParams& p, bool pool, cudaStream_t stream);
int conv_igemm_block_n(int cout);
int launch_conv_igemm(const ConvIgemmParams& p, int block_n, int taps, bool out_f32,
                      cudaStream_t stream);
// conv1_1: 3x3 conv of the raw fp32 input (cin_raw <= 3) as a K=64 GEMM whose A rows
// [hi taps | lo taps | 0] are built in shared memory by producer warps (no im2col buffer).
int launch_conv_igemm_c1(const ConvIgemmParams& p, int cin_raw, cudaStream_t stream);
// conv_c1_sm100.cu: same layer with the fp32 input patch (tile + halo) staged by TMA and the
// weights resident in shared memory; tmap_in = fp32 input viewed as (W*Cin, H, N) with box
// {patch_w, th + 2, 1}.  conv_c1_patch_fits tells whether a tile shape is supported.
bool conv_c1_patch_fits(int th, int tw, int cin_raw, int* patch_w);
int launch_conv_c1(const ConvIgemmParams& p, int cin_raw, cudaStream_t stream);

// conv_igemm_2cta_sm100.cu: CTA-pair (cta_group::2) 3x3 kernel for Cout multiples of 256, halo
// operands; tmap_in box {64, tw, th + 2, 1}, tmap_w box {64, 128}, n_blocks = CoutPad / 256.
int launch_conv_igemm_2cta(const ConvIgemmParams& p, int block_n, cudaStream_t stream);

// conv_igemm_t_sm100.cu: 3x3, Cout <= 128 per block of 128, 16x16 pixel tiles, optional fused
// 2x2 max pool (then tmap_out describes the pooled tensor, box {64,8,8,1}; else {64,16,8,1}).
int launch_conv_igemm_t(const ConvIgemmParams& p, bool pool, cudaStream_t stream);
// any filter geometry (p.taps / kw / dil / pad), no pool
int launch_conv_igemm_t_generic(const ConvIgemmParams& p, cudaStream_t stream);

// conv_wgrad_sm100.cu: tensor-core weight gradient of a 3x3 'same' convolution
struct ConvWgradParams {
  CUtensorMap tmap_x;     // bf16 [N,H,W,Cin]  box {64, tw, th, 1}, SW128
  CUtensorMap tmap_dy;    // bf16 [N,H,W,Cout] box {64, tw, th, 1}, SW128
  float* dw;              // fp32 [9][Cin][Cout], accumulated into
  int N, H, W, cin, cout;
  int th, tw, tiles_x, tiles_y;
  int m_blocks;           // ceil(Cout / 128)
  int taps;               // 9 (3x3 'same') or 1 (1x1)
  int total_atoms;        // taps * Cin / 64  ((tap, 64-channel chunk) pairs)
  int n_groups;           // ceil(total_atoms / 4)
  int k_splits;           // pixel tiles are dealt round-robin to k_splits CTAs
};
int launch_conv_wgrad_tc(const ConvWgradParams& p, cudaStream_t stream);
// conv_wgrad_2cta_sm100.cu: the same for Cout % 256 == 0, Cin % 128 == 0 on CTA pairs
// (cta_group::2, M = 256); m_blocks = Cout / 256
int launch_conv_wgrad_2cta(const ConvWgradParams& p, cudaStream_t stream);

// ------------------------------------------------------------- layers.cu
int launch_im2col_c1(const float* x, __nv_bfloat16* out, int N, int H, int W, int cin,
                     cudaStream_t s);
int launch_to_f32(const void* in, int src_dtype, float* out, size_t n, cudaStream_t s);
int launch_f32_to_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t s);
int launch_bf16_to_f32(const __nv_bfloat16* in, float* out, size_t n, cudaStream_t s);
int launch_maxpool_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int N, int H, int W, int C,
                        cudaStream_t s);
int launch_maxpool_f32(const float* in, float* out, int N, int H, int W, int C, cudaStream_t s);

struct DropoutSpec {
  float rate = 0.f;               // drop probability; keep = 1 - rate
  const uint8_t* ext_mask = nullptr;  // optional keep-mask, one byte per OUTPUT element
  uint64_t seed = 0;              // Philox key
  uint64_t offset = 0;            // Philox counter offset (distinct per site)
  size_t pass_elems = 0;          // leading output elements copied through without dropout;
                                  // ext_mask / Philox indices start after them
};
// out[r * n + i] = dropout(in[i]) for r in [0, replicate); n elements per copy.
int launch_dropout_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, size_t n, int replicate,
                        const DropoutSpec& d, cudaStream_t s);
int launch_dropout_f32(const float* in, float* out, size_t n, int replicate,
                       const DropoutSpec& d, cudaStream_t s);

// generic fp32 reference-order layers (validation mode, any shapes)
int launch_conv_f32(const float* x, const float* w_hwio, const float* bias, float* out, int N,
                    int H, int W, int cin, int cout, int k, int relu, cudaStream_t s);
int launch_deconv_f32(const float* x, const float* w_khkwoi, float* out, int N, int hin, int win,
                      int cin, int cout, int k, int stride, int relu, const float* addend,
                      cudaStream_t s);
int launch_concat_bf16(const __nv_bfloat16* src, __nv_bfloat16* dst, size_t npix, int c_src,
                       int c_dst, int offset, cudaStream_t s);
int launch_add_f32(const float* a, const float* b, float* out, size_t n, cudaStream_t s);
int launch_affine_f32(float* x, const float* scale, const float* shift, size_t npix, int C,
                      int relu, cudaStream_t s);

// adapnet_kernels.cu: the memory-bound steps between Adapnet's convolutions
int launch_add_relu_bf16(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* out,
                         size_t n, cudaStream_t s);
int launch_add_relu_f32(const float* a, const float* b, float* out, size_t n, cudaStream_t s);
// transposed convolution, gather half: col [B,hin,win,k*k*cout] (fp32 or bf16) -> out
// [B,hin*stride,win*stride,cout] = sum of the overlapping taps + shift[co] (+ addend); fp32 output
// and / or bf16 output zero-padded to pad_c channels
int launch_col2im(const void* col, bool col_bf16, const float* shift, const float* addend,
                  float* out_f32, __nv_bfloat16* out_bf16, int pad_c, int B, int hin, int win,
                  int cout, int k, int stride, cudaStream_t s);
// d2s bf16 [B,h8,w8,64*C], channel (py*8+px)*C+c -> full-resolution score / softmax / argmax
int launch_d2s_softmax_argmax(const __nv_bfloat16* d2s, int B, int h8, int w8, int C, float* score,
                              float* prob, int64_t* label_i64, uint8_t* label_u8, cudaStream_t s);
// fp32 convolution with TF 'SAME' padding for any stride / dilation (validation mode)
int launch_conv_f32_ex(const float* x, const float* w_hwio, const float* bias, float* out, int N,
                       int H, int W, int cin, int cout, int k, int stride, int dil, int relu,
                       cudaStream_t s);

// fast decoder pieces (channel-diagonal transposed convolutions)
// fused = s4 + relu(sum_taps g[ky,kx,u] * s5[tap,u]);  s4 [N,2h,2w,nu], s5 [N,h,w,nu]
int launch_upscore2_add(const float* s5, const float* s4, const float* g_4x4xnu, float* fused,
                        int N, int h, int w, int nu, cudaStream_t s, float* up5 = nullptr);
// low[n,y,x,c] = sum_u fused[n,y,x,u] * w[u,c]
int launch_score_lowres(const float* fused, const float* w_nuxc, float* low, size_t npix, int nu,
                        int C, cudaStream_t s);

// both of the above in one pass: fused = s4 + relu(up2(s5)) (written) and low = fused x w
bool head_fused_supported(int nu, int C);
int launch_head_fused(const float* s5, const float* s4, const float* g_4x4xnu, const float* w_nuxc,
                      float* fused, float* low, int N, int h, int w, int nu, int C, cudaStream_t s);

struct DecodeOut {
  uint8_t* label_u8 = nullptr;    // [N,H,W]
  int64_t* label_i64 = nullptr;   // [N,H,W]
  float* prob = nullptr;          // [N,H,W,C]
  float* score = nullptr;         // [N,H,W,C]
};
// score = x8 bilinear-like upsample (shared 16x16 kernel g) of `low` + bias; softmax; argmax
int launch_decode_upsample8(const float* low, const float* g_16x16, const float* bias, int N,
                            int h, int w, int C, const DecodeOut& out, cudaStream_t s);
// batch-normalised decoder in one pass: upscore = relu(up_scale * up8(feat) + up_shift),
// score = sc_scale * (upscore x w + bias) + sc_shift, softmax, argmax
bool decode_bn_supported(int nu, int C);
int launch_decode_bn_upsample8(const float* feat, const float* g_16x16, const float* up_scale,
                               const float* up_shift, const float* w_nuxc, const float* bias,
                               const float* sc_scale, const float* sc_shift, int N, int h, int w,
                               int nu, int C, const DecodeOut& out, cudaStream_t s);
// label-only decode of M experts + decision-table lookup + confusion-matrix accumulation in one
// pass (BayesFusion.score()); fused_out (uint8 [N,8h,8w]) may be NULL
int launch_decode_bayes_confusion(const float* const* low, const float* const* g,
                                  const float* const* bias, int M, const int32_t* lut, int C,
                                  int N, int h, int w, const int32_t* gt, long long* cm,
                                  uint8_t* fused_out, cudaStream_t s);
// MC variant: `low` holds T sample maps [T,N,h,w,C]; accumulates population mean / variance
// of the per-sample softmax without materialising the samples.
int launch_decode_upsample8_mc(const float* low, const float* g_16x16, const float* bias, int T,
                               int N, int h, int w, int C, float* mean_prob, float* var_prob,
                               float* mean_var, cudaStream_t s);

// ------------------------------------------------------------- fusion.cu
int launch_softmax_argmax(const float* score, int64_t npix, int C, float* prob, int64_t* label64,
                          uint8_t* label8, cudaStream_t s);
int launch_bayes_lut(const void* const* labels_dev, int M, int label_bytes, const int32_t* lut,
                     int C, int64_t npix, void* out, cudaStream_t s);
int launch_bayes_score(const void* const* labels_dev, int M, int label_bytes,
                       const float* logcond /*[M,C,C]*/, const float* logprior /*[C]*/, int C,
                       int64_t npix, float* score, void* label_out, cudaStream_t s);
int launch_dirichlet_fuse(const float* const* probs, int M, const float* alpha_m1 /*[M,C,C]*/,
                          const float* lognorm /*[M,C]*/, const float* logprior /*[C]*/, int C,
                          int64_t npix, float* score, void* label_out, int label_bytes,
                          float exact_amax /* < 0: fast arithmetic only */, float exact_tail,
                          unsigned long long* n_exact /* += pixels re-evaluated exactly */,
                          cudaStream_t s);
int launch_average_fuse(const float* const* probs, int M, int C, int64_t npix, float* score,
                        void* label_out, int label_bytes, cudaStream_t s);
int launch_variance_fuse(const float* const* probs, const float* const* vars, int M, int C,
                         int64_t npix, float* score, void* label_out, int label_bytes,
                         cudaStream_t s);
int launch_mc_moments(const float* samples /*[T,npix,C]*/, int T, int64_t npix, int C, float* mean,
                      float* var, float* mean_var, float* entropy, float* cond_entropy,
                      float* sum_var, cudaStream_t s);
int launch_suffstats(const float* prob, const int32_t* labels, int64_t npix, int C,
                     double* S /*[C,C] +=*/, long long* n /*[C] +=*/, cudaStream_t s);
int launch_confusion(const void* pred, int pred_bytes, const int32_t* labels, int64_t npix, int C,
                     long long* cm /*[C,C] +=*/, cudaStream_t s);

// ------------------------------------------------------------- decode_dirichlet.cu
// fused tail of DirichletFusion for two experts: x8 decode + softmax of both -> Dirichlet fusion
// (fast form, exact re-evaluation of near-ties when exact_amax >= 0) -> label_out (may be NULL)
// and / or confusion-matrix accumulation against gt into cm (both NULL or both set)
int launch_decode_dirichlet(const float* const* low, const float* const* g,
                            const float* const* bias, int M, const float* alpha_m1,
                            const float* lognorm, const float* logprior, float exact_amax,
                            float exact_tail, int C, int N, int h, int w, const int32_t* gt,
                            long long* cm, void* label_out, int label_bytes,
                            unsigned long long* n_exact, cudaStream_t s);

// ------------------------------------------------------------- mc_dirichlet.cu
int launch_dirichlet_fit_samples(const float* samples, int T, int64_t npix, int C, float tol,
                                 int maxiter, float* alpha, int* iters, cudaStream_t s);
int launch_dirichlet_uncertainty_fuse(const float* const* probs, const float* const* vars,
                                      const float* const* max_var, int M, const float* cond,
                                      const float* logprior, int C, int64_t npix, float* score,
                                      void* label_out, int label_bytes, cudaStream_t s);
int launch_reduce_max(const float* x, int64_t n, float* out, cudaStream_t s);

// ------------------------------------------------------------- train.cu
int launch_ce_grad(float* score, const int32_t* labels, int64_t npix, int C, double* loss,
                   float* dbias, cudaStream_t s);
// decode + softmax - onehot + loss + transposed x8 upsampling in one pass (bilinear fast path);
// dlow [N,h,w,C] is overwritten, loss[0..1] / dbias[C] are accumulated into
int launch_loss_lowres_grad(const float* low, const float* g_16x16, const float* bias,
                            const int32_t* labels, int N, int h, int w, int C, float* dlow,
                            double* loss, float* dbias, cudaStream_t s);
int launch_upsample8_transpose(const float* dscore, const float* g, float* dlow, int N, int h,
                               int w, int C, cudaStream_t s);
int launch_score_bwd(const float* dlow, const float* fused, const float* w, float* dfused,
                     float* dw, size_t npix, int nu, int C, cudaStream_t s);
int launch_outer_sum_bf16(const __nv_bfloat16* a, const float* b, float* dw, size_t npix, int I,
                          int J, cudaStream_t s);
int launch_upscore2_bwd(const float* dfused, const float* up5, const float* g, float* ds5, int N,
                        int h, int w, int nu, cudaStream_t s);
int launch_relu_mask_f32(const float* dy, const float* y, float* dpre, __nv_bfloat16* dpre_bf16,
                         size_t npix, int c, int c_pad, cudaStream_t s);
int launch_relu_bwd_bf16(const __nv_bfloat16* da, const __nv_bfloat16* db, const __nv_bfloat16* y,
                         __nv_bfloat16* dy, size_t n, int cout, float* bias_grad, cudaStream_t s);
int launch_pool_relu_bwd_bf16(const __nv_bfloat16* dp, const __nv_bfloat16* y,
                              const __nv_bfloat16* p, const __nv_bfloat16* extra,
                              __nv_bfloat16* dy, int N, int H, int W, int C, float* bias_grad,
                              cudaStream_t s);
int launch_bias_grad_bf16(const __nv_bfloat16* dy, float* db, size_t npix, int cout,
                          cudaStream_t s);
int launch_bias_grad_f32(const float* dy, float* db, size_t npix, int cout, cudaStream_t s);
int launch_conv_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int N, int H,
                      int W, int cin, int cout, cudaStream_t s);
int launch_conv_wgrad_c1(const float* x, const __nv_bfloat16* dy, float* dw, int N, int H, int W,
                         int cin, int cout, cudaStream_t s);
int launch_scale_by_count(float* g, size_t n, const double* loss, cudaStream_t s);
int launch_adam(float* w, const float* g, float* m, float* v, size_t n, float lr_t, float b1,
                float b2, float eps, cudaStream_t s);
int launch_adagrad(float* w, const float* g, float* acc, size_t n, float lr, cudaStream_t s);
int launch_rmsprop(float* w, const float* g, float* ms, float* mom, size_t n, float lr,
                   float decay, float momentum, float eps, cudaStream_t s);
int launch_fill_f32(float* x, size_t n, float value, cudaStream_t s);
// every conv layer of a network re-packed in one launch (see launch_pack_weights for the layouts)
struct PackLayer {
  const float* w;          // fp32 HWIO master kernel
  const float* b;          // fp32 master bias [cout]
  __nv_bfloat16* fwd;      // forward operand rows
  __nv_bfloat16* bwd;      // data-gradient operand rows (may be NULL)
  float* bias_pad;         // padded bias of the forward layer
  unsigned long long first, elems;   // range of this layer in the concatenated element index
  int taps, cin, cout, fwd_kdim, bwd_kdim, c1_layout;
};
struct PackAllParams {
  PackLayer layer[20];
  int num_layers;
  unsigned long long total;
};
int launch_pack_all(const PackAllParams& p, cudaStream_t s);
int launch_pack_weights(const float* w, __nv_bfloat16* fwd, __nv_bfloat16* bwd, int taps, int cin,
                        int cout, int fwd_kdim, int bwd_kdim, int c1_layout, cudaStream_t s);

// ------------------------------------------------------------- bn_train.cu
// batch normalisation on batch statistics (training mode); sums: device double[2*C] scratch
int launch_bn_stats(const void* z, bool bf16, size_t npix, int C, double* sums, cudaStream_t s);
int launch_bn_finalize(const double* sums, size_t npix, int C, float eps, float momentum,
                       const float* bias_extra, float* mean, float* rstd, float* moving_mean,
                       float* moving_var, cudaStream_t s);
int launch_bn_apply(const void* z, bool bf16, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, size_t npix, int C, int relu, void* y,
                    cudaStream_t s);
// g: gradient wrt the layer output; y_mask != NULL: masked by (y_mask > 0) (ReLU) on the fly.
// Writes dz (same dtype as z), dgamma[C], dbeta[C] and optionally (fp32 only) dz as bf16 padded
// to c_pad channels.
int launch_bn_backward(const void* g, const void* y_mask, const void* z, bool bf16,
                       const float* mean, const float* rstd, const float* gamma, size_t npix,
                       int C, double* sums, void* dz, __nv_bfloat16* dz_pad, int c_pad,
                       float* dgamma, float* dbeta, cudaStream_t s);
// channel-diagonal transposed convolution (k = 2 * stride, 'same') and its transpose, fp32 NHWC
int launch_upsample_diag(const float* in, const float* g, bool per_channel, int N, int h, int w,
                         int C, int k, int stride, float* out, cudaStream_t s);
int launch_upsample_diag_transpose(const float* dout, const float* g, bool per_channel, int N,
                                   int h, int w, int C, int k, int stride, float* din,
                                   cudaStream_t s);
int launch_add_f32_inplace(float* a, const float* b, size_t n, cudaStream_t s);

constexpr int kMaxClasses = 24;

}  // namespace xv
