// Shared between abi.cu (FCN expert, C ABI) and adapnet.cu: device buffers, the packed
// convolution layer record, the network handle and the helpers that build TMA descriptors and
// launch the tensor-core convolutions.
#pragma once
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/xview_b200.h"
#include "../../include/xview_b200_measure.h"
#include "common.cuh"
#include "kernels.h"

namespace xv {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int ensure(size_t n) {
    if (n <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    XV_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return 0;
  }
  template <typename T>
  int upload(const std::vector<T>& v) {
    XV_TRY(ensure(v.size() * sizeof(T)));
    XV_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
  }
};

struct Arena {
  char* base = nullptr;
  size_t off = 0;
  void* alloc(size_t bytes) {
    off = (off + 1023) & ~static_cast<size_t>(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

struct HostParam {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct ConvLayer {
  std::string name;
  int k = 3, cin = 0, cout = 0, relu = 1;
  // bf16 path
  int taps = 9, kdim = 0, cin_gemm = 0, cout_pad = 0, block_n = 0;
  bool use_t = false;   // few output channels: transposed-role kernel (conv_igemm_t_sm100.cu)
  // generic geometry (Adapnet): tap t reads the input at ((t / k) * dil - pad, (t % k) * dil - pad);
  // `sample` = 2 reads every second input pixel (stride-2 1x1 convolutions, adapnet.py:39,45)
  bool generic = false;
  int dil = 1, pad = 0, sample = 1;
  bool stride2 = false;        // stride-2 filter read through four parity descriptors
  double macs_per_pixel = 0;   // real multiply-adds per output pixel (profiling)
  DevBuf w_packed, bias_pad;
  // fp32 path
  DevBuf w_f32, bias_f32, bn_scale, bn_shift;
  bool has_bn = false;
};

enum class DType { F32, BF16, U8 };
struct Act {
  void* p = nullptr;
  DType dt = DType::F32;
  int B = 0, H = 0, W = 0, C = 0;
  int pitch = 0;   // physical channels per pixel when the tensor is a channel slice (0: == C)
  size_t elems() const { return static_cast<size_t>(B) * H * W * C; }
  int stride_c() const { return pitch ? pitch : C; }
};

}  // namespace xv

using namespace xv;

struct xv_fcn {
  int cin = 0, nu = 0, C = 0, precision = 0;
  // bit 0: batch norm on every layer (simple_fcn.py `batchnorm`), bit 1 (XV_BN_DECODER): on the
  // decoder's upscore / score only - fusion_fcn.py:39 calls decoder() with its default
  // batchnorm=True while the towers and heads are built without
  int batchnorm = 0;
  bool bn_all() const { return (batchnorm & 1) != 0; }
  bool bn_decoder() const { return (batchnorm & 3) != 0; }
  int arch = 0;        // 0 = VGG16-FCN (simple_fcn.py), 1 = Adapnet (adapnet.py)
  int role = 0;        // 0 = whole expert, 1 = VGG16 encoder only, 2 = head + decoder only
  int head_cin = 512;  // input channels of score_conv4/5 (1024 for the two-tower fusion_fcn head)
  bool finalized = false;
  std::map<std::string, HostParam> params;
  std::vector<std::unique_ptr<ConvLayer>> convs;   // conv1_1..conv5_3, score_conv4, score_conv5, score
  // decoder
  bool fast_up5 = false, fast_up = false;
  // the loaded transposed-conv kernels are channel-diagonal (g4 / g16 hold the diagonals); the
  // fast inference paths additionally need the layer to be free of batch norm
  bool diag_up5 = false, diag_up = false;
  DevBuf g4, g16, w_score_nuxc, b_score;           // fast paths
  DevBuf w_up5, w_up, up5_scale, up5_shift, up_scale, up_shift;   // generic paths
  DevBuf arena_buf;
  std::map<std::string, Act> layers;
  // Adapnet tail (adapnet.py:156-166): BN shifts of the two transposed convolutions, fp32-mode
  // transposed kernels and BN factors
  DevBuf up1_shift, up2_shift, up1_scale, up2_scale, w_up1, w_up2;
  std::map<std::array<long long, 9>, CUtensorMap> tmaps;

  ConvLayer* conv(const std::string& n) {
    for (auto& c : convs)
      if (c->name == n) return c.get();
    return nullptr;
  }
};


namespace xv {

int ensure_init();
int debug_flags();
int get_param(xv_fcn* net, const std::string& name, std::vector<int64_t> shape,
              const HostParam** out);
// test-time batch norm of scope `layer` as y = scale * x + shift (identity when `enabled` is false)
int bn_factors_of(xv_fcn* net, const std::string& layer, int cout, bool enabled,
                  std::vector<float>* scale, std::vector<float>* shift);
int pack_conv(xv_fcn* net, ConvLayer* L, const float* w_hwio, const float* bias,
              const std::vector<float>& scale, const std::vector<float>& shift, bool bn);
// activation descriptor: C channels of a [N,H,W,pitch] bf16 tensor starting at `ptr`, every
// `sample`-th pixel in both spatial directions, box {64, tw, th, 1}
int get_tmap_ex(xv_fcn* net, CUtensorMap* out, const void* ptr, int N, int H, int W, int C,
                int pitch, int sample, int th, int tw);
int get_tmap(xv_fcn* net, CUtensorMap* out, const void* ptr, int N, int H, int W, int C, int th,
             int tw);
int get_tmap_w(xv_fcn* net, CUtensorMap* out, const void* ptr, int kdim, int cout_pad,
               int block_n);
// conv1_1-style first layer on the raw fp32 input
int run_igemm_c1(xv_fcn* net, const ConvLayer& L, const float* x, int B, int H, int W, void* out,
                 cudaStream_t s);
// Generic-geometry tensor-core convolution.  in: bf16 [B,H,W,in_pitch] (first cin_gemm channels);
// out: bf16 [B,Ho,Wo,out_pitch] (first `cout` channels written) or fp32 [B,Ho,Wo,cout];
// Ho = ceil(H / sample).
// `residual` (bf16 [B,Ho,Wo,cout], pixel-major bf16 epilogue only): out = relu(conv + residual).
int run_conv_generic(xv_fcn* net, const ConvLayer& L, const void* in, int B, int H, int W,
                     int in_pitch, void* out, int out_pitch, bool out_f32, cudaStream_t s,
                     const void* residual = nullptr);

// adapnet.cu
int adapnet_finalize(xv_fcn* net);
int adapnet_forward(xv_fcn* net, const float* x, int n, int h, int w, const xv_fcn_outputs* o,
                    cudaStream_t s);

}  // namespace xv
