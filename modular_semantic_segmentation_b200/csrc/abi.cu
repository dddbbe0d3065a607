// C-ABI layer of libxview_b200: handle management, weight packing, the layer schedule of the
// FCN expert and thin wrappers around the fusion kernels.  See include/xview_b200.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "net.h"

namespace xv {

// ------------------------------------------------------------------ error + device state
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
int fail(const std::string& msg) {
  g_error = msg;
  return -1;
}

static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }

// optional per-launch timing of the tensor-core convolutions (bench.py roofline)
struct IgemmSample {
  cudaEvent_t e0, e1;
  double flops;
  int block_n;
};
static bool g_profile = false;
static std::vector<IgemmSample> g_samples;

static int g_debug_flags = 0;   // bit0: no pool fusion, bit1: no transposed kernel, bit2: conv1_1 via im2col buffer, bit3: CUDA-core weight gradients, bit10: conv1_2 without the row-pair kernel, bit11: single-CTA weight gradients, bit12: three-kernel loss head

static DeviceInfo g_dev;
const DeviceInfo& device_info() { return g_dev; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int debug_flags() { return g_debug_flags; }

int ensure_init() {
  if (g_dev.device >= 0) return 0;
  int dev = 0;
  XV_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  XV_CUDA(cudaGetDeviceProperties(&prop, dev));
  XV_CHECK(prop.major == 10, "xview_b200 needs an sm_100-class GPU (found sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor) + ")");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  XV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  XV_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess,
           "cuTensorMapEncodeTiled not available from the driver");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  g_dev.num_sms = prop.multiProcessorCount;
  g_dev.smem_optin = prop.sharedMemPerBlockOptin;
  g_dev.device = dev;
  return 0;
}

// ------------------------------------------------------------------ TMA descriptors
static int make_tmap_act(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int pitch,
                         int sample, int th, int tw) {
  const cuuint64_t px = static_cast<cuuint64_t>(pitch) * 2;       // bytes per physical pixel
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(div_up(W, sample)),
                        static_cast<cuuint64_t>(div_up(H, sample)), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {px * sample, px * W * sample, px * W * H};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string(r));
  return 0;
}

static int make_tmap_w(CUtensorMap* m, const void* ptr, int kdim, int cout_pad, int block_n);

// fp32 input [N,H,W,cin] viewed as (W*cin, H, N): box = one input patch {patch_w, rows, 1}
static int make_tmap_patch(CUtensorMap* m, const void* ptr, int N, int H, int W, int cin,
                           int patch_w, int rows) {
  const cuuint64_t row_bytes = static_cast<cuuint64_t>(W) * cin * 4;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(W) * cin, static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[2] = {row_bytes, row_bytes * H};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(patch_w), static_cast<cuuint32_t>(rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(input patch) failed: " + std::to_string(r));
  return 0;
}

// in: bf16 [B,H,W,cin]; out: bf16 [B,H,W,cout] or, with pool, [B,H/2,W/2,cout]
static int run_igemm_t(xv_fcn* net, const ConvLayer& L, const void* in, int B, int H, int W,
                       void* out, bool pool, cudaStream_t s);
static bool pair_kernel_applies(const ConvLayer& L, int H, int W);

static int make_tmap_w(CUtensorMap* m, const void* ptr, int kdim, int cout_pad, int block_n) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(kdim), static_cast<cuuint64_t>(cout_pad)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(kdim) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(block_n)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string(r));
  return 0;
}

// 128-pixel spatial tile minimising padded area, then preferring square-ish patches.
static void choose_tile(int H, int W, int* th, int* tw) {
  long best_area = -1;
  int best_skew = 0;
  for (int t = 1; t <= 128; t *= 2) {
    const int a = t, b = 128 / t;   // th = a, tw = b
    const long area = static_cast<long>(div_up(H, a)) * a * div_up(W, b) * b;
    const int skew = a > b ? a / b : b / a;
    if (best_area < 0 || area < best_area || (area == best_area && skew < best_skew)) {
      best_area = area;
      best_skew = skew;
      *th = a;
      *tw = b;
    }
  }
}

}  // namespace xv

using namespace xv;

namespace xv {

static const char* kConvNames[13] = {"conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1",
                                     "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
                                     "conv5_1", "conv5_2", "conv5_3"};
static const int kConvCout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};

int get_param(xv_fcn* net, const std::string& name, std::vector<int64_t> shape,
                     const HostParam** out) {
  auto it = net->params.find(name);
  XV_CHECK(it != net->params.end(), "parameter '" + name + "' was never set");
  XV_CHECK(it->second.shape == shape, "parameter '" + name + "' has the wrong shape");
  *out = &it->second;
  return 0;
}

// BN folding factors: y = scale * conv + shift (test-time tf.layers.batch_normalization,
// epsilon 1e-3, custom_layers.py:116,132-134)
int bn_factors_of(xv_fcn* net, const std::string& layer, int cout, bool enabled,
                  std::vector<float>* scale, std::vector<float>* shift) {
  scale->assign(cout, 1.f);
  shift->assign(cout, 0.f);
  if (!enabled) return 0;
  const HostParam *g, *b, *m, *v;
  XV_TRY(get_param(net, layer + "/gamma", {cout}, &g));
  XV_TRY(get_param(net, layer + "/beta", {cout}, &b));
  XV_TRY(get_param(net, layer + "/moving_mean", {cout}, &m));
  XV_TRY(get_param(net, layer + "/moving_variance", {cout}, &v));
  for (int c = 0; c < cout; ++c) {
    const float s = g->data[c] / std::sqrt(v->data[c] + 1e-3f);
    (*scale)[c] = s;
    (*shift)[c] = b->data[c] - m->data[c] * s;
  }
  return 0;
}

static int bn_factors(xv_fcn* net, const std::string& layer, int cout, std::vector<float>* scale,
                      std::vector<float>* shift) {
  const bool decoder_layer = layer == "score" || layer == "upscore";
  return bn_factors_of(net, layer, cout, decoder_layer ? net->bn_decoder() : net->bn_all(), scale,
                       shift);
}

static inline uint16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

// Packs one conv layer for both precisions.  `special_c1`: conv1_1 operand layout
// [hi taps | lo taps | 0] with K = 64 (see layers.cu im2col_c1_kernel).
int pack_conv(xv_fcn* net, ConvLayer* L, const float* w_hwio, const float* bias,
                     const std::vector<float>& scale, const std::vector<float>& shift,
                     bool bn) {
  const int k = L->k, cin = L->cin, cout = L->cout, taps = k * k;
  if (net->precision == XV_PRECISION_FP32) {
    std::vector<float> w(w_hwio, w_hwio + static_cast<size_t>(taps) * cin * cout);
    std::vector<float> b(cout, 0.f);
    if (bias) b.assign(bias, bias + cout);
    XV_TRY(L->w_f32.upload(w));
    XV_TRY(L->bias_f32.upload(b));
    L->has_bn = bn;
    if (bn) {
      XV_TRY(L->bn_scale.upload(scale));
      XV_TRY(L->bn_shift.upload(shift));
    }
    return 0;
  }
  const bool special_c1 = (k == 3 && cin <= 3);
  XV_CHECK(special_c1 || cin % 64 == 0,
           "bf16 path: Cin of '" + L->name + "' must be a multiple of 64 (or <= 3 for a 3x3 layer)");
  L->taps = special_c1 ? 1 : taps;
  L->cin_gemm = special_c1 ? 64 : cin;
  L->kdim = L->taps * L->cin_gemm;
  L->block_n = conv_igemm_block_n(cout);
  L->cout_pad = div_up(cout, L->block_n) * L->block_n;
  L->use_t = ((k == 3 || L->generic) && !special_c1 && cout <= 128 && cout % 64 == 0);
  L->macs_per_pixel = static_cast<double>(taps) * cin * cout;
  if (L->use_t) L->cout_pad = div_up(cout, 128) * 128;   // weight rows padded to the M block
  std::vector<uint16_t> wp(static_cast<size_t>(L->cout_pad) * L->kdim, 0);
  std::vector<float> bp(L->cout_pad, 0.f);
  for (int co = 0; co < cout; ++co) {
    for (int t = 0; t < taps; ++t)
      for (int ci = 0; ci < cin; ++ci) {
        const float v = w_hwio[(static_cast<size_t>(t) * cin + ci) * cout + co] * scale[co];
        const uint16_t q = f2bf(v);
        if (special_c1) {
          wp[static_cast<size_t>(co) * 64 + t * cin + ci] = q;             // hi part
          wp[static_cast<size_t>(co) * 64 + 9 * cin + t * cin + ci] = q;   // lo part
        } else {
          wp[static_cast<size_t>(co) * L->kdim + t * cin + ci] = q;
        }
      }
    bp[co] = (bias ? bias[co] : 0.f) * scale[co] + shift[co];
  }
  XV_TRY(L->w_packed.upload(wp));
  XV_TRY(L->bias_pad.upload(bp));
  return 0;
}

// Descriptors are cached per network, keyed by address and geometry.  Activations live in the
// network's arena (stable addresses); caller-owned inputs may move, so the cache is bounded.
static void cache_tmap(xv_fcn* net, const std::array<long long, 9>& key, const CUtensorMap& m) {
  if (net->tmaps.size() >= 4096) net->tmaps.clear();
  net->tmaps[key] = m;
}

int get_tmap_ex(xv_fcn* net, CUtensorMap* out, const void* ptr, int N, int H, int W, int C,
                int pitch, int sample, int th, int tw) {
  const std::array<long long, 9> key = {static_cast<long long>(reinterpret_cast<uintptr_t>(ptr)),
                                        N, H, W, C, pitch, sample, th, tw};
  if (net) {
    auto it = net->tmaps.find(key);
    if (it != net->tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  XV_TRY(make_tmap_act(out, ptr, N, H, W, C, pitch, sample, th, tw));
  if (net) cache_tmap(net, key, *out);
  return 0;
}

// Every second row of a bf16 [N,H,W,C] activation (parity 0: rows 0, 2, ..; 1: rows 1, 3, ..)
// viewed as (C, W, rows, N), box {64, 16, 17, 1}: the operand patches of the row-pair kernel.
static int get_tmap_rows2(xv_fcn* net, CUtensorMap* out, const void* ptr, int N, int H, int W,
                          int C, int parity) {
  const std::array<long long, 9> key = {static_cast<long long>(reinterpret_cast<uintptr_t>(ptr)),
                                        N, H, W, C, -2, parity, 17, 16};
  if (net) {
    auto it = net->tmaps.find(key);
    if (it != net->tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  const cuuint64_t px = static_cast<cuuint64_t>(C) * 2;
  const char* base = static_cast<const char*>(ptr) + (parity ? px * W : 0);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W),
                        static_cast<cuuint64_t>((H + 1 - parity) / 2), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {px, 2 * px * W, px * W * H};
  cuuint32_t box[4] = {64, 16, 17, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(row parity) failed: " + std::to_string(r));
  if (net) cache_tmap(net, key, *out);
  return 0;
}

int get_tmap(xv_fcn* net, CUtensorMap* out, const void* ptr, int N, int H, int W, int C, int th,
             int tw) {
  return get_tmap_ex(net, out, ptr, N, H, W, C, C, 1, th, tw);
}

int get_tmap_w(xv_fcn* net, CUtensorMap* out, const void* ptr, int kdim, int cout_pad,
               int block_n) {
  const std::array<long long, 9> key = {static_cast<long long>(reinterpret_cast<uintptr_t>(ptr)),
                                        kdim, cout_pad, block_n, -1, -1, -1, -1, -1};
  if (net) {
    auto it = net->tmaps.find(key);
    if (it != net->tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  XV_TRY(make_tmap_w(out, ptr, kdim, cout_pad, block_n));
  if (net) cache_tmap(net, key, *out);
  return 0;
}

// Whether run_igemm sends this 3x3 layer to the CTA-pair kernel (the one with the fused pool).
static bool pair_kernel_applies(const ConvLayer& L, int H, int W) {
  int th, tw;
  choose_tile(H, W, &th, &tw);
  const bool halo = L.taps == 9 && !(g_debug_flags & 64) && (tw == 8 || tw == 16) &&
                    (th + 2) * tw * 128 <= 20480;
  return halo && L.block_n == 256 && !L.use_t && !(g_debug_flags & 128);
}

// in: bf16 [B,H,W,cin_gemm]; out: bf16 [B,H,W,cout] (cout % 64 == 0) or fp32 [B,H,W,cout].
// pool_mode 1 / 2 (CTA-pair kernel only, see pair_kernel_applies): the epilogue also (2) or only
// (1, `out` unused) stores the 2x2 max-pooled result into pool_out [B,H/2,W/2,cout].
static int run_igemm(xv_fcn* net, const ConvLayer& L, const void* in, int B, int H, int W,
                     void* out, bool out_f32, cudaStream_t s, void* pool_out = nullptr,
                     int pool_mode = 0) {
  if (L.use_t && !out_f32 && !(g_debug_flags & 2))
    return run_igemm_t(net, L, in, B, H, W, out, false, s);
  ConvIgemmParams p;
  std::memset(&p, 0, sizeof(p));
  choose_tile(H, W, &p.th, &p.tw);
  // halo variant (debug bit6 selects the nine-shifted-tiles one): three (th + 2) x tw patch copies
  p.halo = (L.taps == 9 && !out_f32 && !(g_debug_flags & 64) && (p.tw == 8 || p.tw == 16) &&
            (p.th + 2) * p.tw * 128 <= 20480)
               ? 1
               : 0;
  XV_TRY(get_tmap(net, &p.tmap_in, in, B, H, W, L.cin_gemm, p.halo ? p.th + 2 : p.th, p.tw));
  // CTA-pair kernel (cta_group::2) for the 256-wide layers; debug bit7 keeps the single-CTA one
  const bool pair = p.halo && L.block_n == 256 && !(g_debug_flags & 128);
  XV_TRY(get_tmap_w(net, &p.tmap_w, L.w_packed.p, L.kdim, L.cout_pad, pair ? 128 : L.block_n));
  if (pool_mode) {
    XV_CHECK(pair && pool_out != nullptr, "fused pooling needs the CTA-pair kernel");
    p.pool_mode = pool_mode;
    XV_TRY(get_tmap(net, &p.tmap_pool, pool_out, B, H / 2, W / 2, L.cout, p.th / 2, p.tw / 2));
  }
  if (!out_f32 && pool_mode == 1) {
    p.tmap_out = p.tmap_pool;          // never dereferenced in that mode
  } else if (!out_f32) {
    XV_CHECK(L.cout % 64 == 0, "bf16 epilogue needs Cout % 64 == 0");
    XV_TRY(get_tmap(net, &p.tmap_out, out, B, H, W, L.cout, p.th, p.tw));
  } else {
    p.tmap_out = p.tmap_in;
    p.out_f32 = static_cast<float*>(out);
  }
  p.bias = static_cast<const float*>(L.bias_pad.p);
  p.N = B;
  p.H = H;
  p.W = W;
  p.cin = L.cin_gemm;
  p.cout = L.cout;
  p.tiles_x = div_up(W, p.tw);
  p.tiles_y = div_up(H, p.th);
  p.n_blocks = L.cout_pad / L.block_n;
  p.relu = L.relu;
  auto launch = [&]() -> int {
    if (pair) return launch_conv_igemm_2cta(p, 256, s);
    return launch_conv_igemm(p, L.block_n, L.taps, out_f32, s);
  };
  if (!g_profile) return launch();
  IgemmSample smp;
  XV_CUDA(cudaEventCreate(&smp.e0));
  XV_CUDA(cudaEventCreate(&smp.e1));
  // algorithmic FLOPs of the layer (2 * MACs, real channels only)
  smp.flops = 2.0 * B * H * W * static_cast<double>(L.cout) * L.k * L.k * L.cin;
  smp.block_n = L.block_n;
  XV_CUDA(cudaEventRecord(smp.e0, s));
  const int rc = launch();
  XV_CUDA(cudaEventRecord(smp.e1, s));
  g_samples.push_back(smp);
  return rc;
}

static int run_igemm_t(xv_fcn* net, const ConvLayer& L, const void* in, int B, int H, int W,
                       void* out, bool pool, cudaStream_t s) {
  ConvIgemmParams p;
  std::memset(&p, 0, sizeof(p));
  // conv1_2 + pool1: both halves of the 128 accumulator lanes carry an output row
  // (conv_igemm_rowpair_sm100.cu); debug bit10 keeps the half-empty transposed-role kernel
  const bool rowpair_pool = pool && L.cout <= 64 && H % 2 == 0 && W % 2 == 0;
  const bool rowpair_full = !pool && L.cout == 64;      // fit() forward / data gradient of conv1_2
  if ((rowpair_pool || rowpair_full) && L.taps == 9 && L.cin_gemm == 64 &&
      !(g_debug_flags & (64 | 1024))) {
    XV_TRY(get_tmap_rows2(net, &p.tmap_in_par[0], in, B, H, W, 64, 0));
    XV_TRY(get_tmap_rows2(net, &p.tmap_in_par[1], in, B, H, W, 64, 1));
    XV_TRY(get_tmap_w(net, &p.tmap_w, L.w_packed.p, L.kdim, L.cout_pad, 64));
    if (pool)
      XV_TRY(get_tmap(net, &p.tmap_out, out, B, H / 2, W / 2, L.cout, 16, 8));
    else
      XV_TRY(get_tmap(net, &p.tmap_out, out, B, H, W, L.cout, 8, 16));
    p.bias = static_cast<const float*>(L.bias_pad.p);
    p.N = B;
    p.H = H;
    p.W = W;
    p.cin = 64;
    p.cout = L.cout;
    p.tiles_x = div_up(W, 16);
    p.tiles_y = div_up(H, 32);
    p.n_blocks = 1;
    p.relu = L.relu;
    if (!g_profile) return launch_conv_igemm_rowpair(p, pool, s);
    IgemmSample smp;
    XV_CUDA(cudaEventCreate(&smp.e0));
    XV_CUDA(cudaEventCreate(&smp.e1));
    smp.flops = 2.0 * B * H * W * static_cast<double>(L.cout) * L.k * L.k * L.cin;
    smp.block_n = 0;
    XV_CUDA(cudaEventRecord(smp.e0, s));
    const int rc = launch_conv_igemm_rowpair(p, pool, s);
    XV_CUDA(cudaEventRecord(smp.e1, s));
    g_samples.push_back(smp);
    return rc;
  }
  p.th = p.tw = 16;
  // debug bit6: previous variant (nine shifted 16x16 tiles per channel chunk instead of three
  // column-shifted 18x16 patches)
  p.halo = (g_debug_flags & 64) ? 0 : 1;
  XV_TRY(get_tmap(net, &p.tmap_in, in, B, H, W, L.cin_gemm, p.halo ? 18 : 16, 16));
  p.w_rows = L.cout <= 64 ? 64 : 128;
  XV_TRY(get_tmap_w(net, &p.tmap_w, L.w_packed.p, L.kdim, L.cout_pad, p.w_rows));
  if (pool) {
    XV_TRY(get_tmap(net, &p.tmap_out, out, B, H / 2, W / 2, L.cout, 8, 8));
  } else {
    XV_TRY(get_tmap(net, &p.tmap_out, out, B, H, W, L.cout, 8, 16));
  }
  p.bias = static_cast<const float*>(L.bias_pad.p);
  p.N = B;
  p.H = H;
  p.W = W;
  p.cin = L.cin_gemm;
  p.cout = L.cout;
  p.tiles_x = div_up(W, 16);
  p.tiles_y = div_up(H, 16);
  p.n_blocks = L.cout_pad / 128;
  p.relu = L.relu;
  if (!g_profile) return launch_conv_igemm_t(p, pool, s);
  IgemmSample smp;
  XV_CUDA(cudaEventCreate(&smp.e0));
  XV_CUDA(cudaEventCreate(&smp.e1));
  smp.flops = 2.0 * B * H * W * static_cast<double>(L.cout) * L.k * L.k * L.cin;
  smp.block_n = 0;
  XV_CUDA(cudaEventRecord(smp.e0, s));
  const int rc = launch_conv_igemm_t(p, pool, s);
  XV_CUDA(cudaEventRecord(smp.e1, s));
  g_samples.push_back(smp);
  return rc;
}

// conv1_1 on the raw fp32 input: operand rows are packed inside the kernel
int run_igemm_c1(xv_fcn* net, const ConvLayer& L, const float* x, int B, int H, int W, void* out,
                 cudaStream_t s) {
  ConvIgemmParams p;
  std::memset(&p, 0, sizeof(p));
  choose_tile(H, W, &p.th, &p.tw);
  // debug bit5: previous variant (operand rows built from global loads inside conv_igemm_kernel)
  const bool staged = !(g_debug_flags & 32) && (W * L.cin) % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      conv_c1_patch_fits(p.th, p.tw, L.cin, &p.patch_w);
  XV_TRY(get_tmap_w(net, &p.tmap_w, L.w_packed.p, L.kdim, L.cout_pad, L.block_n));
  XV_TRY(get_tmap(net, &p.tmap_out, out, B, H, W, L.cout, p.th, p.tw));
  if (staged) {
    const std::array<long long, 9> key = {
        static_cast<long long>(reinterpret_cast<uintptr_t>(x)), B, H, W, L.cin, p.patch_w,
        p.th + 2, -2, -2};
    bool cached = false;
    if (net) {
      auto it = net->tmaps.find(key);
      if (it != net->tmaps.end()) {
        p.tmap_in = it->second;
        cached = true;
      }
    }
    if (!cached) {
      XV_TRY(make_tmap_patch(&p.tmap_in, x, B, H, W, L.cin, p.patch_w, p.th + 2));
      if (net) cache_tmap(net, key, p.tmap_in);
    }
  } else {
    p.tmap_in = p.tmap_out;          // unused in that mode, kept valid for the descriptor prefetch
  }
  p.bias = static_cast<const float*>(L.bias_pad.p);
  p.x_raw = x;
  p.N = B;
  p.H = H;
  p.W = W;
  p.cin = 64;
  p.cout = L.cout;
  p.tiles_x = div_up(W, p.tw);
  p.tiles_y = div_up(H, p.th);
  p.n_blocks = 1;
  p.relu = L.relu;
  auto launch = [&]() -> int {
    return staged ? launch_conv_c1(p, L.cin, s) : launch_conv_igemm_c1(p, L.cin, s);
  };
  if (!g_profile) return launch();
  IgemmSample smp;
  XV_CUDA(cudaEventCreate(&smp.e0));
  XV_CUDA(cudaEventCreate(&smp.e1));
  smp.flops = 2.0 * B * H * W * static_cast<double>(L.cout) * L.k * L.k * L.cin;
  smp.block_n = 64;
  XV_CUDA(cudaEventRecord(smp.e0, s));
  const int rc = launch();
  XV_CUDA(cudaEventRecord(smp.e1, s));
  g_samples.push_back(smp);
  return rc;
}

// Adapnet layers: any square filter with dilation / padding offset, optional stride-2 sampling
// of the input and channel-sliced input / output tensors (concat without a copy).
int run_conv_generic(xv_fcn* net, const ConvLayer& L, const void* in, int B, int H, int W,
                     int in_pitch, void* out, int out_pitch, bool out_f32, cudaStream_t s,
                     const void* residual) {
  const int step = L.stride2 ? 2 : L.sample;
  const int Ho = div_up(H, step), Wo = div_up(W, step);
  const bool t_kernel = L.use_t && !out_f32 && (L.stride2 || !(g_debug_flags & 2));
  XV_CHECK(!L.stride2 || (t_kernel && H % 2 == 0 && W % 2 == 0),
           "stride-2 filters run on the transposed-role kernel (Cout <= 128) on even sizes");
  XV_CHECK(residual == nullptr || (!t_kernel && !out_f32 && L.block_n == 256),
           "the residual epilogue needs the pixel-major bf16 kernel with BLOCK_N = 256");
  ConvIgemmParams p;
  std::memset(&p, 0, sizeof(p));
  if (t_kernel) {
    p.th = p.tw = 16;
  } else {
    choose_tile(Ho, Wo, &p.th, &p.tw);
  }
  if (L.stride2) {
    // parity (vy, vx): pixels (2i + vy, 2j + vx) of the input
    for (int v = 0; v < 4; ++v) {
      const char* base = static_cast<const char*>(in) +
                         (static_cast<size_t>(v >> 1) * W + (v & 1)) * in_pitch * 2;
      XV_TRY(get_tmap_ex(net, &p.tmap_in_par[v], base, B, H, W, L.cin_gemm, in_pitch, 2, p.th,
                         p.tw));
    }
    p.tmap_in = p.tmap_in_par[0];
    p.stride2 = 1;
  } else {
    XV_TRY(get_tmap_ex(net, &p.tmap_in, in, B, H, W, L.cin_gemm, in_pitch, L.sample, p.th, p.tw));
  }
  if (residual) {
    p.has_residual = 1;
    XV_TRY(get_tmap_ex(net, &p.tmap_res, residual, B, Ho, Wo, L.cout, L.cout, 1, p.th, p.tw));
  }
  p.w_rows = L.cout <= 64 ? 64 : 128;
  XV_TRY(get_tmap_w(net, &p.tmap_w, L.w_packed.p, L.kdim, L.cout_pad,
                    t_kernel ? p.w_rows : L.block_n));
  if (out_f32) {
    p.tmap_out = p.tmap_in;
    p.out_f32 = static_cast<float*>(out);
  } else if (t_kernel) {
    XV_TRY(get_tmap_ex(net, &p.tmap_out, out, B, Ho, Wo, L.cout, out_pitch, 1, 8, 16));
  } else {
    XV_CHECK(L.cout % 8 == 0, "bf16 epilogue needs Cout % 8 == 0");
    XV_TRY(get_tmap_ex(net, &p.tmap_out, out, B, Ho, Wo, L.cout, out_pitch, 1, p.th, p.tw));
  }
  p.bias = static_cast<const float*>(L.bias_pad.p);
  p.N = B;
  p.H = Ho;
  p.W = Wo;
  p.cin = L.cin_gemm;
  p.cout = L.cout;
  p.tiles_x = div_up(Wo, p.tw);
  p.tiles_y = div_up(Ho, p.th);
  p.n_blocks = L.cout_pad / (t_kernel ? 128 : L.block_n);
  p.relu = L.relu;
  p.taps = L.taps;
  p.kw = L.k;
  p.dil = L.dil;
  p.pad = L.pad;
  auto launch = [&]() -> int {
    return t_kernel ? launch_conv_igemm_t_generic(p, s)
                    : launch_conv_igemm(p, L.block_n, 0, out_f32, s);
  };
  if (!g_profile) return launch();
  IgemmSample smp;
  XV_CUDA(cudaEventCreate(&smp.e0));
  XV_CUDA(cudaEventCreate(&smp.e1));
  smp.flops = 2.0 * B * Ho * Wo * L.macs_per_pixel;
  smp.block_n = t_kernel ? 0 : L.block_n;
  XV_CUDA(cudaEventRecord(smp.e0, s));
  const int rc = launch();
  XV_CUDA(cudaEventRecord(smp.e1, s));
  g_samples.push_back(smp);
  return rc;
}

// ------------------------------------------------------------------ forward schedule
struct Forward {
  xv_fcn* net;
  Arena arena;
  cudaStream_t s;
  bool dry;
  int T = 1;                 // copies made at the first active dropout site (MC samples + the
                             // optional dropout-free leading sample)
  const xv_dropout_cfg* drop = nullptr;
  bool training = false;     // keep every activation (no pool fusion), materialise score + up5
  int N0 = 0;                // images of the call (set by run)
  bool keep_first() const { return drop && (drop->flags & XV_DROP_FLAG_KEEP_FIRST); }
  int mc_samples() const { return keep_first() ? T - 1 : T; }

  bool bf16() const { return net->precision == XV_PRECISION_BF16; }
  size_t esize(DType d) const { return d == DType::F32 ? 4 : (d == DType::BF16 ? 2 : 1); }

  Act make(const std::string& name, DType dt, int B, int H, int W, int C) {
    Act a;
    a.dt = dt;
    a.B = B;
    a.H = H;
    a.W = W;
    a.C = C;
    a.p = arena.alloc(a.elems() * esize(dt));
    if (!name.empty()) net->layers[name] = a;   // the dry pass records shapes (null pointers)
    return a;
  }

  int conv(const std::string& name, const Act& in, Act* out, bool force_f32_out = false) {
    ConvLayer* L = net->conv(name);
    XV_CHECK(L != nullptr, "unknown conv layer " + name);
    if (bf16()) {
      const bool f32o = force_f32_out;
      *out = make(name, f32o ? DType::F32 : DType::BF16, in.B, in.H, in.W, L->cout);
      if (dry) return 0;
      return run_igemm(net, *L, in.p, in.B, in.H, in.W, out->p, f32o, s);
    }
    *out = make(name, DType::F32, in.B, in.H, in.W, L->cout);
    if (dry) return 0;
    const bool relu_in_conv = L->relu && !L->has_bn;
    XV_TRY(launch_conv_f32(static_cast<const float*>(in.p), static_cast<const float*>(L->w_f32.p),
                           static_cast<const float*>(L->bias_f32.p), static_cast<float*>(out->p),
                           in.B, in.H, in.W, L->cin, L->cout, L->k, relu_in_conv, s));
    if (L->has_bn)
      XV_TRY(launch_affine_f32(static_cast<float*>(out->p),
                               static_cast<const float*>(L->bn_scale.p),
                               static_cast<const float*>(L->bn_shift.p),
                               static_cast<size_t>(in.B) * in.H * in.W, L->cout, L->relu, s));
    return 0;
  }

  // conv followed by the 2x2 max pool of simple_fcn.py:41,44,48,58 - fused into the conv epilogue
  // of the transposed-role kernel (conv1_2, conv2_2) or of the CTA-pair kernel (conv3_3, conv4_3).
  // `full_out` == nullptr: the unpooled activation is not needed and, when fused, never
  // materialised; otherwise it is stored as well (conv4_3 also feeds score_conv4).
  int conv_pool(const std::string& name, const std::string& pool_name, const Act& in, Act* out,
                Act* full_out = nullptr) {
    ConvLayer* L = net->conv(name);
    XV_CHECK(L != nullptr, "unknown conv layer " + name);
    if (bf16() && L->use_t && !(g_debug_flags & 3) && !training && !full_out) {
      *out = make(pool_name, DType::BF16, in.B, in.H / 2, in.W / 2, L->cout);
      if (dry) return 0;
      return run_igemm_t(net, *L, in.p, in.B, in.H, in.W, out->p, true, s);
    }
    if (bf16() && !(g_debug_flags & 1) && !training && in.H % 2 == 0 && in.W % 2 == 0 &&
        pair_kernel_applies(*L, in.H, in.W)) {
      if (full_out) *full_out = make(name, DType::BF16, in.B, in.H, in.W, L->cout);
      *out = make(pool_name, DType::BF16, in.B, in.H / 2, in.W / 2, L->cout);
      if (dry) return 0;
      return run_igemm(net, *L, in.p, in.B, in.H, in.W, full_out ? full_out->p : nullptr, false, s,
                       out->p, full_out ? 2 : 1);
    }
    Act full;
    XV_TRY(conv(name, in, &full));
    if (full_out) *full_out = full;
    return pool(pool_name, full, out);
  }

  int pool(const std::string& name, const Act& in, Act* out) {
    *out = make(name, in.dt, in.B, in.H / 2, in.W / 2, in.C);
    if (dry) return 0;
    if (in.dt == DType::BF16)
      return launch_maxpool_bf16(static_cast<const __nv_bfloat16*>(in.p),
                                 static_cast<__nv_bfloat16*>(out->p), in.B, in.H, in.W, in.C, s);
    return launch_maxpool_f32(static_cast<const float*>(in.p), static_cast<float*>(out->p), in.B,
                              in.H, in.W, in.C, s);
  }

  // dropout site `idx` (0 pool3, 1 pool4, 2 conv4_3, 3 conv5_3, 4 features); `active` false
  // with replicate > 1 degenerates to a plain T-fold copy.
  int dropout(const std::string& name, int idx, bool active, const Act& in, int replicate,
              Act* out) {
    *out = make(name, in.dt, in.B * replicate, in.H, in.W, in.C);
    if (dry) return 0;
    DropoutSpec d;
    d.rate = active ? drop->rate : 0.f;
    d.ext_mask = active ? drop->ext_mask[idx] : nullptr;
    d.seed = drop ? drop->seed : 0;
    d.offset = static_cast<uint64_t>(idx + 1) << 40;
    // the dropout-free leading sample: the first N0 images of the (replicated) output
    if (keep_first() && active)
      d.pass_elems = in.elems() / static_cast<size_t>(in.B) * static_cast<size_t>(N0);
    if (!active) {
      static const uint8_t* none = nullptr;
      d.ext_mask = none;
    }
    if (in.dt == DType::BF16)
      return launch_dropout_bf16(static_cast<const __nv_bfloat16*>(in.p),
                                 static_cast<__nv_bfloat16*>(out->p), in.elems(), replicate, d, s);
    return launch_dropout_f32(static_cast<const float*>(in.p), static_cast<float*>(out->p),
                              in.elems(), replicate, d, s);
  }

  int run(const float* x, int N, int H, int W, const xv_fcn_outputs* o);
  int run_encoder(const float* x, int N, int H, int W, Act* c43_out, Act* c53_out,
                  bool* replicated_out);
  int run_head(Act c43, Act c53, bool replicated, const xv_fcn_outputs* o);
};

int Forward::run(const float* x, int N, int H, int W, const xv_fcn_outputs* o) {
  N0 = N;
  Act c43, c53;
  bool replicated = false;
  XV_TRY(run_encoder(x, N, H, W, &c43, &c53, &replicated));
  if (net->role == 1) return 0;
  return run_head(c43, c53, replicated, o);
}

int Forward::run_encoder(const float* x, int N, int H, int W, Act* c43_out, Act* c53_out,
                         bool* replicated_out) {
  const uint32_t sites = drop ? drop->sites : 0u;
  Act cur, t;
  if (bf16() && !(g_debug_flags & 4)) {
    // conv1_1 straight from the raw fp32 input (operand packing fused into the GEMM producer)
    ConvLayer* L = net->conv("conv1_1");
    t = make("conv1_1", DType::BF16, N, H, W, L->cout);
    if (!dry) XV_TRY(run_igemm_c1(net, *L, x, N, H, W, t.p, s));
  } else if (bf16()) {
    Act a0 = make("conv1_1_operand", DType::BF16, N, H, W, 64);
    if (!dry)
      XV_TRY(launch_im2col_c1(x, static_cast<__nv_bfloat16*>(a0.p), N, H, W, net->cin, s));
    XV_TRY(conv("conv1_1", a0, &t));
  } else {
    cur.p = const_cast<float*>(x);
    cur.dt = DType::F32;
    cur.B = N;
    cur.H = H;
    cur.W = W;
    cur.C = net->cin;
    XV_TRY(conv("conv1_1", cur, &t));
  }
  XV_TRY(conv_pool("conv1_2", "pool1", t, &cur));
  t = cur;
  XV_TRY(conv("conv2_1", t, &cur));
  XV_TRY(conv_pool("conv2_2", "pool2", cur, &t));
  cur = t;
  XV_TRY(conv("conv3_1", cur, &t));
  XV_TRY(conv("conv3_2", t, &cur));
  XV_TRY(conv_pool("conv3_3", "pool3", cur, &t));
  cur = t;
  bool replicated = false;
  if (sites & XV_DROP_POOL3) {
    XV_TRY(dropout("pool3_drop", 0, true, cur, T, &t));
    cur = t;
    replicated = true;
  }
  XV_TRY(conv("conv4_1", cur, &t));
  XV_TRY(conv("conv4_2", t, &cur));
  Act c43;
  XV_TRY(conv_pool("conv4_3", "pool4", cur, &t, &c43));
  cur = t;
  if (sites & XV_DROP_POOL3) {   // sic: simple_fcn.py:61 gates pool4 dropout on 'pool3'
    XV_TRY(dropout("pool4_drop", 1, true, cur, 1, &t));
    cur = t;
  }
  XV_TRY(conv("conv5_1", cur, &t));
  XV_TRY(conv("conv5_2", t, &cur));
  Act c53;
  XV_TRY(conv("conv5_3", cur, &c53));
  *c43_out = c43;
  *c53_out = c53;
  *replicated_out = replicated;
  return 0;
}

int Forward::run_head(Act c43, Act c53, bool replicated, const xv_fcn_outputs* o) {
  const uint32_t sites = drop ? drop->sites : 0u;
  const int Tmc = mc_samples();
  const bool mc = Tmc > 1;
  // with a dropout-free leading sample the per-image outputs describe that sample only and
  // the moments skip it
  const bool lead = keep_first() && N0 > 0;
  Act t;
  Act s4in = c43, s5in = c53;
  const bool branch_sites = (sites & (XV_DROP_CONV4_3 | XV_DROP_CONV5_3)) != 0;
  if (branch_sites) {
    const int rep = replicated ? 1 : T;
    if ((sites & XV_DROP_CONV4_3) || rep > 1)
      XV_TRY(dropout("conv4_3_drop", 2, (sites & XV_DROP_CONV4_3) != 0, c43, rep, &s4in));
    if ((sites & XV_DROP_CONV5_3) || rep > 1)
      XV_TRY(dropout("conv5_3_drop", 3, (sites & XV_DROP_CONV5_3) != 0, c53, rep, &s5in));
    replicated = true;
  }
  Act s4, s5, fused;
  XV_TRY(conv("score_conv4", s4in, &s4, /*force_f32_out=*/true));
  XV_TRY(conv("score_conv5", s5in, &s5, /*force_f32_out=*/true));
  const int nu = net->nu;
  // upscore_conv5 + skip add (simple_fcn.py:82-85)
  fused = make("fused", DType::F32, s4.B, s4.H, s4.W, nu);
  // inference with both bilinear fast paths and no dropout on the features: upscore_conv5 + skip
  // add + the (moved) 1x1 score conv run as ONE kernel further down
  const bool head_fused = net->fast_up5 && net->fast_up && !training &&
                          !(sites & XV_DROP_FEATURES) && head_fused_supported(nu, net->C);
  if (head_fused) {
    // launched below, once `low` exists
  } else if (net->fast_up5) {
    Act up5;
    if (training) up5 = make("upscore_conv5", DType::F32, s4.B, s4.H, s4.W, nu);
    if (!dry)
      XV_TRY(launch_upscore2_add(static_cast<const float*>(s5.p), static_cast<const float*>(s4.p),
                                 static_cast<const float*>(net->g4.p),
                                 static_cast<float*>(fused.p), s5.B, s5.H, s5.W, nu, s,
                                 training ? static_cast<float*>(up5.p) : nullptr));
  } else if (!net->bn_all()) {
    if (!dry)
      XV_TRY(launch_deconv_f32(static_cast<const float*>(s5.p),
                               static_cast<const float*>(net->w_up5.p),
                               static_cast<float*>(fused.p), s5.B, s5.H, s5.W, nu, nu, 4, 2, 1,
                               static_cast<const float*>(s4.p), s));
  } else {
    // deconv -> BN -> ReLU (custom_layers.py:112-119), then the skip add
    Act up5 = make("upscore_conv5", DType::F32, s4.B, s4.H, s4.W, nu);
    if (!dry) {
      XV_TRY(launch_deconv_f32(static_cast<const float*>(s5.p),
                               static_cast<const float*>(net->w_up5.p),
                               static_cast<float*>(up5.p), s5.B, s5.H, s5.W, nu, nu, 4, 2, 0,
                               nullptr, s));
      XV_TRY(launch_affine_f32(static_cast<float*>(up5.p),
                               static_cast<const float*>(net->up5_scale.p),
                               static_cast<const float*>(net->up5_shift.p), up5.elems() / nu, nu,
                               1, s));
      XV_TRY(launch_add_f32(static_cast<const float*>(s4.p), static_cast<const float*>(up5.p),
                            static_cast<float*>(fused.p), fused.elems(), s));
    }
  }
  Act feat = fused;
  if (sites & XV_DROP_FEATURES) {
    XV_TRY(dropout("features_drop", 4, true, fused, replicated ? 1 : T, &feat));
    replicated = true;
  }
  const int B = feat.B;
  const int Hf = feat.H * 8, Wf = feat.W * 8;
  const int C = net->C;
  const size_t npix = static_cast<size_t>(B) * Hf * Wf;
  const bool want_samples = o->score || o->prob || o->label_i64 || o->label_u8;
  const bool want_moments = mc && (o->mean_prob || o->var_prob || o->mean_var);

  if (net->fast_up) {
    Act low = make("score_lowres", DType::F32, B, feat.H, feat.W, C);
    xv_fcn_outputs train_out;
    if (training) {
      // fit(): the loss head works on the low-resolution scores (loss_lowres_grad_kernel); the
      // full-resolution score tensor only exists for the three-kernel form behind debug bit12
      std::memset(&train_out, 0, sizeof(train_out));
      if (g_debug_flags & 4096) {
        Act sc = make("train_score", DType::F32, B, Hf, Wf, C);
        train_out.score = static_cast<float*>(sc.p);
      }
      o = &train_out;
    }
    const bool want_samples = o->score || o->prob || o->label_i64 || o->label_u8;
    const int b_samples = lead ? N0 : B;
    const size_t lead_low = lead ? static_cast<size_t>(N0) * feat.H * feat.W * C : 0;
    if (!dry) {
      if (head_fused)
        XV_TRY(launch_head_fused(static_cast<const float*>(s5.p), static_cast<const float*>(s4.p),
                                 static_cast<const float*>(net->g4.p),
                                 static_cast<const float*>(net->w_score_nuxc.p),
                                 static_cast<float*>(fused.p), static_cast<float*>(low.p), s5.B,
                                 s5.H, s5.W, nu, C, s));
      else
        XV_TRY(launch_score_lowres(static_cast<const float*>(feat.p),
                                   static_cast<const float*>(net->w_score_nuxc.p),
                                   static_cast<float*>(low.p),
                                   static_cast<size_t>(B) * feat.H * feat.W, nu, C, s));
      if (want_samples) {
        DecodeOut d;
        d.label_u8 = o->label_u8;
        d.label_i64 = o->label_i64;
        d.prob = o->prob;
        d.score = o->score;
        XV_TRY(launch_decode_upsample8(static_cast<const float*>(low.p),
                                       static_cast<const float*>(net->g16.p),
                                       static_cast<const float*>(net->b_score.p), b_samples,
                                       feat.H, feat.W, C, d, s));
      }
      if (want_moments)
        XV_TRY(launch_decode_upsample8_mc(static_cast<const float*>(low.p) + lead_low,
                                          static_cast<const float*>(net->g16.p),
                                          static_cast<const float*>(net->b_score.p), Tmc,
                                          (B - (lead ? N0 : 0)) / Tmc, feat.H, feat.W, C,
                                          o->mean_prob, o->var_prob, o->mean_var, s));
    }
    return 0;
  }

  // batch-normalised decoder on a bilinear upscore kernel (fusion_fcn's head, batchnorm=True
  // experts): one fused pass, the upsampled num_units-channel tensor is never materialised
  if (bf16() && net->diag_up && net->bn_decoder() && !want_moments && !training &&
      decode_bn_supported(nu, C)) {
    ConvLayer* L = net->conv("score");
    if (!dry && want_samples) {
      DecodeOut d;
      d.label_u8 = o->label_u8;
      d.label_i64 = o->label_i64;
      d.prob = o->prob;
      d.score = o->score;
      const int b_samples = lead ? N0 : B;
      XV_TRY(launch_decode_bn_upsample8(
          static_cast<const float*>(feat.p), static_cast<const float*>(net->g16.p),
          static_cast<const float*>(net->up_scale.p), static_cast<const float*>(net->up_shift.p),
          static_cast<const float*>(net->w_score_nuxc.p), static_cast<const float*>(net->b_score.p),
          static_cast<const float*>(L->bn_scale.p), static_cast<const float*>(L->bn_shift.p),
          b_samples, feat.H, feat.W, nu, C, d, s));
    }
    return 0;
  }

  // generic decoder, reference op order: upscore (dense transposed conv) -> [BN] -> ReLU ->
  // 1x1 score -> softmax -> argmax (simple_fcn.py:129-133, basic_fusion_model.py:21-22)
  Act up = make("upscore", DType::F32, B, Hf, Wf, nu);
  Act sc;
  const bool bn = net->bn_decoder();
  if (!dry) {
    XV_TRY(launch_deconv_f32(static_cast<const float*>(feat.p),
                             static_cast<const float*>(net->w_up.p), static_cast<float*>(up.p), B,
                             feat.H, feat.W, nu, nu, 16, 8, bn ? 0 : 1, nullptr, s));
    if (bn)
      XV_TRY(launch_affine_f32(static_cast<float*>(up.p),
                               static_cast<const float*>(net->up_scale.p),
                               static_cast<const float*>(net->up_shift.p), npix, nu, 1, s));
  }
  {
    // the final 1x1 score conv always runs in fp32 on the CUDA cores in this path
    ConvLayer* L = net->conv("score");
    sc = make("score", DType::F32, B, Hf, Wf, C);
    if (!dry) {
      XV_TRY(launch_conv_f32(static_cast<const float*>(up.p),
                             static_cast<const float*>(L->w_f32.p),
                             static_cast<const float*>(L->bias_f32.p), static_cast<float*>(sc.p),
                             B, Hf, Wf, nu, C, 1, 0, s));
      if (L->has_bn)
        XV_TRY(launch_affine_f32(static_cast<float*>(sc.p),
                                 static_cast<const float*>(L->bn_scale.p),
                                 static_cast<const float*>(L->bn_shift.p), npix, C, 0, s));
    }
  }
  Act prob_tmp;
  float* prob_ptr = o->prob;
  if (want_moments && (!prob_ptr || lead)) {
    prob_tmp = make("prob_samples", DType::F32, B, Hf, Wf, C);
    prob_ptr = static_cast<float*>(prob_tmp.p);
  }
  if (!dry) {
    const size_t npix_lead = lead ? static_cast<size_t>(N0) * Hf * Wf : 0;
    const size_t npix_out = lead ? npix_lead : npix;
    if (o->score)
      XV_CUDA(cudaMemcpyAsync(o->score, sc.p, npix_out * C * sizeof(float),
                              cudaMemcpyDeviceToDevice, s));
    if (lead) {
      // dropout-free sample -> the per-image outputs; dropout samples -> scratch probabilities
      if (want_samples)
        XV_TRY(launch_softmax_argmax(static_cast<const float*>(sc.p), npix_lead, C, o->prob,
                                     o->label_i64, o->label_u8, s));
      if (want_moments) {
        float* mc_prob = static_cast<float*>(prob_tmp.p);
        XV_TRY(launch_softmax_argmax(static_cast<const float*>(sc.p) + npix_lead * C,
                                     npix - npix_lead, C, mc_prob, nullptr, nullptr, s));
        XV_TRY(launch_mc_moments(mc_prob, Tmc, (npix - npix_lead) / Tmc, C, o->mean_prob,
                                 o->var_prob, o->mean_var, nullptr, nullptr, nullptr, s));
      }
      return 0;
    }
    if (want_samples || want_moments)
      XV_TRY(launch_softmax_argmax(static_cast<const float*>(sc.p), npix, C, prob_ptr,
                                   o->label_i64, o->label_u8, s));
    if (want_moments)
      XV_TRY(launch_mc_moments(prob_ptr, T, npix / T, C, o->mean_prob, o->var_prob, o->mean_var,
                               nullptr, nullptr, nullptr, s));
  }
  return 0;
}

}  // namespace xv

// =================================================================== extern "C"
#define XV_STREAM(s) reinterpret_cast<cudaStream_t>(s)
extern "C" int xv_fcn_train_end(xv_fcn* net);

extern "C" {

int xv_abi_version(void) { return XV_ABI_VERSION; }
const char* xv_last_error(void) { return g_error.c_str(); }

int xv_init(int device) {
  XV_CUDA(cudaSetDevice(device));
  g_dev.device = -1;
  return ensure_init();
}
int xv_device_sm_count(int* out) {
  XV_TRY(ensure_init());
  *out = g_dev.num_sms;
  return 0;
}

int xv_set_debug_flags(int flags) {
  g_debug_flags = flags;
  return 0;
}
int xv_launch_count(int64_t* out) {
  *out = g_launches;
  return 0;
}
int xv_profile_enable(int on) {
  for (auto& smp : g_samples) {
    cudaEventDestroy(smp.e0);
    cudaEventDestroy(smp.e1);
  }
  g_samples.clear();
  g_profile = on != 0;
  return 0;
}
int xv_profile_read(double* ms_out, double* flops_out, int64_t* launches_out) {
  double ms = 0, flops = 0;
  for (auto& smp : g_samples) {
    XV_CUDA(cudaEventSynchronize(smp.e1));
    float t = 0.f;
    XV_CUDA(cudaEventElapsedTime(&t, smp.e0, smp.e1));
    ms += t;
    flops += smp.flops;
  }
  if (ms_out) *ms_out = ms;
  if (flops_out) *flops_out = flops;
  if (launches_out) *launches_out = static_cast<int64_t>(g_samples.size());
  return 0;
}

int xv_malloc(void** out, size_t bytes) {
  XV_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return 0;
}
int xv_free(void* p) {
  XV_CUDA(cudaFree(p));
  return 0;
}
int xv_malloc_host(void** out, size_t bytes) {
  XV_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return 0;
}
int xv_free_host(void* p) {
  XV_CUDA(cudaFreeHost(p));
  return 0;
}
int xv_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  XV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, XV_STREAM(stream)));
  return 0;
}
int xv_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  XV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, XV_STREAM(stream)));
  return 0;
}
int xv_memset(void* dst, int value, size_t bytes, void* stream) {
  XV_CUDA(cudaMemsetAsync(dst, value, bytes, XV_STREAM(stream)));
  return 0;
}
int xv_stream_sync(void* stream) {
  XV_CUDA(cudaStreamSynchronize(XV_STREAM(stream)));
  return 0;
}

// ------------------------------------------------------------------ FCN expert
int xv_fcn_create(xv_fcn** out, int cin, int num_units, int num_classes, int batchnorm,
                  int precision) {
  XV_CHECK(out != nullptr, "xv_fcn_create: out is NULL");
  XV_CHECK(cin >= 1 && num_units >= 1, "xv_fcn_create: bad channel counts");
  XV_CHECK(num_classes >= 2 && num_classes <= kMaxClasses,
           "xv_fcn_create: num_classes must be in [2, 24]");
  XV_CHECK(precision == XV_PRECISION_BF16 || precision == XV_PRECISION_FP32,
           "xv_fcn_create: unknown precision");
  if (precision == XV_PRECISION_BF16)
    XV_CHECK(cin <= 3, "xv_fcn_create: the bf16 path packs conv1_1 for Cin <= 3");
  xv_fcn* net = new xv_fcn();
  net->cin = cin;
  net->nu = num_units;
  net->C = num_classes;
  net->batchnorm = batchnorm;
  net->precision = precision;
  *out = net;
  return 0;
}

int xv_fcn_create_ex(xv_fcn** out, int cin, int num_units, int num_classes, int batchnorm,
                     int precision, int role, int head_cin) {
  XV_CHECK(role >= 0 && role <= 2, "xv_fcn_create_ex: role must be 0, 1 or 2");
  XV_CHECK(head_cin >= 64 && head_cin % 64 == 0, "xv_fcn_create_ex: head_cin must be a multiple of 64");
  XV_TRY(xv_fcn_create(out, role == 2 ? 1 : cin, num_units, num_classes, batchnorm, precision));
  (*out)->role = role;
  (*out)->head_cin = head_cin;
  return 0;
}

// Adapnet expert (adapnet.py:99-173): same handle type and the same set_param / finalize /
// forward / get_layer_host / destroy calls as the FCN expert.
int xv_adapnet_create(xv_fcn** out, int cin, int num_units, int num_classes, int precision) {
  XV_TRY(xv_fcn_create(out, cin, num_units, num_classes, 1, precision));
  (*out)->arch = 1;
  return 0;
}

// One VGG16 tower (role 1): runs conv1_1..conv5_3 and keeps conv4_3 / conv5_3 for the head.
int xv_fcn_forward_encoder(xv_fcn* net, const float* x, int n, int h, int w, void* stream) {
  XV_CHECK(net && x, "xv_fcn_forward_encoder: NULL argument");
  XV_CHECK(net->finalized && net->role == 1, "xv_fcn_forward_encoder: needs a finalized encoder");
  XV_CHECK(n >= 1 && h >= 16 && w >= 16 && h % 16 == 0 && w % 16 == 0,
           "xv_fcn_forward_encoder: H and W must be positive multiples of 16");
  xv_fcn_outputs none;
  std::memset(&none, 0, sizeof(none));
  Forward plan{net, Arena(), XV_STREAM(stream), true, 1, nullptr};
  XV_TRY(plan.run(x, n, h, w, &none));
  XV_TRY(net->arena_buf.ensure(plan.arena.off + 1024));
  Forward real{net, Arena(), XV_STREAM(stream), false, 1, nullptr};
  real.arena.base = static_cast<char*>(net->arena_buf.p);
  net->layers.clear();
  return real.run(x, n, h, w, &none);
}

// Mid-level fusion head (role 2, fusion_fcn.py:24-39): channel-concatenates conv4_3 / conv5_3 of
// the towers' LAST forward_encoder calls, then score_conv4/5 -> upscore_conv5 + add -> decoder.
int xv_fcn_forward_head(xv_fcn* head, xv_fcn* const* towers, int num_towers,
                        const xv_fcn_outputs* outputs, void* stream) {
  XV_CHECK(head && towers && outputs, "xv_fcn_forward_head: NULL argument");
  XV_CHECK(head->finalized && head->role == 2, "xv_fcn_forward_head: needs a finalized head");
  XV_CHECK(num_towers >= 1 && num_towers * 512 == head->head_cin,
           "xv_fcn_forward_head: head_cin must equal 512 * number of towers");
  cudaStream_t s = XV_STREAM(stream);
  Act c43s[4], c53s[4];
  XV_CHECK(num_towers <= 4, "xv_fcn_forward_head: at most 4 towers");
  for (int m = 0; m < num_towers; ++m) {
    auto i4 = towers[m]->layers.find("conv4_3");
    auto i5 = towers[m]->layers.find("conv5_3");
    XV_CHECK(i4 != towers[m]->layers.end() && i5 != towers[m]->layers.end() && i4->second.p,
             "xv_fcn_forward_head: run xv_fcn_forward_encoder on every tower first");
    c43s[m] = i4->second;
    c53s[m] = i5->second;
    XV_CHECK(c43s[m].B == c43s[0].B && c43s[m].H == c43s[0].H && c43s[m].W == c43s[0].W,
             "xv_fcn_forward_head: towers ran on different shapes");
  }
  for (int pass = 0; pass < 2; ++pass) {
    const bool dry = pass == 0;
    Forward f{head, Arena(), s, dry, 1, nullptr};
    if (!dry) {
      f.arena.base = static_cast<char*>(head->arena_buf.p);
      head->layers.clear();
    }
    Act cat4 = f.make("concat_conv4", DType::BF16, c43s[0].B, c43s[0].H, c43s[0].W, head->head_cin);
    Act cat5 = f.make("concat_conv5", DType::BF16, c53s[0].B, c53s[0].H, c53s[0].W, head->head_cin);
    if (!dry) {
      for (int m = 0; m < num_towers; ++m) {
        XV_TRY(launch_concat_bf16(static_cast<const __nv_bfloat16*>(c43s[m].p),
                                  static_cast<__nv_bfloat16*>(cat4.p), cat4.elems() / cat4.C, 512,
                                  head->head_cin, m * 512, s));
        XV_TRY(launch_concat_bf16(static_cast<const __nv_bfloat16*>(c53s[m].p),
                                  static_cast<__nv_bfloat16*>(cat5.p), cat5.elems() / cat5.C, 512,
                                  head->head_cin, m * 512, s));
      }
    }
    XV_TRY(f.run_head(cat4, cat5, false, outputs));
    if (dry) XV_TRY(head->arena_buf.ensure(f.arena.off + 1024));
  }
  return 0;
}

int xv_fcn_destroy(xv_fcn* net) {
  xv_fcn_train_end(net);
  delete net;
  return 0;
}

int xv_fcn_set_param_host(xv_fcn* net, const char* name, const float* data,
                          const int64_t* shape, int ndim) {
  XV_CHECK(net && name && data && shape, "xv_fcn_set_param_host: NULL argument");
  HostParam hp;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    hp.shape.push_back(shape[i]);
    n *= static_cast<size_t>(shape[i]);
  }
  hp.data.assign(data, data + n);
  net->params[name] = std::move(hp);
  net->finalized = false;
  return 0;
}

// A [k,k,Cout,Cin] transposed-conv kernel is "diagonal" when only [:,:,i,i] is non-zero.
static bool deconv_is_diagonal(const HostParam& w, int k, int nu) {
  for (int t = 0; t < k * k; ++t)
    for (int co = 0; co < nu; ++co)
      for (int ci = 0; ci < nu; ++ci)
        if (co != ci && w.data[(static_cast<size_t>(t) * nu + co) * nu + ci] != 0.f) return false;
  return true;
}

int xv_fcn_finalize(xv_fcn* net) {
  XV_CHECK(net != nullptr, "xv_fcn_finalize: NULL handle");
  XV_TRY(ensure_init());
  if (net->arch == 1) return adapnet_finalize(net);
  net->convs.clear();
  net->tmaps.clear();
  const int nu = net->nu, C = net->C;
  int cin = net->cin;
  std::vector<float> scale, shift;
  auto add_conv = [&](const std::string& name, int k, int ci, int co, int relu) -> int {
    const HostParam *w, *b;
    XV_TRY(get_param(net, name + "/kernel", {k, k, ci, co}, &w));
    XV_TRY(get_param(net, name + "/bias", {co}, &b));
    XV_TRY(bn_factors(net, name, co, &scale, &shift));
    std::unique_ptr<ConvLayer> L(new ConvLayer());
    L->name = name;
    L->k = k;
    L->cin = ci;
    L->cout = co;
    L->relu = relu;
    XV_TRY(pack_conv(net, L.get(), w->data.data(), b->data.data(), scale, shift, net->bn_all()));
    net->convs.push_back(std::move(L));
    return 0;
  };
  if (net->role != 2) {
    for (int i = 0; i < 13; ++i) {
      XV_TRY(add_conv(kConvNames[i], 3, cin, kConvCout[i], 1));
      cin = kConvCout[i];
    }
  }
  if (net->role == 1) {     // encoder only (one tower of fusion_fcn): no heads, no decoder
    net->finalized = true;
    return 0;
  }
  XV_TRY(add_conv("score_conv4", 1, net->head_cin, nu, 1));
  XV_TRY(add_conv("score_conv5", 1, net->head_cin, nu, 1));
  {
    // final 1x1 score conv: fp32 weights are always kept (generic decoder); the fast decoder
    // uses the [nu,C] matrix directly
    const HostParam *w, *b;
    XV_TRY(get_param(net, "score/kernel", {1, 1, nu, C}, &w));
    XV_TRY(get_param(net, "score/bias", {C}, &b));
    XV_TRY(bn_factors(net, "score", C, &scale, &shift));
    std::unique_ptr<ConvLayer> L(new ConvLayer());
    L->name = "score";
    L->k = 1;
    L->cin = nu;
    L->cout = C;
    L->relu = 0;
    XV_TRY(L->w_f32.upload(w->data));
    XV_TRY(L->bias_f32.upload(b->data));
    L->has_bn = net->bn_decoder();
    if (L->has_bn) {
      XV_TRY(L->bn_scale.upload(scale));
      XV_TRY(L->bn_shift.upload(shift));
    }
    net->convs.push_back(std::move(L));
    XV_TRY(net->w_score_nuxc.upload(w->data));
    XV_TRY(net->b_score.upload(b->data));
  }
  const HostParam *w5, *w16;
  XV_TRY(get_param(net, "upscore_conv5/kernel", {4, 4, nu, nu}, &w5));
  XV_TRY(get_param(net, "upscore/kernel", {16, 16, nu, nu}, &w16));
  XV_TRY(net->w_up5.upload(w5->data));
  XV_TRY(net->w_up.upload(w16->data));
  if (net->bn_all()) {
    XV_TRY(bn_factors(net, "upscore_conv5", nu, &scale, &shift));
    XV_TRY(net->up5_scale.upload(scale));
    XV_TRY(net->up5_shift.upload(shift));
  }
  if (net->bn_decoder()) {
    XV_TRY(bn_factors(net, "upscore", nu, &scale, &shift));
    XV_TRY(net->up_scale.upload(scale));
    XV_TRY(net->up_shift.upload(shift));
  }
  // Fast decoder paths exist in the bf16 production mode only; the fp32 validation mode keeps
  // the reference op order (dense transposed convolutions).
  net->fast_up5 = net->fast_up = net->diag_up5 = net->diag_up = false;
  if (net->precision == XV_PRECISION_BF16) {
    if (deconv_is_diagonal(*w5, 4, nu)) {
      std::vector<float> g(16 * nu);
      for (int t = 0; t < 16; ++t)
        for (int u = 0; u < nu; ++u) g[t * nu + u] = w5->data[(static_cast<size_t>(t) * nu + u) * nu + u];
      XV_TRY(net->g4.upload(g));
      net->diag_up5 = true;
      net->fast_up5 = !net->bn_all();
    }
    if (deconv_is_diagonal(*w16, 16, nu)) {
      // the 1x1 score conv commutes with the upsampling only if every channel shares one
      // non-negative 16x16 kernel (then ReLU after it is the identity on non-negative input)
      bool shared = true;
      std::vector<float> g(256);
      for (int t = 0; t < 256 && shared; ++t) {
        g[t] = w16->data[static_cast<size_t>(t) * nu * nu];
        if (g[t] < 0.f) shared = false;
        for (int u = 1; u < nu; ++u)
          if (w16->data[(static_cast<size_t>(t) * nu + u) * nu + u] != g[t]) shared = false;
      }
      if (shared) {
        XV_TRY(net->g16.upload(g));
        net->diag_up = true;
        net->fast_up = !net->bn_decoder();
      }
    }
  }
  net->finalized = true;
  return 0;
}

int xv_fcn_forward(xv_fcn* net, const float* x, int n, int h, int w, const xv_dropout_cfg* drop,
                   const xv_fcn_outputs* outputs, void* stream) {
  XV_CHECK(net && x && outputs, "xv_fcn_forward: NULL argument");
  XV_CHECK(net->finalized, "xv_fcn_forward: call xv_fcn_finalize first");
  XV_CHECK(n >= 1 && h >= 16 && w >= 16 && h % 16 == 0 && w % 16 == 0,
           "xv_fcn_forward: H and W must be positive multiples of 16");
  if (net->arch == 1) {
    XV_CHECK(!drop || drop->sites == 0, "xv_fcn_forward: adapnet has no dropout sites");
    return adapnet_forward(net, x, n, h, w, outputs, XV_STREAM(stream));
  }
  xv_dropout_cfg cfg;
  const xv_dropout_cfg* d = nullptr;
  int T = 1;
  if (drop && drop->sites != 0) {
    cfg = *drop;
    XV_CHECK(cfg.rate >= 0.f && cfg.rate < 1.f, "xv_fcn_forward: dropout rate must be in [0,1)");
    XV_CHECK(cfg.num_samples >= 1, "xv_fcn_forward: num_samples must be >= 1");
    T = cfg.num_samples + ((cfg.flags & XV_DROP_FLAG_KEEP_FIRST) ? 1 : 0);
    d = &cfg;
  }
  Forward plan{net, Arena(), XV_STREAM(stream), true, T, d};
  XV_TRY(plan.run(x, n, h, w, outputs));
  XV_TRY(net->arena_buf.ensure(plan.arena.off + 1024));
  Forward real{net, Arena(), XV_STREAM(stream), false, T, d};
  real.arena.base = static_cast<char*>(net->arena_buf.p);
  net->layers.clear();
  return real.run(x, n, h, w, outputs);
}

int xv_fcn_get_layer_host(xv_fcn* net, const char* layer, float* out_host, size_t capacity,
                          int64_t* shape_out, void* stream) {
  XV_CHECK(net && layer && shape_out, "xv_fcn_get_layer_host: NULL argument");
  auto it = net->layers.find(layer);
  XV_CHECK(it != net->layers.end(), std::string("no activation named '") + layer + "'");
  const Act& a = it->second;
  shape_out[0] = a.B;
  shape_out[1] = a.H;
  shape_out[2] = a.W;
  shape_out[3] = a.C;
  if (!out_host) return 0;
  XV_CHECK(capacity >= a.elems(), "xv_fcn_get_layer_host: buffer too small");
  cudaStream_t s = XV_STREAM(stream);
  if (a.dt == DType::F32) {
    XV_CUDA(cudaMemcpyAsync(out_host, a.p, a.elems() * 4, cudaMemcpyDeviceToHost, s));
  } else {
    DevBuf tmp;
    XV_TRY(tmp.ensure(a.elems() * 4));
    XV_TRY(launch_bf16_to_f32(static_cast<const __nv_bfloat16*>(a.p), static_cast<float*>(tmp.p),
                              a.elems(), s));
    XV_CUDA(cudaMemcpyAsync(out_host, tmp.p, a.elems() * 4, cudaMemcpyDeviceToHost, s));
    XV_CUDA(cudaStreamSynchronize(s));
  }
  XV_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// ------------------------------------------------------------------ single layers
int xv_conv2d(const float* x, const float* w_host, const float* bias_host, int n, int h, int w,
              int cin, int cout, int k, int relu, int precision, float* out, void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(x && w_host && out, "xv_conv2d: NULL argument");
  XV_CHECK(k == 1 || k == 3, "xv_conv2d: k must be 1 or 3");
  cudaStream_t s = XV_STREAM(stream);
  xv_fcn fake;
  fake.precision = precision;
  ConvLayer L;
  L.name = "conv2d";
  L.k = k;
  L.cin = cin;
  L.cout = cout;
  L.relu = relu;
  std::vector<float> scale(cout, 1.f), shift(cout, 0.f);
  XV_TRY(pack_conv(&fake, &L, w_host, bias_host, scale, shift, false));
  const size_t npix = static_cast<size_t>(n) * h * w;
  if (precision == XV_PRECISION_FP32) {
    XV_TRY(launch_conv_f32(x, static_cast<const float*>(L.w_f32.p),
                           static_cast<const float*>(L.bias_f32.p), out, n, h, w, cin, cout, k,
                           relu, s));
    XV_CUDA(cudaStreamSynchronize(s));
    return 0;
  }
  DevBuf in_bf16;
  XV_TRY(in_bf16.ensure(npix * L.cin_gemm * 2));
  if (L.taps == 1 && k == 3 && cout == 64 && !(g_debug_flags & 4)) {
    DevBuf out_bf16;
    XV_TRY(out_bf16.ensure(npix * cout * 2));
    XV_TRY(run_igemm_c1(nullptr, L, x, n, h, w, out_bf16.p, s));
    XV_TRY(launch_bf16_to_f32(static_cast<const __nv_bfloat16*>(out_bf16.p), out, npix * cout, s));
    XV_CUDA(cudaStreamSynchronize(s));
    return 0;
  }
  if (L.taps == 1 && k == 3) {
    XV_TRY(launch_im2col_c1(x, static_cast<__nv_bfloat16*>(in_bf16.p), n, h, w, cin, s));
  } else {
    XV_TRY(launch_f32_to_bf16(x, static_cast<__nv_bfloat16*>(in_bf16.p), npix * cin, s));
  }
  if (cout % 64 == 0) {
    DevBuf out_bf16;
    XV_TRY(out_bf16.ensure(npix * cout * 2));
    XV_TRY(run_igemm(nullptr, L, in_bf16.p, n, h, w, out_bf16.p, false, s));
    XV_TRY(launch_bf16_to_f32(static_cast<const __nv_bfloat16*>(out_bf16.p), out, npix * cout, s));
    XV_CUDA(cudaStreamSynchronize(s));
  } else {
    XV_TRY(run_igemm(nullptr, L, in_bf16.p, n, h, w, out, true, s));
    XV_CUDA(cudaStreamSynchronize(s));
  }
  return 0;
}

int xv_deconv2d(const float* x, const float* w_host, int n, int h, int w, int cin, int cout, int k,
                int stride, int relu, float* out, void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(x && w_host && out, "xv_deconv2d: NULL argument");
  cudaStream_t s = XV_STREAM(stream);
  DevBuf wd;
  std::vector<float> wv(w_host, w_host + static_cast<size_t>(k) * k * cout * cin);
  XV_TRY(wd.upload(wv));
  XV_TRY(launch_deconv_f32(x, static_cast<const float*>(wd.p), out, n, h, w, cin, cout, k, stride,
                           relu, nullptr, s));
  XV_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int xv_convert_to_f32(const void* src, int src_dtype, int64_t n, float* dst, void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(src && dst && n >= 0, "xv_convert_to_f32: bad argument");
  return launch_to_f32(src, src_dtype, dst, static_cast<size_t>(n), XV_STREAM(stream));
}

int xv_maxpool2x2(const float* x, int n, int h, int w, int c, float* out, void* stream) {
  XV_TRY(ensure_init());
  return launch_maxpool_f32(x, out, n, h, w, c, XV_STREAM(stream));
}

// Timing harness for one tensor-core conv layer on synthetic bf16 data (not a reference
// entry point; used by tools/conv_bench.py to study the kernel in isolation).
int xv_batchnorm_train(const float* x, const float* gamma, const float* beta, int n, int h, int w,
                       int c, int relu, float* y, float* mean_out, float* var_out,
                       float* moving_mean, float* moving_var, void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(x && gamma && beta && y && n > 0 && h > 0 && w > 0 && c > 0, "xv_batchnorm_train: bad arguments");
  XV_CHECK((moving_mean == nullptr) == (moving_var == nullptr),
           "xv_batchnorm_train: pass both moving statistics or none");
  cudaStream_t s = XV_STREAM(stream);
  const size_t npix = static_cast<size_t>(n) * h * w;
  DevBuf scratch;
  XV_TRY(scratch.ensure(2 * c * sizeof(double) + 2 * c * sizeof(float)));
  double* sums = static_cast<double*>(scratch.p);
  float* mean = reinterpret_cast<float*>(sums + 2 * c);
  float* rstd = mean + c;
  XV_TRY(launch_bn_stats(x, false, npix, c, sums, s));
  XV_TRY(launch_bn_finalize(sums, npix, c, 1e-3f, 0.99f, nullptr, mean, rstd, moving_mean,
                            moving_var, s));
  XV_TRY(launch_bn_apply(x, false, mean, rstd, gamma, beta, npix, c, relu, y, s));
  if (mean_out) XV_CUDA(cudaMemcpyAsync(mean_out, mean, c * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (var_out) {
    // var = 1 / rstd^2 - eps, recomputed on the host side of the scratch to keep one code path
    std::vector<float> r(c);
    XV_CUDA(cudaMemcpyAsync(r.data(), rstd, c * sizeof(float), cudaMemcpyDeviceToHost, s));
    XV_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < c; ++i) r[i] = 1.f / (r[i] * r[i]) - 1e-3f;
    XV_CUDA(cudaMemcpyAsync(var_out, r.data(), c * sizeof(float), cudaMemcpyHostToDevice, s));
  }
  XV_CUDA(cudaStreamSynchronize(s));      // the scratch buffer dies with this call
  return 0;
}

int xv_batchnorm_train_backward(const float* x, const float* y, const float* dy,
                                const float* gamma, int n, int h, int w, int c, int relu,
                                float* dx, float* dgamma, float* dbeta, void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(x && dy && gamma && dx && dgamma && dbeta && (y || !relu),
           "xv_batchnorm_train_backward: bad arguments");
  cudaStream_t s = XV_STREAM(stream);
  const size_t npix = static_cast<size_t>(n) * h * w;
  DevBuf scratch;
  XV_TRY(scratch.ensure(2 * c * sizeof(double) + 2 * c * sizeof(float)));
  double* sums = static_cast<double*>(scratch.p);
  float* mean = reinterpret_cast<float*>(sums + 2 * c);
  float* rstd = mean + c;
  XV_TRY(launch_bn_stats(x, false, npix, c, sums, s));
  XV_TRY(launch_bn_finalize(sums, npix, c, 1e-3f, 0.99f, nullptr, mean, rstd, nullptr, nullptr, s));
  XV_TRY(launch_bn_backward(dy, relu ? y : nullptr, x, false, mean, rstd, gamma, npix, c, sums, dx,
                            nullptr, 0, dgamma, dbeta, s));
  XV_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int xv_bench_conv_igemm(int n, int h, int w, int cin, int cout, int k, int iters, int flags,
                        float* ms_out) {
  XV_TRY(ensure_init());
  xv_fcn fake;
  fake.precision = XV_PRECISION_BF16;
  ConvLayer L;
  L.name = "bench";
  L.k = k;
  L.cin = cin;
  L.cout = cout;
  L.relu = 1;
  std::vector<float> wv(static_cast<size_t>(k) * k * cin * cout), bv(cout, 0.1f);
  for (size_t i = 0; i < wv.size(); ++i) wv[i] = 0.01f * static_cast<float>((i * 2654435761u) % 17) - 0.08f;
  std::vector<float> scale(cout, 1.f), shift(cout, 0.f);
  XV_TRY(pack_conv(&fake, &L, wv.data(), bv.data(), scale, shift, false));
  const size_t npix = static_cast<size_t>(n) * h * w;
  DevBuf in, out;
  XV_TRY(in.ensure(npix * cin * 2));
  XV_TRY(out.ensure(npix * cout * 2));
  XV_CUDA(cudaMemset(in.p, 0x3c, npix * cin * 2));
  ConvIgemmParams p;
  std::memset(&p, 0, sizeof(p));
  choose_tile(h, w, &p.th, &p.tw);
  // flag 8192: halo variant of the pixel-major kernel
  p.halo = ((flags & 8192) && k == 3 && cin % 64 == 0 && (p.tw == 8 || p.tw == 16)) ? 1 : 0;
  if (cin % 64 == 0)
    XV_TRY(make_tmap_act(&p.tmap_in, in.p, n, h, w, cin, cin, 1, p.halo ? p.th + 2 : p.th, p.tw));
  XV_TRY(make_tmap_w(&p.tmap_w, L.w_packed.p, L.kdim, L.cout_pad, L.block_n));
  XV_TRY(make_tmap_act(&p.tmap_out, out.p, n, h, w, cout, cout, 1, p.th, p.tw));
  p.bias = static_cast<const float*>(L.bias_pad.p);
  p.N = n; p.H = h; p.W = w; p.cin = cin; p.cout = cout;
  p.tiles_x = div_up(w, p.tw); p.tiles_y = div_up(h, p.th);
  p.n_blocks = L.cout_pad / L.block_n; p.relu = 1; p.debug_flags = flags & ~8192;
  cudaEvent_t e0, e1;
  XV_CUDA(cudaEventCreate(&e0));
  XV_CUDA(cudaEventCreate(&e1));
  if (k == 3 && cin <= 3) {   // conv1_1 path: raw fp32 input, operand packing inside the kernel
    DevBuf xraw;
    XV_TRY(xraw.ensure(npix * cin * 4));
    XV_CUDA(cudaMemset(xraw.p, 0x3c, npix * cin * 4));
    p.x_raw = static_cast<const float*>(xraw.p);
    p.tmap_in = p.tmap_out;
    p.cin = 64;
    p.n_blocks = 1;
    // flag 4096: the previous variant (global loads in the packers); else conv_c1_sm100.cu
    const bool staged = !(flags & 4096) && conv_c1_patch_fits(p.th, p.tw, cin, &p.patch_w);
    if (staged) XV_TRY(make_tmap_patch(&p.tmap_in, xraw.p, n, h, w, cin, p.patch_w, p.th + 2));
    auto run = [&]() -> int {
      return staged ? launch_conv_c1(p, cin, 0) : launch_conv_igemm_c1(p, cin, 0);
    };
    XV_TRY(run());
    XV_CUDA(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i) XV_TRY(run());
    XV_CUDA(cudaEventRecord(e1, 0));
    XV_CUDA(cudaEventSynchronize(e1));
    float ms1 = 0.f;
    XV_CUDA(cudaEventElapsedTime(&ms1, e0, e1));
    *ms_out = ms1 / iters;
    return 0;
  }
  const bool use_t = (flags & 512) != 0 && L.use_t;
  const bool t_pool = (flags & 1024) != 0;
  // flag 16384 (with 8192): CTA-pair kernel
  const bool pair = (flags & 16384) && p.halo;
  if (pair) {
    XV_TRY(make_tmap_w(&p.tmap_w, L.w_packed.p, L.kdim, L.cout_pad, L.block_n / 2));
    p.n_blocks = div_up(cout, L.block_n);
  }
  p.debug_flags = flags & ~(8192 | 16384);
  auto once = [&]() -> int {
    if (pair) return launch_conv_igemm_2cta(p, L.block_n, 0);
    if (use_t) return run_igemm_t(nullptr, L, in.p, n, h, w, out.p, t_pool, 0);
    return launch_conv_igemm(p, L.block_n, L.taps, false, 0);
  };
  XV_TRY(once());
  XV_CUDA(cudaEventRecord(e0, 0));
  for (int i = 0; i < iters; ++i) XV_TRY(once());
  XV_CUDA(cudaEventRecord(e1, 0));
  XV_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  XV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = ms / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

// ------------------------------------------------------------------ fusion stage
static int check_label_bytes(int b) {
  XV_CHECK(b == 8 || b == 1, "label_bytes must be 8 (int64) or 1 (uint8)");
  return 0;
}

int xv_softmax_argmax(const float* score, int64_t npix, int C, float* prob, void* label,
                      int label_bytes, void* stream) {
  XV_TRY(ensure_init());
  XV_TRY(check_label_bytes(label_bytes));
  return launch_softmax_argmax(score, npix, C, prob,
                               label_bytes == 8 ? static_cast<int64_t*>(label) : nullptr,
                               label_bytes == 1 ? static_cast<uint8_t*>(label) : nullptr,
                               XV_STREAM(stream));
}

int xv_bayes_fuse_lut(const void* const* labels, int M, int label_bytes, const int32_t* lut, int C,
                      int64_t npix, void* out, void* stream) {
  XV_TRY(ensure_init());
  return launch_bayes_lut(labels, M, label_bytes, lut, C, npix, out, XV_STREAM(stream));
}

int xv_bayes_fuse_score(const void* const* labels, int M, int label_bytes, const float* log_cond,
                        const float* log_prior, int C, int64_t npix, float* score, void* label,
                        void* stream) {
  XV_TRY(ensure_init());
  XV_TRY(check_label_bytes(label_bytes));
  return launch_bayes_score(labels, M, label_bytes, log_cond, log_prior, C, npix, score, label,
                            XV_STREAM(stream));
}

int xv_bayes_decode_score(xv_fcn* const* experts, int M, const int32_t* lut, int C,
                          const int32_t* gt_labels, int64_t* cm, uint8_t* fused_out,
                          void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(experts && lut && gt_labels && cm && M >= 1 && M <= 4,
           "xv_bayes_decode_score: bad arguments");
  const float *low[4], *g[4], *bias[4];
  int N = 0, h = 0, w = 0;
  for (int m = 0; m < M; ++m) {
    xv_fcn* e = experts[m];
    XV_CHECK(e && e->finalized && e->fast_up && e->C == C,
             "xv_bayes_decode_score: experts need the bilinear decoder fast path and C classes");
    auto it = e->layers.find("score_lowres");
    XV_CHECK(it != e->layers.end() && it->second.p != nullptr,
             "xv_bayes_decode_score: run xv_fcn_forward on every expert first");
    const Act& a = it->second;
    if (m == 0) {
      N = a.B;
      h = a.H;
      w = a.W;
    }
    XV_CHECK(a.B == N && a.H == h && a.W == w, "xv_bayes_decode_score: expert shapes differ");
    low[m] = static_cast<const float*>(a.p);
    g[m] = static_cast<const float*>(e->g16.p);
    bias[m] = static_cast<const float*>(e->b_score.p);
  }
  return launch_decode_bayes_confusion(low, g, bias, M, lut, C, N, h, w, gt_labels,
                                       reinterpret_cast<long long*>(cm), fused_out,
                                       XV_STREAM(stream));
}

int xv_dirichlet_decode_score(xv_fcn* const* experts, int M, const float* alpha_m1,
                              const float* log_norm, const float* log_prior, int C,
                              float abs_alpha_m1_max, float abs_norm_max, const int32_t* gt_labels,
                              int64_t* cm, void* label, int label_bytes, int64_t* num_exact,
                              void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(experts && alpha_m1 && log_norm && log_prior && M == 2,
           "xv_dirichlet_decode_score: two experts and their tables are needed");
  if (label) XV_TRY(check_label_bytes(label_bytes));
  XV_CHECK(label != nullptr || cm != nullptr, "xv_dirichlet_decode_score: nothing to compute");
  const float *low[2], *g[2], *bias[2];
  int N = 0, h = 0, w = 0;
  for (int m = 0; m < M; ++m) {
    xv_fcn* e = experts[m];
    XV_CHECK(e && e->finalized && e->fast_up && e->C == C,
             "xv_dirichlet_decode_score: experts need the bilinear decoder fast path and C classes");
    auto it = e->layers.find("score_lowres");
    XV_CHECK(it != e->layers.end() && it->second.p != nullptr,
             "xv_dirichlet_decode_score: run xv_fcn_forward on every expert first");
    const Act& a = it->second;
    if (m == 0) {
      N = a.B;
      h = a.H;
      w = a.W;
    }
    XV_CHECK(a.B == N && a.H == h && a.W == w, "xv_dirichlet_decode_score: expert shapes differ");
    low[m] = static_cast<const float*>(a.p);
    g[m] = static_cast<const float*>(e->g16.p);
    bias[m] = static_cast<const float*>(e->b_score.p);
  }
  return launch_decode_dirichlet(low, g, bias, M, alpha_m1, log_norm, log_prior, abs_alpha_m1_max,
                                 abs_norm_max, C, N, h, w, gt_labels,
                                 reinterpret_cast<long long*>(cm), label, label_bytes,
                                 reinterpret_cast<unsigned long long*>(num_exact),
                                 XV_STREAM(stream));
}

int xv_dirichlet_fuse(const float* const* probs, int M, const float* alpha_m1,
                      const float* log_norm, const float* log_prior, int C, int64_t npix,
                      float* score, void* label, int label_bytes, void* stream) {
  XV_TRY(ensure_init());
  XV_TRY(check_label_bytes(label_bytes));
  return launch_dirichlet_fuse(probs, M, alpha_m1, log_norm, log_prior, C, npix, score, label,
                               label_bytes, -1.f, 0.f, nullptr, XV_STREAM(stream));
}

int xv_dirichlet_fuse_exact(const float* const* probs, int M, const float* alpha_m1,
                            const float* log_norm, const float* log_prior, int C, int64_t npix,
                            float abs_alpha_m1_max, float abs_norm_max, float* score, void* label,
                            int label_bytes, int64_t* num_exact, void* stream) {
  XV_TRY(ensure_init());
  XV_TRY(check_label_bytes(label_bytes));
  XV_CHECK(abs_alpha_m1_max >= 0.f && abs_norm_max >= 0.f,
           "xv_dirichlet_fuse_exact: the table magnitudes must be non-negative");
  return launch_dirichlet_fuse(probs, M, alpha_m1, log_norm, log_prior, C, npix, score, label,
                               label_bytes, abs_alpha_m1_max, abs_norm_max,
                               reinterpret_cast<unsigned long long*>(num_exact),
                               XV_STREAM(stream));
}

int xv_average_fuse(const float* const* probs, int M, int C, int64_t npix, float* score,
                    void* label, int label_bytes, void* stream) {
  XV_TRY(ensure_init());
  XV_TRY(check_label_bytes(label_bytes));
  return launch_average_fuse(probs, M, C, npix, score, label, label_bytes, XV_STREAM(stream));
}

int xv_variance_fuse(const float* const* probs, const float* const* vars, int M, int C,
                     int64_t npix, float* score, void* label, int label_bytes, void* stream) {
  XV_TRY(ensure_init());
  XV_TRY(check_label_bytes(label_bytes));
  return launch_variance_fuse(probs, vars, M, C, npix, score, label, label_bytes,
                              XV_STREAM(stream));
}

int xv_mc_moments(const float* samples, int T, int64_t npix, int C, float* mean, float* var,
                  float* mean_var, float* entropy, float* cond_entropy, float* sum_var,
                  void* stream) {
  XV_TRY(ensure_init());
  return launch_mc_moments(samples, T, npix, C, mean, var, mean_var, entropy, cond_entropy,
                           sum_var, XV_STREAM(stream));
}

int xv_dirichlet_fit_samples(const float* samples, int T, int64_t npix, int C, float tol,
                             int maxiter, float* alpha, int32_t* iterations, void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(samples && alpha, "xv_dirichlet_fit_samples: NULL argument");
  return launch_dirichlet_fit_samples(samples, T, npix, C, tol, maxiter, alpha, iterations,
                                      XV_STREAM(stream));
}

int xv_dirichlet_uncertainty_fuse(const float* const* probs, const float* const* vars, int M,
                                  const float* cond_params, const float* log_prior, int C,
                                  int64_t npix, float* score, void* label, int label_bytes,
                                  void* stream) {
  XV_TRY(ensure_init());
  XV_TRY(check_label_bytes(label_bytes));
  XV_CHECK(M >= 1 && M <= 4, "number of experts must be in [1, 4]");
  static DevBuf maxbuf;
  XV_TRY(maxbuf.ensure(4 * sizeof(float)));
  const float* maxp[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int m = 0; m < M; ++m) {
    float* slot = static_cast<float*>(maxbuf.p) + m;
    XV_TRY(launch_reduce_max(vars[m], npix * C, slot, XV_STREAM(stream)));
    maxp[m] = slot;
  }
  return launch_dirichlet_uncertainty_fuse(probs, vars, maxp, M, cond_params, log_prior, C, npix,
                                           score, label, label_bytes, XV_STREAM(stream));
}

int xv_dirichlet_suffstats(const float* prob, const int32_t* labels, int64_t npix, int C,
                           double* stats, int64_t* counts, void* stream) {
  XV_TRY(ensure_init());
  return launch_suffstats(prob, labels, npix, C, stats, reinterpret_cast<long long*>(counts),
                          XV_STREAM(stream));
}

int xv_confusion_accumulate(const void* pred, int pred_bytes, const int32_t* labels, int64_t npix,
                            int C, int64_t* cm, void* stream) {
  XV_TRY(ensure_init());
  return launch_confusion(pred, pred_bytes, labels, npix, C, reinterpret_cast<long long*>(cm),
                          XV_STREAM(stream));
}

}  // extern "C"

// ------------------------------------------------------------------ training (fit)
namespace xv {

struct TrainLayer {
  std::string name;
  int k, cin, cout;
  size_t w_off, b_off;               // offsets into the flat parameter / gradient buffers
  std::unique_ptr<ConvLayer> bwd;    // data-gradient conv (flipped + transposed weights)
  // batch-norm training: the forward conv with the RAW weights, no ReLU (the inference layer of
  // the same name carries the folded moving statistics)
  std::unique_ptr<ConvLayer> fwd;
};

// batch-norm variables of one scope inside the flat vector (gamma / beta are trained; the moving
// statistics ride along with zero gradients and are updated by the forward pass)
struct BnParam {
  std::string name;
  int C;
  size_t g_off, b_off, mm_off, mv_off;
};

struct TrainState {
  std::vector<TrainLayer> layers;    // conv1_1..conv5_3, score_conv4, score_conv5, score
  size_t total = 0;
  DevBuf master, m, v;               // fp32 flat
  DevBuf loss;                       // double[2]: sum of -log p, #valid pixels
  int64_t step = 0;
  bool bn = false;                   // batch_normalization=True: batch statistics in training
  std::vector<BnParam> bns;
  int opt_kind = XV_OPT_ADAM;        // whose slot values m / v currently hold
  // gradient buckets for the overlapped all-reduce: suffixes of the flat vector in the order
  // the backward pass completes them (conv5_x + heads, conv4_x, conv3_x, conv1_x + conv2_x)
  std::vector<size_t> bucket_off;    // ascending offsets, bucket_off.back() == total
  std::vector<cudaEvent_t> bucket_events;   // set per call, recorded when a bucket is complete
};

static std::map<xv_fcn*, std::unique_ptr<TrainState>> g_train;

static TrainLayer* find_layer(TrainState* ts, const std::string& n) {
  for (auto& l : ts->layers)
    if (l.name == n) return &l;
  return nullptr;
}
static BnParam* find_bn(TrainState* ts, const std::string& n) {
  for (auto& b : ts->bns)
    if (b.name == n) return &b;
  return nullptr;
}

// master fp32 -> bf16 operand copies (forward + data-gradient) and padded biases of one layer
static int repack_layer(xv_fcn* net, TrainState* ts, TrainLayer& tl, cudaStream_t s) {
  // batch-norm training: the raw weights go to the layer's own forward copy; the final score conv
  // reads its fp32 master values directly
  if (ts->bn && tl.name == "score") return 0;
  ConvLayer* L = ts->bn ? tl.fwd.get() : net->conv(tl.name);
  const float* w = static_cast<const float*>(ts->master.p) + tl.w_off;
  const float* b = static_cast<const float*>(ts->master.p) + tl.b_off;
  if (tl.name == "score") {
    XV_CUDA(cudaMemcpyAsync(L->w_f32.p, w, sizeof(float) * tl.cin * tl.cout,
                            cudaMemcpyDeviceToDevice, s));
    XV_CUDA(cudaMemcpyAsync(L->bias_f32.p, b, sizeof(float) * tl.cout, cudaMemcpyDeviceToDevice, s));
    XV_CUDA(cudaMemcpyAsync(net->w_score_nuxc.p, w, sizeof(float) * tl.cin * tl.cout,
                            cudaMemcpyDeviceToDevice, s));
    XV_CUDA(cudaMemcpyAsync(net->b_score.p, b, sizeof(float) * tl.cout, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const bool c1 = (L->taps == 1 && L->k == 3);
  __nv_bfloat16* bwd = tl.bwd ? static_cast<__nv_bfloat16*>(tl.bwd->w_packed.p) : nullptr;
  const int bwd_kdim = tl.bwd ? tl.bwd->kdim : 0;
  XV_TRY(launch_pack_weights(w, static_cast<__nv_bfloat16*>(L->w_packed.p), bwd, tl.k * tl.k, tl.cin,
                             tl.cout, L->kdim, bwd_kdim, c1 ? 1 : 0, s));
  XV_CUDA(cudaMemcpyAsync(L->bias_pad.p, b, sizeof(float) * tl.cout, cudaMemcpyDeviceToDevice, s));
  return 0;
}

// all layers after an optimizer step: one launch for every conv kernel + bias (pack_all_kernel),
// the final score layer's fp32 copies separately
static int repack_all(xv_fcn* net, TrainState* ts, cudaStream_t s) {
  PackAllParams p;
  std::memset(&p, 0, sizeof(p));
  for (auto& tl : ts->layers) {
    if (tl.name == "score" || p.num_layers == 20) {
      XV_TRY(repack_layer(net, ts, tl, s));
      continue;
    }
    ConvLayer* L = ts->bn ? tl.fwd.get() : net->conv(tl.name);
    PackLayer& d = p.layer[p.num_layers++];
    d.w = static_cast<const float*>(ts->master.p) + tl.w_off;
    d.b = static_cast<const float*>(ts->master.p) + tl.b_off;
    d.fwd = static_cast<__nv_bfloat16*>(L->w_packed.p);
    d.bwd = tl.bwd ? static_cast<__nv_bfloat16*>(tl.bwd->w_packed.p) : nullptr;
    d.bias_pad = static_cast<float*>(L->bias_pad.p);
    d.taps = tl.k * tl.k;
    d.cin = tl.cin;
    d.cout = tl.cout;
    d.fwd_kdim = L->kdim;
    d.bwd_kdim = tl.bwd ? tl.bwd->kdim : 0;
    d.c1_layout = (L->taps == 1 && L->k == 3) ? 1 : 0;
    d.first = p.total;
    d.elems = static_cast<unsigned long long>(d.taps) * d.cin * d.cout;
    p.total += d.elems;
  }
  if (p.num_layers > 0) XV_TRY(launch_pack_all(p, s));
  return 0;
}

// tensor-core weight gradient; the split of the pixel tiles over CTAs is chosen to fill whole waves
// dy: bf16 [B,H,W,dy_channels] (dy_channels = 0: == cout; the 1x1 heads pass their gradient
// zero-padded to 64 channels); taps = 9 (3x3) or 1 (1x1)
static int run_wgrad_tc(xv_fcn* net, const void* x, const void* dy, float* dw, int B, int H, int W,
                        int cin, int cout, cudaStream_t s, int taps = 9, int dy_channels = 0) {
  ConvWgradParams p;
  std::memset(&p, 0, sizeof(p));
  choose_tile(H, W, &p.th, &p.tw);
  XV_TRY(get_tmap(net, &p.tmap_x, x, B, H, W, cin, p.th, p.tw));
  XV_TRY(get_tmap(net, &p.tmap_dy, dy, B, H, W, dy_channels ? dy_channels : cout, p.th, p.tw));
  p.taps = taps;
  p.dw = dw;
  p.N = B;
  p.H = H;
  p.W = W;
  p.cin = cin;
  p.cout = cout;
  p.tiles_x = div_up(W, p.tw);
  p.tiles_y = div_up(H, p.th);
  // CTA pairs (M = 256 per MMA, the x operand split over the pair) where the shapes allow it;
  // debug bit11 keeps the single-CTA kernel
  const bool pair = taps == 9 && cin % 128 == 0 && cout % 256 == 0 && dy_channels == 0 &&
                    !(g_debug_flags & 2048);
  p.m_blocks = pair ? cout / 256 : div_up(cout, 128);
  p.total_atoms = taps * cin / 64;
  p.n_groups = div_up(p.total_atoms, 4);
  const int base = p.m_blocks * p.n_groups;
  const int tiles = B * p.tiles_x * p.tiles_y;
  const int sms = pair ? g_dev.num_sms / 2 : g_dev.num_sms;      // schedulable CTAs / clusters
  int best_k = 1;
  double best_cost = 1e30;
  for (int k = 1; k <= 64 && k <= tiles; ++k) {
    const double waves = std::ceil(static_cast<double>(base) * k / sms);
    const double cost = waves * (std::ceil(static_cast<double>(tiles) / k) + 6.0);  // + epilogue
    if (cost < best_cost) {
      best_cost = cost;
      best_k = k;
    }
  }
  p.k_splits = best_k;
  return pair ? launch_conv_wgrad_2cta(p, s) : launch_conv_wgrad_tc(p, s);
}

struct Backward {
  xv_fcn* net;
  TrainState* ts;
  Arena* arena;
  cudaStream_t s;
  bool dry;
  float* grads;

  Act make(DType dt, int B, int H, int W, int C) {
    Act a;
    a.dt = dt;
    a.B = B;
    a.H = H;
    a.W = W;
    a.C = C;
    a.p = arena->alloc(a.elems() * (dt == DType::F32 ? 4 : 2));
    return a;
  }
  Act layer(const std::string& n) { return net->layers.at(n); }
  // dY (already masked by the layer's ReLU) -> weight/bias gradients, optional data gradient
  int conv_bwd(const std::string& name, const Act& x, const Act& dy, bool need_dx, Act* dx) {
    TrainLayer* tl = find_layer(ts, name);
    XV_CHECK(tl != nullptr, "no train layer " + name);
    if (need_dx) *dx = make(DType::BF16, dy.B, dy.H, dy.W, tl->cin);
    if (dry) return 0;
    // (the bias gradient was accumulated by the kernel that produced dy)
    if (tl->cin % 64 == 0 && !(g_debug_flags & 8)) {
      XV_TRY(run_wgrad_tc(net, x.p, dy.p, grads + tl->w_off, dy.B, dy.H, dy.W, tl->cin, tl->cout, s));
    } else {
      XV_TRY(launch_conv_wgrad(static_cast<const __nv_bfloat16*>(x.p),
                               static_cast<const __nv_bfloat16*>(dy.p), grads + tl->w_off, dy.B,
                               dy.H, dy.W, tl->cin, tl->cout, s));
    }
    if (need_dx) XV_TRY(run_igemm(net, *tl->bwd, dy.p, dy.B, dy.H, dy.W, dx->p, false, s));
    return 0;
  }
  // gradient wrt the pre-ReLU output of conv `name` (+ its bias gradient) from the gradient wrt
  // its post-ReLU output `da`
  int relu_bwd(const std::string& name, const Act& da, const Act& y, Act* out) {
    TrainLayer* tl = find_layer(ts, name);
    *out = make(DType::BF16, y.B, y.H, y.W, y.C);
    if (dry) return 0;
    return launch_relu_bwd_bf16(static_cast<const __nv_bfloat16*>(da.p), nullptr,
                                static_cast<const __nv_bfloat16*>(y.p),
                                static_cast<__nv_bfloat16*>(out->p), y.elems(), y.C,
                                grads + tl->b_off, s);
  }
  // same through the 2x2 max pool that follows conv `name` (dp = gradient wrt the pooled map,
  // extra = optional second gradient wrt the un-pooled map)
  int pool_relu_bwd(const std::string& name, const Act& dp, const Act& y, const Act& p,
                    const Act* extra, Act* out) {
    TrainLayer* tl = find_layer(ts, name);
    *out = make(DType::BF16, y.B, y.H, y.W, y.C);
    if (dry) return 0;
    return launch_pool_relu_bwd_bf16(static_cast<const __nv_bfloat16*>(dp.p),
                                     static_cast<const __nv_bfloat16*>(y.p),
                                     static_cast<const __nv_bfloat16*>(p.p),
                                     extra ? static_cast<const __nv_bfloat16*>(extra->p) : nullptr,
                                     static_cast<__nv_bfloat16*>(out->p), y.B, y.H, y.W, y.C,
                                     grads + tl->b_off, s);
  }
  // 1x1 head conv with fp32 ReLU output y: returns the un-masked gradient wrt its bf16 input
  int head_bwd(const std::string& name, const Act& x, const Act& y, const Act& dy, Act* dx) {
    TrainLayer* tl = find_layer(ts, name);
    const int nu = tl->cout, nu_pad = tl->bwd->cin_gemm;
    Act dpre = make(DType::F32, y.B, y.H, y.W, nu);
    Act dpre16 = make(DType::BF16, y.B, y.H, y.W, nu_pad);
    *dx = make(DType::BF16, y.B, y.H, y.W, tl->cin);
    if (dry) return 0;
    const size_t npix = static_cast<size_t>(y.B) * y.H * y.W;
    XV_TRY(launch_relu_mask_f32(static_cast<const float*>(dy.p), static_cast<const float*>(y.p),
                                static_cast<float*>(dpre.p),
                                static_cast<__nv_bfloat16*>(dpre16.p), npix, nu, nu_pad, s));
    XV_TRY(launch_bias_grad_f32(static_cast<const float*>(dpre.p), grads + tl->b_off, npix, nu, s));
    // weight gradient [512, nu] = x^T . dpre on the tensor cores (K = pixels), from the bf16 copy
    if (tl->cin % 64 == 0 && !(g_debug_flags & 8))
      XV_TRY(run_wgrad_tc(net, x.p, dpre16.p, grads + tl->w_off, y.B, y.H, y.W, tl->cin, nu, s, 1,
                          nu_pad));
    else
      XV_TRY(launch_outer_sum_bf16(static_cast<const __nv_bfloat16*>(x.p),
                                   static_cast<const float*>(dpre.p), grads + tl->w_off, npix,
                                   tl->cin, nu, s));
    return run_igemm(net, *tl->bwd, dpre16.p, y.B, y.H, y.W, dx->p, false, s);
  }

  // bucket b of the flat gradient is final: record the caller's event for it (if any)
  int bucket_done(int b) {
    if (dry || b >= static_cast<int>(ts->bucket_events.size()) || !ts->bucket_events[b]) return 0;
    XV_CUDA(cudaEventRecord(ts->bucket_events[b], s));
    return 0;
  }

  int run(const float* x, const int32_t* labels, int N, int H, int W, int train_encoder);
};

int Backward::run(const float* x, const int32_t* labels, int N, int H, int W, int train_encoder) {
  const int nu = net->nu, C = net->C;
  const int h8 = H / 8, w8 = W / 8, h16 = H / 16, w16 = W / 16;
  TrainLayer* tscore = find_layer(ts, "score");
  const bool fused_loss = !(g_debug_flags & 4096);
  Act low = layer("score_lowres"), fused = layer("fused");
  Act score = fused_loss ? low : layer("train_score");
  Act up5 = layer("upscore_conv5"), s4 = layer("score_conv4"), s5 = layer("score_conv5");
  const size_t npix = static_cast<size_t>(N) * H * W;
  const size_t nlow = static_cast<size_t>(N) * h8 * w8;
  // loss + d(score); transposes of the two fixed bilinear layers; score 1x1
  Act dlow = make(DType::F32, N, h8, w8, C);
  Act dfused = make(DType::F32, N, h8, w8, nu);
  Act ds5 = make(DType::F32, N, h16, w16, nu);
  if (!dry) {
    if (fused_loss) {
      XV_TRY(launch_loss_lowres_grad(static_cast<const float*>(low.p),
                                     static_cast<const float*>(net->g16.p),
                                     static_cast<const float*>(net->b_score.p), labels, N, h8, w8,
                                     C, static_cast<float*>(dlow.p),
                                     static_cast<double*>(ts->loss.p), grads + tscore->b_off, s));
    } else {
      XV_TRY(launch_ce_grad(static_cast<float*>(score.p), labels, static_cast<int64_t>(npix), C,
                            static_cast<double*>(ts->loss.p), grads + tscore->b_off, s));
      XV_TRY(launch_upsample8_transpose(static_cast<const float*>(score.p),
                                        static_cast<const float*>(net->g16.p),
                                        static_cast<float*>(dlow.p), N, h8, w8, C, s));
    }
    XV_TRY(launch_score_bwd(static_cast<const float*>(dlow.p), static_cast<const float*>(fused.p),
                            static_cast<const float*>(net->w_score_nuxc.p),
                            static_cast<float*>(dfused.p), grads + tscore->w_off, nlow, nu, C, s));
    XV_TRY(launch_upscore2_bwd(static_cast<const float*>(dfused.p),
                               static_cast<const float*>(up5.p),
                               static_cast<const float*>(net->g4.p), static_cast<float*>(ds5.p), N,
                               h16, w16, nu, s));
  }
  (void)low;
  // heads: d(fused) is also d(score_conv4 output)
  Act c43 = layer("conv4_3"), c53 = layer("conv5_3");
  Act d43a, d53;
  XV_TRY(head_bwd("score_conv4", c43, s4, dfused, &d43a));
  XV_TRY(head_bwd("score_conv5", c53, s5, ds5, &d53));
  if (!train_encoder) {
    for (int b = 0; b < 4; ++b) XV_TRY(bucket_done(b));
    return 0;
  }
  // encoder, last layer first
  Act g, dx, dp;
  XV_TRY(relu_bwd("conv5_3", d53, c53, &g));
  XV_TRY(conv_bwd("conv5_3", layer("conv5_2"), g, true, &dx));
  XV_TRY(relu_bwd("conv5_2", dx, layer("conv5_2"), &g));
  XV_TRY(conv_bwd("conv5_2", layer("conv5_1"), g, true, &dx));
  XV_TRY(relu_bwd("conv5_1", dx, layer("conv5_1"), &g));
  XV_TRY(conv_bwd("conv5_1", layer("pool4"), g, true, &dp));
  XV_TRY(bucket_done(0));
  XV_TRY(pool_relu_bwd("conv4_3", dp, c43, layer("pool4"), &d43a, &g));
  XV_TRY(conv_bwd("conv4_3", layer("conv4_2"), g, true, &dx));
  XV_TRY(relu_bwd("conv4_2", dx, layer("conv4_2"), &g));
  XV_TRY(conv_bwd("conv4_2", layer("conv4_1"), g, true, &dx));
  XV_TRY(relu_bwd("conv4_1", dx, layer("conv4_1"), &g));
  XV_TRY(conv_bwd("conv4_1", layer("pool3"), g, true, &dp));
  XV_TRY(bucket_done(1));
  XV_TRY(pool_relu_bwd("conv3_3", dp, layer("conv3_3"), layer("pool3"), nullptr, &g));
  XV_TRY(conv_bwd("conv3_3", layer("conv3_2"), g, true, &dx));
  XV_TRY(relu_bwd("conv3_2", dx, layer("conv3_2"), &g));
  XV_TRY(conv_bwd("conv3_2", layer("conv3_1"), g, true, &dx));
  XV_TRY(relu_bwd("conv3_1", dx, layer("conv3_1"), &g));
  XV_TRY(conv_bwd("conv3_1", layer("pool2"), g, true, &dp));
  XV_TRY(bucket_done(2));
  XV_TRY(pool_relu_bwd("conv2_2", dp, layer("conv2_2"), layer("pool2"), nullptr, &g));
  XV_TRY(conv_bwd("conv2_2", layer("conv2_1"), g, true, &dx));
  XV_TRY(relu_bwd("conv2_1", dx, layer("conv2_1"), &g));
  XV_TRY(conv_bwd("conv2_1", layer("pool1"), g, true, &dp));
  XV_TRY(pool_relu_bwd("conv1_2", dp, layer("conv1_2"), layer("pool1"), nullptr, &g));
  XV_TRY(conv_bwd("conv1_2", layer("conv1_1"), g, true, &dx));
  XV_TRY(relu_bwd("conv1_1", dx, layer("conv1_1"), &g));
  // conv1_1: weight + bias gradients only (its input is the image)
  TrainLayer* t11 = find_layer(ts, "conv1_1");
  if (!dry) {
    XV_TRY(launch_conv_wgrad_c1(x, static_cast<const __nv_bfloat16*>(g.p), grads + t11->w_off, N, H,
                                W, t11->cin, t11->cout, s));
  }
  return bucket_done(3);
}

// ------------------------------------------------------------------ fit() with batch norm
// One training step of the batch-normalised expert (simple_fcn.py:201-215 with batchnorm=True,
// is_training=True): every conv / transposed conv -> batch norm on batch statistics -> ReLU; the
// final score conv -> batch norm without activation.  Encoder tensors are bf16 (pre-norm z and
// post-ReLU y are both kept for the backward pass), heads and decoder fp32.  The decoder runs at
// full resolution because its batch statistics are taken there (upscore: [N,H,W,num_units]).
struct BnTrain {
  xv_fcn* net;
  TrainState* ts;
  Arena* arena;
  cudaStream_t s;
  bool dry;
  float* grads;

  struct Rec {          // what the backward pass needs from one normalised layer
    Act z, y;
    float *mean = nullptr, *rstd = nullptr;
  };

  float* master() { return static_cast<float*>(ts->master.p); }
  Act make(DType dt, int B, int H, int W, int C) {
    Act a;
    a.dt = dt;
    a.B = B;
    a.H = H;
    a.W = W;
    a.C = C;
    a.p = arena->alloc(a.elems() * (dt == DType::F32 ? 4 : 2));
    return a;
  }
  template <typename T>
  T* scratch(size_t n) { return static_cast<T*>(arena->alloc(n * sizeof(T))); }
  static size_t npix(const Act& a) { return static_cast<size_t>(a.B) * a.H * a.W; }

  int bn_fwd(const std::string& name, const Act& z, int relu, const float* bias_extra, Rec* r) {
    BnParam* bp = find_bn(ts, name);
    XV_CHECK(bp != nullptr && bp->C == z.C, "no batch-norm variables for " + name);
    r->z = z;
    r->y = make(z.dt, z.B, z.H, z.W, z.C);
    r->mean = scratch<float>(z.C);
    r->rstd = scratch<float>(z.C);
    double* sums = scratch<double>(2 * z.C);
    if (dry) return 0;
    const bool bf = z.dt == DType::BF16;
    XV_TRY(launch_bn_stats(z.p, bf, npix(z), z.C, sums, s));
    XV_TRY(launch_bn_finalize(sums, npix(z), z.C, 1e-3f, 0.99f, bias_extra, r->mean, r->rstd,
                              master() + bp->mm_off, master() + bp->mv_off, s));
    return launch_bn_apply(z.p, bf, r->mean, r->rstd, master() + bp->g_off, master() + bp->b_off,
                           npix(z), z.C, relu, r->y.p, s);
  }
  // g: gradient wrt the layer output (same dtype / shape as r.y); mask: apply the ReLU mask of r.y
  int bn_bwd(const std::string& name, const Rec& r, const void* g, bool mask, Act* dz,
             __nv_bfloat16* dz_pad = nullptr, int c_pad = 0) {
    BnParam* bp = find_bn(ts, name);
    *dz = make(r.z.dt, r.z.B, r.z.H, r.z.W, r.z.C);
    double* sums = scratch<double>(2 * r.z.C);
    if (dry) return 0;
    return launch_bn_backward(g, mask ? r.y.p : nullptr, r.z.p, r.z.dt == DType::BF16, r.mean,
                              r.rstd, master() + bp->g_off, npix(r.z), r.z.C, sums, dz->p, dz_pad,
                              c_pad, grads + bp->g_off, grads + bp->b_off, s);
  }
  int bucket_done(int b) {
    if (dry || b >= static_cast<int>(ts->bucket_events.size()) || !ts->bucket_events[b]) return 0;
    XV_CUDA(cudaEventRecord(ts->bucket_events[b], s));
    return 0;
  }
  // 1x1 head: weight gradient + data gradient from dz (fp32) and its padded bf16 copy
  int head_bwd(const std::string& name, const Rec& r, const Act& x, const Act& g_out, Act* dx) {
    TrainLayer* tl = find_layer(ts, name);
    const int c_pad = tl->bwd->cin_gemm;
    Act dz16 = make(DType::BF16, r.z.B, r.z.H, r.z.W, c_pad);
    Act dz;
    XV_TRY(bn_bwd(name, r, g_out.p, true, &dz, static_cast<__nv_bfloat16*>(dz16.p), c_pad));
    *dx = make(DType::BF16, r.z.B, r.z.H, r.z.W, tl->cin);
    if (dry) return 0;
    if (tl->cin % 64 == 0 && !(g_debug_flags & 8))
      XV_TRY(run_wgrad_tc(net, x.p, dz16.p, grads + tl->w_off, r.z.B, r.z.H, r.z.W, tl->cin,
                          tl->cout, s, 1, c_pad));
    else
      XV_TRY(launch_outer_sum_bf16(static_cast<const __nv_bfloat16*>(x.p),
                                   static_cast<const float*>(dz.p), grads + tl->w_off, npix(r.z),
                                   tl->cin, tl->cout, s));
    return run_igemm(net, *tl->bwd, dz16.p, r.z.B, r.z.H, r.z.W, dx->p, false, s);
  }

  int run(const float* x, const int32_t* labels, int N, int H, int W, int train_encoder);
};

int BnTrain::run(const float* x, const int32_t* labels, int N, int H, int W, int train_encoder) {
  const int nu = net->nu, C = net->C;
  Rec rec[13];
  Act input[13], pooled[13];
  bool has_pool[13];
  // ---------------------------------------------------------------- forward, encoder
  Act cur;
  int h = H, w = W;
  for (int i = 0; i < 13; ++i) {
    TrainLayer* tl = find_layer(ts, kConvNames[i]);
    Act z = make(DType::BF16, N, h, w, tl->cout);
    input[i] = cur;
    if (!dry) {
      if (i == 0)
        XV_TRY(run_igemm_c1(net, *tl->fwd, x, N, h, w, z.p, s));
      else
        XV_TRY(run_igemm(net, *tl->fwd, cur.p, N, h, w, z.p, false, s));
    }
    XV_TRY(bn_fwd(kConvNames[i], z, 1, nullptr, &rec[i]));
    cur = rec[i].y;
    has_pool[i] = (i == 1 || i == 3 || i == 6 || i == 9);
    if (has_pool[i]) {
      pooled[i] = make(DType::BF16, N, h / 2, w / 2, tl->cout);
      if (!dry)
        XV_TRY(launch_maxpool_bf16(static_cast<const __nv_bfloat16*>(cur.p),
                                   static_cast<__nv_bfloat16*>(pooled[i].p), N, h, w, tl->cout, s));
      cur = pooled[i];
      h /= 2;
      w /= 2;
    }
  }
  const int h8 = H / 8, w8 = W / 8, h16 = H / 16, w16 = W / 16;
  const Act c43 = rec[9].y, c53 = rec[12].y;
  // ---------------------------------------------------------------- forward, heads + decoder
  Rec r4, r5, rup5, rup, rsc;
  TrainLayer* t4 = find_layer(ts, "score_conv4");
  TrainLayer* t5 = find_layer(ts, "score_conv5");
  TrainLayer* tsc = find_layer(ts, "score");
  Act s4z = make(DType::F32, N, h8, w8, nu), s5z = make(DType::F32, N, h16, w16, nu);
  if (!dry) {
    XV_TRY(run_igemm(net, *t4->fwd, c43.p, N, h8, w8, s4z.p, true, s));
    XV_TRY(run_igemm(net, *t5->fwd, c53.p, N, h16, w16, s5z.p, true, s));
  }
  XV_TRY(bn_fwd("score_conv4", s4z, 1, nullptr, &r4));
  XV_TRY(bn_fwd("score_conv5", s5z, 1, nullptr, &r5));
  Act up5z = make(DType::F32, N, h8, w8, nu);
  if (!dry)
    XV_TRY(launch_upsample_diag(static_cast<const float*>(r5.y.p),
                                static_cast<const float*>(net->g4.p), true, N, h16, w16, nu, 4, 2,
                                static_cast<float*>(up5z.p), s));
  XV_TRY(bn_fwd("upscore_conv5", up5z, 1, nullptr, &rup5));
  Act fused = make(DType::F32, N, h8, w8, nu);
  if (!dry)
    XV_TRY(launch_add_f32(static_cast<const float*>(r4.y.p), static_cast<const float*>(rup5.y.p),
                          static_cast<float*>(fused.p), fused.elems(), s));
  Act upz = make(DType::F32, N, H, W, nu);
  if (!dry)
    XV_TRY(launch_upsample_diag(static_cast<const float*>(fused.p),
                                static_cast<const float*>(net->g16.p), false, N, h8, w8, nu, 16, 8,
                                static_cast<float*>(upz.p), s));
  XV_TRY(bn_fwd("upscore", upz, 1, nullptr, &rup));
  // score conv: the bias is left out of z (batch norm cancels it), it only enters the moving mean
  Act scz = make(DType::F32, N, H, W, C);
  const size_t full = static_cast<size_t>(N) * H * W;
  if (!dry)
    XV_TRY(launch_score_lowres(static_cast<const float*>(rup.y.p), master() + tsc->w_off,
                               static_cast<float*>(scz.p), full, nu, C, s));
  XV_TRY(bn_fwd("score", scz, 0, master() + tsc->b_off, &rsc));
  // ---------------------------------------------------------------- loss + backward, decoder
  float* dbias_scratch = scratch<float>(C);
  Act dscz, dup = make(DType::F32, N, H, W, nu), dupz;
  if (!dry) {
    XV_CUDA(cudaMemsetAsync(dbias_scratch, 0, C * sizeof(float), s));
    XV_TRY(launch_ce_grad(static_cast<float*>(rsc.y.p), labels, static_cast<int64_t>(full), C,
                          static_cast<double*>(ts->loss.p), dbias_scratch, s));
  }
  XV_TRY(bn_bwd("score", rsc, rsc.y.p, false, &dscz));          // rsc.y now holds dL/dscore
  if (!dry)
    XV_TRY(launch_score_bwd(static_cast<const float*>(dscz.p), static_cast<const float*>(rup.y.p),
                            master() + tsc->w_off, static_cast<float*>(dup.p), grads + tsc->w_off,
                            full, nu, C, s));
  XV_TRY(bn_bwd("upscore", rup, dup.p, true, &dupz));
  Act dfused = make(DType::F32, N, h8, w8, nu);
  if (!dry)
    XV_TRY(launch_upsample_diag_transpose(static_cast<const float*>(dupz.p),
                                          static_cast<const float*>(net->g16.p), false, N, h8, w8,
                                          nu, 16, 8, static_cast<float*>(dfused.p), s));
  Act dup5z, ds5 = make(DType::F32, N, h16, w16, nu);
  XV_TRY(bn_bwd("upscore_conv5", rup5, dfused.p, true, &dup5z));
  if (!dry)
    XV_TRY(launch_upsample_diag_transpose(static_cast<const float*>(dup5z.p),
                                          static_cast<const float*>(net->g4.p), true, N, h16, w16,
                                          nu, 4, 2, static_cast<float*>(ds5.p), s));
  Act d43a, d53;
  XV_TRY(head_bwd("score_conv4", r4, c43, dfused, &d43a));
  XV_TRY(head_bwd("score_conv5", r5, c53, ds5, &d53));
  if (!train_encoder) {
    for (int b = 0; b < 4; ++b) XV_TRY(bucket_done(b));
    return 0;
  }
  // ---------------------------------------------------------------- backward, encoder
  Act da = d53;           // gradient wrt the output y of layer i (or wrt its pooled map)
  for (int i = 12; i >= 0; --i) {
    TrainLayer* tl = find_layer(ts, kConvNames[i]);
    const Rec& r = rec[i];
    Act dz;
    if (has_pool[i]) {
      // route the pooled gradient to the argmax positions and apply the ReLU mask (+ the second
      // gradient source of conv4_3: its score_conv4 branch)
      Act ghat = make(DType::BF16, r.y.B, r.y.H, r.y.W, r.y.C);
      if (!dry)
        XV_TRY(launch_pool_relu_bwd_bf16(
            static_cast<const __nv_bfloat16*>(da.p), static_cast<const __nv_bfloat16*>(r.y.p),
            static_cast<const __nv_bfloat16*>(pooled[i].p),
            i == 9 ? static_cast<const __nv_bfloat16*>(d43a.p) : nullptr,
            static_cast<__nv_bfloat16*>(ghat.p), r.y.B, r.y.H, r.y.W, r.y.C, nullptr, s));
      XV_TRY(bn_bwd(kConvNames[i], r, ghat.p, false, &dz));
    } else {
      XV_TRY(bn_bwd(kConvNames[i], r, da.p, true, &dz));
    }
    const int lh = r.z.H, lw = r.z.W;
    Act dx;
    if (i > 0) dx = make(DType::BF16, N, lh, lw, tl->cin);
    if (!dry) {
      if (i == 0) {
        XV_TRY(launch_conv_wgrad_c1(x, static_cast<const __nv_bfloat16*>(dz.p), grads + tl->w_off,
                                    N, lh, lw, tl->cin, tl->cout, s));
      } else {
        if (tl->cin % 64 == 0 && !(g_debug_flags & 8))
          XV_TRY(run_wgrad_tc(net, input[i].p, dz.p, grads + tl->w_off, N, lh, lw, tl->cin,
                              tl->cout, s));
        else
          XV_TRY(launch_conv_wgrad(static_cast<const __nv_bfloat16*>(input[i].p),
                                   static_cast<const __nv_bfloat16*>(dz.p), grads + tl->w_off, N,
                                   lh, lw, tl->cin, tl->cout, s));
        XV_TRY(run_igemm(net, *tl->bwd, dz.p, N, lh, lw, dx.p, false, s));
      }
    }
    da = dx;
    if (i == 10) XV_TRY(bucket_done(0));
    if (i == 7) XV_TRY(bucket_done(1));
    if (i == 4) XV_TRY(bucket_done(2));
  }
  return bucket_done(3);
}

}  // namespace xv

extern "C" {

// Weight gradient of ONE 3x3 'same' convolution layer, the kernels fit() uses:
// dw[tap][ci][co] = sum_p bf16(x)[p + shift(tap)][ci] * bf16(dy)[p][co], fp32 accumulation.
int xv_conv2d_weight_gradient(const float* x, const float* dy, int n, int h, int w, int cin,
                              int cout, int use_tensor_cores, float* dw, void* stream) {
  XV_TRY(ensure_init());
  XV_CHECK(x && dy && dw, "xv_conv2d_weight_gradient: NULL argument");
  XV_CHECK(cout % 64 == 0, "xv_conv2d_weight_gradient: Cout % 64 == 0 required");
  XV_CHECK(use_tensor_cores ? cin % 64 == 0 : cin % 32 == 0,
           "xv_conv2d_weight_gradient: Cin % 64 (tensor cores) / % 32 (CUDA cores) required");
  cudaStream_t s = XV_STREAM(stream);
  const size_t npix = static_cast<size_t>(n) * h * w;
  DevBuf x16, dy16;
  XV_TRY(x16.ensure(npix * cin * 2));
  XV_TRY(dy16.ensure(npix * cout * 2));
  XV_TRY(launch_f32_to_bf16(x, static_cast<__nv_bfloat16*>(x16.p), npix * cin, s));
  XV_TRY(launch_f32_to_bf16(dy, static_cast<__nv_bfloat16*>(dy16.p), npix * cout, s));
  XV_CUDA(cudaMemsetAsync(dw, 0, static_cast<size_t>(9) * cin * cout * sizeof(float), s));
  if (use_tensor_cores)
    XV_TRY(run_wgrad_tc(nullptr, x16.p, dy16.p, dw, n, h, w, cin, cout, s, 9, 0));
  else
    XV_TRY(launch_conv_wgrad(static_cast<const __nv_bfloat16*>(x16.p),
                             static_cast<const __nv_bfloat16*>(dy16.p), dw, n, h, w, cin, cout, s));
  XV_CUDA(cudaStreamSynchronize(s));
  return 0;
}


int xv_fcn_train_begin(xv_fcn* net, int64_t* num_params_out) {
  XV_CHECK(net && net->finalized, "xv_fcn_train_begin: finalize the expert first");
  XV_CHECK(net->precision == XV_PRECISION_BF16 && (net->batchnorm == 0 || net->batchnorm == 1),
           "fit() runs on the bf16 path (batch norm on every layer or on none)");
  XV_CHECK(net->diag_up5 && net->diag_up,
           "fit() needs the (non-trainable) bilinear transposed-conv kernels");
  std::unique_ptr<TrainState> ts(new TrainState());
  ts->bn = net->bn_all();
  auto add_bn = [&](const std::string& name, int c) {
    BnParam b;
    b.name = name;
    b.C = c;
    b.g_off = ts->total;
    b.b_off = ts->total + c;
    b.mm_off = ts->total + 2 * static_cast<size_t>(c);
    b.mv_off = ts->total + 3 * static_cast<size_t>(c);
    ts->total += 4 * static_cast<size_t>(c);
    ts->bns.push_back(b);
  };
  auto add = [&](const std::string& name, int k, int cin, int cout, bool need_bwd) -> int {
    TrainLayer tl;
    tl.name = name;
    tl.k = k;
    tl.cin = cin;
    tl.cout = cout;
    tl.w_off = ts->total;
    ts->total += static_cast<size_t>(k) * k * cin * cout;
    tl.b_off = ts->total;
    ts->total += cout;
    if (ts->bn) {
      add_bn(name, cout);
      if (name != "score") {
        const HostParam *w, *b;
        XV_TRY(get_param(net, name + "/kernel", {k, k, cin, cout}, &w));
        XV_TRY(get_param(net, name + "/bias", {cout}, &b));
        tl.fwd.reset(new ConvLayer());
        tl.fwd->name = name + "_train";
        tl.fwd->k = k;
        tl.fwd->cin = cin;
        tl.fwd->cout = cout;
        tl.fwd->relu = 0;
        const std::vector<float> ones(cout, 1.f), zeros(cout, 0.f);
        XV_TRY(pack_conv(net, tl.fwd.get(), w->data.data(), b->data.data(), ones, zeros, false));
      }
    }
    if (need_bwd) {
      // data-gradient conv: Cin' = Cout (padded to 64 for the 1x1 heads), Cout' = Cin, no bias/ReLU
      tl.bwd.reset(new ConvLayer());
      ConvLayer* B = tl.bwd.get();
      B->name = name + "_bwd";
      B->k = k;
      B->cin = div_up(cout, 64) * 64;
      B->cout = cin;
      B->relu = 0;
      B->taps = k * k;
      B->cin_gemm = B->cin;
      B->kdim = B->taps * B->cin;
      B->block_n = conv_igemm_block_n(cin);
      B->cout_pad = div_up(cin, B->block_n) * B->block_n;
      B->use_t = (k == 3 && cin <= 128 && cin % 64 == 0);
      if (B->use_t) B->cout_pad = div_up(cin, 128) * 128;
      std::vector<uint16_t> zw(static_cast<size_t>(B->cout_pad) * B->kdim, 0);
      std::vector<float> zb(B->cout_pad, 0.f);
      XV_TRY(B->w_packed.upload(zw));
      XV_TRY(B->bias_pad.upload(zb));
    }
    ts->layers.push_back(std::move(tl));
    return 0;
  };
  int cin = net->cin;
  for (int i = 0; i < 13; ++i) {
    XV_TRY(add(kConvNames[i], 3, cin, kConvCout[i], i > 0));
    cin = kConvCout[i];
  }
  XV_TRY(add("score_conv4", 1, 512, net->nu, true));
  XV_TRY(add("score_conv5", 1, 512, net->nu, true));
  XV_TRY(add("score", 1, net->nu, net->C, false));
  if (ts->bn) {     // the two bilinear transposed convs: fixed kernels, trained batch norm
    add_bn("upscore_conv5", net->nu);
    add_bn("upscore", net->nu);
  }
  // the 1x1 heads' data-gradient operand has K = padded nu: kdim of bwd = cin (padded)
  std::vector<float> flat(ts->total, 0.f);
  for (auto& tl : ts->layers) {
    const HostParam *w, *b;
    XV_TRY(get_param(net, tl.name + "/kernel", {tl.k, tl.k, tl.cin, tl.cout}, &w));
    XV_TRY(get_param(net, tl.name + "/bias", {tl.cout}, &b));
    std::copy(w->data.begin(), w->data.end(), flat.begin() + tl.w_off);
    std::copy(b->data.begin(), b->data.end(), flat.begin() + tl.b_off);
  }
  for (auto& bp : ts->bns) {
    const char* leaves[4] = {"/gamma", "/beta", "/moving_mean", "/moving_variance"};
    const size_t offs[4] = {bp.g_off, bp.b_off, bp.mm_off, bp.mv_off};
    for (int i = 0; i < 4; ++i) {
      const HostParam* v;
      XV_TRY(get_param(net, bp.name + leaves[i], {bp.C}, &v));
      std::copy(v->data.begin(), v->data.end(), flat.begin() + offs[i]);
    }
  }
  ts->bucket_off = {0, find_layer(ts.get(), "conv3_1")->w_off,
                    find_layer(ts.get(), "conv4_1")->w_off,
                    find_layer(ts.get(), "conv5_1")->w_off, ts->total};
  XV_TRY(ts->master.upload(flat));
  std::vector<float> zeros(ts->total, 0.f);
  XV_TRY(ts->m.upload(zeros));
  XV_TRY(ts->v.upload(zeros));
  std::vector<double> lz(2, 0.0);
  XV_TRY(ts->loss.upload(lz));
  for (auto& tl : ts->layers) XV_TRY(repack_layer(net, ts.get(), tl, 0));
  XV_CUDA(cudaStreamSynchronize(0));
  if (num_params_out) *num_params_out = static_cast<int64_t>(ts->total);
  g_train[net] = std::move(ts);
  return 0;
}

int xv_fcn_param_span(xv_fcn* net, const char* name, int64_t* offset, int64_t* size) {
  auto it = g_train.find(net);
  XV_CHECK(it != g_train.end(), "xv_fcn_param_span: call xv_fcn_train_begin first");
  std::string n(name);
  const size_t slash = n.rfind('/');
  XV_CHECK(slash != std::string::npos, "parameter name must look like 'conv1_1/kernel'");
  const std::string scope = n.substr(0, slash), leaf = n.substr(slash + 1);
  if (leaf == "gamma" || leaf == "beta" || leaf == "moving_mean" || leaf == "moving_variance") {
    BnParam* bp = find_bn(it->second.get(), scope);
    XV_CHECK(bp != nullptr, "not a batch-norm variable of this training state: " + n);
    *offset = static_cast<int64_t>(leaf == "gamma" ? bp->g_off : leaf == "beta" ? bp->b_off
                                   : leaf == "moving_mean" ? bp->mm_off : bp->mv_off);
    *size = bp->C;
    return 0;
  }
  TrainLayer* tl = find_layer(it->second.get(), scope);
  XV_CHECK(tl != nullptr && (leaf == "kernel" || leaf == "bias"), "not a trainable parameter: " + n);
  const bool is_w = leaf == "kernel";
  *offset = static_cast<int64_t>(is_w ? tl->w_off : tl->b_off);
  *size = is_w ? static_cast<int64_t>(tl->k) * tl->k * tl->cin * tl->cout : tl->cout;
  return 0;
}

// Forward + backward of one batch.  grads: device float32 [num_params], OVERWRITTEN with the
// gradient of the cross-entropy (mean over valid pixels if normalize != 0, else the plain sum); loss_out: device double[2], OVERWRITTEN with
// {sum of -log p over valid pixels, number of valid pixels}.
int xv_fcn_train_gradients(xv_fcn* net, const float* x, const int32_t* labels, int n, int h, int w,
                           int train_encoder, int normalize, float* grads, double* loss_out,
                           void* stream) {
  return xv_fcn_train_gradients_ex(net, x, labels, n, h, w, train_encoder, normalize, grads,
                                   loss_out, nullptr, 0, stream);
}

int xv_fcn_grad_buckets(xv_fcn* net, int64_t* offsets_out, int capacity, int* num_buckets_out) {
  auto it = g_train.find(net);
  XV_CHECK(it != g_train.end(), "xv_fcn_grad_buckets: call xv_fcn_train_begin first");
  const auto& off = it->second->bucket_off;
  XV_CHECK(offsets_out && num_buckets_out && capacity >= static_cast<int>(off.size()),
           "xv_fcn_grad_buckets: need room for 5 offsets");
  for (size_t i = 0; i < off.size(); ++i) offsets_out[i] = static_cast<int64_t>(off[i]);
  *num_buckets_out = static_cast<int>(off.size()) - 1;
  return 0;
}

int xv_fcn_train_gradients_ex(xv_fcn* net, const float* x, const int32_t* labels, int n, int h,
                              int w, int train_encoder, int normalize, float* grads,
                              double* loss_out, void* const* bucket_events_host, int num_events,
                              void* stream) {
  auto it = g_train.find(net);
  XV_CHECK(it != g_train.end(), "xv_fcn_train_gradients: call xv_fcn_train_begin first");
  XV_CHECK(x && labels && grads, "xv_fcn_train_gradients: NULL argument");
  XV_CHECK(n >= 1 && h % 16 == 0 && w % 16 == 0, "H and W must be multiples of 16");
  TrainState* ts = it->second.get();
  cudaStream_t s = XV_STREAM(stream);
  // events are recorded in completion order: event i <-> flat range
  // [bucket_off[nb-1-i], bucket_off[nb-i])
  ts->bucket_events.clear();
  XV_CHECK(num_events == 0 || (bucket_events_host != nullptr && !normalize),
           "bucket events need un-normalised gradients (scale after the all-reduce)");
  for (int i = 0; i < num_events; ++i)
    ts->bucket_events.push_back(reinterpret_cast<cudaEvent_t>(bucket_events_host[i]));
  if (ts->bn) {
    Arena dry_arena;
    BnTrain plan{net, ts, &dry_arena, s, true, grads};
    XV_TRY(plan.run(x, labels, n, h, w, train_encoder));
    XV_TRY(net->arena_buf.ensure(dry_arena.off + 1024));
    Arena real_arena;
    real_arena.base = static_cast<char*>(net->arena_buf.p);
    XV_CUDA(cudaMemsetAsync(grads, 0, ts->total * sizeof(float), s));
    XV_CUDA(cudaMemsetAsync(ts->loss.p, 0, 2 * sizeof(double), s));
    BnTrain real{net, ts, &real_arena, s, false, grads};
    XV_TRY(real.run(x, labels, n, h, w, train_encoder));
    if (normalize)
      XV_TRY(launch_scale_by_count(grads, ts->total, static_cast<const double*>(ts->loss.p), s));
    if (loss_out)
      XV_CUDA(cudaMemcpyAsync(loss_out, ts->loss.p, 2 * sizeof(double), cudaMemcpyDeviceToDevice,
                              s));
    return 0;
  }
  xv_fcn_outputs none;
  std::memset(&none, 0, sizeof(none));
  Forward plan{net, Arena(), s, true, 1, nullptr};
  plan.training = true;
  XV_TRY(plan.run(x, n, h, w, &none));
  Backward bplan{net, ts, &plan.arena, s, true, grads};
  XV_TRY(bplan.run(x, labels, n, h, w, train_encoder));
  XV_TRY(net->arena_buf.ensure(plan.arena.off + 1024));
  Forward real{net, Arena(), s, false, 1, nullptr};
  real.training = true;
  real.arena.base = static_cast<char*>(net->arena_buf.p);
  net->layers.clear();
  XV_TRY(real.run(x, n, h, w, &none));
  XV_CUDA(cudaMemsetAsync(grads, 0, ts->total * sizeof(float), s));
  XV_CUDA(cudaMemsetAsync(ts->loss.p, 0, 2 * sizeof(double), s));
  Backward breal{net, ts, &real.arena, s, false, grads};
  XV_TRY(breal.run(x, labels, n, h, w, train_encoder));
  if (normalize)
    XV_TRY(launch_scale_by_count(grads, ts->total, static_cast<const double*>(ts->loss.p), s));
  if (loss_out)
    XV_CUDA(cudaMemcpyAsync(loss_out, ts->loss.p, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return 0;
}

// grads *= 1 / (1e-20 + loss[1]) - the mean over the valid pixels counted in loss[1] (device
// double[2] as written by xv_fcn_train_gradients, possibly summed over ranks)
int xv_scale_by_count(float* grads, int64_t n, const double* loss, void* stream) {
  XV_TRY(ensure_init());
  return launch_scale_by_count(grads, static_cast<size_t>(n), loss, XV_STREAM(stream));
}

// tf.train.AdamOptimizer step on the flat parameters, then refresh of the bf16 operand copies.
int xv_fcn_adam_step(xv_fcn* net, const float* grads, float learning_rate, float beta1, float beta2,
                     float epsilon, void* stream) {
  auto it = g_train.find(net);
  XV_CHECK(it != g_train.end(), "xv_fcn_adam_step: call xv_fcn_train_begin first");
  TrainState* ts = it->second.get();
  cudaStream_t s = XV_STREAM(stream);
  if (ts->opt_kind != XV_OPT_ADAM) {
    XV_TRY(launch_fill_f32(static_cast<float*>(ts->v.p), ts->total, 0.f, s));
    XV_TRY(launch_fill_f32(static_cast<float*>(ts->m.p), ts->total, 0.f, s));
    ts->opt_kind = XV_OPT_ADAM;
    ts->step = 0;
  }
  ts->step += 1;
  const double t = static_cast<double>(ts->step);
  const float lr_t = static_cast<float>(learning_rate * std::sqrt(1.0 - std::pow(beta2, t)) /
                                        (1.0 - std::pow(beta1, t)));
  XV_TRY(launch_adam(static_cast<float*>(ts->master.p), grads, static_cast<float*>(ts->m.p),
                     static_cast<float*>(ts->v.p), ts->total, lr_t, beta1, beta2, epsilon, s));
  XV_TRY(repack_all(net, ts, s));
  return 0;
}

// One step of trainers[config['trainer']] (base_model.py:157-162) with TensorFlow 1.x defaults.
int xv_fcn_optimizer_step(xv_fcn* net, const float* grads, int kind, float learning_rate,
                          void* stream) {
  if (kind == XV_OPT_ADAM)
    return xv_fcn_adam_step(net, grads, learning_rate, 0.9f, 0.999f, 1e-8f, stream);
  auto it = g_train.find(net);
  XV_CHECK(it != g_train.end(), "xv_fcn_optimizer_step: call xv_fcn_train_begin first");
  XV_CHECK(kind == XV_OPT_ADAGRAD || kind == XV_OPT_RMSPROP, "unknown optimizer kind");
  TrainState* ts = it->second.get();
  cudaStream_t s = XV_STREAM(stream);
  float* v = static_cast<float*>(ts->v.p);
  float* m = static_cast<float*>(ts->m.p);
  if (ts->opt_kind != kind) {
    // slot initial values: Adagrad accumulator 0.1, RMSProp mean square 1 and momentum 0
    XV_TRY(launch_fill_f32(v, ts->total, kind == XV_OPT_ADAGRAD ? 0.1f : 1.f, s));
    XV_TRY(launch_fill_f32(m, ts->total, 0.f, s));
    ts->opt_kind = kind;
    ts->step = 0;
  }
  ts->step += 1;
  if (kind == XV_OPT_ADAGRAD)
    XV_TRY(launch_adagrad(static_cast<float*>(ts->master.p), grads, v, ts->total, learning_rate,
                          s));
  else
    XV_TRY(launch_rmsprop(static_cast<float*>(ts->master.p), grads, v, m, ts->total,
                          learning_rate, 0.9f, 0.f, 1e-10f, s));
  XV_TRY(repack_all(net, ts, s));
  return 0;
}

// Copies the current fp32 master parameters (flat, layout of xv_fcn_param_span) to the host.
int xv_fcn_get_params_host(xv_fcn* net, float* out_host, int64_t capacity, void* stream) {
  auto it = g_train.find(net);
  XV_CHECK(it != g_train.end(), "xv_fcn_get_params_host: call xv_fcn_train_begin first");
  TrainState* ts = it->second.get();
  XV_CHECK(capacity >= static_cast<int64_t>(ts->total), "xv_fcn_get_params_host: buffer too small");
  XV_CUDA(cudaMemcpyAsync(out_host, ts->master.p, ts->total * sizeof(float), cudaMemcpyDeviceToHost,
                          XV_STREAM(stream)));
  XV_CUDA(cudaStreamSynchronize(XV_STREAM(stream)));
  return 0;
}

int xv_fcn_train_end(xv_fcn* net) {
  g_train.erase(net);
  return 0;
}

}  // extern "C"
