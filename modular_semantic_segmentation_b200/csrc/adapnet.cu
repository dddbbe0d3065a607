// The Adapnet expert (xview/models/adapnet.py:99-173) at test time on the tensor-core
// convolution kernels.  Batch norm (moving statistics) is folded into every convolution.
//
//   block_0_1   3x3 on the raw fp32 input            -> conv1_1-style kernel (operand packed in-kernel)
//   block_0_2   7x7 stride 2                         -> space-to-depth + 4x4 stride-1 (K = 16 x 256)
//   stage_1 / shortcut with strides=2 (1x1)          -> TMA descriptor that samples every 2nd pixel
//   stage_2_1 / stage_2_2 (atrous 3x3, rates 1..16)  -> tap shift = rate x (tap - 1) in the TMA
//                                                       coordinates, both write channel slices of
//                                                       one buffer (tf.concat without a copy)
//   transposed convs (4/2 and 16/8, dense kernels)   -> 1x1 GEMM producing the k*k*Cout products
//                                                       per input pixel, then a gather (col2im)
//                                                       that adds the BN shift and the skip
#include "net.h"

namespace xv {

namespace {

struct ConvSpec {
  std::string name;   // variable scope below the prefix
  int k, cin, cout;
  bool bias, relu;
  int stride, dil;
};

struct BlockSpec {
  const char* name;
  char kind;          // 'a' or 'b'
  int f1, f2, out, stride, d1, d2;
  bool shortcut;
};

// adapnet.py:124-153
const BlockSpec kBlocks[16] = {
    {"block_layer_1", 'a', 64, 64, 256, 1, 1, 1, true},
    {"block_layer_2", 'a', 64, 64, 256, 1, 1, 1, false},
    {"block_layer_3", 'a', 64, 64, 256, 1, 1, 1, false},
    {"block_layer_4", 'a', 128, 128, 512, 2, 1, 1, true},
    {"block_layer_5", 'a', 128, 128, 512, 1, 1, 1, false},
    {"block_layer_6", 'a', 128, 128, 512, 1, 1, 1, false},
    {"block_layer_7", 'b', 128, 64, 512, 1, 1, 2, false},
    {"block_layer_8", 'a', 256, 256, 1024, 2, 1, 1, true},
    {"block_layer_9", 'a', 256, 256, 1024, 1, 1, 1, false},
    {"block_layer_10", 'b', 256, 256, 1024, 1, 1, 2, false},
    {"block_layer_11", 'b', 256, 256, 1024, 1, 1, 4, false},
    {"block_layer_12", 'b', 256, 256, 1024, 1, 1, 8, false},
    {"block_layer_13", 'b', 256, 256, 1024, 1, 1, 16, false},
    {"block_layer_14", 'b', 512, 512, 2048, 1, 2, 4, true},
    {"block_layer_15", 'b', 512, 512, 2048, 1, 2, 8, false},
    {"block_layer_16", 'b', 512, 512, 2048, 1, 2, 16, false},
};

std::vector<ConvSpec> conv_specs(int cin, int nu) {
  std::vector<ConvSpec> v;
  v.push_back({"block_0_1", 3, cin, 64, true, true, 1, 1});
  v.push_back({"block_0_2", 7, 64, 64, true, true, 2, 1});
  int c = 64;
  for (const BlockSpec& b : kBlocks) {
    const std::string scope = b.name;
    v.push_back({scope + "/stage_1", 1, c, b.f1, false, true, b.stride, 1});
    if (b.kind == 'a') {
      v.push_back({scope + "/stage_2", 3, b.f1, b.f1, false, true, 1, 1});
      v.push_back({scope + "/stage_3", 1, b.f1, b.out, false, true, 1, 1});
    } else {
      v.push_back({scope + "/stage_2_1", 3, b.f1, b.f2 / 2, false, true, 1, b.d1});
      v.push_back({scope + "/stage_2_2", 3, b.f1, b.f2 / 2, false, true, 1, b.d2});
      v.push_back({scope + "/stage_3", 1, b.f2, b.out, false, true, 1, 1});
    }
    if (b.shortcut) v.push_back({scope + "/shortcut", 1, c, b.out, false, true, b.stride, 1});
    c = b.out;
    if (scope == "block_layer_7") v.push_back({"shortcut", 1, c, nu, true, false, 1, 1});
  }
  v.push_back({"first_deconvolution_conv", 1, 2048, 2048, true, true, 1, 1});
  return v;
}

const ConvSpec* find_spec(const std::vector<ConvSpec>& specs, const std::string& name) {
  for (const ConvSpec& sp : specs)
    if (sp.name == name) return &sp;
  return nullptr;
}

int new_layer(xv_fcn* net, const std::string& name, int k, int cin, int cout, bool relu,
              ConvLayer** out) {
  std::unique_ptr<ConvLayer> L(new ConvLayer());
  L->name = name;
  L->k = k;
  L->cin = cin;
  L->cout = cout;
  L->relu = relu ? 1 : 0;
  *out = L.get();
  net->convs.push_back(std::move(L));
  return 0;
}

// Transposed conv [k,k,Cout,Cin] (custom_layers.py:92) + BN as a 1x1 layer producing the
// k*k*Cout tap products: w[ci, (ky*k+kx)*Cout + co] = w_t[ky,kx,co,ci] * scale[co]; rows
// ci >= Cin (channel padding of the GEMM input) are zero.
int pack_upconv(xv_fcn* net, const std::string& name, int k, int cout, int cin, int cin_pad,
                bool relu, DevBuf* shift_dev, DevBuf* scale_dev, DevBuf* w_raw_dev) {
  const HostParam* w;
  XV_TRY(get_param(net, name + "/kernel", {k, k, cout, cin}, &w));
  std::vector<float> scale, shift;
  XV_TRY(bn_factors_of(net, name, cout, true, &scale, &shift));
  XV_TRY(shift_dev->upload(shift));
  XV_TRY(scale_dev->upload(scale));
  ConvLayer* L;
  XV_TRY(new_layer(net, name, 1, cin_pad, k * k * cout, relu, &L));
  if (net->precision == XV_PRECISION_FP32) return w_raw_dev->upload(w->data);
  const int n = k * k * cout;
  std::vector<float> hwio(static_cast<size_t>(cin_pad) * n, 0.f);
  for (int t = 0; t < k * k; ++t)
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        hwio[static_cast<size_t>(ci) * n + t * cout + co] =
            w->data[(static_cast<size_t>(t) * cout + co) * cin + ci] * scale[co];
  L->generic = true;
  const std::vector<float> ones(n, 1.f), zeros(n, 0.f);
  XV_TRY(pack_conv(net, L, hwio.data(), nullptr, ones, zeros, false));
  L->macs_per_pixel = static_cast<double>(k) * k * cout * cin;
  return 0;
}

// The 16/8 transposed conv [16,16,C,nu] + BN as a 3x3 stride-1 conv on the 1/8-resolution input
// that produces, per cell, its 8x8 output pixels ("phases"), channel (py*8+px)*C + c:
//   out[8i + py] = sum_ky w[ky] x[(8i + py + 4 - ky) / 8]  ->  input offsets -1 (py < 4, ky = py + 12),
//   0 (ky = py + 4) and +1 (py >= 4, ky = py - 4); unused (offset, phase) pairs get zero weights.
int pack_phase_upconv(xv_fcn* net, const std::string& name, int C, int nu, int nu_pad) {
  const HostParam* w;
  XV_TRY(get_param(net, name + "/kernel", {16, 16, C, nu}, &w));
  std::vector<float> scale, shift;
  XV_TRY(bn_factors_of(net, name, C, true, &scale, &shift));
  XV_TRY(net->up2_shift.upload(shift));
  XV_TRY(net->up2_scale.upload(scale));
  ConvLayer* L;
  XV_TRY(new_layer(net, name, 3, nu_pad, 64 * C, false, &L));
  if (net->precision == XV_PRECISION_FP32) return net->w_up2.upload(w->data);
  const int n = 64 * C;
  std::vector<float> hwio(static_cast<size_t>(9) * nu_pad * n, 0.f);
  std::vector<float> bias(n);
  auto tap_of = [](int phase, int t) {   // kernel index for input offset t - 1, or -1
    if (t == 0) return phase < 4 ? phase + 12 : -1;
    if (t == 1) return phase + 4;
    return phase >= 4 ? phase - 4 : -1;
  };
  for (int py = 0; py < 8; ++py)
    for (int px = 0; px < 8; ++px)
      for (int c = 0; c < C; ++c) {
        const int col = (py * 8 + px) * C + c;
        bias[col] = shift[c];
        for (int ty = 0; ty < 3; ++ty)
          for (int tx = 0; tx < 3; ++tx) {
            const int ky = tap_of(py, ty), kx = tap_of(px, tx);
            if (ky < 0 || kx < 0) continue;
            for (int u = 0; u < nu; ++u)
              hwio[(static_cast<size_t>(ty * 3 + tx) * nu_pad + u) * n + col] =
                  w->data[((static_cast<size_t>(ky) * 16 + kx) * C + c) * nu + u] * scale[c];
          }
      }
  L->generic = true;
  L->pad = 1;
  const std::vector<float> ones(n, 1.f), zeros(n, 0.f);
  XV_TRY(pack_conv(net, L, hwio.data(), bias.data(), ones, zeros, false));
  L->macs_per_pixel = 4.0 * nu * n;     // useful multiply-adds per 1/8-resolution cell
  return 0;
}

struct Plan {
  xv_fcn* net;
  Arena arena;
  cudaStream_t s;
  bool dry;
  std::vector<ConvSpec> specs;

  bool bf16() const { return net->precision == XV_PRECISION_BF16; }
  DType act_type() const { return bf16() ? DType::BF16 : DType::F32; }

  Act make(const std::string& name, DType dt, int B, int H, int W, int C) {
    Act a;
    a.dt = dt;
    a.B = B;
    a.H = H;
    a.W = W;
    a.C = C;
    a.p = arena.alloc(a.elems() * (dt == DType::F32 ? 4 : 2));
    if (!name.empty()) net->layers[name] = a;
    return a;
  }

  // channel slice [offset, offset + C) of `whole`
  Act slice(const Act& whole, int offset, int C) const {
    Act a = whole;
    a.C = C;
    a.pitch = whole.stride_c();
    if (a.p) a.p = static_cast<char*>(a.p) + static_cast<size_t>(offset) * (a.dt == DType::F32 ? 4 : 2);
    return a;
  }

  // conv + BN + (ReLU) of scope `name`; `out` may be a channel slice; out_f32 forces an fp32 result
  int conv(const std::string& name, const Act& in, const Act& out, const Act* residual = nullptr) {
    if (dry) return 0;
    ConvLayer* L = net->conv(name);
    const ConvSpec* sp = find_spec(specs, name);
    XV_CHECK(L != nullptr && sp != nullptr, "adapnet: unknown conv layer " + name);
    if (bf16()) {
      return run_conv_generic(net, *L, in.p, in.B, in.H, in.W, in.stride_c(), out.p,
                              out.stride_c(), out.dt == DType::F32, s,
                              residual ? residual->p : nullptr);
    }
    XV_CHECK(residual == nullptr, "adapnet fp32 path adds the shortcut with its own kernel");
    // fp32 validation path: dense temporaries, then a strided copy into the slice if needed
    XV_CHECK(in.pitch == 0, "adapnet fp32 path: sliced inputs are not used");
    float* dst = static_cast<float*>(out.p);
    DevBuf tmp;
    const size_t npix = static_cast<size_t>(out.B) * out.H * out.W;
    if (out.pitch) {
      XV_TRY(tmp.ensure(npix * out.C * 4));
      dst = static_cast<float*>(tmp.p);
    }
    XV_TRY(launch_conv_f32_ex(static_cast<const float*>(in.p), static_cast<const float*>(L->w_f32.p),
                              static_cast<const float*>(L->bias_f32.p), dst, in.B, in.H, in.W,
                              sp->cin, sp->cout, sp->k, sp->stride, sp->dil, 0, s));
    XV_TRY(launch_affine_f32(dst, static_cast<const float*>(L->bn_scale.p),
                             static_cast<const float*>(L->bn_shift.p), npix, sp->cout,
                             sp->relu ? 1 : 0, s));
    if (out.pitch) {
      XV_CUDA(cudaMemcpy2DAsync(out.p, static_cast<size_t>(out.pitch) * 4, dst,
                                static_cast<size_t>(out.C) * 4, static_cast<size_t>(out.C) * 4,
                                npix, cudaMemcpyDeviceToDevice, s));
      XV_CUDA(cudaStreamSynchronize(s));   // tmp is freed on return
    }
    return 0;
  }

  int add_relu(const Act& a, const Act& b, const Act& out) {
    if (dry) return 0;
    if (bf16())
      return launch_add_relu_bf16(static_cast<const __nv_bfloat16*>(a.p),
                                  static_cast<const __nv_bfloat16*>(b.p),
                                  static_cast<__nv_bfloat16*>(out.p), out.elems(), s);
    return launch_add_relu_f32(static_cast<const float*>(a.p), static_cast<const float*>(b.p),
                               static_cast<float*>(out.p), out.elems(), s);
  }

  int run(const float* x, int N, int H, int W, const xv_fcn_outputs* o);
};

int Plan::run(const float* x, int N, int H, int W, const xv_fcn_outputs* o) {
  const DType dt = act_type();
  const int nu = net->nu, C = net->C;
  Act cur;
  // ---- block_0 (adapnet.py:120-122)
  Act b01 = make("block_0_1", dt, N, H, W, 64);
  Act b02 = make("block_0_2", dt, N, H / 2, W / 2, 64);
  Act pool = make("block_0_pool", dt, N, H / 4, W / 4, 64);
  if (bf16()) {
    if (!dry) {
      XV_TRY(run_igemm_c1(net, *net->conv("block_0_1"), x, N, H, W, b01.p, s));
      XV_TRY(run_conv_generic(net, *net->conv("block_0_2"), b01.p, N, H, W, 64, b02.p, 64, false,
                              s));
      XV_TRY(launch_maxpool_bf16(static_cast<const __nv_bfloat16*>(b02.p),
                                 static_cast<__nv_bfloat16*>(pool.p), N, H / 2, W / 2, 64, s));
    }
  } else {
    Act in;
    in.p = const_cast<float*>(x);
    in.dt = DType::F32;
    in.B = N;
    in.H = H;
    in.W = W;
    in.C = net->cin;
    XV_TRY(conv("block_0_1", in, b01));
    XV_TRY(conv("block_0_2", b01, b02));
    if (!dry)
      XV_TRY(launch_maxpool_f32(static_cast<const float*>(b02.p), static_cast<float*>(pool.p), N,
                                H / 2, W / 2, 64, s));
  }
  cur = pool;

  // ---- residual blocks (adapnet.py:124-153)
  Act skip;
  for (int i = 0; i < 16; ++i) {
    const BlockSpec& b = kBlocks[i];
    const std::string scope = b.name;
    const int Ho = cur.H / b.stride, Wo = cur.W / b.stride;
    Act out = make("block_" + std::to_string(i + 1), dt, N, Ho, Wo, b.out);
    const size_t mark = arena.off;             // block-local temporaries are released below
    Act s1 = make("", dt, N, Ho, Wo, b.f1);
    XV_TRY(conv(scope + "/stage_1", cur, s1));
    Act s2 = make("", dt, N, Ho, Wo, b.f2);
    if (b.kind == 'a') {
      XV_TRY(conv(scope + "/stage_2", s1, s2));
    } else {
      XV_TRY(conv(scope + "/stage_2_1", s1, slice(s2, 0, b.f2 / 2)));
      XV_TRY(conv(scope + "/stage_2_2", s1, slice(s2, b.f2 / 2, b.f2 / 2)));
    }
    Act sc = cur;
    if (b.shortcut) {
      sc = make("", dt, N, Ho, Wo, b.out);
      XV_TRY(conv(scope + "/shortcut", cur, sc));
    }
    if (bf16()) {
      // relu(stage_3 + shortcut) in the epilogue of the stage_3 convolution
      XV_TRY(conv(scope + "/stage_3", s2, out, &sc));
    } else {
      Act s3 = make("", dt, N, Ho, Wo, b.out);
      XV_TRY(conv(scope + "/stage_3", s2, s3));
      XV_TRY(add_relu(s3, sc, out));
    }
    arena.off = mark;
    cur = out;
    if (i == 6) {                              // skip branch (adapnet.py:135-137)
      skip = make("shortcut", DType::F32, N, cur.H, cur.W, nu);
      XV_TRY(conv("shortcut", cur, skip));
    }
  }

  // ---- decoder (adapnet.py:154-166)
  const int h16 = cur.H, w16 = cur.W, h8 = skip.H, w8 = skip.W;
  Act d = make("first_deconvolution_conv", dt, N, h16, w16, 2048);
  XV_TRY(conv("first_deconvolution_conv", cur, d));
  Act merge = make("merge", DType::F32, N, h8, w8, nu);
  Act score;
  if (o->score) {
    score.p = o->score;
    score.dt = DType::F32;
    score.B = N;
    score.H = H;
    score.W = W;
    score.C = C;
  } else if (!bf16()) {
    score = make("score", DType::F32, N, H, W, C);
  }
  if (bf16()) {
    const int nu_pad = div_up(nu, 64) * 64;
    const bool col_bf16 = (16 * nu) % 64 == 0;      // TMA-store epilogue needs 64-channel chunks
    Act col1 = make("", col_bf16 ? DType::BF16 : DType::F32, N, h16, w16, 16 * nu);
    Act merge_bf = make("", DType::BF16, N, h8, w8, nu_pad);
    Act d2s = make("", DType::BF16, N, h8, w8, 64 * C);
    if (!dry) {
      XV_TRY(run_conv_generic(net, *net->conv("first_deconvolution_upconv"), d.p, N, h16, w16,
                              2048, col1.p, 16 * nu, !col_bf16, s));
      XV_TRY(launch_col2im(col1.p, col_bf16, static_cast<const float*>(net->up1_shift.p),
                           static_cast<const float*>(skip.p), static_cast<float*>(merge.p),
                           static_cast<__nv_bfloat16*>(merge_bf.p), nu_pad, N, h16, w16, nu, 4, 2,
                           s));
      // 16/8 transposed conv as a 3x3 conv producing the 8x8 output phases of every cell
      XV_TRY(run_conv_generic(net, *net->conv("second_deconvolution_upconv"), merge_bf.p, N, h8,
                              w8, nu_pad, d2s.p, 64 * C, false, s));
      XV_TRY(launch_d2s_softmax_argmax(static_cast<const __nv_bfloat16*>(d2s.p), N, h8, w8, C,
                                       o->score, o->prob, o->label_i64, o->label_u8, s));
    }
    return 0;
  } else if (!dry) {
    Act up1 = make("", DType::F32, N, h8, w8, nu);
    XV_TRY(launch_deconv_f32(static_cast<const float*>(d.p), static_cast<const float*>(net->w_up1.p),
                             static_cast<float*>(up1.p), N, h16, w16, 2048, nu, 4, 2, 0, nullptr,
                             s));
    XV_TRY(launch_affine_f32(static_cast<float*>(up1.p), static_cast<const float*>(net->up1_scale.p),
                             static_cast<const float*>(net->up1_shift.p),
                             static_cast<size_t>(N) * h8 * w8, nu, 0, s));
    XV_TRY(launch_add_f32(static_cast<const float*>(up1.p), static_cast<const float*>(skip.p),
                          static_cast<float*>(merge.p), merge.elems(), s));
    XV_TRY(launch_deconv_f32(static_cast<const float*>(merge.p),
                             static_cast<const float*>(net->w_up2.p), static_cast<float*>(score.p),
                             N, h8, w8, nu, C, 16, 8, 0, nullptr, s));
    XV_TRY(launch_affine_f32(static_cast<float*>(score.p),
                             static_cast<const float*>(net->up2_scale.p),
                             static_cast<const float*>(net->up2_shift.p),
                             static_cast<size_t>(N) * H * W, C, 0, s));
  } else {
    make("", DType::F32, N, h8, w8, nu);
  }
  if (!dry && (o->prob || o->label_i64 || o->label_u8))
    XV_TRY(launch_softmax_argmax(static_cast<const float*>(score.p),
                                 static_cast<size_t>(N) * H * W, C, o->prob, o->label_i64,
                                 o->label_u8, s));
  return 0;
}

}  // namespace

int adapnet_finalize(xv_fcn* net) {
  XV_TRY(ensure_init());
  net->convs.clear();
  net->tmaps.clear();
  const int nu = net->nu, C = net->C;
  const bool bf16 = net->precision == XV_PRECISION_BF16;
  std::vector<float> scale, shift;
  for (const ConvSpec& sp : conv_specs(net->cin, nu)) {
    const HostParam *w, *b = nullptr;
    XV_TRY(get_param(net, sp.name + "/kernel", {sp.k, sp.k, sp.cin, sp.cout}, &w));
    if (sp.bias) XV_TRY(get_param(net, sp.name + "/bias", {sp.cout}, &b));
    XV_TRY(bn_factors_of(net, sp.name, sp.cout, true, &scale, &shift));
    ConvLayer* L;
    XV_TRY(new_layer(net, sp.name, sp.k, sp.cin, sp.cout, sp.relu, &L));
    const float* bias = b ? b->data.data() : nullptr;
    if (!bf16) {
      XV_TRY(pack_conv(net, L, w->data.data(), bias, scale, shift, true));
      continue;
    }
    if (sp.name == "block_0_1") {
      // 3x3 on <= 3 raw channels: conv1_1 packing (hi + lo bf16 split of the input)
      XV_TRY(pack_conv(net, L, w->data.data(), bias, scale, shift, true));
      continue;
    }
    L->generic = true;
    L->dil = sp.dil;
    L->sample = sp.k == 1 ? sp.stride : 1;
    if (sp.stride == 2 && sp.k > 1) {
      // 7x7 stride 2 (adapnet.py:121), TF 'SAME' on even sizes pads (k - 2) / 2 = 2 before:
      //   out[i] = sum_a w[a] x[2i + a - 2]
      // the transposed-role kernel reads the four pixel parities through their own descriptors
      L->stride2 = true;
      L->pad = (sp.k - 2) / 2;
      XV_TRY(pack_conv(net, L, w->data.data(), bias, scale, shift, true));
      XV_CHECK(L->use_t, "adapnet: the stride-2 filter needs Cout <= 128");
      continue;
    }
    L->pad = sp.k == 3 ? sp.dil : 0;
    XV_TRY(pack_conv(net, L, w->data.data(), bias, scale, shift, true));
  }
  const int nu_pad = div_up(nu, 64) * 64;
  XV_TRY(pack_upconv(net, "first_deconvolution_upconv", 4, nu, 2048, 2048, false, &net->up1_shift,
                     &net->up1_scale, &net->w_up1));
  XV_TRY(pack_phase_upconv(net, "second_deconvolution_upconv", C, nu, nu_pad));
  net->finalized = true;
  return 0;
}

int adapnet_forward(xv_fcn* net, const float* x, int n, int h, int w, const xv_fcn_outputs* o,
                    cudaStream_t s) {
  XV_CHECK(!o->mean_prob && !o->var_prob && !o->mean_var,
           "adapnet has no dropout layers: Monte-Carlo outputs are not available");
  Plan plan{net, Arena(), s, true, conv_specs(net->cin, net->nu)};
  XV_TRY(plan.run(x, n, h, w, o));
  XV_TRY(net->arena_buf.ensure(plan.arena.off + 1024));
  Plan real{net, Arena(), s, false, plan.specs};
  real.arena.base = static_cast<char*>(net->arena_buf.p);
  net->layers.clear();
  return real.run(x, n, h, w, o);
}

}  // namespace xv
