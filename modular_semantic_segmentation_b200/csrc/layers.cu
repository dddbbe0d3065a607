// Non-GEMM layers of the FCN expert: conv1_1 operand packing, pooling, MC-dropout, the
// channel-diagonal (bilinear) transposed convolutions and the fused decoder tail, plus the
// generic fp32 "validation mode" layers that follow the reference op order literally.
//
// Reference: xview/models/simple_fcn.py:10-134, xview/models/custom_layers.py:8-25,71-139.
#include "argmax.cuh"
#include "common.cuh"
#include "kernels.h"

namespace xv {
int debug_flags();   // abi.cu; bit 8: scalar dropout kernel, bit 9: per-sample-barrier MC decode


namespace {

constexpr int kThreads = 256;

inline int grid_for(size_t work, int threads = kThreads) {
  size_t g = (work + threads - 1) / threads;
  size_t cap = static_cast<size_t>(device_info().num_sms) * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------ conv1_1 operand
// Packs the 3x3 neighbourhood of every pixel of the raw fp32 input (Cin <= 3) into one
// 64-element bf16 row [hi(9*Cin) | lo(9*Cin) | 0...] so conv1_1 runs on the tensor cores as
// a 1x1 GEMM with K = 64.  hi + lo carries 16 mantissa bits, enough for raw uint16 depth.
template <int CIN>
__global__ void __launch_bounds__(256)
im2col_c1_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H,
                 int W) {
  // one thread = one pixel: gathers its 3x3xCIN neighbourhood once, emits the 128-byte row
  constexpr int K9 = 9 * CIN;
  const size_t total = static_cast<size_t>(N) * H * W;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(pix % W);
    const int py = static_cast<int>((pix / W) % H);
    const size_t img = pix / (static_cast<size_t>(W) * H);
    float hi[K9], lo[K9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const float* src = x + ((img * H + (in ? yy : py)) * W + (in ? xx : px)) * CIN;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float raw = in ? __ldg(src + ci) : 0.f;
        const float h = __bfloat162float(__float2bfloat16_rn(raw));
        hi[tap * CIN + ci] = h;
        lo[tap * CIN + ci] = raw - h;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out) + pix * 8;
#pragma unroll
    for (int piece = 0; piece < 8; ++piece) {
      uint32_t packed[4];
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        float v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = piece * 8 + e2 * 2 + h;   // compile-time after unrolling
          v[h] = k < K9 ? hi[k < K9 ? k : 0] : (k < 2 * K9 ? lo[k < 2 * K9 ? k - K9 : 0] : 0.f);
        }
        packed[e2] = pack_bf16x2(v[0], v[1]);
      }
      dst[piece] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
}

// raw sensor dtypes -> float32 (the cast of data_baseclass.py:77-78, done after the H2D copy so
// that uint8 rgb / uint16 depth cross PCIe at 1 / 2 bytes per value)
template <typename T>
__global__ void to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<float>(in[i]);
}
__global__ void u8x16_to_f32_kernel(const uint4* __restrict__ in, float4* __restrict__ out,
                                    size_t n16) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(in + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      out[i * 4 + j] = make_float4(static_cast<float>(w[j] & 0xffu),
                                   static_cast<float>((w[j] >> 8) & 0xffu),
                                   static_cast<float>((w[j] >> 16) & 0xffu),
                                   static_cast<float>(w[j] >> 24));
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                   size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                   size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

// ------------------------------------------------------------------ 2x2/2 max pooling
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a),
                             *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  return make_uint4(bf16x2_max(a.x, b.x), bf16x2_max(a.y, b.y), bf16x2_max(a.z, b.z),
                    bf16x2_max(a.w, b.w));
}
// one thread = one output pixel x 8 channels (16 B)
__global__ void maxpool_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N,
                                    int H, int W, int C8) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C8;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    size_t t = idx / C8;
    const int xo = static_cast<int>(t % Wo);
    t /= Wo;
    const int yo = static_cast<int>(t % Ho);
    const size_t n = t / Ho;
    const size_t base = ((n * H + 2 * yo) * W + 2 * xo) * C8 + c;
    const size_t row = static_cast<size_t>(W) * C8;
    uint4 a = __ldg(in + base), b = __ldg(in + base + C8);
    uint4 cc = __ldg(in + base + row), d = __ldg(in + base + row + C8);
    out[idx] = bf16x8_max(bf16x8_max(a, b), bf16x8_max(cc, d));
  }
}
__global__ void maxpool_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int N,
                                   int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    size_t t = idx / C;
    const int xo = static_cast<int>(t % Wo);
    t /= Wo;
    const int yo = static_cast<int>(t % Ho);
    const size_t n = t / Ho;
    const size_t base = ((n * H + 2 * yo) * W + 2 * xo) * C + c;
    const size_t row = static_cast<size_t>(W) * C;
    out[idx] = fmaxf(fmaxf(in[base], in[base + C]), fmaxf(in[base + row], in[base + row + C]));
  }
}

// ------------------------------------------------------------------ MC dropout
// Philox4x32-10 (Salmon et al. 2011): counter = element-group index (+offset), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t r) { return (r >> 8) * (1.0f / 16777216.0f); }

// tf.nn.dropout: y = (x / keep) * floor(keep + U[0,1)); element kept iff U >= rate.
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ in, T* __restrict__ out, size_t n,
                               int replicate, float rate, const uint8_t* __restrict__ ext_mask,
                               uint64_t seed, uint64_t offset, size_t pass_elems) {
  const float keep = 1.f - rate;
  const size_t total = n * static_cast<size_t>(replicate);
  // leading `pass_elems` outputs are plain copies (the dropout-free sample); the dropout region
  // behind them has its own group / mask indexing, so its masks do not depend on pass_elems
  const size_t groups_a = (pass_elems + 3) / 4;
  const size_t total_b = total - pass_elems;
  const size_t groups = groups_a + (total_b + 3) / 4;
  for (size_t gidx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; gidx < groups;
       gidx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (gidx < groups_a) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const size_t i = gidx * 4 + e;
        if (i < pass_elems) out[i] = in[i % n];
      }
      continue;
    }
    const size_t g = gidx - groups_a;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (ext_mask == nullptr) {
      const uint64_t c = g + offset;
      rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32),
                                     0x58564231u, 0u),
                          make_uint2(static_cast<uint32_t>(seed),
                                     static_cast<uint32_t>(seed >> 32)));
    }
    const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const size_t j = g * 4 + e;
      if (j < total_b) {
        const size_t i = j + pass_elems;
        const bool kept = ext_mask ? (ext_mask[j] != 0) : (u01(rr[e]) >= rate);
        const float v = static_cast<float>(in[i % n]);
        out[i] = static_cast<T>(kept ? v / keep : 0.f);
      }
    }
  }
}

// bf16, 8 elements per thread.  nvec = n / 8 vectors per copy, total = vectors of the whole output,
// pass_vecs = leading vectors copied through; element j of the dropout region belongs to Philox
// group j / 4 exactly as in dropout_kernel.
__global__ void dropout_bf16x8_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                      size_t nvec, size_t total, float rate,
                                      const uint8_t* __restrict__ ext_mask, uint64_t seed,
                                      uint64_t offset, size_t pass_vecs) {
  // x / keep as tf.nn.dropout computes it; when keep is a power of two (rate 0.5, the usual MC
  // setting) the product with 1 / keep is the same float and saves eight divisions per thread
  const float keep = 1.f - rate;
  const float inv_keep = 1.f / keep;
  int expo;
  const bool pow2 = frexpf(keep, &expo) == 0.5f;
  const uint32_t thresh = rate <= 0.f ? 0u : static_cast<uint32_t>(ceilf(rate * 16777216.f));
  for (size_t v = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 x = __ldg(in + v % nvec);
    if (v < pass_vecs) {
      out[v] = x;
      continue;
    }
    const size_t j = v - pass_vecs;              // vector index inside the dropout region
    bool kept[8];
    if (ext_mask != nullptr) {
      const uint2 m = __ldg(reinterpret_cast<const uint2*>(ext_mask) + j);
#pragma unroll
      for (int e = 0; e < 8; ++e) kept[e] = (((e < 4 ? m.x : m.y) >> (8 * (e & 3))) & 0xffu) != 0;
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t c = 2 * j + h + offset;
        const uint4 rnd = philox4x32_10(
            make_uint4(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), 0x58564231u, 0u),
            make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
        // u01(r) >= rate  <=>  (r >> 8) >= ceil(rate * 2^24): both sides are exact
        kept[4 * h] = (rnd.x >> 8) >= thresh;
        kept[4 * h + 1] = (rnd.y >> 8) >= thresh;
        kept[4 * h + 2] = (rnd.z >> 8) >= thresh;
        kept[4 * h + 3] = (rnd.w >> 8) >= thresh;
      }
    }
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
      // same arithmetic as the scalar kernel: kept ? v / keep : 0, rounded to bf16
      const float fa = __bfloat162float(t.x), fb = __bfloat162float(t.y);
      const float a = kept[2 * q] ? (pow2 ? fa * inv_keep : fa / keep) : 0.f;
      const float b = kept[2 * q + 1] ? (pow2 ? fb * inv_keep : fb / keep) : 0.f;
      o[q] = pack_bf16x2(a, b);
    }
    out[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------ generic fp32 layers
// one thread = one output pixel x 8 output channels; validation mode only.
__global__ void conv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ out, int N,
                                int H, int W, int cin, int cout, int k, int relu) {
  const int cg = (cout + 7) / 8;
  const size_t total = static_cast<size_t>(N) * H * W * cg;
  const int pad = (k - 1) / 2;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % cg);
    const size_t pix = idx / cg;
    const int px = static_cast<int>(pix % W);
    const int py = static_cast<int>((pix / W) % H);
    const size_t img = pix / (static_cast<size_t>(W) * H);
    const int co0 = g * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int yy = py + ky - pad;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int xx = px + kx - pad;
        if (xx < 0 || xx >= W) continue;
        const float* xp = x + ((img * H + yy) * W + xx) * cin;
        const float* wp = w + static_cast<size_t>(ky * k + kx) * cin * cout + co0;
        for (int ci = 0; ci < cin; ++ci) {
          const float xv = __ldg(xp + ci);
          const float* wr = wp + static_cast<size_t>(ci) * cout;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (co0 + j < cout) acc[j] = fmaf(xv, __ldg(wr + j), acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (co0 + j < cout) {
        float v = acc[j] + (bias ? bias[co0 + j] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        out[pix * cout + co0 + j] = v;
      }
    }
  }
}

// conv2d_transpose 'same': out = in*stride, pad = (k-stride)/2, w[ky,kx,co,ci];
// out[oy,ox,co] = sum over (iy,ky): oy = iy*stride - pad + ky.  Optional ReLU, then + addend.
__global__ void deconv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  float* __restrict__ out, int N, int hin, int win, int cin,
                                  int cout, int k, int stride, int relu,
                                  const float* __restrict__ addend) {
  const int ho = hin * stride, wo = win * stride, pad = (k - stride) / 2;
  const size_t total = static_cast<size_t>(N) * ho * wo * cout;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % cout);
    size_t t = idx / cout;
    const int ox = static_cast<int>(t % wo);
    t /= wo;
    const int oy = static_cast<int>(t % ho);
    const size_t img = t / ho;
    float acc = 0.f;
    for (int ky = (oy + pad) % stride; ky < k; ky += stride) {
      const int iy = (oy + pad - ky) / stride;
      if (oy + pad - ky < 0 || iy >= hin) continue;
      for (int kx = (ox + pad) % stride; kx < k; kx += stride) {
        const int ix = (ox + pad - kx) / stride;
        if (ox + pad - kx < 0 || ix >= win) continue;
        const float* xp = x + ((img * hin + iy) * win + ix) * cin;
        const float* wp = w + (static_cast<size_t>(ky * k + kx) * cout + co) * cin;
        for (int ci = 0; ci < cin; ++ci) acc = fmaf(__ldg(xp + ci), __ldg(wp + ci), acc);
      }
    }
    if (relu) acc = fmaxf(acc, 0.f);
    if (addend) acc += addend[idx];
    out[idx] = acc;
  }
}

// per-channel affine (+ReLU): test-time batch norm for layers where it cannot be folded.
__global__ void affine_f32_kernel(float* __restrict__ x, const float* __restrict__ scale,
                                  const float* __restrict__ shift, size_t total, int C,
                                  int relu) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    float v = x[i] * scale[c] + shift[c];
    if (relu) v = fmaxf(v, 0.f);
    x[i] = v;
  }
}

// dst[p][offset .. offset+c_src) = src[p][0 .. c_src)   (bf16, 16-byte pieces)
__global__ void concat_bf16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                   size_t npix, int src8, int dst8, int off8) {
  const size_t total = npix * src8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = i / src8;
    const int c = static_cast<int>(i - pix * src8);
    dst[pix * dst8 + off8 + c] = __ldg(src + i);
  }
}

__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b,
                               float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = a[i] + b[i];
}

// ------------------------------------------------------------------ fast decoder pieces
// fused[n,Y,X,u] = s4[n,Y,X,u] + relu( sum_{ky,kx} g[ky,kx,u] * s5[n,iy,ix,u] ),
// 4x4 stride-2 channel-diagonal transposed conv (simple_fcn.py:82-85).
__global__ void upscore2_add_kernel(const float* __restrict__ s5, const float* __restrict__ s4,
                                    const float* __restrict__ g, float* __restrict__ fused,
                                    float* __restrict__ up5, int N, int h, int w, int nu) {
  const int ho = 2 * h, wo = 2 * w;
  const size_t total = static_cast<size_t>(N) * ho * wo * nu;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % nu);
    size_t t = idx / nu;
    const int ox = static_cast<int>(t % wo);
    t /= wo;
    const int oy = static_cast<int>(t % ho);
    const size_t img = t / ho;
    float acc = 0.f;
    // oy = 2*iy - 1 + ky  ->  ky in {(oy+1)%2, (oy+1)%2 + 2}
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ((oy + 1) & 1) + 2 * a;
      const int iy = (oy + 1 - ky) / 2;
      if (oy + 1 - ky < 0 || iy >= h) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kx = ((ox + 1) & 1) + 2 * b;
        const int ix = (ox + 1 - kx) / 2;
        if (ox + 1 - kx < 0 || ix >= w) continue;
        acc = fmaf(__ldg(g + (ky * 4 + kx) * nu + u),
                   __ldg(s5 + ((img * h + iy) * w + ix) * nu + u), acc);
      }
    }
    fused[idx] = s4[idx] + fmaxf(acc, 0.f);
    if (up5) up5[idx] = fmaxf(acc, 0.f);     // kept for the backward pass (ReLU mask)
  }
}

// nu % 4 == 0: one thread = one output pixel x 4 channels, 16-byte accesses
__global__ void upscore2_add_v4_kernel(const float4* __restrict__ s5, const float4* __restrict__ s4,
                                       const float4* __restrict__ g, float4* __restrict__ fused,
                                       int N, int h, int w, int nu4) {
  const int ho = 2 * h, wo = 2 * w;
  const size_t total = static_cast<size_t>(N) * ho * wo * nu4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % nu4);
    size_t t = idx / nu4;
    const int ox = static_cast<int>(t % wo);
    t /= wo;
    const int oy = static_cast<int>(t % ho);
    const size_t img = t / ho;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ((oy + 1) & 1) + 2 * a;
      const int iy = (oy + 1 - ky) / 2;
      if (oy + 1 - ky < 0 || iy >= h) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kx = ((ox + 1) & 1) + 2 * b;
        const int ix = (ox + 1 - kx) / 2;
        if (ox + 1 - kx < 0 || ix >= w) continue;
        const float4 gv = __ldg(g + (ky * 4 + kx) * nu4 + u);
        const float4 sv = __ldg(s5 + ((img * h + iy) * w + ix) * nu4 + u);
        acc.x = fmaf(gv.x, sv.x, acc.x);
        acc.y = fmaf(gv.y, sv.y, acc.y);
        acc.z = fmaf(gv.z, sv.z, acc.z);
        acc.w = fmaf(gv.w, sv.w, acc.w);
      }
    }
    const float4 base = __ldg(s4 + idx);
    fused[idx] = make_float4(base.x + fmaxf(acc.x, 0.f), base.y + fmaxf(acc.y, 0.f),
                             base.z + fmaxf(acc.z, 0.f), base.w + fmaxf(acc.w, 0.f));
  }
}

// low[p,c] = sum_u fused[p,u] * w[u,c]   (the 1x1 `score` conv applied BEFORE the x8 upsampling;
// valid because the upsampling kernel is shared by all channels and ReLU is the identity on
// its non-negative output - see DESIGN.md "decoder reordering").
__global__ void score_lowres_kernel(const float* __restrict__ fused, const float* __restrict__ w,
                                    float* __restrict__ low, size_t npix, int nu, int C) {
  const size_t total = npix * C;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const size_t p = idx / C;
    const float* f = fused + p * nu;
    float acc = 0.f;
    for (int u = 0; u < nu; ++u) acc = fmaf(__ldg(f + u), __ldg(w + u * C + c), acc);
    low[idx] = acc;
  }
}

// nu % 4 == 0: 4 threads per pixel, each reduces a quarter of the units with float4 loads and
// the [nu,C] matrix in shared memory, quad shuffle-reduction at the end.
template <int C>
__global__ void __launch_bounds__(256)
score_lowres_v4_kernel(const float4* __restrict__ fused, const float* __restrict__ w,
                       float* __restrict__ low, size_t npix, int nu) {
  extern __shared__ float s_w[];
  for (int i = threadIdx.x; i < nu * C; i += 256) s_w[i] = w[i];
  __syncthreads();
  const int nu4 = nu >> 2;
  const int part = threadIdx.x & 3;
  const size_t quads = static_cast<size_t>(gridDim.x) * 64;
  // uniform trip count per warp so the full-mask shuffles below are legal
  const size_t iters = (npix + quads - 1) / quads;
  for (size_t it = 0; it < iters; ++it) {
    const size_t p = it * quads + blockIdx.x * 64 + (threadIdx.x >> 2);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    if (p < npix) {
      const float4* f = fused + p * nu4;
      for (int u4 = part; u4 < nu4; u4 += 4) {
        const float4 v = __ldg(f + u4);
        const float* wr = s_w + u4 * 4 * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          acc[c] = fmaf(v.x, wr[c], acc[c]);
          acc[c] = fmaf(v.y, wr[C + c], acc[c]);
          acc[c] = fmaf(v.z, wr[2 * C + c], acc[c]);
          acc[c] = fmaf(v.w, wr[3 * C + c], acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
    }
    if (p < npix && part == 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) low[p * C + c] = acc[c];
    }
  }
}

// Decoder tail: 16x16 stride-8 upsampling of the low-res class scores + bias + softmax +
// argmax in one pass; block = 16x16 output pixels, the <= 4x4 contributing low-res pixels
// are staged in shared memory.  T > 1: loop over MC samples and accumulate moments.
// SOFTMAX == false: labels (and raw scores) only - the argmax of the softmax is derived from the
// scores (argmax.cuh), the exponentials are evaluated only on near-ties.
template <int C, bool MC, bool SOFTMAX = true>
__global__ void __launch_bounds__(256)
decode_upsample8_kernel(const float* __restrict__ low, const float* __restrict__ g,
                        const float* __restrict__ bias, int T, int N, int h, int w,
                        DecodeOut out, float* __restrict__ mean_prob,
                        float* __restrict__ var_prob, float* __restrict__ mean_var) {
  __shared__ float s_low[16 * C];
  __shared__ float s_g[256];
  const int H = 8 * h, W = 8 * w;
  const int bx = blockIdx.x * 16, by = blockIdx.y * 16;
  const int img = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ox = bx + tx, oy = by + ty;
  const bool valid = ox < W && oy < H;
  s_g[threadIdx.x] = g[threadIdx.x];
  // 16 output rows/cols touch low-res rows/cols iy0..iy0+3 with iy0 = by/8 - 1
  const int iy0 = (by + 4) / 8 - 1, ix0 = (bx + 4) / 8 - 1;
  const int ay = (oy + 4) >> 3, ry = (oy + 4) & 7;   // taps: (iy=ay, ky=ry), (iy=ay-1, ky=ry+8)
  const int ax = (ox + 4) >> 3, rx = (ox + 4) & 7;

  float mean[C], m2[C];
  if (MC) {
#pragma unroll
    for (int c = 0; c < C; ++c) mean[c] = m2[c] = 0.f;
  }
  for (int t = 0; t < T; ++t) {
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * C; i += 256) {
      const int c = i % C, cell = i / C;
      const int iy = iy0 + cell / 4, ix = ix0 + cell % 4;
      float v = 0.f;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w)
        v = __ldg(low + (((static_cast<size_t>(t) * N + img) * h + iy) * w + ix) * C + c);
      s_low[i] = v;
    }
    __syncthreads();
    if (!valid) continue;
    float s[C];
#pragma unroll
    for (int c = 0; c < C; ++c) s[c] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int iy = ay - a, ky = ry + 8 * a;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ix = ax - b, kx = rx + 8 * b;
        if (ix < 0 || ix >= w) continue;
        const float wgt = s_g[ky * 16 + kx];
        const float* lp = s_low + ((iy - iy0) * 4 + (ix - ix0)) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) s[c] = fmaf(wgt, lp[c], s[c]);
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      s[c] += __ldg(bias + c);
      mx = fmaxf(mx, s[c]);
    }
    const size_t pix = (static_cast<size_t>(img) * H + oy) * W + ox;
    if (!MC && out.score) {
#pragma unroll
      for (int c = 0; c < C; ++c) out.score[pix * C + c] = s[c];
    }
    if (!MC && !SOFTMAX) {
      const int best = argmax_of_softmax<C>(s);
      if (out.label_u8) out.label_u8[pix] = static_cast<uint8_t>(best);
      if (out.label_i64) out.label_i64[pix] = best;
      continue;
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      s[c] = expf(s[c] - mx);
      sum += s[c];
    }
    if (MC) {
      // Welford update of the per-class mean / sum of squared deviations
      const float inv_n = 1.f / static_cast<float>(t + 1);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float pr = s[c] / sum;
        const float d = pr - mean[c];
        mean[c] += d * inv_n;
        m2[c] = fmaf(d, pr - mean[c], m2[c]);
      }
    } else {
      int best = 0;
      float bestv = -1.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float pr = s[c] / sum;
        if (out.prob) out.prob[pix * C + c] = pr;
        if (pr > bestv) {
          bestv = pr;
          best = c;
        }
      }
      if (out.label_u8) out.label_u8[pix] = static_cast<uint8_t>(best);
      if (out.label_i64) out.label_i64[pix] = best;
    }
  }
  if (MC && valid) {
    const size_t pix = (static_cast<size_t>(img) * H + oy) * W + ox;
    float mv = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = m2[c] / static_cast<float>(T);
      if (mean_prob) mean_prob[pix * C + c] = mean[c];
      if (var_prob) var_prob[pix * C + c] = v;
      mv += v;
    }
    if (mean_var) mean_var[pix] = mv / static_cast<float>(C);
  }
}

// MC-dropout decode: mean / population variance over the T per-sample softmax outputs of a pixel
// (variance_mix.py:62-66) without materialising the samples.  The low-resolution cells of
// kTB samples are staged per barrier pair (the generic kernel above synchronises twice per
// sample, which made this pass latency-bound: 1.3 ms per modality for 8 frames x 20 samples), the
// softmax uses ex2.approx and one reciprocal per sample - the moments are statistics over random
// masks, their consumers compare them at 1e-4 - and the Welford update stays in registers.
template <int C>
__global__ void __launch_bounds__(256)
decode_upsample8_mc_kernel(const float* __restrict__ low, const float* __restrict__ g,
                           const float* __restrict__ bias, int T, int N, int h, int w,
                           float* __restrict__ mean_prob, float* __restrict__ var_prob,
                           float* __restrict__ mean_var) {
  constexpr int kTB = 8;
  __shared__ float s_low[kTB][16 * C];
  __shared__ float s_g[256];
  __shared__ float s_b[C];
  const int H = 8 * h, W = 8 * w;
  const int bx = blockIdx.x * 16, by = blockIdx.y * 16;
  const int img = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ox = bx + tx, oy = by + ty;
  const bool valid = ox < W && oy < H;
  s_g[threadIdx.x] = g[threadIdx.x];
  if (threadIdx.x < C) s_b[threadIdx.x] = bias[threadIdx.x];
  const int iy0 = (by + 4) / 8 - 1, ix0 = (bx + 4) / 8 - 1;
  const int ay = (oy + 4) >> 3, ry = (oy + 4) & 7;
  const int ax = (ox + 4) >> 3, rx = (ox + 4) & 7;
  float wgt[4];
  int cell[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int iy = ay - a, ix = ax - b;
      const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
      wgt[a * 2 + b] = 0.f;
      cell[a * 2 + b] = in ? ((iy - iy0) * 4 + (ix - ix0)) * C : 0;
      if (in) cell[a * 2 + b] |= 0x40000000;      // mark: weight read after the barrier
    }
  float mean[C], m2[C];
#pragma unroll
  for (int c = 0; c < C; ++c) mean[c] = m2[c] = 0.f;
  for (int t0 = 0; t0 < T; t0 += kTB) {
    const int tb = T - t0 < kTB ? T - t0 : kTB;
    __syncthreads();
    for (int i = threadIdx.x; i < tb * 16 * C; i += 256) {
      const int tt = i / (16 * C), rem = i - tt * 16 * C;
      const int c = rem % C, cl = rem / C;
      const int iy = iy0 + cl / 4, ix = ix0 + cl % 4;
      float v = 0.f;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w)
        v = __ldg(low + (((static_cast<size_t>(t0 + tt) * N + img) * h + iy) * w + ix) * C + c);
      s_low[tt][rem] = v;
    }
    __syncthreads();
    if (t0 == 0) {
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
          if (cell[a * 2 + b] & 0x40000000) {
            wgt[a * 2 + b] = s_g[(ry + 8 * a) * 16 + rx + 8 * b];
            cell[a * 2 + b] &= 0x3fffffff;
          }
    }
    if (!valid) continue;
    for (int tt = 0; tt < tb; ++tt) {
      float s[C];
#pragma unroll
      for (int c = 0; c < C; ++c) s[c] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* lp = s_low[tt] + cell[k];
#pragma unroll
        for (int c = 0; c < C; ++c) s[c] = fmaf(wgt[k], lp[c], s[c]);
      }
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        s[c] += s_b[c];
        mx = fmaxf(mx, s[c]);
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        s[c] = __expf(s[c] - mx);
        sum += s[c];
      }
      const float inv = 1.f / sum;
      const float inv_n = 1.f / static_cast<float>(t0 + tt + 1);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float pr = s[c] * inv;
        const float d = pr - mean[c];
        mean[c] = fmaf(d, inv_n, mean[c]);
        m2[c] = fmaf(d, pr - mean[c], m2[c]);
      }
    }
  }
  if (valid) {
    const size_t pix = (static_cast<size_t>(img) * H + oy) * W + ox;
    const float inv_t = 1.f / static_cast<float>(T);
    float mv = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      // explicit roundings: the result must not depend on which outputs were asked for (the
      // compiler otherwise contracts v into an fma only in the variant that does not store it)
      const float v = __fmul_rn(m2[c], inv_t);
      if (mean_prob) mean_prob[pix * C + c] = mean[c];
      if (var_prob) var_prob[pix * C + c] = v;
      mv = __fadd_rn(mv, v);
    }
    if (mean_var) mean_var[pix] = mv / static_cast<float>(C);
  }
}

// the C class scores of one low-resolution cell: 16-byte loads where the row is 16-byte aligned
// (C % 4 == 0; the map itself comes from an arena slot aligned to 1 KB)
template <int C>
__device__ __forceinline__ void load_cell(const float* __restrict__ lp, float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(lp) + q);
      v[4 * q] = t.x;
      v[4 * q + 1] = t.y;
      v[4 * q + 2] = t.z;
      v[4 * q + 3] = t.w;
    }
  } else if constexpr (C % 2 == 0) {
#pragma unroll
    for (int q = 0; q < C / 2; ++q) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(lp) + q);
      v[2 * q] = t.x;
      v[2 * q + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = __ldg(lp + c);
  }
}

// Label-only decode (what score() asks for).  One thread = 4 horizontally adjacent output pixels:
// they share their four low-resolution cells ((ox + 4) >> 3 is constant over an aligned group of
// 4), so the cells are read once per group - the one-pixel-per-thread version was bound by L1
// bandwidth.  Same tap order and fmaf chain as decode_upsample8_kernel: identical scores / argmax.
template <int C>
__global__ void __launch_bounds__(256)
decode_upsample8_labels_kernel(const float* __restrict__ low, const float* __restrict__ g,
                               const float* __restrict__ bias, int h, int w,
                               uint8_t* __restrict__ label_u8, int64_t* __restrict__ label_i64) {
  const int H = 8 * h, W = 8 * w;
  const int ox0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
  const int oy = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int img = blockIdx.z;
  if (ox0 >= W || oy >= H) return;
  const int ay = (oy + 4) >> 3, ry = (oy + 4) & 7;   // taps: (iy=ay, ky=ry), (iy=ay-1, ky=ry+8)
  const int ax = (ox0 + 4) >> 3, rx0 = (ox0 + 4) & 7;   // rx0 is 0 or 4
  float s[4][C];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < C; ++c) s[i][c] = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int iy = ay - a, ky = ry + 8 * a;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ix = ax - b;
      if (ix < 0 || ix >= w) continue;
      const float4 wg = __ldg(reinterpret_cast<const float4*>(g + ky * 16 + rx0 + 8 * b));
      const float wgt[4] = {wg.x, wg.y, wg.z, wg.w};
      const float* lp = low + ((static_cast<size_t>(img) * h + iy) * w + ix) * C;
      float cellv[C];
      load_cell<C>(lp, cellv);
#pragma unroll
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i][c] = fmaf(wgt[i], cellv[c], s[i][c]);
      }
    }
  }
  uint32_t packed = 0;
  const size_t pix = (static_cast<size_t>(img) * H + oy) * W + ox0;
  float b[C];
#pragma unroll
  for (int c = 0; c < C; ++c) b[c] = __ldg(bias + c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int c = 0; c < C; ++c) s[i][c] += b[c];
    const int best = argmax_of_softmax<C>(s[i]);
    packed |= static_cast<uint32_t>(best) << (8 * i);
    if (label_i64) label_i64[pix + i] = best;
  }
  if (label_u8) *reinterpret_cast<uint32_t*>(label_u8 + pix) = packed;
}

// Decoder tail with batch norm (decoder(batchnorm=True), simple_fcn.py:115-133, as fusion_fcn.py:39
// builds it): upscore = relu(BN(bilinear x8 upsampling)) per feature channel, score = BN(1x1 conv),
// softmax, argmax - in one pass from the 1/8-resolution features.  The ReLU sits between the
// upsampling and the 1x1 conv, so the conv cannot move below the upsampling as in the plain
// decoder; the upsampled num_units-channel tensor still never exists in memory: a block handles
// 16 x 16 output pixels from the <= 4 x 4 feature cells it touches.
template <int C>
__global__ void __launch_bounds__(256)
decode_bn_upsample8_kernel(const float* __restrict__ feat, const float* __restrict__ g,
                           const float* __restrict__ up_scale, const float* __restrict__ up_shift,
                           const float* __restrict__ w, const float* __restrict__ bias,
                           const float* __restrict__ sc_scale, const float* __restrict__ sc_shift,
                           int h, int wd, int nu, DecodeOut out) {
  extern __shared__ float s_dyn_f[];
  float* s_feat = s_dyn_f;                 // [16 cells][nu]
  float* s_w = s_feat + 16 * nu;           // [nu][C]
  float* s_aff = s_w + nu * C;             // [2][nu] scale, shift of the upscore batch norm
  __shared__ float s_g[256];
  const int H = 8 * h, W = 8 * wd;
  const int bx = blockIdx.x * 16, by = blockIdx.y * 16;
  const int img = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ox = bx + tx, oy = by + ty;
  s_g[threadIdx.x] = g[threadIdx.x];
  const int iy0 = (by + 4) / 8 - 1, ix0 = (bx + 4) / 8 - 1;
  for (int i = threadIdx.x; i < 16 * nu; i += 256) {
    const int u = i % nu, cell = i / nu;
    const int iy = iy0 + cell / 4, ix = ix0 + cell % 4;
    float v = 0.f;
    if (iy >= 0 && iy < h && ix >= 0 && ix < wd)
      v = __ldg(feat + ((static_cast<size_t>(img) * h + iy) * wd + ix) * nu + u);
    s_feat[i] = v;
  }
  for (int i = threadIdx.x; i < nu * C; i += 256) s_w[i] = w[i];
  for (int i = threadIdx.x; i < nu; i += 256) {
    s_aff[i] = up_scale[i];
    s_aff[nu + i] = up_shift[i];
  }
  __syncthreads();
  if (ox >= W || oy >= H) return;
  const int ay = (oy + 4) >> 3, ry = (oy + 4) & 7;
  const int ax = (ox + 4) >> 3, rx = (ox + 4) & 7;
  float wgt[4];
  int cell[4];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int iy = ay - a, ix = ax - b;
      const bool in = iy >= 0 && iy < h && ix >= 0 && ix < wd;
      wgt[a * 2 + b] = in ? s_g[(ry + 8 * a) * 16 + rx + 8 * b] : 0.f;
      cell[a * 2 + b] = in ? ((iy - iy0) * 4 + (ix - ix0)) * nu : 0;
    }
  }
  float sc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) sc[c] = 0.f;
  for (int u = 0; u < nu; ++u) {
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) v = fmaf(wgt[t], s_feat[cell[t] + u], v);
    v = fmaxf(fmaf(v, s_aff[u], s_aff[nu + u]), 0.f);
    const float* wr = s_w + u * C;
#pragma unroll
    for (int c = 0; c < C; ++c) sc[c] = fmaf(v, wr[c], sc[c]);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    sc[c] = fmaf(sc[c] + __ldg(bias + c), __ldg(sc_scale + c), __ldg(sc_shift + c));
    mx = fmaxf(mx, sc[c]);
  }
  const size_t pix = (static_cast<size_t>(img) * H + oy) * W + ox;
  if (out.score) {
#pragma unroll
    for (int c = 0; c < C; ++c) out.score[pix * C + c] = sc[c];
  }
  if (!out.prob) {
    const int best = argmax_of_softmax<C>(sc);
    if (out.label_u8) out.label_u8[pix] = static_cast<uint8_t>(best);
    if (out.label_i64) out.label_i64[pix] = best;
    return;
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    sc[c] = expf(sc[c] - mx);
    sum += sc[c];
  }
  int best = 0;
  float bestv = -1.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float pr = sc[c] / sum;
    out.prob[pix * C + c] = pr;
    if (pr > bestv) {
      bestv = pr;
      best = c;
    }
  }
  if (out.label_u8) out.label_u8[pix] = static_cast<uint8_t>(best);
  if (out.label_i64) out.label_i64[pix] = best;
}

template <int C>
int decode_bn_dispatch(const float* feat, const float* g, const float* up_scale,
                       const float* up_shift, const float* w, const float* bias,
                       const float* sc_scale, const float* sc_shift, int N, int h, int wd, int nu,
                       const DecodeOut& out, cudaStream_t s) {
  dim3 grid(div_up(8 * wd, 16), div_up(8 * h, 16), N);
  const size_t smem = (16 * nu + nu * C + 2 * nu) * sizeof(float);
  decode_bn_upsample8_kernel<C><<<grid, 256, smem, s>>>(feat, g, up_scale, up_shift, w, bias,
                                                        sc_scale, sc_shift, h, wd, nu, out);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// Fused tail of BayesFusion.score(): label-only decode of up to 4 experts (same arithmetic as
// decode_upsample8_labels_kernel, hence the same labels) -> decision-table lookup
// (bayes_mix.py:61-112) -> confusion-matrix accumulation (base_model.py:140-151), one pass, no
// label map in HBM unless `fused_out` asks for it.  One thread = 4 horizontally adjacent pixels.
struct DecodeSrc {
  const float* low[4];      // [N,h,w,C] low-resolution class scores of each expert
  const float* bias[4];     // [C]
  const float* g[4];        // [16,16] shared bilinear kernel
};
template <int C>
__global__ void __launch_bounds__(256, 3)
decode_bayes_confusion_kernel(DecodeSrc src, int M, const int32_t* __restrict__ lut, int lut_size,
                              int N, int h, int w, const int32_t* __restrict__ gt,
                              unsigned long long* __restrict__ cm,
                              uint8_t* __restrict__ fused_out) {
  extern __shared__ int32_t s_dyn[];
  int32_t* s_lut = s_dyn;
  unsigned int* s_cm = reinterpret_cast<unsigned int*>(s_dyn + lut_size);
  for (int i = threadIdx.x; i < lut_size; i += 256) s_lut[i] = lut[i];
  for (int i = threadIdx.x; i < C * C; i += 256) s_cm[i] = 0u;
  __syncthreads();
  const int H = 8 * h, W = 8 * w;
  const int lane = threadIdx.x & 31;
  // few, fat blocks (each ends with C*C global atomics on the same addresses): a block walks
  // over many 256 x 4 pixel tiles and keeps its histogram in shared memory
  const int tiles_x = (W + 255) / 256, tiles_y = (H + 3) / 4;
  const int total = tiles_x * tiles_y * N;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
  const int bxi = tile % tiles_x;
  const int rest = tile / tiles_x;
  const int byi = rest % tiles_y;
  const int img = rest / tiles_y;
  const int ox0 = (bxi * 64 + (threadIdx.x & 63)) * 4;
  const int oy = byi * 4 + (threadIdx.x >> 6);
  const bool live = ox0 < W && oy < H;
  int key[4] = {-1, -1, -1, -1};
  if (live) {
    const int ay = (oy + 4) >> 3, ry = (oy + 4) & 7;
    const int ax = (ox0 + 4) >> 3, rx0 = (ox0 + 4) & 7;
    int idx[4] = {0, 0, 0, 0};
    for (int m = 0; m < M; ++m) {
      const float* __restrict__ low = src.low[m];
      const float* __restrict__ g = src.g[m];
      const float* __restrict__ bias = src.bias[m];
      float s[4][C];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < C; ++c) s[i][c] = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int iy = ay - a, ky = ry + 8 * a;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int ix = ax - b;
          if (ix < 0 || ix >= w) continue;
          const float4 wg = __ldg(reinterpret_cast<const float4*>(g + ky * 16 + rx0 + 8 * b));
          const float wgt[4] = {wg.x, wg.y, wg.z, wg.w};
          const float* lp = low + ((static_cast<size_t>(img) * h + iy) * w + ix) * C;
          float cellv[C];
          load_cell<C>(lp, cellv);
#pragma unroll
          for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s[i][c] = fmaf(wgt[i], cellv[c], s[i][c]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int c = 0; c < C; ++c) s[i][c] += __ldg(bias + c);
        idx[i] = idx[i] * C + argmax_of_softmax<C>(s[i]);
      }
    }
    const size_t pix = (static_cast<size_t>(img) * H + oy) * W + ox0;
    const int4 lv = __ldg(reinterpret_cast<const int4*>(gt + pix));
    const int l[4] = {lv.x, lv.y, lv.z, lv.w};
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int fused = s_lut[idx[i]];
      packed |= static_cast<uint32_t>(fused & 0xff) << (8 * i);
      if (l[i] >= 0 && l[i] < C) key[i] = l[i] * C + fused;
    }
    if (fused_out) *reinterpret_cast<uint32_t*>(fused_out + pix) = packed;
  }
  // warp-aggregated shared-memory histogram (same scheme as confusion_kernel in fusion.cu)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int key0 = __shfl_sync(0xffffffffu, key[j], 0);
    if (__all_sync(0xffffffffu, key[j] == key0)) {
      if (lane == 0 && key0 >= 0) atomicAdd(&s_cm[key0], 32u);
    } else {
      const unsigned peers = __match_any_sync(0xffffffffu, key[j]);
      if (key[j] >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&s_cm[key[j]], __popc(peers));
    }
  }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += 256)
    if (s_cm[i]) atomicAdd(cm + i, static_cast<unsigned long long>(s_cm[i]));
}

template <int C>
int decode_bayes_dispatch(const DecodeSrc& src, int M, const int32_t* lut, int lut_size, int N,
                          int h, int w, const int32_t* gt, long long* cm, uint8_t* fused_out,
                          cudaStream_t s) {
  const long long tiles = static_cast<long long>(div_up(8 * w, 256)) * div_up(8 * h, 4) * N;
  const long long cap = static_cast<long long>(device_info().num_sms) * 3;   // resident blocks
  const int grid = static_cast<int>(tiles < cap ? tiles : cap);
  decode_bayes_confusion_kernel<C><<<grid, 256, (lut_size + C * C) * sizeof(int32_t), s>>>(
      src, M, lut, lut_size, N, h, w, gt, reinterpret_cast<unsigned long long*>(cm), fused_out);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

template <int C>
int decode_dispatch(bool mc, const float* low, const float* g, const float* bias, int T, int N,
                    int h, int w, const DecodeOut& out, float* mean_prob, float* var_prob,
                    float* mean_var, cudaStream_t s) {
  dim3 grid(div_up(8 * w, 16), div_up(8 * h, 16), N);
  if (!mc && !out.prob && !out.score) {
    dim3 lgrid(div_up(8 * w, 256), div_up(8 * h, 4), N);
    decode_upsample8_labels_kernel<C><<<lgrid, 256, 0, s>>>(low, g, bias, h, w, out.label_u8,
                                                            out.label_i64);
  } else if (mc && (debug_flags() & 512))
    decode_upsample8_kernel<C, true><<<grid, 256, 0, s>>>(low, g, bias, T, N, h, w, out,
                                                           mean_prob, var_prob, mean_var);
  else if (mc)
    decode_upsample8_mc_kernel<C><<<grid, 256, 0, s>>>(low, g, bias, T, N, h, w, mean_prob,
                                                       var_prob, mean_var);
  else if (out.prob)
    decode_upsample8_kernel<C, false><<<grid, 256, 0, s>>>(low, g, bias, 1, N, h, w, out,
                                                            nullptr, nullptr, nullptr);
  else
    decode_upsample8_kernel<C, false, false><<<grid, 256, 0, s>>>(low, g, bias, 1, N, h, w, out,
                                                                   nullptr, nullptr, nullptr);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

#define XV_DISPATCH_C(C, CALL)                                                          \
  switch (C) {                                                                          \
    case 2: { constexpr int kC = 2; return CALL; }                                      \
    case 3: { constexpr int kC = 3; return CALL; }                                      \
    case 4: { constexpr int kC = 4; return CALL; }                                      \
    case 5: { constexpr int kC = 5; return CALL; }                                      \
    case 6: { constexpr int kC = 6; return CALL; }                                      \
    case 7: { constexpr int kC = 7; return CALL; }                                      \
    case 8: { constexpr int kC = 8; return CALL; }                                      \
    case 9: { constexpr int kC = 9; return CALL; }                                      \
    case 10: { constexpr int kC = 10; return CALL; }                                    \
    case 11: { constexpr int kC = 11; return CALL; }                                    \
    case 12: { constexpr int kC = 12; return CALL; }                                    \
    case 13: { constexpr int kC = 13; return CALL; }                                    \
    case 14: { constexpr int kC = 14; return CALL; }                                    \
    case 15: { constexpr int kC = 15; return CALL; }                                    \
    case 16: { constexpr int kC = 16; return CALL; }                                    \
    case 17: { constexpr int kC = 17; return CALL; }                                    \
    case 18: { constexpr int kC = 18; return CALL; }                                    \
    case 19: { constexpr int kC = 19; return CALL; }                                    \
    case 20: { constexpr int kC = 20; return CALL; }                                    \
    case 21: { constexpr int kC = 21; return CALL; }                                    \
    case 22: { constexpr int kC = 22; return CALL; }                                    \
    case 23: { constexpr int kC = 23; return CALL; }                                    \
    case 24: { constexpr int kC = 24; return CALL; }                                    \
    default: return fail("num_classes must be in [2, 24]");                             \
  }

}  // namespace

int launch_im2col_c1(const float* x, __nv_bfloat16* out, int N, int H, int W, int cin,
                     cudaStream_t s) {
  XV_CHECK(cin >= 1 && cin <= 3, "conv1_1 packing supports Cin <= 3");
  const size_t total = static_cast<size_t>(N) * H * W;
  const int grid = grid_for(total);
  if (cin == 1) im2col_c1_kernel<1><<<grid, kThreads, 0, s>>>(x, out, N, H, W);
  if (cin == 2) im2col_c1_kernel<2><<<grid, kThreads, 0, s>>>(x, out, N, H, W);
  if (cin == 3) im2col_c1_kernel<3><<<grid, kThreads, 0, s>>>(x, out, N, H, W);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_to_f32(const void* in, int src_dtype, float* out, size_t n, cudaStream_t s) {
  if (src_dtype == 0) {
    if (n % 16 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
      u8x16_to_f32_kernel<<<grid_for(n / 16), kThreads, 0, s>>>(
          reinterpret_cast<const uint4*>(in), reinterpret_cast<float4*>(out), n / 16);
    } else {
      to_f32_kernel<uint8_t><<<grid_for(n), kThreads, 0, s>>>(static_cast<const uint8_t*>(in), out, n);
    }
  } else if (src_dtype == 1) {
    to_f32_kernel<uint16_t><<<grid_for(n), kThreads, 0, s>>>(static_cast<const uint16_t*>(in), out, n);
  } else if (src_dtype == 2) {
    to_f32_kernel<int16_t><<<grid_for(n), kThreads, 0, s>>>(static_cast<const int16_t*>(in), out, n);
  } else if (src_dtype == 3) {
    to_f32_kernel<int32_t><<<grid_for(n), kThreads, 0, s>>>(static_cast<const int32_t*>(in), out, n);
  } else {
    return fail("to_f32: src_dtype must be 0 (uint8), 1 (uint16), 2 (int16) or 3 (int32)");
  }
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_f32_to_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t s) {
  f32_to_bf16_kernel<<<grid_for(n), kThreads, 0, s>>>(in, out, n);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_bf16_to_f32(const __nv_bfloat16* in, float* out, size_t n, cudaStream_t s) {
  bf16_to_f32_kernel<<<grid_for(n), kThreads, 0, s>>>(in, out, n);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_maxpool_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int N, int H, int W, int C,
                        cudaStream_t s) {
  XV_CHECK(C % 8 == 0, "maxpool_bf16: C must be a multiple of 8");
  const size_t total = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
  maxpool_bf16_kernel<<<grid_for(total), kThreads, 0, s>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), N, H, W, C / 8);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_maxpool_f32(const float* in, float* out, int N, int H, int W, int C, cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * (H / 2) * (W / 2) * C;
  maxpool_f32_kernel<<<grid_for(total), kThreads, 0, s>>>(in, out, N, H, W, C);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_dropout_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, size_t n, int replicate,
                        const DropoutSpec& d, cudaStream_t s) {
  // fast path: 16-byte accesses, 8 elements (two Philox groups) per thread; same masks as the
  // scalar kernel (the Philox counter is the index of a group of four elements)
  if (!(debug_flags() & 256) && n % 8 == 0 && d.pass_elems % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      (d.ext_mask == nullptr || (reinterpret_cast<uintptr_t>(d.ext_mask) & 7) == 0)) {
    const size_t vecs = n * replicate / 8;
    dropout_bf16x8_kernel<<<grid_for(vecs), kThreads, 0, s>>>(
        reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), n / 8, vecs, d.rate,
        d.ext_mask, d.seed, d.offset, d.pass_elems / 8);
    XV_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  const size_t groups = (n * replicate + 3) / 4 + 1;
  dropout_kernel<__nv_bfloat16><<<grid_for(groups), kThreads, 0, s>>>(
      in, out, n, replicate, d.rate, d.ext_mask, d.seed, d.offset, d.pass_elems);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_dropout_f32(const float* in, float* out, size_t n, int replicate,
                       const DropoutSpec& d, cudaStream_t s) {
  const size_t groups = (n * replicate + 3) / 4 + 1;
  dropout_kernel<float><<<grid_for(groups), kThreads, 0, s>>>(
      in, out, n, replicate, d.rate, d.ext_mask, d.seed, d.offset, d.pass_elems);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_conv_f32(const float* x, const float* w, const float* bias, float* out, int N, int H,
                    int W, int cin, int cout, int k, int relu, cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * H * W * ((cout + 7) / 8);
  conv_f32_kernel<<<grid_for(total, 128), 128, 0, s>>>(x, w, bias, out, N, H, W, cin, cout, k,
                                                       relu);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_deconv_f32(const float* x, const float* w, float* out, int N, int hin, int win, int cin,
                      int cout, int k, int stride, int relu, const float* addend,
                      cudaStream_t s) {
  XV_CHECK((k - stride) % 2 == 0 && k >= stride, "deconv: (k - stride) must be even");
  const size_t total = static_cast<size_t>(N) * hin * stride * win * stride * cout;
  deconv_f32_kernel<<<grid_for(total), kThreads, 0, s>>>(x, w, out, N, hin, win, cin, cout, k,
                                                         stride, relu, addend);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_concat_bf16(const __nv_bfloat16* src, __nv_bfloat16* dst, size_t npix, int c_src,
                       int c_dst, int offset, cudaStream_t s) {
  XV_CHECK(c_src % 8 == 0 && c_dst % 8 == 0 && offset % 8 == 0, "concat: channels must be multiples of 8");
  concat_bf16_kernel<<<grid_for(npix * (c_src / 8)), kThreads, 0, s>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), npix, c_src / 8, c_dst / 8,
      offset / 8);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_add_f32(const float* a, const float* b, float* out, size_t n, cudaStream_t s) {
  add_f32_kernel<<<grid_for(n), kThreads, 0, s>>>(a, b, out, n);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_affine_f32(float* x, const float* scale, const float* shift, size_t npix, int C,
                      int relu, cudaStream_t s) {
  affine_f32_kernel<<<grid_for(npix * C), kThreads, 0, s>>>(x, scale, shift, npix * C, C, relu);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_upscore2_add(const float* s5, const float* s4, const float* g, float* fused, int N,
                        int h, int w, int nu, cudaStream_t s, float* up5) {
  const size_t total = static_cast<size_t>(N) * 4 * h * w * nu;
  if (nu % 4 == 0 && up5 == nullptr)
    upscore2_add_v4_kernel<<<grid_for(total / 4), kThreads, 0, s>>>(
        reinterpret_cast<const float4*>(s5), reinterpret_cast<const float4*>(s4),
        reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(fused), N, h, w, nu / 4);
  else
    upscore2_add_kernel<<<grid_for(total), kThreads, 0, s>>>(s5, s4, g, fused, up5, N, h, w, nu);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
// Fused head tail (simple_fcn.py:82-85 + the 1x1 score conv moved below the upsampling, DESIGN.md
// 4.2): fused = s4 + relu(up2(s5)) and low = fused x W[nu, C] in one pass over the 1/8-resolution
// pixels.  Four threads share a pixel (nu / 4 channels each as float4 pieces); `fused` is written
// too (small, and fit() / the layer diagnostics read it).
template <int C>
__global__ void __launch_bounds__(256)
head_fused_kernel(const float4* __restrict__ s5, const float4* __restrict__ s4,
                  const float* __restrict__ g, const float* __restrict__ w,
                  float4* __restrict__ fused, float* __restrict__ low, int N, int h, int wd,
                  int nu) {
  extern __shared__ float s_hw[];
  float* s_w = s_hw;                       // [nu][C]
  float* s_g = s_hw + nu * C;              // [16][nu]
  for (int i = threadIdx.x; i < nu * C; i += 256) s_w[i] = w[i];
  for (int i = threadIdx.x; i < 16 * nu; i += 256) s_g[i] = g[i];
  __syncthreads();
  const int nu4 = nu >> 2;
  const int part = threadIdx.x & 3;
  const int ho = 2 * h, wo = 2 * wd;
  const size_t npix = static_cast<size_t>(N) * ho * wo;
  const size_t quads = static_cast<size_t>(gridDim.x) * 64;
  const size_t iters = (npix + quads - 1) / quads;       // uniform trip count per warp
  for (size_t it = 0; it < iters; ++it) {
    const size_t p = it * quads + blockIdx.x * 64 + (threadIdx.x >> 2);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    if (p < npix) {
      const int ox = static_cast<int>(p % wo);
      const size_t t = p / wo;
      const int oy = static_cast<int>(t % ho);
      const size_t img = t / ho;
      // oy = 2*iy - 1 + ky  ->  ky in {(oy+1)%2, (oy+1)%2 + 2}; same for x
      int ky[2], iy[2], kx[2], ix[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        ky[a] = ((oy + 1) & 1) + 2 * a;
        iy[a] = (oy + 1 - ky[a]) >= 0 ? (oy + 1 - ky[a]) / 2 : -1;
        if (iy[a] >= h) iy[a] = -1;
        kx[a] = ((ox + 1) & 1) + 2 * a;
        ix[a] = (ox + 1 - kx[a]) >= 0 ? (ox + 1 - kx[a]) / 2 : -1;
        if (ix[a] >= wd) ix[a] = -1;
      }
      for (int u4 = part; u4 < nu4; u4 += 4) {
        float up[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (iy[a] < 0) continue;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            if (ix[b] < 0) continue;
            const float4 v = __ldg(s5 + ((img * h + iy[a]) * wd + ix[b]) * nu4 + u4);
            const float* gr = s_g + (ky[a] * 4 + kx[b]) * nu + u4 * 4;
            up[0] = fmaf(gr[0], v.x, up[0]);
            up[1] = fmaf(gr[1], v.y, up[1]);
            up[2] = fmaf(gr[2], v.z, up[2]);
            up[3] = fmaf(gr[3], v.w, up[3]);
          }
        }
        const float4 a4 = __ldg(s4 + p * nu4 + u4);
        const float4 f = make_float4(a4.x + fmaxf(up[0], 0.f), a4.y + fmaxf(up[1], 0.f),
                                     a4.z + fmaxf(up[2], 0.f), a4.w + fmaxf(up[3], 0.f));
        fused[p * nu4 + u4] = f;
        const float* wr = s_w + u4 * 4 * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          acc[c] = fmaf(f.x, wr[c], acc[c]);
          acc[c] = fmaf(f.y, wr[C + c], acc[c]);
          acc[c] = fmaf(f.z, wr[2 * C + c], acc[c]);
          acc[c] = fmaf(f.w, wr[3 * C + c], acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
    }
    if (p < npix && part == 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) low[p * C + c] = acc[c];
    }
  }
}

template <int C>
int head_fused_dispatch(const float* s5, const float* s4, const float* g, const float* w,
                        float* fused, float* low, int N, int h, int wd, int nu, cudaStream_t s) {
  const size_t npix = static_cast<size_t>(N) * 4 * h * wd;
  const size_t blocks = (npix + 63) / 64;
  const size_t cap = static_cast<size_t>(device_info().num_sms) * 8;
  head_fused_kernel<C><<<static_cast<int>(blocks < cap ? (blocks ? blocks : 1) : cap), 256,
                         (nu * C + 16 * nu) * sizeof(float), s>>>(
      reinterpret_cast<const float4*>(s5), reinterpret_cast<const float4*>(s4), g, w,
      reinterpret_cast<float4*>(fused), low, N, h, wd, nu);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

template <int C>
int score_lowres_v4_dispatch(const float* fused, const float* w, float* low, size_t npix, int nu,
                             cudaStream_t s) {
  const size_t blocks = (npix + 63) / 64;
  const size_t cap = static_cast<size_t>(device_info().num_sms) * 8;
  score_lowres_v4_kernel<C><<<static_cast<int>(blocks < cap ? (blocks ? blocks : 1) : cap), 256,
                              nu * C * sizeof(float), s>>>(
      reinterpret_cast<const float4*>(fused), w, low, npix, nu);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

bool head_fused_supported(int nu, int C) {
  return nu % 4 == 0 && C >= 2 && C <= kMaxClasses && (nu * C + 16 * nu) * 4 <= 48 * 1024;
}
int launch_head_fused(const float* s5, const float* s4, const float* g_4x4xnu, const float* w_nuxc,
                      float* fused, float* low, int N, int h, int w, int nu, int C,
                      cudaStream_t s) {
  XV_CHECK(head_fused_supported(nu, C), "head_fused: unsupported num_units / num_classes");
  XV_DISPATCH_C(C, (head_fused_dispatch<kC>(s5, s4, g_4x4xnu, w_nuxc, fused, low, N, h, w, nu, s)));
  return 0;
}
int launch_score_lowres(const float* fused, const float* w, float* low, size_t npix, int nu, int C,
                        cudaStream_t s) {
  if (nu % 16 == 0 && nu * C * 4 <= 40 * 1024) {
    XV_DISPATCH_C(C, (score_lowres_v4_dispatch<kC>(fused, w, low, npix, nu, s)));
  }
  score_lowres_kernel<<<grid_for(npix * C), kThreads, 0, s>>>(fused, w, low, npix, nu, C);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_decode_upsample8(const float* low, const float* g, const float* bias, int N, int h,
                            int w, int C, const DecodeOut& out, cudaStream_t s) {
  XV_DISPATCH_C(C, (decode_dispatch<kC>(false, low, g, bias, 1, N, h, w, out, nullptr, nullptr,
                                        nullptr, s)));
}
bool decode_bn_supported(int nu, int C) {
  return C >= 2 && C <= kMaxClasses && (16 * nu + nu * C + 2 * nu) * 4 <= 48 * 1024;
}
int launch_decode_bn_upsample8(const float* feat, const float* g_16x16, const float* up_scale,
                               const float* up_shift, const float* w_nuxc, const float* bias,
                               const float* sc_scale, const float* sc_shift, int N, int h, int w,
                               int nu, int C, const DecodeOut& out, cudaStream_t s) {
  XV_CHECK(decode_bn_supported(nu, C), "decode_bn: unsupported num_units / num_classes");
  XV_DISPATCH_C(C, (decode_bn_dispatch<kC>(feat, g_16x16, up_scale, up_shift, w_nuxc, bias,
                                           sc_scale, sc_shift, N, h, w, nu, out, s)));
  return 0;
}
int launch_decode_bayes_confusion(const float* const* low, const float* const* g,
                                  const float* const* bias, int M, const int32_t* lut, int C,
                                  int N, int h, int w, const int32_t* gt, long long* cm,
                                  uint8_t* fused_out, cudaStream_t s) {
  XV_CHECK(M >= 1 && M <= 4, "decode_bayes_confusion: 1..4 experts");
  XV_CHECK((8 * w) % 4 == 0, "decode_bayes_confusion: width must be a multiple of 4");
  int lut_size = 1;
  for (int m = 0; m < M; ++m) lut_size *= C;
  XV_CHECK((lut_size + C * C) * 4 <= 48 * 1024, "decode_bayes_confusion: decision table too large");
  DecodeSrc src;
  for (int m = 0; m < 4; ++m) {
    src.low[m] = m < M ? low[m] : nullptr;
    src.g[m] = m < M ? g[m] : nullptr;
    src.bias[m] = m < M ? bias[m] : nullptr;
  }
  XV_DISPATCH_C(C, (decode_bayes_dispatch<kC>(src, M, lut, lut_size, N, h, w, gt, cm, fused_out, s)));
  return 0;
}
int launch_decode_upsample8_mc(const float* low, const float* g, const float* bias, int T, int N,
                               int h, int w, int C, float* mean_prob, float* var_prob,
                               float* mean_var, cudaStream_t s) {
  DecodeOut none;
  XV_DISPATCH_C(C, (decode_dispatch<kC>(true, low, g, bias, T, N, h, w, none, mean_prob,
                                        var_prob, mean_var, s)));
}

}  // namespace xv
