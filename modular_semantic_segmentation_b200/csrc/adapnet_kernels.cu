// Memory-bound helper kernels of the Adapnet expert (xview/models/adapnet.py:99-173): everything
// between its tensor-core convolutions.  All of them move each byte once with 16-byte accesses.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kThreads = 256;

inline int grid_for(size_t work, int threads = kThreads) {
  const size_t blocks = (work + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(device_info().num_sms > 0 ? device_info().num_sms : 148) * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

__device__ __forceinline__ uint32_t add_relu_bf16x2(uint32_t a, uint32_t b) {
  const __nv_bfloat162 va = *reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162 vb = *reinterpret_cast<const __nv_bfloat162*>(&b);
  const float2 fa = __bfloat1622float2(va), fb = __bfloat1622float2(vb);
  return pack_bf16x2(fmaxf(fa.x + fb.x, 0.f), fmaxf(fa.y + fb.y, 0.f));
}

// block output of adapnet.py:49,96: relu(stage_3 + shortcut)
__global__ void add_relu_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                     uint4* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 x = __ldg(a + i), y = __ldg(b + i);
    out[i] = make_uint4(add_relu_bf16x2(x.x, y.x), add_relu_bf16x2(x.y, y.y),
                        add_relu_bf16x2(x.z, y.z), add_relu_bf16x2(x.w, y.w));
  }
}

__global__ void add_relu_f32_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                    float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = fmaxf(a[i] + b[i], 0.f);
}

__device__ __forceinline__ float4 col_load4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 col_load4(const __nv_bfloat16* p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Second half of a transposed convolution computed as a GEMM (conv2d_transpose 'same',
// custom_layers.py:71-121): `col` holds, per INPUT pixel, the k*k*cout products
// col[b,iy,ix,(ky*k+kx)*cout+co] = sum_ci x[b,iy,ix,ci] * w[ky,kx,co,ci] (* BN scale); output pixel
// (oy,ox) sums the (k/stride)^2 entries with oy = iy*stride - pad + ky, adds the BN shift and an
// optional addend (the skip connection, adapnet.py:162).  Outputs: fp32 [.., cout] and / or a
// zero-padded bf16 copy [.., pad_c] that feeds the next tensor-core GEMM.
// grid = (x blocks, output row, image); one thread = 4 channels of one output pixel.
template <typename ColT>
__global__ void col2im_kernel(const ColT* __restrict__ col, const float* __restrict__ shift,
                              const float* __restrict__ addend, float* __restrict__ out_f32,
                              __nv_bfloat16* __restrict__ out_bf16, int pad_c, int hin, int win,
                              int cout, int k, int stride) {
  const int ho = hin * stride, wo = win * stride, pad = (k - stride) / 2;
  const int cw4 = (out_bf16 ? pad_c : cout) / 4;         // channel quads enumerated per pixel
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wo * cw4) return;
  const int ox = idx / cw4;
  const int co = (idx - ox * cw4) * 4;
  const int oy = blockIdx.y;
  const size_t b = blockIdx.z;
  const size_t pix = (b * ho + oy) * wo + ox;
  if (co >= cout) {                                      // zero padding of the bf16 copy
    *reinterpret_cast<uint2*>(out_bf16 + pix * pad_c + co) = make_uint2(0u, 0u);
    return;
  }
  const int kkc = k * k * cout;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ky = (oy + pad) % stride; ky < k; ky += stride) {
    const int ny = oy + pad - ky;
    const int iy = ny / stride;
    if (ny < 0 || iy >= hin) continue;
    for (int kx = (ox + pad) % stride; kx < k; kx += stride) {
      const int nx = ox + pad - kx;
      const int ix = nx / stride;
      if (nx < 0 || ix >= win) continue;
      const float4 v =
          col_load4(col + ((b * hin + iy) * win + ix) * kkc + (ky * k + kx) * cout + co);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
  }
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + co));
  acc.x += sh.x;
  acc.y += sh.y;
  acc.z += sh.z;
  acc.w += sh.w;
  if (addend) {
    const float4 ad = __ldg(reinterpret_cast<const float4*>(addend + pix * cout + co));
    acc.x += ad.x;
    acc.y += ad.y;
    acc.z += ad.z;
    acc.w += ad.w;
  }
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + pix * cout + co) = acc;
  if (out_bf16)
    *reinterpret_cast<uint2*>(out_bf16 + pix * pad_c + co) =
        make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
}

// scalar variant for channel counts that are not multiples of 4
template <typename ColT>
__global__ void col2im_scalar_kernel(const ColT* __restrict__ col, const float* __restrict__ shift,
                                     const float* __restrict__ addend, float* __restrict__ out_f32,
                                     __nv_bfloat16* __restrict__ out_bf16, int pad_c, int hin,
                                     int win, int cout, int k, int stride) {
  const int ho = hin * stride, wo = win * stride, pad = (k - stride) / 2;
  const int cw = out_bf16 ? pad_c : cout;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wo * cw) return;
  const int ox = idx / cw;
  const int co = idx - ox * cw;
  const int oy = blockIdx.y;
  const size_t b = blockIdx.z;
  const size_t pix = (b * ho + oy) * wo + ox;
  if (co >= cout) {
    out_bf16[pix * pad_c + co] = __float2bfloat16_rn(0.f);
    return;
  }
  const int kkc = k * k * cout;
  float acc = 0.f;
  for (int ky = (oy + pad) % stride; ky < k; ky += stride) {
    const int ny = oy + pad - ky;
    const int iy = ny / stride;
    if (ny < 0 || iy >= hin) continue;
    for (int kx = (ox + pad) % stride; kx < k; kx += stride) {
      const int nx = ox + pad - kx;
      const int ix = nx / stride;
      if (nx < 0 || ix >= win) continue;
      acc += static_cast<float>(col[((b * hin + iy) * win + ix) * kkc + (ky * k + kx) * cout + co]);
    }
  }
  acc += shift[co];
  if (addend) acc += addend[pix * cout + co];
  if (out_f32) out_f32[pix * cout + co] = acc;
  if (out_bf16) out_bf16[pix * pad_c + co] = __float2bfloat16_rn(acc);
}

// Last step of the 16/8 transposed convolution computed as a 3x3 "phase" convolution: `d2s` holds
// for every 1/8-resolution cell the 8x8 output pixels it owns, channel (py*8+px)*C + c (BN shift
// already added by the GEMM epilogue).  One thread = one output pixel: depth-to-space, then the
// softmax / argmax of basic_fusion_model.py:21-22 (first index wins ties, like tf.argmax).
// Threads enumerate (cell, phase) in memory order, so the bf16 reads are fully coalesced; a group of
// 8 threads writes 8 neighbouring output pixels.
constexpr int kMaxC = 24;
template <bool VEC4>
__global__ void d2s_softmax_argmax_kernel(const __nv_bfloat16* __restrict__ d2s, int B, int h8,
                                          int w8, int C, float* __restrict__ score,
                                          float* __restrict__ prob, int64_t* __restrict__ label_i64,
                                          uint8_t* __restrict__ label_u8) {
  const int W = w8 * 8;
  const unsigned total = static_cast<unsigned>(B) * h8 * w8 * 64;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned cell = i >> 6, phase = i & 63;
  const unsigned cx = cell % w8;
  const unsigned row = cell / w8;                      // b * h8 + cy
  const size_t pix = (static_cast<size_t>(row) * 8 + (phase >> 3)) * W + cx * 8 + (phase & 7);
  const __nv_bfloat16* src = d2s + static_cast<size_t>(i) * C;
  float v[kMaxC];
  if (VEC4) {
#pragma unroll
    for (int c4 = 0; c4 < kMaxC / 4; ++c4) {
      if (c4 * 4 < C) {
        const float4 q = col_load4(src + c4 * 4);
        v[c4 * 4] = q.x;
        v[c4 * 4 + 1] = q.y;
        v[c4 * 4 + 2] = q.z;
        v[c4 * 4 + 3] = q.w;
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) v[c] = __bfloat162float(src[c]);
  }
  float best = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) {
    if (c < C && v[c] > best) {
      best = v[c];
      arg = c;
    }
  }
  if (score) {
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) score[pix * C + c] = v[c];
  }
  if (prob) {
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) {
        v[c] = expf(v[c] - best);
        sum += v[c];
      }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) prob[pix * C + c] = v[c] * inv;
  }
  if (label_i64) label_i64[pix] = arg;
  if (label_u8) label_u8[pix] = static_cast<uint8_t>(arg);
}

// generic fp32 convolution with TF 'SAME' geometry for any stride / dilation (validation mode):
// out[oy,ox] = sum w[ky,kx] * x[oy*stride + ky*dil - pad_t, ox*stride + kx*dil - pad_l]
__global__ void conv_f32_ex_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ out, int N,
                                   int H, int W, int cin, int cout, int k, int stride, int dil,
                                   int pad_t, int pad_l, int relu) {
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const int cg = (cout + 7) / 8;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * cg;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % cg);
    const size_t pix = idx / cg;
    const int ox = static_cast<int>(pix % Wo);
    const int oy = static_cast<int>((pix / Wo) % Ho);
    const size_t img = pix / (static_cast<size_t>(Wo) * Ho);
    const int co0 = g * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int yy = oy * stride + ky * dil - pad_t;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int xx = ox * stride + kx * dil - pad_l;
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

}  // namespace

int launch_add_relu_bf16(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* out,
                         size_t n, cudaStream_t s) {
  XV_CHECK(n % 8 == 0, "add_relu_bf16: element count must be a multiple of 8");
  add_relu_bf16_kernel<<<grid_for(n / 8), kThreads, 0, s>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
      reinterpret_cast<uint4*>(out), n / 8);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_add_relu_f32(const float* a, const float* b, float* out, size_t n, cudaStream_t s) {
  add_relu_f32_kernel<<<grid_for(n), kThreads, 0, s>>>(a, b, out, n);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_col2im(const void* col, bool col_bf16, const float* shift, const float* addend,
                  float* out_f32, __nv_bfloat16* out_bf16, int pad_c, int B, int hin, int win,
                  int cout, int k, int stride, cudaStream_t s) {
  XV_CHECK((k - stride) % 2 == 0 && k >= stride && k % stride == 0,
           "col2im: needs k a multiple of stride and (k - stride) even");
  XV_CHECK(out_f32 || out_bf16, "col2im: no output requested");
  const int ho = hin * stride, wo = win * stride;
  const int cw = out_bf16 ? pad_c : cout;
  const bool vec = cout % 4 == 0 && cw % 4 == 0;
  const int per_row = wo * (vec ? cw / 4 : cw);
  const dim3 grid((per_row + kThreads - 1) / kThreads, ho, B);
#define XV_COL2IM(KERNEL, T)                                                                  \
  KERNEL<T><<<grid, kThreads, 0, s>>>(static_cast<const T*>(col), shift, addend, out_f32,     \
                                      out_bf16, pad_c, hin, win, cout, k, stride)
  if (vec && col_bf16) XV_COL2IM(col2im_kernel, __nv_bfloat16);
  if (vec && !col_bf16) XV_COL2IM(col2im_kernel, float);
  if (!vec && col_bf16) XV_COL2IM(col2im_scalar_kernel, __nv_bfloat16);
  if (!vec && !col_bf16) XV_COL2IM(col2im_scalar_kernel, float);
#undef XV_COL2IM
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_d2s_softmax_argmax(const __nv_bfloat16* d2s, int B, int h8, int w8, int C, float* score,
                              float* prob, int64_t* label_i64, uint8_t* label_u8,
                              cudaStream_t s) {
  XV_CHECK(C >= 2 && C <= kMaxC, "d2s_softmax_argmax: C must be in [2, 24]");
  const size_t total = static_cast<size_t>(B) * h8 * w8 * 64;
  XV_CHECK(total < (1ull << 31), "d2s_softmax_argmax: too many pixels for one launch");
  const int grid = static_cast<int>((total + kThreads - 1) / kThreads);
  if (C % 4 == 0) {
    d2s_softmax_argmax_kernel<true><<<grid, kThreads, 0, s>>>(d2s, B, h8, w8, C, score, prob,
                                                              label_i64, label_u8);
  } else {
    d2s_softmax_argmax_kernel<false><<<grid, kThreads, 0, s>>>(d2s, B, h8, w8, C, score, prob,
                                                               label_i64, label_u8);
  }
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_conv_f32_ex(const float* x, const float* w, const float* bias, float* out, int N, int H,
                       int W, int cin, int cout, int k, int stride, int dil, int relu,
                       cudaStream_t s) {
  // TF 'SAME': pad_total = max((ceil(in/s) - 1) s + (k - 1) d + 1 - in, 0), smaller half first
  auto pad_before = [&](int size) {
    const int o = (size + stride - 1) / stride;
    const int total = (o - 1) * stride + (k - 1) * dil + 1 - size;
    return total > 0 ? total / 2 : 0;
  };
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * ((cout + 7) / 8);
  conv_f32_ex_kernel<<<grid_for(total, 128), 128, 0, s>>>(x, w, bias, out, N, H, W, cin, cout, k,
                                                          stride, dil, pad_before(H),
                                                          pad_before(W), relu);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace xv
