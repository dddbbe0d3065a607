// Backward / optimizer kernels for SimpleFCN.fit() (xview/models/base_model.py:179-261 with the
// loss of simple_fcn.py:205-215 + utils.py:43-53 and tf.train.AdamOptimizer, base_model.py:153-162).
//
// The data gradients of the 3x3 / 1x1 convolutions reuse the forward tcgen05 implicit-GEMM
// kernels with flipped + transposed weights (conv_igemm*_sm100.cu); this file holds the rest:
// softmax-cross-entropy gradient, transposes of the two bilinear transposed convolutions,
// ReLU / max-pool backward, weight + bias gradients (fp32 accumulation on the CUDA cores, round 1)
// and the Adam update with re-packing of the bf16 operand copies.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kThreads = 256;

inline int grid_for(size_t work, int threads = kThreads, int per_sm = 16) {
  size_t g = (work + threads - 1) / threads;
  size_t cap = static_cast<size_t>(device_info().num_sms) * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------ loss
// score [npix,C] -> dscore in place = softmax(score) - onehot(label) for valid labels, 0 else.
// loss[0] += sum -log_softmax(score)[label], loss[1] += #valid, dbias[c] += sum dscore[.,c].
template <int C>
__global__ void __launch_bounds__(kThreads)
ce_grad_kernel(float* __restrict__ score, const int32_t* __restrict__ labels, int64_t npix,
               double* __restrict__ loss, float* __restrict__ dbias) {
  __shared__ float s_db[C];
  __shared__ float s_loss[2];
  if (threadIdx.x < C) s_db[threadIdx.x] = 0.f;
  if (threadIdx.x < 2) s_loss[threadIdx.x] = 0.f;
  __syncthreads();
  float my_loss = 0.f, my_cnt = 0.f;
  float db[C];
#pragma unroll
  for (int c = 0; c < C; ++c) db[c] = 0.f;
  for (int64_t pix = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; pix < npix;
       pix += static_cast<int64_t>(gridDim.x) * kThreads) {
    float* s = score + pix * C;
    const int label = __ldg(labels + pix);
    const bool valid = label >= 0 && label < C;
    float v[C];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      v[c] = s[c];
      mx = fmaxf(mx, v[c]);
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      v[c] = expf(v[c] - mx);
      sum += v[c];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float g = 0.f;
      if (valid) {
        g = v[c] * inv - (c == label ? 1.f : 0.f);
        if (c == label) my_loss -= logf(fmaxf(v[c] * inv, 1e-38f));
      }
      s[c] = g;
      db[c] += g;
    }
    if (valid) my_cnt += 1.f;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float t = db[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_db[c], t);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    my_loss += __shfl_xor_sync(0xffffffffu, my_loss, off);
    my_cnt += __shfl_xor_sync(0xffffffffu, my_cnt, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_loss[0], my_loss);
    atomicAdd(&s_loss[1], my_cnt);
  }
  __syncthreads();
  if (threadIdx.x < C) atomicAdd(dbias + threadIdx.x, s_db[threadIdx.x]);
  if (threadIdx.x < 2) atomicAdd(loss + threadIdx.x, static_cast<double>(s_loss[threadIdx.x]));
}

// Loss head of fit() in one pass over the 1/8-resolution class scores (bilinear-upscore fast path):
// per 16x16 output tile the <= 4x4 contributing low-resolution cells are staged, every thread
// decodes its pixel's scores exactly as decode_upsample8_kernel does, takes softmax - onehot (the
// gradient of simple_fcn.py:205-215 / utils.py:43-53 before the division by the number of
// labelled pixels), and the transposed upsampling
//     dlow[cell, c] += sum_{pixels of the tile} g[ky][kx] * dscore[pixel, c]
// is a [16 cells x 256 pixels] x [256 x C] product out of shared memory, added to dlow with one
// atomic per (cell, class) and tile.  Neither the full-resolution scores nor their gradient
// (2 x 226 MB at batch 16) exist in HBM; the three kernels this replaces (decode, ce_grad,
// upsample8_transpose) took 0.57 ms.  Persistent blocks: loss, count and the score-bias gradient
// are flushed once per block.  dlow must be zero on entry.
template <int C>
__global__ void __launch_bounds__(256)
loss_lowres_grad_kernel(const float* __restrict__ low, const float* __restrict__ g,
                        const float* __restrict__ bias, const int32_t* __restrict__ labels, int N,
                        int h, int w, float* __restrict__ dlow, double* __restrict__ loss,
                        float* __restrict__ dbias) {
  __shared__ float s_low[16 * C];
  __shared__ float s_g[256];
  __shared__ float s_b[C];
  __shared__ float s_d[256 * C];
  __shared__ float s_db[C];
  __shared__ float s_loss[2];
  const int H = 8 * h, W = 8 * w;
  const int tiles_x = (W + 15) / 16, tiles_y = (H + 15) / 16;
  const int total = tiles_x * tiles_y * N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  s_g[threadIdx.x] = g[threadIdx.x];
  if (threadIdx.x < C) {
    s_b[threadIdx.x] = bias[threadIdx.x];
    s_db[threadIdx.x] = 0.f;
  }
  if (threadIdx.x < 2) s_loss[threadIdx.x] = 0.f;
  float my_loss = 0.f, my_cnt = 0.f;
  float db[C];
#pragma unroll
  for (int c = 0; c < C; ++c) db[c] = 0.f;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int bx = (tile % tiles_x) * 16;
    const int by = ((tile / tiles_x) % tiles_y) * 16;
    const int img = tile / (tiles_x * tiles_y);
    const int ox = bx + tx, oy = by + ty;
    const bool inside = ox < W && oy < H;
    const int iy0 = (by + 4) / 8 - 1, ix0 = (bx + 4) / 8 - 1;
    const int ay = (oy + 4) >> 3, ry = (oy + 4) & 7;
    const int ax = (ox + 4) >> 3, rx = (ox + 4) & 7;
    __syncthreads();                                  // previous tile's s_low / s_d are consumed
    for (int i = threadIdx.x; i < 16 * C; i += 256) {
      const int c = i % C, cell = i / C;
      const int iy = iy0 + cell / 4, ix = ix0 + cell % 4;
      float v = 0.f;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w)
        v = __ldg(low + ((static_cast<size_t>(img) * h + iy) * w + ix) * C + c);
      s_low[i] = v;
    }
    __syncthreads();
    float d[C];
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = 0.f;
    if (inside) {
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
        s[c] += s_b[c];
        mx = fmaxf(mx, s[c]);
      }
      const int label = __ldg(labels + (static_cast<size_t>(img) * H + oy) * W + ox);
      if (label >= 0 && label < C) {
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          s[c] = expf(s[c] - mx);
          sum += s[c];
        }
        const float inv = 1.f / sum;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float pr = s[c] * inv;
          d[c] = pr - (c == label ? 1.f : 0.f);
          if (c == label) my_loss -= logf(fmaxf(pr, 1e-38f));
          db[c] += d[c];
        }
        my_cnt += 1.f;
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) s_d[threadIdx.x * C + c] = d[c];
    __syncthreads();
    // transposed upsampling of the tile: output (cell, class) <- its pixels of this tile
    for (int o = threadIdx.x; o < 16 * C; o += 256) {
      const int c = o % C, cell = o / C;
      const int iy = iy0 + cell / 4, ix = ix0 + cell % 4;
      if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
      // pixel (py, px) of the tile reaches the cell with kernel row ky = by + py + 4 - 8 iy
      const int ky0 = by + 4 - 8 * iy, kx0 = bx + 4 - 8 * ix;
      const int py_lo = ky0 < 0 ? -ky0 : 0, py_hi = 16 - ky0 < 16 ? 16 - ky0 : 16;
      const int px_lo = kx0 < 0 ? -kx0 : 0, px_hi = 16 - kx0 < 16 ? 16 - kx0 : 16;
      float acc = 0.f;
      for (int py = py_lo; py < py_hi; ++py) {
        const float* grow = s_g + (ky0 + py) * 16 + kx0;
        const float* drow = s_d + (py * 16) * C + c;
        for (int px = px_lo; px < px_hi; ++px) acc = fmaf(grow[px], drow[px * C], acc);
      }
      atomicAdd(dlow + ((static_cast<size_t>(img) * h + iy) * w + ix) * C + c, acc);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float t = db[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_db[c], t);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    my_loss += __shfl_xor_sync(0xffffffffu, my_loss, off);
    my_cnt += __shfl_xor_sync(0xffffffffu, my_cnt, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_loss[0], my_loss);
    atomicAdd(&s_loss[1], my_cnt);
  }
  __syncthreads();
  if (threadIdx.x < C) atomicAdd(dbias + threadIdx.x, s_db[threadIdx.x]);
  if (threadIdx.x < 2) atomicAdd(loss + threadIdx.x, static_cast<double>(s_loss[threadIdx.x]));
}

// Transpose of the x8 upsampling (16x16 stride-8 shared kernel g): one warp per low-res pixel.
// dlow[n,iy,ix,c] = sum_{ky,kx} g[ky][kx] * dscore[n, 8iy-4+ky, 8ix-4+kx, c]
template <int C>
__global__ void __launch_bounds__(kThreads)
upsample8_transpose_kernel(const float* __restrict__ dscore, const float* __restrict__ g,
                           float* __restrict__ dlow, int N, int h, int w) {
  const int H = 8 * h, W = 8 * w;
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(N) * h * w;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (kThreads / 32);
  for (int64_t cell = blockIdx.x * static_cast<int64_t>(kThreads / 32) + (threadIdx.x >> 5);
       cell < total; cell += warps) {
    const int ix = static_cast<int>(cell % w);
    const int iy = static_cast<int>((cell / w) % h);
    const int64_t img = cell / (static_cast<int64_t>(w) * h);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int t = lane; t < 256; t += 32) {
      const int ky = t >> 4, kx = t & 15;
      const int oy = 8 * iy - 4 + ky, ox = 8 * ix - 4 + kx;
      if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;
      const float wgt = __ldg(g + t);
      const float* src = dscore + ((img * H + oy) * W + ox) * C;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(wgt, __ldg(src + c), acc[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float t = acc[c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      if (lane == 0) dlow[cell * C + c] = t;
    }
  }
}

// 1x1 score conv at low resolution: low = fused . W  ->  dfused = dlow . W^T, dW += fused^T . dlow
__global__ void score_bwd_data_kernel(const float* __restrict__ dlow, const float* __restrict__ w,
                                      float* __restrict__ dfused, size_t npix, int nu, int C) {
  const size_t total = npix * nu;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % nu);
    const size_t p = idx / nu;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(__ldg(dlow + p * C + c), __ldg(w + u * C + c), acc);
    dfused[idx] = acc;
  }
}
// generic small "A^T . B" over pixels: dW[i][j] += sum_p a[p][i] * b[p][j]  (a: [P,I], b: [P,J])
template <typename TA>
__global__ void __launch_bounds__(kThreads)
outer_sum_kernel(const TA* __restrict__ a, const float* __restrict__ b, float* __restrict__ dw,
                 size_t npix, int I, int J) {
  // each block owns a contiguous pixel range; thread t owns entries t, t+256, ... of [I,J]
  const size_t per_block = (npix + gridDim.x - 1) / gridDim.x;
  const size_t p0 = blockIdx.x * per_block;
  const size_t p1 = p0 + per_block < npix ? p0 + per_block : npix;
  for (int e = threadIdx.x; e < I * J; e += kThreads) {
    const int i = e / J, j = e - i * J;
    float acc = 0.f;
    for (size_t p = p0; p < p1; ++p)
      acc = fmaf(static_cast<float>(a[p * I + i]), __ldg(b + p * J + j), acc);
    atomicAdd(dw + e, acc);
  }
}

// Transpose of upscore_conv5 (4x4 stride-2 channel-diagonal kernel g[ky,kx,u]) incl. its ReLU:
// ds5[n,iy,ix,u] = sum_{ky,kx} g * dfused[n,2iy-1+ky,2ix-1+kx,u] * [up5 > 0]
__global__ void upscore2_bwd_kernel(const float* __restrict__ dfused, const float* __restrict__ up5,
                                    const float* __restrict__ g, float* __restrict__ ds5, int N,
                                    int h, int w, int nu) {
  const int ho = 2 * h, wo = 2 * w;
  const size_t total = static_cast<size_t>(N) * h * w * nu;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % nu);
    size_t t = idx / nu;
    const int ix = static_cast<int>(t % w);
    t /= w;
    const int iy = static_cast<int>(t % h);
    const size_t img = t / h;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int oy = 2 * iy - 1 + ky;
      if (oy < 0 || oy >= ho) continue;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ox = 2 * ix - 1 + kx;
        if (ox < 0 || ox >= wo) continue;
        const size_t o = ((img * ho + oy) * wo + ox) * nu + u;
        if (up5[o] > 0.f) acc = fmaf(__ldg(g + (ky * 4 + kx) * nu + u), dfused[o], acc);
      }
    }
    ds5[idx] = acc;
  }
}

// head convs (1x1, ReLU, fp32 output y): dpre = dy * [y > 0], also as bf16 for the data-gradient GEMM
__global__ void relu_mask_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                     float* __restrict__ dpre, __nv_bfloat16* __restrict__ dpre_bf16,
                                     size_t npix, int c, int c_pad) {
  const size_t total = npix * c_pad;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c_pad);
    const size_t p = i / c_pad;
    float v = 0.f;
    if (ch < c) {
      v = y[p * c + ch] > 0.f ? dy[p * c + ch] : 0.f;
      dpre[p * c + ch] = v;
    }
    dpre_bf16[i] = __float2bfloat16_rn(v);
  }
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    f[2 * j] = __bfloat162float(t.x);
    f[2 * j + 1] = __bfloat162float(t.y);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
// per-thread partial channel sums -> block shared memory -> global (fused bias gradient)
__device__ __forceinline__ void flush_bias(const float (&acc)[8], int octet, int cout,
                                           float* __restrict__ bias_grad) {
  extern __shared__ float s_bias[];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) s_bias[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&s_bias[octet * 8 + j], acc[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < cout; i += blockDim.x) atomicAdd(bias_grad + i, s_bias[i]);
}

// dy <- (y > 0) ? (da [+ db]) : 0  (bf16, one thread = 8 channels of one pixel); optionally
// bias_grad[c] += sum over pixels of dy[., c] (a thread keeps the same channel octet because the
// grid stride is a multiple of cout/8).
__global__ void __launch_bounds__(kThreads)
relu_bwd_bf16_kernel(const uint4* __restrict__ da, const uint4* __restrict__ db,
                     const uint4* __restrict__ y, uint4* __restrict__ dy, size_t n8, int cout,
                     float* __restrict__ bias_grad) {
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float a[8], yy[8], o[8];
    unpack8(da[i], a);
    unpack8(y[i], yy);
    if (db) {
      float b[8];
      unpack8(db[i], b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = yy[j] > 0.f ? a[j] : 0.f;
      acc[j] += o[j];
    }
    dy[i] = pack8(o);
  }
  if (bias_grad) flush_bias(acc, threadIdx.x % (cout / 8), cout, bias_grad);
}

// Fused 2x2/2 max-pool backward + ReLU backward (+ optional second gradient source + bias grad):
// the gradient of a pooled element goes to the FIRST window position holding the maximum
// (row-major); dy = (y > 0) ? (pool-routed dp [+ extra]) : 0.  One thread = one pooled pixel x 8
// channels; y / extra / dy are full resolution.
__global__ void __launch_bounds__(kThreads)
pool_relu_bwd_bf16_kernel(const uint4* __restrict__ dp, const uint4* __restrict__ y,
                          const uint4* __restrict__ p, const uint4* __restrict__ extra,
                          uint4* __restrict__ dy, int N, int H, int W, int C8, int cout,
                          float* __restrict__ bias_grad) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    size_t t = idx / C8;
    const int xo = static_cast<int>(t % Wo);
    t /= Wo;
    const int yo = static_cast<int>(t % Ho);
    const size_t n = t / Ho;
    const size_t base = ((n * H + 2 * yo) * W + 2 * xo) * C8 + c;
    const size_t offs[4] = {base, base + C8, base + static_cast<size_t>(W) * C8,
                            base + static_cast<size_t>(W) * C8 + C8};
    float g[8], pv[8];
    unpack8(dp[idx], g);
    unpack8(p[idx], pv);
    bool given[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) given[j] = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float yy[8], o[8], e[8];
      unpack8(y[offs[q]], yy);
      if (extra) unpack8(extra[offs[q]], e);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool hit = !given[j] && yy[j] == pv[j];
        given[j] = given[j] || hit;
        float v = hit ? g[j] : 0.f;
        if (extra) v += e[j];
        o[j] = yy[j] > 0.f ? v : 0.f;
        acc[j] += o[j];
      }
      dy[offs[q]] = pack8(o);
    }
  }
  if (bias_grad) flush_bias(acc, threadIdx.x % C8, cout, bias_grad);
}

// ------------------------------------------------------------------ weight / bias gradients
// db[co] += sum_p dy[p][co]
__global__ void __launch_bounds__(kThreads)
bias_grad_bf16_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db, size_t npix,
                      int cout) {
  // block = 256 threads = (256 / cout_tile) pixel lanes x cout_tile channels, cout_tile = min(cout,256)
  const int ct = cout < 256 ? cout : 256;
  const int lanes = kThreads / ct;
  const int ch = threadIdx.x % ct, lane = threadIdx.x / ct;
  for (int c0 = 0; c0 < cout; c0 += ct) {
    float acc = 0.f;
    for (size_t p = blockIdx.x * static_cast<size_t>(lanes) + lane; p < npix;
         p += static_cast<size_t>(gridDim.x) * lanes)
      acc += __bfloat162float(dy[p * cout + c0 + ch]);
    atomicAdd(db + c0 + ch, acc);
  }
}
__global__ void __launch_bounds__(kThreads)
bias_grad_f32_kernel(const float* __restrict__ dy, float* __restrict__ db, size_t npix, int cout) {
  const int lanes = kThreads / cout;            // cout <= 256
  const int ch = threadIdx.x % cout, lane = threadIdx.x / cout;
  if (lane >= lanes) return;
  float acc = 0.f;
  for (size_t p = blockIdx.x * static_cast<size_t>(lanes) + lane; p < npix;
       p += static_cast<size_t>(gridDim.x) * lanes)
    acc += dy[p * cout + ch];
  atomicAdd(db + ch, acc);
}

// 3x3 weight gradient, fp32 accumulation on the CUDA cores:
//   dW[tap][ci][co] += sum_p x[p + tap][ci] * dy[p][co]        (x, dy bf16 NHWC)
// Block = one (32 ci, 64 co) chunk pair x one slice of 8x8-pixel tiles.  Thread (ci, co-octet)
// keeps 9 taps x 8 channels = 72 accumulators in registers for its whole slice, the halo'd
// input tile and the dy tile are staged in shared memory.
constexpr int kWgCi = 32, kWgCo = 64, kWgTile = 8;
__global__ void __launch_bounds__(kThreads)
conv_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                  float* __restrict__ dw, int N, int H, int W, int cin, int cout, int tiles_x,
                  int tiles_y, int num_slices) {
  __shared__ __nv_bfloat16 s_x[(kWgTile + 2) * (kWgTile + 2) * kWgCi];   // 100 px x 32 ci
  __shared__ __nv_bfloat16 s_dy[kWgTile * kWgTile * kWgCo];              // 64 px x 64 co
  const int ci_chunks = cin / kWgCi, co_chunks = cout / kWgCo;
  const int pair = blockIdx.x % (ci_chunks * co_chunks);
  const int slice = blockIdx.x / (ci_chunks * co_chunks);
  const int ci0 = (pair % ci_chunks) * kWgCi, co0 = (pair / ci_chunks) * kWgCo;
  const int ci = threadIdx.x & 31;             // 0..31
  const int cog = threadIdx.x >> 5;            // 0..7 -> co octet
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int total_tiles = N * tiles_y * tiles_x;
  for (int tile = slice; tile < total_tiles; tile += num_slices) {
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int img = tile / (tiles_x * tiles_y);
    const int y0 = ty * kWgTile, x0 = tx * kWgTile;
    __syncthreads();
    // halo'd input tile: 100 pixels x 32 channels (64 B per pixel = 4 x 16 B)
    for (int i = threadIdx.x; i < (kWgTile + 2) * (kWgTile + 2) * 4; i += kThreads) {
      const int piece = i & 3, pix = i >> 2;
      const int yy = y0 + pix / (kWgTile + 2) - 1, xx = x0 + pix % (kWgTile + 2) - 1;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        v = __ldg(reinterpret_cast<const uint4*>(
            x + ((static_cast<size_t>(img) * H + yy) * W + xx) * cin + ci0 + piece * 8));
      reinterpret_cast<uint4*>(s_x)[i] = v;
    }
    // dy tile: 64 pixels x 64 channels (128 B per pixel = 8 x 16 B)
    for (int i = threadIdx.x; i < kWgTile * kWgTile * 8; i += kThreads) {
      const int piece = i & 7, pix = i >> 3;
      const int yy = y0 + pix / kWgTile, xx = x0 + pix % kWgTile;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (yy < H && xx < W)
        v = __ldg(reinterpret_cast<const uint4*>(
            dy + ((static_cast<size_t>(img) * H + yy) * W + xx) * cout + co0 + piece * 8));
      reinterpret_cast<uint4*>(s_dy)[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int pix = 0; pix < kWgTile * kWgTile; ++pix) {
      const int py = pix / kWgTile, px = pix % kWgTile;
      const uint4 dv = reinterpret_cast<const uint4*>(s_dy)[pix * 8 + cog];   // warp broadcast
      const uint32_t dwv[4] = {dv.x, dv.y, dv.z, dv.w};
      float d[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 t2 = *reinterpret_cast<const __nv_bfloat162*>(&dwv[j]);
        d[2 * j] = __bfloat162float(t2.x);
        d[2 * j + 1] = __bfloat162float(t2.y);
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float xv = __bfloat162float(
            s_x[((py + t / 3) * (kWgTile + 2) + px + t % 3) * kWgCi + ci]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(xv, d[j], acc[t][j]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      atomicAdd(dw + (static_cast<size_t>(t) * cin + ci0 + ci) * cout + co0 + cog * 8 + j, acc[t][j]);
}

// conv1_1 weight gradient (raw fp32 input, CIN <= 3):  dW[tap][ci][co] = sum_p x[p + tap][ci] * dy[p][co].
// Thread = (group of CH output channels, pixel lane): one 16-byte (CH = 8) or 8-byte (CH = 4) dy
// load and 9*CIN input loads feed 9*CIN*CH FMAs, so the kernel is FMA-bound instead of load-bound
// (the one-channel-per-thread version issued ten loads per nine FMAs and took 1.9 - 3 ms at batch
// 16).  Whole image rows per block (no per-pixel index arithmetic); the pixel lanes of a block are
// summed in shared memory and few blocks are launched, because the 9*CIN*Cout global atomics all
// hit the same addresses.
template <int CIN, int CH>
__global__ void __launch_bounds__(kThreads)
conv_wgrad_c1_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                     float* __restrict__ dw, int N, int H, int W, int cout) {
  constexpr int K = 9 * CIN;
  extern __shared__ float s_part[];                        // [K][cout]
  const int groups = cout / CH;                            // channel groups per pixel
  const int lanes = kThreads / groups;                     // pixel lanes per block
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  for (int i = threadIdx.x; i < K * cout; i += kThreads) s_part[i] = 0.f;
  __syncthreads();
  const int rows = N * H;
  const int per_block = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per_block;
  const int r1 = r0 + per_block < rows ? r0 + per_block : rows;
  float acc[K][CH];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < CH; ++j) acc[k][j] = 0.f;
  // the three input rows a gradient row needs live in shared memory (a rotating window over the
  // block's rows): the 9 * CIN input reads per pixel are LDS instead of L1-hitting LDG, whose
  // latency the two resident blocks per SM (72+ accumulators per thread) could not hide
  float* s_x = s_part + K * cout;                          // [3][W * CIN], slot = row mod 3
  const int row_elems = W * CIN;
  auto stage_row = [&](int r) {                            // row r of the stacked images, or zeros
    float* dst = s_x + ((r % 3 + 3) % 3) * row_elems;
    const bool in = r >= 0 && r < rows;
    for (int i = threadIdx.x; i < row_elems; i += kThreads)
      dst[i] = in ? __ldg(x + static_cast<size_t>(r) * row_elems + i) : 0.f;
  };
  if (r0 < r1) {
    stage_row(r0 - 1);
    stage_row(r0);
  }
  // kU pixels per step, their dy vectors loaded up front: one 16-byte load in flight per thread
  // left the kernel bound by memory latency (0.62 ms for 604 MB at batch 16; 8 per step costs a
  // resident block)
  constexpr int kU = 4;
  for (int r = r0; r < r1; ++r) {
    stage_row(r + 1);
    __syncthreads();
    const int py = r % H;
    const __nv_bfloat16* dyrow = dy + static_cast<size_t>(r) * W * cout + g * CH;
    const bool up = py > 0, down = py + 1 < H;
    for (int px0 = lane; lane < lanes && px0 < W; px0 += lanes * kU) {
      uint4 raw[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int px = px0 + u * lanes;
        raw[u] = make_uint4(0u, 0u, 0u, 0u);
        if (px < W) {
          if constexpr (CH == 8) {
            raw[u] = __ldg(reinterpret_cast<const uint4*>(dyrow + static_cast<size_t>(px) * cout));
          } else {
            const uint2 v =
                __ldg(reinterpret_cast<const uint2*>(dyrow + static_cast<size_t>(px) * cout));
            raw[u].x = v.x;
            raw[u].y = v.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int px = px0 + u * lanes;
        if (px >= W) break;
        float d[CH];
        if constexpr (CH == 8) {
          unpack8(raw[u], d);
        } else {
          const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw[u].x);
          const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw[u].y);
          d[0] = __bfloat162float(a.x);
          d[1] = __bfloat162float(a.y);
          d[2] = __bfloat162float(b.x);
          d[3] = __bfloat162float(b.y);
        }
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) {
          if ((ty == 0 && !up) || (ty == 2 && !down)) continue;
          const float* src = s_x + ((r + ty - 1) % 3 + 3) % 3 * row_elems + px * CIN;
#pragma unroll
          for (int tx = 0; tx < 3; ++tx) {
            const int xx = px + tx - 1;
            if (xx < 0 || xx >= W) continue;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
              const float xv = src[(tx - 1) * CIN + ci];
#pragma unroll
              for (int j = 0; j < CH; ++j)
                acc[(ty * 3 + tx) * CIN + ci][j] = fmaf(xv, d[j], acc[(ty * 3 + tx) * CIN + ci][j]);
            }
          }
        }
      }
    }
    __syncthreads();              // row r - 1's slot is overwritten by the next stage_row
  }
  if (lane < lanes) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < CH; ++j) atomicAdd(&s_part[k * cout + g * CH + j], acc[k][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * cout; i += kThreads) atomicAdd(dw + i, s_part[i]);
}

// ------------------------------------------------------------------ optimizer + re-packing
__global__ void scale_kernel(float* __restrict__ g, size_t n, const double* __restrict__ loss) {
  const float s = static_cast<float>(1.0 / (1e-20 + loss[1]));
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    g[i] *= s;
}
// tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t); m, v moments; w -= lr_t m/(sqrt(v)+eps)
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr_t, float b1, float b2,
                            float eps) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    w[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
// fp32 HWIO master -> bf16 operand rows.  forward: [co][tap][ci]; backward (flipped + transposed,
// for the data gradient): [ci][8-tap][co].  conv1_1 layout: [co][hi taps | lo taps | 0] (K = 64).
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                    __nv_bfloat16* __restrict__ bwd, int taps, int cin, int cout,
                                    int fwd_kdim, int bwd_kdim, int c1_layout) {
  const size_t total = static_cast<size_t>(taps) * cin * cout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % cout);
    const int ci = static_cast<int>((i / cout) % cin);
    const int t = static_cast<int>(i / (static_cast<size_t>(cout) * cin));
    const __nv_bfloat16 q = __float2bfloat16_rn(w[i]);
    if (c1_layout) {
      fwd[static_cast<size_t>(co) * 64 + t * cin + ci] = q;
      fwd[static_cast<size_t>(co) * 64 + 9 * cin + t * cin + ci] = q;
    } else {
      fwd[static_cast<size_t>(co) * fwd_kdim + t * cin + ci] = q;
    }
    if (bwd) bwd[static_cast<size_t>(ci) * bwd_kdim + (taps - 1 - t) * cout + co] = q;
  }
}

#define XV_DISPATCH_C(C, CALL)                                                          \
  switch (C) {                                                                          \
    case 2: { constexpr int kC = 2; CALL; break; }                                      \
    case 3: { constexpr int kC = 3; CALL; break; }                                      \
    case 4: { constexpr int kC = 4; CALL; break; }                                      \
    case 5: { constexpr int kC = 5; CALL; break; }                                      \
    case 6: { constexpr int kC = 6; CALL; break; }                                      \
    case 7: { constexpr int kC = 7; CALL; break; }                                      \
    case 8: { constexpr int kC = 8; CALL; break; }                                      \
    case 9: { constexpr int kC = 9; CALL; break; }                                      \
    case 10: { constexpr int kC = 10; CALL; break; }                                    \
    case 11: { constexpr int kC = 11; CALL; break; }                                    \
    case 12: { constexpr int kC = 12; CALL; break; }                                    \
    case 13: { constexpr int kC = 13; CALL; break; }                                    \
    case 14: { constexpr int kC = 14; CALL; break; }                                    \
    case 15: { constexpr int kC = 15; CALL; break; }                                    \
    case 16: { constexpr int kC = 16; CALL; break; }                                    \
    case 17: { constexpr int kC = 17; CALL; break; }                                    \
    case 18: { constexpr int kC = 18; CALL; break; }                                    \
    case 19: { constexpr int kC = 19; CALL; break; }                                    \
    case 20: { constexpr int kC = 20; CALL; break; }                                    \
    case 21: { constexpr int kC = 21; CALL; break; }                                    \
    case 22: { constexpr int kC = 22; CALL; break; }                                    \
    case 23: { constexpr int kC = 23; CALL; break; }                                    \
    case 24: { constexpr int kC = 24; CALL; break; }                                    \
    default: return fail("num_classes must be in [2, 24]");                             \
  }

#define XV_LAUNCHED()                 \
  XV_CUDA(cudaGetLastError());        \
  count_launch();                     \
  return 0

}  // namespace

int launch_ce_grad(float* score, const int32_t* labels, int64_t npix, int C, double* loss,
                   float* dbias, cudaStream_t s) {
  XV_DISPATCH_C(C, (ce_grad_kernel<kC><<<grid_for(npix, kThreads, 4), kThreads, 0, s>>>(
                       score, labels, npix, loss, dbias)));
  XV_LAUNCHED();
}
int launch_loss_lowres_grad(const float* low, const float* g, const float* bias,
                            const int32_t* labels, int N, int h, int w, int C, float* dlow,
                            double* loss, float* dbias, cudaStream_t s) {
  XV_CUDA(cudaMemsetAsync(dlow, 0, static_cast<size_t>(N) * h * w * C * sizeof(float), s));
  const long long tiles = static_cast<long long>(div_up(8 * w, 16)) * div_up(8 * h, 16) * N;
  const long long cap = static_cast<long long>(device_info().num_sms) * 4;
  const int grid = static_cast<int>(tiles < cap ? tiles : cap);
  XV_DISPATCH_C(C, (loss_lowres_grad_kernel<kC><<<grid, 256, 0, s>>>(low, g, bias, labels, N, h, w,
                                                                     dlow, loss, dbias)));
  XV_LAUNCHED();
}
int launch_upsample8_transpose(const float* dscore, const float* g, float* dlow, int N, int h,
                               int w, int C, cudaStream_t s) {
  const int64_t cells = static_cast<int64_t>(N) * h * w;
  XV_DISPATCH_C(C, (upsample8_transpose_kernel<kC><<<grid_for(cells * 32), kThreads, 0, s>>>(
                       dscore, g, dlow, N, h, w)));
  XV_LAUNCHED();
}
int launch_score_bwd(const float* dlow, const float* fused, const float* w, float* dfused,
                     float* dw, size_t npix, int nu, int C, cudaStream_t s) {
  score_bwd_data_kernel<<<grid_for(npix * nu), kThreads, 0, s>>>(dlow, w, dfused, npix, nu, C);
  XV_CUDA(cudaGetLastError());
  outer_sum_kernel<float><<<device_info().num_sms * 2, kThreads, 0, s>>>(fused, dlow, dw, npix, nu, C);
  count_launch();
  XV_LAUNCHED();
}
int launch_outer_sum_bf16(const __nv_bfloat16* a, const float* b, float* dw, size_t npix, int I,
                          int J, cudaStream_t s) {
  outer_sum_kernel<__nv_bfloat16><<<device_info().num_sms * 2, kThreads, 0, s>>>(a, b, dw, npix, I, J);
  XV_LAUNCHED();
}
int launch_upscore2_bwd(const float* dfused, const float* up5, const float* g, float* ds5, int N,
                        int h, int w, int nu, cudaStream_t s) {
  upscore2_bwd_kernel<<<grid_for(static_cast<size_t>(N) * h * w * nu), kThreads, 0, s>>>(
      dfused, up5, g, ds5, N, h, w, nu);
  XV_LAUNCHED();
}
int launch_relu_mask_f32(const float* dy, const float* y, float* dpre, __nv_bfloat16* dpre_bf16,
                         size_t npix, int c, int c_pad, cudaStream_t s) {
  relu_mask_f32_kernel<<<grid_for(npix * c_pad), kThreads, 0, s>>>(dy, y, dpre, dpre_bf16, npix, c,
                                                                  c_pad);
  XV_LAUNCHED();
}
int launch_relu_bwd_bf16(const __nv_bfloat16* da, const __nv_bfloat16* db, const __nv_bfloat16* y,
                         __nv_bfloat16* dy, size_t n, int cout, float* bias_grad, cudaStream_t s) {
  XV_CHECK(n % 8 == 0 && cout % 8 == 0 && kThreads % (cout / 8) == 0,
           "relu_bwd: channel count must be a multiple of 8 dividing 2048");
  relu_bwd_bf16_kernel<<<grid_for(n / 8), kThreads, cout * sizeof(float), s>>>(
      reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(db),
      reinterpret_cast<const uint4*>(y), reinterpret_cast<uint4*>(dy), n / 8, cout, bias_grad);
  XV_LAUNCHED();
}
int launch_pool_relu_bwd_bf16(const __nv_bfloat16* dp, const __nv_bfloat16* y,
                              const __nv_bfloat16* p, const __nv_bfloat16* extra,
                              __nv_bfloat16* dy, int N, int H, int W, int C, float* bias_grad,
                              cudaStream_t s) {
  XV_CHECK(C % 8 == 0 && kThreads % (C / 8) == 0, "pool_relu_bwd: unsupported channel count");
  const size_t total = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
  pool_relu_bwd_bf16_kernel<<<grid_for(total), kThreads, C * sizeof(float), s>>>(
      reinterpret_cast<const uint4*>(dp), reinterpret_cast<const uint4*>(y),
      reinterpret_cast<const uint4*>(p), reinterpret_cast<const uint4*>(extra),
      reinterpret_cast<uint4*>(dy), N, H, W, C / 8, C, bias_grad);
  XV_LAUNCHED();
}
int launch_bias_grad_bf16(const __nv_bfloat16* dy, float* db, size_t npix, int cout,
                          cudaStream_t s) {
  XV_CHECK(cout <= 256 ? 256 % cout == 0 : cout % 256 == 0, "bias_grad: unsupported Cout");
  bias_grad_bf16_kernel<<<device_info().num_sms * 2, kThreads, 0, s>>>(dy, db, npix, cout);
  XV_LAUNCHED();
}
int launch_bias_grad_f32(const float* dy, float* db, size_t npix, int cout, cudaStream_t s) {
  XV_CHECK(cout <= 256, "bias_grad_f32: Cout too large");
  bias_grad_f32_kernel<<<device_info().num_sms * 2, kThreads, 0, s>>>(dy, db, npix, cout);
  XV_LAUNCHED();
}
int launch_conv_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int N, int H,
                      int W, int cin, int cout, cudaStream_t s) {
  XV_CHECK(cin % kWgCi == 0 && cout % kWgCo == 0, "conv_wgrad: Cin % 32 and Cout % 64 required");
  const int tiles_x = div_up(W, kWgTile), tiles_y = div_up(H, kWgTile);
  const int pairs = (cin / kWgCi) * (cout / kWgCo);
  const int total_tiles = N * tiles_x * tiles_y;
  int slices = div_up(device_info().num_sms * 2, pairs);
  if (slices > total_tiles) slices = total_tiles;
  if (slices < 1) slices = 1;
  conv_wgrad_kernel<<<pairs * slices, kThreads, 0, s>>>(x, dy, dw, N, H, W, cin, cout, tiles_x,
                                                        tiles_y, slices);
  XV_LAUNCHED();
}
int launch_conv_wgrad_c1(const float* x, const __nv_bfloat16* dy, float* dw, int N, int H, int W,
                         int cin, int cout, cudaStream_t s) {
  XV_CHECK(cout % 8 == 0 && cout <= 256 && cin >= 1 && cin <= 3,
           "conv_wgrad_c1: Cin <= 3, Cout a multiple of 8 and <= 256");
  XV_CHECK((reinterpret_cast<uintptr_t>(dy) & 15) == 0, "conv_wgrad_c1: dy must be 16-byte aligned");
  int grid = device_info().num_sms * 4;
  if (grid > N * H) grid = N * H;
  const size_t smem = (static_cast<size_t>(9) * cin * cout + static_cast<size_t>(3) * W * cin) *
                      sizeof(float);
  XV_CHECK(smem <= 200 * 1024, "conv_wgrad_c1: image too wide for the staged input rows");
  if (smem > 48 * 1024) {      // beyond the default limit: opt in (idempotent, cheap)
    const int bytes = static_cast<int>(smem);
    XV_CUDA(cudaFuncSetAttribute(conv_wgrad_c1_kernel<1, 8>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    XV_CUDA(cudaFuncSetAttribute(conv_wgrad_c1_kernel<2, 4>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    XV_CUDA(cudaFuncSetAttribute(conv_wgrad_c1_kernel<3, 4>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
  if (cin == 1) conv_wgrad_c1_kernel<1, 8><<<grid, kThreads, smem, s>>>(x, dy, dw, N, H, W, cout);
  if (cin == 2) conv_wgrad_c1_kernel<2, 4><<<grid, kThreads, smem, s>>>(x, dy, dw, N, H, W, cout);
  if (cin == 3) conv_wgrad_c1_kernel<3, 4><<<grid, kThreads, smem, s>>>(x, dy, dw, N, H, W, cout);
  XV_LAUNCHED();
}
int launch_scale_by_count(float* g, size_t n, const double* loss, cudaStream_t s) {
  scale_kernel<<<grid_for(n), kThreads, 0, s>>>(g, n, loss);
  XV_LAUNCHED();
}
// tf.train.AdagradOptimizer: acc += g^2; w -= lr * g / sqrt(acc)   (acc starts at 0.1)
__global__ void adagrad_kernel(float* __restrict__ w, const float* __restrict__ g,
                               float* __restrict__ acc, size_t n, float lr) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float a = acc[i] + gi * gi;
    acc[i] = a;
    w[i] -= lr * gi * rsqrtf(a);
  }
}
// tf.train.RMSPropOptimizer: ms = decay * ms + (1 - decay) * g^2;
// mom = momentum * mom + lr * g / sqrt(ms + eps); w -= mom   (ms starts at 1, mom at 0)
__global__ void rmsprop_kernel(float* __restrict__ w, const float* __restrict__ g,
                               float* __restrict__ ms, float* __restrict__ mom, size_t n, float lr,
                               float decay, float momentum, float eps) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float m2 = decay * ms[i] + (1.f - decay) * gi * gi;
    ms[i] = m2;
    const float step = momentum * mom[i] + lr * gi / sqrtf(m2 + eps);
    mom[i] = step;
    w[i] -= step;
  }
}
__global__ void fill_f32_kernel(float* __restrict__ x, size_t n, float value) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    x[i] = value;
}
int launch_adagrad(float* w, const float* g, float* acc, size_t n, float lr, cudaStream_t s) {
  adagrad_kernel<<<grid_for(n), kThreads, 0, s>>>(w, g, acc, n, lr);
  XV_LAUNCHED();
}
int launch_rmsprop(float* w, const float* g, float* ms, float* mom, size_t n, float lr,
                   float decay, float momentum, float eps, cudaStream_t s) {
  rmsprop_kernel<<<grid_for(n), kThreads, 0, s>>>(w, g, ms, mom, n, lr, decay, momentum, eps);
  XV_LAUNCHED();
}
int launch_fill_f32(float* x, size_t n, float value, cudaStream_t s) {
  fill_f32_kernel<<<grid_for(n), kThreads, 0, s>>>(x, n, value);
  XV_LAUNCHED();
}
int launch_adam(float* w, const float* g, float* m, float* v, size_t n, float lr_t, float b1,
                float b2, float eps, cudaStream_t s) {
  adam_kernel<<<grid_for(n), kThreads, 0, s>>>(w, g, m, v, n, lr_t, b1, b2, eps);
  XV_LAUNCHED();
}
// All layers of a network in ONE launch (after every optimizer step: 15 pack launches + 15 bias
// copies per step otherwise).  Element i of the concatenated HWIO kernels belongs to the layer
// whose [first, first + elems) range holds it; the biases ride along at the front of the grid.
__global__ void pack_all_kernel(const __grid_constant__ PackAllParams p) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (int l = 0; l < p.num_layers; ++l)
    for (size_t j = tid; j < static_cast<size_t>(p.layer[l].cout); j += stride)
      p.layer[l].bias_pad[j] = p.layer[l].b[j];
  int l = 0;
  for (size_t i = tid; i < p.total; i += stride) {
    while (i >= p.layer[l].first + p.layer[l].elems) ++l;      // i grows: l only moves forward
    const PackLayer& d = p.layer[l];
    const size_t k = i - d.first;
    const int co = static_cast<int>(k % d.cout);
    const int ci = static_cast<int>((k / d.cout) % d.cin);
    const int t = static_cast<int>(k / (static_cast<size_t>(d.cout) * d.cin));
    const __nv_bfloat16 q = __float2bfloat16_rn(d.w[k]);
    if (d.c1_layout) {
      d.fwd[static_cast<size_t>(co) * 64 + t * d.cin + ci] = q;
      d.fwd[static_cast<size_t>(co) * 64 + 9 * d.cin + t * d.cin + ci] = q;
    } else {
      d.fwd[static_cast<size_t>(co) * d.fwd_kdim + t * d.cin + ci] = q;
    }
    if (d.bwd) d.bwd[static_cast<size_t>(ci) * d.bwd_kdim + (d.taps - 1 - t) * d.cout + co] = q;
  }
}

int launch_pack_all(const PackAllParams& p, cudaStream_t s) {
  pack_all_kernel<<<grid_for(p.total), kThreads, 0, s>>>(p);
  XV_LAUNCHED();
}
int launch_pack_weights(const float* w, __nv_bfloat16* fwd, __nv_bfloat16* bwd, int taps, int cin,
                        int cout, int fwd_kdim, int bwd_kdim, int c1_layout, cudaStream_t s) {
  pack_weights_kernel<<<grid_for(static_cast<size_t>(taps) * cin * cout), kThreads, 0, s>>>(
      w, fwd, bwd, taps, cin, cout, fwd_kdim, bwd_kdim, c1_layout);
  XV_LAUNCHED();
}

}  // namespace xv
