// Batch normalisation in TRAINING mode for SimpleFCN.fit() with batch_normalization=True
// (xview/models/custom_layers.py:112-119,127-136 -> tf.layers.batch_normalization(training=True),
// update ops run with every step at base_model.py:155-156), plus the dense-free forms of the two
// channel-diagonal transposed convolutions that the batch-normalised decoder needs at full
// resolution (simple_fcn.py:82-83,129-130).
//
//   forward   z (conv output, bf16 or fp32, NHWC) -> per-channel batch mean / biased variance
//             (float64 accumulation) -> y = [relu](gamma * (z - mean) * rstd + beta);
//             moving_mean / moving_variance <- 0.99 * moving + 0.01 * batch statistic (the
//             variance fed to the moving average is Bessel-corrected, as TF's fused kernel does).
//   backward  ghat = dL/dy masked by the ReLU; dbeta = sum ghat, dgamma = sum ghat * zhat,
//             dz = gamma * rstd * (ghat - dbeta / n - zhat * dgamma / n).
// All kernels are plain HBM-streaming passes with channels as the fastest index.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kThreads = 256;

inline int grid_for(size_t work, int threads = kThreads, int per_sm = 8) {
  size_t g = (work + threads - 1) / threads;
  size_t cap = static_cast<size_t>(device_info().num_sms) * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

template <typename T>
__device__ __forceinline__ float ld(const T* p, size_t i);
template <>
__device__ __forceinline__ float ld<float>(const float* p, size_t i) {
  return p[i];
}
template <>
__device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}
__device__ __forceinline__ void st(float* p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void st(__nv_bfloat16* p, size_t i, float v) {
  p[i] = __float2bfloat16_rn(v);
}

// Thread layout shared by the per-channel reductions: a block is (256 / ct) pixel lanes x ct
// channels with ct = min(C, 256); channel tiles beyond 256 are walked in a loop.  A thread keeps
// the same channel for all its pixels, so its partial sums live in registers (float64).

// sums[c] += sum_p z[p][c], sums[C + c] += sum_p z[p][c]^2
template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(const T* __restrict__ z, size_t npix, int C, double* __restrict__ sums) {
  const int ct = C < kThreads ? C : kThreads;
  const int lanes = kThreads / ct;
  const int ch = threadIdx.x % ct, lane = threadIdx.x / ct;
  if (lane >= lanes) return;
  for (int c0 = 0; c0 < C; c0 += ct) {
    if (c0 + ch >= C) continue;
    double s1 = 0.0, s2 = 0.0;
    for (size_t p = blockIdx.x * static_cast<size_t>(lanes) + lane; p < npix;
         p += static_cast<size_t>(gridDim.x) * lanes) {
      const float v = ld<T>(z, p * C + c0 + ch);
      s1 += v;
      s2 += static_cast<double>(v) * v;
    }
    atomicAdd(sums + c0 + ch, s1);
    atomicAdd(sums + C + c0 + ch, s2);
  }
}

// mean / rstd of the batch; moving statistics updated in place (bias_extra: a bias that was left
// out of z because batch norm cancels it - it still belongs to the moving mean)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double n, int C, float eps,
                                   float momentum, const float* __restrict__ bias_extra,
                                   float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ moving_mean,
                                   float* __restrict__ moving_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sums[c] / n;
  double var = sums[C + c] / n - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(m);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (moving_mean) {
    const double mb = m + (bias_extra ? static_cast<double>(bias_extra[c]) : 0.0);
    const double unbiased = n > 1.0 ? var * (n / (n - 1.0)) : var;
    moving_mean[c] = static_cast<float>(momentum * moving_mean[c] + (1.0 - momentum) * mb);
    moving_var[c] = static_cast<float>(momentum * moving_var[c] + (1.0 - momentum) * unbiased);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const T* __restrict__ z, const float* __restrict__ mean,
                const float* __restrict__ rstd, const float* __restrict__ gamma,
                const float* __restrict__ beta, size_t total, int C, int relu, T* __restrict__ y) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    float v = (ld<T>(z, i) - mean[c]) * rstd[c] * gamma[c] + beta[c];
    if (relu) v = fmaxf(v, 0.f);
    st(y, i, v);
  }
}

// sums[c] += sum ghat, sums[C + c] += sum ghat * zhat;  ghat = g masked by (y > 0) if y != NULL
template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ y, const T* __restrict__ z,
                     const float* __restrict__ mean, const float* __restrict__ rstd, size_t npix,
                     int C, double* __restrict__ sums) {
  const int ct = C < kThreads ? C : kThreads;
  const int lanes = kThreads / ct;
  const int ch = threadIdx.x % ct, lane = threadIdx.x / ct;
  if (lane >= lanes) return;
  for (int c0 = 0; c0 < C; c0 += ct) {
    const int c = c0 + ch;
    if (c >= C) continue;
    const float m = mean[c], r = rstd[c];
    double s1 = 0.0, s2 = 0.0;
    for (size_t p = blockIdx.x * static_cast<size_t>(lanes) + lane; p < npix;
         p += static_cast<size_t>(gridDim.x) * lanes) {
      const size_t i = p * C + c;
      float gh = ld<T>(g, i);
      if (y != nullptr && !(ld<T>(y, i) > 0.f)) gh = 0.f;
      s1 += gh;
      s2 += static_cast<double>(gh) * ((ld<T>(z, i) - m) * r);
    }
    atomicAdd(sums + c, s1);
    atomicAdd(sums + C + c, s2);
  }
}

// dz = gamma * rstd * (ghat - dbeta / n - zhat * dgamma / n); also the two parameter gradients
// (written by block 0) and, for fp32 tensors, an optional bf16 copy zero-padded to c_pad channels
// (operand of the tensor-core data-gradient GEMM of the 1x1 heads).
template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const T* __restrict__ g, const T* __restrict__ y, const T* __restrict__ z,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const double* __restrict__ sums, double n,
                    size_t npix, int C, T* __restrict__ dz, __nv_bfloat16* __restrict__ dz_pad,
                    int c_pad, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = static_cast<float>(sums[c]);
      if (dgamma) dgamma[c] = static_cast<float>(sums[C + c]);
    }
  }
  const size_t total = npix * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    float gh = ld<T>(g, i);
    if (y != nullptr && !(ld<T>(y, i) > 0.f)) gh = 0.f;
    const float zh = (ld<T>(z, i) - mean[c]) * rstd[c];
    const float db = static_cast<float>(sums[c] / n), dg = static_cast<float>(sums[C + c] / n);
    const float v = gamma[c] * rstd[c] * (gh - db - zh * dg);
    st(dz, i, v);
    if (dz_pad) dz_pad[(i / C) * c_pad + c] = __float2bfloat16_rn(v);
  }
}

// out[n,oy,ox,c] = sum over the (k/s)^2 contributing inputs of g(ky,kx,c) * in[n,iy,ix,c] with
// oy = s*iy - (k-s)/2 + ky: the channel-diagonal transposed convolution ('same', no bias).
// g: per_channel ? [k*k][C] : [k*k].  `addend` (optional) is added to the result.
__global__ void __launch_bounds__(kThreads)
upsample_diag_kernel(const float* __restrict__ in, const float* __restrict__ g, int per_channel,
                     int N, int h, int w, int C, int k, int s, float* __restrict__ out) {
  const int H = h * s, W = w * s, pad = (k - s) / 2;
  const size_t total = static_cast<size_t>(N) * H * W * C;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    size_t t = idx / C;
    const int ox = static_cast<int>(t % W);
    t /= W;
    const int oy = static_cast<int>(t % H);
    const size_t img = t / H;
    float acc = 0.f;
    for (int a = 0; a * s < k; ++a) {
      const int ky = (oy + pad) % s + a * s;
      const int iy = (oy + pad - ky) / s;
      if (oy + pad - ky < 0 || iy >= h) continue;
      for (int b = 0; b * s < k; ++b) {
        const int kx = (ox + pad) % s + b * s;
        const int ix = (ox + pad - kx) / s;
        if (ox + pad - kx < 0 || ix >= w) continue;
        const float wgt = per_channel ? __ldg(g + (ky * k + kx) * C + c) : __ldg(g + ky * k + kx);
        acc = fmaf(wgt, __ldg(in + ((img * h + iy) * w + ix) * C + c), acc);
      }
    }
    out[idx] = acc;
  }
}

// transpose: din[n,iy,ix,c] = sum_{ky,kx} g(ky,kx,c) * dout[n, s*iy - pad + ky, s*ix - pad + kx, c]
__global__ void __launch_bounds__(kThreads)
upsample_diag_transpose_kernel(const float* __restrict__ dout, const float* __restrict__ g,
                               int per_channel, int N, int h, int w, int C, int k, int s,
                               float* __restrict__ din) {
  const int H = h * s, W = w * s, pad = (k - s) / 2;
  const size_t total = static_cast<size_t>(N) * h * w * C;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    size_t t = idx / C;
    const int ix = static_cast<int>(t % w);
    t /= w;
    const int iy = static_cast<int>(t % h);
    const size_t img = t / h;
    float acc = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int oy = s * iy - pad + ky;
      if (oy < 0 || oy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ox = s * ix - pad + kx;
        if (ox < 0 || ox >= W) continue;
        const float wgt = per_channel ? __ldg(g + (ky * k + kx) * C + c) : __ldg(g + ky * k + kx);
        acc = fmaf(wgt, __ldg(dout + ((img * H + oy) * W + ox) * C + c), acc);
      }
    }
    din[idx] = acc;
  }
}

__global__ void add_f32_inplace_kernel(float* __restrict__ a, const float* __restrict__ b,
                                       size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    a[i] += b[i];
}

}  // namespace

#define XV_BN_LAUNCHED()              \
  do {                                \
    XV_CUDA(cudaGetLastError());      \
    count_launch();                   \
    return 0;                         \
  } while (0)

int launch_bn_stats(const void* z, bool bf16, size_t npix, int C, double* sums, cudaStream_t s) {
  XV_CUDA(cudaMemsetAsync(sums, 0, 2 * C * sizeof(double), s));
  const int ct = C < kThreads ? C : kThreads;
  const int grid = grid_for(npix * ct, kThreads, 4);
  if (bf16)
    bn_stats_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(
        static_cast<const __nv_bfloat16*>(z), npix, C, sums);
  else
    bn_stats_kernel<float><<<grid, kThreads, 0, s>>>(static_cast<const float*>(z), npix, C, sums);
  XV_BN_LAUNCHED();
}

int launch_bn_finalize(const double* sums, size_t npix, int C, float eps, float momentum,
                       const float* bias_extra, float* mean, float* rstd, float* moving_mean,
                       float* moving_var, cudaStream_t s) {
  bn_finalize_kernel<<<div_up(C, 128), 128, 0, s>>>(sums, static_cast<double>(npix), C, eps,
                                                    momentum, bias_extra, mean, rstd, moving_mean,
                                                    moving_var);
  XV_BN_LAUNCHED();
}

int launch_bn_apply(const void* z, bool bf16, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, size_t npix, int C, int relu, void* y,
                    cudaStream_t s) {
  const size_t total = npix * C;
  if (bf16)
    bn_apply_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, s>>>(
        static_cast<const __nv_bfloat16*>(z), mean, rstd, gamma, beta, total, C, relu,
        static_cast<__nv_bfloat16*>(y));
  else
    bn_apply_kernel<float><<<grid_for(total), kThreads, 0, s>>>(
        static_cast<const float*>(z), mean, rstd, gamma, beta, total, C, relu,
        static_cast<float*>(y));
  XV_BN_LAUNCHED();
}

int launch_bn_backward(const void* g, const void* y_mask, const void* z, bool bf16,
                       const float* mean, const float* rstd, const float* gamma, size_t npix,
                       int C, double* sums, void* dz, __nv_bfloat16* dz_pad, int c_pad,
                       float* dgamma, float* dbeta, cudaStream_t s) {
  XV_CUDA(cudaMemsetAsync(sums, 0, 2 * C * sizeof(double), s));
  if (dz_pad) XV_CUDA(cudaMemsetAsync(dz_pad, 0, npix * c_pad * sizeof(__nv_bfloat16), s));
  const int ct = C < kThreads ? C : kThreads;
  const int rgrid = grid_for(npix * ct, kThreads, 4);
  const size_t total = npix * C;
  if (bf16) {
    using T = __nv_bfloat16;
    bn_bwd_reduce_kernel<T><<<rgrid, kThreads, 0, s>>>(
        static_cast<const T*>(g), static_cast<const T*>(y_mask), static_cast<const T*>(z), mean,
        rstd, npix, C, sums);
    XV_CUDA(cudaGetLastError());
    count_launch();
    bn_bwd_apply_kernel<T><<<grid_for(total), kThreads, 0, s>>>(
        static_cast<const T*>(g), static_cast<const T*>(y_mask), static_cast<const T*>(z), mean,
        rstd, gamma, sums, static_cast<double>(npix), npix, C, static_cast<T*>(dz), nullptr, 0,
        dgamma, dbeta);
  } else {
    bn_bwd_reduce_kernel<float><<<rgrid, kThreads, 0, s>>>(
        static_cast<const float*>(g), static_cast<const float*>(y_mask),
        static_cast<const float*>(z), mean, rstd, npix, C, sums);
    XV_CUDA(cudaGetLastError());
    count_launch();
    bn_bwd_apply_kernel<float><<<grid_for(total), kThreads, 0, s>>>(
        static_cast<const float*>(g), static_cast<const float*>(y_mask),
        static_cast<const float*>(z), mean, rstd, gamma, sums, static_cast<double>(npix), npix, C,
        static_cast<float*>(dz), dz_pad, c_pad, dgamma, dbeta);
  }
  XV_BN_LAUNCHED();
}

int launch_upsample_diag(const float* in, const float* g, bool per_channel, int N, int h, int w,
                         int C, int k, int stride, float* out, cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * h * stride * w * stride * C;
  upsample_diag_kernel<<<grid_for(total, kThreads, 16), kThreads, 0, s>>>(
      in, g, per_channel ? 1 : 0, N, h, w, C, k, stride, out);
  XV_BN_LAUNCHED();
}

int launch_upsample_diag_transpose(const float* dout, const float* g, bool per_channel, int N,
                                   int h, int w, int C, int k, int stride, float* din,
                                   cudaStream_t s) {
  const size_t total = static_cast<size_t>(N) * h * w * C;
  upsample_diag_transpose_kernel<<<grid_for(total, kThreads, 16), kThreads, 0, s>>>(
      dout, g, per_channel ? 1 : 0, N, h, w, C, k, stride, din);
  XV_BN_LAUNCHED();
}

int launch_add_f32_inplace(float* a, const float* b, size_t n, cudaStream_t s) {
  add_f32_inplace_kernel<<<grid_for(n), kThreads, 0, s>>>(a, b, n);
  XV_BN_LAUNCHED();
}

}  // namespace xv
