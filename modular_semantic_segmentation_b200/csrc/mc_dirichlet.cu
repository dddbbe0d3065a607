// Per-pixel Dirichlet kernels over Monte-Carlo dropout samples.
//
//  * dirichlet_fit_samples: maximum-likelihood Dirichlet fit of the T sample probability vectors
//    of every pixel - moment-matching initialisation (dirichlet_fastfit.py:376-380) followed by
//    Minka's fixed-point iteration a <- ipsi(psi(sum a) + mean_t log p) (dirichlet_fastfit.py:
//    188-204, inverse digamma :382-395).  The reference runs this on the host in float64 for one
//    [T,C] matrix; here one thread owns one pixel (fp32, checked against the float64 oracle).
//  * dirichlet_uncertainty_fuse: uncertainty-mixed Dirichlet fusion of
//    uncertainty_dirichlet_mix.py:18-52 - per pixel the class-conditional parameters are blended
//    with the "uninformative" table I+1 by mix = mean_c var / max var, so the normaliser
//    lbeta(alpha) has to be evaluated per pixel.
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kThreads = 128;

// digamma / trigamma for x > 0: upward recurrence to x >= 6, then the asymptotic series
__device__ __forceinline__ float digammaf(float x) {
  float acc = 0.f;
#pragma unroll 1
  while (x < 6.f) {
    acc -= 1.f / x;
    x += 1.f;
  }
  const float r = 1.f / x, r2 = r * r;
  return acc + logf(x) - 0.5f * r -
         r2 * (1.f / 12.f - r2 * (1.f / 120.f - r2 * (1.f / 252.f - r2 * (1.f / 240.f))));
}
__device__ __forceinline__ float trigammaf(float x) {
  float acc = 0.f;
#pragma unroll 1
  while (x < 6.f) {
    acc += 1.f / (x * x);
    x += 1.f;
  }
  const float r = 1.f / x, r2 = r * r;
  return acc + r + 0.5f * r2 +
         r * r2 * (1.f / 6.f - r2 * (1.f / 30.f - r2 * (1.f / 42.f - r2 * (1.f / 30.f))));
}
// inverse digamma by Newton iterations, dirichlet_fastfit.py:382-395
__device__ __forceinline__ float inv_digammaf(float y) {
  const float euler = 0.5772156649015329f;
  float x = y >= -2.22f ? expf(y) + 0.5f : -1.f / (y + euler);
#pragma unroll 1
  for (int i = 0; i < 6; ++i) {
    const float step = (digammaf(x) - y) / trigammaf(x);
    x -= step;
    if (fabsf(step) < 1e-6f * x) break;
  }
  return x;
}

template <int C>
__device__ __forceinline__ float dirichlet_loglik(const float (&a)[C], const float (&logp)[C],
                                                  float n) {
  float sum = 0.f, acc = 0.f;
#pragma unroll
  for (int k = 0; k < C; ++k) {
    sum += a[k];
    acc += (a[k] - 1.f) * logp[k] - lgammaf(a[k]);
  }
  return n * (lgammaf(sum) + acc);
}

template <int C>
__global__ void __launch_bounds__(kThreads)
dirichlet_fit_samples_kernel(const float* __restrict__ samples, int T, int64_t npix, float tol,
                             int maxiter, float* __restrict__ alpha, int* __restrict__ iters_out) {
  for (int64_t pix = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; pix < npix;
       pix += static_cast<int64_t>(gridDim.x) * kThreads) {
    float e[C], logp[C];
    float e2_0 = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) e[k] = logp[k] = 0.f;
    for (int t = 0; t < T; ++t) {
      const float* p = samples + (static_cast<int64_t>(t) * npix + pix) * C;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float v = __ldg(p + k);
        e[k] += v;
        logp[k] += logf(v);
        if (k == 0) e2_0 = fmaf(v, v, e2_0);
      }
    }
    const float inv_t = 1.f / static_cast<float>(T);
    e2_0 *= inv_t;
    float a[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
      e[k] *= inv_t;
      logp[k] *= inv_t;
    }
    // moment matching on the first component (dirichlet_fastfit.py:376-380)
    const float s0 = (e[0] - e2_0) / (e2_0 - e[0] * e[0]);
#pragma unroll
    for (int k = 0; k < C; ++k) a[k] = s0 * e[k];
    float ll_old = dirichlet_loglik<C>(a, logp, static_cast<float>(T));
    int it = 0;
    for (; it < maxiter; ++it) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) sum += a[k];
      const float psi_sum = digammaf(sum);
#pragma unroll
      for (int k = 0; k < C; ++k) a[k] = inv_digammaf(psi_sum + logp[k]);
      const float ll_new = dirichlet_loglik<C>(a, logp, static_cast<float>(T));
      const bool done = fabsf(ll_new - ll_old) < tol;
      ll_old = ll_new;
      if (done) {
        ++it;
        break;
      }
    }
#pragma unroll
    for (int k = 0; k < C; ++k) alpha[pix * C + k] = a[k];
    if (iters_out) iters_out[pix] = it;
  }
}

struct Ptrs4 {
  const float* p[4];
};

// score[c] = sum_m [ sum_k (a_mkc - 1) log(1e-20 + p_mk) - lbeta(a_m.c) ] + logprior[c],
// a_mkc = cond[m][k][c] (1 - mix_m) + mix_m (delta_kc + 1), mix_m = mean_k var_mk / max var_m
template <int C>
__global__ void __launch_bounds__(kThreads)
dirichlet_uncertainty_fuse_kernel(Ptrs4 probs, Ptrs4 vars, Ptrs4 max_var, int M,
                                  const float* __restrict__ cond, const float* __restrict__ logprior,
                                  int64_t npix, float* __restrict__ score,
                                  void* __restrict__ label_out, int label_bytes) {
  __shared__ float s_cond[4 * C * C];
  __shared__ float s_prior[C];
  for (int i = threadIdx.x; i < M * C * C; i += kThreads) s_cond[i] = cond[i];
  for (int i = threadIdx.x; i < C; i += kThreads) s_prior[i] = logprior[i];
  __syncthreads();
  for (int64_t pix = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; pix < npix;
       pix += static_cast<int64_t>(gridDim.x) * kThreads) {
    float total[C];
#pragma unroll
    for (int c = 0; c < C; ++c) total[c] = 0.f;
    for (int m = 0; m < M; ++m) {
      float lx[C];
      float vmean = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        lx[k] = logf(1e-20f + __ldg(probs.p[m] + pix * C + k));
        vmean += __ldg(vars.p[m] + pix * C + k);
      }
      const float mix = (vmean / static_cast<float>(C)) / __ldg(max_var.p[m]);
      for (int c = 0; c < C; ++c) {
        float asum = 0.f, acc = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const float a = s_cond[(m * C + k) * C + c] * (1.f - mix) + mix * (k == c ? 2.f : 1.f);
          asum += a;
          acc += (a - 1.f) * lx[k] - lgammaf(a);
        }
        total[c] += acc + lgammaf(asum);
      }
    }
    int best = 0;
    float bv = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      total[c] += s_prior[c];
      if (score) score[pix * C + c] = total[c];
      if (total[c] > bv) {
        bv = total[c];
        best = c;
      }
    }
    if (label_out) {
      if (label_bytes == 8)
        reinterpret_cast<int64_t*>(label_out)[pix] = best;
      else
        reinterpret_cast<uint8_t*>(label_out)[pix] = static_cast<uint8_t>(best);
    }
  }
}

// global maximum of a float array (values >= 0): warp shuffle + one atomicMax on the bits
__global__ void reduce_max_kernel(const float* __restrict__ x, int64_t n, unsigned int* out) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, __ldg(x + i));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // monotone for m >= 0
}

inline int grid_px(int64_t npix) {
  int64_t blocks = div_up64(npix, kThreads);
  const int64_t cap = static_cast<int64_t>(device_info().num_sms) * 16;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
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

}  // namespace

int launch_dirichlet_fit_samples(const float* samples, int T, int64_t npix, int C, float tol,
                                 int maxiter, float* alpha, int* iters, cudaStream_t s) {
  XV_CHECK(T >= 2, "dirichlet_fit_samples: need at least two samples");
  XV_DISPATCH_C(C, (dirichlet_fit_samples_kernel<kC><<<grid_px(npix), kThreads, 0, s>>>(
                       samples, T, npix, tol, maxiter, alpha, iters)));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_dirichlet_uncertainty_fuse(const float* const* probs, const float* const* vars,
                                      const float* const* max_var, int M, const float* cond,
                                      const float* logprior, int C, int64_t npix, float* score,
                                      void* label_out, int label_bytes, cudaStream_t s) {
  XV_CHECK(M >= 1 && M <= 4, "number of experts must be in [1, 4]");
  Ptrs4 pp, pv, pm;
  for (int m = 0; m < 4; ++m) {
    pp.p[m] = m < M ? probs[m] : nullptr;
    pv.p[m] = m < M ? vars[m] : nullptr;
    pm.p[m] = m < M ? max_var[m] : nullptr;
  }
  XV_DISPATCH_C(C, (dirichlet_uncertainty_fuse_kernel<kC><<<grid_px(npix), kThreads, 0, s>>>(
                       pp, pv, pm, M, cond, logprior, npix, score, label_out, label_bytes)));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_reduce_max(const float* x, int64_t n, float* out, cudaStream_t s) {
  XV_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
  int64_t blocks = div_up64(n, 1024);
  const int64_t cap = static_cast<int64_t>(device_info().num_sms) * 4;
  blocks = blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
  reduce_max_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(x, n,
                                                              reinterpret_cast<unsigned int*>(out));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace xv
