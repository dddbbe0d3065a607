// Per-pixel fusion + score kernels: HBM-streaming, one pixel per thread, class tables staged in
// shared memory, coalesced 16-byte global traffic through padded shared-memory tiles.
//
// Reference statements (paths relative to the reference repo):
//   softmax/argmax      xview/models/basic_fusion_model.py:21-22
//   Bayes fusion        xview/models/bayes_mix.py:12-58 (score) and :61-112 (decision LUT)
//   Dirichlet fusion    xview/models/dirichlet_mix.py:14-36, :100-113
//   average fusion      xview/models/average_mix.py:18-21
//   variance fusion     xview/models/variance_mix.py:7-15
//   MC moments          xview/models/variance_mix.py:62-66, bayesian_fcn.py:48-57
//   sufficient stats    xview/models/dirichlet_mix.py:142-163
//   confusion matrix    xview/models/base_model.py:140-151
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kPix = 256;   // pixels per block-tile == threads per block
constexpr int kMaxM = 4;    // experts per fusion call

struct PtrPack {
  const void* p[kMaxM];
};

inline int tiles_grid(int64_t npix) {
  int64_t tiles = div_up64(npix, kPix);
  int64_t cap = static_cast<int64_t>(device_info().num_sms) * 8;
  return static_cast<int>(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

template <int C>
struct Tile {
  static constexpr int CP = C | 1;   // odd row pitch -> conflict-free thread-per-pixel access
  static constexpr int kFloats = kPix * CP;
};

// global [pix0 .. pix0+cnt) x C  ->  smem rows of pitch CP (coalesced float4 reads)
template <int C>
__device__ __forceinline__ void tile_load(const float* __restrict__ g, int64_t pix0, int cnt,
                                          float* __restrict__ s) {
  constexpr int CP = Tile<C>::CP;
  const float* base = g + pix0 * C;          // pix0 % 256 == 0 -> 16-byte aligned
  const int n = cnt * C;
  const int n4 = n >> 2;
  const float4* b4 = reinterpret_cast<const float4*>(base);
  for (int i = threadIdx.x; i < n4; i += kPix) {
    const float4 v = __ldg(b4 + i);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    int e = i * 4;
    int p = e / C, k = e - p * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[p * CP + k] = vv[j];
      if (++k == C) {
        k = 0;
        ++p;
      }
    }
  }
  for (int e = n4 * 4 + threadIdx.x; e < n; e += kPix) {
    const int p = e / C, k = e - p * C;
    s[p * CP + k] = __ldg(base + e);
  }
}

// smem rows of pitch CP -> global [pix0 .. pix0+cnt) x C (coalesced float4 writes)
template <int C>
__device__ __forceinline__ void tile_store(float* __restrict__ g, int64_t pix0, int cnt,
                                           const float* __restrict__ s) {
  constexpr int CP = Tile<C>::CP;
  float* base = g + pix0 * C;
  const int n = cnt * C;
  const int n4 = n >> 2;
  float4* b4 = reinterpret_cast<float4*>(base);
  for (int i = threadIdx.x; i < n4; i += kPix) {
    float vv[4];
    int e = i * 4;
    int p = e / C, k = e - p * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      vv[j] = s[p * CP + k];
      if (++k == C) {
        k = 0;
        ++p;
      }
    }
    b4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  for (int e = n4 * 4 + threadIdx.x; e < n; e += kPix) {
    const int p = e / C, k = e - p * C;
    base[e] = s[p * CP + k];
  }
}

template <int C>
__device__ __forceinline__ int argmax_first(const float (&v)[C]) {
  int best = 0;
  float bv = v[0];
#pragma unroll
  for (int c = 1; c < C; ++c) {
    if (v[c] > bv) {   // strict: first maximal index wins (tf.argmax / np.argmax)
      bv = v[c];
      best = c;
    }
  }
  return best;
}

__device__ __forceinline__ void store_label(void* out, int label_bytes, int64_t pix, int label) {
  if (label_bytes == 8)
    reinterpret_cast<int64_t*>(out)[pix] = label;
  else
    reinterpret_cast<uint8_t*>(out)[pix] = static_cast<uint8_t>(label);
}
__device__ __forceinline__ int load_label(const void* in, int label_bytes, int64_t pix) {
  return label_bytes == 8 ? static_cast<int>(reinterpret_cast<const int64_t*>(in)[pix])
                          : static_cast<int>(reinterpret_cast<const uint8_t*>(in)[pix]);
}

// ------------------------------------------------------------------ softmax + argmax
template <int C>
__global__ void __launch_bounds__(kPix)
softmax_argmax_kernel(const float* __restrict__ score, int64_t npix, float* __restrict__ prob,
                      int64_t* __restrict__ label64, uint8_t* __restrict__ label8) {
  __shared__ float s[Tile<C>::kFloats];
  constexpr int CP = Tile<C>::CP;
  const int64_t tiles = (npix + kPix - 1) / kPix;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t pix0 = tile * kPix;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPix), npix - pix0));
    __syncthreads();
    tile_load<C>(score, pix0, cnt, s);
    __syncthreads();
    if (threadIdx.x < cnt) {
      float v[C];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c] = s[threadIdx.x * CP + c];
        mx = fmaxf(mx, v[c]);
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c] = expf(v[c] - mx);
        sum += v[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c] = v[c] / sum;
        s[threadIdx.x * CP + c] = v[c];
      }
      const int best = argmax_first<C>(v);
      if (label64) label64[pix0 + threadIdx.x] = best;
      if (label8) label8[pix0 + threadIdx.x] = static_cast<uint8_t>(best);
    }
    if (prob) {
      __syncthreads();
      tile_store<C>(prob, pix0, cnt, s);
    }
  }
}

// ------------------------------------------------------------------ Bayes fusion
// Decision-table form (bayes_mix.py:61-112): fused = LUT[l_0][l_1]...; pure integer lookups.
__global__ void __launch_bounds__(kPix)
bayes_lut_kernel(PtrPack labels, int M, int label_bytes, const int32_t* __restrict__ lut, int C,
                 int lut_size, int64_t npix, void* __restrict__ out) {
  extern __shared__ int32_t s_lut[];
  for (int i = threadIdx.x; i < lut_size; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  for (int64_t pix = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; pix < npix;
       pix += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int idx = 0;
    bool ok = true;
    for (int m = 0; m < M; ++m) {
      const int l = load_label(labels.p[m], label_bytes, pix);
      ok = ok && (l >= 0 && l < C);
      idx = idx * C + l;
    }
    store_label(out, label_bytes, pix, ok ? s_lut[idx] : 0);
  }
}

// Literal form (bayes_mix.py:12-58): score[c] = sum_m logcond[m][l_m][c] + logprior[c].
template <int C>
__global__ void __launch_bounds__(kPix)
bayes_score_kernel(PtrPack labels, int M, int label_bytes, const float* __restrict__ logcond,
                   const float* __restrict__ logprior, int64_t npix, float* __restrict__ score,
                   void* __restrict__ label_out) {
  __shared__ float s_tab[kMaxM * C * C + C];
  __shared__ float s[Tile<C>::kFloats];
  constexpr int CP = Tile<C>::CP;
  for (int i = threadIdx.x; i < M * C * C; i += kPix) s_tab[i] = logcond[i];
  for (int i = threadIdx.x; i < C; i += kPix) s_tab[kMaxM * C * C + i] = logprior[i];
  const int64_t tiles = (npix + kPix - 1) / kPix;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t pix0 = tile * kPix;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPix), npix - pix0));
    __syncthreads();
    if (threadIdx.x < cnt) {
      float v[C];
      for (int m = 0; m < M; ++m) {
        int l = load_label(labels.p[m], label_bytes, pix0 + threadIdx.x);
        l = min(max(l, 0), C - 1);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = s_tab[(m * C + l) * C + c];
          v[c] = (m == 0) ? t : v[c] + t;   // stack + reduce_sum in expert order
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c] += s_tab[kMaxM * C * C + c];
        s[threadIdx.x * CP + c] = v[c];
      }
      if (label_out) store_label(label_out, label_bytes, pix0 + threadIdx.x, argmax_first<C>(v));
    }
    if (score) {
      __syncthreads();
      tile_store<C>(score, pix0, cnt, s);
    }
  }
}

// ------------------------------------------------------------------ Dirichlet fusion
// score[c] = sum_m ( sum_k am1[m][k][c] * log(1e-20 + p_m[k]/sum p_m) - lognorm[m][c] ) + logprior[c]
template <int C>
__global__ void __launch_bounds__(kPix)
dirichlet_fuse_kernel(PtrPack probs, int M, const float* __restrict__ alpha_m1,
                      const float* __restrict__ lognorm, const float* __restrict__ logprior,
                      int64_t npix, float* __restrict__ score, void* __restrict__ label_out,
                      int label_bytes) {
  __shared__ float s_am1[kMaxM * C * C];
  __shared__ float s_norm[kMaxM * C];
  __shared__ float s_prior[C];
  __shared__ float s[Tile<C>::kFloats];
  constexpr int CP = Tile<C>::CP;
  for (int i = threadIdx.x; i < M * C * C; i += kPix) s_am1[i] = alpha_m1[i];
  for (int i = threadIdx.x; i < M * C; i += kPix) s_norm[i] = lognorm[i];
  for (int i = threadIdx.x; i < C; i += kPix) s_prior[i] = logprior[i];
  const int64_t tiles = (npix + kPix - 1) / kPix;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t pix0 = tile * kPix;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPix), npix - pix0));
    float total[C];
    for (int m = 0; m < M; ++m) {
      __syncthreads();
      tile_load<C>(reinterpret_cast<const float*>(probs.p[m]), pix0, cnt, s);
      __syncthreads();
      if (threadIdx.x < cnt) {
        float lx[C];
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          lx[k] = s[threadIdx.x * CP + k];
          sum += lx[k];
        }
#pragma unroll
        for (int k = 0; k < C; ++k) lx[k] = logf(1e-20f + lx[k] / sum);
        float ll[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ll[c] = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const float* row = s_am1 + (m * C + k) * C;
#pragma unroll
          for (int c = 0; c < C; ++c) ll[c] = fmaf(lx[k], row[c], ll[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = ll[c] - s_norm[m * C + c];
          total[c] = (m == 0) ? t : total[c] + t;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        total[c] += s_prior[c];
        s[threadIdx.x * CP + c] = total[c];
      }
      if (label_out)
        store_label(label_out, label_bytes, pix0 + threadIdx.x, argmax_first<C>(total));
    }
    if (score) {
      __syncthreads();
      tile_store<C>(score, pix0, cnt, s);
    }
  }
}

// ------------------------------------------------------------------ average / variance fusion
template <int C, bool VAR>
__global__ void __launch_bounds__(kPix)
mean_fuse_kernel(PtrPack probs, PtrPack vars, int M, int64_t npix, float* __restrict__ score,
                 void* __restrict__ label_out, int label_bytes) {
  __shared__ float s[Tile<C>::kFloats];
  constexpr int CP = Tile<C>::CP;
  const int64_t tiles = (npix + kPix - 1) / kPix;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t pix0 = tile * kPix;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPix), npix - pix0));
    float acc[C];
    float csum = 0.f;
    for (int m = 0; m < M; ++m) {
      __syncthreads();
      tile_load<C>(reinterpret_cast<const float*>(probs.p[m]), pix0, cnt, s);
      __syncthreads();
      if (threadIdx.x < cnt) {
        float wgt = 1.f;
        if (VAR) {   // certainty = 1 / (1e-20 + variance), variance_mix.py:11
          wgt = 1.f / (1e-20f + __ldg(reinterpret_cast<const float*>(vars.p[m]) + pix0 +
                                      threadIdx.x));
          csum = (m == 0) ? wgt : csum + wgt;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = wgt * s[threadIdx.x * CP + c];
          acc[c] = (m == 0) ? t : acc[c] + t;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        acc[c] = VAR ? acc[c] / csum : acc[c] / static_cast<float>(M);
        s[threadIdx.x * CP + c] = acc[c];
      }
      if (label_out) store_label(label_out, label_bytes, pix0 + threadIdx.x, argmax_first<C>(acc));
    }
    if (score) {
      __syncthreads();
      tile_store<C>(score, pix0, cnt, s);
    }
  }
}

// ------------------------------------------------------------------ MC-dropout moments
// samples [T, npix, C] -> population mean/var per class, mean-over-classes variance,
// normed entropy of the mean, mean normed entropy of the samples, sum-over-classes variance.
template <int C>
__global__ void __launch_bounds__(kPix)
mc_moments_kernel(const float* __restrict__ samples, int T, int64_t npix, float* __restrict__ mean,
                  float* __restrict__ var, float* __restrict__ mean_var,
                  float* __restrict__ entropy, float* __restrict__ cond_entropy,
                  float* __restrict__ sum_var) {
  __shared__ float s[Tile<C>::kFloats];
  constexpr int CP = Tile<C>::CP;
  const float inv_logc = 1.f / logf(static_cast<float>(C));
  const int64_t tiles = (npix + kPix - 1) / kPix;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t pix0 = tile * kPix;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPix), npix - pix0));
    float mu[C], m2[C];
    float ce = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) mu[c] = m2[c] = 0.f;
    for (int t = 0; t < T; ++t) {
      __syncthreads();
      tile_load<C>(samples + static_cast<int64_t>(t) * npix * C, pix0, cnt, s);
      __syncthreads();
      if (threadIdx.x < cnt) {
        const float inv_n = 1.f / static_cast<float>(t + 1);
        float h = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float x = s[threadIdx.x * CP + c];
          const float d = x - mu[c];
          mu[c] += d * inv_n;
          m2[c] = fmaf(d, x - mu[c], m2[c]);
          h -= x * logf(fminf(fmaxf(x, 1e-10f), 1.f));
        }
        ce += h * inv_logc;
      }
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
      float sv = 0.f, h = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        m2[c] = m2[c] / static_cast<float>(T);
        sv += m2[c];
        h -= mu[c] * logf(fminf(fmaxf(mu[c], 1e-10f), 1.f));
      }
      const int64_t pix = pix0 + threadIdx.x;
      if (mean_var) mean_var[pix] = sv / static_cast<float>(C);
      if (sum_var) sum_var[pix] = sv;
      if (entropy) entropy[pix] = h * inv_logc;
      if (cond_entropy) cond_entropy[pix] = ce / static_cast<float>(T);
    }
    if (mean) {
      if (threadIdx.x < cnt) {
#pragma unroll
        for (int c = 0; c < C; ++c) s[threadIdx.x * CP + c] = mu[c];
      }
      __syncthreads();
      tile_store<C>(mean, pix0, cnt, s);
    }
    if (var) {
      __syncthreads();
      if (threadIdx.x < cnt) {
#pragma unroll
        for (int c = 0; c < C; ++c) s[threadIdx.x * CP + c] = m2[c];
      }
      __syncthreads();
      tile_store<C>(var, pix0, cnt, s);
    }
  }
}

// ------------------------------------------------------------------ Dirichlet sufficient statistics
// S[c][k] += sum_{label==c} log(1e-10 + prob[k]);  n[c] += #{label==c}
template <int C>
__global__ void __launch_bounds__(kPix)
suffstats_kernel(const float* __restrict__ prob, const int32_t* __restrict__ labels, int64_t npix,
                 double* __restrict__ S, unsigned long long* __restrict__ n) {
  __shared__ float s[Tile<C>::kFloats];
  __shared__ double s_S[C * C];
  __shared__ unsigned int s_n[C];
  constexpr int CP = Tile<C>::CP;
  for (int i = threadIdx.x; i < C * C; i += kPix) s_S[i] = 0.0;
  for (int i = threadIdx.x; i < C; i += kPix) s_n[i] = 0u;
  const int lane = threadIdx.x & 31;
  const int64_t tiles = (npix + kPix - 1) / kPix;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t pix0 = tile * kPix;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kPix), npix - pix0));
    __syncthreads();
    tile_load<C>(prob, pix0, cnt, s);
    __syncthreads();
    int label = -1;
    float lp[C];
    if (threadIdx.x < cnt) {
      label = __ldg(labels + pix0 + threadIdx.x);
      if (label < 0 || label >= C) label = -1;
#pragma unroll
      for (int k = 0; k < C; ++k) lp[k] = logf(1e-10f + s[threadIdx.x * CP + k]);
    } else {
#pragma unroll
      for (int k = 0; k < C; ++k) lp[k] = 0.f;
    }
    // one pass per distinct class present in the warp: shuffle-reduce, one smem add per (c,k)
    unsigned remaining = __ballot_sync(0xffffffffu, label >= 0);
    while (remaining) {
      const int leader = __ffs(remaining) - 1;
      const int c = __shfl_sync(0xffffffffu, label, leader);
      const unsigned same = __ballot_sync(0xffffffffu, label == c);
#pragma unroll
      for (int k = 0; k < C; ++k) {
        float v = (label == c) ? lp[k] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) atomicAdd(&s_S[c * C + k], static_cast<double>(v));
      }
      if (lane == 0) atomicAdd(&s_n[c], __popc(same));
      remaining &= ~same;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += kPix) atomicAdd(S + i, s_S[i]);
  for (int i = threadIdx.x; i < C; i += kPix)
    atomicAdd(n + i, static_cast<unsigned long long>(s_n[i]));
}

// ------------------------------------------------------------------ confusion matrix
// cm[label][pred] += 1 for 0 <= label < C (negative labels = ignore, base_model.py:140-143)
__global__ void __launch_bounds__(kPix)
confusion_kernel(const void* __restrict__ pred, int pred_bytes, const int32_t* __restrict__ labels,
                 int64_t npix, int C, unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned int s_cm[];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_cm[i] = 0u;
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t start = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  // uniform trip count so the full-mask match below is legal
  const int64_t iters = (npix + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t pix = start + it * stride;
    int key = -1;
    if (pix < npix) {
      const int l = __ldg(labels + pix);
      const int pr = load_label(pred, pred_bytes, pix);
      if (l >= 0 && l < C && pr >= 0 && pr < C) key = l * C + pr;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && (__ffs(peers) - 1) == (threadIdx.x & 31))
      atomicAdd(&s_cm[key], __popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (s_cm[i]) atomicAdd(cm + i, static_cast<unsigned long long>(s_cm[i]));
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

int pack_ptrs(const void* const* src, int M, PtrPack* dst) {
  XV_CHECK(M >= 1 && M <= kMaxM, "number of experts must be in [1, 4]");
  for (int m = 0; m < kMaxM; ++m) dst->p[m] = m < M ? src[m] : nullptr;
  return 0;
}

}  // namespace

int launch_softmax_argmax(const float* score, int64_t npix, int C, float* prob, int64_t* label64,
                          uint8_t* label8, cudaStream_t s) {
  XV_DISPATCH_C(C, (softmax_argmax_kernel<kC><<<tiles_grid(npix), kPix, 0, s>>>(
                       score, npix, prob, label64, label8)));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_bayes_lut(const void* const* labels, int M, int label_bytes, const int32_t* lut, int C,
                     int64_t npix, void* out, cudaStream_t s) {
  PtrPack pk;
  XV_TRY(pack_ptrs(labels, M, &pk));
  XV_CHECK(label_bytes == 8 || label_bytes == 1, "label_bytes must be 8 (int64) or 1 (uint8)");
  int lut_size = 1;
  for (int m = 0; m < M; ++m) lut_size *= C;
  XV_CHECK(lut_size * 4 <= 48 * 1024, "decision table too large for shared memory");
  bayes_lut_kernel<<<tiles_grid(npix), kPix, lut_size * sizeof(int32_t), s>>>(
      pk, M, label_bytes, lut, C, lut_size, npix, out);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_bayes_score(const void* const* labels, int M, int label_bytes, const float* logcond,
                       const float* logprior, int C, int64_t npix, float* score, void* label_out,
                       cudaStream_t s) {
  PtrPack pk;
  XV_TRY(pack_ptrs(labels, M, &pk));
  XV_DISPATCH_C(C, (bayes_score_kernel<kC><<<tiles_grid(npix), kPix, 0, s>>>(
                       pk, M, label_bytes, logcond, logprior, npix, score, label_out)));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_dirichlet_fuse(const float* const* probs, int M, const float* alpha_m1,
                          const float* lognorm, const float* logprior, int C, int64_t npix,
                          float* score, void* label_out, int label_bytes, cudaStream_t s) {
  PtrPack pk;
  XV_TRY(pack_ptrs(reinterpret_cast<const void* const*>(probs), M, &pk));
  XV_DISPATCH_C(C, (dirichlet_fuse_kernel<kC><<<tiles_grid(npix), kPix, 0, s>>>(
                       pk, M, alpha_m1, lognorm, logprior, npix, score, label_out, label_bytes)));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_average_fuse(const float* const* probs, int M, int C, int64_t npix, float* score,
                        void* label_out, int label_bytes, cudaStream_t s) {
  PtrPack pk, none;
  XV_TRY(pack_ptrs(reinterpret_cast<const void* const*>(probs), M, &pk));
  none = pk;
  XV_DISPATCH_C(C, (mean_fuse_kernel<kC, false><<<tiles_grid(npix), kPix, 0, s>>>(
                       pk, none, M, npix, score, label_out, label_bytes)));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_variance_fuse(const float* const* probs, const float* const* vars, int M, int C,
                         int64_t npix, float* score, void* label_out, int label_bytes,
                         cudaStream_t s) {
  PtrPack pk, vk;
  XV_TRY(pack_ptrs(reinterpret_cast<const void* const*>(probs), M, &pk));
  XV_TRY(pack_ptrs(reinterpret_cast<const void* const*>(vars), M, &vk));
  XV_DISPATCH_C(C, (mean_fuse_kernel<kC, true><<<tiles_grid(npix), kPix, 0, s>>>(
                       pk, vk, M, npix, score, label_out, label_bytes)));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_mc_moments(const float* samples, int T, int64_t npix, int C, float* mean, float* var,
                      float* mean_var, float* entropy, float* cond_entropy, float* sum_var,
                      cudaStream_t s) {
  XV_CHECK(T >= 1, "mc_moments: need at least one sample");
  XV_DISPATCH_C(C, (mc_moments_kernel<kC><<<tiles_grid(npix), kPix, 0, s>>>(
                       samples, T, npix, mean, var, mean_var, entropy, cond_entropy, sum_var)));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_suffstats(const float* prob, const int32_t* labels, int64_t npix, int C, double* S,
                     long long* n, cudaStream_t s) {
  XV_DISPATCH_C(C, (suffstats_kernel<kC><<<tiles_grid(npix), kPix, 0, s>>>(
                       prob, labels, npix, S, reinterpret_cast<unsigned long long*>(n))));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_confusion(const void* pred, int pred_bytes, const int32_t* labels, int64_t npix, int C,
                     long long* cm, cudaStream_t s) {
  XV_CHECK(pred_bytes == 8 || pred_bytes == 1, "pred_bytes must be 8 (int64) or 1 (uint8)");
  XV_CHECK(C >= 1 && C * C * 4 <= 48 * 1024, "confusion: too many classes");
  confusion_kernel<<<tiles_grid(npix), kPix, C * C * sizeof(unsigned int), s>>>(
      pred, pred_bytes, labels, npix, C, reinterpret_cast<unsigned long long*>(cm));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace xv
