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
#include "argmax.cuh"
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kPix = 256;   // threads per block, one pixel per thread and iteration
constexpr int kMaxM = 4;    // experts per fusion call

// Bound on |fast - exact| of one Dirichlet class score, used by the two-tier exact mode
// (DESIGN.md 4.3).  With A = max_c sum_{m,k} |alpha_m1[m][k][c]| (passed as exact_amax) and
// Lmax = max |log x| of the pixel:
//   kLogErr  per-logarithm difference: lg2.approx (rel. 2^-22 of |log2 x| <= 66.5, i.e. <= 1.1e-5
//            abs) + p*inv vs p/sum and sum order ((C+3) ulp of x) + the exact log's half ulp
//   kAccErr  per-term rounding of the two accumulation chains (FMA chain: 1 rounding per term,
//            product+sum chain: 2), relative to |a||L| <= |a| Lmax: 3 * 2^-24, doubled for slack
//   kTailErr roundings of  - lognorm, + next expert, + logprior  (identical operations in both
//            modes, each propagating <= 1 ulp of a partial result <= A Lmax + exact_tail, where
//            exact_tail = max_c sum_m |lognorm[m][c]| + max_c |logprior[c]|)
constexpr float kLogErr = 3.0e-5f;
constexpr float kAccErr = 3.6e-7f;
constexpr float kTailErr = 1.0e-6f;

struct PtrPack {
  const void* p[kMaxM];
};

inline int tiles_grid(int64_t npix) {
  int64_t blocks = div_up64(npix, kPix);
  int64_t cap = static_cast<int64_t>(device_info().num_sms) * 16;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// One pixel's C class values straight into registers with the widest aligned vector access
// (C % 4 == 0: 16 B, C % 2 == 0: 8 B).  A warp touches a contiguous 32*C*4-byte span; the
// partially used sectors of one instruction are completed by the next ones out of L1.
template <int C>
__device__ __forceinline__ void load_px(const float* __restrict__ g, int64_t pix, float (&v)[C]) {
  const float* p = g + pix * C;
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      v[4 * i] = t.x;
      v[4 * i + 1] = t.y;
      v[4 * i + 2] = t.z;
      v[4 * i + 3] = t.w;
    }
  } else if constexpr (C % 2 == 0) {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(p) + i);
      v[2 * i] = t.x;
      v[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) v[i] = __ldg(p + i);
  }
}

template <int C>
__device__ __forceinline__ void store_px(float* __restrict__ g, int64_t pix, const float (&v)[C]) {
  float* p = g + pix * C;
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (C % 2 == 0) {
#pragma unroll
    for (int i = 0; i < C / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) p[i] = v[i];
  }
}

// Odd C: rows are only 4-byte aligned, so a warp stages its 32 consecutive pixels through a
// private shared-memory slice with fully coalesced accesses; reading row `lane` back has
// stride C (odd, hence conflict-free).  `base` = first pixel of the warp, `cnt` = valid pixels.
template <int C>
struct WarpStage {
  static constexpr bool kUse = (C % 2) != 0;
  static constexpr int kFloats = kUse ? (kPix / 32) * 32 * C : 1;
};
// Full warps move their 32 * C floats (128 * C bytes, 16-byte aligned because the chunk starts at a
// multiple of 32 pixels) as float4 pieces; the ragged last chunk falls back to scalar accesses.
template <int C>
__device__ __forceinline__ void warp_load(const float* __restrict__ g, int64_t base, int cnt,
                                          float* __restrict__ slice, float (&v)[C]) {
  const int lane = threadIdx.x & 31;
  const float* p = g + base * C;
  __syncwarp();
  if (cnt == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    float4* s4 = reinterpret_cast<float4*>(slice);
#pragma unroll
    for (int i = 0; i < (8 * C + 31) / 32; ++i) {
      const int j = lane + 32 * i;
      if (j < 8 * C) s4[j] = __ldg(p4 + j);
    }
  } else {
    for (int i = lane; i < cnt * C; i += 32) slice[i] = __ldg(p + i);
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < C; ++k) v[k] = lane < cnt ? slice[lane * C + k] : 0.f;
}
template <int C>
__device__ __forceinline__ void warp_store(float* __restrict__ g, int64_t base, int cnt,
                                           float* __restrict__ slice, const float (&v)[C]) {
  const int lane = threadIdx.x & 31;
  float* p = g + base * C;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < C; ++k) slice[lane * C + k] = v[k];
  __syncwarp();
  if (cnt == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* s4 = reinterpret_cast<const float4*>(slice);
#pragma unroll
    for (int i = 0; i < (8 * C + 31) / 32; ++i) {
      const int j = lane + 32 * i;
      if (j < 8 * C) p4[j] = s4[j];
    }
  } else {
    for (int i = lane; i < cnt * C; i += 32) p[i] = slice[i];
  }
}

// Pixel access used by every kernel below: direct vector access for even C, warp staging for
// odd C.  The loops are written over warp-sized groups so both variants share one structure.
#define XV_WARP_LOOP(base, cnt, npix)                                                          \
  const int64_t _groups = ((npix) + 31) / 32;                                                  \
  const int64_t _gstride = static_cast<int64_t>(gridDim.x) * (kPix / 32);                      \
  for (int64_t _g = blockIdx.x * static_cast<int64_t>(kPix / 32) + (threadIdx.x >> 5),         \
               base = _g * 32;                                                                  \
       _g < _groups; _g += _gstride, base = _g * 32)                                           \
    if (const int cnt = static_cast<int>(((npix) - base) < 32 ? ((npix) - base) : 32); true)

template <int C>
__device__ __forceinline__ void px_load(const float* __restrict__ g, int64_t base, int cnt,
                                        float* slice, float (&v)[C]) {
  if constexpr (WarpStage<C>::kUse) {
    warp_load<C>(g, base, cnt, slice, v);
  } else {
    const int lane = threadIdx.x & 31;
    if (lane < cnt) load_px<C>(g, base + lane, v);
  }
}
template <int C>
__device__ __forceinline__ void px_store(float* __restrict__ g, int64_t base, int cnt, float* slice,
                                         const float (&v)[C]) {
  if constexpr (WarpStage<C>::kUse) {
    warp_store<C>(g, base, cnt, slice, v);
  } else {
    const int lane = threadIdx.x & 31;
    if (lane < cnt) store_px<C>(g, base + lane, v);
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

#define XV_PIXEL_LOOP(pix, npix)                                                        \
  for (int64_t pix = blockIdx.x * static_cast<int64_t>(kPix) + threadIdx.x; pix < (npix); \
       pix += static_cast<int64_t>(gridDim.x) * kPix)

// ------------------------------------------------------------------ softmax + argmax
template <int C>
__global__ void __launch_bounds__(kPix)
softmax_argmax_kernel(const float* __restrict__ score, int64_t npix, float* __restrict__ prob,
                      int64_t* __restrict__ label64, uint8_t* __restrict__ label8) {
  __shared__ __align__(16) float s_stage[WarpStage<C>::kFloats];
  float* slice = s_stage + (WarpStage<C>::kUse ? (threadIdx.x >> 5) * 32 * C : 0);
  const int lane = threadIdx.x & 31;
  XV_WARP_LOOP(base, cnt, npix) {
    float v[C];
    px_load<C>(score, base, cnt, slice, v);
    const bool live = lane < cnt;
    int best = 0;
    if (live && prob == nullptr) {
      // label only: the argmax of the softmax follows from the scores (argmax.cuh)
      best = argmax_of_softmax<C>(v);
    } else if (live) {
      float mx = v[0];
#pragma unroll
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, v[c]);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c] = expf(v[c] - mx);
        sum += v[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = v[c] / sum;
      best = argmax_first<C>(v);
    }
    if (live) {
      if (label64) label64[base + lane] = best;
      if (label8) label8[base + lane] = static_cast<uint8_t>(best);
    }
    if (prob) px_store<C>(prob, base, cnt, slice, v);
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
  XV_PIXEL_LOOP(pix, npix) {
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

// uint8 labels, 16 pixels (one 16-byte vector per expert) per thread and iteration
__global__ void __launch_bounds__(kPix)
bayes_lut_u8x16_kernel(PtrPack labels, int M, const int32_t* __restrict__ lut, int C, int lut_size,
                       int64_t nvec, uint4* __restrict__ out) {
  extern __shared__ int32_t s_lut[];
  for (int i = threadIdx.x; i < lut_size; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  XV_PIXEL_LOOP(vec, nvec) {
    uint32_t idx[16];
    bool ok[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      idx[j] = 0;
      ok[j] = true;
    }
    for (int m = 0; m < M; ++m) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(labels.p[m]) + vec);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t l = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
        ok[j] = ok[j] && (l < static_cast<uint32_t>(C));
        idx[j] = idx[j] * C + l;
      }
    }
    uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t f = ok[j] ? static_cast<uint32_t>(s_lut[idx[j]]) & 0xffu : 0u;
      o[j >> 2] |= f << (8 * (j & 3));
    }
    out[vec] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Literal form (bayes_mix.py:12-58): score[c] = sum_m logcond[m][l_m][c] + logprior[c].
template <int C>
__global__ void __launch_bounds__(kPix)
bayes_score_kernel(PtrPack labels, int M, int label_bytes, const float* __restrict__ logcond,
                   const float* __restrict__ logprior, int64_t npix, float* __restrict__ score,
                   void* __restrict__ label_out) {
  __shared__ float s_tab[kMaxM * C * C + C];
  for (int i = threadIdx.x; i < M * C * C; i += kPix) s_tab[i] = logcond[i];
  for (int i = threadIdx.x; i < C; i += kPix) s_tab[kMaxM * C * C + i] = logprior[i];
  __syncthreads();
  XV_PIXEL_LOOP(pix, npix) {
    float v[C];
    for (int m = 0; m < M; ++m) {
      int l = load_label(labels.p[m], label_bytes, pix);
      l = min(max(l, 0), C - 1);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float t = s_tab[(m * C + l) * C + c];
        v[c] = (m == 0) ? t : v[c] + t;   // stack + reduce_sum in expert order
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] += s_tab[kMaxM * C * C + c];
    if (label_out) store_label(label_out, label_bytes, pix, argmax_first<C>(v));
    if (score) store_px<C>(score, pix, v);
  }
}

// ------------------------------------------------------------------ Dirichlet fusion
// score[c] = sum_m ( sum_k am1[m][k][c] * log(1e-20 + p_m[k]/sum p_m) - lognorm[m][c] ) + logprior[c]
// The (alpha-1) tables sit in shared memory with rows padded to a multiple of 4 floats so one
// broadcast LDS.128 feeds four FMAs.
//
// Two arithmetic modes:
//   fast   one reciprocal per expert, MUFU lg2 logarithm, packed fused multiply-adds.
//   exact  the fixed float32 operation order of oracle.dirichlet_fusion_f32 (sequential sums,
//          IEEE division, correctly rounded logarithm, separately rounded products and sums) so
//          that the argmax is bit-exact against that oracle.  The kernel evaluates the fast
//          form first; only pixels whose two best fast scores are closer than a rigorous bound on
//          |fast - exact| are re-evaluated in the exact arithmetic (`exact_amax` >= 0 enables
//          this; see dirichlet_fast_bound).  With `score` requested every pixel takes the exact
//          path so that the returned scores are the exact ones too.

// Correctly rounded float32 logarithm (round-once from the float64 value).
__device__ __forceinline__ float log_rn(float x) { return static_cast<float>(log(static_cast<double>(x))); }

// Exact-mode score of one pixel; p_m read again from global memory (this path is rare).
template <int C>
__device__ __noinline__ void dirichlet_exact_pixel(const PtrPack& probs, int M, int64_t pix,
                                                   const float* __restrict__ s_am1,
                                                   const float* __restrict__ s_norm,
                                                   const float* __restrict__ s_prior,
                                                   float* __restrict__ total) {
  constexpr int CP = (C + 3) & ~3;
  for (int m = 0; m < M; ++m) {
    const float* p = reinterpret_cast<const float*>(probs.p[m]) + pix * C;
    float lx[C];
#pragma unroll
    for (int k = 0; k < C; ++k) lx[k] = __ldg(p + k);
    float sum = lx[0];
#pragma unroll
    for (int k = 1; k < C; ++k) sum = __fadd_rn(sum, lx[k]);
#pragma unroll
    for (int k = 0; k < C; ++k) lx[k] = log_rn(__fadd_rn(1e-20f, __fdiv_rn(lx[k], sum)));
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float acc = __fmul_rn(lx[0], s_am1[(m * C) * CP + c]);
#pragma unroll
      for (int k = 1; k < C; ++k) acc = __fadd_rn(acc, __fmul_rn(lx[k], s_am1[(m * C + k) * CP + c]));
      const float t = __fsub_rn(acc, s_norm[m * C + c]);
      total[c] = (m == 0) ? t : __fadd_rn(total[c], t);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) total[c] = __fadd_rn(total[c], s_prior[c]);
}

template <int C, bool EXACT>
__global__ void __launch_bounds__(kPix)
dirichlet_fuse_kernel(PtrPack probs, int M, const float* __restrict__ alpha_m1,
                      const float* __restrict__ lognorm, const float* __restrict__ logprior,
                      int64_t npix, float* __restrict__ score, void* __restrict__ label_out,
                      int label_bytes, float exact_amax, float exact_tail,
                      unsigned long long* __restrict__ n_exact) {
  constexpr int CP = (C + 3) & ~3;
  __shared__ __align__(16) float s_am1[kMaxM * C * CP];
  __shared__ float s_norm[kMaxM * C];
  __shared__ float s_prior[C];
  for (int i = threadIdx.x; i < M * C * CP; i += kPix) {
    const int c = i % CP, mk = i / CP;
    s_am1[i] = c < C ? alpha_m1[mk * C + c] : 0.f;
  }
  for (int i = threadIdx.x; i < M * C; i += kPix) s_norm[i] = lognorm[i];
  for (int i = threadIdx.x; i < C; i += kPix) s_prior[i] = logprior[i];
  __shared__ __align__(16) float s_stage[WarpStage<C>::kFloats];
  float* slice = s_stage + (WarpStage<C>::kUse ? (threadIdx.x >> 5) * 32 * C : 0);
  const int lane = threadIdx.x & 31;
  constexpr bool exact = EXACT;
  const bool exact_all = exact && score != nullptr;
  unsigned int exact_count = 0;
  __syncthreads();
  // The expert maps are read one (pixel group, expert) step ahead of the arithmetic: the kernel
  // has ~25 MUFU + ~300 FMA per pixel against 97 bytes, and without the rolling prefetch the
  // loads of a step only started once the previous step's logs and FMAs had retired (0.53 of the
  // copy bandwidth).
  const int64_t groups = (npix + 31) / 32;
  const int64_t gstride = static_cast<int64_t>(gridDim.x) * (kPix / 32);
  int64_t g = blockIdx.x * static_cast<int64_t>(kPix / 32) + (threadIdx.x >> 5);
  auto count_of = [&](int64_t b) { return static_cast<int>((npix - b) < 32 ? (npix - b) : 32); };
  float nx[C];
#pragma unroll
  for (int k = 0; k < C; ++k) nx[k] = 1.f;
  if (!exact_all && g < groups)
    px_load<C>(reinterpret_cast<const float*>(probs.p[0]), g * 32, count_of(g * 32), slice, nx);
  for (; g < groups; g += gstride) {
    const int64_t base = g * 32;
    const int cnt = count_of(base);
    const int64_t pix = base + lane;
    const bool live = lane < cnt;
    float total[C];
    float lmax = 0.f;               // largest |log| of the pixel: scales the rounding bound
    if (!exact_all) {
      for (int m = 0; m < M; ++m) {
        float lx[C];
#pragma unroll
        for (int k = 0; k < C; ++k) {
          lx[k] = nx[k];
          nx[k] = 1.f;
        }
        if (m + 1 < M) {
          px_load<C>(reinterpret_cast<const float*>(probs.p[m + 1]), base, cnt, slice, nx);
        } else if (g + gstride < groups) {
          const int64_t nb = (g + gstride) * 32;
          px_load<C>(reinterpret_cast<const float*>(probs.p[0]), nb, count_of(nb), slice, nx);
        }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) sum += lx[k];
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          lx[k] = fast_log_normal(1e-20f + lx[k] * inv);
          if (exact) lmax = fmaxf(lmax, fabsf(lx[k]));
        }
        // packed fp32 FMAs (fma.rn.f32x2, sm_100): one broadcast LDS.128 feeds two 2-wide FMAs
        unsigned long long ll2[CP / 2];
#pragma unroll
        for (int c = 0; c < CP / 2; ++c) ll2[c] = 0ull;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const ulonglong2* row = reinterpret_cast<const ulonglong2*>(s_am1 + (m * C + k) * CP);
          unsigned long long xx;
          asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(lx[k]));
#pragma unroll
          for (int c4 = 0; c4 < CP / 4; ++c4) {
            const ulonglong2 a = row[c4];
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ll2[2 * c4]) : "l"(xx), "l"(a.x));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ll2[2 * c4 + 1]) : "l"(xx), "l"(a.y));
          }
        }
        float ll[CP];
#pragma unroll
        for (int c = 0; c < CP / 2; ++c)
          asm("mov.b64 {%0, %1}, %2;" : "=f"(ll[2 * c]), "=f"(ll[2 * c + 1]) : "l"(ll2[c]));
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = ll[c] - s_norm[m * C + c];
          total[c] = (m == 0) ? t : total[c] + t;
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) total[c] += s_prior[c];
    }
    int best = 0;
    if (live) {
      bool redo = exact_all;
      if (!exact_all) {
        best = argmax_first<C>(total);
        if (exact) {
          // margin between the two best fast scores against 2 * bound on |fast - exact|
          float second = -INFINITY;
#pragma unroll
          for (int c = 0; c < C; ++c)
            if (c != best) second = fmaxf(second, total[c]);
          const float bound = exact_amax * (kLogErr + kAccErr * static_cast<float>(C) * lmax) +
                              kTailErr * (exact_amax * lmax + exact_tail);
          redo = !(total[best] - second > 2.f * bound);   // also catches NaN / inf
        }
      }
      if (redo) {
        dirichlet_exact_pixel<C>(probs, M, pix, s_am1, s_norm, s_prior, total);
        best = argmax_first<C>(total);
        ++exact_count;
      }
      if (label_out) store_label(label_out, label_bytes, pix, best);
    }
    if (score) px_store<C>(score, base, cnt, slice, total);
  }
  if (n_exact != nullptr && exact_count) atomicAdd(n_exact, static_cast<unsigned long long>(exact_count));
}

// ------------------------------------------------------------------ average / variance fusion
template <int C, bool VAR>
__global__ void __launch_bounds__(kPix)
mean_fuse_kernel(PtrPack probs, PtrPack vars, int M, int64_t npix, float* __restrict__ score,
                 void* __restrict__ label_out, int label_bytes) {
  __shared__ __align__(16) float s_stage[WarpStage<C>::kFloats];
  float* slice = s_stage + (WarpStage<C>::kUse ? (threadIdx.x >> 5) * 32 * C : 0);
  const int lane = threadIdx.x & 31;
  XV_WARP_LOOP(base, cnt, npix) {
    const int64_t pix = base + lane;
    const bool live = lane < cnt;
    float acc[C];
    float csum = 0.f;
    for (int m = 0; m < M; ++m) {
      float v[C];
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = 0.f;
      px_load<C>(reinterpret_cast<const float*>(probs.p[m]), base, cnt, slice, v);
      float wgt = 1.f;
      // every product, sum and quotient is rounded on its own (no FMA contraction, IEEE
      // division): bit-exact against the float32 numpy statement of the rule
      if (VAR) {   // certainty = 1 / (1e-20 + variance), variance_mix.py:11
        wgt = __fdiv_rn(1.f, __fadd_rn(1e-20f, live ? __ldg(reinterpret_cast<const float*>(
                                                                 vars.p[m]) + pix)
                                                    : 1.f));
        csum = (m == 0) ? wgt : __fadd_rn(csum, wgt);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float t = VAR ? __fmul_rn(wgt, v[c]) : v[c];
        acc[c] = (m == 0) ? t : __fadd_rn(acc[c], t);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
      acc[c] = VAR ? __fdiv_rn(acc[c], csum) : __fdiv_rn(acc[c], static_cast<float>(M));
    if (label_out && live) store_label(label_out, label_bytes, pix, argmax_first<C>(acc));
    if (score) px_store<C>(score, base, cnt, slice, acc);
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
  const float inv_logc = 1.f / logf(static_cast<float>(C));
  const bool want_ce = cond_entropy != nullptr;
  __shared__ __align__(16) float s_stage[WarpStage<C>::kFloats];
  float* slice = s_stage + (WarpStage<C>::kUse ? (threadIdx.x >> 5) * 32 * C : 0);
  const int lane = threadIdx.x & 31;
  XV_WARP_LOOP(base, cnt, npix) {
    const int64_t pix = base + lane;
    const bool live = lane < cnt;
    float mu[C], m2[C];
    float ce = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) mu[c] = m2[c] = 0.f;
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
      float x[C];
#pragma unroll
      for (int c = 0; c < C; ++c) x[c] = 0.f;
      px_load<C>(samples + static_cast<int64_t>(t) * npix * C, base, cnt, slice, x);
      const float inv_n = 1.f / static_cast<float>(t + 1);
      float h = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float d = x[c] - mu[c];
        mu[c] += d * inv_n;
        m2[c] = fmaf(d, x[c] - mu[c], m2[c]);
        if (want_ce) h -= x[c] * logf(fminf(fmaxf(x[c], 1e-10f), 1.f));
      }
      ce += h * inv_logc;
    }
    float sv = 0.f, h = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      m2[c] = m2[c] / static_cast<float>(T);
      sv += m2[c];
      if (entropy) h -= mu[c] * logf(fminf(fmaxf(mu[c], 1e-10f), 1.f));
    }
    if (live) {
      if (mean_var) mean_var[pix] = sv / static_cast<float>(C);
      if (sum_var) sum_var[pix] = sv;
      if (entropy) entropy[pix] = h * inv_logc;
      if (cond_entropy) cond_entropy[pix] = ce / static_cast<float>(T);
    }
    if (mean) px_store<C>(mean, base, cnt, slice, mu);
    if (var) px_store<C>(var, base, cnt, slice, m2);
  }
}

// ------------------------------------------------------------------ Dirichlet sufficient statistics
// S[c][k] += sum_{label==c} log(1e-10 + prob[k]);  n[c] += #{label==c}
// = onehot(label)^T . log(prob), a [C x P] . [P x C] product with P = pixels.  A block stages the
// logarithms of 256 pixels in shared memory; thread (g, c, k4) then owns the four statistics
// S[c][4*k4 .. 4*k4+3] for every G-th pixel of the tile (G = 256 / (C * ceil(C/4)) groups share
// the tile) and adds a pixel's four logarithms under the predicate label == c: no atomics, no
// shuffles, no divergence - the cost does not depend on how the labels are distributed (the
// warp-aggregated version was 6x slower on random maps than on blocky ones).  Per-thread float64
// accumulators persist over the block's tiles and are added to the global sums once.
template <int C>
__global__ void __launch_bounds__(kPix)
suffstats_kernel(const float* __restrict__ prob, const int32_t* __restrict__ labels, int64_t npix,
                 double* __restrict__ S, unsigned long long* __restrict__ n) {
  constexpr int CP4 = (C + 3) / 4;           // float4 pieces per pixel row
  constexpr int CP = 4 * CP4;
  constexpr int kTeam = C * CP4;             // threads that cover all (c, k4) pairs
  constexpr int G = kPix / kTeam > 0 ? kPix / kTeam : 1;
  static_assert(kTeam <= kPix, "too many classes for one block");
  __shared__ __align__(16) float s_lp[kPix * CP];
  __shared__ int s_lab[kPix];
  __shared__ unsigned int s_n[C];
  // odd C > 16: the staging slices would not fit next to s_lp; those rows are read with scalar
  // loads instead (this kernel runs once per fit(), odd wide class sets are not worth more)
  constexpr bool kStaged = WarpStage<C>::kUse && C <= 16;
  __shared__ __align__(16) float s_stage[kStaged ? WarpStage<C>::kFloats : 1];
  float* slice = s_stage + (kStaged ? (threadIdx.x >> 5) * 32 * C : 0);
  for (int i = threadIdx.x; i < C; i += kPix) s_n[i] = 0u;
  const int lane = threadIdx.x & 31;
  const int team = threadIdx.x / kTeam;                 // pixel group of this thread (< G: active)
  const int c_own = (threadIdx.x % kTeam) / CP4, k4 = (threadIdx.x % kTeam) % CP4;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t tiles = (npix + kPix - 1) / kPix;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // phase 1: logarithms + labels of the tile's pixels -> shared memory
    const int64_t base = tile * kPix + (threadIdx.x >> 5) * 32;
    const int cnt = static_cast<int>(npix - base < 32 ? (npix - base > 0 ? npix - base : 0) : 32);
    float lp[C];
#pragma unroll
    for (int k = 0; k < C; ++k) lp[k] = 1.f;
    if constexpr (WarpStage<C>::kUse && !kStaged) {
      if (lane < cnt) {
#pragma unroll
        for (int k = 0; k < C; ++k) lp[k] = __ldg(prob + (base + lane) * C + k);
      }
    } else {
      px_load<C>(prob, base, cnt, slice, lp);
    }
    int label = -1;
    if (lane < cnt) {
      label = __ldg(labels + base + lane);
      if (label < 0 || label >= C) label = -1;
    }
    __syncthreads();                                    // previous tile fully consumed
    s_lab[threadIdx.x] = label;
#pragma unroll
    for (int q = 0; q < CP4; ++q) {
      float4 v;
      v.x = 4 * q < C ? logf(1e-10f + lp[4 * q < C ? 4 * q : 0]) : 0.f;
      v.y = 4 * q + 1 < C ? logf(1e-10f + lp[4 * q + 1 < C ? 4 * q + 1 : 0]) : 0.f;
      v.z = 4 * q + 2 < C ? logf(1e-10f + lp[4 * q + 2 < C ? 4 * q + 2 : 0]) : 0.f;
      v.w = 4 * q + 3 < C ? logf(1e-10f + lp[4 * q + 3 < C ? 4 * q + 3 : 0]) : 0.f;
      reinterpret_cast<float4*>(s_lp + threadIdx.x * CP)[q] = v;
    }
    if (label >= 0) atomicAdd(&s_n[label], 1u);
    __syncthreads();
    // phase 2: predicated accumulation of this thread's (class, 4 statistics) over its pixels
    if (team < G) {
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      for (int p = team; p < kPix; p += G) {
        const float4 v = reinterpret_cast<const float4*>(s_lp + p * CP)[k4];
        const bool hit = s_lab[p] == c_own;
        part[0] += hit ? v.x : 0.f;
        part[1] += hit ? v.y : 0.f;
        part[2] += hit ? v.z : 0.f;
        part[3] += hit ? v.w : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += static_cast<double>(part[j]);
    }
  }
  // the G teams of the block are summed in shared memory first: every block ends with only C*C
  // global atomics on the same addresses (and few blocks are launched)
  __syncthreads();
  double* s_S = reinterpret_cast<double*>(s_lp);          // the tile buffer is free now
  for (int i = threadIdx.x; i < C * C; i += kPix) s_S[i] = 0.0;
  __syncthreads();
  if (team < G) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * k4 + j < C && acc[j] != 0.0) atomicAdd(&s_S[c_own * C + 4 * k4 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += kPix)
    if (s_S[i] != 0.0) atomicAdd(S + i, s_S[i]);
  for (int i = threadIdx.x; i < C; i += kPix)
    if (s_n[i]) atomicAdd(n + i, static_cast<unsigned long long>(s_n[i]));
}

// ------------------------------------------------------------------ confusion matrix
// cm[label][pred] += 1 for 0 <= label < C (negative labels = ignore, base_model.py:140-143).
// Label maps are spatially coherent: a warp whose 32 pixels share one (label, pred) pair adds
// 32 with a single shared-memory atomic; mixed warps aggregate equal keys with match_any.
__global__ void __launch_bounds__(1024)
confusion_kernel(const void* __restrict__ pred, int pred_bytes, const int32_t* __restrict__ labels,
                 int64_t npix, int C, unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned int s_cm[];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_cm[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // each thread owns 4 consecutive pixels (16-byte label load); uniform trip count per warp
  const int64_t nquad = (npix + 3) / 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t start = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t iters = (nquad + stride - 1) / stride;
  const bool vec = (npix % 4 == 0);
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t quad = start + it * stride;
    int key[4] = {-1, -1, -1, -1};
    if (quad < nquad) {
      int l[4], pr[4];
      const int64_t p0 = quad * 4;
      if (vec) {
        const int4 lv = __ldg(reinterpret_cast<const int4*>(labels) + quad);
        l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w;
        if (pred_bytes == 1) {
          const uchar4 pv = __ldg(reinterpret_cast<const uchar4*>(pred) + quad);
          pr[0] = pv.x; pr[1] = pv.y; pr[2] = pv.z; pr[3] = pv.w;
        } else {
          const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(pred) + 2 * quad);
          const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(pred) + 2 * quad + 1);
          pr[0] = static_cast<int>(a.x); pr[1] = static_cast<int>(a.y);
          pr[2] = static_cast<int>(b.x); pr[3] = static_cast<int>(b.y);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool in = p0 + j < npix;
          l[j] = in ? __ldg(labels + p0 + j) : -1;
          pr[j] = in ? load_label(pred, pred_bytes, p0 + j) : 0;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (l[j] >= 0 && l[j] < C && pr[j] >= 0 && pr[j] < C) key[j] = l[j] * C + pr[j];
    }
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
  bool vec_ok = label_bytes == 1 && npix % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int m = 0; m < M; ++m) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(labels[m]) & 15) == 0;
  if (vec_ok)
    bayes_lut_u8x16_kernel<<<tiles_grid(npix / 16), kPix, lut_size * sizeof(int32_t), s>>>(
        pk, M, lut, C, lut_size, npix / 16, reinterpret_cast<uint4*>(out));
  else
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
                          float* score, void* label_out, int label_bytes, float exact_amax,
                          float exact_tail, unsigned long long* n_exact, cudaStream_t s) {
  PtrPack pk;
  XV_TRY(pack_ptrs(reinterpret_cast<const void* const*>(probs), M, &pk));
  if (exact_amax >= 0.f) {
    XV_DISPATCH_C(C, (dirichlet_fuse_kernel<kC, true><<<tiles_grid(npix), kPix, 0, s>>>(
                         pk, M, alpha_m1, lognorm, logprior, npix, score, label_out, label_bytes,
                         exact_amax, exact_tail, n_exact)));
  } else {
    XV_DISPATCH_C(C, (dirichlet_fuse_kernel<kC, false><<<tiles_grid(npix), kPix, 0, s>>>(
                         pk, M, alpha_m1, lognorm, logprior, npix, score, label_out, label_bytes,
                         exact_amax, exact_tail, n_exact)));
  }
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
  // few blocks: every block ends with C*C float64 global atomics on the same addresses
  int64_t sblocks = div_up64(npix, kPix);
  const int64_t scap = static_cast<int64_t>(device_info().num_sms) * 4;
  sblocks = sblocks < scap ? (sblocks > 0 ? sblocks : 1) : scap;
  XV_DISPATCH_C(C, (suffstats_kernel<kC><<<static_cast<int>(sblocks), kPix, 0, s>>>(
                       prob, labels, npix, S, reinterpret_cast<unsigned long long*>(n))));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_confusion(const void* pred, int pred_bytes, const int32_t* labels, int64_t npix, int C,
                     long long* cm, cudaStream_t s) {
  XV_CHECK(pred_bytes == 8 || pred_bytes == 1, "pred_bytes must be 8 (int64) or 1 (uint8)");
  XV_CHECK(C >= 1 && C * C * 4 <= 48 * 1024, "confusion: too many classes");
  // few, fat blocks: every block ends with C*C global atomics on the same addresses
  const int64_t quads = (npix + 3) / 4;
  int64_t blocks = div_up64(quads, 1024);
  const int64_t cap = static_cast<int64_t>(device_info().num_sms) * 2;
  blocks = blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
  confusion_kernel<<<static_cast<int>(blocks), 1024, C * C * sizeof(unsigned int), s>>>(
      pred, pred_bytes, labels, npix, C, reinterpret_cast<unsigned long long*>(cm));
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace xv
