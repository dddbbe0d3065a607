// Fused tail of DirichletFusion (dirichlet_mix.py:96-136 behind two FCN experts): per output pixel
//   upsample x8 + bias + softmax of BOTH experts' 1/8-resolution class scores   (decoder tail,
//                                                       simple_fcn.py:129-133, basic_fusion_model.py:21)
//   -> Dirichlet fusion with the bit-exact argmax                (dirichlet_mix.py:14-36,100-113)
//   -> fused label and / or confusion-matrix accumulation        (base_model.py:140-151)
// in ONE pass: the two [N,H,W,C] probability tensors (2 x 14.2 MB written + 28.3 MB read per
// 768x384 frame at C = 12) never exist.  Every arithmetic step repeats the unfused kernels
// operation for operation (decode_upsample8_kernel's softmax, dirichlet_fuse_kernel's fast form and
// its exact re-evaluation), so labels are identical to the unfused path bit for bit.
#include "argmax.cuh"
#include "common.cuh"
#include "kernels.h"

namespace xv {

namespace {

constexpr int kMaxM = 2;     // the fused tail is built for the RGB-D pair
// bound constants: see fusion.cu (kLogErr, kAccErr, kTailErr) - same values
constexpr float kLogErr = 3.0e-5f;
constexpr float kAccErr = 3.6e-7f;
constexpr float kTailErr = 1.0e-6f;

struct DirSrc {
  const float* low[kMaxM];
  const float* bias[kMaxM];
  const float* g[kMaxM];
};

__device__ __forceinline__ float log_rn(float x) {
  return static_cast<float>(log(static_cast<double>(x)));
}

template <int C>
__device__ __forceinline__ int argmax_first(const float (&v)[C]) {
  int best = 0;
  float bv = v[0];
#pragma unroll
  for (int c = 1; c < C; ++c) {
    if (v[c] > bv) {
      bv = v[c];
      best = c;
    }
  }
  return best;
}

// exact-mode score of one pixel from its register probabilities (operation order of
// oracle.dirichlet_fusion_f32 / dirichlet_exact_pixel in fusion.cu)
template <int C>
__device__ __noinline__ void exact_pixel(const float (&p0)[C], const float (&p1)[C],
                                         const float* __restrict__ s_am1,
                                         const float* __restrict__ s_norm,
                                         const float* __restrict__ s_prior,
                                         float* __restrict__ total) {
  constexpr int CP = (C + 3) & ~3;
#pragma unroll 1
  for (int m = 0; m < 2; ++m) {
    float lx[C];
#pragma unroll
    for (int k = 0; k < C; ++k) lx[k] = m == 0 ? p0[k] : p1[k];
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

template <int C>
__global__ void __launch_bounds__(256)
decode_dirichlet_kernel(DirSrc src, const float* __restrict__ alpha_m1,
                        const float* __restrict__ lognorm, const float* __restrict__ logprior,
                        float exact_amax, float exact_tail, int N, int h, int w,
                        const int32_t* __restrict__ gt, unsigned long long* __restrict__ cm,
                        void* __restrict__ label_out, int label_bytes,
                        unsigned long long* __restrict__ n_exact) {
  constexpr int CP = (C + 3) & ~3;
  __shared__ __align__(16) float s_am1[kMaxM * C * CP];
  __shared__ float s_norm[kMaxM * C];
  __shared__ float s_prior[C];
  __shared__ float s_low[kMaxM][16 * C];
  __shared__ float s_g[kMaxM][256];
  __shared__ float s_bias[kMaxM][C];
  __shared__ unsigned int s_cm[C * C];
  for (int i = threadIdx.x; i < kMaxM * C * CP; i += 256) {
    const int c = i % CP, mk = i / CP;
    s_am1[i] = c < C ? alpha_m1[mk * C + c] : 0.f;
  }
  for (int i = threadIdx.x; i < kMaxM * C; i += 256) s_norm[i] = lognorm[i];
  for (int i = threadIdx.x; i < C; i += 256) s_prior[i] = logprior[i];
  for (int i = threadIdx.x; i < C * C; i += 256) s_cm[i] = 0u;
  for (int m = 0; m < kMaxM; ++m) {
    s_g[m][threadIdx.x] = src.g[m][threadIdx.x];
    if (threadIdx.x < C) s_bias[m][threadIdx.x] = src.bias[m][threadIdx.x];
  }
  const bool exact = exact_amax >= 0.f;
  const int H = 8 * h, W = 8 * w;
  const int tiles_x = (W + 15) / 16, tiles_y = (H + 15) / 16;
  const int total_tiles = tiles_x * tiles_y * N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lane = threadIdx.x & 31;
  unsigned int redo_count = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int bx = (tile % tiles_x) * 16;
    const int rest = tile / tiles_x;
    const int by = (rest % tiles_y) * 16;
    const int img = rest / tiles_y;
    const int iy0 = (by + 4) / 8 - 1, ix0 = (bx + 4) / 8 - 1;
    __syncthreads();
    for (int m = 0; m < kMaxM; ++m) {
      for (int i = threadIdx.x; i < 16 * C; i += 256) {
        const int c = i % C, cell = i / C;
        const int iy = iy0 + cell / 4, ix = ix0 + cell % 4;
        float v = 0.f;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w)
          v = __ldg(src.low[m] + ((static_cast<size_t>(img) * h + iy) * w + ix) * C + c);
        s_low[m][i] = v;
      }
    }
    __syncthreads();
    const int ox = bx + tx, oy = by + ty;
    const bool valid = ox < W && oy < H;
    int key = -1;
    if (valid) {
      const int ay = (oy + 4) >> 3, ry = (oy + 4) & 7;
      const int ax = (ox + 4) >> 3, rx = (ox + 4) & 7;
      float p[kMaxM][C];
      // ---- decode: same taps, fmaf chain, bias add, softmax as decode_upsample8_kernel
#pragma unroll
      for (int m = 0; m < kMaxM; ++m) {
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
            const float wgt = s_g[m][ky * 16 + kx];
            const float* lp = s_low[m] + ((iy - iy0) * 4 + (ix - ix0)) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) s[c] = fmaf(wgt, lp[c], s[c]);
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          s[c] += s_bias[m][c];
          mx = fmaxf(mx, s[c]);
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          s[c] = expf(s[c] - mx);
          sum += s[c];
        }
#pragma unroll
        for (int c = 0; c < C; ++c) p[m][c] = s[c] / sum;
      }
      // ---- Dirichlet fusion, fast form (dirichlet_fuse_kernel) + exact re-evaluation on near-ties
      float total[C];
      float lmax = 0.f;
#pragma unroll
      for (int m = 0; m < kMaxM; ++m) {
        float lx[C];
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) sum += p[m][k];
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 0; k < C; ++k) {
          lx[k] = fast_log_normal(1e-20f + p[m][k] * inv);
          lmax = fmaxf(lmax, fabsf(lx[k]));
        }
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
      int best = argmax_first<C>(total);
      if (exact) {
        float second = -INFINITY;
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (c != best) second = fmaxf(second, total[c]);
        const float bound = exact_amax * (kLogErr + kAccErr * static_cast<float>(C) * lmax) +
                            kTailErr * (exact_amax * lmax + exact_tail);
        if (!(total[best] - second > 2.f * bound)) {
          exact_pixel<C>(p[0], p[1], s_am1, s_norm, s_prior, total);
          best = argmax_first<C>(total);
          ++redo_count;
        }
      }
      const size_t pix = (static_cast<size_t>(img) * H + oy) * W + ox;
      if (label_out) {
        if (label_bytes == 8)
          reinterpret_cast<int64_t*>(label_out)[pix] = best;
        else
          reinterpret_cast<uint8_t*>(label_out)[pix] = static_cast<uint8_t>(best);
      }
      if (gt) {
        const int l = __ldg(gt + pix);
        if (l >= 0 && l < C) key = l * C + best;
      }
    }
    if (gt) {      // warp-aggregated shared-memory histogram (uniform branch: gt is a kernel argument)
      const int key0 = __shfl_sync(0xffffffffu, key, 0);
      if (__all_sync(0xffffffffu, key == key0)) {
        if (lane == 0 && key0 >= 0) atomicAdd(&s_cm[key0], 32u);
      } else {
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&s_cm[key], __popc(peers));
      }
    }
  }
  __syncthreads();
  if (cm) {
    for (int i = threadIdx.x; i < C * C; i += 256)
      if (s_cm[i]) atomicAdd(cm + i, static_cast<unsigned long long>(s_cm[i]));
  }
  if (n_exact != nullptr && redo_count)
    atomicAdd(n_exact, static_cast<unsigned long long>(redo_count));
}

template <int C>
int dispatch(const DirSrc& src, const float* alpha_m1, const float* lognorm, const float* logprior,
             float amax, float tail, int N, int h, int w, const int32_t* gt, long long* cm,
             void* label_out, int label_bytes, unsigned long long* n_exact, cudaStream_t s) {
  const long long tiles = static_cast<long long>(div_up(8 * w, 16)) * div_up(8 * h, 16) * N;
  const long long cap = static_cast<long long>(device_info().num_sms) * 4;
  const int grid = static_cast<int>(tiles < cap ? tiles : cap);
  decode_dirichlet_kernel<C><<<grid, 256, 0, s>>>(src, alpha_m1, lognorm, logprior, amax, tail, N,
                                                  h, w, gt,
                                                  reinterpret_cast<unsigned long long*>(cm),
                                                  label_out, label_bytes, n_exact);
  XV_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

int launch_decode_dirichlet(const float* const* low, const float* const* g,
                            const float* const* bias, int M, const float* alpha_m1,
                            const float* lognorm, const float* logprior, float exact_amax,
                            float exact_tail, int C, int N, int h, int w, const int32_t* gt,
                            long long* cm, void* label_out, int label_bytes,
                            unsigned long long* n_exact, cudaStream_t s) {
  XV_CHECK(M == kMaxM, "decode_dirichlet: built for two experts");
  XV_CHECK(C >= 2 && C <= 16, "decode_dirichlet: 2..16 classes");
  XV_CHECK((gt == nullptr) == (cm == nullptr), "decode_dirichlet: pass labels and matrix together");
  DirSrc src;
  for (int m = 0; m < kMaxM; ++m) {
    src.low[m] = low[m];
    src.g[m] = g[m];
    src.bias[m] = bias[m];
  }
  switch (C) {
#define XV_CASE(K) \
  case K:          \
    return dispatch<K>(src, alpha_m1, lognorm, logprior, exact_amax, exact_tail, N, h, w, gt, cm, \
                       label_out, label_bytes, n_exact, s);
    XV_CASE(2) XV_CASE(3) XV_CASE(4) XV_CASE(5) XV_CASE(6) XV_CASE(7) XV_CASE(8) XV_CASE(9)
    XV_CASE(10) XV_CASE(11) XV_CASE(12) XV_CASE(13) XV_CASE(14) XV_CASE(15) XV_CASE(16)
#undef XV_CASE
    default:
      return fail("decode_dirichlet: unsupported class count");
  }
}

}  // namespace xv
