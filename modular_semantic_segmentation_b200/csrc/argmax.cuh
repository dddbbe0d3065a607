// argmax(softmax(score)) without evaluating the softmax (basic_fusion_model.py:21-22).
#pragma once
#include <cuda_runtime.h>

namespace xv {

// Label of one pixel = first index of the maximum of p_c = fl(fl(expf(s_c - mx)) / sum), i.e.
// exactly what the probability-writing kernels return, computed from the scores alone:
//  * p is a non-decreasing function of s, so a class AFTER the first score maximum can never win
//    (ties go to the lower index);
//  * a class BEFORE it wins only if its probability rounds to the same float as the maximum's.
//    With e_best = expf(0) = 1 that needs expf(s_c - mx) >= 1 - 3 ulp; expf is within 2 ulp, so
//    s_c - mx >= -4.8e-7 is necessary.  Only then (practically never) is the softmax evaluated,
//    with the same operation order as the kernels that write probabilities.
template <int C>
__device__ __forceinline__ int argmax_of_softmax(const float (&s)[C]) {
  // one pass: `before` is the largest score in front of the running maximum (the value the
  // maximum had when it was last replaced), which is all the tie test needs
  float mx = s[0], before = -INFINITY;
  int best = 0;
#pragma unroll
  for (int c = 1; c < C; ++c) {
    const bool up = s[c] > mx;
    before = up ? mx : before;
    best = up ? c : best;
    mx = up ? s[c] : mx;
  }
  const bool close = before - mx >= -4.8e-7f;
  if (!close) return best;
  float e[C];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    e[c] = expf(s[c] - mx);
    sum += e[c];
  }
  int b = 0;
  float bv = -1.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float pr = e[c] / sum;
    if (pr > bv) {
      bv = pr;
      b = c;
    }
  }
  return b;
}

// Natural log of a NORMAL positive float on the MUFU: lg2.approx.ftz * ln 2.  Same values as
// __logf for normal inputs, without its subnormal rescue (five extra ALU instructions per call);
// the Dirichlet kernels only call it on 1e-20 + p with p >= 0.
__device__ __forceinline__ float fast_log_normal(float x) {
  float t;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t * 0.693147180559945309f;
}

}  // namespace xv
