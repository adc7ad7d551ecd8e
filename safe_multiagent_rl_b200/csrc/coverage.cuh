// CoverageDiscrete device code shared by the step kernel and the fused rollout kernel.
#pragma once
#include "common.cuh"

namespace smarl {

// Largest penalty table kept in shared memory (entries).  fv^2 <= this, i.e. fieldview <= 110.
constexpr int kCoverageMaxLut = 12287;

// Sum of pair penalties of ONE env whose agents are packed as p[a] = x | y << 8.
//
// Replaces the i<j double loop over scipy's distance_matrix in CoverageContinuous.reward
// (envs/coverage.py:76-83).  For integer coordinates the distance is sqrt(q) with
// q = dx^2 + dy^2 an integer, so the penalty is a table lookup: |dx|,|dy| for the pair come
// from one SIMD byte abs-diff, q from one dp4a, and q is clamped to lut_len, whose entry is 0.
// Canonical summation order (every kernel that produces rewards uses this function, so
// step-mode and fused-mode rewards are bit-identical): pair n in the reference's i-major
// order goes to accumulator n % 4, result = (acc0 + acc1) + (acc2 + acc3), all f32.
template <int A>
__device__ __forceinline__ float coverage_pair_penalty(const uint32_t (&p)[A],
                                                       const float* __restrict__ s_lut,
                                                       uint32_t lut_len) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int n = 0;
#pragma unroll
  for (int i = 0; i < A; ++i) {
#pragma unroll
    for (int j = i + 1; j < A; ++j) {
      const uint32_t v = __vabsdiffu4(p[i], p[j]);
      const uint32_t q = min(__dp4a(v, v, 0u), lut_len);
      acc[n & 3] += s_lut[q];
      ++n;
    }
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// Stage the penalty table: s_lut[0..lut_len) = lut, s_lut[lut_len] = 0.
__device__ __forceinline__ void coverage_load_lut(float* s_lut, const float* __restrict__ lut,
                                                  int lut_len) {
  for (int i = threadIdx.x; i <= lut_len; i += blockDim.x) s_lut[i] = i < lut_len ? __ldg(lut + i) : 0.f;
  __syncthreads();
}

}  // namespace smarl
