// Cross-env statistics feeding MetaAgent.update (safe_multi_agent_RL/meta_agent.py:32-36) and
// Buffer.mean_score (buffer.py:45-48).  Layout of a stats vector for (A agents, K constraints):
//   [0,K) sum_e C_k   [K,2K) #{e: C_k > thr_k}   [2K,2K+A) sum_e R_a   [2K+A,2K+2A) sum_e modR_a
//   [2K+2A] episode count
// Producers write one row of block partials per CTA; launch_stats_finalize reduces the rows in
// a fixed order (bit-reproducible for a given launch shape, no float atomics).
#pragma once
#include "common.cuh"

namespace smarl {

__host__ __device__ constexpr int stats_len(int A, int K) { return 2 * K + 2 * A + 1; }

int launch_stats_finalize(const double* partials, int64_t n_rows, int n_agents, int n_constraints,
                          int64_t n_envs, double* stats, cudaStream_t stream);

#ifdef __CUDACC__
// Sum over the CTA; the result is valid on thread 0 only.  s_red: THREADS/32 doubles.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* s_red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) r += s_red[w];
  }
  __syncthreads();
  return r;
}
#endif

}  // namespace smarl
