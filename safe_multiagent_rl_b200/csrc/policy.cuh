// Shared by the two builds of the fused per-agent discrete policy (policy.cu: FP32 pipes, policy_tc.cu: fc1 on the
// tensor cores): arguments, the Philox stream, and the softmax / inverse-CDF sampling head.
#pragma once

#include "common.cuh"

namespace smarl {

constexpr int kPolHidden = 16;
constexpr int kPolActions = 5;
constexpr float kLog2e = 1.4426950408889634f;

struct PolicyArgs {
  const uint8_t* pos_x;
  const uint8_t* pos_y;
  uint8_t* actions;
  float* logp;
  const float* w1;     // [A][2A][16]
  const float* b1;     // [A][16]
  const float* w2;     // [A][16][5]
  const float* b2;     // [A][5]
  uint64_t seed;
  int64_t env_offset;
  int64_t n_envs;
  int64_t ld;
  int64_t n_tiles;
  uint32_t t_word;     // t | episode << 16
  const uint32_t* episode_dev;
};

// tensor-core build (policy_tc.cu); SMARL_EUNSUPPORTED when the shape does not fit it
int launch_policy_tc(const PolicyArgs& a, int n_agents, int group_max, int sms, cudaStream_t st);

#ifdef __CUDACC__
__device__ __forceinline__ uint2 policy_key(uint64_t seed) {
  return make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x504C4359u);   // "PLCY"
}
// One Philox block serves four consecutive agents of one env and step: counter (env id lo, hi, t | episode << 16,
// agent >> 2), agent a uses word a & 3 (oracle/philox.policy_uniforms restates it).
__device__ __forceinline__ uint4 policy_words(uint64_t env_id, uint32_t t_word, int agent_quad, uint2 key) {
  return philox4x32_10(make_uint4((uint32_t)env_id, (uint32_t)(env_id >> 32), t_word, (uint32_t)agent_quad), key);
}
__device__ __forceinline__ uint32_t word_of(const uint4& o, int i) {
  return i == 0 ? o.x : (i == 1 ? o.y : (i == 2 ? o.z : o.w));
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// softmax (agent.py:35), Categorical sample by inverse CDF and log_prob (:44-46) of one agent and env.  The logits
// arrive in log2 units, l2_c = log2(e) * logit_c (both builds scale their shared-memory copies of fc2 by log2(e)), so
// e_c = exp(logit_c - max) is one subtraction and one MUFU.EX2:
// u = ((word >> 8) + 0.5) * 2^-24; action = #{c < 4 : sum_{c' <= c} e_c' <= u * sum e}.
__device__ __forceinline__ void policy_head(const float (&l2)[kPolActions], uint32_t word, int& pick, float& logp) {
  float m = l2[0];
#pragma unroll
  for (int c = 1; c < kPolActions; ++c) m = fmaxf(m, l2[c]);
  float ex[kPolActions], sum = 0.f;
#pragma unroll
  for (int c = 0; c < kPolActions; ++c) {
    ex[c] = ex2_ftz(l2[c] - m);
    sum += ex[c];
  }
  const float target = ((float)(word >> 8) + 0.5f) * (1.0f / 16777216.0f) * sum;
  float cum = 0.f, l_pick = l2[0];
  pick = 0;
#pragma unroll
  for (int c = 0; c < kPolActions - 1; ++c) {
    cum += ex[c];
    if (cum <= target) {
      pick = c + 1;
      l_pick = l2[c + 1];
    }
  }
  logp = ((l_pick - m) - lg2_ftz(sum)) * 0.693147180559945309f;     // log_softmax at the sampled action (sum >= 1)
}
#endif

}  // namespace smarl
