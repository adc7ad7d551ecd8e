// Congestion device code shared by the step kernel and the fused rollout kernel.
#pragma once
#include "common.cuh"

namespace smarl {

// envs/congestion.py:7-10
#define SMARL_HOURLY_COMPENSATION 30.0
#define SMARL_AVERAGE_RIDE_COMPENSATION 7.5
#define SMARL_AVERAGE_RIDE_COST 4.0
#define SMARL_CONGESTION_COST 2.0

// Effective (post-noise) moves of four envs for every agent, from Philox4x32-10:
//   counter = (env_id lo, env_id hi, t, agent >> 2), key = seed; agent a reads word w = out[a & 3];
//   move = action if w < keep_threshold else w mod 5,
// the integer form of congestion.py:64-67 (u1 < 1 - noise ? a : int(u2 * 5)) with u1 = w * 2^-32 and
// u2 = (w mod 5 + 0.5) / 5.  Given w >= keep_threshold, w mod 5 is uniform on 0..4 up to one count in
// 2^32 per outcome -- the same resolution as drawing a second 32-bit word -- so one generator call serves
// four agents.
template <int A>
__device__ __forceinline__ void congestion_noise_moves(const uint32_t (&aw)[A], uint32_t (&mw)[A],
                                                       uint64_t seed, uint64_t keep_threshold,
                                                       int64_t env0, uint32_t t) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int i = 0; i < A; ++i) mw[i] = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t id = (uint64_t)(env0 + k);
#pragma unroll
    for (int j = 0; j < (A + 3) / 4; ++j) {
      const uint4 o = philox4x32_10(make_uint4((uint32_t)id, (uint32_t)(id >> 32), t, (uint32_t)j), key);
      const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (4 * j + q < A) {
          const uint32_t act = (aw[4 * j + q] >> (8 * k)) & 0xFFu;
          const uint32_t mv = ((uint64_t)w[q] < keep_threshold) ? act : w[q] % 5u;
          mw[4 * j + q] |= mv << (8 * k);
        }
      }
    }
  }
}

// Congestion._congestions (congestion.py:113-137) for env lane k of the packed words, in the
// closed form proved equivalent in tests/test_oracle_vs_reference.py: agents sharing the same
// directed edge form a class; with L = lowest index in the class whose INTENDED action is a
// move (<4), members >= L get (#members >= L) - 1, everything else 0.
//   xw,yw  new positions   dcw  displacement code (nx-ox+1) | (ny-oy+1) << 2   aw intended
// The edge (ox,oy,nx,ny) is identified by the 3-byte key (nx, ny, dcode).
// Writes byte k of conw[i]; returns the number of agents standing on node (0,0).
template <int A>
__device__ __forceinline__ int congestion_env(const uint32_t (&xw)[A], const uint32_t (&yw)[A],
                                              const uint32_t (&dcw)[A], const uint32_t (&aw)[A],
                                              uint32_t (&conw)[A], int k) {
  const uint32_t sel_xy = (uint32_t)k | ((uint32_t)(k + 4) << 4);          // (x_k, y_k, ., .)
  const uint32_t sel_key = 0x0010u | ((uint32_t)(k + 4) << 8) | (4u << 12); // (b0, b1, dcode_k, dc byte0)
  uint32_t key[A], same[A];
  uint32_t movers = 0u;
  int at_origin = 0;
#pragma unroll
  for (int i = 0; i < A; ++i) {
    const uint32_t xy = __byte_perm(xw[i], yw[i], sel_xy);
    key[i] = __byte_perm(xy, dcw[i], sel_key) & 0x00FFFFFFu;
    same[i] = 0u;
    movers |= (((aw[i] >> (8 * k)) & 0xFFu) < 4u) ? (1u << i) : 0u;
    at_origin += ((key[i] & 0xFFFFu) == 0u) ? 1 : 0;                       // congestion.py:97
  }
#pragma unroll
  for (int i = 0; i < A; ++i) {
#pragma unroll
    for (int j = i + 1; j < A; ++j) {
      // one compare + two predicated ORs (the select form costs two more instructions per pair)
      asm("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %2, %3;\n\t@q or.b32 %0, %0, %4;\n\t@q or.b32 %1, %1, %5;\n\t}"
          : "+r"(same[i]), "+r"(same[j])
          : "r"(key[i]), "r"(key[j]), "r"(1u << j), "r"(1u << i));
    }
  }
  uint32_t active = 0u;
#pragma unroll
  for (int i = 0; i < A; ++i) {
    const uint32_t below_or_self = (same[i] & ((1u << i) - 1u)) | (1u << i);
    active |= (below_or_self & movers) ? (1u << i) : 0u;
  }
#pragma unroll
  for (int i = 0; i < A; ++i) {
    const uint32_t cls = (same[i] | (1u << i)) & active;
    const uint32_t con = ((active >> i) & 1u) ? (uint32_t)__popc(cls) - 1u : 0u;
    conw[i] |= con << (8 * k);
  }
  return at_origin;
}

// Waiting-branch reward in f64, the reference's operation order (congestion.py:86-87).
__device__ __forceinline__ double congestion_reward_f64(uint32_t con, uint32_t nx, uint32_t ny,
                                                        const double* __restrict__ demand, int W) {
  const double d = __ldg(demand + nx * W + ny);
  const double q = __ddiv_rn(-SMARL_HOURLY_COMPENSATION * (double)(con + 1u), d);
  return __dadd_rn(__dadd_rn(q, SMARL_AVERAGE_RIDE_COMPENSATION), -SMARL_AVERAGE_RIDE_COST);
}

// Congestion.reward (congestion.py:82-87) for one agent, rounded to f32 exactly once.  The waiting
// branch comes from the host-built table wait[con][cell] when given (same f64 expression, same single
// rounding), otherwise it is evaluated here in f64 in the reference's operation order.
__device__ __forceinline__ float congestion_reward(uint32_t act, uint32_t con, uint32_t nx, uint32_t ny,
                                                   const double* __restrict__ demand, int W,
                                                   const float* __restrict__ wait) {
  if (act < 4u) return -4.0f - 2.0f * (float)con;                   // exact: small integers
  if (wait) return __ldg(wait + (con * W + nx) * W + ny);
  return (float)congestion_reward_f64(con, nx, ny, demand, W);
}

}  // namespace smarl
