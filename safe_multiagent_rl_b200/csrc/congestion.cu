// Congestion on sm_100a: step kernel (+ fused rollout).  Same thread mapping and SoA layout as
// coverage.cu: one thread owns four consecutive envs, one byte per env in every state word.
#include "congestion.cuh"
#include "stats.cuh"

namespace smarl {

struct CongestionStepArgs {
  uint8_t* pos_x;
  uint8_t* pos_y;
  const uint8_t* actions;
  uint8_t* moves;
  float* obs;
  float* reward;
  int32_t* cost;
  uint8_t* done;
  const double* lambdas;
  float* penalty;
  const double* demand;
  uint64_t keep_threshold;
  uint64_t seed;
  int64_t env_offset;
  int64_t n_groups;
  int64_t ld;
  int32_t size;
  int32_t t;
};

constexpr int kCongThreads = 128;

// One Congestion transition for four envs (congestion.py:49-75): applies the effective moves,
// leaves the new positions in xw/yw and the displacement codes in dcw.
template <int A>
__device__ __forceinline__ void congestion_transition(uint32_t (&xw)[A], uint32_t (&yw)[A],
                                                      const uint32_t (&mw)[A], uint32_t (&dcw)[A],
                                                      uint32_t size4) {
  const uint32_t k1 = 0x01010101u;
#pragma unroll
  for (int i = 0; i < A; ++i) {
    const uint32_t ox = xw[i], oy = yw[i];
    grid_move4(xw[i], yw[i], mw[i], size4);
    dcw[i] = ((xw[i] + k1) - ox) | (((yw[i] + k1) - oy) << 2);
  }
}

template <int A, int MODE>
__global__ void __launch_bounds__(kCongThreads) congestion_step_kernel(const CongestionStepArgs a) {
  const int64_t g = (int64_t)blockIdx.x * kCongThreads + threadIdx.x;
  if (g >= a.n_groups) return;
  const int64_t e0 = g * 4;
  const int64_t ld = a.ld;

  uint32_t xw[A], yw[A], aw[A], mw[A];
#pragma unroll
  for (int i = 0; i < A; ++i) {
    xw[i] = ld_stream_u32(a.pos_x + i * ld + e0);
    yw[i] = ld_stream_u32(a.pos_y + i * ld + e0);
    aw[i] = ld_stream_u32(a.actions + i * ld + e0);
    if (MODE == 1) mw[i] = ld_stream_u32(a.moves + i * ld + e0);
    if (MODE == 0) mw[i] = aw[i];
  }
  if (MODE == 2) congestion_noise_moves<A>(aw, mw, a.seed, a.keep_threshold, a.env_offset + e0, (uint32_t)a.t);

  uint32_t dcw[A];
  congestion_transition<A>(xw, yw, mw, dcw, (uint32_t)a.size * 0x01010101u);
#pragma unroll
  for (int i = 0; i < A; ++i) {
    st_stream_u32(a.pos_x + i * ld + e0, xw[i]);
    st_stream_u32(a.pos_y + i * ld + e0, yw[i]);
    if (MODE != 1 && a.moves) st_stream_u32(a.moves + i * ld + e0, mw[i]);
    if (a.done) st_stream_u32(a.done + i * ld + e0, 0u);                 // congestion.py:103-104
    if (a.obs) {
      st_stream_f4(a.obs + (2 * i) * ld + e0, bytes_to_float4(xw[i]));
      st_stream_f4(a.obs + (2 * i + 1) * ld + e0, bytes_to_float4(yw[i]));
    }
  }

  uint32_t conw[A];
#pragma unroll
  for (int i = 0; i < A; ++i) conw[i] = 0u;
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int at_origin = congestion_env<A>(xw, yw, dcw, aw, conw, k);
    const int c = max(0, A / 3 - at_origin);                             // congestion.py:93-100
    c0 = k == 0 ? c : c0;
    c1 = k == 1 ? c : c1;
    c2 = k == 2 ? c : c2;
    c3 = k == 3 ? c : c3;
  }
  st_stream_i4(a.cost + e0, make_int4(c0, c1, c2, c3));
  if (a.penalty) {                                                       // meta_agent.py:21-22
    const double lam = __ldg(a.lambdas);
    st_stream_f4(a.penalty + e0, make_float4((float)(lam * c0), (float)(lam * c1), (float)(lam * c2),
                                             (float)(lam * c3)));
  }
  const int W = a.size + 1;
#pragma unroll
  for (int i = 0; i < A; ++i) {
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      r[k] = (float)congestion_reward((aw[i] >> (8 * k)) & 0xFFu, (conw[i] >> (8 * k)) & 0xFFu,
                                      (xw[i] >> (8 * k)) & 0xFFu, (yw[i] >> (8 * k)) & 0xFFu, a.demand, W);
    st_stream_f4(a.reward + i * ld + e0, make_float4(r[0], r[1], r[2], r[3]));
  }
}

static int check_congestion(const SmarlCongestionParams* p) {
  SMARL_REQUIRE(p != nullptr, "params is NULL");
  SMARL_REQUIRE(p->size >= 1 && p->size <= 254, "size=%d outside 1..254", p->size);
  SMARL_REQUIRE(p->demand != nullptr, "demand table is NULL");
  SMARL_REQUIRE(p->noise_mode >= 0 && p->noise_mode <= 2, "bad noise_mode %d", p->noise_mode);
  SMARL_REQUIRE(p->keep_threshold <= (1ull << 32), "keep_threshold must be <= 2^32");
  return SMARL_OK;
}

}  // namespace smarl

using namespace smarl;

extern "C" int smarl_congestion_step(const SmarlCongestionParams* p, uint8_t* pos_x, uint8_t* pos_y,
                                     const uint8_t* actions, uint8_t* moves, float* obs, float* reward,
                                     int32_t* cost, uint8_t* done, const double* lambdas, float* penalty,
                                     int32_t t, int64_t n_envs, int64_t ld, smarl_stream_t stream) {
  if (int rc = check_congestion(p)) return rc;
  if (int rc = check_layout(n_envs, ld)) return rc;
  SMARL_REQUIRE(pos_x && pos_y && actions && reward && cost, "null required pointer");
  SMARL_REQUIRE(p->noise_mode != 1 || moves, "noise_mode 1 needs the recorded moves");
  SMARL_REQUIRE((lambdas == nullptr) == (penalty == nullptr), "lambdas and penalty go together");
  SMARL_REQUIRE(aligned16(pos_x) && aligned16(pos_y) && aligned16(actions) && aligned16(moves) &&
                    aligned16(obs) && aligned16(reward) && aligned16(cost) && aligned16(done) &&
                    aligned16(penalty), "pointers must be 16-byte aligned");
  CongestionStepArgs a;
  a.pos_x = pos_x; a.pos_y = pos_y; a.actions = actions; a.moves = moves; a.obs = obs;
  a.reward = reward; a.cost = cost; a.done = done; a.lambdas = lambdas; a.penalty = penalty;
  a.demand = p->demand; a.keep_threshold = p->keep_threshold; a.seed = p->seed;
  a.env_offset = p->env_offset; a.n_groups = (n_envs + 3) / 4; a.ld = ld; a.size = p->size; a.t = t;
  const unsigned grid = (unsigned)((a.n_groups + kCongThreads - 1) / kCongThreads);
  cudaStream_t s = (cudaStream_t)stream;
  switch (p->noise_mode) {
    case 0: SMARL_DISPATCH_A(p->n_agents, congestion_step_kernel<kA, 0><<<grid, kCongThreads, 0, s>>>(a)); break;
    case 1: SMARL_DISPATCH_A(p->n_agents, congestion_step_kernel<kA, 1><<<grid, kCongThreads, 0, s>>>(a)); break;
    default: SMARL_DISPATCH_A(p->n_agents, congestion_step_kernel<kA, 2><<<grid, kCongThreads, 0, s>>>(a)); break;
  }
  SMARL_CUDA(cudaGetLastError());
  return SMARL_OK;
}

extern "C" int smarl_congestion_rollout(const SmarlCongestionParams* p, const SmarlAccounting* acc,
                                        const uint8_t* start_x, const uint8_t* start_y,
                                        const uint8_t* actions, const uint8_t* moves,
                                        const double* lambdas, uint8_t* final_x, uint8_t* final_y,
                                        float* R, float* modR, int32_t* C, float* G, float* g_scratch,
                                        double* stats, double* stats_scratch, int64_t n_envs,
                                        int64_t ld, smarl_stream_t stream) {
  set_error("smarl_congestion_rollout: fused Congestion rollout not built yet; use the step + returns path");
  return SMARL_EUNSUPPORTED;
}
