// CoverageContinuous / CoverageDiscretized on sm_100a (float64 positions).  One thread owns one
// env; all arrays agent-major SoA.  Compiled with -fmad=false: positions follow the reference's
// float64 operations one by one (bit-exact); the pair penalty uses a multiply where the reference's
// numpy scalar power calls libm pow (<= 1 ulp per term, invisible after the f32 rounding of rewards).
#include "common.cuh"
#include "stats.cuh"

#ifndef SMARL_TU
#define SMARL_TU -1
#endif
#define SMARL_TU_IS(k) (SMARL_TU == -1 || SMARL_TU == (k))

namespace smarl {

struct CoverageFloatArgs {
  double* pos_x;
  double* pos_y;
  const void* actions;
  float* obs;
  float* reward;
  float* cost;
  uint8_t* done;
  const double* lambdas;
  float* penalty;
  const float* weights;
  int64_t n_envs;
  int64_t ld;
  double size, fieldview, max_norm, zoom, hi, cost_axis, cost_diag;
  double m4_lo, m4_hi;        // max_norm^4 * (1 -+ 1e-9): outside this band q decides sqrt(sqrt(q)) > max_norm
  double rzoom;               // RN(1 / zoom)
  int32_t has_coarseness;
};

constexpr int kCovFThreads = 128;

int launch_coverage_float_step(int mode, int A, const CoverageFloatArgs& a, unsigned grid, cudaStream_t s);

#if SMARL_TU_IS(0) || SMARL_TU_IS(1) || SMARL_TU_IS(3) || SMARL_TU_IS(4)
// x / zoom, correctly rounded, for the kernel-invariant divisor zoom with rzoom = RN(1 / zoom) from the host: a
// multiply and two exact-residual corrections (5 instructions) instead of the ~45 of a float64 division, of which
// CoverageDiscretized.transition needs two per agent-step (coverage.py:230).  q1 is a faithful quotient, and a faithful
// quotient corrected once through the exact FMA residual with a correctly rounded reciprocal IS the rounded quotient
// (Markstein); x is 0 or a normal number of magnitude >= ~1e-16 here, so nothing underflows.  tools/div_probe.cu checked
// 3.3e10 operands of exactly this shape against __ddiv_rn bit for bit (profiles/r02/div_probe.txt).
__device__ __forceinline__ double div_by_zoom(double x, double zoom, double rzoom) {
  const double q0 = __dmul_rn(x, rzoom);
  const double q1 = __fma_rn(__fma_rn(-q0, zoom, x), rzoom, q0);
  return __fma_rn(__fma_rn(-q1, zoom, x), rzoom, q1);
}

// One step of one env held in registers: moves every agent, returns the per-agent costs (f64) and the
// unweighted env reward.  `act` points at this env's lane of step t's action rows.
template <int A, int MODE>
__device__ __forceinline__ void coverage_float_env_step(double (&px)[A], double (&py)[A], const void* act, int64_t ld,
                                                        const CoverageFloatArgs& a, double (&cost)[A], double& rew) {
  // all action loads first: issued back to back with the caller's position loads, one memory latency per step
  // instead of one per agent (ncu on the A = 3 step kernel: 6.9 of 13 stall cycles per issue were long_scoreboard)
  float fa[MODE == 0 ? 2 * A : 1];
  uint32_t ma[MODE == 1 ? A : 1];
#pragma unroll
  for (int i = 0; i < A; ++i) {
    if (MODE == 0) {
      fa[2 * i] = static_cast<const float*>(act)[(2 * i) * ld];
      fa[2 * i + 1] = static_cast<const float*>(act)[(2 * i + 1) * ld];
    } else {
      ma[i] = static_cast<const uint8_t*>(act)[i * ld];
    }
  }
#pragma unroll
  for (int i = 0; i < A; ++i) {
    if (MODE == 0) {                                   // CoverageContinuous.transition, coverage.py:54-74
      double dx = (double)fa[2 * i], dy = (double)fa[2 * i + 1];
      cost[i] = __dsqrt_rn(__fma_rn(dy, dy, __dmul_rn(dx, dx)));       // :94 np.linalg.norm(action)
      if (a.has_coarseness) {
        // :65-69  norm = sqrt(dx^2 + dy^2); if sqrt(norm) > max_norm (sic: sqrt of the norm) rescale.  sqrt_rn is
        // monotonic, so sqrt(sqrt(q)) > max_norm is decided on q outside a 1e-9 band around max_norm^4 and the two
        // square roots are only evaluated for the few actions that are rescaled (or fall inside the band).
        const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (q > a.m4_lo) {
          const double norm = __dsqrt_rn(q);
          if (q > a.m4_hi || __dsqrt_rn(norm) > a.max_norm) {
            dx = __dmul_rn(__ddiv_rn(dx, norm), a.max_norm);
            dy = __dmul_rn(__ddiv_rn(dy, norm), a.max_norm);
          }
        }
      }
      px[i] = fmax(0.0, fmin(a.size, __dadd_rn(px[i], dx)));            // :70
      py[i] = fmax(0.0, fmin(a.size, __dadd_rn(py[i], dy)));
    } else {                                           // CoverageDiscretized.transition, coverage.py:219-234
      const uint32_t m = ma[i];
      // directions (:221): x +1,-1,0,0,+1,+1,-1,-1,0 ; y 0,0,-1,+1,+1,-1,+1,-1,0  -- two bits per action, value + 1
      const int dxi = (int)((0x10A52u >> (2 * m)) & 3u) - 1;
      const int dyi = (int)((0x12285u >> (2 * m)) & 3u) - 1;
      px[i] = div_by_zoom(fmax(0.0, fmin(a.hi, __dadd_rn(__dmul_rn(px[i], a.zoom), (double)dxi))), a.zoom, a.rzoom);   // :230
      py[i] = div_by_zoom(fmax(0.0, fmin(a.hi, __dadd_rn(__dmul_rn(py[i], a.zoom), (double)dyi))), a.zoom, a.rzoom);
      cost[i] = m < 4 ? a.cost_axis : (m < 8 ? a.cost_diag : 0.0);     // :237
    }
  }
  // reward, coverage.py:76-89: i-major sequential f64 sum over overlapping pairs
  rew = 0.0;
#pragma unroll
  for (int i = 0; i < A; ++i) {
#pragma unroll
    for (int j = i + 1; j < A; ++j) {
      const double dx = __dadd_rn(px[j], -px[i]), dy = __dadd_rn(py[j], -py[i]);
      const double gap = __dadd_rn(a.fieldview, -__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
      if (gap > 0.0) rew = __dadd_rn(rew, -__dmul_rn(gap, gap));
    }
  }
}
#endif

#if SMARL_TU_IS(0) || SMARL_TU_IS(1)
template <int A, int MODE>
__global__ void __launch_bounds__(kCovFThreads) coverage_float_step_kernel(const CoverageFloatArgs a) {
  pdl_prologue();   // programmatic dependent launch: the previous grid has completed past this point (common.cuh)
  const int64_t e = (int64_t)blockIdx.x * kCovFThreads + threadIdx.x;
  if (e >= a.n_envs) return;
  const int64_t ld = a.ld;
  double px[A], py[A], cost[A];
  double pen = 0.0, rew;
#pragma unroll
  for (int i = 0; i < A; ++i) {
    px[i] = a.pos_x[i * ld + e];
    py[i] = a.pos_y[i * ld + e];
  }
  const void* act = MODE == 0 ? static_cast<const void*>(static_cast<const float*>(a.actions) + e)
                              : static_cast<const void*>(static_cast<const uint8_t*>(a.actions) + e);
  coverage_float_env_step<A, MODE>(px, py, act, ld, a, cost, rew);
#pragma unroll
  for (int i = 0; i < A; ++i) {
    a.pos_x[i * ld + e] = px[i];
    a.pos_y[i * ld + e] = py[i];
    a.cost[i * ld + e] = (float)cost[i];
    if (a.done) a.done[i * ld + e] = 0;                                 // :97-98
    if (a.obs) {
      a.obs[(2 * i) * ld + e] = (float)px[i];
      a.obs[(2 * i + 1) * ld + e] = (float)py[i];
    }
    if (a.penalty) pen += __ldg(a.lambdas + i) * cost[i];               // meta_agent.py:21-22
    const double w = a.weights ? (double)__ldg(a.weights + i) : 1.0;
    a.reward[i * ld + e] = (float)__dmul_rn(rew, w);
  }
  if (a.penalty) a.penalty[e] = (float)pen;
}

#define SMARL_DEFINE_COVF_STEP(M)                                                                        \
  int launch_coverage_float_step_m##M(int A, const CoverageFloatArgs& a, unsigned grid, cudaStream_t s) { \
    SMARL_DISPATCH_A(A, SMARL_CUDA(launch_pdl(coverage_float_step_kernel<kA, M>, grid, kCovFThreads, 0, s, a)));             \
    SMARL_CUDA(cudaGetLastError());                                                                      \
    return SMARL_OK;                                                                                     \
  }
#endif
int launch_coverage_float_step_m0(int A, const CoverageFloatArgs& a, unsigned grid, cudaStream_t s);
int launch_coverage_float_step_m1(int A, const CoverageFloatArgs& a, unsigned grid, cudaStream_t s);
#if SMARL_TU_IS(0)
SMARL_DEFINE_COVF_STEP(0)
#endif
#if SMARL_TU_IS(1)
SMARL_DEFINE_COVF_STEP(1)
#endif

struct CoverageFloatRolloutArgs {
  CoverageFloatArgs env;      // env.pos_x/pos_y = start positions (read only), env.actions = [T][rows][ld]
  double* final_x;
  double* final_y;
  float* R;
  float* modR;
  float* C;                   // [A][ld] f32 cost sums
  float* G;
  float* g_scratch;           // [2][T][ld]
  double* partials;
  const double* thresholds;
  double gamma;
  int32_t n_steps;
  int32_t g_mode;
};
int launch_coverage_float_rollout_m0(int A, const CoverageFloatRolloutArgs& a, unsigned grid, cudaStream_t s);
int launch_coverage_float_rollout_m1(int A, const CoverageFloatRolloutArgs& a, unsigned grid, cudaStream_t s);

#if SMARL_TU_IS(3) || SMARL_TU_IS(4)
// Fused open-loop episode of the float-position Coverage envs: positions, per-agent cost sums and the
// (S_rew, S_pen) discounted sums stay in registers; reward_a = w_a * rew, so one pair serves every agent.
// CAP: 64 registers (8 CTAs per SM) for A <= 4 on batches >= 2^18 envs (2^20 envs: Continuous 1.69 -> 1.26 ms, Discretized
// 2.34 -> 2.00 ms per batch); a one-wave batch of 65 536 envs is latency-bound per thread and loses 8 % to the spills.
template <int A, int MODE, bool CAP>
__global__ void __launch_bounds__(kCovFThreads, (CAP ? 8 : 0)) coverage_float_rollout_kernel(const CoverageFloatRolloutArgs r) {
  __shared__ double s_red[kCovFThreads / 32];
  const CoverageFloatArgs& a = r.env;
  const int64_t eg = (int64_t)blockIdx.x * kCovFThreads + threadIdx.x;
  const bool live = eg < a.n_envs;
  const int64_t e = live ? eg : 0;
  const int64_t ld = a.ld;
  const int T = r.n_steps;
  constexpr int kRows = MODE == 0 ? 2 * A : A;
  double px[A], py[A], csum[A], lam[A];
#pragma unroll
  for (int i = 0; i < A; ++i) {
    px[i] = a.pos_x[i * ld + e];
    py[i] = a.pos_y[i * ld + e];
    csum[i] = 0.0;
    lam[i] = a.lambdas ? __ldg(a.lambdas + i) : 0.0;
  }
  double s_rew = 0.0, s_pen = 0.0, disc = 1.0;
  for (int t = 0; t < T; ++t) {
    const void* act = MODE == 0
        ? static_cast<const void*>(static_cast<const float*>(a.actions) + (int64_t)t * kRows * ld + e)
        : static_cast<const void*>(static_cast<const uint8_t*>(a.actions) + (int64_t)t * kRows * ld + e);
    double cost[A], rew, pen = 0.0;
    coverage_float_env_step<A, MODE>(px, py, act, ld, a, cost, rew);
#pragma unroll
    for (int i = 0; i < A; ++i) {
      csum[i] += (double)(float)cost[i];             // the step kernel publishes costs as f32
      pen += lam[i] * cost[i];
    }
    const float pf = (float)pen;
    s_rew += disc * rew;
    s_pen += disc * (double)pf;
    if (r.g_mode == 1 && live) {
      r.g_scratch[(int64_t)t * ld + e] = (float)rew;
      r.g_scratch[((int64_t)T + t) * ld + e] = pf;
    } else if (r.g_mode == 2 && live) {
#pragma unroll
      for (int i = 0; i < A; ++i) {
        const double w = a.weights ? (double)__ldg(a.weights + i) : 1.0;
        r.G[((int64_t)t * A + i) * ld + e] = (float)(disc * ((double)(float)__dmul_rn(rew, w) - (double)pf));
      }
    }
    disc *= r.gamma;
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < A; ++i) {
      const double w = a.weights ? (double)__ldg(a.weights + i) : 1.0;
      if (r.final_x) r.final_x[i * ld + e] = px[i];
      if (r.final_y) r.final_y[i * ld + e] = py[i];
      r.R[i * ld + e] = (float)(w * s_rew);
      r.modR[i * ld + e] = (float)(w * s_rew - s_pen);
      r.C[i * ld + e] = (float)csum[i];
    }
  }
  if (r.partials) {
    double* out = r.partials + (int64_t)blockIdx.x * stats_len(A, A);
    const double b_rew = block_sum<kCovFThreads>(live ? s_rew : 0.0, s_red);
    const double b_pen = block_sum<kCovFThreads>(live ? s_pen : 0.0, s_red);
#pragma unroll
    for (int i = 0; i < A; ++i) {
      const double thr = r.thresholds ? __ldg(r.thresholds + i) : 0.0;
      const double c = (double)(float)csum[i];
      const double bc = block_sum<kCovFThreads>(live ? c : 0.0, s_red);
      const double bv = block_sum<kCovFThreads>((live && r.thresholds && c > thr) ? 1.0 : 0.0, s_red);
      if (threadIdx.x == 0) {
        const double w = a.weights ? (double)__ldg(a.weights + i) : 1.0;
        out[i] = bc;
        out[A + i] = bv;
        out[2 * A + i] = w * b_rew;
        out[3 * A + i] = w * b_rew - b_pen;
      }
    }
    if (threadIdx.x == 0) out[4 * A] = 0.0;
  }
  if (r.g_mode == 1 && live) {                          // agent.py:200-206
    double g_rew = 0.0, g_pen = 0.0;
    for (int t = T - 1; t >= 0; --t) {
      g_rew = (double)r.g_scratch[(int64_t)t * ld + e] + r.gamma * g_rew;
      g_pen = (double)r.g_scratch[((int64_t)T + t) * ld + e] + r.gamma * g_pen;
#pragma unroll
      for (int i = 0; i < A; ++i) {
        const double w = a.weights ? (double)__ldg(a.weights + i) : 1.0;
        r.G[((int64_t)t * A + i) * ld + e] = (float)(w * g_rew - g_pen);
      }
    }
  }
}
#define SMARL_DEFINE_COVF_ROLLOUT(M)                                                                              \
  int launch_coverage_float_rollout_m##M(int A, const CoverageFloatRolloutArgs& a, unsigned grid, cudaStream_t s) { \
    if (A <= 4 && a.env.n_envs >= (1 << 18)) {                                                                      \
      switch (A) {                                                                                                  \
        case 1: coverage_float_rollout_kernel<1, M, true><<<grid, kCovFThreads, 0, s>>>(a); break;                  \
        case 2: coverage_float_rollout_kernel<2, M, true><<<grid, kCovFThreads, 0, s>>>(a); break;                  \
        case 3: coverage_float_rollout_kernel<3, M, true><<<grid, kCovFThreads, 0, s>>>(a); break;                  \
        default: coverage_float_rollout_kernel<4, M, true><<<grid, kCovFThreads, 0, s>>>(a); break;                 \
      }                                                                                                             \
    } else {                                                                                                        \
      SMARL_DISPATCH_A(A, coverage_float_rollout_kernel<kA, M, false><<<grid, kCovFThreads, 0, s>>>(a));            \
    }                                                                                                               \
    SMARL_CUDA(cudaGetLastError());                                                                               \
    return SMARL_OK;                                                                                              \
  }
#endif
#if SMARL_TU_IS(3)
SMARL_DEFINE_COVF_ROLLOUT(0)
#endif
#if SMARL_TU_IS(4)
SMARL_DEFINE_COVF_ROLLOUT(1)
#endif

#if SMARL_TU_IS(2)
__global__ void coverage_float_reset_kernel(const double* __restrict__ start_x, const double* __restrict__ start_y,
                                            double* __restrict__ pos_x, double* __restrict__ pos_y,
                                            float* __restrict__ obs, int A, int64_t n_envs, int64_t ld) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  for (int i = 0; i < A; ++i) {
    const double x = start_x[i * ld + e], y = start_y[i * ld + e];
    pos_x[i * ld + e] = x;
    pos_y[i * ld + e] = y;
    if (obs) {
      obs[(2 * i) * ld + e] = (float)x;
      obs[(2 * i + 1) * ld + e] = (float)y;
    }
  }
}
#endif

}  // namespace smarl

using namespace smarl;

#if SMARL_TU_IS(2)
extern "C" int smarl_coverage_float_reset(const double* start_x, const double* start_y, double* pos_x,
                                          double* pos_y, float* obs, int32_t n_agents, int64_t n_envs,
                                          int64_t ld, smarl_stream_t stream) {
  if (int rc = check_layout(n_envs, ld)) return rc;
  SMARL_REQUIRE(start_x && start_y && pos_x && pos_y, "null state pointer");
  SMARL_REQUIRE(n_agents >= 1 && n_agents <= SMARL_MAX_AGENTS, "n_agents=%d outside 1..32", n_agents);
  coverage_float_reset_kernel<<<(unsigned)((n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      start_x, start_y, pos_x, pos_y, obs, n_agents, n_envs, ld);
  SMARL_CUDA(cudaGetLastError());
  return SMARL_OK;
}

extern "C" int smarl_coverage_float_step(const SmarlCoverageFloatParams* p, double* pos_x, double* pos_y,
                                         const void* actions, float* obs, float* reward, float* cost,
                                         uint8_t* done, const double* lambdas, float* penalty,
                                         int64_t n_envs, int64_t ld, smarl_stream_t stream) {
  SMARL_REQUIRE(p != nullptr, "params is NULL");
  if (int rc = check_layout(n_envs, ld)) return rc;
  SMARL_REQUIRE(p->size >= 1, "size=%d must be >= 1", p->size);
  SMARL_REQUIRE(p->mode == 0 || p->mode == 1, "bad mode %d", p->mode);
  SMARL_REQUIRE(p->mode == 0 || p->zoom > 0.0, "mode 1 needs zoom > 0");
  SMARL_REQUIRE(pos_x && pos_y && actions && reward && cost, "null required pointer");
  SMARL_REQUIRE((lambdas == nullptr) == (penalty == nullptr), "lambdas and penalty go together");
  CoverageFloatArgs a;
  a.pos_x = pos_x; a.pos_y = pos_y; a.actions = actions; a.obs = obs; a.reward = reward; a.cost = cost;
  a.done = done; a.lambdas = lambdas; a.penalty = penalty; a.weights = p->weights; a.n_envs = n_envs;
  a.ld = ld; a.size = (double)p->size; a.fieldview = p->fieldview; a.max_norm = p->max_norm;
  a.zoom = p->zoom; a.hi = p->hi; a.cost_axis = p->cost_axis; a.cost_diag = p->cost_diag;
  a.has_coarseness = p->has_coarseness;
  {
    const double m4 = (a.max_norm * a.max_norm) * (a.max_norm * a.max_norm);
    a.m4_lo = m4 * (1.0 - 1e-9);
    a.m4_hi = m4 * (1.0 + 1e-9);
    a.rzoom = 1.0 / a.zoom;
  }
  const unsigned grid = (unsigned)((n_envs + kCovFThreads - 1) / kCovFThreads);
  return p->mode == 0 ? launch_coverage_float_step_m0(p->n_agents, a, grid, (cudaStream_t)stream)
                      : launch_coverage_float_step_m1(p->n_agents, a, grid, (cudaStream_t)stream);
}

extern "C" int smarl_coverage_float_rollout(const SmarlCoverageFloatParams* p, const SmarlAccounting* acc,
                                            const double* start_x, const double* start_y, const void* actions,
                                            const double* lambdas, double* final_x, double* final_y, float* R,
                                            float* modR, float* C, float* G, float* g_scratch, double* stats,
                                            double* stats_scratch, int64_t n_envs, int64_t ld,
                                            smarl_stream_t stream) {
  SMARL_REQUIRE(p != nullptr && acc != nullptr, "params is NULL");
  if (int rc = check_layout(n_envs, ld)) return rc;
  SMARL_REQUIRE(p->size >= 1 && (p->mode == 0 || p->mode == 1), "bad size / mode");
  SMARL_REQUIRE(p->mode == 0 || p->zoom > 0.0, "mode 1 needs zoom > 0");
  SMARL_REQUIRE(acc->n_steps >= 1 && acc->g_mode >= 0 && acc->g_mode <= 2, "bad n_steps / g_mode");
  SMARL_REQUIRE(start_x && start_y && actions && R && modR && C, "null required pointer");
  SMARL_REQUIRE(acc->g_mode == 0 || G, "g_mode != 0 needs G");
  SMARL_REQUIRE(acc->g_mode != 1 || g_scratch, "g_mode 1 needs g_scratch [2][T][ld]");
  SMARL_REQUIRE((stats == nullptr) == (stats_scratch == nullptr), "stats and stats_scratch go together");
  CoverageFloatRolloutArgs r;
  CoverageFloatArgs& a = r.env;
  a.pos_x = const_cast<double*>(start_x); a.pos_y = const_cast<double*>(start_y); a.actions = actions;
  a.obs = nullptr; a.reward = nullptr; a.cost = nullptr; a.done = nullptr; a.lambdas = lambdas; a.penalty = nullptr;
  a.weights = p->weights; a.n_envs = n_envs; a.ld = ld; a.size = (double)p->size; a.fieldview = p->fieldview;
  a.max_norm = p->max_norm; a.zoom = p->zoom; a.hi = p->hi; a.cost_axis = p->cost_axis; a.cost_diag = p->cost_diag;
  a.has_coarseness = p->has_coarseness;
  {
    const double m4 = (a.max_norm * a.max_norm) * (a.max_norm * a.max_norm);
    a.m4_lo = m4 * (1.0 - 1e-9);
    a.m4_hi = m4 * (1.0 + 1e-9);
    a.rzoom = 1.0 / a.zoom;
  }
  r.final_x = final_x; r.final_y = final_y; r.R = R; r.modR = modR; r.C = C; r.G = G; r.g_scratch = g_scratch;
  r.partials = stats_scratch; r.thresholds = acc->thresholds; r.gamma = acc->gamma; r.n_steps = acc->n_steps;
  r.g_mode = acc->g_mode;
  const unsigned grid = (unsigned)((n_envs + kCovFThreads - 1) / kCovFThreads);
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = p->mode == 0 ? launch_coverage_float_rollout_m0(p->n_agents, r, grid, st)
                            : launch_coverage_float_rollout_m1(p->n_agents, r, grid, st))
    return rc;
  if (stats) return launch_stats_finalize(stats_scratch, grid, p->n_agents, p->n_agents, n_envs, stats, st);
  return SMARL_OK;
}
#endif
