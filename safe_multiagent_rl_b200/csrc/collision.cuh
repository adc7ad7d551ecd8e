// CollisionAvoidance device code shared by the one-thread-per-env kernels (collision.cu) and the
// lane-cooperative kernels (collision_coop.cu).  Everything here restates the reference's float64
// arithmetic operation by operation; both files are compiled with -fmad=false.
#pragma once
#include "common.cuh"
#include "stats.cuh"

namespace smarl {

struct CollisionStepArgs {
  double* pos_x;
  double* pos_y;
  uint8_t* done;
  const float* actions;       // [2A][ld]
  const double* landmarks;    // [2L][ld]
  float* obs;
  float* reward;
  int32_t* cost;
  uint8_t* done_out;
  int32_t* episode_len;
  const double* lambdas;
  float* penalty;
  int64_t n_envs;
  int64_t ld;
  double size;
  double agents_size;
  int32_t L;
  int32_t obs_landmarks;
  int32_t normalize;
  int32_t reward_rows;
};

constexpr int kCollThreads = 128;
constexpr int64_t kCollCapMinEnvs = 1 << 18;   // batches from which the A <= 4 kernels run register-capped (see below)

// observation value of a coordinate: the state itself, or state / size with normalize_state
// (collision_avoidance.py:164-165, a float64 division), rounded once to f32.
__device__ __forceinline__ float obs_value(double v, double size, int normalize) {
  return (float)(normalize ? __ddiv_rn(v, size) : v);
}

static __device__ __noinline__ float obs_normalized(double v, double size) { return (float)__ddiv_rn(v, size); }

// numpy's pairwise float64 sum of n <= 128 contiguous values (what np.sum does to the A
// per-agent minima at collision_avoidance.py:161): n < 8 sequential; otherwise 8 running
// accumulators over blocks of 8, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail.
template <int N>
__device__ __forceinline__ double numpy_sum(const double (&v)[N]) {
  if (N < 8) {
    double s = 0.0;   // np.sum starts from the first element; 0.0 + v0 == v0 exactly (v0 >= 0)
#pragma unroll
    for (int i = 0; i < N; ++i) s = __dadd_rn(s, v[i]);
    return s;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = v[j < N ? j : 0];
  constexpr int kFull = N - (N % 8);
#pragma unroll
  for (int i = 8; i < kFull; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], v[(i + j) < N ? (i + j) : 0]);
  }
  double s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                       __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
#pragma unroll
  for (int i = kFull; i < N; ++i) s = __dadd_rn(s, v[i]);
  return s;
}

// sqrt_rn(q) < lim, out of line: only ever reached inside the 1e-9 / 1e-6 bands around lim^2 (practically
// never), so the call sites stay three instructions instead of an inlined f64 square root each.
static __device__ __noinline__ bool sqrt_below(double q, double lim) { return __dsqrt_rn(q) < lim; }

// Exact recount of the colliding pairs (sqrt on every pair) for the rare env with a pair inside the
// 1e-6 band around (2 agents_size)^2; rolled loops over a local-memory copy keep it out of the hot code.
static __device__ __noinline__ int collisions_exact(const double* x, const double* y, int A, uint32_t alive,
                                                    double lim) {
  int n = 0;
  for (int i = 0; i < A; ++i)
    for (int j = i + 1; j < A; ++j) {
      const double dx = __dadd_rn(x[i], -x[j]), dy = __dadd_rn(y[i], -y[j]);
      const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      n += (__dsqrt_rn(q) < lim && ((alive >> i) & (alive >> j) & 1u)) ? 1 : 0;
    }
  return n;
}

// Exact test of one pair (the f64 arithmetic of the pair loop below) for the f32-screened path.
static __device__ __noinline__ int pair_collides(const double* x, const double* y, int i, int j, double lim2_lo,
                                                 double lim2_hi, double lim) {
  const double dx = __dadd_rn(x[i], -x[j]), dy = __dadd_rn(y[i], -y[j]);
  const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  if (q < lim2_lo) return 1;
  return (q < lim2_hi && __dsqrt_rn(q) < lim) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// Fused open-loop episode: positions / done mask / discounted sums in registers for all T
// steps (main.py:28-57 minus the policy nets, incl. the early break at :51).  The env reward
// is shared by all agents, so one (S_rew, S_pen) pair per env serves every agent.
// ---------------------------------------------------------------------------------------
struct CollisionRolloutArgs {
  const double* start_x;
  const double* start_y;
  const double* landmarks;
  const float* actions;      // [T][2A][ld]
  const double* lambdas;
  double* final_x;
  double* final_y;
  uint8_t* final_done;
  int32_t* n_active;
  float* R;
  float* modR;
  int32_t* C;
  float* G;
  float* g_scratch;          // [2][T][ld]
  double* partials;
  const double* thresholds;
  double gamma;
  int64_t n_envs;
  int64_t ld;
  double size;
  double agents_size;
  int32_t L;
  int32_t n_steps;
  int32_t g_mode;
};


// One agent's move (collision_avoidance.py:111-119): clip the action to unit norm, add, clamp to [0,size].
__device__ __forceinline__ void collision_move_agent(double& px, double& py, float adx, float ady, double size) {
  double dx = (double)adx, dy = (double)ady;
  // :113  fp32-origin components: dx**2 == dx*dx exactly, so this is the reference's norm.
  // sqrt_rn is monotonic with sqrt_rn(1) == 1, so norm > 1 needs q > 1: the sqrt (and the two
  // divisions) are only evaluated for actions that can actually be clipped.
  const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  if (q > 1.0) {
    const double norm = __dsqrt_rn(q);
    if (norm > 1.0) {                                     // :114-117
      dx = __ddiv_rn(dx, norm);
      dy = __ddiv_rn(dy, norm);
    }
  }
  px = fmax(0.0, fmin(size, __dadd_rn(px, dx)));         // :118
  py = fmax(0.0, fmin(size, __dadd_rn(py, dy)));         // :119
}

// One (agent, landmark) pair: landmark reach (:122-124, np.linalg.norm = sqrt(fma(ay, ay, ax*ax)), probed on OpenBLAS)
// decided on the squared norm outside a 1e-9 band around agents_size^2, and the running minimum of the squared
// distance_matrix entry (:158-161; sqrt_rn is monotonic, so min_l sqrt(q_l) = sqrt(min_l q_l)).
__device__ __forceinline__ bool collision_landmark(double px, double py, double lx, double ly, double agents_size,
                                                   double as2_lo, double as2_hi, double& minq) {
  const double ax = __dadd_rn(px, -lx), ay = __dadd_rn(py, -ly);
  const double axx = __dmul_rn(ax, ax);
  const double qn = __fma_rn(ay, ay, axx);
  bool hit = qn < as2_lo;
  if (!hit && qn < as2_hi) hit = sqrt_below(qn, agents_size);
  minq = fmin(minq, __dadd_rn(axx, __dmul_rn(ay, ay)));
  return hit;
}

// f32 screen threshold of the colliding-pair test: coordinates <= size carry an absolute f32 rounding error
// <= size * 2^-24 each, so an f32 pair distance is off by < size * 2^-22; the screen keeps (lim + size * 2^-21)^2
// plus the rounding of the squared sum itself, which is rigorous for any agents_size (the 1 % relative margin used
// before missed pairs for agents_size <~ 1e-3 on large fields).
__host__ __device__ inline float collision_screen_q(double lim, double size) {
  const double d = lim + size * (1.0 / 2097152.0);
  return (float)(d * d * 1.000001) * 1.000001f;
}

int launch_collision_step(int A, const CollisionStepArgs& a, unsigned grid, cudaStream_t s);
int launch_collision_rollout(int A, const CollisionRolloutArgs& a, unsigned grid, cudaStream_t s);
int launch_collision_coop_step(int A, int S, const CollisionStepArgs& a, cudaStream_t s);
int launch_collision_coop_rollout(int A, int S, const CollisionRolloutArgs& a, cudaStream_t s);
// lanes per env of the lane-cooperative kernels for this agent count / batch (0 = one thread per env)
int collision_coop_lanes(int A, int L, int64_t ld);

}  // namespace smarl
