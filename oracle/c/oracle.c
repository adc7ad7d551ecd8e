/*
 * Plain-C restatement of the reference env-step hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py): the checker for full-size parity runs and a strong multi-core CPU
 * baseline -- never part of the product.
 *
 * Same arithmetic as oracle/numpy_oracle.py (float64 / integers, the reference's operation
 * order; compile with -ffp-contract=off), which is pinned against the live reference; this file
 * is in turn diff-tested bit-exactly against the numpy oracle (tests/test_c_oracle.py).
 * Arrays use the product's agent-major layout ([row][ld], env fastest) so device buffers can be
 * compared directly.  A small pthread parallel-for splits the env range over the host cores.  Citations: paths under /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- minimal parallel-for over env ranges (pthreads; this image's gcc has no libgomp) ---- */
#include <pthread.h>
#include <unistd.h>
typedef void (*range_fn)(void* ctx, int64_t lo, int64_t hi);
typedef struct { range_fn fn; void* ctx; int64_t lo, hi; } par_job;
static void* par_entry(void* p) { par_job* j = (par_job*)p; j->fn(j->ctx, j->lo, j->hi); return NULL; }
int oracle_num_threads(void) {
  const char* s = getenv("ORACLE_THREADS");
  long n = s ? atol(s) : sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}
static void parallel_for(range_fn fn, void* ctx, int64_t n) {
  int nt = oracle_num_threads();
  if (n < 64 || nt == 1) { fn(ctx, 0, n); return; }
  if (nt > n) nt = (int)n;
  pthread_t th[256];
  par_job jobs[256];
  for (int i = 0; i < nt; ++i) {
    jobs[i].fn = fn; jobs[i].ctx = ctx; jobs[i].lo = n * i / nt; jobs[i].hi = n * (i + 1) / nt;
    pthread_create(&th[i], NULL, par_entry, &jobs[i]);
  }
  for (int i = 0; i < nt; ++i) pthread_join(th[i], NULL);
}

static const int DIR_X[5] = {1, -1, 0, 0, 0};   /* envs/coverage.py:176, envs/congestion.py:55 */
static const int DIR_Y[5] = {0, 0, -1, 1, 0};

static inline int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------
 * CoverageDiscrete episode(s): envs/coverage.py:174-196 (transition, constraint), :76-89 (reward),
 * meta_agent.py:21-22, buffer.py:31-39, agent.py:200-206.
 *   start_x/y u8 [A][ld], actions u8 [T][A][ld], lut f64 [lut_len] (pen(q), 0 beyond), weights f64 [A]
 *   out: final_x/y u8 [A][ld], R/modR f64 [A][ld], C i32 [A][ld], G f64 [T][A][ld] (NULL ok),
 *        reward_last f64 [A][ld] = per-agent rewards of the final step (NULL ok)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int size, A; int64_t E, ld; int T; const uint8_t *start_x, *start_y, *actions; const double* lut; int lut_len;
  const double *weights, *lambdas; double gamma; uint8_t *final_x, *final_y; double *R, *modR; int32_t* C;
  double *G, *reward_last, *disc;
} coverage_ctx;

static void coverage_range(void* vctx, int64_t lo, int64_t hi) {
  const coverage_ctx* c_ = (const coverage_ctx*)vctx;
  const int size = c_->size, A = c_->A, T = c_->T, lut_len = c_->lut_len;
  const int64_t ld = c_->ld;
  const uint8_t *start_x = c_->start_x, *start_y = c_->start_y, *actions = c_->actions;
  const double *lut = c_->lut, *weights = c_->weights, *lambdas = c_->lambdas, *disc = c_->disc;
  const double gamma = c_->gamma;
  uint8_t *final_x = c_->final_x, *final_y = c_->final_y;
  double *R = c_->R, *modR = c_->modR, *G = c_->G, *reward_last = c_->reward_last;
  int32_t* C = c_->C;
  {
    double* rew_t = (double*)malloc(sizeof(double) * (size_t)T);
    double* pen_t = (double*)malloc(sizeof(double) * (size_t)T);
    for (int64_t e = lo; e < hi; ++e) {
      int x[32], y[32], cnt[32];
      for (int i = 0; i < A; ++i) { x[i] = start_x[i * ld + e]; y[i] = start_y[i * ld + e]; cnt[i] = 0; }
      for (int t = 0; t < T; ++t) {
        double pen = 0.0;
        for (int i = 0; i < A; ++i) {
          const int a = actions[((int64_t)t * A + i) * ld + e];
          x[i] = clampi(x[i] + DIR_X[a], size);
          y[i] = clampi(y[i] + DIR_Y[a], size);
          const int c = a != 4;                                            /* coverage.py:192 */
          cnt[i] += c;
          pen += lambdas ? lambdas[i] * c : 0.0;
        }
        double rew = 0.0;
        for (int i = 0; i < A; ++i)
          for (int j = i + 1; j < A; ++j) {
            const int dx = x[i] - x[j], dy = y[i] - y[j];
            const int q = dx * dx + dy * dy;
            rew = rew - (q < lut_len ? lut[q] : 0.0);                       /* coverage.py:82-83 */
          }
        rew_t[t] = rew;
        pen_t[t] = pen;
      }
      for (int i = 0; i < A; ++i) {
        const double w = weights ? weights[i] : 1.0;
        double r_sum = 0.0, m_sum = 0.0;
        for (int t = 0; t < T; ++t) {
          const double r = rew_t[t] * w;                                     /* coverage.py:87 */
          r_sum = r_sum + disc[t] * r;
          m_sum = m_sum + disc[t] * (-pen_t[t] + r);
        }
        R[i * ld + e] = r_sum;
        modR[i * ld + e] = m_sum;
        C[i * ld + e] = cnt[i];
        final_x[i * ld + e] = (uint8_t)x[i];
        final_y[i * ld + e] = (uint8_t)y[i];
        if (reward_last) reward_last[i * ld + e] = rew_t[T - 1] * w;
        if (G) {
          double run = 0.0;
          for (int t = T - 1; t >= 0; --t) {                                 /* agent.py:203-205 */
            run = (-pen_t[t] + rew_t[t] * w) + gamma * run;
            G[((int64_t)t * A + i) * ld + e] = run;
          }
        }
      }
    }
    free(rew_t);
    free(pen_t);
  }
}

int oracle_coverage_rollout(int size, int A, int64_t E, int64_t ld, int T, const uint8_t* start_x,
                            const uint8_t* start_y, const uint8_t* actions, const double* lut, int lut_len,
                            const double* weights, const double* lambdas, double gamma, uint8_t* final_x,
                            uint8_t* final_y, double* R, double* modR, int32_t* C, double* G,
                            double* reward_last) {
  double* disc = (double*)malloc(sizeof(double) * (size_t)T);
  for (int t = 0; t < T; ++t) disc[t] = pow(gamma, (double)t);            /* buffer.py:31 gamma ** i */
  coverage_ctx c = {size, A, E, ld, T, start_x, start_y, actions, lut, lut_len, weights, lambdas, gamma,
                    final_x, final_y, R, modR, C, G, reward_last, disc};
  parallel_for(coverage_range, &c, E);
  free(disc);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 and the Congestion noise rule (oracle/philox.py; envs/congestion.py:64-67)
 * ---------------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

/* Congestion episode(s): envs/congestion.py:49-137, accounting as above.
 *   noise_mode 0 none, 1 recorded moves u8 [T][A][ld], 2 Philox(seed, env_offset+e, t, agent>>2)
 *   demand f64 [(size+1)^2]; rewards are rounded to f32 before accounting when round_f32 != 0
 *   (the product publishes f32 rewards).  R/modR f64 [A][ld], C i32 [ld], G f64 [T][A][ld] (NULL ok) */
typedef struct {
  int size, A; int64_t E, ld; int T; const uint8_t *start_x, *start_y, *actions, *moves; int noise_mode;
  uint64_t keep_threshold, seed; int64_t env_offset; const double *demand, *lambdas; double gamma; int round_f32;
  uint8_t *final_x, *final_y; double *R, *modR; int32_t* C; double *G, *disc;
} congestion_ctx;

static void congestion_range(void* vctx, int64_t lo, int64_t hi) {
  const congestion_ctx* c_ = (const congestion_ctx*)vctx;
  const int size = c_->size, A = c_->A, T = c_->T, noise_mode = c_->noise_mode, round_f32 = c_->round_f32;
  const int64_t ld = c_->ld, env_offset = c_->env_offset;
  const uint8_t *start_x = c_->start_x, *start_y = c_->start_y, *actions = c_->actions, *moves = c_->moves;
  const uint64_t keep_threshold = c_->keep_threshold, seed = c_->seed;
  const double *demand = c_->demand, *disc = c_->disc;
  const double gamma = c_->gamma;
  uint8_t *final_x = c_->final_x, *final_y = c_->final_y;
  double *R = c_->R, *modR = c_->modR, *G = c_->G;
  int32_t* C = c_->C;
  const int W = size + 1;
  const double lam = c_->lambdas ? c_->lambdas[0] : 0.0;
  {
    double* rew = (double*)malloc(sizeof(double) * (size_t)T * 32);
    double* pen_t = (double*)malloc(sizeof(double) * (size_t)T);
    for (int64_t e = lo; e < hi; ++e) {
      int x[32], y[32], csum = 0;
      for (int i = 0; i < A; ++i) { x[i] = start_x[i * ld + e]; y[i] = start_y[i * ld + e]; }
      const uint64_t id = (uint64_t)(env_offset + e);
      for (int t = 0; t < T; ++t) {
        int act[32], mv[32], con[32];
        int64_t key[32];
        for (int i = 0; i < A; ++i) act[i] = actions[((int64_t)t * A + i) * ld + e];
        if (noise_mode == 1) {
          for (int i = 0; i < A; ++i) mv[i] = moves[((int64_t)t * A + i) * ld + e];
        } else if (noise_mode == 2) {
          for (int j = 0; j < (A + 3) / 4; ++j) {
            uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (uint32_t)j};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            for (int q = 0; q < 4 && 4 * j + q < A; ++q)      /* u1 = w 2^-32, int(u2 5) = w mod 5 (oracle/philox.py) */
              mv[4 * j + q] = ((uint64_t)c[q] < keep_threshold) ? act[4 * j + q] : (int)(c[q] % 5u);
          }
        } else {
          for (int i = 0; i < A; ++i) mv[i] = act[i];
        }
        for (int i = 0; i < A; ++i) {
          const int nx = clampi(x[i] + DIR_X[mv[i]], size), ny = clampi(y[i] + DIR_Y[mv[i]], size);
          key[i] = (((int64_t)x[i] * W + y[i]) * W + nx) * W + ny;          /* agent.edge, :72 */
          x[i] = nx; y[i] = ny;
          con[i] = 0;
        }
        /* Congestion._congestions (:113-137), literal control flow */
        for (int i = 0; i < A; ++i) {
          if (con[i]) continue;
          if (act[i] < 4) {
            int n = 1;
            for (int j = i + 1; j < A; ++j) n += key[j] == key[i];
            for (int j = i + 1; j < A; ++j) if (key[j] == key[i]) con[j] = n - 1;
            con[i] = n - 1;
          }                                                                 /* else: the ==5 branch never groups */
        }
        int at_origin = 0;
        for (int i = 0; i < A; ++i) {
          double r;
          if (act[i] < 4) r = -4.0 - con[i] * 2.0;                          /* :84 */
          else r = -30.0 * (con[i] + 1) / demand[x[i] * W + y[i]] + 7.5 - 4.0;   /* :86-87 */
          if (round_f32) r = (double)(float)r;
          rew[(size_t)t * 32 + i] = r;
          at_origin += (x[i] == 0 && y[i] == 0);
        }
        int cost = A / 3 - at_origin;                                       /* :94-99 */
        if (cost < 0) cost = 0;
        csum += cost;
        pen_t[t] = round_f32 ? (double)(float)(lam * cost) : lam * cost;
      }
      C[e] = csum;
      for (int i = 0; i < A; ++i) {
        double r_sum = 0.0, m_sum = 0.0;
        for (int t = 0; t < T; ++t) {
          const double r = rew[(size_t)t * 32 + i];
          r_sum = r_sum + disc[t] * r;
          m_sum = m_sum + disc[t] * (-pen_t[t] + r);
        }
        R[i * ld + e] = r_sum;
        modR[i * ld + e] = m_sum;
        final_x[i * ld + e] = (uint8_t)x[i];
        final_y[i * ld + e] = (uint8_t)y[i];
        if (G) {
          double run = 0.0;
          for (int t = T - 1; t >= 0; --t) {
            run = (-pen_t[t] + rew[(size_t)t * 32 + i]) + gamma * run;
            G[((int64_t)t * A + i) * ld + e] = run;
          }
        }
      }
    }
    free(rew);
    free(pen_t);
  }
}

int oracle_congestion_rollout(int size, int A, int64_t E, int64_t ld, int T, const uint8_t* start_x,
                              const uint8_t* start_y, const uint8_t* actions, const uint8_t* moves,
                              int noise_mode, uint64_t keep_threshold, uint64_t seed, int64_t env_offset,
                              const double* demand, const double* lambdas, double gamma, int round_f32,
                              uint8_t* final_x, uint8_t* final_y, double* R, double* modR, int32_t* C,
                              double* G) {
  double* disc = (double*)malloc(sizeof(double) * (size_t)T);
  for (int t = 0; t < T; ++t) disc[t] = pow(gamma, (double)t);
  congestion_ctx c = {size, A, E, ld, T, start_x, start_y, actions, moves, noise_mode, keep_threshold, seed,
                      env_offset, demand, lambdas, gamma, round_f32, final_x, final_y, R, modR, C, G, disc};
  parallel_for(congestion_range, &c, E);
  free(disc);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * CollisionAvoidance episode(s): envs/collision_avoidance.py:103-162, main.py:51 early break.
 * ---------------------------------------------------------------------------------------- */
static double numpy_sum(const double* v, int n) {        /* numpy pairwise sum, n <= 128 */
  if (n < 8) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i];
    return s;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = v[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] += v[i + j];
  double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) s += v[i];
  return s;
}

/*   start_x/y f64 [A][ld], landmarks f64 [2L][ld], actions f32 [T][2A][ld]
 *   out: final_x/y f64 [A][ld], final_done u8 [A][ld], n_active i32 [ld], R/modR f64 [A][ld] (same value
 *   per agent), C i32 [ld], G f64 [T][A][ld] (NULL ok) */
typedef struct {
  int size, A, L; int64_t E, ld; int T; const double *start_x, *start_y, *landmarks; const float* actions;
  const double* lambdas; double gamma, agents_size; int round_f32; double *final_x, *final_y; uint8_t* final_done;
  int32_t* n_active; double *R, *modR; int32_t* C; double *G, *disc;
} collision_ctx;

static void collision_range(void* vctx, int64_t lo, int64_t hi) {
  const collision_ctx* c_ = (const collision_ctx*)vctx;
  const int A = c_->A, L = c_->L, T = c_->T, round_f32 = c_->round_f32;
  const int64_t ld = c_->ld;
  const double *start_x = c_->start_x, *start_y = c_->start_y, *landmarks = c_->landmarks, *disc = c_->disc;
  const float* actions = c_->actions;
  const double gamma = c_->gamma, agents_size = c_->agents_size;
  double *final_x = c_->final_x, *final_y = c_->final_y, *R = c_->R, *modR = c_->modR, *G = c_->G;
  uint8_t* final_done = c_->final_done;
  int32_t *n_active = c_->n_active, *C = c_->C;
  const double lam = c_->lambdas ? c_->lambdas[0] : 0.0;
  const double S = (double)c_->size;
  {
    double* rew_t = (double*)malloc(sizeof(double) * (size_t)T);
    double* pen_t = (double*)malloc(sizeof(double) * (size_t)T);
    for (int64_t e = lo; e < hi; ++e) {
      double px[32], py[32], mind[32];
      int done[32], csum = 0, steps = 0;
      for (int i = 0; i < A; ++i) { px[i] = start_x[i * ld + e]; py[i] = start_y[i * ld + e]; done[i] = 0; }
      for (int t = 0; t < T; ++t) {
        int all_done = 1;
        for (int i = 0; i < A; ++i) all_done &= done[i];
        if (all_done) { rew_t[t] = 0.0; pen_t[t] = 0.0; continue; }            /* main.py:51 */
        ++steps;
        for (int i = 0; i < A; ++i) {
          if (done[i]) continue;
          double dx = (double)actions[((int64_t)t * 2 * A + 2 * i) * ld + e];
          double dy = (double)actions[((int64_t)t * 2 * A + 2 * i + 1) * ld + e];
          const double norm = sqrt(dx * dx + dy * dy);                           /* :113 */
          if (norm > 1) { dx = dx / norm; dy = dy / norm; }
          double nx = px[i] + dx, ny = py[i] + dy;
          nx = nx > S ? S : nx; nx = nx < 0 ? 0 : nx;                            /* :118 max(0, min(S, .)) */
          ny = ny > S ? S : ny; ny = ny < 0 ? 0 : ny;
          px[i] = nx; py[i] = ny;
          for (int l = 0; l < L; ++l) {
            const double ax = px[i] - landmarks[(2 * l) * ld + e], ay = py[i] - landmarks[(2 * l + 1) * ld + e];
            if (sqrt(fma(ay, ay, ax * ax)) < agents_size) done[i] = 1;           /* :123 np.linalg.norm */
          }
        }
        for (int i = 0; i < A; ++i) {
          double m = 1e300;
          for (int l = 0; l < L; ++l) {
            const double bx = landmarks[(2 * l) * ld + e] - px[i], by = landmarks[(2 * l + 1) * ld + e] - py[i];
            const double d = sqrt(bx * bx + by * by);
            m = d < m ? d : m;
          }
          mind[i] = m;
        }
        double r = -numpy_sum(mind, A);                                          /* :129,:161 */
        int coll = 0;
        for (int i = 0; i < A; ++i)
          for (int j = i + 1; j < A; ++j) {
            if (done[i] || done[j]) continue;
            const double dx = px[i] - px[j], dy = py[i] - py[j];
            coll += sqrt(dx * dx + dy * dy) < 2 * agents_size;                   /* :155 */
          }
        if (round_f32) r = (double)(float)r;
        rew_t[t] = r;
        pen_t[t] = round_f32 ? (double)(float)(lam * coll) : lam * coll;
        csum += coll;
      }
      double r_sum = 0.0, m_sum = 0.0;
      for (int t = 0; t < T; ++t) {
        r_sum = r_sum + disc[t] * rew_t[t];
        m_sum = m_sum + disc[t] * (-pen_t[t] + rew_t[t]);
      }
      C[e] = csum;
      n_active[e] = steps;
      for (int i = 0; i < A; ++i) {
        final_x[i * ld + e] = px[i]; final_y[i * ld + e] = py[i]; final_done[i * ld + e] = (uint8_t)done[i];
        R[i * ld + e] = r_sum; modR[i * ld + e] = m_sum;
      }
      if (G) {
        double run = 0.0;
        for (int t = T - 1; t >= 0; --t) {
          run = (-pen_t[t] + rew_t[t]) + gamma * run;
          for (int i = 0; i < A; ++i) G[((int64_t)t * A + i) * ld + e] = run;
        }
      }
    }
    free(rew_t);
    free(pen_t);
  }
}

int oracle_collision_rollout(int size, int A, int L, int64_t E, int64_t ld, int T, const double* start_x,
                             const double* start_y, const double* landmarks, const float* actions,
                             const double* lambdas, double gamma, double agents_size, int round_f32,
                             double* final_x, double* final_y, uint8_t* final_done, int32_t* n_active,
                             double* R, double* modR, int32_t* C, double* G) {
  double* disc = (double*)malloc(sizeof(double) * (size_t)T);
  for (int t = 0; t < T; ++t) disc[t] = pow(gamma, (double)t);
  collision_ctx c = {size, A, L, E, ld, T, start_x, start_y, landmarks, actions, lambdas, gamma, agents_size,
                     round_f32, final_x, final_y, final_done, n_active, R, modR, C, G, disc};
  parallel_for(collision_range, &c, E);
  free(disc);
  return 0;
}
