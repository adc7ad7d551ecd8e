// Host-buffer entry points: what a caller holding numpy arrays binds.
// The env dimension is cut into chunks that are pipelined over two streams so that the
// host->device copy of chunk c+1, the fused rollout of chunk c and the device->host copy of
// chunk c-1 overlap (PCIe is full duplex; the kernel is far shorter than either copy).
#include <new>
#include <vector>

#include "stats.cuh"

struct SmarlHostSession {
  int32_t A, K, T;
  int64_t n_envs, ld;
  int n_chunks;
  int64_t chunk;            // envs per chunk (multiple of 16)
  cudaStream_t streams[2];
  uint8_t* d_start_x;
  uint8_t* d_start_y;
  uint8_t* d_actions;       // [T][A][ld]
  float* d_R;
  float* d_modR;
  int32_t* d_C;
  double* d_stats;          // [n_chunks][stats_len]
  double* d_scratch;        // [n_chunks][scratch_len(chunk)]
  int64_t scratch_per_chunk;
  float* d_lut;
  float* d_weights;
  double* d_lambdas;
  double* d_thresholds;
  double* h_stats;          // pinned [n_chunks][stats_len]
};

using namespace smarl;

static void free_session(SmarlHostSession* s) {
  if (!s) return;
  for (auto st : s->streams)
    if (st) cudaStreamDestroy(st);
  cudaFree(s->d_start_x); cudaFree(s->d_start_y); cudaFree(s->d_actions); cudaFree(s->d_R);
  cudaFree(s->d_modR); cudaFree(s->d_C); cudaFree(s->d_stats); cudaFree(s->d_scratch);
  cudaFree(s->d_lut); cudaFree(s->d_weights); cudaFree(s->d_lambdas); cudaFree(s->d_thresholds);
  if (s->h_stats) cudaFreeHost(s->h_stats);
  delete s;
}

extern "C" int smarl_host_session_create(SmarlHostSession** out, int32_t A, int32_t K, int32_t T,
                                         int64_t n_envs) {
  SMARL_REQUIRE(out != nullptr, "out is NULL");
  SMARL_REQUIRE(A >= 1 && A <= SMARL_MAX_AGENTS && K >= 1 && K <= SMARL_MAX_AGENTS, "bad A=%d / K=%d", A, K);
  SMARL_REQUIRE(T >= 1 && T <= 255 && n_envs >= 1, "bad T=%d or n_envs=%lld", T, (long long)n_envs);
  SmarlHostSession* s = new (std::nothrow) SmarlHostSession();
  SMARL_REQUIRE(s != nullptr, "out of host memory");
  s->A = A; s->K = K; s->T = T; s->n_envs = n_envs;
  s->ld = (n_envs + 15) / 16 * 16;
  // ~8 chunks, each a multiple of 16 envs and at least 64Ki envs so launches stay large.
  int64_t chunk = (s->ld / 8 + 15) / 16 * 16;
  if (chunk < 65536) chunk = 65536;
  if (chunk > s->ld) chunk = s->ld;
  s->chunk = chunk;
  s->n_chunks = (int)((n_envs + chunk - 1) / chunk);
  const int sl = stats_len(A, K);
  s->scratch_per_chunk = smarl_stats_scratch_len(A, K, chunk);
#define SMARL_TRY(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      set_error("%s failed: %s", #call, cudaGetErrorString(e__));                   \
      free_session(s);                                                               \
      return SMARL_ECUDA;                                                            \
    }                                                                                \
  } while (0)
  for (auto& st : s->streams) SMARL_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  SMARL_TRY(cudaMalloc(&s->d_start_x, (size_t)A * s->ld));
  SMARL_TRY(cudaMalloc(&s->d_start_y, (size_t)A * s->ld));
  SMARL_TRY(cudaMalloc(&s->d_actions, (size_t)T * A * s->ld));
  SMARL_TRY(cudaMalloc(&s->d_R, sizeof(float) * A * s->ld));
  SMARL_TRY(cudaMalloc(&s->d_modR, sizeof(float) * A * s->ld));
  SMARL_TRY(cudaMalloc(&s->d_C, sizeof(int32_t) * K * s->ld));
  SMARL_TRY(cudaMalloc(&s->d_stats, sizeof(double) * sl * s->n_chunks));
  SMARL_TRY(cudaMalloc(&s->d_scratch, sizeof(double) * s->scratch_per_chunk * s->n_chunks));
  SMARL_TRY(cudaMalloc(&s->d_lut, sizeof(float) * 12288));
  SMARL_TRY(cudaMalloc(&s->d_weights, sizeof(float) * SMARL_MAX_AGENTS));
  SMARL_TRY(cudaMalloc(&s->d_lambdas, sizeof(double) * SMARL_MAX_AGENTS));
  SMARL_TRY(cudaMalloc(&s->d_thresholds, sizeof(double) * SMARL_MAX_AGENTS));
  SMARL_TRY(cudaMallocHost(&s->h_stats, sizeof(double) * sl * s->n_chunks));
#undef SMARL_TRY
  *out = s;
  return SMARL_OK;
}

extern "C" void smarl_host_session_destroy(SmarlHostSession* s) { free_session(s); }

extern "C" int64_t smarl_host_session_ld(const SmarlHostSession* s) { return s ? s->ld : 0; }

extern "C" int smarl_host_coverage_rollout(SmarlHostSession* s, const SmarlCoverageParams* p,
                                           const SmarlAccounting* acc, const uint8_t* start_x_h,
                                           const uint8_t* start_y_h, const uint8_t* actions_h,
                                           const double* lambdas_h, float* R_h, float* modR_h,
                                           int32_t* C_h, double* stats_h) {
  SMARL_REQUIRE(s && p && acc, "null session / params");
  SMARL_REQUIRE(p->n_agents == s->A && s->K == s->A && acc->n_steps == s->T,
                "session was created for A=%d K=%d T=%d", s->A, s->K, s->T);
  SMARL_REQUIRE(acc->g_mode == 0, "host rollout returns episode products only (g_mode 0)");
  SMARL_REQUIRE(p->lut_len >= 0 && p->lut_len <= 12287, "lut_len=%d outside 0..12287", p->lut_len);
  SMARL_REQUIRE(start_x_h && start_y_h && actions_h && R_h && modR_h && C_h, "null host buffer");
  const int A = s->A, T = s->T, sl = stats_len(A, A);
  const int64_t ld = s->ld;
  cudaStream_t s0 = s->streams[0];
  // small parameters
  if (p->lut_len) SMARL_CUDA(cudaMemcpyAsync(s->d_lut, p->lut, sizeof(float) * p->lut_len, cudaMemcpyHostToDevice, s0));
  if (p->weights) SMARL_CUDA(cudaMemcpyAsync(s->d_weights, p->weights, sizeof(float) * A, cudaMemcpyHostToDevice, s0));
  if (lambdas_h) SMARL_CUDA(cudaMemcpyAsync(s->d_lambdas, lambdas_h, sizeof(double) * A, cudaMemcpyHostToDevice, s0));
  if (acc->thresholds) SMARL_CUDA(cudaMemcpyAsync(s->d_thresholds, acc->thresholds, sizeof(double) * A, cudaMemcpyHostToDevice, s0));
  SMARL_CUDA(cudaStreamSynchronize(s0));
  SmarlCoverageParams dp = *p;
  dp.lut = s->d_lut;
  dp.weights = p->weights ? s->d_weights : nullptr;
  SmarlAccounting dacc = *acc;
  dacc.thresholds = acc->thresholds ? s->d_thresholds : nullptr;

  for (int c = 0; c < s->n_chunks; ++c) {
    cudaStream_t st = s->streams[c & 1];
    const int64_t e0 = (int64_t)c * s->chunk;
    const int64_t n = (e0 + s->chunk <= s->n_envs) ? s->chunk : (s->n_envs - e0);
    const int64_t w = (n + 15) / 16 * 16;          // bytes (u8) / elements copied per row
    SMARL_CUDA(cudaMemcpy2DAsync(s->d_start_x + e0, ld, start_x_h + e0, ld, w, A, cudaMemcpyHostToDevice, st));
    SMARL_CUDA(cudaMemcpy2DAsync(s->d_start_y + e0, ld, start_y_h + e0, ld, w, A, cudaMemcpyHostToDevice, st));
    SMARL_CUDA(cudaMemcpy2DAsync(s->d_actions + e0, ld, actions_h + e0, ld, w, (size_t)T * A, cudaMemcpyHostToDevice, st));
    int rc = smarl_coverage_rollout(&dp, &dacc, s->d_start_x + e0, s->d_start_y + e0, s->d_actions + e0,
                                    lambdas_h ? s->d_lambdas : nullptr, nullptr, nullptr, s->d_R + e0,
                                    s->d_modR + e0, s->d_C + e0, nullptr, nullptr,
                                    s->d_stats + (int64_t)c * sl, s->d_scratch + (int64_t)c * s->scratch_per_chunk,
                                    n, ld, st);
    if (rc) return rc;
    SMARL_CUDA(cudaMemcpy2DAsync(R_h + e0, ld * 4, s->d_R + e0, ld * 4, w * 4, A, cudaMemcpyDeviceToHost, st));
    SMARL_CUDA(cudaMemcpy2DAsync(modR_h + e0, ld * 4, s->d_modR + e0, ld * 4, w * 4, A, cudaMemcpyDeviceToHost, st));
    SMARL_CUDA(cudaMemcpy2DAsync(C_h + e0, ld * 4, s->d_C + e0, ld * 4, w * 4, A, cudaMemcpyDeviceToHost, st));
    SMARL_CUDA(cudaMemcpyAsync(s->h_stats + (int64_t)c * sl, s->d_stats + (int64_t)c * sl, sizeof(double) * sl,
                               cudaMemcpyDeviceToHost, st));
  }
  SMARL_CUDA(cudaStreamSynchronize(s->streams[0]));
  SMARL_CUDA(cudaStreamSynchronize(s->streams[1]));
  if (stats_h) {
    for (int j = 0; j < sl; ++j) {
      double v = 0.0;
      for (int c = 0; c < s->n_chunks; ++c) v += s->h_stats[(int64_t)c * sl + j];
      stats_h[j] = v;
    }
  }
  return SMARL_OK;
}
