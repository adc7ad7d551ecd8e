// shuffle=True: per-episode re-randomised starts / landmarks (Agent.reset in envs/coverage.py:266-273,
// envs/congestion.py:211-217, envs/collision_avoidance.py:178-179; _reset_landmarks :100-101).
// The reference draws from numpy's global MT19937 stream, which a batched, sharded run cannot
// replay; here every coordinate pair comes from Philox4x32-10 with counter
// (global env id lo, hi, episode, row) and key (seed lo, seed hi ^ 0x52534554 "RSET"), turned into
// 53-bit uniforms exactly like numpy's random_sample: u = ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53.
#include "common.cuh"

namespace smarl {

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void start_uniforms(uint64_t seed, uint64_t env, uint32_t episode, uint32_t row,
                                               double& ux, double& uy) {
  const uint4 o = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)(env >> 32), episode, row),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x52534554u));
  ux = u53(o.x, o.y);
  uy = u53(o.z, o.w);
}

// kind 0: floor(u*size); kind 1: same, row 0 pinned to (0,0) (Congestion agent 0)
__global__ void random_starts_u8_kernel(int kind, double size, uint64_t seed, uint32_t episode, int64_t env_offset,
                                        uint8_t* __restrict__ sx, uint8_t* __restrict__ sy, int64_t n_envs, int64_t ld) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  const uint32_t row = blockIdx.y;
  double ux, uy;
  start_uniforms(seed, (uint64_t)(env_offset + e), episode, row, ux, uy);
  uint8_t x = (uint8_t)floor(__dmul_rn(ux, size)), y = (uint8_t)floor(__dmul_rn(uy, size));
  if (kind == 1 && row == 0) { x = 0; y = 0; }
  sx[row * ld + e] = x;
  sy[row * ld + e] = y;
}

// kind 2: u*size; kind 3: floor((u*size)*zoom)/zoom
__global__ void random_starts_f64_kernel(int kind, double size, double zoom, uint64_t seed, uint32_t episode,
                                         int64_t env_offset, uint32_t row_offset, double* __restrict__ sx,
                                         double* __restrict__ sy, int64_t row_stride, int64_t n_envs) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  const uint32_t row = blockIdx.y;
  double ux, uy;
  start_uniforms(seed, (uint64_t)(env_offset + e), episode, row + row_offset, ux, uy);
  double x = __dmul_rn(ux, size), y = __dmul_rn(uy, size);
  if (kind == 3) {
    x = __ddiv_rn(floor(__dmul_rn(x, zoom)), zoom);
    y = __ddiv_rn(floor(__dmul_rn(y, zoom)), zoom);
  }
  sx[row * row_stride + e] = x;
  sy[row * row_stride + e] = y;
}

}  // namespace smarl

using namespace smarl;

extern "C" int smarl_random_starts_u8(int32_t kind, int32_t size, uint64_t seed, int64_t episode,
                                      int64_t env_offset, uint8_t* start_x, uint8_t* start_y, int32_t n_agents,
                                      int64_t n_envs, int64_t ld, smarl_stream_t stream) {
  if (int rc = check_layout(n_envs, ld)) return rc;
  SMARL_REQUIRE(kind == 0 || kind == 1, "kind=%d must be 0 or 1", kind);
  SMARL_REQUIRE(size >= 1 && size <= 254 && start_x && start_y, "bad size or null pointer");
  SMARL_REQUIRE(n_agents >= 1 && n_agents <= SMARL_MAX_AGENTS, "n_agents=%d outside 1..32", n_agents);
  dim3 grid((unsigned)((n_envs + 255) / 256), (unsigned)n_agents);
  random_starts_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kind, (double)size, seed, (uint32_t)episode,
                                                                  env_offset, start_x, start_y, n_envs, ld);
  SMARL_CUDA(cudaGetLastError());
  return SMARL_OK;
}

extern "C" int smarl_random_starts_f64(int32_t kind, int32_t size, double zoom, uint64_t seed, int64_t episode,
                                       int64_t env_offset, int32_t row_offset, double* x, double* y,
                                       int64_t row_stride, int32_t n_rows, int64_t n_envs, smarl_stream_t stream) {
  SMARL_REQUIRE(kind == 2 || kind == 3, "kind=%d must be 2 or 3", kind);
  SMARL_REQUIRE(size >= 1 && x && y && n_rows >= 1 && n_envs >= 1 && row_stride >= n_envs, "bad arguments");
  SMARL_REQUIRE(kind == 2 || zoom > 0.0, "kind 3 needs zoom > 0");
  dim3 grid((unsigned)((n_envs + 255) / 256), (unsigned)n_rows);
  random_starts_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kind, (double)size, zoom, seed, (uint32_t)episode,
                                                                   env_offset, (uint32_t)row_offset, x, y, row_stride,
                                                                   n_envs);
  SMARL_CUDA(cudaGetLastError());
  return SMARL_OK;
}
