// Access-pattern sweep for the step kernel's traffic mix (no arithmetic): envs per thread V in {4,8,16},
// CTA size, and store cache policy.  Per env group and agent: read 3 u8 rows, write 4 u8 rows + 3 f32 rows.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int POL> __device__ __forceinline__ void st32(void* p, uint32_t v) {
  if (POL == 0) asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (POL == 1) asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (POL == 2) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (POL == 3) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int POL> __device__ __forceinline__ void st128(void* p, uint4 v) {
  if (POL == 0) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (POL == 1) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (POL == 2) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (POL == 3) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t ld32(const void* p) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

// V = envs per thread (multiple of 4): u8 rows move V bytes per thread, f32 rows V*4 bytes.
template <int A, int V, int POL>
__global__ void step_shaped(uint8_t* px, uint8_t* py, const uint8_t* act, float* obs, float* rew, uint8_t* cost,
                            uint8_t* done, int64_t ng, int64_t ld) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  int64_t e0 = g * V;
  constexpr int W = V / 4;
  uint32_t x[A][W], y[A][W], a[A][W];
#pragma unroll
  for (int i = 0; i < A; ++i)
#pragma unroll
    for (int w = 0; w < W; ++w) {
      x[i][w] = ld32(px + i * ld + e0 + 4 * w); y[i][w] = ld32(py + i * ld + e0 + 4 * w); a[i][w] = ld32(act + i * ld + e0 + 4 * w);
    }
#pragma unroll
  for (int i = 0; i < A; ++i)
#pragma unroll
    for (int w = 0; w < W; ++w) {
      uint32_t s = x[i][w] ^ y[i][w] ^ a[i][w];
      int64_t o = i * ld + e0 + 4 * w;
      st32<POL>(px + o, s); st32<POL>(py + o, s + 1); st32<POL>(cost + o, s & 0x01010101u); st32<POL>(done + o, 0u);
      uint4 f = make_uint4(s, s, s, s);
      st128<POL>(obs + (2 * i) * ld + e0 + 4 * w, f);
      st128<POL>(obs + (2 * i + 1) * ld + e0 + 4 * w, f);
      st128<POL>(rew + o, f);
    }
}
template <class F> float timeit(F f, int it = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a); for (int i = 0; i < it; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / it;
}
uint8_t *px, *py, *act, *cost, *done_; float *obs, *rew;
const int A = 16; const int64_t E = 1 << 22, ld = E;
template <int V, int POL> void run(int block) {
  int64_t ng = E / V;
  float ms = timeit([&] { step_shaped<A, V, POL><<<(unsigned)((ng + block - 1) / block), block>>>(px, py, act, obs, rew, cost, done_, ng, ld); });
  printf("V=%2d block=%4d pol=%d : %.1f us  %.0f GB/s   (%s)\n", V, block, POL, ms * 1e3, 19.0 * A * E / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  cudaMalloc(&px, A * ld); cudaMalloc(&py, A * ld); cudaMalloc(&act, A * ld); cudaMalloc(&cost, A * ld); cudaMalloc(&done_, A * ld);
  cudaMalloc(&obs, 2 * A * ld * 4); cudaMalloc(&rew, A * ld * 4);
  cudaMemset(px, 1, A * ld); cudaMemset(py, 2, A * ld); cudaMemset(act, 3, A * ld);
  run<4, 0>(64); run<4, 0>(128); run<4, 0>(256); run<4, 0>(512); run<4, 0>(1024);
  run<4, 1>(128); run<4, 2>(128); run<4, 3>(128);
  run<8, 0>(128); run<8, 0>(256); run<8, 1>(128);
  run<16, 0>(64); run<16, 0>(128); run<16, 1>(128);
  return 0;
}
