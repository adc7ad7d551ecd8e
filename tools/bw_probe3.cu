// Does the 4 B/lane access width of the u8 rows limit the step kernel?  Same byte volume as the step
// kernel's traffic, but (variant 1) the u8 rows are moved with 16 B per lane (every 4th thread handles 16 envs
// of each u8 row), f32 rows unchanged; variant 0 = 4 B per lane as in the kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t ld32(const void* p) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint4 ld128(const void* p) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ void st32(void* p, uint32_t v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st128(void* p, uint4 v) { asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
template <int A, int WIDE>
__global__ void __launch_bounds__(128) k(uint8_t* px, uint8_t* py, const uint8_t* act, float* obs, float* rew, uint8_t* cost, uint8_t* done, int64_t ng, int64_t ld) {
  int64_t g = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (g >= ng) return;
  int64_t e0 = g * 4;
  uint32_t acc = 0;
  if (WIDE) {
    if ((threadIdx.x & 3) == 0) {
#pragma unroll
      for (int i = 0; i < A; ++i) {
        uint4 x = ld128(px + i * ld + e0), y = ld128(py + i * ld + e0), a = ld128(act + i * ld + e0);
        uint4 s = make_uint4(x.x ^ y.x ^ a.x, x.y ^ y.y ^ a.y, x.z ^ y.z ^ a.z, x.w ^ y.w ^ a.w);
        st128(px + i * ld + e0, s); st128(py + i * ld + e0, s); st128(cost + i * ld + e0, s); st128(done + i * ld + e0, make_uint4(0, 0, 0, 0));
        acc ^= s.x;
      }
    }
    acc = __shfl_sync(0xffffffffu, acc, threadIdx.x & ~3);
  } else {
#pragma unroll
    for (int i = 0; i < A; ++i) {
      uint32_t s = ld32(px + i * ld + e0) ^ ld32(py + i * ld + e0) ^ ld32(act + i * ld + e0);
      st32(px + i * ld + e0, s); st32(py + i * ld + e0, s + 1); st32(cost + i * ld + e0, s & 0x01010101u); st32(done + i * ld + e0, 0u);
      acc ^= s;
    }
  }
#pragma unroll
  for (int i = 0; i < A; ++i) {
    uint4 f = make_uint4(acc, acc + i, acc, acc);
    st128(obs + (2 * i) * ld + e0, f); st128(obs + (2 * i + 1) * ld + e0, f); st128(rew + i * ld + e0, f);
  }
}
template <class F> float timeit(F f, int it = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a); for (int i = 0; i < it; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / it;
}
int main() {
  const int A = 16; const int64_t E = 1 << 22, ld = E, ng = E / 4;
  uint8_t *px, *py, *act, *cost, *done; float *obs, *rew;
  cudaMalloc(&px, A * ld); cudaMalloc(&py, A * ld); cudaMalloc(&act, A * ld); cudaMalloc(&cost, A * ld); cudaMalloc(&done, A * ld);
  cudaMalloc(&obs, 2 * A * ld * 4); cudaMalloc(&rew, A * ld * 4);
  cudaMemset(px, 1, A * ld); cudaMemset(py, 2, A * ld); cudaMemset(act, 3, A * ld);
  float ms = timeit([&] { k<A, 0><<<(unsigned)((ng + 127) / 128), 128>>>(px, py, act, obs, rew, cost, done, ng, ld); });
  printf("u8 rows 4 B/lane : %.1f us  %.0f GB/s\n", ms * 1e3, 19.0 * A * E / ms / 1e6);
  ms = timeit([&] { k<A, 1><<<(unsigned)((ng + 127) / 128), 128>>>(px, py, act, obs, rew, cost, done, ng, ld); });
  printf("u8 rows 16 B/lane: %.1f us  %.0f GB/s\n", ms * 1e3, 19.0 * A * E / ms / 1e6);
  // f32-only part (12 B of 19) and u8-only part
  return 0;
}
