// HBM bandwidth probe for the step kernel's traffic mix (build: nvcc -arch=sm_100a -O3 -o bw_probe bw_probe.cu).
// Measures copy, write-only, read-only, and "step-shaped" traffic: per group of 4 envs and per agent
// read 3 x 4 B, write 4 x 4 B + 3 x 16 B  (pos_x,pos_y,actions in; pos_x,pos_y,cost,done u8 + obs x2, reward f32 out)
// in the same agent-major layout, with no arithmetic.  Tells how much of the copy peak a write-heavy
// (84 % stores) stream can reach.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ldu(const void* p) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ void stu(void* p, uint32_t v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void stf4(void* p, float4 v) { asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

template <int A>
__global__ void __launch_bounds__(128) step_shaped(uint8_t* px, uint8_t* py, const uint8_t* act, float* obs, float* rew,
                                                    uint8_t* cost, uint8_t* done, int64_t ng, int64_t ld) {
  int64_t g = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (g >= ng) return;
  int64_t e0 = g * 4;
  uint32_t x[A], y[A], a[A];
#pragma unroll
  for (int i = 0; i < A; ++i) { x[i] = ldu(px + i * ld + e0); y[i] = ldu(py + i * ld + e0); a[i] = ldu(act + i * ld + e0); }
#pragma unroll
  for (int i = 0; i < A; ++i) {
    uint32_t s = x[i] ^ y[i] ^ a[i];
    stu(px + i * ld + e0, s); stu(py + i * ld + e0, s + 1); stu(cost + i * ld + e0, s & 0x01010101u); stu(done + i * ld + e0, 0u);
    float f = __uint_as_float(s & 0x3fffffffu);
    stf4(obs + (2 * i) * ld + e0, make_float4(f, f, f, f));
    stf4(obs + (2 * i + 1) * ld + e0, make_float4(f, f, f, f));
    stf4(rew + i * ld + e0, make_float4(f, f, f, f));
  }
}
__global__ void copy_k(const float4* __restrict__ a, float4* __restrict__ b, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] = a[i];
}
__global__ void write_k(float4* __restrict__ b, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf4(b + i, make_float4(1.f, 2.f, 3.f, 4.f));
}
__global__ void read_k(const float4* __restrict__ a, float* out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0;
  if (i < n) { float4 v = a[i]; s = v.x + v.y + v.z + v.w; }
  if (s == 123.456f) out[0] = s;
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
  float ms = timeit([&] { step_shaped<A><<<(unsigned)((ng + 127) / 128), 128>>>(px, py, act, obs, rew, cost, done, ng, ld); });
  double bytes = 19.0 * A * E;
  printf("step_shaped  : %.1f us  %.0f GB/s (19 B x A x E = %.3f GB)\n", ms * 1e3, bytes / ms / 1e6, bytes / 1e9);
  int64_t n = (int64_t)1 << 26;  // 1 GiB of float4
  float4 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 0, n * 16);
  ms = timeit([&] { copy_k<<<(unsigned)(n / 256), 256>>>(a, b, n); });
  printf("copy 1+1 GiB : %.1f us  %.0f GB/s\n", ms * 1e3, 2.0 * n * 16 / ms / 1e6);
  ms = timeit([&] { write_k<<<(unsigned)(n / 256), 256>>>(b, n); });
  printf("write 1 GiB  : %.1f us  %.0f GB/s\n", ms * 1e3, 1.0 * n * 16 / ms / 1e6);
  ms = timeit([&] { read_k<<<(unsigned)(n / 256), 256>>>(a, (float*)b, n); });
  printf("read 1 GiB   : %.1f us  %.0f GB/s\n", ms * 1e3, 1.0 * n * 16 / ms / 1e6);
  return 0;
}
