// Issue-rate probe for the FP32 pipe forms the policy epilogue can use (B200, sm_100a): scalar FFMA with three
// register operands, FFMA with a constant-bank operand, packed FFMA2 (fma.rn.f32x2) with register operands, and LDS.128
// broadcast loads.  Prints cycles per warp-instruction per SM sub-partition with all four sub-partitions busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_probe fma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__constant__ float c_w[64];

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)),
        "l"(*reinterpret_cast<const uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

constexpr int ITER = 4096, CH = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) probe(float* out, long long* cyc, const float* gw) {
  __shared__ float4 s_w[64];
  if (threadIdx.x < 64) s_w[threadIdx.x] = make_float4(gw[threadIdx.x], 1.f, 2.f, 3.f);
  __syncthreads();
  float acc[CH];
  float2 acc2[CH];
  const float x = gw[threadIdx.x & 31], y = gw[(threadIdx.x + 1) & 31];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    acc[i] = i;
    acc2[i] = make_float2(i, -i);
  }
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) acc[i] = fmaf(acc[i], x, y);                                    // FFMA R, R, R, R
      if (MODE == 1) acc[i] = fmaf(x, c_w[i], acc[i]);                               // FFMA R, R, c[][], R
      if (MODE == 2) acc2[i] = ffma2(acc2[i], make_float2(x, y), make_float2(y, x)); // FFMA2 (3 register pairs)
      if (MODE == 3) {                                                               // LDS.128 broadcast + 2 FFMA2
        const float4 w = s_w[(it + i) & 63];
        acc2[i] = ffma2(make_float2(x, y), make_float2(w.x, w.y), acc2[i]);
        acc2[i] = ffma2(make_float2(y, x), make_float2(w.z, w.w), acc2[i]);
      }
      if (MODE == 4) acc[i] = fmaf(x, acc[i], 1.25f);                                // FFMA with an immediate
    }
  }
  const long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) r += acc[i] + acc2[i].x + acc2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, int threads) {
  float *out, *gw;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&gw, 64 * 4);
  cudaMalloc(&cyc, 148 * 8);
  float h[64];
  for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-3f;
  cudaMemcpy(gw, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(c_w, h, sizeof(h));
  probe<MODE><<<148, threads>>>(out, cyc, gw);
  probe<MODE><<<148, threads>>>(out, cyc, gw);
  cudaDeviceSynchronize();
  long long hc[148];
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += hc[i];
  avg /= 148;
  const double warps_per_smsp = threads / 32 / 4.0;
  printf("%-34s %4d threads/SM: %.2f cycles per warp-instruction per sub-partition (%s)\n", name, threads,
         avg / ((double)ITER * instr_per_iter * warps_per_smsp), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(gw); cudaFree(cyc);
}

int main() {
  for (int threads : {256, 512, 1024}) {
    run<0>("FFMA reg,reg,reg", CH, threads);
    run<4>("FFMA reg,reg,imm", CH, threads);
    run<1>("FFMA reg,const,reg", CH, threads);
    run<2>("FFMA2 (3 register pairs)", CH, threads);
    run<3>("LDS.128 broadcast + 2 FFMA2 (per 3)", 3 * CH, threads);
  }
  return 0;
}
