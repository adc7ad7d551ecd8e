// Probe for the tcgen05 wrappers in csrc/tc.cuh: one CTA, D[128 x N] = A[128 x K] * B[N x K]^T with bf16 operands in the
// canonical K-major no-swizzle shared-memory layout, accumulators in TMEM, read back with tcgen05.ld and compared with
// a host reference.  Run for both readings of the descriptor's leading / stride byte offsets so that the layout the
// policy kernel relies on is pinned by measurement:   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../safe_multiagent_rl_b200/csrc/tc.cuh"

using namespace smarl;

// image offset (bytes) of element (row, k) of a [rows x K] K-major operand: [k chunk of 8][row group of 8][8 rows][16 B]
__host__ __device__ inline size_t img_off(int row, int k, int rows) {
  return (size_t)(k >> 3) * rows * 16 + (size_t)(row >> 3) * 128 + (size_t)(row & 7) * 16 + (size_t)(k & 7) * 2;
}

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, float* d, int N, int K,
                                                    int variant, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sa = smem;
  uint8_t* sb = smem + (size_t)128 * K * 2;
  const int tid = threadIdx.x;
  for (int i = tid; i < 128 * K * 2 / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (int i = tid; i < N * K * 2 / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  if (tid == 0) tc::mbar_init(&bar, 1);
  if (tid < 32) tc::tmem_alloc(&tmem_slot, tmem_cols);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_bf16_f32(128, N);
    const uint32_t a_chunk = 128 * 16, b_chunk = N * 16;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a0 = tc::smem_u32(sa) + 2 * ks * a_chunk, b0 = tc::smem_u32(sb) + 2 * ks * b_chunk;
      const uint64_t da = variant == 0 ? tc::smem_desc(a0, a_chunk, 128) : tc::smem_desc(a0, 128, a_chunk);
      const uint64_t db = variant == 0 ? tc::smem_desc(b0, b_chunk, 128) : tc::smem_desc(b0, 128, b_chunk);
      tc::mma_bf16(tmem, da, db, idesc, ks > 0);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  const int warp = tid >> 5;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) d[(size_t)tid * N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, tmem_cols);
}

__global__ void ffma2_kernel(float* out) {
  const float2 r = tc::ffma2(make_float2(1.5f, -2.f), make_float2(4.f, 0.25f), make_float2(0.5f, 10.f));
  out[0] = r.x;
  out[1] = r.y;
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main(int argc, char** argv) {
  const int n_variants = argc > 1 ? 2 : 1;   // any argument: also try the swapped LBO / SBO reading (faults on B200)
  float* f2;
  cudaMalloc(&f2, 8);
  ffma2_kernel<<<1, 1>>>(f2);
  float h2[2];
  cudaMemcpy(h2, f2, 8, cudaMemcpyDeviceToHost);
  printf("ffma2: %g %g (want 6.5 9.5) %s\n", h2[0], h2[1], cudaGetErrorString(cudaGetLastError()));
  int bad = !(h2[0] == 6.5f && h2[1] == 9.5f);
  const int shapes[][2] = {{128, 16}, {128, 48}, {256, 112}, {48, 32}, {16, 16}};
  for (auto& sh : shapes) {
    const int N = sh[0], K = sh[1];
    std::vector<float> A(128 * K), B((size_t)N * K);
    srand(7 + N + K);
    for (auto& x : A) x = (float)(rand() % 200);                                  // u8 positions: exact in bf16
    for (auto& x : B) x = bf16_round((rand() / (float)RAND_MAX - 0.5f) * 0.7f);   // one bf16 split of a weight
    std::vector<uint8_t> ai((size_t)128 * K * 2), bi((size_t)N * K * 2);
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < K; ++k) *reinterpret_cast<__nv_bfloat16*>(&ai[img_off(m, k, 128)]) = __float2bfloat16_rn(A[m * K + k]);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) *reinterpret_cast<__nv_bfloat16*>(&bi[img_off(n, k, N)]) = __float2bfloat16_rn(B[(size_t)n * K + k]);
    uint8_t *da, *db;
    float* dd;
    cudaMalloc(&da, ai.size());
    cudaMalloc(&db, bi.size());
    cudaMalloc(&dd, (size_t)128 * N * 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    uint32_t cols = 32;
    while ((int)cols < N) cols *= 2;
    const size_t smem = ai.size() + bi.size();
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int variant = 0; variant < n_variants; ++variant) {
      cudaMemset(dd, 0xff, (size_t)128 * N * 4);
      probe_kernel<<<1, 128, smem>>>(da, db, dd, N, K, variant, cols);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("N=%d K=%d variant %d: CUDA error %s\n", N, K, variant, cudaGetErrorString(e));
        return 2;
      }
      std::vector<float> D((size_t)128 * N);
      cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[(size_t)n * K + k];
          maxerr = fmax(maxerr, fabs(ref - D[(size_t)m * N + n]));
          maxref = fmax(maxref, fabs(ref));
        }
      printf("N=%d K=%d variant %d (%s): max |D - ref| = %.3e (max |ref| %.1f) %s\n", N, K, variant,
             variant == 0 ? "LBO = K-chunk stride, SBO = 8-row-group stride" : "swapped", maxerr, maxref,
             maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
      if (variant == 0 && !(maxerr < 1e-3 * maxref)) bad = 1;
    }
    cudaFree(da); cudaFree(db); cudaFree(dd);
  }
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
