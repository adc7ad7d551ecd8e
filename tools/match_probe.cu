// Throughput probe for the warp primitives the agents-as-lanes Congestion kernel leans on (B200, sm_100a):
// MATCH.ANY with few / many distinct keys, VOTE (ballot), SHFL, against a LOP3 baseline.  Prints warp-instructions
// per cycle per SM for 1..16 resident warps per SM quadrant-filling launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/match_probe tools/match_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(uint32_t* out, int iters, uint32_t seed, uint32_t modulo) {
  uint32_t v = (threadIdx.x * 2654435761u + seed) % modulo, acc = 0;
  uint32_t a0 = v, a1 = v + 1, a2 = v + 2, a3 = v + 3;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) {         // MATCH.ANY, 4 independent per iteration
      a0 = __match_any_sync(0xffffffffu, a0 % modulo) + i;
      a1 = __match_any_sync(0xffffffffu, a1 % modulo) + i;
      a2 = __match_any_sync(0xffffffffu, a2 % modulo) + i;
      a3 = __match_any_sync(0xffffffffu, a3 % modulo) + i;
    } else if (OP == 1) {  // ballot
      a0 = __ballot_sync(0xffffffffu, (a0 + i) & 1) ^ v;
      a1 = __ballot_sync(0xffffffffu, (a1 + i) & 2) ^ v;
      a2 = __ballot_sync(0xffffffffu, (a2 + i) & 4) ^ v;
      a3 = __ballot_sync(0xffffffffu, (a3 + i) & 8) ^ v;
    } else if (OP == 2) {  // shuffle
      a0 = __shfl_xor_sync(0xffffffffu, a0, 1) + i;
      a1 = __shfl_xor_sync(0xffffffffu, a1, 2) + i;
      a2 = __shfl_xor_sync(0xffffffffu, a2, 4) + i;
      a3 = __shfl_xor_sync(0xffffffffu, a3, 8) + i;
    } else {               // LOP3 / IADD baseline
      a0 = (a0 ^ (a1 + i)) + v;
      a1 = (a1 ^ (a2 + i)) + v;
      a2 = (a2 ^ (a3 + i)) + v;
      a3 = (a3 ^ (a0 + i)) + v;
    }
  }
  acc = a0 ^ a1 ^ a2 ^ a3;
  if (acc == 0x12345678u) out[0] = acc;
}

template <int OP>
static void run(const char* name, uint32_t modulo, int warps_per_sm) {
  int sms = 148;
  uint32_t* out;
  cudaMalloc(&out, 4);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<OP><<<sms, 32 * warps_per_sm>>>(out, 100, 1, modulo);
  cudaEventRecord(e0);
  probe<OP><<<sms, 32 * warps_per_sm>>>(out, iters, 1, modulo);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk_khz * 1e3;
  const double per_sm = 4.0 * iters * warps_per_sm / cycles;     // primitive warp-instructions per cycle per SM
  printf("%-22s modulo %6u  warps/SM %2d  %.3f ms  %.3f warp-instr/cycle/SM  (%.1f cycles each per SM)\n", name, modulo,
         warps_per_sm, ms, per_sm, 1.0 / per_sm);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<0>("match.any few keys", 3, w);
    run<0>("match.any 8 keys", 8, w);
    run<0>("match.any all distinct", 1u << 30, w);
    run<1>("ballot", 16, w);
    run<2>("shfl.xor", 16, w);
    run<3>("lop3+iadd (x2)", 16, w);
  }
  return 0;
}
