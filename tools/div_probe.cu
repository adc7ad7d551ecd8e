// Division by a kernel-invariant divisor without the f64 division sequence: is
//     q0 = x * r;  e0 = fma(-q0, y, x);  q1 = fma(e0, r, q0);  e1 = fma(-q1, y, x);  q2 = fma(e1, r, q1)
// with r = RN(1 / y) always the correctly rounded x / y?  (Markstein: a faithful quotient plus one exact-residual
// correction with a correctly rounded reciprocal rounds correctly; q1 is faithful.)  This probe checks it bit for bit
// against __ddiv_rn on random operands of the ranges coverage_float.cu divides: x in [0, hi] incl. values a few ulps off
// the lattice points, y = coarseness / size for many (coarseness, size).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/div_probe tools/div_probe.cu && tools/div_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ double div_inv(double x, double y, double r) {
  const double q0 = __dmul_rn(x, r);
  const double e0 = __fma_rn(-q0, y, x);
  const double q1 = __fma_rn(e0, r, q0);
  const double e1 = __fma_rn(-q1, y, x);
  return __fma_rn(e1, r, q1);
}

__global__ void probe(double y, double hi, int mode, uint64_t seed, unsigned long long* bad, unsigned long long* bad1, int iters) {
  uint64_t s = seed + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x632BE59BD9B4E019ull;
  const double r = 1.0 / y;
  unsigned long long nb = 0, nb1 = 0;
  for (int i = 0; i < iters; ++i) {
    const uint64_t w = splitmix(s);
    double x;
    if (mode == 0) {                     // uniform in [0, hi]
      x = (double)(w >> 11) * (1.0 / 9007199254740992.0) * hi;
    } else if (mode == 1) {              // lattice points k / y * y +- a few ulps, + d: what the transition produces
      const int k = (int)(w % (uint64_t)(hi + 1.0));
      double p = __ddiv_rn((double)k, y);
      // (not around 0: bit patterns next to 0.0 are denormals / NaNs, which positions never are -- a position is 0 or
      // at least ~1e-16 / zoom, because every step adds an integer before the clamp)
      if (k > 0) p = __longlong_as_double(__double_as_longlong(p) + (long long)((w >> 20) % 9) - 4);
      const int d = (int)((w >> 40) % 3) - 1;
      x = fmax(0.0, fmin(hi, __dadd_rn(__dmul_rn(p, y), (double)d)));
    } else {                             // random bit patterns with exponents -60 .. +10
      const uint64_t m = w & 0x000FFFFFFFFFFFFFull;
      const uint64_t e = 1023 - 60 + ((w >> 52) % 71);
      x = __longlong_as_double((long long)((e << 52) | m));
    }
    const double want = __ddiv_rn(x, y);
    const double got = div_inv(x, y, r);
    const double q0 = __dmul_rn(x, r), e0 = __fma_rn(-q0, y, x), q1 = __fma_rn(e0, r, q0);
    nb += __double_as_longlong(want) != __double_as_longlong(got);
    nb1 += __double_as_longlong(want) != __double_as_longlong(q1);
  }
  atomicAdd(bad, nb);
  atomicAdd(bad1, nb1);
}

int main() {
  unsigned long long *bad, *bad1;
  cudaMallocManaged(&bad, 8);
  cudaMallocManaged(&bad1, 8);
  const int sizes[] = {1, 2, 3, 5, 7, 8, 10, 16, 32, 33, 64, 100, 127};
  const int coarse[] = {1, 2, 3, 4, 6, 7, 10, 12, 25, 50, 100};
  unsigned long long total = 0, total_bad = 0, total_bad1 = 0;
  for (int s : sizes)
    for (int c : coarse)
      for (int mode = 0; mode < 3; ++mode) {
        *bad = 0; *bad1 = 0;
        const double y = (double)c / (double)s;
        probe<<<592, 256>>>(y, (double)c, mode, 1234567ull * s + 89ull * c + mode, bad, bad1, 512);
        cudaDeviceSynchronize();
        total += 592ull * 256 * 512;
        total_bad += *bad;
        total_bad1 += *bad1;
        if (*bad) printf("size %d coarseness %d mode %d: %llu mismatches\n", s, c, mode, *bad);
      }
  printf("%llu divisions checked against __ddiv_rn: %llu mismatches with two corrections, %llu with one\n", total, total_bad, total_bad1);
  return total_bad != 0;
}
