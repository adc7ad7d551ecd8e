// Fused per-agent discrete policies, tensor-core build (sm_100a: tcgen05.mma + tensor memory).
//
// Same contract as policy.cu (DiscretePolicy of safe_multi_agent_RL/agent.py:23-47 for all envs and agents, called
// from main.py:30-35): u8 position rows in, u8 action rows + f32 log-probabilities out.  Of the 2A*16 + 80
// multiply-adds per agent-step, fc1 (2A*16: 512 of 592 at A = 16) is ONE GEMM per tile of envs,
//
//     H[128 envs x 16*GA] = X[128 envs x 2A] * W1[2A x 16*GA]          (all agents of a group side by side along N)
//
// because every agent reads the same joint observation X.  It runs on the tensor cores with f32-grade accuracy:
//   * X holds grid coordinates 0..254: exact in bf16.
//   * every fp32 weight is split into three bf16 pieces w = hi + mid + lo (8 + 8 + 8 mantissa bits: the split is exact
//     for normal weights), so X*W1 = X*hi + X*mid + X*lo with exact products and f32 accumulation in tensor memory --
//     three passes over the same A operand with three B operands, smallest pieces first.
//   * the bias rides along as one more K step: X gets a constant block (1,1,1,0,...) and W1 the rows (b_hi, b_mid, b_lo).
// Measured error of the pre-activations against float64: a few 1e-7 relative (tools/umma_probe.cu, tests).
// relu, fc2 (80 multiply-adds, packed FFMA2), softmax, the Philox inverse-CDF sample and log_prob run on the FP32 pipes
// straight out of tensor memory: a thread owns one env (= one TMEM lane) and walks over its warpgroup's agents, 16
// accumulator columns (tcgen05.ld.32x32b.x16) per agent.
//
// CTA = NWG warpgroups; one accumulator tile (128 lanes x 16*GA columns) per CTA; several CTAs per SM overlap one
// CTA's staging + MMA with the others' epilogues (512 TMEM columns per SM).  Persistent over env tiles; the bf16
// images of the group's weights are built once per CTA in shared memory from the fp32 parameters (no prepared
// buffers in the ABI).  Operand layout: canonical K-major, no swizzle -- [K chunk of 8][row group of 8][8 rows][16 B].
#include <cuda_bf16.h>

#include "policy.cuh"
#include "tc.cuh"

namespace smarl {

template <int A, int GMAX>
struct TcCfg {
  static constexpr int H = kPolHidden, NA = kPolActions;
  static constexpr int NGROUPS = (A + GMAX - 1) / GMAX;                    // agent groups; a CTA serves one
  static constexpr int GA = NGROUPS == 1 ? A : (((A + NGROUPS - 1) / NGROUPS + 3) & ~3);   // agents per group (multiple of 4 when split)
  static constexpr int NWG = GA > 8 ? 2 : 1;                               // warpgroups sharing the accumulator tile
  static constexpr int N = H * GA;                                         // accumulator columns = MMA N
  static constexpr int COLS = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  static constexpr int KST = (2 * A + 15) / 16;                            // K steps (16 inputs each) per weight piece
  static constexpr int ACH = 2 * KST + 2;                                  // K chunks of the X image (+ the bias step)
  static constexpr int BCH = 6 * KST + 2;                                  // K chunks of the W1 image: 3 pieces + bias step
  static constexpr int THREADS = 128 * NWG;
  static constexpr int AW = NWG == 1 ? GA : ((((GA + NWG - 1) / NWG) + 3) & ~3);   // agents per warpgroup
  static constexpr int W2S = H * NA + 8;                                   // floats per agent: w2t[c][u], b2[5], pad
  static constexpr uint32_t A_CHUNK = 128 * 16, B_CHUNK = N * 16;          // bytes per K chunk
  static constexpr size_t kSmemA = (size_t)ACH * A_CHUNK;
  static constexpr size_t kSmemB = (size_t)BCH * B_CHUNK;
  static constexpr size_t kSmemW2 = (size_t)GA * W2S * sizeof(float);
  static constexpr size_t kSmem = kSmemA + kSmemB + kSmemW2 + 128;
  static_assert(N % 16 == 0 && N >= 16 && N <= 256, "MMA N out of range");
};

__device__ __forceinline__ uint16_t bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }

template <int A, int GMAX>
__global__ void __launch_bounds__(TcCfg<A, GMAX>::THREADS) policy_act_discrete_tc_kernel(const PolicyArgs a) {
  using C = TcCfg<A, GMAX>;
  constexpr int NWG = C::NWG, H = C::H, NA = C::NA, N = C::N, KST = C::KST, GA = C::GA, AW = C::AW;
  extern __shared__ __align__(128) uint8_t s_raw[];
  uint8_t* s_a = s_raw;                                             // X image   [ACH][16 row groups][8][16 B]
  uint8_t* s_b = s_a + C::kSmemA;                                   // W1 image  [BCH][N / 8][8][16 B]
  float* s_w2 = reinterpret_cast<float*>(s_b + C::kSmemB);          // [GA][W2S]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_w2) + C::kSmemW2);
  uint32_t* s_slot = reinterpret_cast<uint32_t*>(s_bar + 1);
  const int tid = threadIdx.x;
  const int a0 = blockIdx.y * GA;                                   // first agent of this CTA's group
  const int n_real = (A - a0) < GA ? (A - a0) : GA;                 // agents of the group that exist

  // ---- one-time setup: barrier, tensor memory, the bf16 images of the group's weights ---------------------------
  if (tid == 0) tc::mbar_init(s_bar, 1);
  if (tid < 32) tc::tmem_alloc(s_slot, C::COLS);
  {
    // W1 pieces: column n = (local agent, unit), K chunk = piece * 2 KST + k / 8
    constexpr int KP = 16 * KST;
    for (int i = tid; i < GA * KP * H; i += C::THREADS) {
      const int u = i % H, k = (i / H) % KP, j = i / (H * KP);
      float w = (j < n_real && k < 2 * A) ? __ldg(a.w1 + ((size_t)(a0 + j) * (2 * A) + k) * H + u) : 0.f;
      const int n = j * H + u;
      uint8_t* dst = s_b + (size_t)(k >> 3) * C::B_CHUNK + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
#pragma unroll
      for (int piece = 0; piece < 3; ++piece) {
        const __nv_bfloat16 p = __float2bfloat16_rn(w);
        *reinterpret_cast<uint16_t*>(dst + (size_t)piece * 2 * KST * C::B_CHUNK) = __bfloat16_as_ushort(p);
        w -= __bfloat162float(p);                                   // exact: the remainder of an fp32 value
      }
    }
    // bias step: K chunk 6 KST holds (b_hi, b_mid, b_lo, 0, ...) per column, chunk 6 KST + 1 zeros
    for (int i = tid; i < N * 16; i += C::THREADS) {
      const int kk = i % 16, n = i / 16, j = n / H, u = n % H;
      float v = 0.f;
      if (kk < 3 && j < n_real) {
        float w = __ldg(a.b1 + (size_t)(a0 + j) * H + u);
        for (int piece = 0; piece < kk; ++piece) w -= __bfloat162float(__float2bfloat16_rn(w));
        v = w;
      }
      uint8_t* dst = s_b + (size_t)(6 * KST + (kk >> 3)) * C::B_CHUNK + (n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
      *reinterpret_cast<uint16_t*>(dst) = bf16_bits(v);
    }
    // X image, constant part: chunk 2 KST = (1, 1, 1, 0, ...) for every env row, chunk 2 KST + 1 zeros
    for (int i = tid; i < 2 * 128; i += C::THREADS) {
      const uint32_t one2 = 0x3F803F80u;                            // two bf16 ones
      const uint4 v = i < 128 ? make_uint4(one2, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(s_a + (size_t)2 * KST * C::A_CHUNK + (size_t)i * 16) = v;
    }
    // fc2: w2t[c][u] (pairs along u feed FFMA2), then b2
    for (int i = tid; i < GA * C::W2S; i += C::THREADS) {
      const int j = i / C::W2S, r = i % C::W2S;
      float v = 0.f;
      if (j < n_real) {
        if (r < H * NA) v = __ldg(a.w2 + ((size_t)(a0 + j) * H + r % H) * NA + r / H);
        else if (r < H * NA + NA) v = __ldg(a.b2 + (size_t)(a0 + j) * NA + (r - H * NA));
      }
      s_w2[i] = v;
    }
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *s_slot;

  const int wg = tid >> 7, row = tid & 127;
  const uint32_t t_lane = tmem + ((uint32_t)(row & ~31) << 16);     // this warp's 32 TMEM lanes
  const uint2 key = policy_key(a.seed);
  const uint32_t t_word = a.t_word + ((a.episode_dev ? __ldg(a.episode_dev) : 0u) << 16);
  const uint32_t idesc = tc::idesc_bf16_f32(128, N);
  const uint32_t sa_addr = tc::smem_u32(s_a), sb_addr = tc::smem_u32(s_b);
  const int j_lo = wg * AW;                                         // first local agent of this warpgroup
  uint32_t phase = 0;

  // The observation bytes of a tile (4 agents x (x, y) per K chunk, this thread's env row) are fetched one tile ahead
  // into registers, so their global-memory latency hides behind the previous tile's epilogue.
  constexpr int NCH = (2 * KST + NWG - 1) / NWG;                    // K chunks staged by one thread
  uint32_t raw[NCH][8];
  auto fetch = [&](int64_t tile) {
    const int64_t e = tile * 128 + row;
    const bool live = e < a.ld;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = wg + i * NWG;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ag = 4 * c + q;
        const bool on = c < 2 * KST && ag < A && live;
        raw[i][2 * q] = on ? (uint32_t)__ldg(a.pos_x + (int64_t)ag * a.ld + e) : 0u;
        raw[i][2 * q + 1] = on ? (uint32_t)__ldg(a.pos_y + (int64_t)ag * a.ld + e) : 0u;
      }
    }
  };
  if ((int64_t)blockIdx.x < a.n_tiles) fetch(blockIdx.x);

  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t e = tile * 128 + row;
    const bool live = e < a.ld;
    // ---- stage the tile's joint observation as bf16: element k = 2 i is x_i, k = 2 i + 1 is y_i (main.py:33) -----
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = wg + i * NWG;
      if (c < 2 * KST) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 p = __floats2bfloat162_rn((float)raw[i][2 * q], (float)raw[i][2 * q + 1]);
          w[q] = *reinterpret_cast<const uint32_t*>(&p);
        }
        *reinterpret_cast<uint4*>(s_a + (size_t)c * C::A_CHUNK + (size_t)row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();              // also orders the previous tile's tcgen05.ld before the MMAs that overwrite it
    __syncthreads();
    // ---- fc1 on the tensor cores: one elected thread issues, pieces lo -> mid -> hi, then the bias step ---------
    if (tid == 0) {
      tc::fence_after_sync();
#pragma unroll
      for (int piece = 2; piece >= 0; --piece)
#pragma unroll
        for (int ks = 0; ks < KST; ++ks)
          tc::mma_bf16(tmem, tc::smem_desc(sa_addr + 2 * ks * C::A_CHUNK, C::A_CHUNK, 128),
                       tc::smem_desc(sb_addr + (piece * 2 * KST + 2 * ks) * C::B_CHUNK, C::B_CHUNK, 128), idesc,
                       !(piece == 2 && ks == 0));
      tc::mma_bf16(tmem, tc::smem_desc(sa_addr + 2 * KST * C::A_CHUNK, C::A_CHUNK, 128),
                   tc::smem_desc(sb_addr + 6 * KST * C::B_CHUNK, C::B_CHUNK, 128), idesc, true);
      tc::mma_commit(s_bar);
    }
    if (tile + gridDim.x < a.n_tiles) fetch(tile + gridDim.x);
    tc::mbar_wait(s_bar, phase);
    phase ^= 1u;
    tc::fence_after_sync();
    // ---- epilogue out of tensor memory: relu, fc2, softmax, sample, log_prob ---------------------------------
    uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
    uint32_t hr[2][16];
    if (j_lo < n_real) tc::tmem_ld16_raw(t_lane + (uint32_t)(j_lo * H), hr[0]);
#pragma unroll
    for (int jj = 0; jj < AW; ++jj) {
      const int j = j_lo + jj;
      if (j < n_real) {
        tc::tmem_ld_wait_regs(hr[jj & 1]);
        if (jj + 1 < AW && j + 1 < n_real) tc::tmem_ld16_raw(t_lane + (uint32_t)((j + 1) * H), hr[(jj + 1) & 1]);
        const int ag = a0 + j;
        if ((jj & 3) == 0) rnd = policy_words((uint64_t)(a.env_offset + e), t_word, ag >> 2, key);
        const float* w2 = s_w2 + j * C::W2S;
        float2 acc[NA];
#pragma unroll
        for (int c = 0; c < NA; ++c) acc[c] = make_float2(w2[H * NA + c], 0.f);
#pragma unroll
        for (int p4 = 0; p4 < H / 4; ++p4) {
          const float r0 = fmaxf(__uint_as_float(hr[jj & 1][4 * p4 + 0]), 0.f), r1 = fmaxf(__uint_as_float(hr[jj & 1][4 * p4 + 1]), 0.f);
          const float r2 = fmaxf(__uint_as_float(hr[jj & 1][4 * p4 + 2]), 0.f), r3 = fmaxf(__uint_as_float(hr[jj & 1][4 * p4 + 3]), 0.f);
#pragma unroll
          for (int c = 0; c < NA; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(w2 + c * H + 4 * p4);
            acc[c] = tc::ffma2(make_float2(r0, r1), make_float2(w.x, w.y), acc[c]);
            acc[c] = tc::ffma2(make_float2(r2, r3), make_float2(w.z, w.w), acc[c]);
          }
        }
        float l[NA];
#pragma unroll
        for (int c = 0; c < NA; ++c) l[c] = acc[c].x + acc[c].y;
        int pick;
        float lp;
        policy_head(l, word_of(rnd, jj & 3), pick, lp);
        if (live) {
          a.actions[(int64_t)ag * a.ld + e] = (uint8_t)pick;
          if (a.logp) a.logp[(int64_t)ag * a.ld + e] = lp;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, C::COLS);
}

template <int A, int GMAX>
static int launch_tc(const PolicyArgs& a_in, int sms, cudaStream_t st) {
  using C = TcCfg<A, GMAX>;
  PolicyArgs a = a_in;
  auto kern = policy_act_discrete_tc_kernel<A, GMAX>;
  // At most 512 / COLS CTAs of tensor memory fit an SM: size the shared-memory request so that no more than that
  // become resident (a surplus CTA would spin in tcgen05.alloc until a resident one exits).
  const int tmem_ctas = 512 / C::COLS;
  size_t smem = C::kSmem;
  const size_t floor_smem = (size_t)(227 * 1024) / (tmem_ctas + 1) + 1;
  if (smem < floor_smem) smem = floor_smem;
  if (smem > 227 * 1024) {
    set_error("tensor-core policy: %d agents need %zu bytes of shared memory", A, smem);
    return SMARL_EUNSUPPORTED;
  }
  SMARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SMARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // resident CTAs per SM: tensor memory (by construction of the shared-memory request) and registers; the occupancy
  // calculator is not asked because it assumes the default carveout and reported 1 where 2 CTAs run
  cudaFuncAttributes fa;
  SMARL_CUDA(cudaFuncGetAttributes(&fa, kern));
  int per_sm = tmem_ctas;
  const int by_regs = 65536 / (((fa.numRegs + 7) & ~7) * C::THREADS);
  if (per_sm > by_regs) per_sm = by_regs;
  if (per_sm > 2048 / C::THREADS) per_sm = 2048 / C::THREADS;
  if (per_sm < 1) per_sm = 1;
  a.n_tiles = (a.n_envs + 127) / 128;
  int64_t gx = (int64_t)sms * per_sm / C::NGROUPS;
  if (gx < 1) gx = 1;
  if (gx > a.n_tiles) gx = a.n_tiles;
  kern<<<dim3((unsigned)gx, C::NGROUPS), C::THREADS, smem, st>>>(a);
  SMARL_CUDA(cudaGetLastError());
  return SMARL_OK;
}

// group_max: 16 = groups of up to 16 agents (256 accumulator columns, two warpgroups per CTA), 8 = groups of up to 8
// (128 columns, one warpgroup, more CTAs per SM; the tile's observation is staged once per group).
template <int A>
static int launch_tc_any(const PolicyArgs& a, int group_max, int sms, cudaStream_t st) {
  if constexpr (A > 8) {
    if (group_max == 8) return launch_tc<A, 8>(a, sms, st);
  }
  return launch_tc<A, 16>(a, sms, st);
}

int launch_policy_tc(const PolicyArgs& a, int n_agents, int group_max, int sms, cudaStream_t st) {
  SMARL_DISPATCH_A(n_agents, return launch_tc_any<kA>(a, group_max, sms, st));
  return SMARL_OK;
}

}  // namespace smarl
