// Congestion, lane-cooperative step kernel for large agent counts (sm_100a).
//
// congestion.cu keeps all A agents of four envs in one thread; from A = 16 that is 168-255 registers (2-3 CTAs per
// SM, spills at A = 32) and the kernel falls to 0.55 / 0.31 of the HBM roofline at A = 16 / 32.  Here a group of
// S = 2 or 4 lanes shares one quad of four envs, lane s owning agents i = j*S + s as byte-SIMD words.  A warp covers
// 32/S quads (lane = s*(32/S) + q), so every load / store instruction touches S rows x (32/S) consecutive words
// or 16-byte vectors of the agent-major SoA arrays: whole 32-byte sectors.
//
// _congestions (congestion.py:113-137) is order dependent, so its A(A-1)/2 pair tests cannot be shared between
// the two partners' lanes without sending partial results back.  Instead every agent SCANS all agents in index
// order, as the reference's sequential loop does, over records published to shared memory:
//     r      running number of agents on the same directed edge           (class members so far, incl. self)
//     seen   an INTENDED mover with index <= own was on the edge          (= "active" in the closed form)
//     z      r at the moment the first such mover appears                 (= inactive members before the leader)
//     con = seen ? r_final - z - 1 : 0                                    (#members >= leader, minus self)
// which is the closed form proved against the literal loop in tests/test_oracle_vs_reference.py: members below
// the lowest-index intended mover L of a class get 0, members >= L get #{members >= L} - 1.  Lower partners cost
// 12 integer ops for the four envs, higher ones 8; nothing flows back between lanes.
//
// Philox noise (MODE 2): one generator block serves four agents of one env, so the 4 * ceil(A/4) blocks of a
// quad are spread over the S lanes and handed out through shared memory (no lane evaluates a block twice).
#include "congestion.cuh"

// Built as three translation units, one per noise mode (build.py passes -DSMARL_TU=0|1|2).
#ifndef SMARL_TU
#define SMARL_TU -1
#endif
#define SMARL_TU_IS(k) (SMARL_TU == -1 || SMARL_TU == (k))

namespace smarl {

constexpr int kCongCoopThreads = 128;

template <int A, int S>
struct CongCoop {
  static constexpr int B = (A + S - 1) / S;          // agents per lane
  static constexpr int EPW = 32 / S;                 // env quads per warp
  static constexpr int EPC = kCongCoopThreads / S;   // env quads per CTA
  static constexpr int NQ = (A + 3) / 4;             // Philox blocks per env (four agents each)
  static constexpr bool kGhost = (A % S) != 0;
  static constexpr int PR = (B * S) | 1;             // uint4 records per quad (incl. padding agents): odd => conflict-free LDS.128
  static constexpr int PW = (4 * NQ) | 1;            // uint4 Philox blocks per quad
  static constexpr int PQ = PR > PW ? PR : PW;       // the Philox blocks are consumed before the records are written: one buffer
  static constexpr size_t kSmem = (size_t)EPC * PQ * sizeof(uint4);
};

// r = (e80 >> 7) + r per env byte as ONE multiply-add on the FMA pipe (the kernel is bound by the integer ALU
// pipe; written in C, ptxas turns the same expression into LEA.HI, an ALU-pipe instruction).
__device__ __forceinline__ uint32_t add_flags80(uint32_t e80, uint32_t r) {
  uint32_t o;
  // multiplier 2^25 + 1, not 2^25: the low 7 bits of e80 are clear, so e80 * 2^25 has a zero low word and the
  // extra e80 cannot carry into the high word -- same result, but not a shift ptxas would strength-reduce
  asm("mad.hi.u32 %0, %1, 33554433, %2;" : "=r"(o) : "r"(e80), "r"(r));
  return o;
}
// 0xFF in every env byte whose bit 7 is set in f80 (FMA pipe).
__device__ __forceinline__ uint32_t expand_flags80(uint32_t f80) { return __umulhi(f80, 0xFFu << 25); }

// The scan: all agents in index order against every own agent (see the file header).  Flags live in bit 7 of
// each env byte (0x80), so equality needs no final shift and counting / mask expansion are IMAD.HI.  The loop
// over partner ROWS (S partners each) is rolled -- fully unrolled, the scan alone was ~3000 instructions and the
// kernel stalled 3.6 cycles per issue on instruction fetch -- with the own agents unrolled inside; "partner <= own"
// is uniform per (row, own agent) except on the own row, where a per-lane mask decides.
//   KW    words of the edge key compared (2: the (x_old + 2 x', y_old + 2 y') encoding; 3: (x', y', dcode))
//   LOW7  every key byte is <= 0x7F, which saves two instructions of the zero-byte test
template <int A, int S, int KW, bool LOW7>
__device__ __forceinline__ void congestion_scan(const uint4* __restrict__ rec, const uint32_t (&k0)[CongCoop<A, S>::B],
                                                const uint32_t (&k1)[CongCoop<A, S>::B],
                                                const uint32_t (&k2)[CongCoop<A, S>::B], int s,
                                                uint32_t (&r)[CongCoop<A, S>::B], uint32_t (&seen)[CongCoop<A, S>::B],
                                                uint32_t (&z)[CongCoop<A, S>::B]) {
  constexpr int B = CongCoop<A, S>::B;
  const uint32_t k7f = 0x7F7F7F7Fu, k80 = 0x80808080u;
  uint32_t le80[S];
#pragma unroll
  for (int j = 0; j < B; ++j) r[j] = seen[j] = z[j] = 0u;
#pragma unroll
  for (int t = 0; t < S; ++t) le80[t] = (t <= s) ? k80 : 0u;
#pragma unroll 1
  for (int jr = 0; jr < B; ++jr) {
    uint4 p[S];
#pragma unroll
    for (int t = 0; t < S; ++t) p[t] = rec[jr * S + t];
#pragma unroll
    for (int j = 0; j < B; ++j) {
      uint32_t e[S];
#pragma unroll
      for (int t = 0; t < S; ++t) {
        uint32_t v = (p[t].x ^ k0[j]) | (p[t].y ^ k1[j]);
        if (KW == 3) v |= p[t].z ^ k2[j];
        // 0x80 in every env byte whose edge keys are equal
        e[t] = LOW7 ? (~(v + k7f) & k80) : (~(((v & k7f) + k7f) | v) & k80);
      }
      if (jr <= j) {                                                // uniform: this row holds partners <= own (or the own row)
#pragma unroll
        for (int t = 0; t < S; ++t) {
          const uint32_t w = jr < j ? p[t].w : (p[t].w & le80[t]);  // intended movers with index <= own
          const uint32_t first = e[t] & w & ~seen[j];
          z[j] |= expand_flags80(first) & r[j];                     // r before this partner: members below the leader
          seen[j] |= e[t] & w;
          r[j] = add_flags80(e[t], r[j]);
        }
      } else {
#pragma unroll
        for (int t = 0; t < S; ++t) r[j] = add_flags80(e[t], r[j]);
      }
    }
  }
}

template <int A, int S, int MODE>
__global__ void __launch_bounds__(kCongCoopThreads, 4) congestion_coop_step_kernel(const CongestionStepArgs a) {
  pdl_prologue();   // programmatic dependent launch: the previous grid has completed past this point (common.cuh)
  using C = CongCoop<A, S>;
  constexpr int B = C::B, EPW = C::EPW, EPC = C::EPC, NQ = C::NQ;
  extern __shared__ uint4 s_rec[];                                 // [EPC][PQ]: Philox blocks (MODE 2), then the
                                                                   // (x', y', dcode, intended mover) records
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane % EPW, s = lane / EPW;
  const int ql = warp * EPW + q;
  const int64_t gq = (int64_t)blockIdx.x * EPC + ql;
  const bool live = gq < a.n_groups;
  const uint32_t ld = (uint32_t)a.ld;
  const uint32_t e0 = (uint32_t)(live ? gq : 0) * 4u;
  const uint32_t row0 = (uint32_t)s * ld + e0;                     // 32-bit element offsets (host-checked)
  const uint32_t k01 = 0x01010101u;

  uint32_t xw[B], yw[B], aw[B], mw[B];
#pragma unroll
  for (int j = 0; j < B; ++j) {
    const uint32_t off = row0 + (uint32_t)(j * S) * ld;
    if (!C::kGhost || j < B - 1 || j * S + s < A) {
      xw[j] = ld_stream_u32(a.pos_x + off);
      yw[j] = ld_stream_u32(a.pos_y + off);
      aw[j] = ld_stream_u32(a.actions + off);
      mw[j] = MODE == 1 ? ld_stream_u32(a.moves + off) : aw[j];
    } else {
      xw[j] = yw[j] = 0xFFFFFFFFu;
      aw[j] = mw[j] = 0x04040404u;
    }
  }
  if (MODE == 2) {
    // blocks b = jq*4 + k (agent quad jq, env lane k) of this env quad, b = s, s + S, ...: counter and key as in
    // congestion_noise_moves (congestion.cuh)
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const uint32_t episode = a.episode + (a.episode_dev ? __ldg(a.episode_dev) : 0u);
    const int64_t env0 = a.env_offset + (int64_t)(live ? gq : 0) * 4;
    uint4* wq = s_rec + ql * C::PQ;
#pragma unroll
    for (int m = 0; m * S < 4 * NQ; ++m) {
      const int b = m * S + s;
      if (4 * NQ % S == 0 || b < 4 * NQ) {
        const uint64_t id = (uint64_t)(env0 + (b & 3));
        wq[b] = philox4x32_10(make_uint4((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)a.t,
                                         (uint32_t)(b >> 2) | (episode << 3)), key);
      }
    }
    __syncwarp();
    const uint32_t* ww = reinterpret_cast<const uint32_t*>(wq);
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int i = j * S + s;                                      // word (i & 3) of block (i >> 2, k)
      uint32_t m4 = 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t w = ww[(((i >> 2) * 4 + k) << 2) + (i & 3)];
        const uint32_t act = (aw[j] >> (8 * k)) & 0xFFu;
        const uint32_t mv = ((uint64_t)w < a.keep_threshold) ? act : w % 5u;
        m4 |= mv << (8 * k);
      }
      if (!C::kGhost || j < B - 1 || i < A) mw[j] = m4;
    }
    __syncwarp();                                                   // the records below reuse the block buffer
  }

  // transition (congestion.py:49-75) and the records every lane of the group scans
  const uint32_t size4 = (uint32_t)a.size * k01;
  // edge key (x_old, y_old, x', y') per env byte: up to size 84 the two bytes (x_old + 2 x', y_old + 2 y') identify
  // the directed edge (3 x' - dx with dx in {-1, 0, 1} determines both); larger grids compare (x', y', dcode)
  const bool small_key = a.size <= 84;
  uint32_t k0[B], k1[B], k2[B];
  uint32_t at_origin = 0u;
  uint4* rec = s_rec + ql * C::PQ;
#pragma unroll
  for (int j = 0; j < B; ++j) {
    const bool own = !C::kGhost || j < B - 1 || j * S + s < A;
    const uint32_t ox = xw[j], oy = yw[j];
    grid_move4(xw[j], yw[j], mw[j], size4);
    k0[j] = small_key ? ox + 2u * xw[j] : xw[j];
    k1[j] = small_key ? oy + 2u * yw[j] : yw[j];
    k2[j] = small_key ? 0u : (((xw[j] + k01) - ox) | (((yw[j] + k01) - oy) << 2));
    if (!own) {                                                      // a padding agent's record matches no real edge:
      const uint32_t never = (a.size <= 42 || (a.size > 84 && a.size <= 127)) ? 0x7F7F7F7Fu : 0xFFFFFFFFu;
      if (small_key) k0[j] = never;                                 // real bytes are <= 126 (size <= 42) / <= 252
      else k2[j] = never;                                           // real displacement codes are <= 0x0A
    }
    // record: edge key words, intended mover (action < 4) as 0x80 per env byte
    rec[j * S + s] = make_uint4(k0[j], k1[j], k2[j], own ? ((((aw[j] >> 2) & k01) ^ k01) << 7) : 0u);
    if (own) {
      at_origin += zero_bytes01(xw[j] | yw[j]);                    // congestion.py:97
      if (live) {
        const uint32_t off = row0 + (uint32_t)(j * S) * ld;
        st_stream_u32(a.pos_x + off, xw[j]);
        st_stream_u32(a.pos_y + off, yw[j]);
        if (MODE != 1 && a.moves) st_stream_u32(a.moves + off, mw[j]);
        if (a.done) st_stream_u32(a.done + off, 0u);                // congestion.py:103-104
        if (a.obs) {
          const uint32_t o2 = 2u * off - e0;
          st_stream_f4(a.obs + o2, bytes_to_float4(xw[j]));
          st_stream_f4(a.obs + (o2 + ld), bytes_to_float4(yw[j]));
        }
      }
    }
  }
  __syncwarp();

  // the scan (congestion_scan above), specialised on the width of the edge key
  uint32_t r[B], seen[B], z[B];
  if (a.size <= 42) congestion_scan<A, S, 2, true>(rec, k0, k1, k2, s, r, seen, z);
  else if (a.size <= 84) congestion_scan<A, S, 2, false>(rec, k0, k1, k2, s, r, seen, z);
  else if (a.size <= 127) congestion_scan<A, S, 3, true>(rec, k0, k1, k2, s, r, seen, z);
  else congestion_scan<A, S, 3, false>(rec, k0, k1, k2, s, r, seen, z);
#pragma unroll
  for (int off = 1; off < S; off <<= 1) at_origin += __shfl_xor_sync(0xffffffffu, at_origin, off * EPW);
  if (!live) return;

  // congestion.py:93-100: cost = max(0, A // 3 - #agents on node (0,0)) per env lane
  const int c0 = max(0, A / 3 - (int)(at_origin & 0xFFu)), c1 = max(0, A / 3 - (int)((at_origin >> 8) & 0xFFu)),
            c2 = max(0, A / 3 - (int)((at_origin >> 16) & 0xFFu)), c3 = max(0, A / 3 - (int)(at_origin >> 24));
  if (s == 0) {
    st_stream_i4(a.cost + e0, make_int4(c0, c1, c2, c3));
    if (a.penalty) {                                                // meta_agent.py:21-22
      const double lam = __ldg(a.lambdas);
      st_stream_f4(a.penalty + e0, make_float4((float)(lam * c0), (float)(lam * c1), (float)(lam * c2), (float)(lam * c3)));
    }
  }
  const int W = a.size + 1;
#pragma unroll
  for (int j = 0; j < B; ++j) {
    if (C::kGhost && j == B - 1 && j * S + s >= A) continue;
    const uint32_t conw = (r[j] - z[j] - k01) & __umulhi(seen[j], 0xFFu << 25);
    const float4 rw = a.wait_reward ? congestion_reward4<true>(aw[j], conw, xw[j], yw[j], a.demand, W, a.wait_reward)
                                    : congestion_reward4<false>(aw[j], conw, xw[j], yw[j], a.demand, W, nullptr);
    st_stream_f4(a.reward + (row0 + (uint32_t)(j * S) * ld), rw);
  }
}

template <int S, int MODE>
static int launch_sm(int A, const CongestionStepArgs& a, cudaStream_t st) {
  switch (A) {
#define SMARL_COOP_CASE(N)                                                                              \
  case N: {                                                                                             \
    using C = CongCoop<N, S>;                                                                           \
    auto kern = congestion_coop_step_kernel<N, S, MODE>;                                                \
    const size_t smem = C::kSmem;                                                                       \
    if (smem > 48 * 1024)                                                                               \
      SMARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    const unsigned grid = (unsigned)((a.n_groups + C::EPC - 1) / C::EPC);                               \
    SMARL_CUDA(launch_pdl(kern, grid, kCongCoopThreads, smem, st, a));                                                      \
  } break;
    SMARL_COOP_CASE(9) SMARL_COOP_CASE(10) SMARL_COOP_CASE(11) SMARL_COOP_CASE(12) SMARL_COOP_CASE(13)
    SMARL_COOP_CASE(14) SMARL_COOP_CASE(15) SMARL_COOP_CASE(16) SMARL_COOP_CASE(17) SMARL_COOP_CASE(18)
    SMARL_COOP_CASE(19) SMARL_COOP_CASE(20) SMARL_COOP_CASE(21) SMARL_COOP_CASE(22) SMARL_COOP_CASE(23)
    SMARL_COOP_CASE(24) SMARL_COOP_CASE(25) SMARL_COOP_CASE(26) SMARL_COOP_CASE(27) SMARL_COOP_CASE(28)
    SMARL_COOP_CASE(29) SMARL_COOP_CASE(30) SMARL_COOP_CASE(31) SMARL_COOP_CASE(32)
#undef SMARL_COOP_CASE
    default:
      set_error("cooperative Congestion kernels cover n_agents 9..32 (got %d)", A);
      return SMARL_EUNSUPPORTED;
  }
  SMARL_CUDA(cudaGetLastError());
  return SMARL_OK;
}

#define SMARL_DEFINE_CONG_COOP(M)                                                                          \
  int launch_congestion_coop_step_m##M(int A, int S, const CongestionStepArgs& a, cudaStream_t st) {      \
    return S == 2 ? launch_sm<2, M>(A, a, st) : launch_sm<4, M>(A, a, st);                                 \
  }
#if SMARL_TU_IS(0)
SMARL_DEFINE_CONG_COOP(0)
#endif
#if SMARL_TU_IS(1)
SMARL_DEFINE_CONG_COOP(1)
#endif
#if SMARL_TU_IS(2)
SMARL_DEFINE_CONG_COOP(2)
#endif

}  // namespace smarl
