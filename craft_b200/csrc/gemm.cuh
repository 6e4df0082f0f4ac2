// craft_b200 -- "shift-GEMM": the one tcgen05 GEMM main loop behind every projection and every
// convolution of the CRAFT update block.
//
//   D[m, n] = sum_{t < T} sum_{k < K}  A[m + tap_off[t], a_koff + k] * B[t * Npad + n, b_koff + k]
//
// A and B are bf16, K-major, addressed through TMA tensor maps; D accumulates in fp32 in TMEM.
// A convolution over the padded-flat token grid (DESIGN.md "geometry") is the case T = kh*kw with
// tap_off[t] = dy * Wp + dx: the shifted A tile is the same tensor map with a different row
// coordinate, rows that fall off either end are zero-filled by TMA, and the two halo cells at the
// end of every grid row hold zeros, so no im2col buffer ever exists.  Projections are T = 1.
//
// Thread-block clusters: CL consecutive M tiles form a cluster.  They all need the same weight
// tile B(t, k), so each CTA fetches 1/CL of it and TMA-multicasts that slice into the shared
// memory of every CTA of the cluster (L2 -> SM traffic for weights drops by CL and the L2 slices
// holding the weights are no longer hit by every SM at once).  A stage is recycled when all CL
// consumers have released it: the MMA warp's tcgen05.commit arrives on the `empty` barrier of
// every CTA in the cluster.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (thread == accumulator row == TMEM lane; the two warps that share a lane
// quadrant split the BN columns).  Epilogue operands that do not depend on the accumulator
// (bias tile, GRU state) are fetched while the main loop runs.
#pragma once
#include "common.cuh"

namespace cb {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;       // 64 bf16 = one 128-byte swizzle atom row
constexpr int kGemmThreads = 320;
constexpr int kMaxTaps = 49;

enum GemmEpilogue : int {
  EPI_STORE = 0,    // v = act(alpha*acc + bias[n]) -> bf16 and/or f32 row-major
  EPI_GRU_ZR = 1,   // n<128: z=sigmoid -> Z(f32);  n>=128: r=sigmoid, (r*h) -> bf16
  EPI_GRU_Q = 2,    // q=tanh; h=(1-z)h+zq -> Hm(f32) and bf16 copy
  EPI_MOTION = 3,   // n<126: relu(acc+bias) ; n in {126,127}: flow channel copy
  EPI_FLOW = 4,     // flow head's last conv: columns 0,1 = delta -> f32 store, and coords1 += delta,
                    // flow = coords1 - coords0 in the same epilogue (core/network.py:247,236)
};

struct GemmParams {
  int M;        // rows of D (and of A)
  int Npad;     // columns of D computed (multiple of BN); B holds T*Npad rows
  int K;        // reduction length per tap (multiple of 64)
  int T;        // number of taps
  int a_koff;   // first A column used
  int b_koff;   // first B column used
  // b_blocked != 0: B rows are tokens of a padded-flat grid and the n-th tile of BN rows is the
  // 8 x (BN/8) spatial block (by, bx) = (n / b_nbx, n % b_nbx), fetched with a 3-D TMA box (tmB is
  // then a [H][W][C] map).  Used to emit V^T with keys in the block order attn_pv consumes.
  int b_blocked, b_nbx;
  int tap_off[kMaxTaps];
  int stages;   // pipeline depth actually used (stages * kc <= GemmSmem<BN>::kStages slots); tuning knob
  int kc;       // 64-column K atoms per pipeline stage (1 or 2): 2 halves the per-stage handshake cost
  // ashare != 0: the taps come in groups of `gsize` consecutive row offsets (the kw taps of one kernel
  // row); the A rows of a group are loaded ONCE (128 + 8 rows) and every tap reads them through a
  // descriptor whose start address is shifted by whole rows, so only the weights are fetched per tap.
  // (Measured: the 128-byte swizzle is a function of the absolute smem address, so the shifted start
  // needs NO base-offset correction in the descriptor -- setting bits 49-51 gives wrong results.)
  int ashare, gsize;
  // bigbox != 0 (experimental, untested on hardware at the end of round 1 -- profiles/r01_mb_tma.txt shows
  // a TMA box costs ~257 clk whatever its size): tmA / tmB are 3-D maps over [64][rows][K/64] and ONE box
  // per operand brings both K atoms of a stage; stage layout [A0][A1][B0][B1].  bigbox == 2: A and B are
  // issued by two lanes of the producer warp.  Needs kc == 2, no cluster, no blocked B.
  int bigbox;
  // row validity: Wp > 0 => row m is a real token iff (m % Wp) < W and (m / Wp) < H;
  //               Wp == 0 => every m < M is valid.
  int Wp, W, H;
  // epilogue operands
  float alpha;
  int act;                       // 0 none, 1 relu
  const float* bias;             // [Npad] or nullptr
  act_t* out_b;          // bf16 destination (row-major), may be nullptr
  int ldb, colb;
  float* out_f;                  // f32 destination, may be nullptr
  int ldf, colf;
  float* aux_f0;                 // EPI_GRU_*: Z   [M,128] f32 ; EPI_FLOW: coords1 [M,2] f32
  float* aux_f1;                 // EPI_GRU_*: Hm  [M,128] f32 ; EPI_MOTION / EPI_FLOW: flow [M,2] f32
  // CRAFT_GEMM_TRACE (profiling): clock64 of CTA (0,0)'s phases in [0,16), then per CTA {globaltimer start, end, smid}
  long long* trace;
};

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 6 : (BN >= 64 ? 8 : 10));
  static constexpr int kABytes = kGemmBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // tap-shared ("grouped") stages: one 136-row A tile + up to 5 weight tiles under one barrier; at least two of
  // them must fit (BN = 128: 2 x 99,328 B; BN = 256 does not use the mode)
  static constexpr int kGroupStage5 = 136 * 128 + 5 * kBBytes;
  static constexpr int kRing = (BN <= 128 && 2 * kGroupStage5 > kStages * kStageBytes) ? 2 * kGroupStage5 : kStages * kStageBytes;
  static constexpr int kTotal = kRing + 1024 /*align*/ + 256 /*barriers*/ + BN * 4 /*bias*/;
  static_assert(kTotal <= 227 * 1024, "shift_gemm: shared memory budget");
};

__device__ __forceinline__ float sigmoid_fast(float x) {
  return fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x));
}
__device__ __forceinline__ float tanh_fast(float x) {
  // 1 - 2/(1 + e^{2x}); saturates correctly for |x| large (ex2 -> inf or 0)
  return 1.0f - 2.0f * fast_rcp(1.0f + fast_ex2(2.8853900817779268f * x));
}

template <int BN, int EPI, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
shift_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ GemmParams p) {
  using S = GemmSmem<BN>;
  static_assert(BN % (8 * CL) == 0, "weight slice per CTA must be whole 8-row swizzle groups");
  pdl_launch_dependents();
  const bool tr0 = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0;
#define G_TRACE(slot) do { if (tr0) p.trace[slot] = clock64(); } while (0)
  // marks taken before griddepcontrol.wait stay in registers until after it (no global access may precede the wait)
  long long tr_c0 = 0, tr_c1 = 0, tr_g0 = 0;
  if (p.trace != nullptr && threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_g0));
    tr_c0 = clock64();
  }
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms.  The dynamic smem window starts at
  // the same offset in every CTA of the cluster, so the aligned addresses agree too (multicast
  // writes land at identical offsets).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // barrier block (256 B): [acc_bar][tmem_slot][ring barriers ...]
  uint64_t* acc_bar = reinterpret_cast<uint64_t*>(smem + S::kRing);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  uint64_t* full_bar = acc_bar + 2;
  uint64_t* empty_bar = full_bar + S::kStages;
  float* s_bias = reinterpret_cast<float*>(smem + S::kRing + 256);
  // shared-A ("grouped") mode: a stage = one 136-row A tile (every tap of a kernel row is a row shift of it) followed
  // by the gsize weight tiles of that kernel row, all under ONE full / empty barrier pair -- the issuing warp pays
  // one wait + one commit per 4 * gsize MMAs, and the SM ingests 17 KB + gsize * BN * 128 B per gsize taps instead of
  // gsize * (16 KB + BN * 128 B).  (Round 1 built this with a barrier per weight tile: exact, less ingest, slower.)
  constexpr int kAsBytes = 136 * 128;
  const int g_stage_bytes = kAsBytes + p.gsize * S::kBBytes;

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * kGemmBM;
  const int n0 = blockIdx.y * BN;
  const int kchunks = p.K / kGemmBK;
  const int nk = p.T * kchunks;
  constexpr uint32_t kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));   // power of two >= BN
  const int nstages = p.stages;
  uint32_t cta_rank = 0;
  if constexpr (CL > 1) cta_rank = cluster_ctarank();
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], p.bigbox == 2 ? 2 : 1);     // one arrive.expect_tx per producer lane
      mbar_init(&empty_bar[s], CL);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  if (p.trace != nullptr && threadIdx.x == 0) tr_c1 = clock64();
  pdl_wait();                                    // A, bias and the GRU state come from earlier kernels
  if (p.trace != nullptr && threadIdx.x == 0) {
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    long long* e = p.trace + 16 + 3 * (blockIdx.y * gridDim.x + blockIdx.x);
    e[0] = tr_g0; e[2] = sm;
    if (tr0) { p.trace[0] = tr_c0; p.trace[1] = tr_c1; }
  }
  if (threadIdx.x == 0) G_TRACE(2);
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) G_TRACE(3);

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // A pipeline stage is `kc` consecutive smem slots (one slot = A atom + B atom of 64 K columns).
    if (p.ashare) {
      if (elect_one()) {
        const int GS = p.gsize, ngr = p.T / GS;
        int stage = 0;
        uint32_t phase = 0;
        for (int gi = 0; gi < ngr; ++gi)
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(g_stage_bytes));
            uint8_t* sa = smem + stage * g_stage_bytes;
            // rows m0 + off(first tap of the group) .. +136: every tap of the group is a row shift of it
            tma_load_2d(sa, &tmA, &full_bar[stage], p.a_koff + kc * kGemmBK, m0 + p.tap_off[gi * GS]);
            for (int j = 0; j < GS; ++j)
              tma_load_2d(sa + kAsBytes + j * S::kBBytes, &tmB, &full_bar[stage], p.b_koff + kc * kGemmBK,
                          (gi * GS + j) * p.Npad + n0);
            if (gi == 0 && kc == 0) G_TRACE(4);
            if (++stage == nstages) { stage = 0; phase ^= 1u; }
          }
        G_TRACE(5);
      }
    } else if (p.bigbox) {
      const uint32_t lane = lane_id();
      const bool two = p.bigbox == 2;
      if (lane < (two ? 2u : 1u)) {
        const bool do_a = !two || lane == 0, do_b = !two || lane == 1;
        const uint32_t bytes = (do_a ? 2u * S::kABytes : 0u) + (do_b ? 2u * S::kBBytes : 0u);
        const int nst = nk / 2;
        int stage = 0, t = 0, kc = 0;
        uint32_t phase = 0;
        for (int it = 0; it < nst; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          uint8_t* sa = smem + stage * 2 * S::kStageBytes;       // [A atom 0][A atom 1][B atom 0][B atom 1]
          if (do_a) tma_load_3d(sa, &tmA, &full_bar[stage], 0, m0 + p.tap_off[t], (p.a_koff >> 6) + kc);
          if (do_b) tma_load_3d(sa + 2 * S::kABytes, &tmB, &full_bar[stage], 0, t * p.Npad + n0, (p.b_koff >> 6) + kc);
          kc += 2;
          if (kc == kchunks) { kc = 0; ++t; }
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (elect_one()) {
      const int KC = p.kc;
      const int nst = nk / KC;
      int stage = 0;
      uint32_t phase = 0;
      int t = 0, kc = 0;                                // tap and 64-column chunk of the next slot
      for (int it = 0; it < nst; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);     // all CL consumers released this stage
        mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(KC) * S::kStageBytes);
        for (int a = 0; a < KC; ++a) {
          uint8_t* sa = smem + (stage * KC + a) * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          tma_load_2d(sa, &tmA, &full_bar[stage], p.a_koff + kc * kGemmBK, m0 + p.tap_off[t]);
          if (p.b_blocked) {
            const int by = static_cast<int>(blockIdx.y) / p.b_nbx, bx = static_cast<int>(blockIdx.y) - by * p.b_nbx;
            tma_load_3d(sb, &tmB, &full_bar[stage], p.b_koff + kc * kGemmBK, bx * (BN / 8), by * 8);
          } else if constexpr (CL > 1) {
            constexpr int kSlice = BN / CL;
            tma_load_2d_mcast(sb + cta_rank * kSlice * 128, &tmB, &full_bar[stage], p.b_koff + kc * kGemmBK,
                              t * p.Npad + n0 + static_cast<int>(cta_rank) * kSlice, kMask);
          } else {
            tma_load_2d(sb, &tmB, &full_bar[stage], p.b_koff + kc * kGemmBK, t * p.Npad + n0);
          }
          if (++kc == kchunks) { kc = 0; ++t; }
        }
        if (it == 0) G_TRACE(4);
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
      G_TRACE(5);
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    // This warp's instruction latency is the pacing item of the whole kernel (a tcgen05.mma is one
    // instruction, the 64 clk it keeps the tensor pipe busy are easily lost to descriptor arithmetic
    // and barrier round trips), so the loop is kept minimal: descriptors advance by constant adds,
    // the next stage's barrier is probed while the current MMAs execute, one commit per stage.
    constexpr uint32_t idesc = umma_idesc_f16<kGemmBM, (BN < 16 ? 16 : BN)>();
    constexpr uint64_t kSlotStep = static_cast<uint64_t>(S::kStageBytes >> 4);
    constexpr uint64_t kBOff = static_cast<uint64_t>(S::kABytes >> 4);
    const int KC = p.kc;
    const int nst = nk / KC;
    const uint64_t desc_base = umma_desc_sw128(smem_u32(smem));
    const bool leader = elect_one();
    if (p.ashare) {
      const int GS = p.gsize;
      const int nst_g = (p.T / GS) * kchunks;
      const uint64_t g_step = static_cast<uint64_t>(g_stage_bytes >> 4);
      constexpr uint64_t kBStep = static_cast<uint64_t>(S::kBBytes >> 4);
      constexpr uint64_t kAOff = static_cast<uint64_t>(kAsBytes >> 4);
      int stage = 0;
      uint32_t phase = 0;
      uint64_t da = desc_base;
      bool ready = mbar_try_wait_nohint(&full_bar[0], 0);
      for (int it = 0; it < nst_g; ++it) {
        if (!ready) mbar_wait(&full_bar[stage], phase);
        if (it == 0) G_TRACE(6);
        tc_fence_after();
        if (leader) {
          for (int j = 0; j < GS; ++j) {
            // tap j of the group reads the A rows j * 128 bytes further down the same smem tile (the 128-byte
            // swizzle follows the absolute smem address, so a row-shifted start needs no descriptor fix-up)
            const uint64_t d = da + static_cast<uint64_t>(j) * 8u;
            const uint64_t e = da + kAOff + static_cast<uint64_t>(j) * kBStep;
            umma_f16(tmem_base, d, e, idesc, (it | j) != 0 ? 1u : 0u);
            umma_f16_acc(tmem_base, d + 2u, e + 2u, idesc);
            umma_f16_acc(tmem_base, d + 4u, e + 4u, idesc);
            umma_f16_acc(tmem_base, d + 6u, e + 6u, idesc);
          }
          umma_commit(&empty_bar[stage]);
          if (it == nst_g - 1) umma_commit(acc_bar);
        }
        if (++stage == nstages) { stage = 0; phase ^= 1u; da = desc_base; }
        else da += g_step;
        ready = (it + 1 < nst_g) && mbar_try_wait_nohint(&full_bar[stage], phase);
      }
      G_TRACE(7);
      __syncwarp();
    } else {
    int stage = 0;
    uint32_t phase = 0;
    uint64_t da = desc_base;
    bool ready = mbar_try_wait_nohint(&full_bar[0], 0);
    for (int it = 0; it < nst; ++it) {
      if (!ready) mbar_wait(&full_bar[stage], phase);
      if (it == 0) G_TRACE(6);
      tc_fence_after();
      if (leader) {
        for (int a = 0; a < KC; ++a) {
          // legacy stage: [A0 B0][A1 B1]; big-box stage: [A0 A1][B0 B1]
          const uint64_t d = da + static_cast<uint64_t>(a) * (p.bigbox ? kBOff : kSlotStep);
          const uint64_t e = p.bigbox ? da + 2u * kBOff + static_cast<uint64_t>(a) * static_cast<uint64_t>(S::kBBytes >> 4)
                                      : d + kBOff;
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
          umma_f16(tmem_base, d, e, idesc, (it | a) != 0 ? 1u : 0u);
          umma_f16_acc(tmem_base, d + 2u, e + 2u, idesc);
          umma_f16_acc(tmem_base, d + 4u, e + 4u, idesc);
          umma_f16_acc(tmem_base, d + 6u, e + 6u, idesc);
        }
        // frees the smem stage (in every CTA of the cluster) when these MMAs retire
        if constexpr (CL > 1) umma_commit_mcast(&empty_bar[stage], kMask);
        else umma_commit(&empty_bar[stage]);
        if (it == nst - 1) umma_commit(acc_bar);  // accumulator complete
      }
      if (++stage == nstages) { stage = 0; phase ^= 1u; da = desc_base; }
      else da += static_cast<uint64_t>(KC) * kSlotStep;
      ready = (it + 1 < nst) && mbar_try_wait_nohint(&full_bar[stage], phase);
    }
    G_TRACE(7);
    __syncwarp();
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    constexpr int HALF = BN / 2;                       // columns per thread
    constexpr int CW = (HALF % 32 == 0) ? 32 : 16;     // columns per TMEM load (BN = 96 / 32: 16)
    const int et = threadIdx.x - 64;                   // 0..255
    const int lane_grp = warp & 3;                     // TMEM lanes this warp may read
    const int half = (warp - 2) >> 2;                  // which half of the BN columns
    const int row = lane_grp * 32 + (threadIdx.x & 31);
    const int m = m0 + row;
    bool valid = m < p.M;
    if (p.Wp > 0) {
      const int y = m / p.Wp;
      const int x = m - y * p.Wp;
      valid = valid && (x < p.W) && (y < p.H);
    }
    // ---- work that does not depend on the accumulator overlaps the main loop
    for (int i = et; i < BN; i += 256) s_bias[i] = p.bias ? p.bias[n0 + i] : 0.0f;
    const int nbase = n0 + half * HALF;                // first global column of this thread
    float pre_h[(EPI == EPI_GRU_ZR || EPI == EPI_GRU_Q) ? HALF : 1];
    float pre_z[(EPI == EPI_GRU_Q) ? HALF : 1];
    if constexpr (EPI == EPI_GRU_ZR) {
      if (valid && nbase >= 128) {
        const float4* hs = reinterpret_cast<const float4*>(p.aux_f1 + static_cast<size_t>(m) * 128 + (nbase - 128));
#pragma unroll
        for (int q = 0; q < HALF / 4; ++q) {
          const float4 h = hs[q];
          pre_h[4 * q] = h.x; pre_h[4 * q + 1] = h.y; pre_h[4 * q + 2] = h.z; pre_h[4 * q + 3] = h.w;
        }
      }
    }
    if constexpr (EPI == EPI_GRU_Q) {
      if (valid) {
        const float4* hs = reinterpret_cast<const float4*>(p.aux_f1 + static_cast<size_t>(m) * 128 + nbase);
        const float4* zs = reinterpret_cast<const float4*>(p.aux_f0 + static_cast<size_t>(m) * 128 + nbase);
#pragma unroll
        for (int q = 0; q < HALF / 4; ++q) {
          const float4 h = hs[q], z = zs[q];
          pre_h[4 * q] = h.x; pre_h[4 * q + 1] = h.y; pre_h[4 * q + 2] = h.z; pre_h[4 * q + 3] = h.w;
          pre_z[4 * q] = z.x; pre_z[4 * q + 1] = z.y; pre_z[4 * q + 2] = z.z; pre_z[4 * q + 3] = z.w;
        }
      }
    }
    asm volatile("bar.sync 1, 256;");                  // s_bias visible to all epilogue warps
    if (warp == 2) G_TRACE(8);

    mbar_wait(acc_bar, 0);
    if (warp == 2) G_TRACE(9);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + half * HALF;
#pragma unroll
    for (int c = 0; c < HALF; c += CW) {
      uint32_t raw[CW];
      if constexpr (CW == 32) tmem_ld32(trow + c, raw); else tmem_ld16(trow + c, raw);
      tmem_ld_wait();
      if (valid) {
        const int n = nbase + c;                       // global column of raw[0]
        const int nl = half * HALF + c;                // column inside the CTA tile
        float v[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] = fmaf(__uint_as_float(raw[j]), p.alpha, s_bias[nl + j]);
        if constexpr (EPI == EPI_FLOW) {
          if (n == 0) {          // this thread holds (dx, dy) of its token
            float2* cp = reinterpret_cast<float2*>(p.aux_f0) + m;
            float2 c1 = *cp;
            c1.x += v[0];
            c1.y += v[1];
            *cp = c1;
            const int y = m / p.Wp, x = m - y * p.Wp;
            reinterpret_cast<float2*>(p.aux_f1)[m] = make_float2(c1.x - static_cast<float>(x), c1.y - static_cast<float>(y));
          }
        }
        if constexpr (EPI == EPI_STORE || EPI == EPI_FLOW) {
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          if (p.out_b) {
            uint4* dst = reinterpret_cast<uint4*>(p.out_b + static_cast<size_t>(m) * p.ldb + p.colb + n);
#pragma unroll
            for (int q = 0; q < CW / 8; ++q) {
              uint4 u;
              u.x = pack_act2(v[8 * q + 0], v[8 * q + 1]);
              u.y = pack_act2(v[8 * q + 2], v[8 * q + 3]);
              u.z = pack_act2(v[8 * q + 4], v[8 * q + 5]);
              u.w = pack_act2(v[8 * q + 6], v[8 * q + 7]);
              dst[q] = u;
            }
          }
          if (p.out_f) {
            float4* dst = reinterpret_cast<float4*>(p.out_f + static_cast<size_t>(m) * p.ldf + p.colf + n);
#pragma unroll
            for (int q = 0; q < CW / 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        } else if constexpr (EPI == EPI_GRU_ZR) {
          if (n < 128) {
            float4* dst = reinterpret_cast<float4*>(p.aux_f0 + static_cast<size_t>(m) * 128 + n);
#pragma unroll
            for (int q = 0; q < CW / 4; ++q)
              dst[q] = make_float4(sigmoid_fast(v[4 * q]), sigmoid_fast(v[4 * q + 1]),
                                   sigmoid_fast(v[4 * q + 2]), sigmoid_fast(v[4 * q + 3]));
          } else {
            uint4* dst = reinterpret_cast<uint4*>(p.out_b + static_cast<size_t>(m) * p.ldb + p.colb + (n - 128));
#pragma unroll
            for (int q = 0; q < CW / 8; ++q) {
              uint4 u;
              u.x = pack_act2(sigmoid_fast(v[8 * q + 0]) * pre_h[c + 8 * q + 0], sigmoid_fast(v[8 * q + 1]) * pre_h[c + 8 * q + 1]);
              u.y = pack_act2(sigmoid_fast(v[8 * q + 2]) * pre_h[c + 8 * q + 2], sigmoid_fast(v[8 * q + 3]) * pre_h[c + 8 * q + 3]);
              u.z = pack_act2(sigmoid_fast(v[8 * q + 4]) * pre_h[c + 8 * q + 4], sigmoid_fast(v[8 * q + 5]) * pre_h[c + 8 * q + 5]);
              u.w = pack_act2(sigmoid_fast(v[8 * q + 6]) * pre_h[c + 8 * q + 6], sigmoid_fast(v[8 * q + 7]) * pre_h[c + 8 * q + 7]);
              dst[q] = u;
            }
          }
        } else if constexpr (EPI == EPI_GRU_Q) {
          float4* hptr = reinterpret_cast<float4*>(p.aux_f1 + static_cast<size_t>(m) * 128 + n);
          uint4* dst = reinterpret_cast<uint4*>(p.out_b + static_cast<size_t>(m) * p.ldb + p.colb + n);
          float hn[CW];
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const float z = pre_z[c + j];
            hn[j] = (1.0f - z) * pre_h[c + j] + z * tanh_fast(v[j]);
          }
#pragma unroll
          for (int q = 0; q < CW / 4; ++q) hptr[q] = make_float4(hn[4 * q], hn[4 * q + 1], hn[4 * q + 2], hn[4 * q + 3]);
#pragma unroll
          for (int q = 0; q < CW / 8; ++q) {
            uint4 u;
            u.x = pack_act2(hn[8 * q + 0], hn[8 * q + 1]);
            u.y = pack_act2(hn[8 * q + 2], hn[8 * q + 3]);
            u.z = pack_act2(hn[8 * q + 4], hn[8 * q + 5]);
            u.w = pack_act2(hn[8 * q + 6], hn[8 * q + 7]);
            dst[q] = u;
          }
        } else if constexpr (EPI == EPI_MOTION) {
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.0f);
          if (n + CW == 128) {   // last chunk: channels 126,127 carry the flow itself
            v[CW - 2] = p.aux_f1[static_cast<size_t>(m) * 2 + 0];
            v[CW - 1] = p.aux_f1[static_cast<size_t>(m) * 2 + 1];
          }
          uint4* dst = reinterpret_cast<uint4*>(p.out_b + static_cast<size_t>(m) * p.ldb + p.colb + n);
#pragma unroll
          for (int q = 0; q < CW / 8; ++q) {
            uint4 u;
            u.x = pack_act2(v[8 * q + 0], v[8 * q + 1]);
            u.y = pack_act2(v[8 * q + 2], v[8 * q + 3]);
            u.z = pack_act2(v[8 * q + 4], v[8 * q + 5]);
            u.w = pack_act2(v[8 * q + 6], v[8 * q + 7]);
            dst[q] = u;
          }
          if (p.out_f) {
            float4* dstf = reinterpret_cast<float4*>(p.out_f + static_cast<size_t>(m) * p.ldf + p.colf + n);
#pragma unroll
            for (int q = 0; q < CW / 4; ++q) dstf[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
      __syncwarp();
    }
    if (warp == 2) G_TRACE(10);
    tc_fence_before();
  }

  // No CTA may leave while a peer can still multicast into its smem or arrive on its barriers.
  if constexpr (CL > 1) cluster_sync(); else __syncthreads();
  if (threadIdx.x == 0) G_TRACE(11);
  if (p.trace != nullptr && threadIdx.x == 0) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[16 + 3 * (blockIdx.y * gridDim.x + blockIdx.x) + 1] = static_cast<long long>(t);
  }
#undef G_TRACE
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace cb
