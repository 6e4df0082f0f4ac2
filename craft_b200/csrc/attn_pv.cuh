// craft_b200 -- flash-style P.V for the multi-mode attentions (never materialises [M,U,U]).
//
//   O_m[q, :] = sum_k  exp(S_m[q,k] - lse_m[q]) * V_m[k, :]        S_m = clamp(Q_m K_m^T / sqrt(d)) + w_pos*bias
//
// reference: CrossAttFeatTrans.forward core/setrans.py:514-557 (scores, clamp, bias, softmax) and
// ExpandedFeatTrans.forward core/setrans.py:373-383 (V = first_linear(x) split into M modes,
// out_m = P_m V_m); GMA: Attention.forward core/gma.py:78-100 + Aggregate.forward :131-134.
//
// The row log-sum-exp is known up front (scores.cuh SC_LSE), so no online rescaling is needed:
// every key tile contributes an exact, final slice of P.  The work is the list of (unit, key tile)
// pairs, unit = (128-query tile, mode); it is cut into gridDim.x equal contiguous ranges, one per
// PERSISTENT CTA (one CTA per SM), so every SM gets the same number of key tiles and the ~7 us a
// CTA needs to start up / drain (TMEM allocation, pipeline fill, O write-back) is paid once per SM
// instead of once per wave.  A CTA's range touches 2-3 units ("segments"); a unit cut by a range
// boundary is finished by the next CTA and the partial O's land in different slots of `out`
// (summed by modes_finalize_kernel).  Keys are visited in spatial blocks of 8 x (BK/8) tokens (3-D TMA box), so that the
// positional-bias window (|dy|,|dx| <= R) touches only the few blocks around the query and every
// other block takes the 3-instruction fast path.  V^T arrives in the same block order
// (gemm.cuh b_blocked).
//
// Data flow per key tile j (two softmax groups alternate tiles):
//     warp 1   : S = Q K_j^T                        tcgen05.mma SS, fp32 accumulator in TMEM
//     softmax  : P = exp2(S*c - lse) as bf16        written back to TMEM (tcgen05.st)
//     warp 2   : O += P V_j                          tcgen05.mma with the A operand read from TMEM
// P never touches shared memory: no 32 KB/tile of st.shared, no proxy fence, and the P.V MMA reads
// only V from smem.  Two TMEM layouts (template SPLIT), chosen per shape by measurement:
//   ring  (SPLIT = false): three S/P buffers, P overwrites columns of the S it came from; S(j+3) waits for
//         P.V(j) to retire.  The ring S issue -> softmax -> arrival skew -> P.V issue -> retire (~4300 clk for
//         3 tiles) staggers the two groups, which is what the 128-key tiles of the aggregator want (83 us
//         against 89 us): a lone group reaches only ~73 % of the MUFU rate, two groups in step idle together.
//   split (SPLIT = true):  S 2 x BK columns | P 2 x BK/2 | O; group g owns S[g] and P[g].  S[g] is handed back
//         as soon as the group has LOADED it, so S(j+2) is computed under tile j, and P[g] is free long before
//         it is needed.  Wins for the 64-key tiles of the F2 transformer (118 -> 104 us), whose per-tile
//         hand-offs weigh twice as much.
// (Also tried on the split layout: an "exponential token" that makes the groups take turns on the MUFU unit --
//  115 us: exclusive access runs at the lone-group rate.  profiles/r02_pv_split_token.txt.)
// The two MMA streams are issued by different warps because the instruction latency of an issuing warp
// (barrier probes, descriptor arithmetic, one UTCHMMA per 16 K columns) is what paces a tile.
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

namespace cb {

constexpr int kPvThreads = 128 + 512;   // TMA warp, S-MMA warp, PV-MMA warp, spare warp, 16 softmax warps

struct PvParams {
  Grid2 g;
  int M;                   // modes
  int nslots;              // partial-sum slots in `out` (>= CTAs that can share one unit)
  int zero_fill;           // != 0: slots a unit does not use are zero-filled (readers that sum all slots blindly)
  int nqt;                 // query tiles
  float scale;             // 1/sqrt(d)
  float w_pos;
  const float* pos_table;  // [(2R+1)^2] or nullptr
  int R;
  const float* clip;       // device scalar (+inf or attn_clip)
  const float* lse2;       // [M][Mp] log2-domain log-sum-exp
  float* out;              // [nslots][M][F/8][Mp][8] f32 partial sums (8-column chunks, see the write-back)
  int nkt, nbx;            // key tiles (blocks) in total / per block-row
  int mask_radius;         // > 0: keys farther than this (Chebyshev) from the query get probability 0 (--f2radius)
  int nostore;             // timing experiment only (CRAFT_PV_NOSTORE=1): skip the O write-back stores -- WRONG results
  long long* trace;        // CRAFT_PV_TRACE: clock64 timeline of CTA 0: [role 4][tile 64][slot 8], then globaltimer (start, end) of every CTA
};

// BULK: 32 KB of shared memory through which the O write-back goes (see the write-back below)
template <int D, int F, int BK, int KS, int VS, bool BULK = false>
struct PvSmem {
  static constexpr int kQAtoms = D > 64 ? D / 64 : 1;
  static constexpr int kQBytes = kQAtoms * 128 * 128;
  static constexpr int kKBytes = kQAtoms * BK * 128;
  static constexpr int kVBytes = (BK / 64) * F * 128;
  static constexpr int kStageBytes = BULK ? 8 * 128 * 32 : 0;       // 8 chunks x 128 rows x 8 floats
  static constexpr int kTotal = 2 * kQBytes + KS * kKBytes + VS * kVBytes + kStageBytes + 1024 + 512 + 4096;   // + align, barriers, bias table
  static_assert(kTotal <= 227 * 1024, "attn_pv: shared memory budget");
};

// TRACE: the clock64 timeline instrumentation (CRAFT_PV_TRACE) is a separate instantiation -- even predicated off,
// its ~30 instructions per tile and the registers they pin sit in the issue slots of the loops being measured.
template <int D, int F, int BK, int KS, int VS, int POLY = 0, bool MASKED = false, bool TRACE = false, bool SPLIT = false,
          bool BULK = false>
__global__ void __launch_bounds__(kPvThreads, 1)
attn_pv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ PvParams p) {
  using S = PvSmem<D, F, BK, KS, VS, BULK>;
  static_assert(!BULK || F == 128, "attn_pv: the bulk write-back is built for F = 128 (one 32-column quarter per warp)");
  constexpr int BW = BK / 8;             // block width in tokens (block height is 8)
  constexpr int HALF = BK / 2;           // S columns per softmax thread
  constexpr int PW = BK / 4;             // packed P columns (2 bf16 each) per softmax thread
  constexpr int NSB = SPLIT ? 2 : 3;     // S buffers in TMEM (split: one S and one P buffer per softmax group)
  constexpr uint32_t kTmemP = NSB * BK;  // split: P buffers (BK/2 packed columns each) follow the S buffers
  constexpr uint32_t kPOff = BK / 4;     // ring: P sits at S + BK/4: every warp overwrites only columns it has read itself
  constexpr uint32_t kTmemO = SPLIT ? NSB * BK + NSB * (BK / 2) : NSB * BK;   // O accumulator
  static_assert(kTmemO + F <= 512, "attn_pv: TMEM budget");
  static_assert(HALF == 32 || HALF == 64, "attn_pv: BK must be 64 or 128");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * S::kQBytes;
  uint8_t* sV = sK + KS * S::kKBytes;
  uint8_t* sStage = sV + VS * S::kVBytes;                 // BULK: O staging, [8 chunks][128 rows][8 floats]
  uint8_t* tail = sStage + S::kStageBytes;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);   // [2] Q of segment s lives in buffer s & 1
  uint64_t* q_free = q_full + 2;                           // [2] all S MMAs of that segment retired
  uint64_t* k_full = q_free + 2;
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;
  uint64_t* v_empty = v_full + VS;
  uint64_t* s_full = v_empty + VS;       // [NSB] S(j) complete in TMEM                      (tcgen05.commit, count 1)
  uint64_t* s_free = s_full + NSB;       // [NSB] S buffer reusable -- split: S(j) loaded into registers (8 softmax
                                         //       warps); ring: P.V(j) retired (tcgen05.commit, count 1)
  uint64_t* p_full = s_free + NSB;       // [NSB] P(j) stored to TMEM                        (8 softmax warps)
  uint64_t* p_free = p_full + NSB;       // [NSB] split only: P.V(j) retired, P buffer reusable (tcgen05.commit, count 1)
  uint64_t* o_full = p_free + NSB;       // O of the current segment complete                 (tcgen05.commit)
  uint64_t* o_free = o_full + 1;         // O read back by the 16 epilogue warps         (count 16)
  uint64_t* stage_free = o_free + 1;     // BULK: the previous user's bulk copies have read the staging buffer (count 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + 1);
  // positional-bias table, zero padded so that a thread's 4 x BW window can be read without range
  // checks: row iy+3 (iy in [-3, 2R+3]), column ix+BW-1 (ix in [-(BW-1), 2R+BW-1]); log2 domain
  float* s_table = reinterpret_cast<float*>(tail + 512);
  const int TW = 2 * p.R + 1 + 2 * (BW - 1);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  // this CTA's contiguous range of the (unit, key tile) list
  const long long NT = static_cast<long long>(p.nqt) * p.M * p.nkt;
  const long long lin_begin = NT * blockIdx.x / gridDim.x;
  const long long lin_end = NT * (blockIdx.x + 1) / gridDim.x;
  // Key tiles of a unit are visited block-COLUMN major (kt = bx * nby + by): the blocks inside the positional-
  // bias window (|dy| <= R: ~3 of the nby block rows, all columns), which cost ~1.4x a plain tile, then occur
  // 3 per nby everywhere in the list and every CTA's equal-length range carries the same share of them
  // (block-row major order packed them into one contiguous run per unit: CTA finish times spread 72..86 us).
  const int nby = p.nkt / p.nbx;
  // The range is cut into segments, one per unit it touches.  ONE thread works the list out (64-bit divisions,
  // the search for the first CTA of a shared unit) and leaves it in shared memory; every role then walks the
  // table.  Doing this arithmetic in all 512 softmax threads at every segment boundary cost ~4000 clk per
  // boundary (clock64 timeline, profiles/r02_pv_timeline_before.txt).
  struct Seg { int qt, mode, t0, nt, slot, last; };
  constexpr int kMaxSegs = 16;
  __shared__ Seg s_segs[kMaxSegs];
  __shared__ int s_nseg;
  if (threadIdx.x == 32) {          // lane 0 of warp 1; thread 0 initialises the barriers meanwhile
    // CTA that owns list position x (ranges are [NT*c/G, NT*(c+1)/G))
    auto cta_of = [&](long long x) {
      long long c = x * gridDim.x / NT;
      while (c + 1 < static_cast<long long>(gridDim.x) && NT * (c + 1) / gridDim.x <= x) ++c;
      while (c > 0 && NT * c / gridDim.x > x) --c;
      return static_cast<int>(c);
    };
    int n = 0;
    for (long long lin = lin_begin; lin < lin_end; ++n) {
      if (n == kMaxSegs) __trap();                 // the host sizes the grid so that this cannot happen
      Seg sg;
      const int unit = static_cast<int>(lin / p.nkt);
      sg.t0 = static_cast<int>(lin - static_cast<long long>(unit) * p.nkt);
      const long long left = lin_end - lin;
      sg.nt = static_cast<int>(left < p.nkt - sg.t0 ? left : p.nkt - sg.t0);
      sg.qt = unit / p.M;
      sg.mode = unit - sg.qt * p.M;
      sg.slot = static_cast<int>(blockIdx.x) - cta_of(lin - sg.t0);      // position among the CTAs sharing the unit
      sg.last = (sg.t0 + sg.nt == p.nkt) ? 1 : 0;
      s_segs[n] = sg;
      lin += sg.nt;
    }
    s_nseg = n;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int b = 0; b < 2; ++b) { mbar_init(&q_full[b], 1); mbar_init(&q_free[b], 1); }
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int b = 0; b < NSB; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_free[b], SPLIT ? 8 : 1);
      mbar_init(&p_full[b], 8);      // one arrival per softmax warp of the group
      mbar_init(&p_free[b], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, 16);
    mbar_init(stage_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();                                    // everything below reads tensors of earlier kernels
  if (p.pos_table) {
    const int TDp = 2 * p.R + 1;
    const int n = (TDp + 6) * TW;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int iy = i / TW - 3, ix = i % TW - (BW - 1);
      const bool in = (iy >= 0) && (iy < TDp) && (ix >= 0) && (ix < TDp);
      s_table[i] = in ? p.pos_table[iy * TDp + ix] * p.w_pos * 1.4426950408889634f : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = s_nseg;

  const bool tr = TRACE && p.trace != nullptr && blockIdx.x == 0;
  auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return static_cast<long long>(t); };
  if constexpr (TRACE) {
    if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 256) p.trace[2048 + 2 * blockIdx.x] = gtime();
  }
#define PV_TRACE(role, tile, slot)                                                              \
  do {                                                                                          \
    if constexpr (TRACE) {                                                                      \
      if (tr && (threadIdx.x & 31) == 0 && (tile) < 64) p.trace[((role) * 64 + (tile)) * 8 + (slot)] = clock64(); \
    }                                                                                           \
  } while (0)

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      int seg = 0, g0 = 0;
      for (; seg < nseg; ++seg) {
        const Seg sgm = s_segs[seg];
        const int ch0 = sgm.mode * D;               // first channel of this mode in the Q/K rows
        const int qk_col = (ch0 >> 6) << 6;         // TMA column of the 64-channel atom holding it
        const int qb = seg & 1;
        mbar_wait(&q_free[qb], ((static_cast<uint32_t>(seg) >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full[qb], S::kQBytes);
        for (int a = 0; a < S::kQAtoms; ++a)
          tma_load_2d(sQ + qb * S::kQBytes + a * 128 * 128, &tmQ, &q_full[qb], qk_col + a * 64, sgm.qt * 128);
        // K and V are two independent pipelines (K feeds S up to three tiles ahead of the P.V that
        // frees a V stage), so one thread serves both with non-blocking probes.
        int ki = 0, vi = 0;
        uint32_t idle = 0;
        // block coordinates of the next K / V tile, advanced incrementally (no division in the loop: it would
        // go through MUFU.RCP and queue behind the softmax warps' exponentials)
        int bx = sgm.t0 / nby, by = sgm.t0 - bx * nby;
        int vbx = bx, vby = by;
        while (ki < sgm.nt || vi < sgm.nt) {
          bool moved = false;
          if (ki < sgm.nt && mbar_test(&k_empty[ks], kph ^ 1u)) {
            mbar_arrive_expect_tx(&k_full[ks], S::kKBytes);
            for (int a = 0; a < S::kQAtoms; ++a)
              tma_load_3d(sK + ks * S::kKBytes + a * BK * 128, &tmK, &k_full[ks], qk_col + a * 64, bx * BW, by * 8);
            PV_TRACE(3, g0 + ki, 0);
            if (++ks == KS) { ks = 0; kph ^= 1u; }
            if (++by == nby) { by = 0; ++bx; }
            ++ki; moved = true;
          }
          if (vi < sgm.nt && mbar_test(&v_empty[vs], vph ^ 1u)) {
            const int vcol = (vby * p.nbx + vbx) * BK;          // V^T columns are in block-row-major block order
            mbar_arrive_expect_tx(&v_full[vs], S::kVBytes);
            for (int a = 0; a < BK / 64; ++a)
              tma_load_2d(sV + vs * S::kVBytes + a * F * 128, &tmV, &v_full[vs], vcol + a * 64, sgm.mode * F);
            PV_TRACE(3, g0 + vi, 1);
            if (++vs == VS) { vs = 0; vph ^= 1u; }
            if (++vby == nby) { vby = 0; ++vbx; }
            ++vi; moved = true;
          }
          if (moved) idle = 0;
          else {
            __nanosleep(32);
            if (++idle > (1u << 24)) __trap();     // protocol bug -> launch failure, never a hang
          }
        }
        g0 += sgm.nt;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ S = Q K^T issuer --------------------------------
    constexpr uint32_t idesc_s = umma_idesc_f16<128, BK>();
    constexpr uint64_t kKStep = static_cast<uint64_t>(S::kKBytes >> 4);
    const bool leader = elect_one();
    int ks = 0, b = 0, seg = 0, g = 0;
    uint32_t kph = 0, bpar = 1;            // s_free[b] parity for "buffer b is reusable" (first use: free)
    for (; seg < nseg; ++seg) {
      const Seg sgm = s_segs[seg];
      const uint32_t qk_inner = static_cast<uint32_t>((sgm.mode * D) & 63) * 2u;
      const int qb = seg & 1;
      const uint64_t dq0 = umma_desc_sw128(smem_u32(sQ + qb * S::kQBytes) + qk_inner);
      const uint64_t dk0 = umma_desc_sw128(smem_u32(sK) + qk_inner);
      mbar_wait(&q_full[qb], (static_cast<uint32_t>(seg) >> 1) & 1u);
      for (int i = 0; i < sgm.nt; ++i, ++g) {
        const bool r1 = mbar_try_wait_nohint(&k_full[ks], kph);
        const bool r2 = mbar_try_wait_nohint(&s_free[b], bpar);
        if (!r1) mbar_wait(&k_full[ks], kph);
        if (!r2) mbar_wait(&s_free[b], bpar);
        tc_fence_after();
        if (leader) {
          const uint32_t ts = tmem_base + static_cast<uint32_t>(b * BK);
          const uint64_t dk = dk0 + static_cast<uint64_t>(ks) * kKStep;
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const int ka = (k * 16) >> 6, kin = (k * 16) & 63;
            const uint64_t oq = static_cast<uint64_t>((ka * 128 * 128 + kin * 2) >> 4);
            const uint64_t ok = static_cast<uint64_t>((ka * BK * 128 + kin * 2) >> 4);
            if (k == 0) umma_f16(ts, dq0 + oq, dk + ok, idesc_s, 0u);
            else umma_f16_acc(ts, dq0 + oq, dk + ok, idesc_s);
          }
          umma_commit(&k_empty[ks]);
          umma_commit(&s_full[b]);
          if (i == sgm.nt - 1) umma_commit(&q_free[qb]);    // this segment's Q may be overwritten
        }
        PV_TRACE(0, g, 0);
        if (++ks == KS) { ks = 0; kph ^= 1u; }
        if (++b == NSB) { b = 0; bpar ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------ O += P V issuer (A = P from TMEM) ---------------
    constexpr uint32_t idesc_o = umma_idesc_f16<128, F>();
    constexpr uint64_t kVStep = static_cast<uint64_t>(S::kVBytes >> 4);
    const uint64_t dv0 = umma_desc_sw128(smem_u32(sV));
    const bool leader = elect_one();
    int vs = 0, b = 0, seg = 0, g = 0;
    uint32_t vph = 0, bpar = 0;
    for (; seg < nseg; ++seg) {
      const Seg sgm = s_segs[seg];
      // the previous segment's O must have been read back before its columns are overwritten
      mbar_wait(o_free, (static_cast<uint32_t>(seg) & 1u) ^ 1u);
      for (int i = 0; i < sgm.nt; ++i, ++g) {
        const bool r1 = mbar_try_wait_nohint(&p_full[b], bpar);
        const bool r2 = mbar_try_wait_nohint(&v_full[vs], vph);
        if (!r1) mbar_wait(&p_full[b], bpar);
        PV_TRACE(0, g, 2);
        if (!r2) mbar_wait(&v_full[vs], vph);
        PV_TRACE(0, g, 3);
        tc_fence_after();
        if (leader) {
          const uint32_t tp = SPLIT ? tmem_base + kTmemP + static_cast<uint32_t>(b * (BK / 2))
                                    : tmem_base + static_cast<uint32_t>(b * BK) + kPOff;
          const uint64_t dv = dv0 + static_cast<uint64_t>(vs) * kVStep;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const int ka = (k * 16) >> 6, kin = (k * 16) & 63;
            const uint64_t ov = static_cast<uint64_t>((ka * F * 128 + kin * 2) >> 4);
            // 16 bf16 of P per row = 8 packed TMEM columns per K step
            if (k == 0) umma_f16_ts(tmem_base + kTmemO, tp, dv + ov, idesc_o, i != 0 ? 1u : 0u);
            else umma_f16_ts(tmem_base + kTmemO, tp + 8u * k, dv + ov, idesc_o, 1u);
          }
          umma_commit(SPLIT ? &p_free[b] : &s_free[b]);
          umma_commit(&v_empty[vs]);
          if (i == sgm.nt - 1) umma_commit(o_full);
        }
        PV_TRACE(0, g, 4);
        if (++vs == VS) { vs = 0; vph ^= 1u; }
        if (++b == NSB) { b = 0; bpar ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------ softmax groups ----------------------------------
    // warp w (4..19): tile parity sg = ((w-4)>>2)&1, column half ch = (w-4)>>3, TMEM lane quadrant w&3
    const int sg = ((warp - 4) >> 2) & 1;
    const int ch = (warp - 4) >> 3;
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + (threadIdx.x & 31);
    const float clipv = *p.clip;
    const bool clamped = clipv < INFINITY;
    const float sc2 = p.scale * 1.4426950408889634f;
    const float clip2 = clipv * 1.4426950408889634f;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
    const int R = p.R;
    const bool has_bias = p.pos_table != nullptr;
    const int trole = TRACE ? ((warp == 4 || warp == 8) ? 1 + sg : 99) : 99;

    int seg = 0, g0 = 0;                 // g = g0 + i: CTA-wide tile counter; tile g uses buffer g % NSB, group g & 1
    float lse_next = 0.f;
    if (nseg > 0) {
      const int qn = s_segs[0].qt * 128 + row;
      lse_next = (qn < p.g.Mp) ? p.lse2[static_cast<size_t>(s_segs[0].mode) * p.g.Mp + qn] : 0.f;
    }
    Seg sg_next = s_segs[0];
    for (; seg < nseg; ++seg) {
      const Seg sgm = sg_next;
      const int q = sgm.qt * 128 + row;
      const int qy = q / p.g.Wp, qx = q - qy * p.g.Wp;
      const float lse = lse_next;
      // block coordinates of this group's tiles are advanced incrementally: an integer division by the runtime
      // nby goes through MUFU.RCP, i.e. it queues behind the other group's exponentials on the same XU pipe
      // (~450 clk per tile between a group's arrive and its next s_full wait in the clock64 timeline)
      const int i0 = (g0 & 1) ^ sg;
      int bx = (sgm.t0 + i0) / nby, by = (sgm.t0 + i0) - bx * nby;
      // this group's tiles are the ones with CTA-wide index g == sg (mod 2).  split: buffers S[sg], P[sg], whose
      // phase flips with every tile of the group; ring: buffer g % 3, advanced by two tiles per iteration
      int b = SPLIT ? sg : (g0 + i0) % NSB;
      uint32_t bpar = static_cast<uint32_t>(SPLIT ? (g0 + i0) >> 1 : (g0 + i0) / NSB) & 1u;
      for (int i = i0; i < sgm.nt; i += 2) {
        const int g = g0 + i;
        // this thread's half block: block rows [ch*4, ch*4+4), all BW columns
        const int iy0 = by * 8 + ch * 4 - qy + R;       // table row of the first block row
        const int ix0 = bx * BW - qx + R;               // table column of the first block column
        const bool near = has_bias && (iy0 + 3 >= 0) && (iy0 <= 2 * R) && (ix0 + BW - 1 >= 0) && (ix0 <= 2 * R);
        if (trole < 4) PV_TRACE(trole, g, 0);
        mbar_wait(&s_full[b], bpar);
        if (trole < 4) PV_TRACE(trole, g, 1);
        tc_fence_after();
        const uint32_t tS = tlane + static_cast<uint32_t>(b * BK + ch * HALF);
        const uint32_t tP = SPLIT ? tlane + kTmemP + static_cast<uint32_t>(b * (BK / 2) + ch * PW)
                                  : tlane + static_cast<uint32_t>(b * BK) + kPOff + static_cast<uint32_t>(ch * PW);
        uint32_t raw_all[HALF];
#pragma unroll
        for (int c = 0; c < HALF; c += 32)
          tmem_ld32(tS + c, *reinterpret_cast<uint32_t(*)[32]>(&raw_all[c]));
        tmem_ld_wait();
        if constexpr (SPLIT) {
          // S is in registers: hand the buffer back, so that S of this group's NEXT tile is computed under this one
          tc_fence_before();
          mbar_arrive_warp(&s_free[b]);
        }
        if (trole < 4) PV_TRACE(trole, g, 2);
        uint32_t pk[PW];
#pragma unroll
        for (int c = 0; c < HALF; c += 32) {
          const uint32_t* raw = &raw_all[c];
          float x[32];
          if (!clamped) {
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = fmaf(__uint_as_float(raw[e]), sc2, -lse);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              x[e] = fminf(fmaxf(__uint_as_float(raw[e]) * sc2, -clip2), clip2) - lse;
          }
          if (near) {
            const float* trow = s_table + (iy0 + 3) * TW + (ix0 + BW - 1);
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int col = c + e;                       // column inside this half: row-major (4 x BW)
              x[e] += trow[(col / BW) * TW + (col % BW)];
            }
          }
          if constexpr (MASKED) {     // --f2radius (default off): exp2(-inf) = 0, as exp(s - 1e9 - max) is in the reference
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int col = c + e;
              const int dy = by * 8 + ch * 4 + col / BW - qy, dx = bx * BW + col % BW - qx;
              if (abs(dy) > p.mask_radius || abs(dx) > p.mask_radius) x[e] = -INFINITY;
            }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e)
            pk[c / 2 + e] = pack_act2(ex2_mix<POLY>(x[2 * e], 2 * e), ex2_mix<POLY>(x[2 * e + 1], 2 * e + 1));
        }
        if (trole < 4) PV_TRACE(trole, g, 3);
        // P -> TMEM, then hand it to the P.V issuer.  split: the group's P buffer is free once the P.V of its
        // previous tile has retired -- a whole tile ago; ring: the columns of S this warp has just read
        if constexpr (SPLIT) {
          mbar_wait(&p_free[b], bpar ^ 1u);
          tc_fence_after();
        }
        if constexpr (PW == 32) tmem_st32(tP, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
        else tmem_st16(tP, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_warp(&p_full[b]);
        if (trole < 4) PV_TRACE(trole, g, 4);
        by += 2;
        while (by >= nby) { by -= nby; ++bx; }
        if constexpr (SPLIT) bpar ^= 1u;
        else { b += 2; if (b >= NSB) { b -= NSB; bpar ^= 1u; } }
      }

      // -------------------------------- O write-back of this segment ------------------------
      // the four (group, half) warp sets split the F columns in quarters.  Slot = position of this
      // CTA among the CTAs that share the unit; the CTA finishing a unit zero-fills the unused slots.
      const int slot = sgm.slot;
      const bool last_part = sgm.last != 0;
      const int g_last = g0 + sgm.nt - 1;
      // The next segment's table entry and log-sum-exp are fetched HERE: the 64 KB of write-back stores below
      // take ~2000 clk to drain from the load/store unit and any load issued behind them -- global OR shared
      // (the timeline showed ~2200 clk between "O written" and the next loop top, spent on the s_segs read) --
      // stalls the warp for that long; these complete while the warp waits for the last P.V anyway.
      if (seg + 1 < nseg) {
        sg_next = s_segs[seg + 1];
        const int qn = sg_next.qt * 128 + row;
        lse_next = (qn < p.g.Mp) ? p.lse2[static_cast<size_t>(sg_next.mode) * p.g.Mp + qn] : 0.f;
      }
      if (trole < 4) PV_TRACE(trole, g_last - (g_last & 1) + (trole - 1), 5);
      mbar_wait(o_full, static_cast<uint32_t>(seg) & 1u);
      if (trole < 4) PV_TRACE(trole, g_last - (g_last & 1) + (trole - 1), 6);
      tc_fence_after();
      __syncwarp();
      // out[slot][mode][F/8][Mp][8]: a thread (= query row) stores one 32-byte sector per chunk with a
      // single 256-bit store and consecutive lanes hit consecutive sectors, so a store instruction
      // covers 1 KB contiguous (row-major [Mp][F] would scatter 32 half sectors 4*F bytes apart and
      // the load/store unit then needs ~8000 clk per 64 KB tile -- measured, profiles/README.md)
      const size_t slot_stride = static_cast<size_t>(p.M) * p.g.Mp * F;
      float* dst = p.out + (static_cast<size_t>(slot) * p.M + sgm.mode) * p.g.Mp * F + static_cast<size_t>(q) * 8;
      constexpr int kQuarter = F / 4;
      const int c_begin = (sg * 2 + ch) * kQuarter;
      if constexpr (BULK) {
        // Write-back through shared memory and the bulk-copy engine.  Stored straight from registers, the 64 KB of
        // a segment keep the load/store unit busy for ~4400 clk (2200 to issue, 2200 more to drain, during which
        // the next barrier operation of these warps cannot issue: ~5900 clk from the last P of a segment to the
        // first tile of the next, 2.5 times per CTA -- profiles/r02_pv_timeline_ring.txt).  Here a warp reads its
        // 32 O columns (the accumulator is free again as soon as they are in registers), parks them in a 32 KB
        // staging buffer in the layout of `out` ([chunk][row][8]: 4 KB contiguous per chunk) and one thread per
        // group hands the eight chunks to cp.async.bulk.  The two groups use the buffer in turn.
        uint32_t raw[32];
        tmem_ld32(tlane + kTmemO + c_begin, raw);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(o_free);
        if (trole < 4) PV_TRACE(trole, g_last - (g_last & 1) + (trole - 1), 7);
        // use u = 2*seg + sg of the staging buffer waits for use u-1 to have been read out
        if (sg == 1) mbar_wait(stage_free, 0u);
        else if (seg > 0) mbar_wait(stage_free, 1u);
        float4* srow = reinterpret_cast<float4*>(sStage) + (static_cast<size_t>(ch * 4) * 128 + row) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          srow[static_cast<size_t>(e) * 256] = make_float4(__uint_as_float(raw[8 * e]), __uint_as_float(raw[8 * e + 1]),
                                                           __uint_as_float(raw[8 * e + 2]), __uint_as_float(raw[8 * e + 3]));
          srow[static_cast<size_t>(e) * 256 + 1] = make_float4(__uint_as_float(raw[8 * e + 4]), __uint_as_float(raw[8 * e + 5]),
                                                               __uint_as_float(raw[8 * e + 6]), __uint_as_float(raw[8 * e + 7]));
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 256;" ::"r"(1 + sg) : "memory");
        if (ch == 0 && lane_grp == 0 && (threadIdx.x & 31) == 0) {
          const int q0 = sgm.qt * 128;
          const int rows = min(128, p.g.Mp - q0);
          if (rows > 0 && !p.nostore) {
            float* gdst = p.out + (static_cast<size_t>(slot) * p.M + sgm.mode) * p.g.Mp * F + static_cast<size_t>(q0) * 8;
#pragma unroll 1
            for (int cg = 0; cg < 8; ++cg)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                               gdst + static_cast<size_t>(sg * 8 + cg) * p.g.Mp * 8),
                           "r"(smem_u32(sStage + cg * 4096)), "r"(rows * 32)
                           : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(stage_free);
        }
      } else {
#pragma unroll
      for (int c = 0; c < kQuarter; c += 32) {
        uint32_t raw[32];
        tmem_ld32(tlane + kTmemO + c_begin + c, raw);
        tmem_ld_wait();
        if (q < p.g.Mp && !p.nostore) {
#pragma unroll
          for (int e = 0; e < 4; ++e)      // one 256-bit store per chunk: consecutive lanes -> 1 KB contiguous
            st_global_v8(dst + static_cast<size_t>((c_begin + c) / 8 + e) * p.g.Mp * 8,
                         *reinterpret_cast<const uint32_t(*)[8]>(&raw[8 * e]));
        }
      }
      tc_fence_before();
      mbar_arrive_warp(o_free);
      if (trole < 4) PV_TRACE(trole, g_last - (g_last & 1) + (trole - 1), 7);
      }
      if (p.zero_fill && last_part && q < p.g.Mp) {
        for (int sl = slot + 1; sl < p.nslots; ++sl) {
#pragma unroll
          for (int e = 0; e < kQuarter / 8; ++e) {
            float4* z4 = reinterpret_cast<float4*>(dst + static_cast<size_t>(sl - slot) * slot_stride +
                                                   static_cast<size_t>(c_begin / 8 + e) * p.g.Mp * 8);
            z4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            z4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      if (trole == 1) PV_TRACE(3, g_last, 2);      // boundary: write-back (+ zero fill) done
      g0 += sgm.nt;
    }
    if constexpr (BULK) {      // the bulk copies must have completed (writes performed) before the CTA exits
      if (ch == 0 && lane_grp == 0 && (threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
#undef PV_TRACE

  __syncthreads();
  if constexpr (TRACE) {
    if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 256) p.trace[2048 + 2 * blockIdx.x + 1] = gtime();
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cb
