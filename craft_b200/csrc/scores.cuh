// craft_b200 -- multi-mode score tiles S_m = Q_m K_m^T on tcgen05, with two fused epilogues:
//
//  SC_CORR  (TransCorrBlock.update, core/corr.py:148-207 + CrossAttFeatTrans scores-only path
//            core/setrans.py:501-550 + LearnedSoftAggregate(1) core/setrans.py:289-300):
//            clamp -> soft-aggregate the M modes -> + w_pos*bias -> accumulate the global
//            layer-norm statistics (sum, sum^2) and the global raw max (clamp gate,
//            core/setrans.py:520-529) -> 2x2/4x4/8x8 average pooling of the 8x8 key block
//            -> store pyramid levels 1..3 (level 0 optional).  The level-0 volume
//            (U x U) therefore never has to reach HBM.
//  SC_LSE   (CrossAttFeatTrans softmax, core/setrans.py:540-553): per (mode, query) running
//            max / sum-exp over the key range -> partial log-sum-exp, merged by lse_merge_kernel.
//            The P.V kernel (attn_pv.cuh) then recomputes P = exp(S - lse) tile by tile.
//
// Tile: 128 queries (TMEM lanes) x 64 keys (an 8x8 spatial block fetched by a 3-D TMA box) x M
// modes (M*64 = 256 TMEM columns, double buffered).
//
// Schedule: the (query tile, key tile) list is cut into gridDim.x equal contiguous ranges, one per
// PERSISTENT CTA (one per SM): every SM gets the same number of tiles and pays the start-up once.
// A range touches 1-2 query tiles ("segments"); in SC_LSE a query tile shared by several CTAs
// leaves one partial (max, sum) per CTA in consecutive slots of lse_part, the CTA that finishes
// the query tile fills the unused slots with the neutral element.
//
// Warps (640 threads): 0 = TMA, 1 = MMA issuer (+TMEM alloc), 2..3 idle (they complete the producer warpgroup, whose
// registers setmaxnreg hands to the epilogue), 4..19 = epilogue:
// group eg = tile parity (TMEM buffer), half ch = key-block rows 0-3 / 4-7 (columns 0-31 / 32-63 of
// every mode), TMEM lane quadrant = warp & 3.  Thread = one query x 32 keys x M modes per tile.
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

namespace cb {

constexpr int kScThreads = 128 + 512;  // producer warpgroup (TMA, MMA, two idle warps) + 16 epilogue warps; 640 threads are
                                       // launched with 96 registers each, setmaxnreg then moves 64 per producer thread to
                                       // the epilogue warpgroups (32 / 112)
constexpr int kScKStages = 3;
constexpr int kScTailBytes = 512 /*barriers*/ + 2560 /*bias table*/ + 2048 /*level-3 exchange*/ + 16384 /*lse merge*/;

enum ScoreMode : int { SC_CORR = 0, SC_LSE = 1, SC_LSE_MASKED = 2 };   // MASKED: SC_LSE with the --f2radius key mask

struct ScoreParams {
  Grid2 g;                 // token grid (queries and keys share it)
  int C;                   // channels of Q/K rows (M * d)
  int M;                   // modes (1, 2, 4)
  int d;                   // per-mode dim (32, 64, 128)
  float scale;             // 1/sqrt(d)
  float w_pos;             // pos_code_weight
  const float* pos_table;  // [(2R+1)^2] f32 or nullptr
  int R;                   // pos bias radius
  const float* clip;       // device scalar: +inf (no clamp) or attn_clip
  const int* run_flag;     // optional: whole kernel is a no-op when *run_flag == 0
  int nkt_y, nkt_x;        // key tiles (8x8 blocks) in y / x
  int nqt;                 // query tiles
  int nslots;              // SC_LSE: partial slots in lse_part
  int mask_radius;         // SC_LSE: > 0 masks keys farther than this (Chebyshev) from the query (--f2radius)
  // SC_CORR
  float w_agg;             // LearnedSoftAggregate(1).feat2score.weight
  double* stat_sum;        // [2]: sum, sumsq   (atomics)
  float* stat_max;         // [1]: global max of raw scaled scores (pre-bias)
  float* lvl[4];           // pooled volumes [Mp][h_l*w_l]; lvl[0] optional
  int hl[4], wl[4];
  // level 0 in 16 bits, BLOCKED: [Mp][nkt_y*nkt_x][64] fp16 (both precision tiers), block (by, bx) = the 8x8 key
  // block of one tile, cell (y&7)*8 + (x&7) -- a thread's half block is 64 contiguous bytes (two 256-bit stores,
  // whole sectors), and a lookup window row inside a block is one 16-byte run.  104 MB at 448x1024.
  // The halves are DELTAS against the fp32 mean of their 4x8 half block (lvl0_base [Mp][nkt][2], 6.5 MB): the
  // rounding error is then relative to the local variation of the volume, not to its magnitude -- a volume
  // whose mean is 100x its standard deviation (a saturated clamp) loses nothing under the global layer-norm.
  __half* lvl0h;
  float* lvl0_base;
  long long l0_qstride;    // halves between consecutive query rows (nkt_y*nkt_x*64)
  // SC_LSE
  float2* lse_part;        // [nslots][M][Mp] (max, sumexp) natural-exp domain
};

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// SC_CORR, unclamped, M = 4 or 1: one key ROW (8 keys) of a thread's 4 x 8 window, raw accumulators of the modes ->
// soft-aggregated score.  The generic epilogue below (any M, clamp) spent 58 instructions per key, a third of them
// local-memory traffic and per-key branches on runtime constants; this form is 22 per key for M = 4, registers only:
//   softmax over modes of w*s in the exp2 domain on the UNSCALED accumulators: exponent a_m*c - ext*c (c = w*scale*
//   log2e; ext = max or, for c < 0, min of the a_m: NEG) is one FFMA per mode, the denominator is in [1, 4] (no
//   range fix-up around the reciprocal) and the 1/sqrt(d) scale is applied once to the quotient.
template <int MT, bool NEG>
__device__ __forceinline__ void corr_row_fast(uint32_t taddr, float wc2, float scale, float& rawmax, float* __restrict__ out8,
                                              bool release, uint64_t* acc_empty) {
  uint32_t r0[8], r1[8], r2[8], r3[8];
  tmem_ld8(taddr, r0);
  if (MT == 4) {
    tmem_ld8(taddr + 64, r1);
    tmem_ld8(taddr + 128, r2);
    tmem_ld8(taddr + 192, r3);
  }
  tmem_ld_wait();
  if (release) {       // last row: the TMEM buffer is drained -> back to the MMA warp before the math
    tc_fence_before();
    mbar_arrive_warp(acc_empty);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a0 = __uint_as_float(r0[j]);
    if (MT == 1) {
      rawmax = fmaxf(rawmax, a0);
      out8[j] = a0 * scale;
    } else {
      const float a1 = __uint_as_float(r1[j]), a2 = __uint_as_float(r2[j]), a3 = __uint_as_float(r3[j]);
      const float amax = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
      rawmax = fmaxf(rawmax, amax);
      const float ext = NEG ? fminf(fminf(a0, a1), fminf(a2, a3)) : amax;
      const float ntm = -ext * wc2;
      const float e0 = fast_ex2(fmaf(a0, wc2, ntm)), e1 = fast_ex2(fmaf(a1, wc2, ntm));
      const float e2 = fast_ex2(fmaf(a2, wc2, ntm)), e3 = fast_ex2(fmaf(a3, wc2, ntm));
      const float num = fmaf(e3, a3, fmaf(e2, a2, fmaf(e1, a1, e0 * a0)));
      const float den = (e0 + e1) + (e2 + e3);
      out8[j] = (num * fast_rcp(den)) * scale;
    }
  }
}

// Any M (1, 2, 4), with or without the clamp (core/setrans.py:520-529): the rare configurations and the clamped
// re-pass.  Same row-at-a-time shape as corr_row_fast (8 accumulator columns per mode in flight).
__device__ __forceinline__ void corr_row_generic(uint32_t taddr, int M, float wl2, float scale, bool clamped, float clipv,
                                              float& rawmax, float* __restrict__ out8, bool release, uint64_t* acc_empty) {
  uint32_t r0[8], r1[8], r2[8], r3[8];
  tmem_ld8(taddr, r0);
  if (M > 1) tmem_ld8(taddr + 64, r1);
  if (M > 2) {
    tmem_ld8(taddr + 128, r2);
    tmem_ld8(taddr + 192, r3);
  }
  tmem_ld_wait();
  if (release) {
    tc_fence_before();
    mbar_arrive_warp(acc_empty);
  }
  const float lim = clamped ? clipv : INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a0 = __uint_as_float(r0[j]);
    // M == 2: the unused modes duplicate mode 0/1, so that neither the max nor the (signed!) soft-aggregation
    // weight ever sees an infinity; M == 1: all four are mode 0 and the aggregate is s0 itself
    const float a1 = (M > 1) ? __uint_as_float(r1[j]) : a0;
    const float a2 = (M > 2) ? __uint_as_float(r2[j]) : a0;
    const float a3 = (M > 2) ? __uint_as_float(r3[j]) : a1;
    rawmax = fmaxf(rawmax, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
    const float s0 = fminf(fmaxf(a0 * scale, -lim), lim), s1 = fminf(fmaxf(a1 * scale, -lim), lim);
    const float s2 = fminf(fmaxf(a2 * scale, -lim), lim), s3 = fminf(fmaxf(a3 * scale, -lim), lim);
    // softmax over modes of w*s (the Linear(1,1) bias cancels), exp2 domain
    const float t0 = s0 * wl2, t1 = s1 * wl2, t2 = s2 * wl2, t3 = s3 * wl2;
    const float tm = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
    const float e0 = fast_ex2(t0 - tm), e1 = fast_ex2(t1 - tm), e2 = fast_ex2(t2 - tm), e3 = fast_ex2(t3 - tm);
    const float num = (e0 * s0 + e1 * s1) + (e2 * s2 + e3 * s3);
    const float den = (e0 + e1) + (e2 + e3);
    out8[j] = (M == 1) ? s0 : num * fast_rcp(den);
  }
}

// smem: Q tile (C/64 atoms x 16 KB) + K stages (C/64 atoms x 8 KB each) + tail
template <int MODE_>
__global__ void __launch_bounds__(kScThreads, 1)
scores_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ ScoreParams p) {
  constexpr int MODE = (MODE_ == SC_LSE_MASKED) ? SC_LSE : MODE_;
  constexpr bool masking = (MODE_ == SC_LSE_MASKED);
  pdl_launch_dependents();
  if (p.run_flag != nullptr) {
    pdl_wait();                                             // the flag is written by the preceding gate kernel
    if (*p.run_flag == 0) return;                           // clamp re-pass not needed (uniform)
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int atoms = p.C / 64;              // 64-channel swizzle atoms per row
  const int q_bytes = atoms * 128 * 128;   // 128 queries
  const int k_bytes = atoms * 64 * 128;    // 64 keys
  uint8_t* sQ = smem;
  uint8_t* sK = smem + q_bytes;
  uint8_t* tail = sK + kScKStages * k_bytes;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* q_free = q_full + 1;               // all MMAs of the segment retired: Q may be overwritten
  uint64_t* k_full = q_free + 1;
  uint64_t* k_empty = k_full + kScKStages;
  uint64_t* acc_full = k_empty + kScKStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2] count 8 (warps of the group)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // positional-bias table, zero padded so that a thread's 4 x 8 window needs no range checks:
  // row iy+4 (iy in [-4, 2R+3]; rows 0..3 are all zero and double as the window of lanes that are
  // outside the bias range), column ix+7 (ix in [-7, 2R+7]); natural domain, times w_pos
  float* s_table = reinterpret_cast<float*>(tail + 512);
  float* s_l3 = reinterpret_cast<float*>(tail + 512 + 2560);             // [2 slots][2 eg][128 rows]
  float2* s_merge = reinterpret_cast<float2*>(tail + 512 + 2560 + 2048);  // [4 sets][4 modes][128 rows]
  const int TW = 2 * p.R + 1 + 14;

  const int warp = threadIdx.x >> 5;
  const int nkt = p.nkt_y * p.nkt_x;
  // this CTA's contiguous range of the (query tile, key tile) list
  // 32-bit tile indices (the host rejects grids with more than 2^31 / gridDim tiles): the 64-bit range variables
  // cost the epilogue warps four registers and two spills in the tile loop
  const int NT = p.nqt * nkt;
  const int lin_begin = static_cast<int>(static_cast<long long>(NT) * blockIdx.x / gridDim.x);
  const int lin_end = static_cast<int>(static_cast<long long>(NT) * (blockIdx.x + 1) / gridDim.x);
  auto cta_of = [&](int x) {
    long long c = static_cast<long long>(x) * gridDim.x / NT;
    while (c + 1 < static_cast<long long>(gridDim.x) && static_cast<long long>(NT) * (c + 1) / gridDim.x <= x) ++c;
    while (c > 0 && static_cast<long long>(NT) * c / gridDim.x > x) --c;
    return static_cast<int>(c);
  };
  struct Seg { int qt, t0, nt; };
  auto seg_at = [&](int lin) {
    Seg s;
    s.qt = lin / nkt;
    s.t0 = lin - s.qt * nkt;
    const int left = lin_end - lin;
    s.nt = left < nkt - s.t0 ? left : nkt - s.t0;
    return s;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    mbar_init(q_full, 1);
    mbar_init(q_free, 1);
    for (int s = 0; s < kScKStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);       // one arrival per epilogue warp of the group
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();                                    // everything below reads tensors of earlier kernels
  if (p.pos_table) {
    const int TDp = 2 * p.R + 1;
    const int n = (TDp + 7) * TW;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int iy = i / TW - 4, ix = i % TW - 7;
      const bool in = (iy >= 0) && (iy < TDp) && (ix >= 0) && (ix < TDp);
      s_table[i] = in ? p.pos_table[iy * TDp + ix] * p.w_pos : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      int stage = 0, seg = 0;
      uint32_t phase = 0;
      for (int lin = lin_begin; lin < lin_end; ++seg) {
        const Seg sg = seg_at(lin);
        mbar_wait(q_free, (static_cast<uint32_t>(seg) & 1u) ^ 1u);       // previous segment's MMAs retired
        mbar_arrive_expect_tx(q_full, q_bytes);
        for (int a = 0; a < atoms; ++a) tma_load_2d(sQ + a * 128 * 128, &tmQ, q_full, a * 64, sg.qt * 128);
        // block-column major key tiles (see attn_pv.cuh); coordinates advanced without a division (an integer
        // division by a runtime value goes through MUFU.RCP and queues behind the epilogue warps' exponentials)
        int bx = sg.t0 / p.nkt_y, by = sg.t0 - bx * p.nkt_y;
        for (int i = 0; i < sg.nt; ++i) {
          mbar_wait(&k_empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&k_full[stage], k_bytes);
          uint8_t* dst = sK + stage * k_bytes;
          for (int a = 0; a < atoms; ++a)
            tma_load_3d(dst + a * 64 * 128, &tmK, &k_full[stage], a * 64, bx * 8, by * 8);
          if (++stage == kScKStages) { stage = 0; phase ^= 1u; }
          if (++by == p.nkt_y) { by = 0; ++bx; }
        }
        lin += sg.nt;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer --------------------------------------
    // lean issue loop (see gemm.cuh): constant-add descriptors, both barriers probed up front
    constexpr uint32_t idesc = umma_idesc_f16<128, 64>();
    const bool leader = elect_one();
    const uint64_t dq0 = umma_desc_sw128(smem_u32(sQ));
    const uint64_t dk0 = umma_desc_sw128(smem_u32(sK));
    const uint64_t k_step = static_cast<uint64_t>(k_bytes >> 4);
    const int ksteps = p.d / 16;
    int stage = 0, seg = 0, g = 0;
    uint32_t phase = 0;
    for (int lin = lin_begin; lin < lin_end; ++seg) {
      const Seg sg = seg_at(lin);
      mbar_wait(q_full, static_cast<uint32_t>(seg) & 1u);
      for (int i = 0; i < sg.nt; ++i, ++g) {
        const int b = g & 1;
        const uint32_t par = ((static_cast<uint32_t>(g) >> 1) & 1u) ^ 1u;
        const bool r1 = mbar_try_wait_nohint(&k_full[stage], phase);
        const bool r2 = mbar_try_wait_nohint(&acc_empty[b], par);
        if (!r1) mbar_wait(&k_full[stage], phase);
        if (!r2) mbar_wait(&acc_empty[b], par);
        tc_fence_after();
        if (leader) {
          const uint64_t dk = dk0 + static_cast<uint64_t>(stage) * k_step;
          for (int m = 0; m < p.M; ++m) {
            const int ch0 = m * p.d;                 // first channel of this mode
            const int atom = ch0 >> 6;
            const uint32_t inner = static_cast<uint32_t>(ch0 & 63) * 2u;   // byte offset inside the 128-B row
            const uint64_t dqm = dq0 + static_cast<uint64_t>((atom * 128 * 128 + inner) >> 4);
            const uint64_t dkm = dk + static_cast<uint64_t>((atom * 64 * 128 + inner) >> 4);
            const uint32_t tcol = tmem_base + static_cast<uint32_t>(b * 256 + m * 64);
            for (int k = 0; k < ksteps; ++k) {
              // K steps past the first 64 channels of a mode (d = 128) move to the next atom.
              const int ka = (k * 16) >> 6, kin = (k * 16) & 63;
              const uint64_t oq = static_cast<uint64_t>((ka * 128 * 128 + kin * 2) >> 4);
              const uint64_t ok = static_cast<uint64_t>((ka * 64 * 128 + kin * 2) >> 4);
              umma_f16(tcol, dqm + oq, dkm + ok, idesc, k != 0 ? 1u : 0u);
            }
          }
          umma_commit(&k_empty[stage]);
          umma_commit(&acc_full[b]);
          if (i == sg.nt - 1) umma_commit(q_free);
        }
        if (++stage == kScKStages) { stage = 0; phase ^= 1u; }
      }
      lin += sg.nt;
    }
    __syncwarp();
  }
  } else {
    // ------------------------------------ epilogue ----------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int eg = ((warp - 4) >> 2) & 1;           // epilogue group <-> TMEM buffer / tile parity
    const int ch = (warp - 4) >> 3;                 // key-block rows ch*4 .. ch*4+3
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + (threadIdx.x & 31);
    const float clipv = *p.clip;
    const bool clamped = clipv < INFINITY;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + eg * 256 + ch * 32;
    const int R = p.R;
    const bool has_bias = p.pos_table != nullptr;
    const uint32_t s_table_u32 = smem_u32(s_table);
    constexpr float kLog2e = 1.4426950408889634f;

    float st_sum = 0.f, st_sq = 0.f, st_rawmax = -INFINITY;   // st_rawmax: max of the UNSCALED accumulators
    const float wl2 = p.w_agg * kLog2e;
    const float sc2 = p.scale * kLog2e;
    const float wc2 = p.scale * wl2;             // exponent per unit of raw accumulator
    const bool fast_corr = (MODE == SC_CORR) && !clamped && (p.M == 4 || p.M == 1);

    int seg = 0, g0 = 0;
    for (int lin = lin_begin; lin < lin_end; ++seg) {
      const Seg sgm = seg_at(lin);
      const int q = sgm.qt * 128 + row;
      const int qy = q / p.g.Wp, qx = q - qy * p.g.Wp;
      const bool qvalid = (q < p.g.Mp) && (qx < p.g.W);
      if constexpr (MODE == SC_LSE) {     // this thread's running (max, sum) per mode: its own smem cells
        for (int m = 0; m < 4; ++m) s_merge[((eg * 2 + ch) * 4 + m) * 128 + row] = make_float2(-INFINITY, 0.f);
      }
      float seg_rawmax = -INFINITY;              // merged into st_rawmax only for real query rows

      const int i0 = (g0 & 1) ^ eg;
      int bx = (sgm.t0 + i0) / p.nkt_y, by = (sgm.t0 + i0) - bx * p.nkt_y;      // block-column major, incremental
      auto advance = [&] { by += 2; while (by >= p.nkt_y) { by -= p.nkt_y; ++bx; } };
      for (int i = i0; i < sgm.nt; i += 2, advance()) {
        const int g = g0 + i;
        const uint32_t par = (static_cast<uint32_t>(g) >> 1) & 1u;
        const int ky0 = by * 8 + ch * 4, kx0 = bx * 8;          // this thread's 4 x 8 key window
        const int iy0 = ky0 - qy + R, ix0 = kx0 - qx + R;       // table coordinates of its first key
        const bool near = has_bias && (iy0 + 3 >= 0) && (iy0 <= 2 * R) && (ix0 + 7 >= 0) && (ix0 <= 2 * R);
        const bool full = (ky0 + 4 <= p.g.H) && (kx0 + 8 <= p.g.W);
        // lanes outside the bias range read the all-zero rows, so the window path can be taken by the
        // whole warp at once (no divergence between the plain and the window path)
        // 32-bit shared-memory address + ld.shared: a generic pointer derived from the aligned smem base costs two
        // registers (it was one of the values spilled in this loop) and a generic LD per table read
        const uint32_t tab_addr = s_table_u32 + 4u * static_cast<uint32_t>(near ? (iy0 + 4) * TW + (ix0 + 7) : 0);
        auto trow_tab = [&](int idx) {
          float v;
          asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(tab_addr + 4u * static_cast<uint32_t>(idx)));
          return v;
        };
        const bool near_w = __any_sync(0xffffffffu, near);
        mbar_wait(&acc_full[eg], par);
        tc_fence_after();

        if constexpr (MODE == SC_CORR) {
          float agg[32];
          if (fast_corr) {
            // warp-uniform choice of one specialised row routine (runtime constants of the launch)
            if (p.M == 1) {
#pragma unroll
              for (int r = 0; r < 4; ++r) corr_row_fast<1, false>(trow + 8 * r, wc2, p.scale, seg_rawmax, agg + 8 * r, r == 3, &acc_empty[eg]);
            } else if (wc2 >= 0.f) {
#pragma unroll
              for (int r = 0; r < 4; ++r) corr_row_fast<4, false>(trow + 8 * r, wc2, p.scale, seg_rawmax, agg + 8 * r, r == 3, &acc_empty[eg]);
            } else {
#pragma unroll
              for (int r = 0; r < 4; ++r) corr_row_fast<4, true>(trow + 8 * r, wc2, p.scale, seg_rawmax, agg + 8 * r, r == 3, &acc_empty[eg]);
            }
          } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
              corr_row_generic(trow + 8 * r, p.M, wl2, p.scale, clamped, clipv, seg_rawmax, agg + 8 * r, r == 3, &acc_empty[eg]);
          }

          if (qvalid) {
            if (near_w) {      // warp-uniform: lanes outside the range add the zero rows
#pragma unroll
              for (int e = 0; e < 32; ++e) agg[e] += trow_tab((e >> 3) * TW + (e & 7));
            }
            if (full) {
#pragma unroll
              for (int e = 0; e < 32; ++e) { st_sum += agg[e]; st_sq = fmaf(agg[e], agg[e], st_sq); }
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                if ((ky0 + (e >> 3) < p.g.H) && (kx0 + (e & 7) < p.g.W)) { st_sum += agg[e]; st_sq = fmaf(agg[e], agg[e], st_sq); }
              }
            }
            // level 0 in fp32, row-major (optional, debugging / SAVECORR)
            if (p.lvl[0]) {
              float* dst = p.lvl[0] + static_cast<size_t>(q) * (p.hl[0] * p.wl[0]);
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const int yy = ky0 + (e >> 3), xx = kx0 + (e & 7);
                if (yy < p.hl[0] && xx < p.wl[0]) dst[yy * p.wl[0] + xx] = agg[e];
              }
            }
          }
          // level 1: 2 x 4 cells, level 2: 1 x 2 cells of this half (floor-mode avg_pool2d chain, corr.py:186-189)
          float l1[8];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              l1[r * 4 + c] = 0.25f * (agg[(2 * r) * 8 + 2 * c] + agg[(2 * r) * 8 + 2 * c + 1] +
                                       agg[(2 * r + 1) * 8 + 2 * c] + agg[(2 * r + 1) * 8 + 2 * c + 1]);
          if (qvalid) {
            const int w1 = p.wl[1];
            float* dst = p.lvl[1] + static_cast<size_t>(q) * (p.hl[1] * w1);
            const int yy0 = by * 4 + ch * 2, xx0 = bx * 4;
            const bool vec = ((w1 & 3) == 0) && (((p.hl[1] * w1) & 3) == 0) && (xx0 + 4 <= w1);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int yy = yy0 + r;
              if (yy >= p.hl[1]) continue;
              if (vec) {
                *reinterpret_cast<float4*>(dst + yy * w1 + xx0) = make_float4(l1[r * 4], l1[r * 4 + 1], l1[r * 4 + 2], l1[r * 4 + 3]);
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if (xx0 + c < w1) dst[yy * w1 + xx0 + c] = l1[r * 4 + c];
              }
            }
          }
          float l2[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) l2[c] = 0.25f * (l1[2 * c] + l1[2 * c + 1] + l1[4 + 2 * c] + l1[4 + 2 * c + 1]);
          if (qvalid && p.lvl0h) {
            const float base = 0.5f * (l2[0] + l2[1]);      // mean of this thread's 4x8 half block (its two 4x4 cells)
            uint32_t w16[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const __half2 h2 = __floats2half2_rn(agg[2 * e] - base, agg[2 * e + 1] - base);
              w16[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            const size_t blk = static_cast<size_t>(by * p.nkt_x + bx);
            __half* dst = p.lvl0h + static_cast<size_t>(q) * p.l0_qstride + (blk * 64 + ch * 32);
            st_global_v8(dst, *reinterpret_cast<const uint32_t(*)[8]>(&w16[0]));
            st_global_v8(dst + 16, *reinterpret_cast<const uint32_t(*)[8]>(&w16[8]));
            p.lvl0_base[(static_cast<size_t>(q) * (p.l0_qstride >> 6) + blk) * 2 + ch] = base;
          }
          if (qvalid) {
            const int w2 = p.wl[2];
            float* dst = p.lvl[2] + static_cast<size_t>(q) * (p.hl[2] * w2);
            const int yy = by * 2 + ch, xx0 = bx * 2;
            if (yy < p.hl[2]) {
              if (((w2 & 1) == 0) && (((p.hl[2] * w2) & 1) == 0) && (xx0 + 2 <= w2))
                *reinterpret_cast<float2*>(dst + yy * w2 + xx0) = make_float2(l2[0], l2[1]);
              else {
                if (xx0 < w2) dst[yy * w2 + xx0] = l2[0];
                if (xx0 + 1 < w2) dst[yy * w2 + xx0 + 1] = l2[1];
              }
            }
          }
          // level 3 needs both halves of the block: the lower half hands its share over through smem
          // (pair barrier: the two warps of one (group, lane quadrant); ids 2..9)
          const float half3 = 0.25f * (l2[0] + l2[1]);
          float* xs = s_l3 + ((g >> 1) & 1) * 256 + eg * 128 + row;
          if (ch == 1) *xs = half3;
          asm volatile("bar.sync %0, 64;" ::"r"(2 + eg * 4 + lane_grp) : "memory");
          if (ch == 0 && qvalid && by < p.hl[3] && bx < p.wl[3])
            p.lvl[3][static_cast<size_t>(q) * (p.hl[3] * p.wl[3]) + by * p.wl[3] + bx] = half3 + *xs;
        } else {
          // ------------------------------ SC_LSE ------------------------------
          // one mode at a time: 32 accumulator registers live (two modes in flight spill at 96 regs/thread)
          // (the mode loop stays rolled: unrolled, the tile loop no longer fits the instruction cache;
          //  the running (max, sum) of each mode therefore lives in shared memory, not in an indexed
          //  register array)
#pragma unroll 1
          for (int m = 0; m < p.M; ++m) {
            {
              float2* rs = s_merge + ((eg * 2 + ch) * 4 + m) * 128 + row;
              const float2 run = *rs;
              float run_m_new = run.x, run_l_new = run.y;
              uint32_t raw[32];
              tmem_ld32(trow + m * 64, raw);
              tmem_ld_wait();
              if (m == p.M - 1) {
                tc_fence_before();
                mbar_arrive_warp(&acc_empty[eg]);
              }
              float rmax = __uint_as_float(raw[0]);
#pragma unroll
              for (int e = 1; e < 32; ++e) rmax = fmaxf(rmax, __uint_as_float(raw[e]));
              seg_rawmax = fmaxf(seg_rawmax, rmax);
              if (!near_w && !clamped && full && !masking) {
                // fast path: exp2(raw * scale*log2e - max*log2e), one FFMA + one MUFU per key
                const float nm = fmaxf(run.x, rmax * p.scale);
                const float nm2 = nm * kLog2e;
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                  acc0 += fast_ex2(fmaf(__uint_as_float(raw[e]), sc2, -nm2));
                  acc1 += fast_ex2(fmaf(__uint_as_float(raw[e + 1]), sc2, -nm2));
                }
                run_l_new = run.y * fast_ex2((run.x - nm) * kLog2e) + (acc0 + acc1);
                run_m_new = nm;
              } else if (!clamped && full && !masking) {
                // near the query: the positional bias enters the max; values are built in place
                float x[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) x[e] = fmaf(__uint_as_float(raw[e]), p.scale, trow_tab((e >> 3) * TW + (e & 7)));
                float tmax = x[0];
#pragma unroll
                for (int e = 1; e < 32; ++e) tmax = fmaxf(tmax, x[e]);
                const float nm = fmaxf(run.x, tmax);
                const float nm2 = nm * kLog2e;
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                  acc0 += fast_ex2(fmaf(x[e], kLog2e, -nm2));
                  acc1 += fast_ex2(fmaf(x[e + 1], kLog2e, -nm2));
                }
                run_l_new = run.y * fast_ex2((run.x - nm) * kLog2e) + (acc0 + acc1);
                run_m_new = nm;
              } else {
                // rare path (clamped re-pass or a ragged edge block): the value of a key is
                // recomputed in both passes instead of being kept in a second 32-register array
                auto val = [&](int e) {
                  float s = fminf(fmaxf(__uint_as_float(raw[e]) * p.scale, -clipv), clipv);
                  if (near) s += trow_tab((e >> 3) * TW + (e & 7));
                  if (!full && !((ky0 + (e >> 3) < p.g.H) && (kx0 + (e & 7) < p.g.W))) s = -INFINITY;
                  // --f2radius: the reference adds -1e9 (core/setrans.py:583), whose exp is exactly 0 in fp32
                  if (masking && (abs(ky0 + (e >> 3) - qy) > p.mask_radius || abs(kx0 + (e & 7) - qx) > p.mask_radius)) s = -INFINITY;
                  return s;
                };
                float tmax = val(0);
#pragma unroll
                for (int e = 1; e < 32; ++e) tmax = fmaxf(tmax, val(e));
                const float nm = fmaxf(run.x, tmax);
                if (nm > -INFINITY) {
                  float acc = 0.f;
#pragma unroll
                  for (int e = 0; e < 32; ++e) acc += fast_ex2((val(e) - nm) * kLog2e);
                  run_l_new = run.y * fast_ex2((run.x - nm) * kLog2e) + acc;
                  run_m_new = nm;
                }
              }
              *rs = make_float2(run_m_new, run_l_new);
            }
          }
        }
      }

      // ------------------------- end of segment: SC_LSE partial of this query tile ----------
      if constexpr (MODE == SC_LSE) {
        // the four (group, half) sets saw disjoint keys: merge through shared memory
        const int set = eg * 2 + ch;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (set == 0 && q < p.g.Mp) {
          const int slot = static_cast<int>(blockIdx.x) - cta_of(lin - sgm.t0);
          const bool last_part = (sgm.t0 + sgm.nt == nkt);
          for (int m = 0; m < p.M; ++m) {
            float nm = -INFINITY;
#pragma unroll
            for (int s = 0; s < 4; ++s) nm = fmaxf(nm, s_merge[(s * 4 + m) * 128 + row].x);
            float l = 0.f;
            if (nm > -INFINITY) {
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const float2 o = s_merge[(s * 4 + m) * 128 + row];
                if (o.x > -INFINITY) l += o.y * __expf(o.x - nm);
              }
            }
            p.lse_part[(static_cast<size_t>(slot) * p.M + m) * p.g.Mp + q] = make_float2(nm, l);
            if (last_part)
              for (int sl = slot + 1; sl < p.nslots; ++sl)
                p.lse_part[(static_cast<size_t>(sl) * p.M + m) * p.g.Mp + q] = make_float2(-INFINITY, 0.f);
          }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");   // s_merge is reused by the next segment
      }
      if (qvalid) st_rawmax = fmaxf(st_rawmax, seg_rawmax);
      lin += sgm.nt;
      g0 += sgm.nt;
    }

    // ------------------------- per-CTA reductions -----------------------------------------
    // global max of raw scores (clamp gate) -- both modes
    {
      const float wm = warp_max(st_rawmax) * p.scale;
      if ((threadIdx.x & 31) == 0 && wm > -INFINITY) atomic_max_float(p.stat_max, wm);
    }
    if constexpr (MODE == SC_CORR) {
      const float ws = warp_sum(st_sum), wq = warp_sum(st_sq);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&p.stat_sum[0], static_cast<double>(ws));
        atomicAdd(&p.stat_sum[1], static_cast<double>(wq));
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// lse_merge: combine the partial (max, sumexp) pairs of all slots -> lse in the log2 domain.
//   lse2[m][q] = log2( sum_k exp(s_k) ) = (mx + ln(sum)) * log2(e)
__global__ void lse_merge_kernel(const float2* __restrict__ part, int ksplit, int M, int Mp,
                                 float* __restrict__ lse2) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * Mp) return;
  float mx = -INFINITY;
  for (int s = 0; s < ksplit; ++s) mx = fmaxf(mx, part[static_cast<size_t>(s) * M * Mp + i].x);
  float l = 0.f;
  for (int s = 0; s < ksplit; ++s) {
    const float2 v = part[static_cast<size_t>(s) * M * Mp + i];
    if (v.x > -INFINITY) l += v.y * __expf(v.x - mx);
  }
  lse2[i] = (l > 0.f) ? (mx + __logf(l)) * 1.4426950408889634f : 0.f;
}

// corr_stats_finalize: {sum, sumsq} over n elements -> {mean, rstd} (biased variance, eps 1e-12,
// F.layer_norm semantics, core/corr.py:202) and the clamp gate: clip = (max > attn_clip) ? attn_clip : +inf.
// sums: [2][2] -- slot 0 from the unclamped pass, slot 1 from the clamped re-pass, selected by *flag.
__global__ void corr_stats_finalize_kernel(const double* sums, const int* flag, double n, float* mean_rstd) {
  pdl_launch_dependents();
  pdl_wait();
  if (flag != nullptr && *reinterpret_cast<const volatile int*>(flag) != 0) sums += 2;
  const volatile double* vs = sums;          // no invariant loads: see clip_gate_kernel
  const double mean = vs[0] / n;
  double var = vs[1] / n - mean * mean;
  if (var < 0) var = 0;
  mean_rstd[0] = static_cast<float>(mean);
  mean_rstd[1] = static_cast<float>(1.0 / sqrt(var + 1e-12));
}
// diag (optional, [2] f32): the module's running diagnostics {max_attn, clamp_count} that the reference
// keeps on the host with two .item() syncs per call (core/setrans.py:520-529); here they stay on the device
// and are read lazily.
// NOTE (found under CUDA-graph replay in round 2): the inputs are deliberately NOT `const __restrict__`.
// nvcc marks loads through such pointers invariant and hoisted this kernel's only load above
// griddepcontrol.wait, so the gate read the score maximum before the scores kernel had produced it.
// profiles/audit_pdl_hoist.py (run by tests/test_capi_symbols.py) checks the SASS of every kernel for this.
__global__ void clip_gate_kernel(const float* stat_max, float attn_clip, float* clip, int* flag, float* diag) {
  pdl_launch_dependents();
  pdl_wait();
  const float mx = *reinterpret_cast<const volatile float*>(stat_max);
  const bool hit = mx > attn_clip;
  clip[0] = hit ? attn_clip : INFINITY;
  flag[0] = hit ? 1 : 0;
  if (diag) {
    diag[0] = fmaxf(diag[0], mx);
    if (hit) diag[1] += 1.0f;
  }
}

}  // namespace cb
