// craft_b200 -- flash-style P.V for the multi-mode attentions (never materialises [M,U,U]).
//
//   O_m[q, :] = sum_k  exp(S_m[q,k] - lse_m[q]) * V_m[k, :]        S_m = clamp(Q_m K_m^T / sqrt(d)) + w_pos*bias
//
// reference: CrossAttFeatTrans.forward core/setrans.py:514-557 (scores, clamp, bias, softmax) and
// ExpandedFeatTrans.forward core/setrans.py:373-383 (V = first_linear(x) split into M modes,
// out_m = P_m V_m); GMA: Attention.forward core/gma.py:78-100 + Aggregate.forward :131-134.
//
// The row log-sum-exp is known up front (scores.cuh SC_LSE), so no online rescaling is needed:
// every key tile contributes an exact, final slice of P.  One CTA = (128 queries, one mode, one
// key split).  Keys are visited in spatial blocks of 8 x (BK/8) tokens (3-D TMA box), so that the
// positional-bias window (|dy|,|dx| <= R) touches only the few blocks around the query and every
// other block takes the 3-instruction fast path.  V^T arrives in the same block order
// (gemm.cuh b_blocked).  Pipeline per key tile j:
//     MMA warp : S[j&1] = Q K_j^T                     (tcgen05, accumulator in TMEM)
//     softmax  : P[j&1] = exp2(...) as bf16 -> smem    (4 groups of 4 warps: tile parity x column half)
//     MMA warp : O += P[j&1] V_j                       (A = P from smem, B = V^T tile from TMA)
// S(j+2) is issued before PV(j): neither the tensor pipe nor the two softmax groups wait for each other.
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

namespace cb {

constexpr int kPvThreads = 64 + 512;   // TMA warp, MMA warp, 16 softmax warps

struct PvParams {
  Grid2 g;
  int M;                   // modes
  int ksplit;              // gridDim.z
  float scale;             // 1/sqrt(d)
  float w_pos;
  const float* pos_table;  // [(2R+1)^2] or nullptr
  int R;
  const float* clip;       // device scalar (+inf or attn_clip)
  const float* lse2;       // [M][Mp] log2-domain log-sum-exp
  float* out;              // [ksplit][M][Mp][F] f32 partial sums
  int nkt, nbx;            // key tiles (blocks) in total / per block-row
  long long* trace;        // CRAFT_PV_TRACE: clock64 timeline of CTA (0,0,0): [role 4][tile 64][slot 8]
  int dbg;                 // timing experiments only (CRAFT_PV_DBG): bit0 no exp, bit1 no TMEM ld, bit2 no P store, bit3 no PV MMA, bit4 no S MMA
};

template <int D, int F, int BK, int KS, int VS>
struct PvSmem {
  static constexpr int kQAtoms = D > 64 ? D / 64 : 1;
  static constexpr int kQBytes = kQAtoms * 128 * 128;
  static constexpr int kKBytes = kQAtoms * BK * 128;
  static constexpr int kVBytes = (BK / 64) * F * 128;
  static constexpr int kPBytes = (BK / 64) * 128 * 128;
  static constexpr int kTotal = kQBytes + KS * kKBytes + VS * kVBytes + 2 * kPBytes + 1024 + 2048;
};

template <int D, int F, int BK, int KS, int VS>
__global__ void __launch_bounds__(kPvThreads, 1)
attn_pv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ PvParams p) {
  using S = PvSmem<D, F, BK, KS, VS>;
  constexpr int BW = BK / 8;             // block width in tokens (block height is 8)
  constexpr int HALF = BK / 2;           // columns per softmax thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + S::kQBytes;
  uint8_t* sV = sK + KS * S::kKBytes;
  uint8_t* sP = sV + VS * S::kVBytes;
  uint8_t* tail = sP + 2 * S::kPBytes;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;
  uint64_t* v_empty = v_full + VS;
  uint64_t* s_full = v_empty + VS;     // [2]
  uint64_t* s_empty = s_full + 2;      // [2] count 8
  uint64_t* p_full = s_empty + 2;      // [2] count 8
  uint64_t* p_empty = p_full + 2;      // [2]
  uint64_t* o_full = p_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* s_table = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 128;
  const int mode = blockIdx.y;
  const int kt_begin = static_cast<int>((static_cast<long long>(p.nkt) * blockIdx.z) / p.ksplit);
  const int kt_end = static_cast<int>((static_cast<long long>(p.nkt) * (blockIdx.z + 1)) / p.ksplit);
  const int ntiles = kt_end - kt_begin;
  const int ch0 = mode * D;               // first channel of this mode in the Q/K rows
  const int qk_col = (ch0 >> 6) << 6;     // TMA column of the 64-channel atom holding it
  const uint32_t qk_inner = static_cast<uint32_t>(ch0 & 63) * 2u;
  constexpr uint32_t kTmemO = 2 * BK;     // O accumulator starts after the two S buffers

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], 8);     // one arrival per softmax warp of the group
      mbar_init(&p_full[b], 8);
      mbar_init(&p_empty[b], 1);
    }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (p.pos_table) {
    const int n = (2 * p.R + 1) * (2 * p.R + 1);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      s_table[i] = p.pos_table[i] * p.w_pos * 1.4426950408889634f;   // pre-scaled to the log2 domain
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int WM = (p.dbg >> 8) & 3;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define PV_TRACE(role, tile, slot) do { if (tr && (threadIdx.x & 31) == 0 && (tile) < 64) p.trace[((role) * 64 + (tile)) * 8 + (slot)] = clock64(); } while (0)
  if (ntiles > 0) {
    if (warp == 0) {
      // ------------------------------------ TMA producer ------------------------------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, S::kQBytes);
        for (int a = 0; a < S::kQAtoms; ++a)
          tma_load_2d(sQ + a * 128 * 128, &tmQ, q_full, qk_col + a * 64, q0);
        // K and V are two independent pipelines (K feeds S two tiles ahead of the P.V that frees a V
        // stage), so one thread serves both with non-blocking probes instead of waiting on either.
        int ks = 0, vs = 0, ki = 0, vi = 0;
        uint32_t kph = 0, vph = 0, idle = 0;
        while (ki < ntiles || vi < ntiles) {
          bool moved = false;
          if (ki < ntiles && mbar_test(&k_empty[ks], kph ^ 1u)) {
            const int kt = kt_begin + ki;
            const int by = kt / p.nbx, bx = kt - by * p.nbx;
            if (p.dbg & 64) mbar_arrive(&k_full[ks]);
            else {
              mbar_arrive_expect_tx(&k_full[ks], S::kKBytes);
              for (int a = 0; a < S::kQAtoms; ++a)
                tma_load_3d(sK + ks * S::kKBytes + a * BK * 128, &tmK, &k_full[ks], qk_col + a * 64, bx * BW, by * 8);
            }
            PV_TRACE(3, ki, 0);
            if (++ks == KS) { ks = 0; kph ^= 1u; }
            ++ki; moved = true;
          }
          if (vi < ntiles && mbar_test(&v_empty[vs], vph ^ 1u)) {
            const int kt = kt_begin + vi;
            if (p.dbg & 32) mbar_arrive(&v_full[vs]);
            else {
              mbar_arrive_expect_tx(&v_full[vs], S::kVBytes);
              for (int a = 0; a < BK / 64; ++a)
                tma_load_2d(sV + vs * S::kVBytes + a * F * 128, &tmV, &v_full[vs], kt * BK + a * 64, mode * F);
            }
            PV_TRACE(3, vi, 1);
            if (++vs == VS) { vs = 0; vph ^= 1u; }
            ++vi; moved = true;
          }
          if (moved) idle = 0;
          else {
            __nanosleep(32);
            if (++idle > (1u << 24)) __trap();     // protocol bug -> launch failure, never a hang
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------------ MMA issuer --------------------------------------
      constexpr uint32_t idesc_s = umma_idesc_f16<128, BK>();
      constexpr uint32_t idesc_o = umma_idesc_f16<128, F>();
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      mbar_wait_mode(q_full, 0, WM);
      auto issue_s = [&](int j) {
        const int b = j & 1;
        const uint32_t use = static_cast<uint32_t>(j >> 1);
        mbar_wait_mode(&k_full[ks], kph, WM);
        mbar_wait_mode(&s_empty[b], (use & 1u) ^ 1u, WM);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sq = smem_u32(sQ) + qk_inner;
          const uint32_t sk = smem_u32(sK + ks * S::kKBytes) + qk_inner;
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const int ka = (k * 16) >> 6, kin = (k * 16) & 63;
            const uint64_t dq = umma_desc_sw128(sq + ka * 128 * 128 + kin * 2);
            const uint64_t dk = umma_desc_sw128(sk + ka * BK * 128 + kin * 2);
            if (!(p.dbg & 16)) umma_f16(tmem_base + b * BK, dq, dk, idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&k_empty[ks]);
          umma_commit(&s_full[b]);
        }
        __syncwarp();
        if (++ks == KS) { ks = 0; kph ^= 1u; }
      };
      // Tensor-pipe order: S(0) S(1) | S(2) PV(0) | S(3) PV(1) | ...  S(j+2) only needs the softmax group
      // of tile j to have pulled S(j) into registers (s_empty), which happens right after S(j) lands, so
      // it runs ahead of PV(j) and both softmax groups always find their next S tile ready.
      issue_s(0);
      if (ntiles > 1) issue_s(1);
      for (int j = 0; j < ntiles; ++j) {
        PV_TRACE(0, j, 0);
        if (j + 2 < ntiles) issue_s(j + 2);
        PV_TRACE(0, j, 1);
        const int b = j & 1;
        const uint32_t use = static_cast<uint32_t>(j >> 1);
        mbar_wait_mode(&p_full[b], use & 1u, WM);
        PV_TRACE(0, j, 2);
        mbar_wait_mode(&v_full[vs], vph, WM);
        PV_TRACE(0, j, 3);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sp = smem_u32(sP + b * S::kPBytes);
          const uint32_t sv = smem_u32(sV + vs * S::kVBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const int ka = (k * 16) >> 6, kin = (k * 16) & 63;
            const uint64_t dp = umma_desc_sw128(sp + ka * 128 * 128 + kin * 2);
            const uint64_t dv = umma_desc_sw128(sv + ka * F * 128 + kin * 2);
            if (!(p.dbg & 8)) umma_f16(tmem_base + kTmemO, dp, dv, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&p_empty[b]);
          umma_commit(&v_empty[vs]);
          if (j == ntiles - 1) umma_commit(o_full);
        }
        __syncwarp();
        PV_TRACE(0, j, 4);
        if (++vs == VS) { vs = 0; vph ^= 1u; }
      }
    } else {
      // ------------------------------------ softmax groups ----------------------------------
      // warp w (2..17): tile parity sg = ((w-2)>>2)&1, column half ch = (w-2)>>3, TMEM lane quadrant w&3
      const int sg = ((warp - 2) >> 2) & 1;
      const int ch = (warp - 2) >> 3;
      const int lane_grp = warp & 3;
      const int row = lane_grp * 32 + (threadIdx.x & 31);
      const int q = q0 + row;
      const int qy = q / p.g.Wp, qx = q - qy * p.g.Wp;
      const float clipv = *p.clip;
      const bool clamped = clipv < INFINITY;
      const float lse = (q < p.g.Mp) ? p.lse2[static_cast<size_t>(mode) * p.g.Mp + q] : 0.f;
      const float sc2 = p.scale * 1.4426950408889634f;
      const float clip2 = clipv * 1.4426950408889634f;
      const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
      const int R = p.R, TD = 2 * R + 1;
      const bool has_bias = p.pos_table != nullptr;

      for (int j = sg; j < ntiles; j += 2) {
        const uint32_t use = static_cast<uint32_t>(j >> 1);
        const int kt = kt_begin + j;
        const int by = kt / p.nbx, bx = kt - by * p.nbx;
        // this thread's half block: block rows [ch*4, ch*4+4), all BW columns
        const int iy0 = by * 8 + ch * 4 - qy + R;       // table row of the first block row
        const int ix0 = bx * BW - qx + R;               // table column of the first block column
        const bool near = has_bias && (iy0 + 3 >= 0) && (iy0 <= 2 * R) && (ix0 + BW - 1 >= 0) && (ix0 <= 2 * R);
        const int trole = (warp == 2 || warp == 6) ? 1 + sg : 99;
        if (trole < 4) PV_TRACE(trole, j, 0);
        mbar_wait_mode(&s_full[sg], use & 1u, WM);
        if (trole < 4) PV_TRACE(trole, j, 1);
        mbar_wait_mode(&p_empty[sg], (use & 1u) ^ 1u, WM);
        if (trole < 4) PV_TRACE(trole, j, 2);
        tc_fence_after();
        __syncwarp();
        uint8_t* pbuf = sP + sg * S::kPBytes;
        // all of this thread's S columns are requested up front (one wait instead of one per chunk),
        // then the S buffer is handed back to the MMA warp before any math happens
        uint32_t raw_all[HALF];
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int c = 0; c < HALF; c += 32)
            tmem_ld32(tlane + sg * BK + ch * HALF + c, *reinterpret_cast<uint32_t(*)[32]>(&raw_all[c]));
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int c = 0; c < HALF; ++c) raw_all[c] = static_cast<uint32_t>(j + c);
        }
        tc_fence_before();
        mbar_arrive_warp(&s_empty[sg]);
        if (trole < 4) PV_TRACE(trole, j, 3);
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
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int col = c + e;                       // column inside this half: row-major (4 x BW)
              const int iy = iy0 + col / BW, ix = ix0 + col % BW;
              if (static_cast<unsigned>(iy) <= static_cast<unsigned>(2 * R) &&
                  static_cast<unsigned>(ix) <= static_cast<unsigned>(2 * R))
                x[e] += s_table[iy * TD + ix];
            }
          }
          const int kcol = ch * HALF + c;                  // key column inside the tile
          const int atom = kcol >> 6;
          const int chunk0 = (kcol & 63) >> 3;
          if (p.dbg & 1) {
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              uint4 u;
              u.x = pack_bf16x2(x[8 * v4 + 0], x[8 * v4 + 1]);
              u.y = pack_bf16x2(x[8 * v4 + 2], x[8 * v4 + 3]);
              u.z = pack_bf16x2(x[8 * v4 + 4], x[8 * v4 + 5]);
              u.w = pack_bf16x2(x[8 * v4 + 6], x[8 * v4 + 7]);
              *reinterpret_cast<uint4*>(pbuf + atom * 128 * 128 + swz128_offset(row, chunk0 + v4)) = u;
            }
            continue;
          }
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 u;
            u.x = pack_bf16x2(fast_ex2(x[8 * v4 + 0]), fast_ex2(x[8 * v4 + 1]));
            u.y = pack_bf16x2(fast_ex2(x[8 * v4 + 2]), fast_ex2(x[8 * v4 + 3]));
            u.z = pack_bf16x2(fast_ex2(x[8 * v4 + 4]), fast_ex2(x[8 * v4 + 5]));
            u.w = pack_bf16x2(fast_ex2(x[8 * v4 + 6]), fast_ex2(x[8 * v4 + 7]));
            if (!(p.dbg & 4) || u.x == 0x12345678u)
              *reinterpret_cast<uint4*>(pbuf + atom * 128 * 128 + swz128_offset(row, chunk0 + v4)) = u;
          }
        }
        if (trole < 4) PV_TRACE(trole, j, 4);
        if (!(p.dbg & 1024)) fence_proxy_async_smem();      // st.shared -> visible to the tensor core's async proxy
        mbar_arrive_warp(&p_full[sg]);
        if (trole < 4) PV_TRACE(trole, j, 5);
      }

      // ------------------------------------ O epilogue --------------------------------------
      // the four groups split the F columns in quarters
      mbar_wait_mode(o_full, 0, WM);
      tc_fence_after();
      __syncwarp();
      float* dst = p.out + ((static_cast<size_t>(blockIdx.z) * p.M + mode) * p.g.Mp + q) * F;
      constexpr int kQuarter = F / 4;
      const int c_begin = (sg * 2 + ch) * kQuarter;
#pragma unroll
      for (int c = 0; c < kQuarter; c += 32) {
        uint32_t raw[32];
        tmem_ld32(tlane + kTmemO + c_begin + c, raw);
        tmem_ld_wait();
        if (q < p.g.Mp && !(p.dbg & 128)) {
          float4* d4 = reinterpret_cast<float4*>(dst + c_begin + c);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            d4[e] = make_float4(__uint_as_float(raw[4 * e]), __uint_as_float(raw[4 * e + 1]),
                                __uint_as_float(raw[4 * e + 2]), __uint_as_float(raw[4 * e + 3]));
        }
      }
      tc_fence_before();
    }
  } else {
    // empty key range: this split contributes zeros
    if (warp >= 2) {
      const int row = (warp & 3) * 32 + (threadIdx.x & 31);
      const int part = (((warp - 2) >> 2) & 1) * 2 + ((warp - 2) >> 3);
      const int q = q0 + row;
      if (q < p.g.Mp) {
        float* dst = p.out + ((static_cast<size_t>(blockIdx.z) * p.M + mode) * p.g.Mp + q) * F;
        for (int c = part * (F / 4); c < (part + 1) * (F / 4); ++c) dst[c] = 0.f;
      }
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cb
