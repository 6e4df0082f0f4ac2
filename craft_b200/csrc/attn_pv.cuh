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
// key split).  Pipeline per key tile j:
//     MMA warp : S[j&1] = Q K_j^T                     (tcgen05, accumulator in TMEM)
//     softmax  : P[j&1] = exp2(...) as bf16 -> smem    (two 128-thread groups alternate tiles)
//     MMA warp : O += P[j&1] V_j                       (A = P from smem, B = V^T tile from TMA)
// S(j+1) is issued before PV(j) so the tensor pipe never waits for the softmax group.
// V is consumed as V^T ([M*F, keys], keys contiguous) so that every operand is K-major.
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

namespace cb {

constexpr int kPvThreads = 320;

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
  int nkeys;               // number of key rows to visit (Mp)
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
  uint64_t* s_empty = s_full + 2;      // [2] count 128
  uint64_t* p_full = s_empty + 2;      // [2] count 128
  uint64_t* p_empty = p_full + 2;      // [2]
  uint64_t* o_full = p_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* s_table = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 128;
  const int mode = blockIdx.y;
  const int nkt = (p.nkeys + BK - 1) / BK;
  const int kt_begin = static_cast<int>((static_cast<long long>(nkt) * blockIdx.z) / p.ksplit);
  const int kt_end = static_cast<int>((static_cast<long long>(nkt) * (blockIdx.z + 1)) / p.ksplit);
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
      mbar_init(&s_empty[b], 128);
      mbar_init(&p_full[b], 128);
      mbar_init(&p_empty[b], 1);
    }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (p.pos_table) {
    const int n = (2 * p.R + 1) * (2 * p.R + 1);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      s_table[i] = p.pos_table[i] * p.w_pos * 1.4426950408889634f;   // pre-scaled to log2 domain
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (ntiles > 0) {
    if (warp == 0) {
      // ------------------------------------ TMA producer ------------------------------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, S::kQBytes);
        for (int a = 0; a < S::kQAtoms; ++a)
          tma_load_2d(sQ + a * 128 * 128, &tmQ, q_full, qk_col + a * 64, q0);
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int i = 0; i < ntiles; ++i) {
          const int k0 = (kt_begin + i) * BK;
          mbar_wait(&k_empty[ks], kph ^ 1u);
          mbar_arrive_expect_tx(&k_full[ks], S::kKBytes);
          for (int a = 0; a < S::kQAtoms; ++a)
            tma_load_2d(sK + ks * S::kKBytes + a * BK * 128, &tmK, &k_full[ks], qk_col + a * 64, k0);
          if (++ks == KS) { ks = 0; kph ^= 1u; }
          mbar_wait(&v_empty[vs], vph ^ 1u);
          mbar_arrive_expect_tx(&v_full[vs], S::kVBytes);
          for (int a = 0; a < BK / 64; ++a)
            tma_load_2d(sV + vs * S::kVBytes + a * F * 128, &tmV, &v_full[vs], k0 + a * 64, mode * F);
          if (++vs == VS) { vs = 0; vph ^= 1u; }
        }
      }
    } else if (warp == 1) {
      // ------------------------------------ MMA issuer --------------------------------------
      constexpr uint32_t idesc_s = umma_idesc_f16<128, BK>();
      constexpr uint32_t idesc_o = umma_idesc_f16<128, F>();
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int b = j & 1;
        const uint32_t use = static_cast<uint32_t>(j >> 1);
        mbar_wait(&k_full[ks], kph);
        mbar_wait(&s_empty[b], (use & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sq = smem_u32(sQ) + qk_inner;
          const uint32_t sk = smem_u32(sK + ks * S::kKBytes) + qk_inner;
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const int ka = (k * 16) >> 6, kin = (k * 16) & 63;
            const uint64_t dq = umma_desc_sw128(sq + ka * 128 * 128 + kin * 2);
            const uint64_t dk = umma_desc_sw128(sk + ka * BK * 128 + kin * 2);
            umma_f16(tmem_base + b * BK, dq, dk, idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&k_empty[ks]);
          umma_commit(&s_full[b]);
        }
        __syncwarp();
        if (++ks == KS) { ks = 0; kph ^= 1u; }
      };
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);
        const int b = j & 1;
        const uint32_t use = static_cast<uint32_t>(j >> 1);
        mbar_wait(&p_full[b], use & 1u);
        mbar_wait(&v_full[vs], vph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sp = smem_u32(sP + b * S::kPBytes);
          const uint32_t sv = smem_u32(sV + vs * S::kVBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const int ka = (k * 16) >> 6, kin = (k * 16) & 63;
            const uint64_t dp = umma_desc_sw128(sp + ka * 128 * 128 + kin * 2);
            const uint64_t dv = umma_desc_sw128(sv + ka * F * 128 + kin * 2);
            umma_f16(tmem_base + kTmemO, dp, dv, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&p_empty[b]);
          umma_commit(&v_empty[vs]);
          if (j == ntiles - 1) umma_commit(o_full);
        }
        __syncwarp();
        if (++vs == VS) { vs = 0; vph ^= 1u; }
      }
    } else {
      // ------------------------------------ softmax groups ----------------------------------
      const int sg = (warp - 2) >> 2;
      const int lane_grp = warp & 3;
      const int row = lane_grp * 32 + (threadIdx.x & 31);
      const int q = q0 + row;
      const int qy = q / p.g.Wp, qx = q - qy * p.g.Wp;
      const float clipv = *p.clip;
      const float lse = (q < p.g.Mp) ? p.lse2[static_cast<size_t>(mode) * p.g.Mp + q] : 0.f;
      const float sc2 = p.scale * 1.4426950408889634f;
      const float clip2 = clipv * 1.4426950408889634f;
      const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
      const int R = p.R, TD = 2 * R + 1;
      // query rows covered by this CTA (for the tile-level "near the diagonal" test)
      const int qy_lo = q0 / p.g.Wp, qy_hi = (q0 + 127) / p.g.Wp;

      for (int j = sg; j < ntiles; j += 2) {
        const uint32_t use = static_cast<uint32_t>(j >> 1);
        const int k0 = (kt_begin + j) * BK;
        const int ky_lo = k0 / p.g.Wp, ky_hi = (k0 + BK - 1) / p.g.Wp;
        const bool near = p.pos_table && (ky_hi >= qy_lo - R) && (ky_lo <= qy_hi + R);
        mbar_wait(&s_full[sg], use & 1u);
        mbar_wait(&p_empty[sg], (use & 1u) ^ 1u);
        tc_fence_after();
        __syncwarp();
        uint8_t* pbuf = sP + sg * S::kPBytes;
#pragma unroll 1
        for (int c = 0; c < BK; c += 32) {
          uint32_t raw[32];
          tmem_ld32(tlane + sg * BK + c, raw);
          tmem_ld_wait();
          if (c + 32 >= BK) {      // S buffer fully read -> MMA warp may overwrite it
            tc_fence_before();
            mbar_arrive(&s_empty[sg]);
          }
          float x[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float s = __uint_as_float(raw[e]) * sc2;
            x[e] = fminf(fmaxf(s, -clip2), clip2) - lse;
          }
          if (near) {
            int k = k0 + c;
            int ky = k / p.g.Wp, kx = k - ky * p.g.Wp;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int dy = ky - qy, dx = kx - qx;
              if (dy >= -R && dy <= R && dx >= -R && dx <= R) x[e] += s_table[(dy + R) * TD + dx + R];
              if (++kx == p.g.Wp) { kx = 0; ++ky; }
            }
          }
          const int atom = c >> 6;
          const int chunk0 = (c & 63) >> 3;
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 u;
            u.x = pack_bf16x2(fast_ex2(x[8 * v4 + 0]), fast_ex2(x[8 * v4 + 1]));
            u.y = pack_bf16x2(fast_ex2(x[8 * v4 + 2]), fast_ex2(x[8 * v4 + 3]));
            u.z = pack_bf16x2(fast_ex2(x[8 * v4 + 4]), fast_ex2(x[8 * v4 + 5]));
            u.w = pack_bf16x2(fast_ex2(x[8 * v4 + 6]), fast_ex2(x[8 * v4 + 7]));
            *reinterpret_cast<uint4*>(pbuf + atom * 128 * 128 + swz128_offset(row, chunk0 + v4)) = u;
          }
        }
        fence_proxy_async_smem();      // st.shared -> visible to the tensor core's async proxy
        mbar_arrive(&p_full[sg]);
      }

      // ------------------------------------ O epilogue --------------------------------------
      // both groups split the F columns in halves
      mbar_wait(o_full, 0);
      tc_fence_after();
      __syncwarp();
      float* dst = p.out + ((static_cast<size_t>(blockIdx.z) * p.M + mode) * p.g.Mp + q) * F;
      constexpr int kHalf = F / 2;
#pragma unroll 1
      for (int c = sg * kHalf; c < (sg + 1) * kHalf; c += 32) {
        uint32_t raw[32];
        tmem_ld32(tlane + kTmemO + c, raw);
        tmem_ld_wait();
        if (q < p.g.Mp) {
          float4* d4 = reinterpret_cast<float4*>(dst + c);
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
      const int sg = (warp - 2) >> 2;
      const int q = q0 + row;
      if (q < p.g.Mp) {
        float* dst = p.out + ((static_cast<size_t>(blockIdx.z) * p.M + mode) * p.g.Mp + q) * F;
        for (int c = sg * (F / 2); c < (sg + 1) * (F / 2); ++c) dst[c] = 0.f;
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
