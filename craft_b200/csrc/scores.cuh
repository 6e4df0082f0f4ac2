// craft_b200 -- multi-mode score tiles S_m = Q_m K_m^T on tcgen05, with two fused epilogues:
//
//  SC_CORR  (TransCorrBlock.update, core/corr.py:148-207 + CrossAttFeatTrans scores-only path
//            core/setrans.py:501-550 + LearnedSoftAggregate(1) core/setrans.py:289-300):
//            clamp -> soft-aggregate the M modes -> + w_pos*bias -> accumulate the global
//            layer-norm statistics (sum, sum^2) and the global raw max (clamp gate,
//            core/setrans.py:520-529) -> 2x2/4x4/8x8 average pooling of each thread's private
//            8x8 key block -> store pyramid levels 1..3 (level 0 optional).  The level-0 volume
//            (U x U) therefore never has to reach HBM.
//  SC_LSE   (CrossAttFeatTrans softmax, core/setrans.py:540-553): per (mode, query) running
//            max / sum-exp over the key range -> partial log-sum-exp, merged by lse_merge_kernel.
//            The P.V kernel (attn_pv.cuh) then recomputes P = exp(S - lse) tile by tile.
//
// Tile: 128 queries (TMEM lanes) x 64 keys (an 8x8 spatial block fetched by a 3-D TMA box) x M
// modes (M*64 = 256 TMEM columns, double buffered).  grid = (query tiles, key splits).
// Warps: 0 = TMA, 1 = MMA (+TMEM alloc), 2..5 = epilogue group 0, 6..9 = epilogue group 1
// (groups alternate key tiles / TMEM buffers).
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

namespace cb {

constexpr int kScThreads = 320;
constexpr int kScKStages = 3;

enum ScoreMode : int { SC_CORR = 0, SC_LSE = 1 };

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
  int ksplit;              // gridDim.y
  // SC_CORR
  float w_agg;             // LearnedSoftAggregate(1).feat2score.weight
  double* stat_sum;        // [2]: sum, sumsq   (atomics)
  float* stat_max;         // [1]: global max of raw scaled scores (pre-bias)
  float* lvl[4];           // pooled volumes [Mp][h_l*w_l]; lvl[0] optional
  int hl[4], wl[4];
  // SC_LSE
  float2* lse_part;        // [ksplit][M][Mp] (max, sumexp) natural-exp domain
};

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// smem: Q tile (C/64 atoms x 16 KB) + K stages (C/64 atoms x 8 KB each) + barriers + table
template <int MODE>
__global__ void __launch_bounds__(kScThreads, 1)
scores_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ ScoreParams p) {
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
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + kScKStages;
  uint64_t* acc_full = k_empty + kScKStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_table = reinterpret_cast<float*>(tmem_slot + 4);   // (2R+1)^2 floats
  float* s_red = s_table + 232;                                // 8 warps x 4

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 128;
  const int nkt = p.nkt_y * p.nkt_x;
  const int kt_begin = static_cast<int>((static_cast<long long>(nkt) * blockIdx.y) / p.ksplit);
  const int kt_end = static_cast<int>((static_cast<long long>(nkt) * (blockIdx.y + 1)) / p.ksplit);
  const int ntiles = kt_end - kt_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    mbar_init(q_full, 1);
    for (int s = 0; s < kScKStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);       // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();                                    // everything below reads tensors of earlier kernels
  if (p.pos_table) {
    const int n = (2 * p.R + 1) * (2 * p.R + 1);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_table[i] = p.pos_table[i] * p.w_pos;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, q_bytes);
      for (int a = 0; a < atoms; ++a) tma_load_2d(sQ + a * 128 * 128, &tmQ, q_full, a * 64, q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < ntiles; ++i) {
        const int kt = kt_begin + i;
        const int by = kt / p.nkt_x, bx = kt - by * p.nkt_x;
        mbar_wait(&k_empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&k_full[stage], k_bytes);
        uint8_t* dst = sK + stage * k_bytes;
        for (int a = 0; a < atoms; ++a)
          tma_load_3d(dst + a * 64 * 128, &tmK, &k_full[stage], a * 64, bx * 8, by * 8);
        if (++stage == kScKStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer --------------------------------------
    constexpr uint32_t idesc = umma_idesc_f16<128, 64>();
    mbar_wait(q_full, 0);
    int stage = 0;
    uint32_t phase = 0;
    const int ksteps = p.d / 16;
    for (int i = 0; i < ntiles; ++i) {
      const int b = i & 1;
      const uint32_t use = static_cast<uint32_t>(i >> 1);
      mbar_wait(&k_full[stage], phase);
      mbar_wait(&acc_empty[b], (use & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sq = smem_u32(sQ);
        const uint32_t sk = smem_u32(sK + stage * k_bytes);
        for (int m = 0; m < p.M; ++m) {
          const int ch0 = m * p.d;                 // first channel of this mode
          const int atom = ch0 >> 6;
          const uint32_t inner = static_cast<uint32_t>(ch0 & 63) * 2u;   // byte offset inside the 128-B row
          const uint64_t dq = umma_desc_sw128(sq + atom * 128 * 128 + inner);
          const uint64_t dk = umma_desc_sw128(sk + atom * 64 * 128 + inner);
          const uint32_t tcol = tmem_base + static_cast<uint32_t>(b * 256 + m * 64);
          for (int k = 0; k < ksteps; ++k) {
            // K steps past the first 64 channels of a mode (d = 128) move to the next atom.
            const int ka = (k * 16) >> 6, kin = (k * 16) & 63;
            const uint64_t oq = static_cast<uint64_t>((ka * 128 * 128 + kin * 2) >> 4);
            const uint64_t ok = static_cast<uint64_t>((ka * 64 * 128 + kin * 2) >> 4);
            umma_f16(tcol, dq + oq, dk + ok, idesc, k != 0 ? 1u : 0u);
          }
        }
        umma_commit(&k_empty[stage]);
        umma_commit(&acc_full[b]);
      }
      __syncwarp();
      if (++stage == kScKStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ------------------------------------ epilogue ----------------------------------------
    const int eg = (warp - 2) >> 2;                 // epilogue group 0/1 <-> TMEM buffer
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + (threadIdx.x & 31);
    const int q = q0 + row;
    const int qy = q / p.g.Wp, qx = q - qy * p.g.Wp;
    const bool qvalid = (q < p.g.Mp) && (qx < p.g.W);
    const float clipv = *p.clip;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + eg * 256;
    const int R = p.R;
    const int TD = 2 * R + 1;

    float st_sum = 0.f, st_sq = 0.f, st_max = -INFINITY;
    float run_m[4], run_l[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) { run_m[m] = -INFINITY; run_l[m] = 0.f; }
    const float wl2 = p.w_agg * 1.4426950408889634f;

    for (int i = eg; i < ntiles; i += 2) {
      const uint32_t use = static_cast<uint32_t>(i >> 1);
      const int kt = kt_begin + i;
      const int by = kt / p.nkt_x, bx = kt - by * p.nkt_x;
      mbar_wait(&acc_full[eg], use & 1u);
      tc_fence_after();
      __syncwarp();

      const bool near = p.pos_table && (by * 8 + 7 >= qy - R) && (by * 8 <= qy + R) &&
                        (bx * 8 + 7 >= qx - R) && (bx * 8 <= qx + R);

      if constexpr (MODE == SC_CORR) {
        float agg[64];
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          uint32_t r0[16], r1[16], r2[16], r3[16];
          tmem_ld16(trow + c, r0);
          if (p.M > 1) tmem_ld16(trow + 64 + c, r1);
          if (p.M > 2) {
            tmem_ld16(trow + 128 + c, r2);
            tmem_ld16(trow + 192 + c, r3);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float s0 = __uint_as_float(r0[j]) * p.scale;
            float v;
            if (p.M == 1) {
              st_max = fmaxf(st_max, s0);
              v = fminf(fmaxf(s0, -clipv), clipv);
            } else {
              float s1 = __uint_as_float(r1[j]) * p.scale;
              float s2 = (p.M > 2) ? __uint_as_float(r2[j]) * p.scale : -INFINITY;
              float s3 = (p.M > 2) ? __uint_as_float(r3[j]) * p.scale : -INFINITY;
              st_max = fmaxf(st_max, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
              s0 = fminf(fmaxf(s0, -clipv), clipv);
              s1 = fminf(fmaxf(s1, -clipv), clipv);
              if (p.M > 2) {
                s2 = fminf(fmaxf(s2, -clipv), clipv);
                s3 = fminf(fmaxf(s3, -clipv), clipv);
              }
              // softmax over modes of w*s (the Linear(1,1) bias cancels), in the exp2 domain.
              const float t0 = s0 * wl2, t1 = s1 * wl2, t2 = s2 * wl2, t3 = s3 * wl2;
              const float tm = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
              const float e0 = fast_ex2(t0 - tm), e1 = fast_ex2(t1 - tm);
              float num = e0 * s0 + e1 * s1, den = e0 + e1;
              if (p.M > 2) {
                const float e2 = fast_ex2(t2 - tm), e3 = fast_ex2(t3 - tm);
                num += e2 * s2 + e3 * s3;
                den += e2 + e3;
              }
              v = __fdividef(num, den);
            }
            agg[c + j] = v;
          }
        }
        // TMEM buffer drained -> hand it back to the MMA warp before the slow part.
        tc_fence_before();
        mbar_arrive_warp(&acc_empty[eg]);

        if (qvalid) {
          const int ky0 = by * 8, kx0 = bx * 8;
          if (near) {
#pragma unroll
            for (int e = 0; e < 64; ++e) {
              const int dy = ky0 + (e >> 3) - qy, dx = kx0 + (e & 7) - qx;
              if (dy >= -R && dy <= R && dx >= -R && dx <= R) agg[e] += s_table[(dy + R) * TD + dx + R];
            }
          }
          const bool full = (ky0 + 8 <= p.g.H) && (kx0 + 8 <= p.g.W);
#pragma unroll
          for (int e = 0; e < 64; ++e) {
            const bool kv = full || ((ky0 + (e >> 3) < p.g.H) && (kx0 + (e & 7) < p.g.W));
            if (kv) {
              st_sum += agg[e];
              st_sq += agg[e] * agg[e];
            }
          }
          // level 0 (optional, debugging / SAVECORR)
          if (p.lvl[0]) {
            float* dst = p.lvl[0] + static_cast<size_t>(q) * (p.hl[0] * p.wl[0]);
#pragma unroll
            for (int e = 0; e < 64; ++e) {
              const int yy = ky0 + (e >> 3), xx = kx0 + (e & 7);
              if (yy < p.hl[0] && xx < p.wl[0]) dst[yy * p.wl[0] + xx] = agg[e];
            }
          }
          // level 1: 4x4 cells, level 2: 2x2, level 3: 1 (floor-mode avg_pool2d chain, corr.py:186-189)
          float l1[16];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              l1[r * 4 + c] = 0.25f * (agg[(2 * r) * 8 + 2 * c] + agg[(2 * r) * 8 + 2 * c + 1] +
                                       agg[(2 * r + 1) * 8 + 2 * c] + agg[(2 * r + 1) * 8 + 2 * c + 1]);
          {
            float* dst = p.lvl[1] + static_cast<size_t>(q) * (p.hl[1] * p.wl[1]);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int yy = by * 4 + (e >> 2), xx = bx * 4 + (e & 3);
              if (yy < p.hl[1] && xx < p.wl[1]) dst[yy * p.wl[1] + xx] = l1[e];
            }
          }
          float l2[4];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c)
              l2[r * 2 + c] = 0.25f * (l1[(2 * r) * 4 + 2 * c] + l1[(2 * r) * 4 + 2 * c + 1] +
                                       l1[(2 * r + 1) * 4 + 2 * c] + l1[(2 * r + 1) * 4 + 2 * c + 1]);
          {
            float* dst = p.lvl[2] + static_cast<size_t>(q) * (p.hl[2] * p.wl[2]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int yy = by * 2 + (e >> 1), xx = bx * 2 + (e & 1);
              if (yy < p.hl[2] && xx < p.wl[2]) dst[yy * p.wl[2] + xx] = l2[e];
            }
          }
          if (by < p.hl[3] && bx < p.wl[3])
            p.lvl[3][static_cast<size_t>(q) * (p.hl[3] * p.wl[3]) + by * p.wl[3] + bx] =
                0.25f * (l2[0] + l2[1] + l2[2] + l2[3]);
        }
      } else {
        // ------------------------------ SC_LSE ------------------------------
        const int ky0 = by * 8, kx0 = bx * 8;
        const bool full = (ky0 + 8 <= p.g.H) && (kx0 + 8 <= p.g.W);
        for (int m = 0; m < p.M; ++m) {
          float sv[64];
#pragma unroll
          for (int c = 0; c < 64; c += 16) {
            uint32_t r0[16];
            tmem_ld16(trow + m * 64 + c, r0);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float s = __uint_as_float(r0[j]) * p.scale;
              st_max = fmaxf(st_max, s);
              sv[c + j] = fminf(fmaxf(s, -clipv), clipv);
            }
          }
          if (m == p.M - 1) {
            tc_fence_before();
            mbar_arrive_warp(&acc_empty[eg]);
          }
          if (near) {
#pragma unroll
            for (int e = 0; e < 64; ++e) {
              const int dy = ky0 + (e >> 3) - qy, dx = kx0 + (e & 7) - qx;
              if (dy >= -R && dy <= R && dx >= -R && dx <= R) sv[e] += s_table[(dy + R) * TD + dx + R];
            }
          }
          if (!full) {
#pragma unroll
            for (int e = 0; e < 64; ++e)
              if (!((ky0 + (e >> 3) < p.g.H) && (kx0 + (e & 7) < p.g.W))) sv[e] = -INFINITY;
          }
          float tmax = sv[0];
#pragma unroll
          for (int e = 1; e < 64; ++e) tmax = fmaxf(tmax, sv[e]);
          const float nm = fmaxf(run_m[m], tmax);
          float acc = 0.f;
#pragma unroll
          for (int e = 0; e < 64; ++e) acc += fast_ex2((sv[e] - nm) * 1.4426950408889634f);
          run_l[m] = run_l[m] * fast_ex2((run_m[m] - nm) * 1.4426950408889634f) + acc;
          run_m[m] = nm;
        }
      }
    }

    // ------------------------- per-CTA reductions / partial writes -------------------------
    if constexpr (MODE == SC_LSE) {
      // the two epilogue groups saw alternate key tiles: merge through shared memory.
      float2* xch = reinterpret_cast<float2*>(sK);   // K stages are idle now (all MMAs retired)
      // make sure every MMA that reads sK has retired: the last acc_full wait above implies it
      // for this group's tiles; the other group's tiles are covered by the named barrier below.
      asm volatile("bar.sync 1, 256;");
      if (eg == 1) {
        for (int m = 0; m < p.M; ++m) xch[m * 128 + row] = make_float2(run_m[m], run_l[m]);
      }
      asm volatile("bar.sync 1, 256;");
      if (eg == 0 && q < p.g.Mp) {
        for (int m = 0; m < p.M; ++m) {
          const float2 o = xch[m * 128 + row];
          const float nm = fmaxf(run_m[m], o.x);
          float l = 0.f;
          if (nm > -INFINITY) l = run_l[m] * __expf(run_m[m] - nm) + o.y * __expf(o.x - nm);
          p.lse_part[(static_cast<size_t>(blockIdx.y) * p.M + m) * p.g.Mp + q] = make_float2(nm, l);
        }
      }
    }
    // global max of raw scores (clamp gate) -- both modes
    {
      const float wm = warp_max(qvalid ? st_max : -INFINITY);
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
  (void)s_red;
}

// lse_merge: combine ksplit partial (max, sumexp) pairs -> lse in the log2 domain.
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
__global__ void corr_stats_finalize_kernel(const double* __restrict__ sums, double n,
                                           float* __restrict__ mean_rstd) {
  pdl_launch_dependents();
  pdl_wait();

  const double mean = sums[0] / n;
  double var = sums[1] / n - mean * mean;
  if (var < 0) var = 0;
  mean_rstd[0] = static_cast<float>(mean);
  mean_rstd[1] = static_cast<float>(1.0 / sqrt(var + 1e-12));
}
__global__ void clip_gate_kernel(const float* __restrict__ stat_max, float attn_clip,
                                 float* __restrict__ clip, int* __restrict__ flag) {
  pdl_launch_dependents();
  pdl_wait();

  const bool hit = stat_max[0] > attn_clip;
  clip[0] = hit ? attn_clip : INFINITY;
  flag[0] = hit ? 1 : 0;
}

}  // namespace cb
