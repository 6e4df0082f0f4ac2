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
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (thread == accumulator row == TMEM lane).
#pragma once
#include "common.cuh"

namespace cb {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;       // 64 bf16 = one 128-byte swizzle atom row
constexpr int kGemmThreads = 192;
constexpr int kMaxTaps = 49;

enum GemmEpilogue : int {
  EPI_STORE = 0,    // v = act(alpha*acc + bias[n]) -> bf16 and/or f32 row-major
  EPI_GRU_ZR = 1,   // n<128: z=sigmoid -> Z(f32);  n>=128: r=sigmoid, (r*h) -> bf16
  EPI_GRU_Q = 2,    // q=tanh; h=(1-z)h+zq -> Hm(f32) and bf16 copy
  EPI_MOTION = 3,   // n<126: relu(acc+bias) ; n in {126,127}: flow channel copy
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
  // row validity: Wp > 0 => row m is a real token iff (m % Wp) < W and (m / Wp) < H;
  //               Wp == 0 => every m < M is valid.
  int Wp, W, H;
  // epilogue operands
  float alpha;
  int act;                       // 0 none, 1 relu
  const float* bias;             // [Npad] or nullptr
  __nv_bfloat16* out_b;          // bf16 destination (row-major), may be nullptr
  int ldb, colb;
  float* out_f;                  // f32 destination, may be nullptr
  int ldf, colf;
  float* aux_f0;                 // EPI_GRU_*: Z   [M,128] f32
  float* aux_f1;                 // EPI_GRU_*: Hm  [M,128] f32 ; EPI_MOTION: flow [M,2] f32
};

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN >= 256) ? 4 : 6;
  static constexpr int kABytes = kGemmBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
shift_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ GemmParams p) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* acc_bar = empty_bar + S::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * kGemmBM;
  const int n0 = blockIdx.y * BN;
  const int kchunks = p.K / kGemmBK;
  const int nk = p.T * kchunks;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nk; ++it) {
        const int t = it / kchunks;
        const int kc = it - t * kchunks;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
        tma_load_2d(sa, &tmA, &full_bar[stage], p.a_koff + kc * kGemmBK, m0 + p.tap_off[t]);
        if (p.b_blocked) {
          const int by = static_cast<int>(blockIdx.y) / p.b_nbx, bx = static_cast<int>(blockIdx.y) - by * p.b_nbx;
          tma_load_3d(sb, &tmB, &full_bar[stage], p.b_koff + kc * kGemmBK, bx * (BN / 8), by * 8);
        } else {
          tma_load_2d(sb, &tmB, &full_bar[stage], p.b_koff + kc * kGemmBK, t * p.Npad + n0);
        }
        if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = umma_idesc_f16<kGemmBM, (BN < 16 ? 16 : BN)>();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nk; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sa + S::kABytes);
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
          umma_f16(tmem_base, da + 2u * k, db + 2u * k, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs retire
        if (it == nk - 1) umma_commit(acc_bar);  // accumulator complete
      }
      __syncwarp();
      if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int lane_grp = warp & 3;                 // TMEM lanes this warp may read
    const int row = lane_grp * 32 + (threadIdx.x & 31);
    const int m = m0 + row;
    bool valid = m < p.M;
    if (p.Wp > 0) {
      const int y = m / p.Wp;
      const int x = m - y * p.Wp;
      valid = valid && (x < p.W) && (y < p.H);
    }
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t raw[32];
      tmem_ld32(trow + c, raw);
      tmem_ld_wait();
      if (!valid) continue;
      const int n = n0 + c;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a = __uint_as_float(raw[j]) * p.alpha;
        if (p.bias) a += __ldg(p.bias + n + j);
        v[j] = a;
      }
      if constexpr (EPI == EPI_STORE) {
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (p.out_b) {
          uint4* dst = reinterpret_cast<uint4*>(p.out_b + static_cast<size_t>(m) * p.ldb + p.colb + n);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
            u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
            u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
            u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
            dst[q] = u;
          }
        }
        if (p.out_f) {
          float4* dst = reinterpret_cast<float4*>(p.out_f + static_cast<size_t>(m) * p.ldf + p.colf + n);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      } else if constexpr (EPI == EPI_GRU_ZR) {
        if (n < 128) {
          float4* dst = reinterpret_cast<float4*>(p.aux_f0 + static_cast<size_t>(m) * 128 + n);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(sigmoidf_acc(v[4 * q]), sigmoidf_acc(v[4 * q + 1]),
                                 sigmoidf_acc(v[4 * q + 2]), sigmoidf_acc(v[4 * q + 3]));
        } else {
          const int nh = n - 128;
          const float4* hsrc = reinterpret_cast<const float4*>(p.aux_f1 + static_cast<size_t>(m) * 128 + nh);
          uint4* dst = reinterpret_cast<uint4*>(p.out_b + static_cast<size_t>(m) * p.ldb + p.colb + nh);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 h0 = hsrc[2 * q], h1 = hsrc[2 * q + 1];
            uint4 u;
            u.x = pack_bf16x2(sigmoidf_acc(v[8 * q + 0]) * h0.x, sigmoidf_acc(v[8 * q + 1]) * h0.y);
            u.y = pack_bf16x2(sigmoidf_acc(v[8 * q + 2]) * h0.z, sigmoidf_acc(v[8 * q + 3]) * h0.w);
            u.z = pack_bf16x2(sigmoidf_acc(v[8 * q + 4]) * h1.x, sigmoidf_acc(v[8 * q + 5]) * h1.y);
            u.w = pack_bf16x2(sigmoidf_acc(v[8 * q + 6]) * h1.z, sigmoidf_acc(v[8 * q + 7]) * h1.w);
            dst[q] = u;
          }
        }
      } else if constexpr (EPI == EPI_GRU_Q) {
        float4* hptr = reinterpret_cast<float4*>(p.aux_f1 + static_cast<size_t>(m) * 128 + n);
        const float4* zptr = reinterpret_cast<const float4*>(p.aux_f0 + static_cast<size_t>(m) * 128 + n);
        uint4* dst = reinterpret_cast<uint4*>(p.out_b + static_cast<size_t>(m) * p.ldb + p.colb + n);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float hn[8];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float4 h = hptr[2 * q + e];
            const float4 z = zptr[2 * q + e];
            const float q0 = tanhf(v[8 * q + 4 * e + 0]), q1 = tanhf(v[8 * q + 4 * e + 1]);
            const float q2 = tanhf(v[8 * q + 4 * e + 2]), q3 = tanhf(v[8 * q + 4 * e + 3]);
            float4 o;
            o.x = (1.0f - z.x) * h.x + z.x * q0;
            o.y = (1.0f - z.y) * h.y + z.y * q1;
            o.z = (1.0f - z.z) * h.z + z.z * q2;
            o.w = (1.0f - z.w) * h.w + z.w * q3;
            hptr[2 * q + e] = o;
            hn[4 * e + 0] = o.x; hn[4 * e + 1] = o.y; hn[4 * e + 2] = o.z; hn[4 * e + 3] = o.w;
          }
          uint4 u;
          u.x = pack_bf16x2(hn[0], hn[1]);
          u.y = pack_bf16x2(hn[2], hn[3]);
          u.z = pack_bf16x2(hn[4], hn[5]);
          u.w = pack_bf16x2(hn[6], hn[7]);
          dst[q] = u;
        }
      } else if constexpr (EPI == EPI_MOTION) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        if (n + 32 == 128) {   // last chunk: channels 126,127 carry the flow itself
          v[30] = p.aux_f1[static_cast<size_t>(m) * 2 + 0];
          v[31] = p.aux_f1[static_cast<size_t>(m) * 2 + 1];
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out_b + static_cast<size_t>(m) * p.ldb + p.colb + n);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
          u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
          u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
          u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
          dst[q] = u;
        }
        if (p.out_f) {
          float4* dstf = reinterpret_cast<float4*>(p.out_f + static_cast<size_t>(m) * p.ldf + p.colf + n);
#pragma unroll
          for (int q = 0; q < 8; ++q) dstf[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace cb
