// craft_b200 -- persistent implicit-GEMM 3x3 convolution, 64 -> 64 channels, fp16 activations, for the first
// residual layer of the feature / context encoders (core/extractor.py:24-26,142-146: layer1 = two ResidualBlocks
// of 3x3 stride-1 convolutions at half resolution; SURVEY.md section 8f rank 1, the first row outside the named hot
// path).  cuDNN runs these 64-channel layers at 0.39 of the tensor peak (26 us for two 224x512 images); they are
// 4 of the 15 convolutions of each encoder and the largest ones.
//
// Layout ("padded-flat", the token-grid layout of DESIGN.md section 3 at image scale): an activation is a row-major
// [R, 64] fp16 matrix, row r = (n*(H+1) + y)*(W+2) + x; the two cells x in {W, W+1} of every grid row and the whole
// row y = H of every image hold ZEROS.  A tap (dy, dx) is then the row offset dy*(W+2) + dx, rows beyond either end
// are zero-filled by TMA, and the zero cells provide the convolution's padding between rows and between images: no
// im2col, no border code.
//
// Schedule: one persistent CTA per SM; the CTAs of image n walk its 128-row tiles with stride ctas_per_image.
//   warp 0   TMA producer.  The nine 64x64 weight tiles (72 KB) are loaded ONCE and stay in shared memory.  Per tile
//            (= one pipeline stage, two stages) and kernel row dy one 136-row box of the input (rows m0 + dy*(W+2) - 1 ...): the three taps of that
//            kernel row read it through descriptors shifted by whole rows (the 128-byte swizzle is a function of the
//            absolute shared-memory address, so a row-shifted start needs no fix-up -- gemm.cuh), i.e. 51 KB of
//            operand ingest per 128x64 output tile against 36 MMAs (1152 clk): tensor-bound, not ingest-bound.
//   warp 1   MMA issuer: 3 x 3 x 4 tcgen05.mma (128x64x16, fp16 -> fp32) per tile into one of TWO TMEM accumulators,
//            so that the epilogue of tile i runs under the main loop of tile i+1.
//   warps 2-9  epilogue: thread = output row x 32 channels: tcgen05.ld, (+ bias, ReLU: the folded eval BatchNorm of
//            the context encoder), zeros for halo cells, fp16, two 32-byte stores; with STATS the per-channel sum /
//            sum of squares of the fp32 accumulators (InstanceNorm2d statistics of the feature encoder) are kept
//            in registers across the CTA's tiles and reduced once, in a fixed order, into `part` -- the separate
//            statistics pass over the convolution output disappears.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "encoder.cuh"

namespace cb {

constexpr int kCvThreads = 64 + 256;
constexpr int kCvStages = 2;                    // a stage = the three input boxes (kernel rows) of one output tile
constexpr int kCvBoxBytes = 136 * 128;          // one input box: 136 rows x 64 channels x 2 B
constexpr int kCvABytes = 3 * kCvBoxBytes;
constexpr int kCvWTile = 64 * 128;              // one tap's weights: 64 output rows x 64 input channels x 2 B
constexpr int kCvWBytes = 9 * kCvWTile;
constexpr int kCvSmem = kCvWBytes + kCvStages * kCvABytes + 1024 /*align*/ + 256 /*barriers*/ + 8 * 64 * 4 /*stats*/;
static_assert(kCvSmem <= 227 * 1024, "conv3x3_c64: shared memory budget");

struct ConvEncParams {
  int N, H, W, Wp;          // images, image size, row pitch W + 2
  int rpi;                  // rows per image = (H + 1) * Wp
  int tiles_per_image;      // ceil(rpi / 128)
  int ctas_per_image;       // grid = N * ctas_per_image
  const float* bias;        // [64] or nullptr
  int relu;
  __half* out;              // [N * rpi][64]
  float* part;              // STATS: [ctas_per_image][N][64][2] (sum, sum of squares) over the valid cells
};

template <bool STATS>
__global__ void __launch_bounds__(kCvThreads, 1)
conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ ConvEncParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + kCvWBytes;
  uint8_t* tail = sA + kCvStages * kCvABytes;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full_bar = w_full + 1;
  uint64_t* empty_bar = full_bar + kCvStages;
  uint64_t* acc_full = empty_bar + kCvStages;     // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2], count 8 (epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_red = reinterpret_cast<float*>(tail + 256);     // [8 warps][64]

  const int warp = threadIdx.x >> 5;
  const int n = blockIdx.x / p.ctas_per_image;
  const int j = blockIdx.x - n * p.ctas_per_image;
  const int row_base = n * p.rpi;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    for (int s = 0; s < kCvStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  pdl_wait();                                   // the input is the previous kernel's output
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, kCvWBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * kCvWTile, &tmW, w_full, 0, t * 64);
      int stage = 0;
      uint32_t phase = 0;
      for (int lt = j; lt < p.tiles_per_image; lt += p.ctas_per_image) {
        const int m0 = row_base + lt * 128;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], kCvABytes);
        for (int gi = 0; gi < 3; ++gi)
          tma_load_2d(sA + stage * kCvABytes + gi * kCvBoxBytes, &tmX, &full_bar[stage], 0, m0 + (gi - 1) * p.Wp - 1);
        if (++stage == kCvStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = umma_idesc_f16<128, 64, true>();      // fp16 operands in both builds of the library
    const bool leader = elect_one();
    const uint64_t dW = umma_desc_sw128(smem_u32(sW));
    const uint64_t dA = umma_desc_sw128(smem_u32(sA));
    constexpr uint64_t kAStep = static_cast<uint64_t>(kCvABytes >> 4);
    constexpr uint64_t kBoxStep = static_cast<uint64_t>(kCvBoxBytes >> 4);
    constexpr uint64_t kWStep = static_cast<uint64_t>(kCvWTile >> 4);
    mbar_wait(w_full, 0u);
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int lt = j; lt < p.tiles_per_image; lt += p.ctas_per_image, ++it) {
      const int b = it & 1;
      while (!mbar_try_wait_nohint(&acc_empty[b], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u)) {}
      tc_fence_after();
      const uint32_t tacc = tmem_base + static_cast<uint32_t>(b * 64);
      // one barrier round per TILE (36 MMAs): with a round per kernel row the issuing warp spent two thirds of its
      // time in the three wait / fence / commit sequences (tensor pipe 31 %, profiles/r02_conv64.txt)
      while (!mbar_try_wait_nohint(&full_bar[stage], phase)) {}
      tc_fence_after();
      if (leader) {
        const uint64_t da = dA + static_cast<uint64_t>(stage) * kAStep;
#pragma unroll
        for (int gi = 0; gi < 3; ++gi) {
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const uint64_t d = da + static_cast<uint64_t>(gi) * kBoxStep + static_cast<uint64_t>(t) * 8u;   // tap dx = t - 1: one row (128 B) further
            const uint64_t e = dW + static_cast<uint64_t>(gi * 3 + t) * kWStep;
            if (gi == 0 && t == 0) umma_f16(tacc, d, e, idesc, 0u);
            else umma_f16_acc(tacc, d, e, idesc);
            umma_f16_acc(tacc, d + 2u, e + 2u, idesc);
            umma_f16_acc(tacc, d + 4u, e + 4u, idesc);
            umma_f16_acc(tacc, d + 6u, e + 6u, idesc);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&acc_full[b]);
      }
      if (++stage == kCvStages) { stage = 0; phase ^= 1u; }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue ----------------------------------
    const int lane_grp = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;              // channels half*32 .. half*32+31
    const int row = lane_grp * 32 + (threadIdx.x & 31);
    float bias[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bias[c] = p.bias ? p.bias[half * 32 + c] : 0.f;
    float st_s[STATS ? 32 : 1], st_q[STATS ? 32 : 1];
    if constexpr (STATS) {
#pragma unroll
      for (int c = 0; c < 32; ++c) { st_s[c] = 0.f; st_q[c] = 0.f; }
    }
    int it = 0;
    for (int lt = j; lt < p.tiles_per_image; lt += p.ctas_per_image, ++it) {
      const int b = it & 1;
      const int lr = lt * 128 + row;               // row inside the image
      const int y = lr / p.Wp, x = lr - y * p.Wp;
      const bool in_image = lr < p.rpi;            // rows past the image's end belong to the next image's first tile
      const bool valid = in_image && (y < p.H) && (x < p.W);
      mbar_wait(&acc_full[b], (static_cast<uint32_t>(it) >> 1) & 1u);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(b * 64 + half * 32), raw);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(&acc_empty[b]);             // accumulator drained: the MMA warp may start tile it + 2 in it
      uint32_t w16[16];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        float v0 = __uint_as_float(raw[c]), v1 = __uint_as_float(raw[c + 1]);
        if constexpr (STATS) {
          if (valid) {
            st_s[c] += v0; st_q[c] = fmaf(v0, v0, st_q[c]);
            st_s[c + 1] += v1; st_q[c + 1] = fmaf(v1, v1, st_q[c + 1]);
          }
        }
        v0 += bias[c]; v1 += bias[c + 1];
        if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        if (!valid) { v0 = 0.f; v1 = 0.f; }        // halo cells / the gap row stay zero: they ARE the padding
        const __half2 h2 = __floats2half2_rn(v0, v1);
        w16[c >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      if (in_image) {
        __half* dst = p.out + (static_cast<size_t>(row_base + lr)) * 64 + half * 32;
        st_global_v8(dst, *reinterpret_cast<const uint32_t(*)[8]>(&w16[0]));
        st_global_v8(dst + 16, *reinterpret_cast<const uint32_t(*)[8]>(&w16[8]));
      }
    }
    if constexpr (STATS) {
      // one reduction per CTA, fixed order: lanes (xor tree) -> the four lane-quadrant warps of a channel half
      const int ew = warp - 2;                     // = half * 4 + lane_grp
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float s = warp_sum(st_s[c]), q = warp_sum(st_q[c]);
        if ((threadIdx.x & 31) == 0) { s_red[ew * 64 + c] = s; s_red[ew * 64 + 32 + c] = q; }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = threadIdx.x - 64;             // 0..255
      if (et < 128) {
        const int h = et >> 6, v = et & 63;        // v < 32: sum of channel h*32+v; v >= 32: sum of squares of channel h*32+v-32
        const float tot = (s_red[(h * 4 + 0) * 64 + v] + s_red[(h * 4 + 1) * 64 + v]) +
                          (s_red[(h * 4 + 2) * 64 + v] + s_red[(h * 4 + 3) * 64 + v]);
        const int ch = h * 32 + (v & 31);
        p.part[((static_cast<size_t>(j) * p.N + n) * 64 + ch) * 2 + (v >> 5)] = tot;
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ---------------------------------------------------------------------------------------------
// nhwc_affine_pad_kernel: nhwc_affine (encoder.cuh) between the dense channels-last layout [N][H][W][C] cuDNN uses
// and the padded-flat one above, in any combination:
//     out = relu_out( [ra*res+rb | res] + relu_in(a*v + b) ),   halo cells / gap rows of a padded output := 0
// One thread per 16-byte vector of the OUTPUT.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) nhwc_affine_pad_kernel(const T* __restrict__ v, int v_pad, const float* __restrict__ ab,
                                                              int ab_nstride, const T* __restrict__ res, int res_pad,
                                                              const float* __restrict__ rab, int rab_nstride, int relu_in,
                                                              int relu_out, int N, int H, int W, int C, T* __restrict__ out,
                                                              int out_pad) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int V = 16 / static_cast<int>(sizeof(T));
  const int cq = C / V;
  const int Wp = W + 2;
  const long long rows_out = out_pad ? static_cast<long long>(N) * (H + 1) * Wp : static_cast<long long>(N) * H * W;
  const long long iv = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= rows_out * cq) return;
  const long long r = iv / cq;
  const int c = static_cast<int>(iv - r * cq) * V;
  int n, y, x;
  if (out_pad) {
    const long long rpi = static_cast<long long>(H + 1) * Wp;
    n = static_cast<int>(r / rpi);
    const int lr = static_cast<int>(r - n * rpi);
    y = lr / Wp; x = lr - y * Wp;
  } else {
    n = static_cast<int>(r / (static_cast<long long>(H) * W));
    const int lr = static_cast<int>(r - static_cast<long long>(n) * H * W);
    y = lr / W; x = lr - y * W;
  }
  T* dst = out + r * C + c;
  if (y >= H || x >= W) {                          // only reachable for a padded output
    *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  auto src_row = [&](int pad) -> long long {
    return pad ? (static_cast<long long>(n) * (H + 1) + y) * Wp + x : (static_cast<long long>(n) * H + y) * W + x;
  };
  float a[V];
  ActVec<T>::load(v + src_row(v_pad) * C + c, a);
  if (ab) {
    const float* pp = ab + static_cast<size_t>(n) * ab_nstride + 2 * c;
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] = fmaf(a[k], __ldg(pp + 2 * k), __ldg(pp + 2 * k + 1));
  }
  if (relu_in) {
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] = fmaxf(a[k], 0.f);
  }
  if (res) {
    float s[V];
    ActVec<T>::load(res + src_row(res_pad) * C + c, s);
    if (rab) {
      const float* pp = rab + static_cast<size_t>(n) * rab_nstride + 2 * c;
#pragma unroll
      for (int k = 0; k < V; ++k) s[k] = fmaf(s[k], __ldg(pp + 2 * k), __ldg(pp + 2 * k + 1));
    }
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] += s[k];
  }
  if (relu_out) {
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] = fmaxf(a[k], 0.f);
  }
  ActVec<T>::store(dst, a);
}

}  // namespace cb
