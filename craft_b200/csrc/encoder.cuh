// craft_b200 -- normalisation / activation kernels for the feature and context encoders
// (core/extractor.py BasicEncoder; SURVEY.md section 8f rank 1 -- the first row OUTSIDE the named
// hot path).  Convolutions stay cuDNN; what is replaced here are the memory-bound passes between
// them, which PyTorch runs as 4-6 separate kernels per conv (instance-norm statistics, transform,
// relu, residual add, plus NCHW<->NHWC conversions around every cuDNN call):
//
//   nhwc_stats_kernel  : per (image, channel) sum and sum of squares over H*W   (InstanceNorm2d)
//   nhwc_affine_kernel : out = [relu]( [ra*res + rb] + [relu](a*v + b) )        (norm + relu + residual)
//
// with per-(image,channel) scale/shift (a, b): instance norm a = rstd, b = -mean*rstd (eps 1e-5,
// biased variance, no affine: nn.InstanceNorm2d defaults); eval-mode batch norm a = gamma/sqrt(var+eps),
// b = beta - mean*a.  Activations are channels-last fp32 or fp16 (statistics and the affine arithmetic are
// always fp32): lanes run over channels, so every access is a coalesced 16-byte vector.
#pragma once
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace cb {

// 16-byte vector of activations: 4 floats or 8 halves
template <typename T> struct ActVec;
template <> struct ActVec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct ActVec<__half> {
  static constexpr int N = 8;
  __device__ static void load(const __half* p, float (&v)[8]) {
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&x);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h[k]);
      v[2 * k] = f.x; v[2 * k + 1] = f.y;
    }
  }
  __device__ static void store(__half* p, const float (&v)[8]) {
    uint4 x;
    __half2* h = reinterpret_cast<__half2*>(&x);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    *reinterpret_cast<uint4*>(p) = x;
  }
};

// x: [N, HW, C] channels-last (f32 or f16).  part: [chunks][N][C][2] f32 partial (sum, sum of squares);
// every block owns one slot, so the reduction order -- and therefore the result -- is fixed
// (no atomics).  grid = (chunks, N); block = 256 threads = (C/VEC channel groups) x rows in flight.
// (Folding the reduction of the partials into this kernel -- "last block done" -- was tried in round 2: one block
// summing ~300 partials per channel is L2-latency bound and made the pass 4x slower; it stays a second launch.)
template <typename T>
__global__ void __launch_bounds__(256) nhwc_stats_kernel(const T* __restrict__ x, int HW, int C,
                                                         int rows_per_block, float* __restrict__ part) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int V = ActVec<T>::N;
  __shared__ float red_s[256][V + 1], red_q[256][V + 1];
  const int cq = C / V;                       // vector columns
  const int rpb = 256 / cq;                   // rows processed per pass
  const int tc = threadIdx.x % cq, tr = threadIdx.x / cq;
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  float s[V], q[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { s[k] = 0.f; q[k] = 0.f; }
  if (tr < rpb) {
    const T* base = x + (static_cast<size_t>(n) * HW) * C + tc * V;
    for (int r = r0 + tr; r < r1; r += rpb) {
      float v[V];
      ActVec<T>::load(base + static_cast<size_t>(r) * C, v);
#pragma unroll
      for (int k = 0; k < V; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { red_s[threadIdx.x][k] = s[k]; red_q[threadIdx.x][k] = q[k]; }
  __syncthreads();
  if (tr == 0) {
    for (int j = 1; j < rpb; ++j) {
#pragma unroll
      for (int k = 0; k < V; ++k) { s[k] += red_s[j * cq + tc][k]; q[k] += red_q[j * cq + tc][k]; }
    }
    float* dst = part + ((static_cast<size_t>(blockIdx.x) * gridDim.y + n) * C + V * tc) * 2;
#pragma unroll
    for (int k = 0; k < V; ++k) { dst[2 * k] = s[k]; dst[2 * k + 1] = q[k]; }
  }
}

// part [chunks][N*C][2] -> ab [N,C,2] = (rstd, -mean*rstd).  One warp per (image, channel): lane l adds
// chunks l, l+32, ... in index order and the lanes are combined by a fixed xor tree, so the result
// does not depend on scheduling.
__global__ void __launch_bounds__(256) instnorm_finalize_kernel(const float* __restrict__ part, int chunks, int NC,
                                                                float inv_hw, float eps, float* __restrict__ ab) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= NC) return;
  float s = 0.f, q = 0.f;
  for (int c = lane; c < chunks; c += 32) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(part + (static_cast<size_t>(c) * NC + i) * 2));
    s += v.x; q += v.y;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    const float mean = s * inv_hw;
    const float var = fmaxf(q * inv_hw - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    ab[2 * i] = rstd;
    ab[2 * i + 1] = -mean * rstd;
  }
}

// image_s2d: the encoders' input transform.  [N,3,H,W] f32 frames in 0..255 -> 2*(x/255)-1 (core/network.py:170-171),
// 2x2 space-to-depth, channels-last, zero border of (2 before, 1 after) cells: out [N, H/2+3, W/2+3, 16], channel
// (py*2 + px)*3 + c, channels 12..15 zero.  The 7x7 stride-2 first convolution (core/extractor.py:129) is then a
// 4x4 stride-1 convolution over 16 channels with no padding -- a shape cuDNN runs on the sm_100 tensor-op kernels
// instead of the sm_80 fallback it picks for 3 input channels (105 us -> see profiles/).
template <typename T>
__global__ void __launch_bounds__(256) image_s2d_kernel(const float* __restrict__ img, int N, int H, int W,
                                                        T* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int Hs = H / 2 + 3, Ws = W / 2 + 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(N) * Hs * Ws) return;
  const int xs = static_cast<int>(idx % Ws);
  const int ys = static_cast<int>((idx / Ws) % Hs);
  const int n = static_cast<int>(idx / (static_cast<long long>(Ws) * Hs));
  const int Y = ys - 2, X = xs - 2;
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
  if (Y >= 0 && Y < H / 2 && X >= 0 && X < W / 2) {
    const float* base = img + static_cast<size_t>(n) * 3 * H * W;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const float2 two = __ldg(reinterpret_cast<const float2*>(base + (static_cast<size_t>(c) * H + 2 * Y + py) * W + 2 * X));
        v[(py * 2 + 0) * 3 + c] = 2.0f * (two.x / 255.0f) - 1.0f;      // the reference's rounding order
        v[(py * 2 + 1) * 3 + c] = 2.0f * (two.y / 255.0f) - 1.0f;
      }
  }
  T* dst = out + idx * 16;
  if constexpr (sizeof(T) == 2) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
      w[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) reinterpret_cast<float4*>(dst)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
}

// out = relu_out( res_term + relu_in(a*v + b) ),  res_term = 0 | res | ra*res + rb
// v, res, out: [N, HW, C] (f32 or f16).  ab, rab: [N or 1][C][2] f32 (per-image stride ab_nstride elements, 0 = shared).
template <typename T>
__global__ void __launch_bounds__(256) nhwc_affine_kernel(const T* __restrict__ v, const float* __restrict__ ab,
                                                          int ab_nstride, const T* __restrict__ res,
                                                          const float* __restrict__ rab, int rab_nstride,
                                                          int relu_in, int relu_out, long long per_image /*HW*C*/,
                                                          int C, long long totalv, T* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int V = ActVec<T>::N;
  const long long iv = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= totalv) return;
  const long long e = iv * V;
  const int n = static_cast<int>(e / per_image);
  const int c = static_cast<int>(e % C);
  float r[V];
  ActVec<T>::load(v + e, r);
  if (ab) {
    const float* p = ab + static_cast<size_t>(n) * ab_nstride + 2 * c;
#pragma unroll
    for (int k = 0; k < V; ++k) r[k] = fmaf(r[k], __ldg(p + 2 * k), __ldg(p + 2 * k + 1));
  }
  if (relu_in) {
#pragma unroll
    for (int k = 0; k < V; ++k) r[k] = fmaxf(r[k], 0.f);
  }
  if (res) {
    float sres[V];
    ActVec<T>::load(res + e, sres);
    if (rab) {
      const float* p = rab + static_cast<size_t>(n) * rab_nstride + 2 * c;
#pragma unroll
      for (int k = 0; k < V; ++k) sres[k] = fmaf(sres[k], __ldg(p + 2 * k), __ldg(p + 2 * k + 1));
    }
#pragma unroll
    for (int k = 0; k < V; ++k) r[k] += sres[k];
  }
  if (relu_out) {
#pragma unroll
    for (int k = 0; k < V; ++k) r[k] = fmaxf(r[k], 0.f);
  }
  ActVec<T>::store(out + e, r);
}

// ---------------------------------------------------------------------------------------------
// instnorm_fused_kernel: InstanceNorm2d + ReLU (+ residual) of one channels-last tensor in ONE launch
//     out = relu_out( res_term + relu_in( (x - mean) * rstd ) ),   res_term = 0 | res | ra*res + rb
// replacing nhwc_stats + instnorm_finalize + nhwc_affine (three launches, x read twice from memory).
// A cooperative grid of at most one CTA per SM; CTA j of image n owns a contiguous slab of rows:
//   phase 1  the slab is pulled into shared memory with cp.async.bulk (16 KB pieces, one mbarrier each; the
//            whole slab -- 198 KB for the 64-channel layers at 448x1024 -- is in flight at once) and the
//            per-channel sum / sum of squares are accumulated from shared memory as the pieces land; one
//            (sum, sumsq) pair per channel and CTA goes to `part`;
//   grid.sync()
//   phase 2  every CTA adds the partials of its image in a fixed order (bit-reproducible, no atomics), forms
//            (rstd, -mean*rstd) and applies them to the slab still sitting in shared memory: x is read from
//            memory ONCE, the only other traffic is the residual and the output.
// Rows of a slab beyond the shared-memory capacity (larger images) are read from global memory in both phases.
// Thread layout: a warp covers 32/CQP rows x CQP 16-byte channel groups (CQP = C/V rounded up to a power of
// two), so a lane's channels are fixed and the cross-row reduction is a shuffle tree.
// ---------------------------------------------------------------------------------------------
constexpr int kInThreads = 512;
constexpr int kInMaxPieces = 16;
constexpr int kInPieceBytes = 16384;

struct InFusedParams {
  int N, HW, C;
  int cpi;             // CTAs per image; grid = cpi * N
  int rows_per_cta;
  int smem_rows;       // rows of a slab held in shared memory
  int cqp;             // vector columns rounded up to a power of two (<= 32)
  float inv_hw, eps;
  int relu_in, relu_out, rab_nstride;
  int coop;            // 1: cooperative launch, cooperative_groups grid barrier; 0: plain launch, the barrier below
};

// Grid barrier of a NON-cooperative launch (CRAFT_B200_IN_COOP=0 experiment): sense-reversing, state = {count, sense}
// in global memory, zero at first use and left as {0, sense+1}.  Only safe while every CTA of the grid can become
// resident without another spinning grid holding its SM -- hence the cooperative launch is the default.
__device__ __forceinline__ void grid_barrier_manual(unsigned* state, unsigned nctas, unsigned sense0) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(state, 1u);
    if (old == nctas - 1) {
      state[0] = 0u;
      __threadfence();
      atomicAdd(state + 1, 1u);
    } else {
      unsigned v, spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(state + 1) : "memory");
        if (++spins > (1u << 26)) __trap();
      } while (v == sense0);
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename T> struct InSmemVec;
template <> struct InSmemVec<float> {
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 x = *reinterpret_cast<const float4*>(p);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  }
};
template <> struct InSmemVec<__half> {
  __device__ static void load(const __half* p, float (&v)[8]) {
    const uint4 x = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&x);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h[k]);
      v[2 * k] = f.x; v[2 * k + 1] = f.y;
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(kInThreads, 1) instnorm_fused_kernel(
    const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ rab, float* part, unsigned* sync_state,
    float* __restrict__ ab_out, T* __restrict__ out, const InFusedParams p) {
  constexpr int V = ActVec<T>::N;
  constexpr int NW = kInThreads / 32;
  extern __shared__ __align__(128) uint8_t in_smem[];
  const int C = p.C, C2 = 2 * p.C;
  const int rowbytes = C * static_cast<int>(sizeof(T));
  T* slab = reinterpret_cast<T*>(in_smem);
  float* red = reinterpret_cast<float*>(in_smem + static_cast<size_t>(p.smem_rows) * rowbytes);   // [NW][C2]; later [nsub][C2]
  float* s_ab = red + NW * C2;                                                                     // [C2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ab + C2);                                         // [kInMaxPieces]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cq = C / V, cqp = p.cqp;
  const int rw = 32 / cqp;                         // rows per warp and pass
  const int tc = lane & (cqp - 1), trw = lane / cqp;
  const bool col_ok = tc < cq;
  const int n = blockIdx.x / p.cpi, j = blockIdx.x - n * p.cpi;
  const int r0 = j * p.rows_per_cta;
  const int my_rows = max(0, min(p.HW, r0 + p.rows_per_cta) - r0);
  const int in_smem_rows = min(my_rows, p.smem_rows);
  const int piece_rows = max(1, kInPieceBytes / rowbytes);
  const int npieces = (in_smem_rows + piece_rows - 1) / piece_rows;
  const T* xg = x + (static_cast<size_t>(n) * p.HW + r0) * C;

  unsigned sense0 = 0;
  if (tid == 0) {
    for (int k = 0; k < npieces; ++k) mbar_init(&bars[k], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  if (tid == 0) {
    if (!p.coop) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(sense0) : "l"(sync_state + 1) : "memory");
    for (int k = 0; k < npieces; ++k) {
      const int rk = min(piece_rows, in_smem_rows - k * piece_rows);
      const uint32_t bytes = static_cast<uint32_t>(rk) * rowbytes;
      mbar_arrive_expect_tx(&bars[k], bytes);
      bulk_load_1d(in_smem + static_cast<size_t>(k) * piece_rows * rowbytes, xg + static_cast<size_t>(k) * piece_rows * C, bytes,
                   &bars[k]);
    }
  }

  // ---- phase 1: per-channel sums over the slab
  float s[V], q[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { s[k] = 0.f; q[k] = 0.f; }
  const int rstep = NW * rw;
  const int rfirst = warp * rw + trw;
  if (col_ok) {
    int waited = -1;
    int r = rfirst;
    for (; r < in_smem_rows; r += rstep) {
      const int pc = r / piece_rows;
      if (pc > waited) { mbar_wait(&bars[pc], 0u); waited = pc; }
      float v[V];
      InSmemVec<T>::load(slab + static_cast<size_t>(r) * C + tc * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    }
    for (; r < my_rows; r += 4 * rstep) {            // overflow rows: straight from global, four loads in flight
      float v[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ru = min(r + u * rstep, my_rows - 1);
        ActVec<T>::load(xg + static_cast<size_t>(ru) * C + tc * V, v[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u * rstep < my_rows) {
#pragma unroll
          for (int k = 0; k < V; ++k) { s[k] += v[u][k]; q[k] = fmaf(v[u][k], v[u][k], q[k]); }
        }
      }
    }
  }
  for (int o = cqp; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
      q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
    }
  }
  if (trw == 0 && col_ok) {
    float* dst = red + warp * C2 + 2 * V * tc;
#pragma unroll
    for (int k = 0; k < V; ++k) { dst[2 * k] = s[k]; dst[2 * k + 1] = q[k]; }
  }
  __syncthreads();
  if (tid < C2) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) acc += red[w * C2 + tid];
    part[static_cast<size_t>(blockIdx.x) * C2 + tid] = acc;
  }
  if (p.coop) {
    __threadfence();
    cooperative_groups::this_grid().sync();
  } else {
    grid_barrier_manual(sync_state, gridDim.x, sense0);
  }
  pdl_launch_dependents();

  // ---- totals of this image (fixed order), scale / shift.  16-byte loads: C2/4 vector columns x nsub row groups;
  // a thread's partials (cpi / nsub of them, 5 for the 64-channel layers) are all requested before the first add
  const int vcols = C2 / 4;
  const int nsub = min(kInThreads / vcols, NW);       // red holds NW rows of C2 floats
  {
    const int vc = tid % vcols, sub = tid / vcols;
    if (sub < nsub) {
      const float4* src = reinterpret_cast<const float4*>(part + static_cast<size_t>(n) * p.cpi * C2) + vc;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = sub; c < p.cpi; c += 8 * nsub) {
        float4 t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          t[u] = (c + u * nsub < p.cpi) ? __ldcg(src + static_cast<size_t>(c + u * nsub) * vcols) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w; }
      }
      *reinterpret_cast<float4*>(red + sub * C2 + 4 * vc) = acc;
    }
  }
  __syncthreads();
  if (tid < C) {
    float sm = 0.f, sq = 0.f;
    for (int sub = 0; sub < nsub; ++sub) { sm += red[sub * C2 + 2 * tid]; sq += red[sub * C2 + 2 * tid + 1]; }
    const float mean = sm * p.inv_hw;
    const float var = fmaxf(sq * p.inv_hw - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    s_ab[2 * tid] = rstd;
    s_ab[2 * tid + 1] = -mean * rstd;
    if (j == 0 && ab_out) {
      ab_out[(static_cast<size_t>(n) * C + tid) * 2] = rstd;
      ab_out[(static_cast<size_t>(n) * C + tid) * 2 + 1] = -mean * rstd;
    }
  }
  __syncthreads();

  // ---- phase 2: apply
  if (!col_ok) return;
  float a[V], b[V], ra[V], rb[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    a[k] = s_ab[2 * (tc * V + k)];
    b[k] = s_ab[2 * (tc * V + k) + 1];
    ra[k] = 1.f; rb[k] = 0.f;
  }
  if (res && rab) {
    const float* pr = rab + static_cast<size_t>(n) * p.rab_nstride + 2 * tc * V;
#pragma unroll
    for (int k = 0; k < V; ++k) { ra[k] = __ldg(pr + 2 * k); rb[k] = __ldg(pr + 2 * k + 1); }
  }
  const T* rg = res ? res + (static_cast<size_t>(n) * p.HW + r0) * C + tc * V : nullptr;
  T* og = out + (static_cast<size_t>(n) * p.HW + r0) * C + tc * V;
  for (int r = rfirst; r < my_rows; r += 4 * rstep) {
    float v[4][V], z[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ru = min(r + u * rstep, my_rows - 1);
      if (rg) ActVec<T>::load(rg + static_cast<size_t>(ru) * C, z[u]);
      if (ru < in_smem_rows) InSmemVec<T>::load(slab + static_cast<size_t>(ru) * C + tc * V, v[u]);
      else ActVec<T>::load(xg + static_cast<size_t>(ru) * C + tc * V, v[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ru = r + u * rstep;
      if (ru < my_rows) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float y = fmaf(v[u][k], a[k], b[k]);
          if (p.relu_in) y = fmaxf(y, 0.f);
          if (rg) y += fmaf(z[u][k], ra[k], rb[k]);
          if (p.relu_out) y = fmaxf(y, 0.f);
          v[u][k] = y;
        }
        ActVec<T>::store(og + static_cast<size_t>(ru) * C, v[u]);
      }
    }
  }
}

}  // namespace cb
