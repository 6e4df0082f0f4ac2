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

}  // namespace cb
