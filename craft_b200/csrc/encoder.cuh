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
// b = beta - mean*a.  Activations are channels-last fp32: lanes run over channels, so every access
// is a coalesced float4.
#pragma once
#include "common.cuh"

namespace cb {

// x: [N, HW, C] f32 (channels-last).  sums: [N, C, 2] f32, zeroed by the caller.
// grid = (chunks, N); block = 256 threads = (C/4 channel-quads) x (256/(C/4) rows in flight).
__global__ void __launch_bounds__(256) nhwc_stats_kernel(const float* __restrict__ x, int HW, int C,
                                                         int rows_per_block, float* __restrict__ sums) {
  __shared__ float4 red_s[256], red_q[256];
  const int cq = C >> 2;                      // float4 columns
  const int rpb = 256 / cq;                   // rows processed per pass
  const int tc = threadIdx.x % cq, tr = threadIdx.x / cq;
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (tr < rpb) {
    const float4* base = reinterpret_cast<const float4*>(x + (static_cast<size_t>(n) * HW) * C) + tc;
    for (int r = r0 + tr; r < r1; r += rpb) {
      const float4 v = __ldg(base + static_cast<size_t>(r) * cq);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
    }
  }
  red_s[threadIdx.x] = s;
  red_q[threadIdx.x] = q;
  __syncthreads();
  if (tr == 0) {
    for (int k = 1; k < rpb; ++k) {
      const float4 a = red_s[k * cq + tc], b = red_q[k * cq + tc];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float* dst = sums + (static_cast<size_t>(n) * C + 4 * tc) * 2;
    atomicAdd(dst + 0, s.x); atomicAdd(dst + 1, q.x);
    atomicAdd(dst + 2, s.y); atomicAdd(dst + 3, q.y);
    atomicAdd(dst + 4, s.z); atomicAdd(dst + 5, q.z);
    atomicAdd(dst + 6, s.w); atomicAdd(dst + 7, q.w);
  }
}

// sums [N,C,2] -> ab [N,C,2] = (rstd, -mean*rstd)
__global__ void instnorm_finalize_kernel(const float* __restrict__ sums, int NC, float inv_hw, float eps,
                                         float* __restrict__ ab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const float mean = sums[2 * i] * inv_hw;
  const float var = fmaxf(sums[2 * i + 1] * inv_hw - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  ab[2 * i] = rstd;
  ab[2 * i + 1] = -mean * rstd;
}

// out = relu_out( res_term + relu_in(a*v + b) ),  res_term = 0 | res | ra*res + rb
// v, res, out: [N, HW, C] f32.  ab, rab: [N or 1][C][2] (per-image stride ab_nstride elements, 0 = shared).
__global__ void __launch_bounds__(256) nhwc_affine_kernel(const float* __restrict__ v, const float* __restrict__ ab,
                                                          int ab_nstride, const float* __restrict__ res,
                                                          const float* __restrict__ rab, int rab_nstride,
                                                          int relu_in, int relu_out, long long per_image /*HW*C*/,
                                                          int C, long long total4, float* __restrict__ out) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const long long e = i4 * 4;
  const int n = static_cast<int>(e / per_image);
  const int c = static_cast<int>(e % C);
  const float4 x = __ldg(reinterpret_cast<const float4*>(v) + i4);
  float r[4] = {x.x, x.y, x.z, x.w};
  if (ab) {
    const float* p = ab + static_cast<size_t>(n) * ab_nstride + 2 * c;
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fmaf(r[k], __ldg(p + 2 * k), __ldg(p + 2 * k + 1));
  }
  if (relu_in) {
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fmaxf(r[k], 0.f);
  }
  if (res) {
    const float4 y = __ldg(reinterpret_cast<const float4*>(res) + i4);
    float s[4] = {y.x, y.y, y.z, y.w};
    if (rab) {
      const float* p = rab + static_cast<size_t>(n) * rab_nstride + 2 * c;
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] = fmaf(s[k], __ldg(p + 2 * k), __ldg(p + 2 * k + 1));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] += s[k];
  }
  if (relu_out) {
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fmaxf(r[k], 0.f);
  }
  reinterpret_cast<float4*>(out)[i4] = make_float4(r[0], r[1], r[2], r[3]);
}

}  // namespace cb
