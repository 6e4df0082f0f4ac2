// craft_b200 -- kernels behind the STANDALONE forward()s of the reference's small modules.  Inside
// CRAFT.forward these operations are fused into the epilogues of the tensor-core kernels
// (scores.cuh, modes_finalize); the reference's nn.Module signatures still have to work on their own
// (SURVEY.md section 8b), so the same arithmetic exists here as plain coalesced kernels.
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

namespace cb {

// -------------------------------------------------------------------------------------------
// LearnedSoftAggregate.forward  core/setrans.py:289-300 on a dense tensor with the group
// (mode) axis leading:  x [M][N] (num_feat == 1)  or  x [M][R][F] (num_feat == F).
//   num_feat == 1 :  p_m = softmax_m(w * b_m[i] + b)            out[i]   = sum_m p_m x_m[i]
//   num_feat == F :  p_m = softmax_m(<w, b_m[r,:]> + b)          out[r,:] = sum_m p_m x_m[r,:]
// (b = score_basis, which defaults to x).  M <= 8.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) soft_aggregate_scalar_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ basis, int M,
                                                                    long long N, const float* __restrict__ w,
                                                                    const float* __restrict__ b,
                                                                    float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const float wv = w[0], bv = b[0];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < N;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float xs[8], sc[8];
    float mx = -INFINITY;
    for (int m = 0; m < M; ++m) {
      xs[m] = x[m * N + i];
      sc[m] = fmaf(basis[m * N + i], wv, bv);
      mx = fmaxf(mx, sc[m]);
    }
    float num = 0.f, den = 0.f;
    for (int m = 0; m < M; ++m) {
      const float e = __expf(sc[m] - mx);
      num = fmaf(e, xs[m], num);
      den += e;
    }
    out[i] = num / den;
  }
}

// one warp per row r
__global__ void __launch_bounds__(256) soft_aggregate_feat_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ basis, int M,
                                                                  long long R, int F, const float* __restrict__ w,
                                                                  const float* __restrict__ b,
                                                                  float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long r = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (r >= R) return;
  float sc[8];
  float mx = -INFINITY;
  for (int m = 0; m < M; ++m) {
    const float* row = basis + (m * R + r) * F;
    float s = 0.f;
    for (int f = lane; f < F; f += 32) s = fmaf(row[f], w[f], s);
    sc[m] = warp_sum(s) + b[0];
    mx = fmaxf(mx, sc[m]);
  }
  float den = 0.f;
  for (int m = 0; m < M; ++m) {
    sc[m] = __expf(sc[m] - mx);
    den += sc[m];
  }
  const float inv = 1.0f / den;
  for (int f = lane; f < F; f += 32) {
    float a = 0.f;
    for (int m = 0; m < M; ++m) a = fmaf(sc[m], x[(m * R + r) * F + f], a);
    out[r * F + f] = a * inv;
  }
}

// -------------------------------------------------------------------------------------------
// Dense attention matrix from projected token rows (debugging / standalone CrossAttFeatTrans.forward
// on SMALL grids only: the production path never forms it).
//   S_m[q,k] = clamp(<Q_m[q], K_m[k]> * scale) + w_pos * bias(k - q) + mask(k - q)
//   lse2 == nullptr : out = S                          core/setrans.py:514-542
//   lse2 != nullptr : out = exp2(S*log2e - lse2[m][q])  core/setrans.py:553 (softmax over keys)
// out: [M][U][U] over REAL tokens (no halo).  Block = 8 warps; a warp owns one (m, q) and sweeps
// the keys, lane = key.
// -------------------------------------------------------------------------------------------
struct DenseAttnParams {
  const act_t* Q;
  const act_t* K;
  int C, M, d;
  float scale, w_pos;
  const float* pos_table;
  int R;
  const float* clip;
  const float* lse2;      // [M][Mp] or nullptr
  int mask_radius;        // > 0: keys with max(|dy|,|dx|) > mask_radius get -1e9 (core/setrans.py:580-584)
  float* out;
};

__global__ void __launch_bounds__(256) attn_dense_kernel(DenseAttnParams p, Grid2 g) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int U = g.H * g.W;
  const long long unit = blockIdx.x * 8ll + (threadIdx.x >> 5);     // (m, q) over real tokens
  if (unit >= static_cast<long long>(p.M) * U) return;
  const int m = static_cast<int>(unit / U), qi = static_cast<int>(unit - static_cast<long long>(m) * U);
  const int qy = qi / g.W, qx = qi - qy * g.W;
  const int qrow = qy * g.Wp + qx;
  const act_t* qp = p.Q + static_cast<size_t>(qrow) * p.C + m * p.d;
  const float clipv = *p.clip;
  const int TD = 2 * p.R + 1;
  const float lse = p.lse2 ? p.lse2[static_cast<size_t>(m) * g.Mp + qrow] : 0.f;
  float* dst = p.out + (static_cast<size_t>(m) * U + qi) * U;
  for (int k0 = 0; k0 < U; k0 += 32) {
    const int ki = k0 + lane;
    if (ki >= U) break;
    const int ky = ki / g.W, kx = ki - ky * g.W;
    const act_t* kp = p.K + static_cast<size_t>(ky * g.Wp + kx) * p.C + m * p.d;
    float acc = 0.f;
    for (int c = 0; c < p.d; c += 2) {
      const float2 a = unpack_act2(*reinterpret_cast<const uint32_t*>(qp + c));
      const float2 b = unpack_act2(*reinterpret_cast<const uint32_t*>(kp + c));
      acc = fmaf(a.x, b.x, acc);
      acc = fmaf(a.y, b.y, acc);
    }
    float s = fminf(fmaxf(acc * p.scale, -clipv), clipv);
    const int dy = ky - qy, dx = kx - qx;
    if (p.pos_table && dy >= -p.R && dy <= p.R && dx >= -p.R && dx <= p.R)
      s += p.w_pos * p.pos_table[(dy + p.R) * TD + dx + p.R];
    if (p.mask_radius > 0 && (abs(dy) > p.mask_radius || abs(dx) > p.mask_radius)) s += -1e9f;
    dst[ki] = p.lse2 ? exp2f(s * 1.4426950408889634f - lse) : s;
  }
}

}  // namespace cb
