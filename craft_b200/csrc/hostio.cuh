// craft_b200 -- device side of the host I/O around the hot path: the warm-start
// forward_interpolate (core/utils/utils.py:34-62, scipy griddata on the CPU in the reference) and the byte
// re-packing for the .flo / KITTI-png writers (core/utils/frame_utils.py:70-99, 116-120).  HBM-bound
// integer/byte work: coalesced loads, shared-memory staging, no tensor cores.
#pragma once
#include "common.cuh"

namespace cb {

// forward_interpolate: every pixel p = (x0, y0) of a flow field carries its flow to p + flow(p); the output at
// grid point g is the flow of the NEAREST landed point (Euclidean; points landing outside the open rectangle
// (0,W) x (0,H) are dropped -- utils.py:50).  Brute force over all sources, staged through shared memory in
// tiles of 256: N^2 distance evaluations with N = h*w <= ~8k at 1/8 resolution (51 M for 448x1024), exact.
// Ties go to the lowest source index.  out == zeros when no source is valid (griddata raises there).
__global__ void __launch_bounds__(256) forward_interpolate_kernel(const float* __restrict__ flow, int H, int W,
                                                                  float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sx[256], sy[256], sdx[256], sdy[256];
  const int N = H * W;
  const int t = blockIdx.x * 256 + threadIdx.x;
  const float gx = static_cast<float>(t % W), gy = static_cast<float>(t / W);
  float best = INFINITY, bdx = 0.f, bdy = 0.f;
  for (int s0 = 0; s0 < N; s0 += 256) {
    const int s = s0 + threadIdx.x;
    float x1 = INFINITY, y1 = INFINITY, dx = 0.f, dy = 0.f;
    if (s < N) {
      dx = flow[s];
      dy = flow[N + s];
      const float px = static_cast<float>(s % W) + dx, py = static_cast<float>(s / W) + dy;
      if (px > 0.f && px < static_cast<float>(W) && py > 0.f && py < static_cast<float>(H)) { x1 = px; y1 = py; }
    }
    __syncthreads();
    sx[threadIdx.x] = x1; sy[threadIdx.x] = y1; sdx[threadIdx.x] = dx; sdy[threadIdx.x] = dy;
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < 256; ++j) {
      const float ex = sx[j] - gx, ey = sy[j] - gy;
      const float d = fmaf(ex, ex, ey * ey);          // +inf for dropped sources
      if (d < best) { best = d; bdx = sdx[j]; bdy = sdy[j]; }
    }
  }
  if (t < N) {
    out[t] = bdx;
    out[N + t] = bdy;
  }
}

// flow_encode: [2,H,W] f32 planes ->
//   mode 0 (.flo payload, frame_utils.py:94-98): interleaved f32 [H][W][2] = (u, v)
//   mode 1 (KITTI png, frame_utils.py:116-120):  u16 [H][W][3] in the B,G,R order cv2.imwrite takes =
//                                                (valid = 1, 64 v + 2^15, 64 u + 2^15), C-style truncation
__global__ void __launch_bounds__(256) flow_encode_kernel(const float* __restrict__ flow, int H, int W, int mode,
                                                          void* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int N = H * W;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= N) return;
  const float u = flow[t], v = flow[N + t];
  if (mode == 0) {
    reinterpret_cast<float2*>(out)[t] = make_float2(u, v);
  } else {
    uint16_t* o = reinterpret_cast<uint16_t*>(out) + 3 * static_cast<size_t>(t);
    // frame_utils.py:117 evaluates 64.0 * uv + 2**15 on the float32 array, i.e. in float32 (64 u is exact, the sum
    // rounds to nearest); the later astype(uint16) truncates toward zero and wraps out-of-range values mod 2^16
    const long long qu = static_cast<long long>(__fadd_rn(64.0f * u, 32768.0f));
    const long long qv = static_cast<long long>(__fadd_rn(64.0f * v, 32768.0f));
    o[0] = 1;
    o[1] = static_cast<uint16_t>(qv);
    o[2] = static_cast<uint16_t>(qu);
  }
}

}  // namespace cb
