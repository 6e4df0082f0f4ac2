// craft_b200 -- HBM-bound kernels of the hot path (coalesced, vectorised, warp-shuffle reductions).
//
// Geometry ("padded-flat token grid"): a feature map of h x w tokens is stored
// token-major as rows p = y * Wp + x with Wp = w + 2; the two cells x in {w, w+1} of every grid
// row are halo cells that always hold zeros, so a conv tap (dy,dx), |dx| <= 2, is the row offset
// dy*Wp + dx (gemm.cuh).  Mp = h * Wp rows.
#pragma once
#include <limits.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace cb {

struct Grid2 {
  int H, W, Wp, Mp;
};

// -------------------------------------------------------------------------------------------
// pack_tokens: NCHW fp32 -> token-major rows, with optional per-token LayerNorm / activation.
//   reference: SETransInputFeatEncoder.forward core/setrans.py:791-795 (transpose + LayerNorm,
//   eps 1e-12, no affine; pos_code_weight = 0 for pos_code_type 'bias'), and the tanh / relu
//   split of cnet features core/network.py:209-211.
// mode: 0 copy, 1 layer-norm over C, 2 tanh, 3 relu, 4 relu followed by layer-norm
// Block = 256 threads handles 32 consecutive tokens of one grid row, all C channels.
// -------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) pack_tokens_kernel(const float* __restrict__ src, Grid2 g,
                                                          int mode, act_t* __restrict__ out_b,
                                                          int ldb, int colb, float* __restrict__ out_f,
                                                          int ldf, int colf) {
  pdl_launch_dependents();
  pdl_wait();

  __shared__ float tile[C][33];
  const int tiles_per_row = (g.Wp + 31) / 32;
  const int y = blockIdx.x / tiles_per_row;
  const int x0 = (blockIdx.x - y * tiles_per_row) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps
  // coalesced read along x for each channel
  for (int c = ty; c < C; c += 8) {
    const int x = x0 + tx;
    float v = (x < g.W) ? __ldg(src + (static_cast<size_t>(c) * g.H + y) * g.W + x) : 0.0f;
    if (mode == 4) v = fmaxf(v, 0.0f);
    tile[c][tx] = v;
  }
  if (mode == 4) mode = 1;   // relu then LayerNorm (intra-frame attention input, network.py:211,214)
  __syncthreads();
  // each warp owns 4 tokens; lanes stride channels
  for (int t = ty * 4; t < ty * 4 + 4; ++t) {
    const int x = x0 + t;
    if (x >= g.Wp) break;
    const size_t p = static_cast<size_t>(y) * g.Wp + x;
    const bool halo = x >= g.W;
    float mean = 0.f, rstd = 1.f;
    if (mode == 1 && !halo) {
      float s = 0.f;
      for (int c = tx; c < C; c += 32) s += tile[c][t];
      mean = warp_sum(s) * (1.0f / C);
      float v = 0.f;
      for (int c = tx; c < C; c += 32) {
        const float d = tile[c][t] - mean;
        v += d * d;
      }
      rstd = rsqrtf(warp_sum(v) * (1.0f / C) + 1e-12f);
    }
    for (int c = tx; c < C; c += 32) {
      float v = tile[c][t];
      if (halo) v = 0.f;
      else if (mode == 1) v = (v - mean) * rstd;
      else if (mode == 2) v = tanhf(v);
      else if (mode == 3) v = fmaxf(v, 0.f);
      if (out_b) out_b[p * ldb + colb + c] = f2act(v);
      if (out_f) out_f[p * ldf + colf + c] = v;
    }
  }
}

// pack_tokens_nhwc: the same token packer for a channels-last source [H][W][ldc] (f16 or f32, the layout the
// fused encoders produce): a token's channels are already contiguous, so this is one warp per token with
// 16-byte loads -- no transpose through shared memory and no NHWC -> NCHW -> token round trip
// (core/extractor.py output -> core/setrans.py:791-795 / core/network.py:209-211).
// Channels [c0, c0 + C) of every token are packed; C in {128, 256}; modes as in pack_tokens_kernel.
template <typename TS, int C>
__global__ void __launch_bounds__(256) pack_tokens_nhwc_kernel(const TS* __restrict__ src, int ldc, int c0, Grid2 g,
                                                               int mode, act_t* __restrict__ out_b, int ldb, int colb,
                                                               float* __restrict__ out_f, int ldf, int colf) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int PER = C / 32;                 // channels per lane: 4 or 8, contiguous
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= g.Mp) return;
  const int y = p / g.Wp, x = p - y * g.Wp;
  float v[PER];
  if (x < g.W) {
    const TS* s = src + (static_cast<size_t>(y) * g.W + x) * ldc + c0 + lane * PER;
    if constexpr (sizeof(TS) == 2) {          // PER halves = 8 or 16 bytes, aligned (c0, ldc multiples of 8)
      uint32_t w[PER / 2];
      if constexpr (PER == 8) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(s));
        w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
      } else {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(s));
        w[0] = u.x; w[1] = u.y;
      }
#pragma unroll
      for (int k = 0; k < PER / 2; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        v[2 * k] = f.x; v[2 * k + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < PER; k += 4) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(s + k));
        v[k] = f.x; v[k + 1] = f.y; v[k + 2] = f.z; v[k + 3] = f.w;
      }
    }
    if (mode == 4) {
#pragma unroll
      for (int k = 0; k < PER; ++k) v[k] = fmaxf(v[k], 0.f);
    }
    if (mode == 1 || mode == 4) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < PER; ++k) sum += v[k];
      const float mean = warp_sum(sum) * (1.0f / C);
      float var = 0.f;
#pragma unroll
      for (int k = 0; k < PER; ++k) { const float d = v[k] - mean; var += d * d; }
      const float rstd = rsqrtf(warp_sum(var) * (1.0f / C) + 1e-12f);
#pragma unroll
      for (int k = 0; k < PER; ++k) v[k] = (v[k] - mean) * rstd;
    } else if (mode == 2) {
#pragma unroll
      for (int k = 0; k < PER; ++k) v[k] = tanhf(v[k]);
    } else if (mode == 3) {
#pragma unroll
      for (int k = 0; k < PER; ++k) v[k] = fmaxf(v[k], 0.f);
    }
  } else {
#pragma unroll
    for (int k = 0; k < PER; ++k) v[k] = 0.f;                           // halo cell
  }
  if (out_b) {
    act_t* d = out_b + static_cast<size_t>(p) * ldb + colb + lane * PER;
#pragma unroll
    for (int k = 0; k < PER; k += 2) *reinterpret_cast<uint32_t*>(d + k) = pack_act2(v[k], v[k + 1]);
  }
  if (out_f) {
    float* d = out_f + static_cast<size_t>(p) * ldf + colf + lane * PER;
#pragma unroll
    for (int k = 0; k < PER; k += 4) *reinterpret_cast<float4*>(d + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
  }
}

// unpack_tokens: token-major rows (bf16 or f32) -> NCHW fp32.
template <typename T>
__global__ void __launch_bounds__(256) unpack_tokens_kernel(const T* __restrict__ src, int ld, int col,
                                                            int C, Grid2 g, float* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();

  __shared__ float tile[32][33];
  const int tiles_per_row = (g.W + 31) / 32;
  const int y = blockIdx.x / tiles_per_row;
  const int x0 = (blockIdx.x - y * tiles_per_row) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int t = ty; t < 32; t += 8) {
    const int x = x0 + t;
    const int c = c0 + tx;
    float v = 0.f;
    if (x < g.W && c < C) v = static_cast<float>(src[(static_cast<size_t>(y) * g.Wp + x) * ld + col + c]);
    tile[t][tx] = v;
  }
  __syncthreads();
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc;
    const int x = x0 + tx;
    if (c < C && x < g.W) dst[(static_cast<size_t>(c) * g.H + y) * g.W + x] = tile[tx][cc];
  }
}

// -------------------------------------------------------------------------------------------
// corr_lookup: radius-r bilinear window lookup over the 4-level pyramid with the global
// layer-norm applied as a deferred affine.
//   reference: CorrBlock.__call__ core/corr.py:47-71, bilinear_sampler core/utils/utils.py:65-79
//   (grid_sample bilinear, zeros padding, align_corners=True), global LN core/corr.py:200-204.
// Output channel = l*(2r+1)^2 + i*(2r+1) + j with x-offset = i - r, y-offset = j - r.
// lookup(LN(v)) = rstd * (lookup(v) - mean * sum_of_inbounds_weights)  (LN is affine, pooling
// and bilinear sampling are linear).  stats = {mean, rstd}; pass {0,1} for no normalisation.
// One warp per query token.  R = 4 -> 10x10 cells per level.
// -------------------------------------------------------------------------------------------
struct LookupParams {
  const float* lvl[4];   // level l volume: [Mq(query rows, padded-flat)][h_l * w_l]  (lvl[0] may be null if on-demand)
  int hl[4], wl[4];
  long long qstride[4];  // elements between consecutive query rows
  const float* coords;   // [Mp, 2] (x, y) padded-flat
  const float* stats;    // device {mean, rstd}
  act_t* out_b;  // [Mp, ldb] token-major (cols >= 4*81 left untouched)
  int ldb;
  float* out_nchw;       // [324, H, W] or nullptr
  int first_level;       // levels < first_level are skipped (handled by the on-demand kernel)
  // level 0 stored in 16 bits, blocked [Mp][nby*nbx][64] (scores.cuh lvl0h); used instead of lvl[0] when non-null
  const __half* lvl0h;
  const float* lvl0_base;   // [Mp][nblk][2] fp32 mean of each 4x8 half block; lvl0h holds fp16 deltas against it
  int nbx0;              // 8x8 blocks per block row
  long long qstride0h;
};

__global__ void __launch_bounds__(256) corr_lookup_kernel(LookupParams p, Grid2 g) {
  pdl_launch_dependents();
  pdl_wait();

  constexpr int R = 4, D = 2 * R + 1, WN = D + 1;   // 9 taps, 10 cells
  constexpr int NC = WN * WN, NCP = NC + 4;
  __shared__ float win[8][4][NCP];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + wib;   // padded-flat query index
  if (q >= g.Mp) return;
  const int qy = q / g.Wp, qx = q - qy * g.Wp;
  if (qx >= g.W) return;                // halo token: leave zeros
  const float cx = p.coords[2 * q], cy = p.coords[2 * q + 1];
  const float mean = p.stats[0], rstd = p.stats[1];
  // phase 1: the 10x10 windows of ALL levels are gathered first -- every lane has its 13-16 independent loads in
  // flight at once (one level at a time, each gather paid the full L2 / HBM latency: 12.6 us per call)
  float ax[4], ay[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float inv = 1.0f / static_cast<float>(1 << l);
    const float px = cx * inv, py = cy * inv;
    const float fx0 = floorf(px), fy0 = floorf(py);
    ax[l] = px - fx0;
    ay[l] = py - fy0;
    if (l < p.first_level) continue;
    const int x0 = static_cast<int>(fx0) - R, y0 = static_cast<int>(fy0) - R;
    const int hl = p.hl[l], wl = p.wl[l];
    float* wv = win[wib][l];
    if (l == 0 && p.lvl0h != nullptr) {
      const __half* vol = p.lvl0h + static_cast<long long>(q) * p.qstride0h;
      const float* bases = p.lvl0_base + static_cast<long long>(q) * (p.qstride0h >> 5);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = lane + 32 * k;
        if (e < NC) {
          const int r = e / WN, c = e - r * WN;
          const int yy = y0 + r, xx = x0 + c;
          float v = 0.f;
          if (yy >= 0 && yy < hl && xx >= 0 && xx < wl) {
            const int blk = (yy >> 3) * p.nbx0 + (xx >> 3);
            v = (__ldg(bases + 2 * blk + ((yy >> 2) & 1)) - mean) + __half2float(__ldg(vol + (blk << 6) + ((yy & 7) << 3) + (xx & 7)));
          }
          wv[e] = v;
        }
      }
    } else {
      const float* vol = p.lvl[l] + static_cast<long long>(q) * p.qstride[l];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = lane + 32 * k;
        if (e < NC) {
          const int r = e / WN, c = e - r * WN;
          const int yy = y0 + r, xx = x0 + c;
          float v = 0.f;
          if (yy >= 0 && yy < hl && xx >= 0 && xx < wl) v = __ldg(vol + yy * wl + xx) - mean;
          wv[e] = v;   // (value - mean) inside bounds, 0 outside == deferred-LN numerator
        }
      }
    }
  }
  __syncwarp();
  // phase 2: 4 x 81 bilinear taps
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    if (l < p.first_level) continue;
    const float* wv = win[wib][l];
    const float axl = ax[l], ayl = ay[l];
    for (int e = lane; e < D * D; e += 32) {
      const int i = e / D, j = e - i * D;   // i: x offset index, j: y offset index
      const float v00 = wv[j * WN + i], v01 = wv[j * WN + i + 1];
      const float v10 = wv[(j + 1) * WN + i], v11 = wv[(j + 1) * WN + i + 1];
      const float top = v00 + axl * (v01 - v00);
      const float bot = v10 + axl * (v11 - v10);
      const float val = (top + ayl * (bot - top)) * rstd;
      const int ch = l * D * D + e;
      if (p.out_b) p.out_b[static_cast<size_t>(q) * p.ldb + ch] = f2act(val);
      if (p.out_nchw) p.out_nchw[(static_cast<size_t>(ch) * g.H + qy) * g.W + qx] = val;
    }
  }
}

// -------------------------------------------------------------------------------------------
// corr_lookup0: level-0 window of the lookup computed ON DEMAND from the projected query / key
// rows, so the U x U level-0 correlation volume never exists in HBM.  For each query the 10x10
// cells around coords1 are recomputed exactly as the build kernel would have produced them:
//   s_m = <q_m, k_m>/sqrt(d) -> clamp -> softmax-over-modes aggregation -> + w_pos*bias(dy,dx)
// (core/setrans.py:514-550, 289-300), then the same deferred-LN bilinear taps as corr_lookup.
// 100 cells x 256 channels = 25.6 kMAC per query (0.37 GFLOP per iteration at 448x1024) against a
// key matrix (3.7 MB) that stays in L2.  One warp per query: every window cell is ONE coalesced 512-byte
// row load (lane = 8 channels) followed by a shuffle reduction per mode.
// -------------------------------------------------------------------------------------------
struct Lookup0Params {
  const act_t* Q;   // [Mp, 256] projected queries
  const act_t* K;   // [Mp, 256] projected keys
  int M, d;                 // modes, per-mode dim (M*d == 256)
  float scale, w_agg, w_pos;
  const float* pos_table;   // [(2R+1)^2] or nullptr
  int Rb;                   // bias radius
  const float* clip;        // device scalar
  const float* coords;      // [Mp,2]
  const float* stats;       // {mean, rstd}
  act_t* out_b;     // [Mp, ldb], channels 0..80
  int ldb;
  float* out_nchw;          // [324,H,W] (channels 0..80) or nullptr
  int cap;                  // cells of shared memory available for staging a patch's bounding box (0: read from L2)
};

// Block = 8 warps = a 4 x 2 patch of queries, one warp per query; every window cell is one coalesced 512-byte
// key row from L2 (51 KB per query, 367 MB per launch at 448x1024).
// Optional staging (p.cap > 0, CRAFT_LOOKUP0_CAP): the windows of neighbouring queries overlap almost completely
// wherever the flow is smooth (a 4 x 2 patch with locally constant flow needs 13 x 11 = 143 distinct key rows,
// not 8 x 100), so the block can stage the BOUNDING BOX of its eight windows in shared memory with cp.async and
// serve every warp from there; patches whose windows scatter fall back to L2.  Measured in round 2: bit-identical
// and SLOWER (55 us vs 45.6 us): the shared memory costs occupancy and the kernel is bound by the latency of its
// load -> FMA -> shuffle chains, not by L2 bandwidth.  Kept switchable, off by default.
constexpr int kL0Cap = 192;
constexpr int kL0PX = 4, kL0PY = 2;

__global__ void __launch_bounds__(256) corr_lookup0_kernel(Lookup0Params p, Grid2 g) {
  pdl_launch_dependents();
  pdl_wait();

  constexpr int R = 4, D = 2 * R + 1, WN = D + 1, C = 256, NC = WN * WN;
  extern __shared__ uint4 s_rows[];                 // [kL0Cap][32] staged key rows
  __shared__ float win[8][NC + 4];
  __shared__ float smodes[8][NC][4];
  __shared__ float tbl[232];
  __shared__ int s_org[8][2];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.pos_table) {
    const int n = (2 * p.Rb + 1) * (2 * p.Rb + 1);
    for (int i = threadIdx.x; i < n; i += 256) tbl[i] = p.pos_table[i] * p.w_pos;
  }
  const int npx = (g.W + kL0PX - 1) / kL0PX;
  const int py = blockIdx.x / npx, px = blockIdx.x - py * npx;
  const int qx = px * kL0PX + (wib % kL0PX), qy = py * kL0PY + (wib / kL0PX);
  const bool qvalid = qx < g.W && qy < g.H;
  const int q = qy * g.Wp + qx;
  float cx = 0.f, cy = 0.f;
  if (qvalid) { cx = p.coords[2 * q]; cy = p.coords[2 * q + 1]; }
  const float fx0 = floorf(cx), fy0 = floorf(cy);
  const float ax = cx - fx0, ay = cy - fy0;
  // window origin, clamped so that far-off-image coordinates cannot overflow the box arithmetic (their
  // windows are entirely out of bounds either way)
  const int x0 = static_cast<int>(fminf(fmaxf(fx0, -32768.f), 32768.f)) - R;
  const int y0 = static_cast<int>(fminf(fmaxf(fy0, -32768.f), 32768.f)) - R;
  if (lane == 0) { s_org[wib][0] = qvalid ? x0 : INT_MAX; s_org[wib][1] = qvalid ? y0 : INT_MAX; }
  __syncthreads();
  int xmin = INT_MAX, ymin = INT_MAX, xmax = INT_MIN, ymax = INT_MIN;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ox = s_org[i][0], oy = s_org[i][1];
    if (ox != INT_MAX) { xmin = min(xmin, ox); xmax = max(xmax, ox); ymin = min(ymin, oy); ymax = max(ymax, oy); }
  }
  if (xmin == INT_MAX) return;                      // no real query in this patch (block-uniform)
  // only the part of the box inside the key grid is staged; cells outside it are zeros by definition
  const int bx0 = max(xmin, 0), by0 = max(ymin, 0);
  const int bx1 = min(xmax + WN, g.W), by1 = min(ymax + WN, g.H);     // exclusive
  const int bw = max(bx1 - bx0, 0), bh = max(by1 - by0, 0);
  const long long cells = static_cast<long long>(xmax - xmin + WN) * (ymax - ymin + WN);
  const bool staged = cells <= p.cap;               // block-uniform (conservative: unclipped box); cap 0 = never
  if (staged) {
    // asynchronous copies (LDGSTS): all rows of a warp are in flight at once, no register round trip
    for (int c = wib; c < bw * bh; c += 8) {
      const int ky = by0 + c / bw, kx = bx0 + c % bw;
      const uint4* src = reinterpret_cast<const uint4*>(p.K + (static_cast<size_t>(ky) * g.Wp + kx) * C) + lane;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&s_rows[c * 32 + lane])), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (!qvalid) return;                              // whole warp
  // this lane's 8 query channels, fp32
  float qf[8];
  {
    const uint4 raw = *reinterpret_cast<const uint4*>(p.Q + static_cast<size_t>(q) * C + lane * 8);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 v = unpack_act2(w[k]);
      qf[2 * k] = v.x;
      qf[2 * k + 1] = v.y;
    }
  }
  const float mean = p.stats[0], rstd = p.stats[1];
  const float clipv = *p.clip;
  const int lanes_per_mode = p.d >> 3;    // 8 channels per lane; 8, 16 or 32 (d = 64, 128, 256)
  const bool is_writer = (lane & (lanes_per_mode - 1)) == 0;
  const int mode_of_lane = (lanes_per_mode == 8) ? (lane >> 3) : (lanes_per_mode == 16 ? (lane >> 4) : 0);
  // phase 1: one 512-byte key row per window cell (lane = 8 channels); per-mode dot products by shuffle.
  // The 10 rows of a window row are fetched back to back (memory-level parallelism).
#pragma unroll 1
  for (int r = 0; r < WN; ++r) {
    const int ky = y0 + r;                  // warp-uniform
    if (ky < 0 || ky >= g.H) continue;
    uint4 kk[WN];
    if (staged) {
      const int rowi = (ky - by0) * bw - bx0;      // + kx = cell index inside the staged box
#pragma unroll
      for (int c = 0; c < WN; ++c) {
        const int kx = x0 + c;
        kk[c] = (kx >= 0 && kx < g.W) ? s_rows[(rowi + kx) * 32 + lane] : make_uint4(0, 0, 0, 0);
      }
    } else {
      const uint4* rowbase = reinterpret_cast<const uint4*>(p.K + (static_cast<size_t>(ky) * g.Wp) * C) + lane;
#pragma unroll
      for (int c = 0; c < WN; ++c) {
        const int kx = x0 + c;
        kk[c] = (kx >= 0 && kx < g.W) ? __ldg(rowbase + static_cast<size_t>(kx) * (C / 8)) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int c = 0; c < WN; ++c) {
      const float2 k0 = unpack_act2(kk[c].x), k1 = unpack_act2(kk[c].y), k2 = unpack_act2(kk[c].z), k3 = unpack_act2(kk[c].w);
      float a = qf[0] * k0.x + qf[1] * k0.y + qf[2] * k1.x + qf[3] * k1.y +
                qf[4] * k2.x + qf[5] * k2.y + qf[6] * k3.x + qf[7] * k3.y;
      a += __shfl_xor_sync(0xffffffffu, a, 1);           // d >= 64 always (M <= 4): 8 lanes per mode at least
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      if (lanes_per_mode > 8) a += __shfl_xor_sync(0xffffffffu, a, 8);
      if (lanes_per_mode > 16) a += __shfl_xor_sync(0xffffffffu, a, 16);
      if (is_writer) smodes[wib][r * WN + c][mode_of_lane] = a;
    }
  }
  __syncwarp();
  // phase 2: lane = cell: clamp, soft-aggregate the modes, positional bias, deferred-LN numerator
  const float wl2 = p.w_agg * 1.4426950408889634f;
  const int TD = 2 * p.Rb + 1;
  float* wv = win[wib];
  for (int e = lane; e < NC; e += 32) {
    const int r = e / WN, c = e - r * WN;
    const int ky = y0 + r, kx = x0 + c;
    float val = 0.f;
    if (ky >= 0 && ky < g.H && kx >= 0 && kx < g.W) {
      if (p.M == 1) {
        val = fminf(fmaxf(smodes[wib][e][0] * p.scale, -clipv), clipv);
      } else {
        float s[4], t[4];
        float tm = -INFINITY;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          s[m] = (m < p.M) ? fminf(fmaxf(smodes[wib][e][m] * p.scale, -clipv), clipv) : 0.f;
          t[m] = s[m] * wl2;                      // w_agg is signed: unused modes must not enter the max
          if (m < p.M) tm = fmaxf(tm, t[m]);
        }
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (m < p.M) {
            const float em = fast_ex2(t[m] - tm);
            num += em * s[m];
            den += em;
          }
        }
        val = __fdividef(num, den);
      }
      if (p.pos_table) {
        const int dy = ky - qy, dx = kx - qx;
        if (dy >= -p.Rb && dy <= p.Rb && dx >= -p.Rb && dx <= p.Rb) val += tbl[(dy + p.Rb) * TD + dx + p.Rb];
      }
      val -= mean;
    }
    wv[e] = val;
  }
  __syncwarp();
  for (int e = lane; e < D * D; e += 32) {
    const int i = e / D, j = e - i * D;
    const float v00 = wv[j * WN + i], v01 = wv[j * WN + i + 1];
    const float v10 = wv[(j + 1) * WN + i], v11 = wv[(j + 1) * WN + i + 1];
    const float top = v00 + ax * (v01 - v00);
    const float bot = v10 + ax * (v11 - v10);
    const float val = (top + ay * (bot - top)) * rstd;
    if (p.out_b) p.out_b[static_cast<size_t>(q) * p.ldb + e] = f2act(val);
    if (p.out_nchw) p.out_nchw[(static_cast<size_t>(e) * g.H + qy) * g.W + qx] = val;
  }
}

// -------------------------------------------------------------------------------------------
// upsample_flow: convex 8x upsampling.
//   reference: CRAFT.upsample_flow core/network.py:151-162.  mask channel = k*64 + sy*8 + sx,
//   k = 3x3 neighbour (row-major, zero padded), softmax over k, output pixel (8y+sy, 8x+sx).
// mask: token-major [Mp, ldm] (already scaled by 0.25 by the mask head), flow: [Mp, 2] f32.
// Block = 256 threads = 4 tokens x 64 sub-pixels.
// -------------------------------------------------------------------------------------------
template <typename TM>
__global__ void __launch_bounds__(256) upsample_flow_kernel(const TM* __restrict__ mask, int ldm,
                                                            const float* __restrict__ flow, Grid2 g,
                                                            float* __restrict__ out /*[2,8H,8W]*/) {
  pdl_launch_dependents();
  pdl_wait();

  const int sub = threadIdx.x & 63;
  const int tok = blockIdx.x * 4 + (threadIdx.x >> 6);   // index over real tokens (y*W + x)
  if (tok >= g.H * g.W) return;
  const int y = tok / g.W, x = tok - y * g.W;
  const size_t p = static_cast<size_t>(y) * g.Wp + x;
  float m[9];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    m[k] = static_cast<float>(mask[p * ldm + k * 64 + sub]);
    mx = fmaxf(mx, m[k]);
  }
  float den = 0.f, ox = 0.f, oy = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int dy = k / 3 - 1, dx = k % 3 - 1;
    const int yy = y + dy, xx = x + dx;
    const float e = __expf(m[k] - mx);
    den += e;
    if (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) {
      const size_t pn = static_cast<size_t>(yy) * g.Wp + xx;
      ox += e * flow[2 * pn];
      oy += e * flow[2 * pn + 1];
    }
  }
  const float s = 8.0f / den;
  const int sy = sub >> 3, sx = sub & 7;
  const size_t HW8 = static_cast<size_t>(8 * g.H) * (8 * g.W);
  const size_t o = static_cast<size_t>(8 * y + sy) * (8 * g.W) + 8 * x + sx;
  out[o] = ox * s;
  out[HW8 + o] = oy * s;
}

// -------------------------------------------------------------------------------------------
// convf1: 7x7 conv 2 -> 128 + ReLU on the flow field (K = 98: too thin for a tensor-core tile).
//   reference: BasicMotionEncoder.convf1 core/update.py:75,82.
// weights wt: [98][128] f32 (tap-major: (c*49 + ky*7 + kx), cout), bias [128].
// Block: 128 threads (= cout), 16 consecutive tokens of one grid row.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) convf1_kernel(const float* __restrict__ flow /*[Mp,2]*/,
                                                     const float* __restrict__ wt,
                                                     const float* __restrict__ bias, Grid2 g,
                                                     act_t* __restrict__ out, int ldo, int colo) {
  pdl_launch_dependents();
  pdl_wait();

  constexpr int TX = 8;     // tokens per thread; block covers 16 tokens x 64 couts (gridDim.y = 2)
  __shared__ float in[2][7][16 + 6];
  __shared__ __align__(16) float ws[98 * 64];
  const int tiles_per_row = (g.W + 15) / 16;
  const int y = blockIdx.x / tiles_per_row;
  const int xb = (blockIdx.x - y * tiles_per_row) * 16;
  const int co = blockIdx.y * 64 + (threadIdx.x & 63);
  const int tg = threadIdx.x >> 6;
  const int x0 = xb + tg * TX;
  {
    // this block's half of the weights: 98 rows of 64 contiguous floats.  All 16-byte requests of a thread are
    // issued before the first store (13 per thread): the scalar copy loop this replaces ran its 49 L2 round trips
    // nearly back to back and was most of the kernel's 16 us.
    const float4* src = reinterpret_cast<const float4*>(wt) + blockIdx.y * 16;
    float4* dst = reinterpret_cast<float4*>(ws);
    float4 t[13];
#pragma unroll
    for (int u = 0; u < 13; ++u) {
      const int i = threadIdx.x + u * 128;             // float4 index: row i >> 4, column group i & 15
      t[u] = (i < 98 * 16) ? __ldg(src + (i >> 4) * 32 + (i & 15)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 13; ++u) {
      const int i = threadIdx.x + u * 128;
      if (i < 98 * 16) dst[i] = t[u];
    }
  }
  for (int i = threadIdx.x; i < 2 * 7 * 22; i += 128) {
    const int c = i / (7 * 22);
    const int r = (i / 22) % 7;
    const int cx = i % 22;
    const int yy = y + r - 3, xx = xb + cx - 3;
    float v = 0.f;
    if (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) v = flow[(static_cast<size_t>(yy) * g.Wp + xx) * 2 + c];
    in[c][r][cx] = v;
  }
  __syncthreads();
  float acc[TX];
  const float b = bias[co];
#pragma unroll
  for (int t = 0; t < TX; ++t) acc[t] = b;
  for (int c = 0; c < 2; ++c)
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float w = ws[(c * 49 + ky * 7 + kx) * 64 + (threadIdx.x & 63)];
#pragma unroll
        for (int t = 0; t < TX; ++t) acc[t] += w * in[c][ky][tg * TX + t + kx];
      }
  for (int t = 0; t < TX; ++t) {
    const int x = x0 + t;
    if (x < g.W)
      out[(static_cast<size_t>(y) * g.Wp + x) * ldo + colo + co] = f2act(fmaxf(acc[t], 0.f));
  }
}

// -------------------------------------------------------------------------------------------
// flow_update: coords1 += delta ; flow = coords1 - coords0 (coords0 is the token's own (x,y)).
//   reference: core/network.py:236,247.  delta: [Mp, ldd] f32 (cols 0,1).  Halo rows stay zero.
// -------------------------------------------------------------------------------------------
__global__ void flow_update_kernel(float* __restrict__ coords1, float* __restrict__ flow,
                                   const float* __restrict__ delta, int ldd, Grid2 g) {
  pdl_launch_dependents();
  pdl_wait();

  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.Mp) return;
  const int y = p / g.Wp, x = p - y * g.Wp;
  if (x >= g.W) return;
  float cx = coords1[2 * p], cy = coords1[2 * p + 1];
  if (delta) {
    cx += delta[static_cast<size_t>(p) * ldd];
    cy += delta[static_cast<size_t>(p) * ldd + 1];
    coords1[2 * p] = cx;
    coords1[2 * p + 1] = cy;
  }
  flow[2 * p] = cx - static_cast<float>(x);
  flow[2 * p + 1] = cy - static_cast<float>(y);
}

// init_coords: coords1 = grid (+ flow_init NCHW [2,H,W]); reference core/network.py:142-149,221-222.
__global__ void init_coords_kernel(float* __restrict__ coords1, const float* __restrict__ flow_init,
                                   Grid2 g) {
  pdl_launch_dependents();
  pdl_wait();

  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.Mp) return;
  const int y = p / g.Wp, x = p - y * g.Wp;
  float cx = 0.f, cy = 0.f;
  if (x < g.W) {
    cx = static_cast<float>(x);
    cy = static_cast<float>(y);
    if (flow_init) {
      cx += flow_init[static_cast<size_t>(y) * g.W + x];
      cy += flow_init[static_cast<size_t>(g.H) * g.W + static_cast<size_t>(y) * g.W + x];
    }
  }
  coords1[2 * p] = cx;
  coords1[2 * p + 1] = cy;
}

// -------------------------------------------------------------------------------------------
// modes_finalize: collapse M per-mode aggregated features into one, add the input skip and
// LayerNorm.   reference: ExpandedFeatTrans.forward core/setrans.py:395-407 with
// LearnedSoftAggregate (num_feat = F) core/setrans.py:289-300:
//     s_m = <w, O_m> + b ; p = softmax_m(s) ; agg = sum_m p_m O_m ; y = LN(coeff * x + agg)
// GMA variant (core/gma.py:140): M = 1, y = x + gamma * O  (no LN) -- `gma` != 0.
// O: [slot][M][F/8][Mp][8] f32 (8-column chunks, token-major inside a chunk: the layout attn_pv's
// one-row-per-thread write-back can store with full 32-byte sectors).  x: token-major bf16 (ldx, colx).
// One warp per token; lane l owns columns 8*(l/2 + 16j) + 4*(l%2) + {0..3}, j < F/128.
// -------------------------------------------------------------------------------------------
template <int F>
__global__ void __launch_bounds__(256) modes_finalize_kernel(
    const float* __restrict__ O, int nsum, long long part_stride, int M, long long mode_stride,
    const float* __restrict__ w_score,
    const float* __restrict__ b_score, const float* __restrict__ coeff, int gma,
    const act_t* __restrict__ xb, int ldx, int colx, const float* __restrict__ xf, int ldxf,
    int colxf, Grid2 g, act_t* __restrict__ out_b, int ldb, int colb,
    float* __restrict__ out_f, int ldf, int colf, int pv_G, int pv_nkt) {
  pdl_launch_dependents();
  pdl_wait();

  // pv_G > 0: O was written by attn_pv's persistent schedule with pv_G CTAs and pv_nkt key tiles per
  // unit; a (query tile, mode) unit then owns as many valid slots as CTAs shared it (1 or 2, rarely
  // more) and the other slots hold garbage -- recompute that count instead of reading them.
  constexpr int PER = F / 32;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  // valid slots per mode: the block's 8 tokens share one 128-query tile, so 4 threads work it out once
  __shared__ int s_nvalid[4];
  if (threadIdx.x < 4) {
    const int m = threadIdx.x;
    int nvalid = nsum;
    if (pv_G > 0 && m < M) {
      // the host guarantees NT * pv_G < 2^31 (craft_modes_finalize), so 32-bit arithmetic is exact
      const unsigned NT = static_cast<unsigned>((g.Mp + 127) >> 7) * M * pv_nkt;
      const unsigned G = static_cast<unsigned>(pv_G);
      auto cta_of = [&](unsigned x) {
        unsigned c = x * G / NT;
        while (c + 1 < G && NT * (c + 1) / G <= x) ++c;
        while (c > 0 && NT * c / G > x) --c;
        return static_cast<int>(c);
      };
      const unsigned lin0 = (static_cast<unsigned>((blockIdx.x * 8) >> 7) * M + m) * pv_nkt;
      nvalid = cta_of(lin0 + pv_nkt - 1) - cta_of(lin0) + 1;
    }
    s_nvalid[m] = nvalid;
  }
  __syncthreads();
  if (p >= g.Mp) return;
  const int y = p / g.Wp, x = p - y * g.Wp;
  if (x >= g.W) return;
  // The kernel is bound by its instruction count (ncu: 685 warp instructions per token, issue slots the busiest
  // unit), so everything that does not depend on the mode is hoisted and every access is a vector.
  float o[4][PER];
  float sc[4];
  float wsc[PER];                          // this lane's score weights: the same for every mode
#pragma unroll
  for (int j = 0; j < PER / 4; ++j) {
    // scalar loads: w_score is a module parameter, and nn.DataParallel replicas hold their parameters as views into
    // one flat broadcast buffer -- aligned to 4 bytes only (a float4 load here trapped on the second GPU)
    const float* wp = w_score + 8 * ((lane >> 1) + 16 * j) + 4 * (lane & 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) wsc[4 * j + k] = gma ? 0.f : __ldg(wp + k);
  }
  const float* src0 = O + (static_cast<long long>(lane >> 1) * g.Mp + p) * 8 + 4 * (lane & 1);
  float part_s[4];                         // per-mode partial dot products, reduced over the warp together below
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    sc[m] = -INFINITY;
    part_s[m] = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) o[m][e] = 0.f;
    if (m < M) {
      const int nvalid = s_nvalid[m];      // warp-uniform: 1 or 2, rarely more (nsum <= 4, checked by the host)
#pragma unroll
      for (int j = 0; j < PER / 4; ++j) {
        const float* src = src0 + m * mode_stride + static_cast<long long>(16 * j) * g.Mp * 8;
        float4 acc = __ldg(reinterpret_cast<const float4*>(src));
        for (int sp = 1; sp < nvalid; ++sp) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src + sp * part_stride));
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        o[m][4 * j] = acc.x; o[m][4 * j + 1] = acc.y; o[m][4 * j + 2] = acc.z; o[m][4 * j + 3] = acc.w;
        part_s[m] += acc.x * wsc[4 * j] + acc.y * wsc[4 * j + 1] + acc.z * wsc[4 * j + 2] + acc.w * wsc[4 * j + 3];
      }
    }
  }
  if (!gma) {
    const float bsc = b_score[0];
#pragma unroll
    for (int o_ = 16; o_ > 0; o_ >>= 1) {      // the four reductions share one shuffle ladder
#pragma unroll
      for (int m = 0; m < 4; ++m) part_s[m] += __shfl_xor_sync(0xffffffffu, part_s[m], o_);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) if (m < M) sc[m] = part_s[m] + bsc;
  } else {
#pragma unroll
    for (int m = 0; m < 4; ++m) if (m < M) sc[m] = 0.f;
  }
  const float mx = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
  float den = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    sc[m] = (m < M) ? __expf(sc[m] - mx) : 0.f;
    den += sc[m];
  }
  const float c = coeff[0];
  const float rden = 1.0f / den;
  float yv[PER];
  float s1 = 0.f;
#pragma unroll
  for (int j = 0; j < PER / 4; ++j) {
    const int f0 = 8 * ((lane >> 1) + 16 * j) + 4 * (lane & 1);      // four consecutive channels: one vector access
    float xin[4];
    if (xf) {
      const float4 x4 = *reinterpret_cast<const float4*>(xf + static_cast<size_t>(p) * ldxf + colxf + f0);
      xin[0] = x4.x; xin[1] = x4.y; xin[2] = x4.z; xin[3] = x4.w;
    } else {
      const uint2 x2 = *reinterpret_cast<const uint2*>(xb + static_cast<size_t>(p) * ldx + colx + f0);
      const float2 lo = unpack_act2(x2.x), hi = unpack_act2(x2.y);
      xin[0] = lo.x; xin[1] = lo.y; xin[2] = hi.x; xin[3] = hi.y;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = 4 * j + k;
      float agg = 0.f;
#pragma unroll
      for (int m = 0; m < 4; ++m) agg += sc[m] * o[m][e];
      agg *= rden;
      yv[e] = gma ? (xin[k] + c * agg) : (c * xin[k] + agg);
      s1 += yv[e];
    }
  }
  if (!gma) {
    const float mean = warp_sum(s1) * (1.0f / F);
    float s2 = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const float d = yv[e] - mean;
      s2 += d * d;
    }
    const float rstd = rsqrtf(warp_sum(s2) * (1.0f / F) + 1e-12f);
#pragma unroll
    for (int e = 0; e < PER; ++e) yv[e] = (yv[e] - mean) * rstd;
  }
#pragma unroll
  for (int j = 0; j < PER / 4; ++j) {
    const int f0 = 8 * ((lane >> 1) + 16 * j) + 4 * (lane & 1);
    if (out_b)
      *reinterpret_cast<uint2*>(out_b + static_cast<size_t>(p) * ldb + colb + f0) =
          make_uint2(pack_act2(yv[4 * j], yv[4 * j + 1]), pack_act2(yv[4 * j + 2], yv[4 * j + 3]));
    if (out_f)
      *reinterpret_cast<float4*>(out_f + static_cast<size_t>(p) * ldf + colf + f0) =
          make_float4(yv[4 * j], yv[4 * j + 1], yv[4 * j + 2], yv[4 * j + 3]);
  }
}

}  // namespace cb
