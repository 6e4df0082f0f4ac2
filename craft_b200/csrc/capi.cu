// craft_b200 -- C ABI (include/craft_b200.h): argument checking, TMA tensor-map construction and
// kernel launches.  No torch, no global mutable state besides the thread-local error string.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <atomic>

#include "../../include/craft_b200.h"
#include "attn_pv.cuh"
#include "common.cuh"
#include "encoder.cuh"
#include "conv_enc.cuh"
#include "gemm.cuh"
#include "hostio.cuh"
#include "pointwise.cuh"
#include "scores.cuh"
#include "standalone.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};   // kernels launched by this library (bench.py's gpu_launches)

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Every kernel of the library is launched with programmatic dependent launch (PDL): the next
// kernel of the stream may be scheduled while this one drains, runs its prologue (barrier init,
// TMEM allocation, table loads from constant parameters) and blocks in griddepcontrol.wait until
// the predecessor has completed and flushed.  Inside a captured CUDA graph this becomes a
// programmatic edge.  CRAFT_B200_NO_PDL=1 restores plain stream order.
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CRAFT_B200_NO_PDL"); v = (e && atoi(e) != 0) ? 0 : 1; }
  return v == 1;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

cb::Grid2 make_grid(int H, int W) {
  cb::Grid2 g;
  g.H = H;
  g.W = W;
  g.Wp = W + 2;
  g.Mp = H * (W + 2);
  return g;
}

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed) ------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 map over a row-major [rows, ld] matrix, box = [box_rows, 64 cols], 128B swizzle.
int make_map_2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld,
                int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16) return fail("tensor map: base/stride not 16B aligned");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(2d) failed: %d", static_cast<int>(r));
  return 0;
}
// 3-D bf16 map over a row-major [rows, ld] matrix viewed as [64 cols][rows][ld/64 chunks]: one box
// [64][box_rows][2] lands as two consecutive 128-byte-swizzled K atoms (experimental big-box GEMM mode).
// Returns 1 (not an error) when the driver rejects the encoding, so the caller can fall back.
int make_map_3d_k2(CUtensorMap* m, const void* base, long long rows, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || ld % 64) return 1;
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(ld / 64)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}
std::atomic<long long> g_bigbox_launches{0};   // GEMM launches that really took the big-box path

// 3-D bf16 map over token rows viewed as [H][W][C] with row pitch Wp: box = [8][bw][64].
int make_map_3d_keys(CUtensorMap* m, const void* base, const cb::Grid2& g, int C, int bw = 8) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(g.W), static_cast<cuuint64_t>(g.H)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(g.Wp) * C * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(bw), 8};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(3d) failed: %d", static_cast<int>(r));
  return 0;
}

// Per-device caches: a process may drive several GPUs (nn.DataParallel replicates the module, one
// thread per device), and both the SM count and cudaFuncSetAttribute are per device.
constexpr int kMaxDevices = 64;
int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = cur_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (!v) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// Opt a kernel in to `bytes` of dynamic shared memory once per device; `done` is the call site's
// own bitmap (one bit per device).
template <typename K>
int ensure_smem(K kern, int bytes, std::atomic<unsigned long long>& done, const char* what) {
  const int dev = cur_device();
  if ((done.load(std::memory_order_acquire) >> dev) & 1ull) return 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
    return fail("%s: cannot opt in to %d bytes of dynamic shared memory", what, bytes);
  done.fetch_or(1ull << dev, std::memory_order_release);
  return 0;
}

template <int BN, int EPI, int CL>
int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, const cb::GemmParams& p, cudaStream_t st) {
  using S = cb::GemmSmem<BN>;
  static std::atomic<unsigned long long> attr_set{0};
  auto kern = cb::shift_gemm_kernel<BN, EPI, CL>;
  if (ensure_smem(kern, S::kTotal, attr_set, "shift_gemm")) return -1;
  const int mtiles = (p.M + cb::kGemmBM - 1) / cb::kGemmBM;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(((mtiles + CL - 1) / CL) * CL, p.Npad / BN, 1);   // M tiles padded to whole clusters
  cfg.blockDim = dim3(cb::kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
  if (e != cudaSuccess) return fail("shift_gemm launch: %s", cudaGetErrorString(e));
  return check_launch("shift_gemm");
}

template <int BN, int EPI>
int launch_gemm_cl(int CL, const CUtensorMap& ta, const CUtensorMap& tb, const cb::GemmParams& p, cudaStream_t st) {
  if constexpr (BN >= 32) {
    if (CL == 4) return launch_gemm_t<BN, EPI, 4>(ta, tb, p, st);
  }
  if (CL == 2) return launch_gemm_t<BN, EPI, 2>(ta, tb, p, st);
  return launch_gemm_t<BN, EPI, 1>(ta, tb, p, st);
}

template <int EPI>
int launch_gemm_bn(int BN, int CL, const CUtensorMap& ta, const CUtensorMap& tb, const cb::GemmParams& p, cudaStream_t st) {
  switch (BN) {
    case 32: return launch_gemm_cl<32, EPI>(CL, ta, tb, p, st);
    case 64: return launch_gemm_cl<64, EPI>(CL, ta, tb, p, st);
    case 96:
      if constexpr (EPI == cb::EPI_STORE) return launch_gemm_t<96, EPI, 1>(ta, tb, p, st);   // plain epilogue only, no multicast
      break;
    case 128: return launch_gemm_cl<128, EPI>(CL, ta, tb, p, st);
    case 256: return launch_gemm_cl<256, EPI>(CL, ta, tb, p, st);
  }
  return fail("gemm: unsupported BN %d", BN);
}

}  // namespace

extern "C" {

int craft_b200_abi_version(void) { return CRAFT_B200_ABI_VERSION; }
const char* craft_b200_last_error(void) { return g_err; }
long long craft_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
long long craft_b200_bigbox_gemm_count(void) { return g_bigbox_launches.load(std::memory_order_relaxed); }

int craft_b200_device_info(int* out3) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail("no CUDA device");
  cudaDeviceGetAttribute(&out3[0], cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&out3[1], cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&out3[2], cudaDevAttrComputeCapabilityMinor, dev);
  return 0;
}

int craft_pack_tokens(const float* src, int C, int H, int W, int mode, void* out_b, int ldb, int colb,
                      float* out_f, int ldf, int colf, void* stream) {
  cb::Grid2 g = make_grid(H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(H * ((g.Wp + 31) / 32));
  auto* ob = static_cast<cb::act_t*>(out_b);
  if (C == 128) launch_k(cb::pack_tokens_kernel<128>, dim3(grid), dim3(256), 0, st, src, g, mode, ob, ldb, colb, out_f, ldf, colf);
  else if (C == 256) launch_k(cb::pack_tokens_kernel<256>, dim3(grid), dim3(256), 0, st, src, g, mode, ob, ldb, colb, out_f, ldf, colf);
  else return fail("pack_tokens: C must be 128 or 256 (got %d)", C);
  return check_launch("pack_tokens");
}

int craft_pack_tokens_nhwc(const void* src, int src_is_half, int ldc, int c0, int C, int H, int W, int mode,
                           void* out_b, int ldb, int colb, float* out_f, int ldf, int colf, void* stream) {
  cb::Grid2 g = make_grid(H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (C != 128 && C != 256) return fail("pack_tokens_nhwc: C must be 128 or 256 (got %d)", C);
  if (c0 % 8 || ldc % 8 || c0 + C > ldc) return fail("pack_tokens_nhwc: channel window [%d, %d) of %d must be 8-aligned", c0, c0 + C, ldc);
  if ((out_b && (ldb % 8 || colb % 8)) || (out_f && (ldf % 4 || colf % 4))) return fail("pack_tokens_nhwc: output ld/col alignment");
  dim3 grid((g.Mp + 7) / 8);
  auto* ob = static_cast<cb::act_t*>(out_b);
  if (src_is_half) {
    auto* s = static_cast<const __half*>(src);
    if (C == 128) launch_k(cb::pack_tokens_nhwc_kernel<__half, 128>, grid, dim3(256), 0, st, s, ldc, c0, g, mode, ob, ldb, colb, out_f, ldf, colf);
    else launch_k(cb::pack_tokens_nhwc_kernel<__half, 256>, grid, dim3(256), 0, st, s, ldc, c0, g, mode, ob, ldb, colb, out_f, ldf, colf);
  } else {
    auto* s = static_cast<const float*>(src);
    if (C == 128) launch_k(cb::pack_tokens_nhwc_kernel<float, 128>, grid, dim3(256), 0, st, s, ldc, c0, g, mode, ob, ldb, colb, out_f, ldf, colf);
    else launch_k(cb::pack_tokens_nhwc_kernel<float, 256>, grid, dim3(256), 0, st, s, ldc, c0, g, mode, ob, ldb, colb, out_f, ldf, colf);
  }
  return check_launch("pack_tokens_nhwc");
}

int craft_unpack_tokens(const void* src, int is_bf16, int ld, int col, int C, int H, int W, float* dst,
                        void* stream) {
  cb::Grid2 g = make_grid(H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(H * ((W + 31) / 32), (C + 31) / 32);
  if (is_bf16)
    launch_k(cb::unpack_tokens_kernel<cb::act_t>, dim3(grid), dim3(256), 0, st, static_cast<const cb::act_t*>(src), ld, col, C, g, dst);
  else
    launch_k(cb::unpack_tokens_kernel<float>, dim3(grid), dim3(256), 0, st, static_cast<const float*>(src), ld, col, C, g, dst);
  return check_launch("unpack_tokens");
}

static int launch_gemm_dispatch(const craft_gemm_args* a, int CL, const CUtensorMap& ta, const CUtensorMap& tb,
                                const cb::GemmParams& p, cudaStream_t st);

int craft_shift_gemm(const craft_gemm_args* a, void* stream) {
  if (!a || !a->A || !a->B) return fail("gemm: null operand");
  if (a->K <= 0 || a->K % 64) return fail("gemm: K=%d must be a positive multiple of 64", a->K);
  if (a->T < 1 || a->T > CRAFT_MAX_TAPS) return fail("gemm: T=%d out of range", a->T);
  if (a->Npad % a->BN) return fail("gemm: Npad=%d not a multiple of BN=%d", a->Npad, a->BN);
  if (a->a_koff % 64 || a->b_koff % 64) return fail("gemm: koff must be multiples of 64");
  if (a->a_koff + a->K > a->lda || a->b_koff + a->K > a->ldb_) return fail("gemm: K range exceeds operand width");
  // T == 1: B rows beyond b_rows are zero-filled by TMA (used for V^T = W X^T with a padded ld)
  if (a->T > 1 && a->b_rows < a->T * a->Npad) return fail("gemm: B has %d rows, needs %d", a->b_rows, a->T * a->Npad);
  if (a->out_bf16 && ((a->ldo_b % 8) || (a->colo_b % 8))) return fail("gemm: bf16 output ld/col must be multiples of 8");
  if (a->out_f32 && ((a->ldo_f % 4) || (a->colo_f % 4))) return fail("gemm: f32 output ld/col must be multiples of 4");
  CUtensorMap ta, tb;
  cb::GemmParams p;
  memset(&p, 0, sizeof(p));
  // shared-A mode (gemm.cuh): taps in groups of consecutive row offsets, e.g. the kw taps of a conv row
  {
    // CRAFT_GEMM_ASHARE: 0 = never, otherwise every kernel row of consecutive taps (1x5 and 3x3) -- measured
    // (profiles/r02_gemm_sweep_ashare.txt, r02_gemm_trace*.txt): 13.5 -> 12.1 us (GRU z/r), 10.7 -> 8.9 us (GRU q) for
    // five shared taps; the single-wave 3x3 convolutions 10.4 -> 8.9, 9.4 -> 7.8, 8.1 -> 6.8, 7.3 -> 6.6, 5.4 -> 4.6 us
    // (CTA durations).  Bit-exact either way (same MMAs, same order).
    static int env_mode = -2;
    if (env_mode == -2) { const char* e = getenv("CRAFT_GEMM_ASHARE"); env_mode = e ? (atoi(e) != 0 ? 1 : 0) : -1; }
    int gs = 1;
    while (gs < a->T && a->tap_off[gs] == a->tap_off[gs - 1] + 1) ++gs;
    const int mode = (a->a_share != 0 || env_mode != 0) ? 1 : 0;
    bool ok = mode != 0 && gs > 1 && gs <= 5 && a->T % gs == 0 && !a->b_blocked && a->cluster <= 1 && a->BN <= 128;
    if (ok) {      // at least two grouped stages must fit
      const int ring = a->BN == 32 ? cb::GemmSmem<32>::kRing : a->BN == 64 ? cb::GemmSmem<64>::kRing
                     : a->BN == 96 ? cb::GemmSmem<96>::kRing : cb::GemmSmem<128>::kRing;
      if (ring / (136 * 128 + gs * a->BN * 128) < 2) ok = false;
    }
    for (int t = 1; ok && t < a->T; ++t)
      if (t % gs != 0 && a->tap_off[t] != a->tap_off[t - 1] + 1) ok = false;
    if (ok) { p.ashare = mode; p.gsize = gs; }
  }
  if (make_map_2d(&ta, a->A, a->a_rows, a->lda, a->lda, p.ashare ? 136 : cb::kGemmBM)) return -1;
  if (a->b_blocked) {
    if (a->T != 1 || (a->BN != 64 && a->BN != 128)) return fail("gemm: blocked B needs T=1 and BN in {64,128}");
    cb::Grid2 bg = make_grid(a->b_H, a->b_W);
    if (a->b_rows < bg.Mp) return fail("gemm: blocked B has %d rows, grid needs %d", a->b_rows, bg.Mp);
    const int bw = a->BN / 8;
    p.b_blocked = 1;
    p.b_nbx = (a->b_W + bw - 1) / bw;
    if (a->Npad != p.b_nbx * ((a->b_H + 7) / 8) * a->BN) return fail("gemm: blocked B: Npad must be nblocks*BN");
    if (make_map_3d_keys(&tb, a->B, bg, a->ldb_, bw)) return -1;
  }
  // cluster size: CL consecutive M tiles share (multicast) the weight tile
  int CL = a->cluster > 0 ? a->cluster : 1;   // measured (profiles/gemm_sweep): multicast does not pay at M = 7280
  if (a->b_blocked) CL = 1;
  const int mtiles = (a->M + cb::kGemmBM - 1) / cb::kGemmBM;
  while (CL > 1 && (mtiles < CL || a->BN % (8 * CL))) CL /= 2;
  if (CL != 1 && CL != 2 && CL != 4) return fail("gemm: cluster must be 1, 2 or 4");
  if (!a->b_blocked) {
    if (make_map_2d(&tb, a->B, a->b_rows, a->ldb_, a->ldb_, a->BN / CL)) return -1;
  }
  // experimental big-box mode (CRAFT_GEMM_BIGBOX=1|2, gemm.cuh): replaces both maps by 3-D ones
  {
    static int env_big = -1;
    if (env_big < 0) { const char* e = getenv("CRAFT_GEMM_BIGBOX"); env_big = e ? atoi(e) : 0; }
    if ((env_big == 1 || env_big == 2) && !p.ashare && !a->b_blocked && CL == 1 && a->K % 128 == 0 && a->BN <= 256) {
      CUtensorMap ta3, tb3;
      if (make_map_3d_k2(&ta3, a->A, a->a_rows, a->lda, cb::kGemmBM) == 0 &&
          make_map_3d_k2(&tb3, a->B, a->b_rows, a->ldb_, a->BN) == 0) {
        ta = ta3; tb = tb3;
        p.bigbox = env_big;
        g_bigbox_launches.fetch_add(1, std::memory_order_relaxed);
      }
    }
  }
  p.M = a->M; p.Npad = a->Npad; p.K = a->K; p.T = a->T;
  p.a_koff = a->a_koff; p.b_koff = a->b_koff;
  for (int t = 0; t < a->T; ++t) p.tap_off[t] = a->tap_off[t];
  if (a->H > 0) { p.Wp = a->W + 2; p.W = a->W; p.H = a->H; } else { p.Wp = 0; p.W = 0; p.H = 0; }
  p.alpha = a->alpha; p.act = a->act; p.bias = a->bias;
  p.out_b = static_cast<cb::act_t*>(a->out_bf16); p.ldb = a->ldo_b; p.colb = a->colo_b;
  p.out_f = a->out_f32; p.ldf = a->ldo_f; p.colf = a->colo_f;
  p.aux_f0 = a->aux0; p.aux_f1 = a->aux1;
  {
    // smem holds `slots` (A atom + B atom) pairs; a pipeline stage is kc of them (two when K allows it:
    // the per-stage wait/commit round trip of the MMA warp is then paid once per 128 K columns)
    const int slots = a->BN >= 256 ? 4 : (a->BN >= 128 ? 6 : (a->BN >= 64 ? 8 : 10));
    const char* e = getenv("CRAFT_GEMM_KC");          // tuning/profiling override
    p.kc = (a->K % 128 == 0) ? 2 : 1;
    if (e && atoi(e) == 1 && !p.bigbox) p.kc = 1;
    if (p.bigbox) p.kc = 2;
    int smax = slots / p.kc;
    if (p.ashare) {      // grouped stages: as many (A tile + gsize weight tiles) as fit in the ring
      const int ring = a->BN == 32 ? cb::GemmSmem<32>::kRing : a->BN == 64 ? cb::GemmSmem<64>::kRing
                     : a->BN == 96 ? cb::GemmSmem<96>::kRing : cb::GemmSmem<128>::kRing;
      smax = ring / (136 * 128 + p.gsize * a->BN * 128);
      if (smax > slots) smax = slots;       // barrier block holds `slots` pairs
      if (smax < 2) { p.ashare = 0; p.gsize = 1; smax = slots / p.kc; }
    }
    p.stages = (a->stages > 0 && a->stages < smax) ? a->stages : smax;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // CRAFT_GEMM_TRACE=<file> (profiling aid, eager launches only): appends one record per launch -- the clock64
  // phase marks of CTA (0,0) and {globaltimer start, end, SM id} of every CTA.  Synchronises the stream.
  const char* trace_path = getenv("CRAFT_GEMM_TRACE");
  static long long* d_trace = nullptr;
  constexpr int kTraceLen = 16 + 3 * 1024;
  if (trace_path) {
    if (!d_trace) cudaMalloc(&d_trace, kTraceLen * sizeof(long long));
    cudaMemsetAsync(d_trace, 0, kTraceLen * sizeof(long long), st);
    p.trace = d_trace;
    const int rc = launch_gemm_dispatch(a, CL, ta, tb, p, st);
    static long long h[kTraceLen];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "a")) {
      const int nm = (a->M + cb::kGemmBM - 1) / cb::kGemmBM, nn = a->Npad / a->BN;
      fprintf(f, "gemm BN=%d epi=%d T=%d K=%d Npad=%d ashare=%d gsize=%d stages=%d kc=%d ctas=%d\nmarks", a->BN, a->epilogue, a->T,
              a->K, a->Npad, p.ashare, p.gsize, p.stages, p.kc, nm * nn);
      for (int k = 0; k < 12; ++k) fprintf(f, " %lld", h[k] ? h[k] - h[0] : -1ll);
      fprintf(f, "\n");
      for (int c = 0; c < nm * nn && c < 1024; ++c)
        fprintf(f, "cta %d %lld %lld %lld\n", c, h[16 + 3 * c], h[16 + 3 * c + 1], h[16 + 3 * c + 2]);
      fclose(f);
    }
    return rc;
  }
  return launch_gemm_dispatch(a, CL, ta, tb, p, st);
}

static int launch_gemm_dispatch(const craft_gemm_args* a, int CL, const CUtensorMap& ta, const CUtensorMap& tb,
                                const cb::GemmParams& p, cudaStream_t st) {
  switch (a->epilogue) {
    case cb::EPI_STORE: return launch_gemm_bn<cb::EPI_STORE>(a->BN, CL, ta, tb, p, st);
    case cb::EPI_GRU_ZR:
      if (a->Npad != 256 || !a->aux0 || !a->aux1 || !a->out_bf16) return fail("gemm: gru_zr needs Npad=256, Z, Hm, out");
      if (a->BN > 128) return fail("gemm: gru_zr needs BN <= 128");
      return launch_gemm_bn<cb::EPI_GRU_ZR>(a->BN, CL, ta, tb, p, st);
    case cb::EPI_GRU_Q:
      if (a->Npad != 128 || !a->aux0 || !a->aux1 || !a->out_bf16) return fail("gemm: gru_q needs Npad=128, Z, Hm, out");
      if (a->BN > 64) return fail("gemm: gru_q needs BN <= 64 (state prefetch lives in registers)");
      return launch_gemm_bn<cb::EPI_GRU_Q>(a->BN, CL, ta, tb, p, st);
    case cb::EPI_MOTION:
      if (a->Npad != 128 || !a->aux1 || !a->out_bf16) return fail("gemm: motion needs Npad=128, flow, out");
      return launch_gemm_bn<cb::EPI_MOTION>(a->BN, CL, ta, tb, p, st);
    case cb::EPI_FLOW:
      if (a->BN != 32 || a->Npad != 32 || !a->aux0 || !a->aux1 || !a->out_f32 || a->H <= 0)
        return fail("gemm: flow epilogue needs BN=Npad=32, coords1, flow, an f32 delta output and a grid");
      return launch_gemm_t<32, cb::EPI_FLOW, 1>(ta, tb, p, st);
  }
  return fail("gemm: unknown epilogue %d", a->epilogue);
}

// scores kernels run persistent CTAs over equal contiguous ranges of the (query tile, key tile) list
// (scores.cuh); the number of CTAs that can share one query tile is the number of partial slots.
static int sc_grid(int nqt, int nkt) {
  long long G = sm_count();
  if (G > static_cast<long long>(nqt) * nkt) G = static_cast<long long>(nqt) * nkt;
  return static_cast<int>(G < 1 ? 1 : G);
}
static int sc_slots(int nqt, int nkt) {
  const long long NT = static_cast<long long>(nqt) * nkt, G = sc_grid(nqt, nkt);
  auto cta_of = [&](long long x) {
    long long c = x * G / NT;
    while (c + 1 < G && NT * (c + 1) / G <= x) ++c;
    while (c > 0 && NT * c / G > x) --c;
    return c;
  };
  long long worst = 1;
  for (long long u = 0; u < nqt; ++u) {
    const long long parts = cta_of(u * nkt + nkt - 1) - cta_of(u * nkt) + 1;
    if (parts > worst) worst = parts;
  }
  return static_cast<int>(worst);
}
/* partial (max, sum) slots craft_attn_lse needs in lse_part */
int craft_scores_auto_ksplit(int H, int W) {
  cb::Grid2 g = make_grid(H, W);
  return sc_slots((g.Mp + 127) / 128, ((H + 7) / 8) * ((W + 7) / 8));
}
int craft_pv_block_keys(int d, int F) {
  if (d == 32 && F == 128) return 128;
  if (d == 64 && F == 256) return 64;
  if (d == 128 && F == 128) return 64;
  if (d == 64 && F == 128) return 128;
  return 0;
}
// attn_pv runs persistent CTAs over equal contiguous ranges of the (unit, key tile) list
// (attn_pv.cuh): grid size and the number of CTAs that can share one unit (= partial-sum slots).
static int pv_grid(int nqt, int M, int nkt) {
  const long long nunits = static_cast<long long>(nqt) * M;
  long long G = sm_count();
  if (G > nunits * nkt) G = nunits * nkt;
  if (G > 3 * nunits) G = 3 * nunits;      // keeps a unit within <= 4 CTAs (modes_finalize sums <= 4 slots)
  return static_cast<int>(G < 1 ? 1 : G);
}
static int pv_slots(int nqt, int M, int nkt) {
  const long long NT = static_cast<long long>(nqt) * M * nkt;
  const long long G = pv_grid(nqt, M, nkt);
  auto cta_of = [&](long long x) {
    long long c = x * G / NT;
    while (c + 1 < G && NT * (c + 1) / G <= x) ++c;
    while (c > 0 && NT * c / G > x) --c;
    return c;
  };
  long long worst = 1;
  for (long long u = 0; u < static_cast<long long>(nqt) * M; ++u) {
    const long long parts = cta_of(u * nkt + nkt - 1) - cta_of(u * nkt) + 1;
    if (parts > worst) worst = parts;
  }
  return static_cast<int>(worst);
}
/* number of partial-sum slots craft_attn_pv needs in `out` (covers both key-block widths) */
int craft_pv_auto_ksplit(int H, int W, int M) {
  cb::Grid2 g = make_grid(H, W);
  const int nqt = (g.Mp + 127) / 128;
  const int a = pv_slots(nqt, M, ((H + 7) / 8) * ((W + 15) / 16));
  const int b = pv_slots(nqt, M, ((H + 7) / 8) * ((W + 7) / 8));
  return a > b ? a : b;
}

}  // extern "C" (pause)
static int scores_common(const craft_scores_args* a, int mode, void* stream) {
  if (!a || !a->Q || !a->K) return fail("scores: null operand");
  if (a->C != a->M * a->d || a->C % 64 || a->C > 256) return fail("scores: bad C/M/d = %d/%d/%d", a->C, a->M, a->d);
  if (!(a->M == 1 || a->M == 2 || a->M == 4)) return fail("scores: M must be 1, 2 or 4");
  if (a->d % 16 || (a->d < 64 && 64 % a->d)) return fail("scores: d=%d unsupported", a->d);
  if (a->pos_table && (2 * a->R + 1) * (2 * a->R + 1) > 225) return fail("scores: pos radius too large");
  cb::Grid2 g = make_grid(a->H, a->W);
  CUtensorMap tq, tk;
  if (make_map_2d(&tq, a->Q, g.Mp, a->C, a->C, 128)) return -1;
  if (make_map_3d_keys(&tk, a->K, g, a->C)) return -1;
  cb::ScoreParams p;
  memset(&p, 0, sizeof(p));
  p.g = g; p.C = a->C; p.M = a->M; p.d = a->d; p.scale = a->scale; p.w_pos = a->w_pos;
  p.pos_table = a->pos_table; p.R = a->R; p.clip = a->clip; p.run_flag = a->run_flag;
  p.nkt_y = (a->H + 7) / 8; p.nkt_x = (a->W + 7) / 8;
  p.nqt = (g.Mp + 127) / 128;
  if (static_cast<long long>(p.nqt) * p.nkt_y * p.nkt_x > 0x7fffffffLL) return fail("scores: %d x %d tiles exceed 32-bit tile indices", p.nqt, p.nkt_y * p.nkt_x);
  const int need = sc_slots(p.nqt, p.nkt_y * p.nkt_x);
  p.nslots = a->ksplit > 0 ? a->ksplit : need;
  p.mask_radius = (mode == cb::SC_LSE) ? a->mask_radius : 0;
  if (mode == cb::SC_LSE && p.nslots < need) return fail("attn_lse: lse_part has %d slots, the schedule needs %d", p.nslots, need);
  p.w_agg = a->w_agg; p.stat_sum = a->stat_sum; p.stat_max = a->stat_max;
  int h = a->H, w = a->W;
  for (int l = 0; l < 4; ++l) {
    p.lvl[l] = a->lvl[l]; p.hl[l] = h; p.wl[l] = w;
    h /= 2; w /= 2;
  }
  p.lse_part = static_cast<float2*>(a->lse_part);
  p.lvl0h = static_cast<__half*>(a->lvl0_h16);
  p.l0_qstride = static_cast<long long>(p.nkt_y) * p.nkt_x * 64;
  p.lvl0_base = p.lvl0h ? reinterpret_cast<float*>(p.lvl0h + static_cast<long long>(g.Mp) * p.l0_qstride) : nullptr;
  const int atoms = a->C / 64;
  const int smem = 1024 + atoms * 128 * 128 + cb::kScKStages * atoms * 64 * 128 + cb::kScTailBytes;
  dim3 grid(sc_grid(p.nqt, p.nkt_y * p.nkt_x));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == cb::SC_CORR) {
    if (!a->stat_sum || !a->stat_max || !a->lvl[1] || !a->lvl[2] || !a->lvl[3]) return fail("corr_build: missing outputs");
    auto kern = cb::scores_kernel<cb::SC_CORR>;
    static std::atomic<unsigned long long> set{0};
    if (ensure_smem(kern, 200 * 1024, set, "scores")) return -1;
    launch_k(kern, dim3(grid), dim3(cb::kScThreads), smem, st, tq, tk, p);
    return check_launch("corr_build");
  } else {
    if (!a->lse_part || !a->lse2 || !a->stat_max) return fail("attn_lse: missing outputs");
    if (p.mask_radius > 0) {
      auto kern = cb::scores_kernel<cb::SC_LSE_MASKED>;
      static std::atomic<unsigned long long> set{0};
      if (ensure_smem(kern, 200 * 1024, set, "scores")) return -1;
      launch_k(kern, dim3(grid), dim3(cb::kScThreads), smem, st, tq, tk, p);
    } else {
      auto kern = cb::scores_kernel<cb::SC_LSE>;
      static std::atomic<unsigned long long> set{0};
      if (ensure_smem(kern, 200 * 1024, set, "scores")) return -1;
      launch_k(kern, dim3(grid), dim3(cb::kScThreads), smem, st, tq, tk, p);
    }
    if (check_launch("attn_lse")) return -1;
    const int n = a->M * g.Mp;
    launch_k(cb::lse_merge_kernel, dim3((n + 255) / 256), dim3(256), 0, st, p.lse_part, p.nslots, a->M, g.Mp, a->lse2);
    return check_launch("lse_merge");
  }
}

extern "C" {
int craft_corr_build(const craft_scores_args* a, void* stream) { return scores_common(a, cb::SC_CORR, stream); }
int craft_attn_lse(const craft_scores_args* a, void* stream) { return scores_common(a, cb::SC_LSE, stream); }

int craft_corr_stats_finalize(const double* stat_sum, const int* flag, double n, float* mean_rstd, void* stream) {
  launch_k(cb::corr_stats_finalize_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), stat_sum, flag, n, mean_rstd);
  return check_launch("corr_stats_finalize");
}
int craft_clip_gate(const float* stat_max, float attn_clip, float* clip, int* flag, float* diag, void* stream) {
  launch_k(cb::clip_gate_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), stat_max, attn_clip, clip, flag, diag);
  return check_launch("clip_gate");
}

}  // extern "C" (pause)
template <int D, int F, int BK, int KS, int VS, int POLY = 0, bool MASKED = false, bool TRACE = false, bool SPLIT = false,
          bool BULK = false>
static int launch_pv(const craft_pv_args* a, const cb::Grid2& g, cudaStream_t st) {
  using S = cb::PvSmem<D, F, BK, KS, VS, BULK>;
  CUtensorMap tq, tk, tv;
  if (make_map_2d(&tq, a->Q, g.Mp, a->C, a->C, 128)) return -1;
  constexpr int BW = BK / 8;
  const int nbx = (g.W + BW - 1) / BW;
  const int nkt = nbx * ((g.H + 7) / 8);
  if (a->ldv < nkt * BK) return fail("attn_pv: ldv=%d must cover %d blocked keys", a->ldv, nkt * BK);
  if (make_map_3d_keys(&tk, a->K, g, a->C, BW)) return -1;
  if (make_map_2d(&tv, a->Vt, static_cast<long long>(a->M) * F, static_cast<long long>(nkt) * BK, a->ldv, F)) return -1;
  cb::PvParams p;
  memset(&p, 0, sizeof(p));
  const int nqt = (g.Mp + 127) / 128;
  const int need = pv_slots(nqt, a->M, nkt);
  if (a->ksplit < need) return fail("attn_pv: out has %d partial slots, the persistent schedule needs %d", a->ksplit, need);
  if (a->ksplit > 4) return fail("attn_pv: at most 4 partial slots");
  p.g = g; p.M = a->M; p.nslots = a->ksplit; p.zero_fill = a->zero_fill; p.nqt = nqt; p.scale = a->scale; p.w_pos = a->w_pos;
  p.pos_table = a->pos_table; p.R = a->R; p.clip = a->clip; p.lse2 = a->lse2; p.out = a->out;
  p.nkt = nkt; p.nbx = nbx; p.mask_radius = a->mask_radius;
  { const char* e = getenv("CRAFT_PV_NOSTORE"); p.nostore = (e && atoi(e) != 0) ? 1 : 0; }

  auto kern = cb::attn_pv_kernel<D, F, BK, KS, VS, POLY, MASKED, TRACE, SPLIT, BULK>;
  static std::atomic<unsigned long long> set{0};
  if (ensure_smem(kern, S::kTotal, set, "attn_pv")) return -1;
  dim3 grid(pv_grid(nqt, a->M, nkt));
  static long long* d_trace = nullptr;
  const char* trace_path = TRACE ? getenv("CRAFT_PV_TRACE") : nullptr;     // profiling aid: dumps CTA 0's clock64 timeline
  if (trace_path) {
    if (!d_trace) cudaMalloc(&d_trace, (4 * 64 * 8 + 512) * sizeof(long long));
    cudaMemsetAsync(d_trace, 0, (4 * 64 * 8 + 512) * sizeof(long long), st);
    p.trace = d_trace;
  }
  launch_k(kern, dim3(grid), dim3(cb::kPvThreads), S::kTotal, st, tq, tk, tv, p);
  if (trace_path) {
    static long long h[4 * 64 * 8 + 512];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int r = 0; r < 4; ++r)
        for (int t = 0; t < 64; ++t) {
          fprintf(f, "%d %d", r, t);
          for (int k = 0; k < 8; ++k) fprintf(f, " %lld", h[(r * 64 + t) * 8 + k]);
          fprintf(f, "\n");
        }
      for (int c = 0; c < 256; ++c)      // role 9: per-CTA (start, end) in ns
        if (h[2048 + 2 * c]) fprintf(f, "9 %d %lld %lld\n", c, h[2048 + 2 * c], h[2048 + 2 * c + 1]);
      fclose(f);
    }
  }
  return check_launch("attn_pv");
}

extern "C" {
int craft_attn_pv(const craft_pv_args* a, void* stream) {
  if (!a || !a->Q || !a->K || !a->Vt || !a->out || !a->lse2 || !a->clip) return fail("attn_pv: null operand");
  if (a->C != a->M * a->d) return fail("attn_pv: C != M*d");
  if (a->ksplit < 1) return fail("attn_pv: ksplit must be >= 1");
  if (a->pos_table && (2 * a->R + 1) * (2 * a->R + 1) > 225) return fail("attn_pv: pos radius too large");
  cb::Grid2 g = make_grid(a->H, a->W);
  if (a->ldv % 8) return fail("attn_pv: ldv=%d must be a multiple of 8", a->ldv);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->d == 32 && a->F == 128) {
    // fraction (POLY / 8) of the exponentials evaluated on the FMA pipe instead of the MUFU unit
    static int poly = -1;
    if (poly < 0) { const char* e = getenv("CRAFT_PV_POLY"); poly = e ? atoi(e) : 0; }
    switch (poly) {
      case 1: return launch_pv<32, 128, 128, 3, 4, 1>(a, g, st);
      case 2: return launch_pv<32, 128, 128, 3, 4, 2>(a, g, st);
      case 3: return launch_pv<32, 128, 128, 3, 4, 3>(a, g, st);
      case 4: return launch_pv<32, 128, 128, 3, 4, 4>(a, g, st);
      case 8: return launch_pv<32, 128, 128, 3, 4, 8>(a, g, st);
      default: {
        // CRAFT_PV_BULK=1 (experiment, default off): O write-back through a 32 KB staging buffer + cp.async.bulk,
        // paid for by three V stages instead of four.  Measured slower (86.7 vs 83.3 us alone, 236.1 vs 238.9
        // pairs/s: profiles/r02_pv_bulk_writeback.txt) -- the fourth V stage is worth more than the store drain.
        static int bulk = -1;
        if (bulk < 0) { const char* e = getenv("CRAFT_PV_BULK"); bulk = e ? (atoi(e) != 0) : 0; }
        if (getenv("CRAFT_PV_TRACE")) {      // instrumented builds
          if (bulk) return launch_pv<32, 128, 128, 3, 3, 0, false, true, false, true>(a, g, st);
          return launch_pv<32, 128, 128, 3, 4, 0, false, true>(a, g, st);
        }
        if (bulk) return launch_pv<32, 128, 128, 3, 3, 0, false, false, false, true>(a, g, st);
        return launch_pv<32, 128, 128, 3, 4>(a, g, st);
      }
    }
  }
  if (a->mask_radius > 0) {
    if (a->d == 64 && a->F == 256) return launch_pv<64, 256, 64, 4, 4, 0, true, false, true>(a, g, st);
    return fail("attn_pv: the --f2radius key mask is built for the F2 transformer shape (d=64, F=256) only");
  }
  // 64-key tiles: the split TMEM layout (attn_pv.cuh; F2 transformer 118 -> 104 us); 128-key tiles: the ring
  if (a->d == 64 && a->F == 256) return launch_pv<64, 256, 64, 4, 4, 0, false, false, true>(a, g, st);
  if (a->d == 128 && a->F == 128) return launch_pv<128, 128, 64, 3, 4, 0, false, false, true>(a, g, st);
  if (a->d == 64 && a->F == 128) return launch_pv<64, 128, 128, 3, 4>(a, g, st);
  return fail("attn_pv: unsupported (d=%d, F=%d)", a->d, a->F);
}

int craft_modes_finalize(const float* O, int nsum, int M, int F, const float* w_score, const float* b_score,
                         const float* coeff, int gma, const void* x_bf16, int ldx, int colx,
                         const float* x_f32, int ldxf, int colxf, int H, int W, void* out_bf16, int ldb,
                         int colb, float* out_f32, int ldf, int colf, int pv_bk, void* stream) {
  cb::Grid2 g = make_grid(H, W);
  if (M < 1 || M > 4) return fail("modes_finalize: M out of range");
  int pv_G = 0, pv_nkt = 0;
  if (pv_bk) {     // O comes from craft_attn_pv's persistent schedule with key blocks of pv_bk tokens
    if (pv_bk != 64 && pv_bk != 128) return fail("modes_finalize: pv_bk must be 0, 64 or 128");
    pv_nkt = ((H + 7) / 8) * ((W + pv_bk / 8 - 1) / (pv_bk / 8));
    pv_G = pv_grid((g.Mp + 127) / 128, M, pv_nkt);
    if (static_cast<long long>((g.Mp + 127) / 128) * M * pv_nkt * pv_G >= (1ll << 31)) return fail("modes_finalize: schedule too large");
  }
  if (nsum < 1 || nsum > 4) return fail("modes_finalize: nsum must be in [1,4]");
  if (!x_bf16 && !x_f32) return fail("modes_finalize: need the skip input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long mode_stride = static_cast<long long>(g.Mp) * F;
  const long long part_stride = mode_stride * M;
  dim3 grid((g.Mp + 7) / 8);
  auto* xb = static_cast<const cb::act_t*>(x_bf16);
  auto* ob = static_cast<cb::act_t*>(out_bf16);
  if (F == 128)
    launch_k(cb::modes_finalize_kernel<128>, dim3(grid), dim3(256), 0, st, O, nsum, part_stride, M, mode_stride, w_score, b_score, coeff, gma, xb, ldx, colx, x_f32, ldxf, colxf, g, ob, ldb, colb, out_f32, ldf, colf, pv_G, pv_nkt);
  else if (F == 256)
    launch_k(cb::modes_finalize_kernel<256>, dim3(grid), dim3(256), 0, st, O, nsum, part_stride, M, mode_stride, w_score, b_score, coeff, gma, xb, ldx, colx, x_f32, ldxf, colxf, g, ob, ldb, colb, out_f32, ldf, colf, pv_G, pv_nkt);
  else return fail("modes_finalize: F must be 128 or 256");
  return check_launch("modes_finalize");
}

int craft_soft_aggregate(const float* x, const float* basis, int M, long long n, int F, const float* w,
                         const float* b, float* out, void* stream) {
  if (!x || !w || !b || !out) return fail("soft_aggregate: null operand");
  if (M < 1 || M > 8 || n < 1 || F < 1) return fail("soft_aggregate: bad shape M=%d n=%lld F=%d", M, n, F);
  if (!basis) basis = x;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (F == 1) {
    long long blocks = (n + 255) / 256;
    if (blocks > 8ll * sm_count()) blocks = 8ll * sm_count();
    launch_k(cb::soft_aggregate_scalar_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, x, basis, M, n, w, b, out);
  } else {
    launch_k(cb::soft_aggregate_feat_kernel, dim3(static_cast<unsigned>((n + 7) / 8)), dim3(256), 0, st, x, basis, M, n, F, w, b, out);
  }
  return check_launch("soft_aggregate");
}

int craft_attn_dense(const craft_dense_attn_args* a, void* stream) {
  if (!a || !a->Q || !a->K || !a->clip || !a->out) return fail("attn_dense: null operand");
  if (a->C != a->M * a->d || a->d % 2) return fail("attn_dense: bad C/M/d");
  if (a->pos_table && (2 * a->R + 1) * (2 * a->R + 1) > 225) return fail("attn_dense: pos radius too large");
  cb::Grid2 g = make_grid(a->H, a->W);
  cb::DenseAttnParams p;
  memset(&p, 0, sizeof(p));
  p.Q = static_cast<const cb::act_t*>(a->Q); p.K = static_cast<const cb::act_t*>(a->K);
  p.C = a->C; p.M = a->M; p.d = a->d; p.scale = a->scale; p.w_pos = a->w_pos; p.pos_table = a->pos_table; p.R = a->R;
  p.clip = a->clip; p.lse2 = a->lse2; p.mask_radius = a->mask_radius; p.out = a->out;
  const long long units = static_cast<long long>(a->M) * a->H * a->W;
  launch_k(cb::attn_dense_kernel, dim3(static_cast<unsigned>((units + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), p, g);
  return check_launch("attn_dense");
}

int craft_corr_lookup(const float* const* lvl, const void* lvl0_h16, int H, int W, const float* coords,
                      const float* mean_rstd, void* out_bf16, int ldb, float* out_nchw, int first_level, void* stream) {
  cb::Grid2 g = make_grid(H, W);
  cb::LookupParams p;
  memset(&p, 0, sizeof(p));
  int h = H, w = W;
  for (int l = 0; l < 4; ++l) {
    p.lvl[l] = lvl[l]; p.hl[l] = h; p.wl[l] = w; p.qstride[l] = static_cast<long long>(h) * w;
    if (l >= first_level && !lvl[l] && !(l == 0 && lvl0_h16)) return fail("corr_lookup: level %d missing", l);
    h /= 2; w /= 2;
  }
  p.coords = coords; p.stats = mean_rstd; p.out_b = static_cast<cb::act_t*>(out_bf16); p.ldb = ldb;
  p.out_nchw = out_nchw; p.first_level = first_level;
  p.lvl0h = static_cast<const __half*>(lvl0_h16); p.nbx0 = (W + 7) / 8;
  p.qstride0h = static_cast<long long>((H + 7) / 8) * p.nbx0 * 64;
  p.lvl0_base = p.lvl0h ? reinterpret_cast<const float*>(p.lvl0h + static_cast<long long>(g.Mp) * p.qstride0h) : nullptr;
  launch_k(cb::corr_lookup_kernel, dim3((g.Mp + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), p, g);
  return check_launch("corr_lookup");
}

int craft_corr_lookup0(const void* Q, const void* K, int M, int d, float scale, float w_agg, float w_pos,
                       const float* pos_table, int R, const float* clip, int H, int W, const float* coords,
                       const float* mean_rstd, void* out_bf16, int ldb, float* out_nchw, void* stream) {
  if (M * d != 256 || (d != 64 && d != 128 && d != 256)) return fail("corr_lookup0: needs M*d == 256, d in {64,128,256} (got %d x %d)", M, d);
  if (!Q || !K || !clip || !coords || !mean_rstd) return fail("corr_lookup0: null operand");
  if (pos_table && (2 * R + 1) * (2 * R + 1) > 225) return fail("corr_lookup0: pos radius too large");
  cb::Grid2 g = make_grid(H, W);
  cb::Lookup0Params p;
  memset(&p, 0, sizeof(p));
  p.Q = static_cast<const cb::act_t*>(Q); p.K = static_cast<const cb::act_t*>(K);
  p.M = M; p.d = d; p.scale = scale; p.w_agg = w_agg; p.w_pos = w_pos; p.pos_table = pos_table; p.Rb = R;
  p.clip = clip; p.coords = coords; p.stats = mean_rstd; p.out_b = static_cast<cb::act_t*>(out_bf16);
  p.ldb = ldb; p.out_nchw = out_nchw;
  static std::atomic<unsigned long long> set{0};
  // CRAFT_LOOKUP0_CAP: cells of shared memory for staging a 4x2 query patch's bounding box.  Default 0 = every
  // row straight from L2: measured (profiles/r02_lookup0_cap_sweep.txt) 45.6 us, against 55 us with 144-192 staged
  // cells (2 CTAs/SM) and 87 us with >= 256 (1 CTA/SM) -- the kernel is latency/issue bound, not L2 bound, and
  // the staged variant pays for its shared memory with occupancy.
  static int cap = -1;
  if (cap < 0) { const char* e = getenv("CRAFT_LOOKUP0_CAP"); cap = e ? atoi(e) : 0; if (cap > 400) cap = 400; if (cap < 0) cap = 0; }
  const int smem = cap * 512;
  p.cap = cap;
  if (ensure_smem(cb::corr_lookup0_kernel, 400 * 512, set, "corr_lookup0")) return -1;
  const int patches = ((W + cb::kL0PX - 1) / cb::kL0PX) * ((H + cb::kL0PY - 1) / cb::kL0PY);
  launch_k(cb::corr_lookup0_kernel, dim3(patches), dim3(256), smem, static_cast<cudaStream_t>(stream), p, g);
  return check_launch("corr_lookup0");
}

int craft_convf1(const float* flow, const float* wt, const float* bias, int H, int W, void* out_bf16, int ldo,
                 int colo, void* stream) {
  cb::Grid2 g = make_grid(H, W);
  dim3 grid(H * ((W + 15) / 16), 2);
  launch_k(cb::convf1_kernel, dim3(grid), dim3(128), 0, static_cast<cudaStream_t>(stream), flow, wt, bias, g, static_cast<cb::act_t*>(out_bf16), ldo, colo);
  return check_launch("convf1");
}

int craft_flow_update(float* coords1, float* flow, const float* delta, int ldd, int H, int W, void* stream) {
  cb::Grid2 g = make_grid(H, W);
  launch_k(cb::flow_update_kernel, dim3((g.Mp + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), coords1, flow, delta, ldd, g);
  return check_launch("flow_update");
}
int craft_init_coords(float* coords1, const float* flow_init, int H, int W, void* stream) {
  cb::Grid2 g = make_grid(H, W);
  launch_k(cb::init_coords_kernel, dim3((g.Mp + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), coords1, flow_init, g);
  return check_launch("init_coords");
}

int craft_upsample_flow(const void* mask, int mask_is_bf16, int ldm, const float* flow, int H, int W,
                        float* out, void* stream) {
  cb::Grid2 g = make_grid(H, W);
  dim3 grid((H * W + 3) / 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mask_is_bf16)
    launch_k(cb::upsample_flow_kernel<cb::act_t>, dim3(grid), dim3(256), 0, st, static_cast<const cb::act_t*>(mask), ldm, flow, g, out);
  else
    launch_k(cb::upsample_flow_kernel<float>, dim3(grid), dim3(256), 0, st, static_cast<const float*>(mask), ldm, flow, g, out);
  return check_launch("upsample_flow");
}


int craft_forward_interpolate(const float* flow, int H, int W, float* out, void* stream) {
  if (!flow || !out || H < 1 || W < 1) return fail("forward_interpolate: bad arguments");
  launch_k(cb::forward_interpolate_kernel, dim3((H * W + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), flow, H, W, out);
  return check_launch("forward_interpolate");
}

int craft_flow_encode(const float* flow, int H, int W, int mode, void* out, void* stream) {
  if (!flow || !out || H < 1 || W < 1 || (mode != 0 && mode != 1)) return fail("flow_encode: bad arguments");
  launch_k(cb::flow_encode_kernel, dim3((H * W + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), flow, H, W, mode, out);
  return check_launch("flow_encode");
}

int craft_image_s2d(const float* img, int N, int H, int W, void* out, int out_is_half, void* stream) {
  if (!img || !out) return fail("image_s2d: null operand");
  if (N < 1 || H < 2 || W < 2 || H % 2 || W % 2) return fail("image_s2d: H, W must be even (got %d x %d)", H, W);
  const long long cells = static_cast<long long>(N) * (H / 2 + 3) * (W / 2 + 3);
  const dim3 grid(static_cast<unsigned>((cells + 255) / 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_is_half) launch_k(cb::image_s2d_kernel<__half>, grid, dim3(256), 0, st, img, N, H, W, static_cast<__half*>(out));
  else launch_k(cb::image_s2d_kernel<float>, grid, dim3(256), 0, st, img, N, H, W, static_cast<float*>(out));
  return check_launch("image_s2d");
}

int craft_nhwc_instnorm_stats(const void* x, int is_half, int N, int HW, int C, float eps, float* part,
                              long long part_capacity, float* ab, void* stream) {
  const int V = is_half ? 8 : 4;
  if (C % V || C > 256 * V || C < V) return fail("instnorm_stats: C=%d must be a multiple of %d (<= %d)", C, V, 256 * V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int chunks = (4 * sm_count() + N - 1) / N;          // ~4 blocks per SM in total
  int rows = (HW + chunks - 1) / chunks;
  if (rows < 64) rows = 64;
  chunks = (HW + rows - 1) / rows;
  if (static_cast<long long>(chunks) * N * C * 2 > part_capacity)
    return fail("instnorm_stats: partial buffer holds %lld floats, needs %lld", part_capacity, static_cast<long long>(chunks) * N * C * 2);
  if (is_half) launch_k(cb::nhwc_stats_kernel<__half>, dim3(chunks, N), dim3(256), 0, st, static_cast<const __half*>(x), HW, C, rows, part);
  else launch_k(cb::nhwc_stats_kernel<float>, dim3(chunks, N), dim3(256), 0, st, static_cast<const float*>(x), HW, C, rows, part);
  if (check_launch("nhwc_stats")) return -1;
  launch_k(cb::instnorm_finalize_kernel, dim3((N * C + 7) / 8), dim3(256), 0, st, part, chunks, N * C, 1.0f / static_cast<float>(HW), eps, ab);
  return check_launch("instnorm_finalize");
}

}  // extern "C" (pause)
template <typename T>
static int launch_instnorm_fused(const void* x, int N, int HW, int C, float eps, const void* res, const float* rab,
                                 int rab_nstride, int relu_in, int relu_out, float* part, long long part_capacity,
                                 float* ab_out, void* out, cudaStream_t st) {
  constexpr int V = cb::ActVec<T>::N;
  const int cq = C / V;
  cb::InFusedParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.HW = HW; p.C = C;
  p.cqp = 1;
  while (p.cqp < cq) p.cqp <<= 1;
  const int sms = sm_count();
  p.cpi = sms / N;
  const int rowbytes = C * static_cast<int>(sizeof(T));
  // small tensors: at least 32 rows per CTA, so that a CTA's 16 warps have something to add up
  if (p.cpi > (HW + 31) / 32) p.cpi = (HW + 31) / 32;
  if (p.cpi < 1) p.cpi = 1;
  p.rows_per_cta = (HW + p.cpi - 1) / p.cpi;
  const int overhead = (cb::kInThreads / 32 + 1) * 2 * C * 4 + cb::kInMaxPieces * 8 + 128;
  const int cap_rows = std::min((227 * 1024 - overhead) / rowbytes, cb::kInMaxPieces * std::max(1, cb::kInPieceBytes / rowbytes));
  p.smem_rows = std::min(p.rows_per_cta, cap_rows);
  p.inv_hw = 1.0f / static_cast<float>(HW); p.eps = eps;
  p.relu_in = relu_in; p.relu_out = relu_out; p.rab_nstride = rab_nstride;
  const int grid = p.cpi * N;
  // part: [0, 64) floats = the manual barrier's state (zero at first use), partial sums from float 64 on
  if (static_cast<long long>(grid) * 2 * C + 64 > part_capacity)
    return fail("instnorm_fused: partial buffer holds %lld floats, needs %lld", part_capacity, static_cast<long long>(grid) * 2 * C + 64);
  static int coop = -1;
  if (coop < 0) { const char* e = getenv("CRAFT_B200_IN_COOP"); coop = e ? (atoi(e) != 0) : 1; }
  p.coop = coop;
  const int smem = p.smem_rows * rowbytes + overhead;
  auto kern = cb::instnorm_fused_kernel<T>;
  static std::atomic<unsigned long long> set{0};
  if (ensure_smem(kern, 227 * 1024, set, "instnorm_fused")) return -1;
  // a cooperative launch: the grid barrier needs every CTA resident, and the driver -- not a spin loop of ours --
  // is what guarantees it when other streams hold SMs (the context encoder runs next to this one)
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(cb::kInThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (coop) {
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.numAttrs = 1;
  } else {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
  }
  cfg.attrs = attr;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<const T*>(x), static_cast<const T*>(res), rab, part + 64,
                                     reinterpret_cast<unsigned*>(part), ab_out, static_cast<T*>(out), p);
  if (e != cudaSuccess) return fail("instnorm_fused: launch failed: %s", cudaGetErrorString(e));
  return check_launch("instnorm_fused");
}

extern "C" {
int craft_nhwc_instnorm_apply(const void* x, int is_half, int N, int HW, int C, float eps, const void* res,
                              const float* rab, int rab_nstride, int relu_in, int relu_out, float* part,
                              long long part_capacity, float* ab_out, void* out, void* stream) {
  if (!x || !out || !part) return fail("instnorm_apply: null operand");
  const int V = is_half ? 8 : 4;
  if (C % V || C < V || C > 256 || C / V > 32) return fail("instnorm_apply: C=%d must be a multiple of %d, at most %d", C, V, std::min(256, 32 * V));
  if (N < 1 || N > sm_count()) return fail("instnorm_apply: N=%d images (at most one per SM)", N);
  if (HW < 1) return fail("instnorm_apply: empty image");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_half)
    return launch_instnorm_fused<__half>(x, N, HW, C, eps, res, rab, rab_nstride, relu_in, relu_out, part, part_capacity, ab_out, out, st);
  return launch_instnorm_fused<float>(x, N, HW, C, eps, res, rab, rab_nstride, relu_in, relu_out, part, part_capacity, ab_out, out, st);
}

int craft_conv3x3_c64(const void* x, const void* w, const float* bias, int relu, int N, int H, int W, void* out,
                      float* part, long long part_capacity, float* ab, float eps, void* stream) {
  if (!x || !w || !out) return fail("conv3x3_c64: null operand");
  if (N < 1 || H < 1 || W < 1) return fail("conv3x3_c64: empty input");
  if ((ab != nullptr) != (part != nullptr)) return fail("conv3x3_c64: statistics need both `part` and `ab`");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cb::ConvEncParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.Wp = W + 2;
  const long long rpi = static_cast<long long>(H + 1) * p.Wp;
  if (rpi * N > 0x7fffffffLL - 4096) return fail("conv3x3_c64: %lld rows do not fit 32-bit row coordinates", rpi * N);
  p.rpi = static_cast<int>(rpi);
  p.tiles_per_image = (p.rpi + 127) / 128;
  p.ctas_per_image = std::max(1, std::min(p.tiles_per_image, sm_count() / N));
  p.bias = bias; p.relu = relu; p.out = static_cast<__half*>(out); p.part = part;
  if (part && static_cast<long long>(p.ctas_per_image) * N * 64 * 2 > part_capacity)
    return fail("conv3x3_c64: partial buffer holds %lld floats, needs %lld", part_capacity,
                static_cast<long long>(p.ctas_per_image) * N * 64 * 2);
  CUtensorMap tx, tw;
  if (make_map_2d(&tx, x, rpi * N, 64, 64, 136)) return -1;
  if (make_map_2d(&tw, w, 9 * 64, 64, 64, 64)) return -1;
  const dim3 grid(p.ctas_per_image * N);
  if (part) {
    auto kern = cb::conv3x3_c64_kernel<true>;
    static std::atomic<unsigned long long> set{0};
    if (ensure_smem(kern, cb::kCvSmem, set, "conv3x3_c64")) return -1;
    launch_k(kern, grid, dim3(cb::kCvThreads), cb::kCvSmem, st, tx, tw, p);
    if (check_launch("conv3x3_c64")) return -1;
    // InstanceNorm2d scale / shift from the per-CTA partial sums (fixed order)
    launch_k(cb::instnorm_finalize_kernel, dim3((N * 64 + 7) / 8), dim3(256), 0, st, part, p.ctas_per_image, N * 64,
             1.0f / (static_cast<float>(H) * static_cast<float>(W)), eps, ab);
    return check_launch("instnorm_finalize");
  }
  auto kern = cb::conv3x3_c64_kernel<false>;
  static std::atomic<unsigned long long> set{0};
  if (ensure_smem(kern, cb::kCvSmem, set, "conv3x3_c64")) return -1;
  launch_k(kern, grid, dim3(cb::kCvThreads), cb::kCvSmem, st, tx, tw, p);
  return check_launch("conv3x3_c64");
}

int craft_nhwc_affine_pad(const void* v, int is_half, int v_pad, const float* ab, int ab_nstride, const void* res, int res_pad,
                          const float* rab, int rab_nstride, int relu_in, int relu_out, int N, int H, int W, int C, void* out,
                          int out_pad, void* stream) {
  const int V = is_half ? 8 : 4;
  if (!v || !out) return fail("nhwc_affine_pad: null operand");
  if (C % V) return fail("nhwc_affine_pad: C must be a multiple of %d", V);
  const long long rows = out_pad ? static_cast<long long>(N) * (H + 1) * (W + 2) : static_cast<long long>(N) * H * W;
  const long long totalv = rows * (C / V);
  const dim3 grid(static_cast<unsigned>((totalv + 255) / 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_half)
    launch_k(cb::nhwc_affine_pad_kernel<__half>, grid, dim3(256), 0, st, static_cast<const __half*>(v), v_pad, ab, ab_nstride,
             static_cast<const __half*>(res), res_pad, rab, rab_nstride, relu_in, relu_out, N, H, W, C, static_cast<__half*>(out), out_pad);
  else
    launch_k(cb::nhwc_affine_pad_kernel<float>, grid, dim3(256), 0, st, static_cast<const float*>(v), v_pad, ab, ab_nstride,
             static_cast<const float*>(res), res_pad, rab, rab_nstride, relu_in, relu_out, N, H, W, C, static_cast<float*>(out), out_pad);
  return check_launch("nhwc_affine_pad");
}

int craft_nhwc_affine(const void* v, int is_half, const float* ab, int ab_nstride, const void* res, const float* rab,
                      int rab_nstride, int relu_in, int relu_out, int N, int HW, int C, void* out, void* stream) {
  const int V = is_half ? 8 : 4;
  if (C % V) return fail("nhwc_affine: C must be a multiple of %d", V);
  const long long per_image = static_cast<long long>(HW) * C;
  const long long totalv = per_image * N / V;
  const dim3 grid(static_cast<unsigned>((totalv + 255) / 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_half)
    launch_k(cb::nhwc_affine_kernel<__half>, grid, dim3(256), 0, st, static_cast<const __half*>(v), ab, ab_nstride,
             static_cast<const __half*>(res), rab, rab_nstride, relu_in, relu_out, per_image, C, totalv, static_cast<__half*>(out));
  else
    launch_k(cb::nhwc_affine_kernel<float>, grid, dim3(256), 0, st, static_cast<const float*>(v), ab, ab_nstride,
             static_cast<const float*>(res), rab, rab_nstride, relu_in, relu_out, per_image, C, totalv, static_cast<float*>(out));
  return check_launch("nhwc_affine");
}

}  // extern "C"
