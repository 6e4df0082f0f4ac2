// Micro-benchmark: TMA load latency and per-SM throughput as a function of the bytes kept in flight
// (first results: profiles/r01_mb_tma.txt; the boxes-per-barrier rows were added afterwards and have
// not been run yet).
//
// Question it answers: the shift-GEMM main loop costs ~215 ns per
// 64 K columns (32 KB of A+B per CTA) with no unit above 25 %.  If that is Little's law --
// (bytes in flight) / (TMA round trip) -- then the curve below saturates early and only fewer bytes
// per FLOP help; if the per-SM rate keeps growing with the number of boxes in flight, deeper or
// finer-grained pipelines help.
//
// One CTA per SM (or a single CTA for the unloaded latency); one thread issues 2-D TMA boxes
// [ROWS x 64 bf16] (128-byte swizzle, the GEMM's A-tile shape) from an L2-resident matrix into a ring
// of NS smem slots and consumes them in order through mbarriers, exactly like the GEMM producer/MMA
// pair but with no math.  Reported: ns per box and bytes/clk/SM.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I craft_b200/csrc -o build/mb_tma profiles/microbench_tma.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "common.cuh"
using namespace cb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NS, int ROWS, int BPB>
__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap tm, int iters, int rows_total,
                                           long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int kBox = ROWS * 128;                 // one TMA box
  constexpr int kSlot = BPB * kBox;                // BPB boxes share one barrier (like the GEMM's A+B atoms)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * kSlot);
  uint64_t* empty = full + NS;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm);
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (elect_one()) {               // producer
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], kSlot);
        // different rows / column chunks per step and per CTA, all inside an L2-resident matrix
        const int row = (blockIdx.x * 128 + it * 7) % (rows_total - ROWS);
#pragma unroll
        for (int b = 0; b < BPB; ++b)
          tma_load_2d(smem + s * kSlot + b * kBox, &tm, &full[s], ((it + b) % 8) * 64, row);
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    if (elect_one()) {               // consumer: takes the boxes in order, frees the slot at once
      int s = 0; uint32_t ph = 0;
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full[s], ph);
        mbar_arrive(&empty[s]);
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  }
}

template <int NS, int ROWS, int BPB = 1>
void run(const CUtensorMap& tm, int grid, int rows_total, long long* d_out) {
  constexpr int kBox = ROWS * 128;
  constexpr int kSlot = BPB * kBox;
  const int smem = NS * kSlot + 1024 + 256;
  if (smem > 227 * 1024) return;
  auto kern = k<NS, ROWS, BPB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  kern<<<grid, 64, smem>>>(tm, 200, rows_total, d_out);
  kern<<<grid, 64, smem>>>(tm, iters, rows_total, d_out);
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double clk_per_slot = static_cast<double>(mx) / iters;
  printf("grid=%3d  box=%2d KB x%d per barrier  slots in flight=%2d (%3d KB)  %7.0f clk/barrier  %6.1f B/clk/SM   (%s)\n", grid,
         kBox / 1024, BPB, NS, NS * kSlot / 1024, clk_per_slot, kSlot / clk_per_slot, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int rows = 7280, cols = 640;                      // the update block's X matrix
  void* d = nullptr;
  cudaMalloc(&d, static_cast<size_t>(rows) * cols * 2);
  cudaMemset(d, 0, static_cast<size_t>(rows) * cols * 2);
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 148);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  auto make = [&](int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return m;
  };
  const CUtensorMap m128 = make(128), m256 = make(256), m64 = make(64);
  for (int grid : {1, 148}) {
    run<1, 128>(m128, grid, rows, d_out);
    run<2, 128>(m128, grid, rows, d_out);
    run<3, 128>(m128, grid, rows, d_out);
    run<4, 128>(m128, grid, rows, d_out);
    run<6, 128>(m128, grid, rows, d_out);
    run<8, 128>(m128, grid, rows, d_out);
    run<12, 128>(m128, grid, rows, d_out);
    run<2, 256>(m256, grid, rows, d_out);
    run<4, 256>(m256, grid, rows, d_out);
    run<6, 256>(m256, grid, rows, d_out);
    run<4, 64>(m64, grid, rows, d_out);
    run<8, 64>(m64, grid, rows, d_out);
    run<16, 64>(m64, grid, rows, d_out);
    // several boxes per barrier: separates the cost of a TMA instruction from the cost of a barrier round
    run<3, 128, 2>(m128, grid, rows, d_out);
    run<3, 128, 4>(m128, grid, rows, d_out);
    run<6, 128, 2>(m128, grid, rows, d_out);
    run<6, 64, 4>(m64, grid, rows, d_out);
  }
  return 0;
}
