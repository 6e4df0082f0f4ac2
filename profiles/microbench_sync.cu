// Micro-benchmark: round-trip latencies of the handshakes attn_pv / scores / gemm are built from.
//   A: warp <-> warp through two mbarriers (mbarrier.arrive + try_wait)
//   B: warp -> MMA warp -> tcgen05.commit -> warp   (commit with no MMA outstanding)
//   C: 8 warps arrive (count 8) -> MMA warp -> commit -> 8 warps   (the attn_pv skeleton)
//   D: like C but the MMA warp answers with a plain mbarrier.arrive
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I craft_b200/csrc -o build/mb_sync profiles/microbench_sync.cu
#include <cstdio>
#include "common.cuh"
using namespace cb;

template <int TEST, int HINT>
__global__ void __launch_bounds__(320, 1) k(long long* out, int iters) {
  __shared__ uint64_t barA, barB;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const int nsm = (TEST == 0) ? 1 : 8;      // warps on the "softmax" side
  if (threadIdx.x == 0) {
    mbar_init(&barA, nsm);
    mbar_init(&barB, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  auto wait = [&](uint64_t* b, uint32_t ph) {
    if (HINT) mbar_wait(b, ph);
    else { while (!mbar_try_wait_nohint(b, ph)) {} }
  };
  long long t0 = clock64();
  if (warp == 1) {
    for (int i = 0; i < iters; ++i) {
      wait(&barA, i & 1);
      tc_fence_after();
      if (TEST == 0 || TEST == 3) {
        if (elect_one()) mbar_arrive(&barB);
      } else {
        if (elect_one()) umma_commit(&barB);
      }
      __syncwarp();
    }
  } else if (warp >= 2 && warp < 2 + nsm) {
    for (int i = 0; i < iters; ++i) {
      mbar_arrive_warp(&barA);
      wait(&barB, i & 1);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 64) out[blockIdx.x] = (t1 - t0) / iters;
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_slot, 32);
}

template <int TEST, int HINT>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 8 * 148);
  k<TEST, HINT><<<148, 320>>>(d, 2000);
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-58s %6lld clk / round trip   (%s)\n", name, h[0], cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<0, 1>("A  warp<->warp, mbarrier.arrive both ways, hint wait");
  run<0, 0>("A  warp<->warp, mbarrier.arrive both ways, plain try_wait");
  run<1, 1>("C  8 warps -> MMA warp -> tcgen05.commit -> 8 warps, hint");
  run<1, 0>("C  8 warps -> MMA warp -> tcgen05.commit -> 8 warps, plain");
  run<3, 1>("D  8 warps -> MMA warp -> mbarrier.arrive -> 8 warps, hint");
  run<3, 0>("D  8 warps -> MMA warp -> mbarrier.arrive -> 8 warps, plain");
  return 0;
}
