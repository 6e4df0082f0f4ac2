// Micro-benchmark: per-SM throughput of the candidate exp2 implementations for the softmax warps of
// attn_pv (ex2.approx.f32 on MUFU, ex2.approx.f16x2 on MUFU, a degree-3 polynomial on the FMA pipe,
// and mixes).  Build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb_exp microbench_exp.cu
// Prints elements/clk/SM for 16 resident warps per SM (the attn_pv softmax population).
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t pack_b2(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// 2^x for x <= 0 on the FMA/ALU pipes: round-to-nearest split x = n + f, f in [-0.5, 0.5],
// p(f) ~ 2^f (degree 3, rel. err ~1e-4), result = p * 2^n through the exponent field.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float magic = 12582912.0f;              // 1.5 * 2^23
  const float t = x + magic;
  const float n = t - magic;
  const float f = x - n;
  float p = fmaf(f, 0.05550410866f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  const int bits = __float_as_int(p) + (__float_as_int(t) << 23);
  return __int_as_float(bits);
}

// packed (f32x2) flavour of the same polynomial: two elements per FMA-pipe instruction
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  return (static_cast<uint64_t>(__float_as_uint(b)) << 32) | __float_as_uint(a);
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& y0, float& y1) {
  x0 = fmaxf(x0, -125.0f); x1 = fmaxf(x1, -125.0f);
  const uint64_t magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f);
  const uint64_t x = pk2(x0, x1);
  const uint64_t t = add2(x, magic);
  const uint64_t n = add2(t, nmagic);
  const uint64_t f = fma2(n, pk2(-1.0f, -1.0f), x);
  uint64_t p = fma2(f, pk2(0.05550410866f, 0.05550410866f), pk2(0.2402265070f, 0.2402265070f));
  p = fma2(p, f, pk2(0.6931471806f, 0.6931471806f));
  p = fma2(p, f, pk2(1.0f, 1.0f));
  const uint32_t plo = static_cast<uint32_t>(p), phi = static_cast<uint32_t>(p >> 32);
  const uint32_t tlo = static_cast<uint32_t>(t), thi = static_cast<uint32_t>(t >> 32);
  y0 = __uint_as_float(plo + (tlo << 23));
  y1 = __uint_as_float(phi + (thi << 23));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, const float* in, int iters, float sc, float lse) {
  float acc[32];
  uint32_t pk[16];
#pragma unroll
  for (int e = 0; e < 32; ++e) acc[e] = in[(threadIdx.x * 32 + e) & 1023];
#pragma unroll
  for (int e = 0; e < 16; ++e) pk[e] = 0;
  for (int it = 0; it < iters; ++it) {
    float x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = fmaf(acc[e], sc, -lse);
    if (MODE == 0) {          // f32 MUFU + bf16 pack (today's path)
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] ^= pack_b2(ex2f(x[2 * e]), ex2f(x[2 * e + 1]));
    } else if (MODE == 1) {   // pack to f16x2 first, one MUFU op per pair
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] ^= ex2h2(pack_h2(x[2 * e], x[2 * e + 1]));
    } else if (MODE == 2) {   // polynomial only
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] ^= pack_b2(ex2_poly(x[2 * e]), ex2_poly(x[2 * e + 1]));
    } else if (MODE == 3) {   // 3/4 MUFU f32, 1/4 polynomial
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if ((e & 3) == 3) pk[e] ^= pack_b2(ex2_poly(x[2 * e]), ex2_poly(x[2 * e + 1]));
        else pk[e] ^= pack_b2(ex2f(x[2 * e]), ex2f(x[2 * e + 1]));
      }
    } else if (MODE == 4) {   // 1/2 MUFU f32, 1/2 polynomial
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (e & 1) pk[e] ^= pack_b2(ex2_poly(x[2 * e]), ex2_poly(x[2 * e + 1]));
        else pk[e] ^= pack_b2(ex2f(x[2 * e]), ex2f(x[2 * e + 1]));
      }
    } else if (MODE == 5) {   // bf16x2 MUFU
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] ^= ex2b2(pack_b2(x[2 * e], x[2 * e + 1]));
    } else if (MODE == 8) {   // packed polynomial only
#pragma unroll
      for (int e = 0; e < 16; ++e) { float a, b; ex2_poly2(x[2 * e], x[2 * e + 1], a, b); pk[e] ^= pack_b2(a, b); }
    } else if (MODE == 9) {   // 3/4 MUFU, 1/4 packed polynomial
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if ((e & 3) == 3) { float a, b; ex2_poly2(x[2 * e], x[2 * e + 1], a, b); pk[e] ^= pack_b2(a, b); }
        else pk[e] ^= pack_b2(ex2f(x[2 * e]), ex2f(x[2 * e + 1]));
      }
    } else if (MODE == 10) {  // 5/8 MUFU, 3/8 packed polynomial
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if ((e & 7) < 3) { float a, b; ex2_poly2(x[2 * e], x[2 * e + 1], a, b); pk[e] ^= pack_b2(a, b); }
        else pk[e] ^= pack_b2(ex2f(x[2 * e]), ex2f(x[2 * e + 1]));
      }
    } else if (MODE == 11) {  // 1/2 MUFU, 1/2 packed polynomial
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (e & 1) { float a, b; ex2_poly2(x[2 * e], x[2 * e + 1], a, b); pk[e] ^= pack_b2(a, b); }
        else pk[e] ^= pack_b2(ex2f(x[2 * e]), ex2f(x[2 * e + 1]));
      }
    } else if (MODE == 7) {   // 5/8 MUFU f32, 3/8 polynomial
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if ((e & 7) < 3) pk[e] ^= pack_b2(ex2_poly(x[2 * e]), ex2_poly(x[2 * e + 1]));
        else pk[e] ^= pack_b2(ex2f(x[2 * e]), ex2f(x[2 * e + 1]));
      }
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] += 0.001f * (float)(pk[e & 15] & 1u);
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) s += acc[e];
#pragma unroll
  for (int e = 0; e < 16; ++e) s += (float)pk[e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void acc_check(float* out) {
  // max relative error of the polynomial and of the f16x2 path against exp2f over [-24, 0]
  float worst_p = 0.f, worst_h = 0.f;
  for (int i = threadIdx.x; i < (1 << 20); i += blockDim.x) {
    const float x = -24.0f * (float)i / (float)(1 << 20);
    const float r = exp2f(x);
    const float p = ex2_poly(x);
    uint32_t hh = ex2h2(pack_h2(x, x));
    const float h = __half2float(*reinterpret_cast<__half*>(&hh));
    worst_p = fmaxf(worst_p, fabsf(p - r) / r);
    if (x > -13.f) worst_h = fmaxf(worst_h, fabsf(h - r) / r);
  }
  out[threadIdx.x * 2] = worst_p;
  out[threadIdx.x * 2 + 1] = worst_h;
}

template <int MODE>
void run(const char* name, float* out, float* in) {
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 512>>>(out, in, 10, 0.1f, 3.0f);
  cudaEventRecord(e0);
  k<MODE><<<148, 512>>>(out, in, iters, 0.1f, 3.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double elems = 512.0 * 32 * iters;              // per SM
  const double clks = ms * 1e-3 * clk * 1e3;
  printf("%-34s %8.3f ms  %6.2f elem/clk/SM (at %d MHz nominal)  -> 16384-elem tile = %6.0f clk\n", name, ms,
         elems / clks, clk / 1000, 16384.0 / (elems / clks));
}

int main() {
  float *out, *in;
  cudaMalloc(&out, 148 * 512 * 4 * 2);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  run<0>("f32 MUFU + bf16 pack (current)", out, in);
  run<1>("f16x2 MUFU", out, in);
  run<5>("bf16x2 MUFU", out, in);
  run<2>("poly only", out, in);
  run<3>("3/4 MUFU + 1/4 poly", out, in);
  run<7>("5/8 MUFU + 3/8 poly", out, in);
  run<4>("1/2 MUFU + 1/2 poly", out, in);
  run<8>("packed poly only", out, in);
  run<9>("3/4 MUFU + 1/4 packed poly", out, in);
  run<10>("5/8 MUFU + 3/8 packed poly", out, in);
  run<11>("1/2 MUFU + 1/2 packed poly", out, in);
  acc_check<<<1, 256>>>(out);
  float h[512];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  float wp = 0, wh = 0;
  for (int i = 0; i < 256; ++i) { wp = fmaxf(wp, h[2 * i]); wh = fmaxf(wh, h[2 * i + 1]); }
  printf("max rel err: poly %.3e   f16x2 (x > -13) %.3e\n", wp, wh);
  printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
