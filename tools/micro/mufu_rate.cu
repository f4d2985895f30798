// Microbenchmark: MUFU.EX2 / FFMA issue rates per SM on this GPU (build: nvcc -arch=sm_100a -O3 -o mufu_rate mufu_rate.cu)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      else if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i])); }
      else if (MODE == 3) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(a[i])); a[i] = __uint_as_float(u | 0x3f000000u); }
      else if (MODE == 4) { asm volatile("max.f32 %0, %0, %0, %0;" : "+f"(a[i])); }
      else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(a[i])); a[i] = __uint_as_float(u | 0x3f000000u); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4 * sizeof(float));
  int dev; cudaGetDevice(&dev); int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  for (int warps = 2; warps <= 4; warps *= 2) {
    for (int mode = 0; mode < 6; ++mode) {
      const int iters = 20000, threads = warps * 32 * 4;  // warps per SMSP
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      auto run = [&]() { switch (mode) { case 0: k<0><<<148, threads>>>(d, iters, 0.1f); break; case 1: k<1><<<148, threads>>>(d, iters, 0.1f); break; case 2: k<2><<<148, threads>>>(d, iters, 0.1f); break; case 3: k<3><<<148, threads>>>(d, iters, 0.1f); break; case 4: k<4><<<148, threads>>>(d, iters, 0.1f); break; default: k<5><<<148, threads>>>(d, iters, 0.1f); } };
      run(); cudaDeviceSynchronize();
      cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double ops = (double)iters * 8 * threads;  // per SM
      double cycles = ms * 1e-3 * 1.965e9;
      printf("warps/SMSP=%2d mode=%s: %.2f ops/clk/SM (at 1965 MHz)\n", warps, mode == 0 ? "ex2" : mode == 1 ? "fma" : mode == 2 ? "ex2+fma pairs" : mode == 3 ? "cvt.bf16x2 (+lop)" : mode == 4 ? "max3" : "ex2+cvt pairs", ops / cycles);
    }
  }
  return 0;
}
