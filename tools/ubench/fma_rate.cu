// FP32 FMA issue rate on sm_100a: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), vs warps per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fma_rate tools/ubench/fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d)) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)), "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float a, float b) {
  float acc[16];
  float2 acc2[8];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  for (int i = 0; i < 8; ++i) acc2[i] = make_float2(threadIdx.x + i, i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc2[i] = ffma2(acc2[i], make_float2(a, a), make_float2(b, b));
    } else {   // scalar-broadcast tap form
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc2[i] = ffma2(make_float2(a, a), acc2[(i + 1) & 7], acc2[i]);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  for (int i = 0; i < 8; ++i) s += acc2[i].x + acc2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters, 1.0001f, 0.5f);
        else if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters, 1.0001f, 0.5f);
        else k<2><<<148, warps * 32>>>(out, cyc, iters, 1.0001f, 0.5f);
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double instr = double(iters) * (mode == 0 ? 64 : 32);        // per warp
      const double per_smsp = instr * (warps / 4);
      printf("%s warps/SM %2d: %.2f cycles per warp-instr per SMSP  (%.1f FMA lanes/clk/SM)\n", mode == 0 ? "FFMA " : mode == 1 ? "FFMA2" : "FFMA2b",
             warps, double(c) / per_smsp, (mode == 0 ? 1.0 : 2.0) * 32.0 * per_smsp * 4 / double(c));
    }
  return 0;
}
