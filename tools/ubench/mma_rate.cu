// Microbenchmark: issue rate / latency of small tcgen05.mma shapes from one thread (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../nextgen_uia_b200/csrc -I../../include mma_rate.cu -o mma_rate
#include <cstdio>
#include "common.cuh"
using namespace ngu;

// mode 0: SS K-major x K-major; 1: TS (A in TMEM) x B MN-major; 2: SS A MN-major x B MN-major
// chains: number of distinct accumulators cycled through (1 = fully dependent chain)
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int chains, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  const uint32_t sbar = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(sbar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  // zero smem operands (values irrelevant)
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  if (threadIdx.x < 32) {
    const uint32_t sA = base, sB = base + 32768;
    uint32_t idesc;
    uint64_t da, db;
    if (mode == 0) { idesc = make_idesc_bf16(128, N); da = make_smem_desc_sw128(sA, 16, 1024); db = make_smem_desc_sw128(sB, 16, 1024); }
    else if (mode == 1) { idesc = make_idesc_bf16(128, N, 0, 1); da = 0; db = make_smem_desc_sw128(sB, 0, 1024); }
    else { idesc = make_idesc_bf16(128, N, 1, 1); da = make_smem_desc_sw128(sA, 16384, 1024); db = make_smem_desc_sw128(sB, 0, 1024); }
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        if (chains == 0) {
          // fully unrolled, compile-time offsets, two interleaved accumulators (what a tuned issuer looks like)
          for (int o = 0; o < reps / 16; ++o) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t d = tmem + 256 + uint32_t(i & 1) * 64;
              if (mode == 1) umma_ts(d, tmem + (i & 7) * 8, db + uint64_t((i & 3) * 128), idesc, 1u);
              else umma_ss(d, da + uint64_t((i & 3) * 2), db + uint64_t((i & 3) * 2), idesc, 1u);
            }
          }
        } else
        for (int i = 0; i < reps; ++i) {
          const uint32_t d = tmem + 256 + uint32_t(i % chains) * 64;   // chains <= 4 when N = 64
          if (mode == 1) umma_ts(d, tmem + (i & 7) * 8, db + uint64_t((i & 3) * 128), idesc, 1u);
          else umma_ss(d, da + uint64_t((i & 3) * 2), db + uint64_t((i & 3) * 2), idesc, 1u);
        }
        umma_commit(sbar);
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(sbar, rep & 1);
      t2 = clock64();
    }
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  const int reps = 64;
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {16, 64, 128, 256})
      for (int chains : {0, 1}) {
        if (N > 128 && chains == 0) continue;
        k<<<1, 128, 70 * 1024>>>(mode, N, chains, reps, d);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
        printf("mode %d (%s) N=%3d chains=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", mode,
               mode == 0 ? "SS K/K" : mode == 1 ? "TS B-MN" : "SS MN/MN", N, chains, double(h[0]) / reps, double(h[1]) / reps, N / 2);
      }
  return 0;
}
