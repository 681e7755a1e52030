// CUDA-core GEMM with the same epilogue contract as gemm_tc.cu.  This is the "fp32 check mode" of
// the parity contract (north_star: 1e-4 relative in fp32): operands, accumulation and outputs in
// fp32, no tensor cores.  A bf16-I/O instantiation exists for tests that bisect the tcgen05 path.
//   C[M,N] = epi( alpha * (A[M,K]·B[N,K]^T + A2[M,K2]·B2[N,K2]^T) )
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs a) {
  pdl_prologue();
  __shared__ float sA[TK][TM + 1];
  __shared__ float sB[TK][TN + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int pass = 0; pass < 2; ++pass) {
    const T* A = reinterpret_cast<const T*>(pass == 0 ? a.A : a.A2);
    const T* B = reinterpret_cast<const T*>(pass == 0 ? a.B : a.B2);
    const int lda = pass == 0 ? a.lda : a.lda2;
    const int ldb = pass == 0 ? a.ldb : a.ldb2;
    const int K = pass == 0 ? a.K : a.K2;
    if (K <= 0 || A == nullptr) continue;
    for (int k0 = 0; k0 < K; k0 += TK) {
      for (int i = threadIdx.x; i < TM * TK; i += 256) {
        const int r = i / TK, c = i % TK;
        const int gm = m0 + r, gk = k0 + c;
        sA[c][r] = (gm < a.M && gk < K) ? to_f32<T>(A[size_t(gm) * lda + gk]) : 0.f;
        const int gn = n0 + r;
        sB[c][r] = (gn < a.N && gk < K) ? to_f32<T>(B[size_t(gn) * ldb + gk]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { av[i] = sA[k][ty * 4 + i]; bv[i] = sB[k][tx * 4 + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  T* C = reinterpret_cast<T*>(a.C);
  T* Pre = reinterpret_cast<T*>(a.Pre);
  const T* aux = reinterpret_cast<const T*>(a.aux);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      float x = acc[i][j] * a.alpha;
      if (a.bias) x += a.bias[gn];
      const float ax = (a.aux_mode != NGU_AUX_NONE) ? to_f32<T>(aux[size_t(gm) * a.ldaux + gn]) : 0.f;
      float dact = x;  // what save_pre stores: act'(pre) when there is an activation, else the pre-activation itself
      if (a.aux_mode == NGU_AUX_DACT) {
        x *= ax;  // aux = act'(pre) saved by the forward
      } else {
        if (a.act == NGU_ACT_GELU) {
          const float cdf = 0.5f * (1.f + erff(x * 0.7071067811865476f));
          const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
          dact = cdf + x * pdf;
          x = x * cdf;
        } else if (a.act == NGU_ACT_QUICKGELU) {
          const float sg = 1.f / (1.f + expf(-1.702f * x));
          dact = sg * (1.f + 1.702f * x * (1.f - sg));
          x = x * sg;
        }
        if (a.aux_mode == NGU_AUX_RESIDUAL) x += ax;
      }
      if (a.save_pre) Pre[size_t(gm) * a.ldpre + gn] = from_f32<T>(dact);
      C[size_t(gm) * a.ldc + gn] = from_f32<T>(x);
    }
  }
}

}  // namespace

int gemm_simt(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) { set_last_error("gemm_simt: empty problem"); return NGU_ERR_SHAPE; }
  dim3 grid((a.N + TN - 1) / TN, (a.M + TM - 1) / TM);
  if (grid.y > 65535) { set_last_error("gemm_simt: M too large for check mode"); return NGU_ERR_SHAPE; }
  if (a.dtype == NGU_F32) launch_pdl(gemm_simt_kernel<float>, dim3(grid), dim3(256), size_t(0), stream, a);
  else launch_pdl(gemm_simt_kernel<bf16>, dim3(grid), dim3(256), size_t(0), stream, a);
  return check_launch("gemm_simt");
}

}  // namespace ngu
