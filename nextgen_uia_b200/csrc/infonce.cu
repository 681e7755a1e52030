// Symmetric InfoNCE (CLIP contrastive loss), forward + backward in one call, fp32 math.
// Reference: src/losses/losses.py:23-47
//     Ihat = I / max(||I||, 1e-12), That likewise;  L = Ihat That^T / tau
//     loss = ( CE(L, arange) + CE(L^T, arange) ) / 2          (mean reduction)
// Data-parallel form: every rank holds the all-gathered normalised features
// [Bg, E]; it forms the full L and L^T (redundant, tiny), the row/column log-sum-exps, the global
// loss, and the gradient rows of its own local slice [r0, r0+Bl) with no second collective.
//     dIhat_i = 1/(2 Bg tau) * sum_j ( exp(L_ij - rlse_i) + exp(L_ij - clse_j) - 2 delta_ij ) That_j
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

// one warp per row: xhat = x / max(||x||, eps) (fp32 out), norm saved
template <typename T>
__global__ void normalize_fwd_kernel(const T* __restrict__ x, float* __restrict__ xhat, float* __restrict__ norm, int B, int E) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float s = 0.f;
  for (int c = lane; c < E; c += 32) { const float v = to_f32<T>(x[size_t(row) * E + c]); s += v * v; }
  const float n = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  if (lane == 0) norm[row] = n;
  const float inv = 1.f / n;
  for (int c = lane; c < E; c += 32) xhat[size_t(row) * E + c] = to_f32<T>(x[size_t(row) * E + c]) * inv;
}
// dx = gscale * (dxhat - xhat * <xhat, dxhat>) / norm
template <typename T>
__global__ void normalize_bwd_kernel(const float* __restrict__ dxhat, const float* __restrict__ xhat,
                                     const float* __restrict__ norm, const float* __restrict__ gscale,
                                     T* __restrict__ dx, int B, int E) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float s = 0.f;
  for (int c = lane; c < E; c += 32) s += dxhat[size_t(row) * E + c] * xhat[size_t(row) * E + c];
  s = warp_sum(s);
  const float k = (gscale ? *gscale : 1.f) / norm[row];
  for (int c = lane; c < E; c += 32)
    dx[size_t(row) * E + c] = from_f32<T>(k * (dxhat[size_t(row) * E + c] - xhat[size_t(row) * E + c] * s));
}

// row log-sum-exp of L [n, n] (one warp per row) + accumulate sum_i (lse_i - L_ii) * w into *loss
__global__ void row_lse_kernel(const float* __restrict__ L, float* __restrict__ lse, float* __restrict__ loss, int n, float w) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* r = L + size_t(row) * n;
  float mx = -INFINITY;
  for (int c = lane; c < n; c += 32) mx = fmaxf(mx, r[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s += expf(r[c] - mx);
  s = warp_sum(s);
  if (lane == 0) {
    const float l = mx + logf(s);
    lse[row] = l;
    atomicAdd(loss, (l - r[row]) * w);
  }
}

// dA_i = coef * sum_j ( exp(L_ij - lse_a[i]) + exp(L_ij - lse_b[j]) - 2 delta_ij ) Bf_j   for i in [r0, r0+Bl)
// block per local row, threads over the embedding dim; the weight row is staged in smem.
__global__ void __launch_bounds__(256)
dfeat_kernel(const float* __restrict__ L, const float* __restrict__ lse_a, const float* __restrict__ lse_b,
             const float* __restrict__ Bf, float* __restrict__ dA, int n, int E, int r0, float coef) {
  pdl_prologue();
  extern __shared__ float gw[];  // [n]
  const int i = r0 + blockIdx.x;
  const float la = lse_a[i];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float l = L[size_t(i) * n + j];
    gw[j] = (expf(l - la) + expf(l - lse_b[j]) - (j == i ? 2.f : 0.f)) * coef;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = fmaf(gw[j], Bf[size_t(j) * E + e], acc);
    dA[size_t(blockIdx.x) * E + e] = acc;
  }
}

// bf16 product path: G[i - r0][j] = coef (exp(L_ij - lse_a[i]) + exp(L_ij - lse_b[j]) - 2 delta_ij) as the bf16 A operand of the
// tcgen05 feature-gradient GEMM  dA = G . Bf   (one block per local row, coalesced over j)
__global__ void __launch_bounds__(256)
infonce_g_kernel(const float* __restrict__ L, const float* __restrict__ lse_a, const float* __restrict__ lse_b, bf16* __restrict__ G,
                 int n, int r0, float coef) {
  pdl_prologue();
  const int i = r0 + blockIdx.x;
  const float la = lse_a[i];
  const float* Lr = L + size_t(i) * n;
  bf16* Gr = G + size_t(blockIdx.x) * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float l = Lr[j];
    Gr[j] = __float2bfloat16_rn((expf(l - la) + expf(l - lse_b[j]) - (j == i ? 2.f : 0.f)) * coef);
  }
}

}  // namespace

int infonce_normalize(const void* x, float* xhat, float* norm, int B, int E, int dtype, cudaStream_t st) {
  if (B <= 0 || E <= 0) { set_last_error("infonce_normalize: empty"); return NGU_ERR_SHAPE; }
  const int wpb = 4;
  const int grid = (B + wpb - 1) / wpb;
  if (dtype == NGU_F32) launch_pdl(normalize_fwd_kernel<float>, dim3(grid), dim3(wpb * 32), size_t(0), st, reinterpret_cast<const float*>(x), xhat, norm, B, E);
  else launch_pdl(normalize_fwd_kernel<bf16>, dim3(grid), dim3(wpb * 32), size_t(0), st, reinterpret_cast<const bf16*>(x), xhat, norm, B, E);
  return check_launch("infonce_normalize");
}

int infonce_normalize_bwd(const float* dxhat, const float* xhat, const float* norm, const float* gscale, void* dx, int B, int E,
                          int dtype, cudaStream_t st) {
  const int wpb = 4;
  const int grid = (B + wpb - 1) / wpb;
  if (dtype == NGU_F32) launch_pdl(normalize_bwd_kernel<float>, dim3(grid), dim3(wpb * 32), size_t(0), st, dxhat, xhat, norm, gscale, reinterpret_cast<float*>(dx), B, E);
  else launch_pdl(normalize_bwd_kernel<bf16>, dim3(grid), dim3(wpb * 32), size_t(0), st, dxhat, xhat, norm, gscale, reinterpret_cast<bf16*>(dx), B, E);
  return check_launch("infonce_normalize_bwd");
}

// ws: fp32 workspace of 2*Bg*Bg + 2*Bg floats.  loss must be zeroed by this call.
// Lt = L^T for the [n, n] fp32 logits (32 x 32 smem tiles, conflict-free): the text->image logits are the transpose of the
// image->text logits, so a second n x n x E GEMM is not needed.
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (by + r < n && bx + tx < n) tile[r][tx] = in[size_t(by + r) * n + bx + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (bx + r < n && by + tx < n) out[size_t(bx + r) * n + by + tx] = tile[tx][r];
}

int infonce_core(const ngu_infonce_desc& d, cudaStream_t st) {
  const int n = d.Bg, E = d.E;
  if (n <= 0 || E <= 0 || d.Bl <= 0 || d.r0 < 0 || d.r0 + d.Bl > n) { set_last_error("infonce: bad shape Bg=%d Bl=%d r0=%d", n, d.Bl, d.r0); return NGU_ERR_SHAPE; }
  if (d.temperature <= 0.f) { set_last_error("infonce: temperature must be > 0"); return NGU_ERR_ARG; }
  float* L = d.ws;
  float* Lt = L + size_t(n) * n;
  float* rlse = Lt + size_t(n) * n;
  float* clse = rlse + n;
  cudaError_t e = cudaMemsetAsync(d.loss, 0, sizeof(float), st);
  if (e != cudaSuccess) return cuda_status(e, "infonce memset");
  // bf16 product path (ihat16 given): logits and both feature-gradient contractions on the tcgen05 GEMM (bf16 operands, fp32
  // accumulate / output); fp32 check mode: CUDA cores
  const bool tc = d.ihat16 != nullptr;
  if (tc && (!d.that16 || (d.dihat && (!d.ihat16_t || !d.that16_t || !d.g_ws)) || (E % 8) || (n % 8))) {
    set_last_error("infonce: the bf16 path needs ihat16 / that16 (+ transposes and g_ws for gradients), E and Bg multiples of 8");
    return NGU_ERR_ARG;
  }
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = n; g.N = n; g.K = E; g.lda = E; g.ldb = E; g.ldc = n; g.alpha = 1.f / d.temperature; g.dtype = NGU_F32;
  g.A = d.ihat; g.B = d.that; g.C = L;
  if (tc) {
    g.A = d.ihat16; g.B = d.that16; g.dtype = NGU_BF16; g.c_dtype = NGU_F32;
    if (int rc = gemm_tc(g, st)) return rc;
  } else if (int rc = gemm_simt(g, st)) return rc;
  {
    const dim3 tg((n + 31) / 32, (n + 31) / 32);
    launch_pdl(transpose_kernel, dim3(tg), dim3(256), size_t(0), st, L, Lt, n);
    if (int rc = check_launch("infonce transpose")) return rc;
  }
  const int wpb = 4;
  const int grid = (n + wpb - 1) / wpb;
  launch_pdl(row_lse_kernel, dim3(grid), dim3(wpb * 32), size_t(0), st, L, rlse, d.loss, n, 0.5f / float(n));
  if (int rc = check_launch("infonce row_lse")) return rc;
  launch_pdl(row_lse_kernel, dim3(grid), dim3(wpb * 32), size_t(0), st, Lt, clse, d.loss, n, 0.5f / float(n));
  if (int rc = check_launch("infonce col_lse")) return rc;
  if (d.dihat != nullptr && tc) {
    const float coef = 1.f / (2.f * float(n) * d.temperature);
    bf16* G1 = reinterpret_cast<bf16*>(d.g_ws);
    bf16* G2 = G1 + size_t(d.Bl) * n;
    launch_pdl(infonce_g_kernel, dim3(d.Bl), dim3(256), size_t(0), st, L, rlse, clse, G1, n, d.r0, coef);
    if (int rc = check_launch("infonce G (image rows)")) return rc;
    launch_pdl(infonce_g_kernel, dim3(d.Bl), dim3(256), size_t(0), st, Lt, clse, rlse, G2, n, d.r0, coef);
    if (int rc = check_launch("infonce G (text rows)")) return rc;
    GemmArgs h;
    memset(&h, 0, sizeof(h));
    h.M = d.Bl; h.N = E; h.K = n; h.lda = n; h.ldb = n; h.ldc = E; h.alpha = 1.f; h.dtype = NGU_BF16; h.c_dtype = NGU_F32;
    h.A = G1; h.B = d.that16_t; h.C = d.dihat;          // dIhat = G1 . That
    if (int rc = gemm_tc(h, st)) return rc;
    h.A = G2; h.B = d.ihat16_t; h.C = d.dthat;          // dThat = G2 . Ihat
    if (int rc = gemm_tc(h, st)) return rc;
  } else if (d.dihat != nullptr) {
    const float coef = 1.f / (2.f * float(n) * d.temperature);
    const int smem = n * int(sizeof(float));
    if (smem > 48 * 1024) {
      static bool attr = false;
      if (!attr) {
        cudaError_t ea = cudaFuncSetAttribute(dfeat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ea != cudaSuccess) return cuda_status(ea, "infonce dfeat attr");
        attr = true;
      }
      if (smem > 200 * 1024) { set_last_error("infonce (fp32 check mode): global batch %d exceeds the %d rows the weight-row staging holds", n, 200 * 1024 / 4); return NGU_ERR_SHAPE; }
    }
    launch_pdl(dfeat_kernel, dim3(d.Bl), dim3(256), size_t(smem), st, L, rlse, clse, d.that, d.dihat, n, E, d.r0, coef);
    if (int rc = check_launch("infonce dI")) return rc;
    launch_pdl(dfeat_kernel, dim3(d.Bl), dim3(256), size_t(smem), st, Lt, clse, rlse, d.ihat, d.dthat, n, E, d.r0, coef);
    if (int rc = check_launch("infonce dT")) return rc;
  }
  return NGU_OK;
}

}  // namespace ngu
