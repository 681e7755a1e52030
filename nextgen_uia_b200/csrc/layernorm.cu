// LayerNorm forward/backward and the Mona pre-scale, one warp per token row, 128-bit accesses,
// warp-shuffle reductions, statistics in fp32.  HBM-bound: fwd reads x once and writes y once.
//
// Reference semantics:
//   * timm Block norm1/norm2 (eps 1e-6, pinned dep) and CLIP LayerNorm (fp32 compute, eps 1e-5,
//     src/third_party/openai_clip/model.py:163-169): y = xhat*w + b
//   * Mona pre-scale  src/adapters/mona.py:125:  u = LN(x)*gamma + x*gammax   (LN eps 1e-5)
//   Backward of the frozen-affine LN produces dx only; the Mona pre-scale backward also produces
//   d(norm.weight), d(norm.bias), d(gamma), d(gammax) and the column sum of dy (= d project2.bias).
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

constexpr int kLnWarps = 4;
constexpr int kMaxD = 1024;

template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_fwd_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ b,
              const float* __restrict__ gamma, const float* __restrict__ gammax, T* __restrict__ y, int64_t ldy,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D, float eps) {
  constexpr int V = Vec<T>::N;
  constexpr int kMaxIt = kMaxD / (32 * V);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + warp;
  if (row >= M) return;
  const T* xr = x + int64_t(row) * ldx;
  float v[kMaxIt][V];
  float s = 0.f;
#pragma unroll
  for (int it = 0; it < kMaxIt; ++it) {
    const int c = (it * 32 + lane) * V;
    if (c < D) {
      Vec<T>::load(xr + c, v[it]);
#pragma unroll
      for (int i = 0; i < V; ++i) s += v[it][i];
    }
  }
  const float mean = warp_sum(s) / float(D);
  float q = 0.f;
#pragma unroll
  for (int it = 0; it < kMaxIt; ++it) {
    const int c = (it * 32 + lane) * V;
    if (c < D) {
#pragma unroll
      for (int i = 0; i < V; ++i) { const float d = v[it][i] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / float(D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  T* yr = y + int64_t(row) * ldy;
#pragma unroll
  for (int it = 0; it < kMaxIt; ++it) {
    const int c = (it * 32 + lane) * V;
    if (c < D) {
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float n = (v[it][i] - mean) * rstd * __ldg(w + c + i) + __ldg(b + c + i);
        if (gamma != nullptr) n = n * __ldg(gamma + c + i) + v[it][i] * __ldg(gammax + c + i);
        o[i] = n;
      }
      Vec<T>::store(yr + c, o);
    }
  }
}

// dx = rstd * (gh - mean(gh) - xhat * mean(gh * xhat)) (+ dres),  gh = g * w
template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_bwd_kernel(const T* __restrict__ g, int64_t ldg, const T* __restrict__ x, int64_t ldx, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ w, const T* __restrict__ dres, int64_t ldr,
              T* __restrict__ dx, int64_t lddx, int M, int D) {
  constexpr int V = Vec<T>::N;
  constexpr int kMaxIt = kMaxD / (32 * V);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + warp;
  if (row >= M) return;
  const float mu = mean[row], rs = rstd[row];
  float gh[kMaxIt][V], xh[kMaxIt][V];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int it = 0; it < kMaxIt; ++it) {
    const int c = (it * 32 + lane) * V;
    if (c < D) {
      float gv[V], xv[V];
      Vec<T>::load(g + int64_t(row) * ldg + c, gv);
      Vec<T>::load(x + int64_t(row) * ldx + c, xv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        gh[it][i] = gv[i] * __ldg(w + c + i);
        xh[it][i] = (xv[i] - mu) * rs;
        s1 += gh[it][i];
        s2 += gh[it][i] * xh[it][i];
      }
    }
  }
  s1 = warp_sum(s1) / float(D);
  s2 = warp_sum(s2) / float(D);
#pragma unroll
  for (int it = 0; it < kMaxIt; ++it) {
    const int c = (it * 32 + lane) * V;
    if (c < D) {
      float o[V];
      if (dres != nullptr) Vec<T>::load(dres + int64_t(row) * ldr + c, o);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] += rs * (gh[it][i] - s1 - xh[it][i] * s2);
      Vec<T>::store(dx + int64_t(row) * lddx + c, o);
    }
  }
}

// Backward of u = (xhat*w + b)*gamma + x*gammax with residual y = x + f(u):
//   dx = dy + du*gammax + LNbwd(du*gamma);  column sums -> dw, db, dgamma, dgammax, dcol(dy)
// Persistent: each warp walks rows with a grid stride and keeps per-lane column partials in
// registers; partials are combined through smem and one atomicAdd per column per CTA.
template <typename T, int D>
__global__ void __launch_bounds__(kLnWarps * 32)
mona_pre_bwd_kernel(const T* __restrict__ du, const T* __restrict__ dy, const T* __restrict__ x,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ w, const float* __restrict__ b,
                    const float* __restrict__ gamma, const float* __restrict__ gammax,
                    T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                    float* __restrict__ dgamma, float* __restrict__ dgammax, float* __restrict__ dycol, int M) {
  constexpr int V = Vec<T>::N;
  constexpr int kIt = D / (32 * V);
  static_assert(D % (32 * V) == 0, "D must be a multiple of 32 vectors");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a_w[kIt][V] = {}, a_b[kIt][V] = {}, a_g[kIt][V] = {}, a_gx[kIt][V] = {}, a_dy[kIt][V] = {};
  for (int row = blockIdx.x * kLnWarps + warp; row < M; row += gridDim.x * kLnWarps) {
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    // pass 1: row statistics of the LN backward + column partials
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int c = (it * 32 + lane) * V;
      float duv[V], xv[V];
      Vec<T>::load(du + size_t(row) * D + c, duv);
      Vec<T>::load(x + size_t(row) * D + c, xv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float pw = __ldg(w + c + i), pb = __ldg(b + c + i), pg = __ldg(gamma + c + i);
        const float xhat = (xv[i] - mu) * rs;
        const float n = xhat * pw + pb;
        a_g[it][i] += duv[i] * n;
        a_gx[it][i] += duv[i] * xv[i];
        const float dn = duv[i] * pg;
        a_w[it][i] += dn * xhat;
        a_b[it][i] += dn;
        const float g = dn * pw;
        s1 += g;
        s2 += g * xhat;
      }
    }
    s1 = warp_sum(s1) / float(D);
    s2 = warp_sum(s2) / float(D);
    // pass 2 (row is L1/L2 resident): dx
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int c = (it * 32 + lane) * V;
      float duv[V], dyv[V], xv[V], o[V];
      Vec<T>::load(du + size_t(row) * D + c, duv);
      Vec<T>::load(dy + size_t(row) * D + c, dyv);
      Vec<T>::load(x + size_t(row) * D + c, xv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float pw = __ldg(w + c + i), pg = __ldg(gamma + c + i), pgx = __ldg(gammax + c + i);
        const float xhat = (xv[i] - mu) * rs;
        const float g = duv[i] * pg * pw;
        a_dy[it][i] += dyv[i];
        o[i] = dyv[i] + duv[i] * pgx + rs * (g - s1 - xhat * s2);
      }
      Vec<T>::store(dx + size_t(row) * D + c, o);
    }
  }
  // combine the kLnWarps warps of this CTA, then one atomic per column
  __shared__ float red[kLnWarps][D];
  float* outs[5] = {dw, db, dgamma, dgammax, dycol};
#pragma unroll
  for (int which = 0; which < 5; ++which) {
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kIt; ++it)
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = (it * 32 + lane) * V + i;
        const float val = which == 0 ? a_w[it][i] : which == 1 ? a_b[it][i] : which == 2 ? a_g[it][i] : which == 3 ? a_gx[it][i] : a_dy[it][i];
        red[warp][c] = val;
      }
    __syncthreads();
    if (outs[which] != nullptr) {
      for (int c = threadIdx.x; c < D; c += kLnWarps * 32) {
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < kLnWarps; ++wv) t += red[wv][c];
        atomicAdd(outs[which] + c, t);
      }
    }
  }
}

template <typename T>
int ln_fwd_t(const ngu_ln_desc& d, cudaStream_t s) {
  const int grid = (d.M + kLnWarps - 1) / kLnWarps;
  ln_fwd_kernel<T><<<grid, kLnWarps * 32, 0, s>>>(reinterpret_cast<const T*>(d.x), d.ldx, d.w, d.b, d.gamma, d.gammax,
                                                   reinterpret_cast<T*>(d.y), d.ldy, d.mean, d.rstd, d.M, d.D, d.eps);
  return check_launch("ln_fwd");
}
template <typename T>
int ln_bwd_t(const ngu_ln_bwd_desc& d, cudaStream_t s) {
  const int grid = (d.M + kLnWarps - 1) / kLnWarps;
  ln_bwd_kernel<T><<<grid, kLnWarps * 32, 0, s>>>(reinterpret_cast<const T*>(d.g), d.ldg, reinterpret_cast<const T*>(d.x), d.ldx,
                                                   d.mean, d.rstd, d.w, reinterpret_cast<const T*>(d.dres), d.ldr,
                                                   reinterpret_cast<T*>(d.dx), d.lddx, d.M, d.D);
  return check_launch("ln_bwd");
}
template <typename T, int D>
int mona_pre_bwd_t(const ngu_mona_pre_bwd_desc& d, cudaStream_t s) {
  int grid = sm_count() * 4;
  const int need = (d.M + kLnWarps - 1) / kLnWarps;
  if (grid > need) grid = need;
  mona_pre_bwd_kernel<T, D><<<grid, kLnWarps * 32, 0, s>>>(
      reinterpret_cast<const T*>(d.du), reinterpret_cast<const T*>(d.dy), reinterpret_cast<const T*>(d.x), d.mean, d.rstd,
      d.w, d.b, d.gamma, d.gammax, reinterpret_cast<T*>(d.dx), d.dw, d.db, d.dgamma, d.dgammax, d.dycol, d.M);
  return check_launch("mona_pre_bwd");
}

int check_ln_shape(int M, int D, int dtype, const char* what) {
  const int V = dtype == NGU_F32 ? 4 : 8;
  if (M <= 0 || D <= 0 || D > kMaxD || (D % V)) {
    set_last_error("%s: need 0 < D <= %d and D %% %d == 0 (got M=%d D=%d)", what, kMaxD, V, M, D);
    return NGU_ERR_SHAPE;
  }
  return NGU_OK;
}

}  // namespace

int ln_fwd(const ngu_ln_desc& d, cudaStream_t s) {
  if (int rc = check_ln_shape(d.M, d.D, d.dtype, "ln_fwd")) return rc;
  if ((d.gamma == nullptr) != (d.gammax == nullptr)) { set_last_error("ln_fwd: gamma and gammax go together"); return NGU_ERR_ARG; }
  const int V = d.dtype == NGU_F32 ? 4 : 8;
  if ((d.ldx % V) || (d.ldy % V)) { set_last_error("ln_fwd: row strides must keep 16-byte alignment"); return NGU_ERR_ALIGN; }
  return d.dtype == NGU_F32 ? ln_fwd_t<float>(d, s) : ln_fwd_t<bf16>(d, s);
}
int ln_bwd(const ngu_ln_bwd_desc& d, cudaStream_t s) {
  if (int rc = check_ln_shape(d.M, d.D, d.dtype, "ln_bwd")) return rc;
  const int V = d.dtype == NGU_F32 ? 4 : 8;
  if ((d.ldx % V) || (d.ldg % V) || (d.lddx % V) || (d.dres && (d.ldr % V))) { set_last_error("ln_bwd: row strides must keep 16-byte alignment"); return NGU_ERR_ALIGN; }
  return d.dtype == NGU_F32 ? ln_bwd_t<float>(d, s) : ln_bwd_t<bf16>(d, s);
}
int mona_pre_bwd(const ngu_mona_pre_bwd_desc& d, cudaStream_t s) {
  if (d.M <= 0) { set_last_error("mona_pre_bwd: empty"); return NGU_ERR_SHAPE; }
  if (d.D == 768) return d.dtype == NGU_F32 ? mona_pre_bwd_t<float, 768>(d, s) : mona_pre_bwd_t<bf16, 768>(d, s);
  if (d.D == 1024) return d.dtype == NGU_F32 ? mona_pre_bwd_t<float, 1024>(d, s) : mona_pre_bwd_t<bf16, 1024>(d, s);
  if (d.D == 256) return d.dtype == NGU_F32 ? mona_pre_bwd_t<float, 256>(d, s) : mona_pre_bwd_t<bf16, 256>(d, s);
  set_last_error("mona_pre_bwd: embed dim %d not instantiated (256, 768, 1024)", d.D);
  return NGU_ERR_SHAPE;
}

}  // namespace ngu
