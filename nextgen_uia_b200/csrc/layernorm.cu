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
  pdl_prologue();
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
  pdl_prologue();
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
// CTA = D/2 threads (D/64 warps), persistent over tiles of (D/64) rows.
//   phase A (warp per row, 128-bit accesses): row reductions s1/s2, dx written, the row's du/x/dy staged in smem
//   phase B (thread per column pair): walks the staged rows and accumulates the five column sums in registers
// so the row data is read from HBM once and the column reductions never touch registers-per-row budgets.
template <typename T> struct Pair;
template <> struct Pair<bf16> { typedef uint32_t type; static NGU_DEVINL float2 get(uint32_t u) { return unpack_bf16x2(u); } };
template <> struct Pair<float> { typedef float2 type; static NGU_DEVINL float2 get(float2 u) { return u; } };

template <typename T, int D>
__global__ void __launch_bounds__(D / 2, (D <= 768 ? 2 : 1))
mona_pre_bwd_kernel(const T* __restrict__ du, const T* __restrict__ dy, const T* __restrict__ x,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ w, const float* __restrict__ b,
                    const float* __restrict__ gamma, const float* __restrict__ gammax,
                    T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                    float* __restrict__ dgamma, float* __restrict__ dgammax, float* __restrict__ dycol, int M) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  constexpr int kIt = D / (32 * V);
  constexpr int R = D / 64;  // warps per CTA = rows per tile
  static_assert(D % (32 * V) == 0 && D % 64 == 0, "unsupported width");
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  float* sp = reinterpret_cast<float*>(smem_dyn);            // [5][D]: w, b, gamma, gammax, gamma*w
  float* srow = sp + 5 * D;                                  // [R][2]: mean, rstd
  T* sdu = reinterpret_cast<T*>(srow + 2 * R);               // [R][D]
  T* sx = sdu + R * D;
  T* sdy = sx + R * D;
  for (int c = threadIdx.x; c < D; c += D / 2) {
    sp[c] = w[c]; sp[D + c] = b[c]; sp[2 * D + c] = gamma[c]; sp[3 * D + c] = gammax[c]; sp[4 * D + c] = gamma[c] * w[c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c2 = threadIdx.x * 2;  // this thread's column pair in phase B
  const float pw0 = sp[c2], pw1 = sp[c2 + 1], pb0 = sp[D + c2], pb1 = sp[D + c2 + 1], pg0 = sp[2 * D + c2], pg1 = sp[2 * D + c2 + 1];
  float aw0 = 0, aw1 = 0, ab0 = 0, ab1 = 0, ag0 = 0, ag1 = 0, agx0 = 0, agx1 = 0, ady0 = 0, ady1 = 0;
  const int ntiles = (M + R - 1) / R;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row = tile * R + warp;
    // ---------------- phase A
    if (row < M) {
      const float mu = mean[row], rs = rstd[row];
      if (lane == 0) { srow[2 * warp] = mu; srow[2 * warp + 1] = rs; }
      // issue every global load of the row first (3 tensors x kIt vectors in flight), stage them, then compute from smem
      uint4 raw[3][kIt];
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int c = (it * 32 + lane) * V;
        raw[0][it] = *reinterpret_cast<const uint4*>(du + size_t(row) * D + c);
        raw[1][it] = *reinterpret_cast<const uint4*>(x + size_t(row) * D + c);
        raw[2][it] = *reinterpret_cast<const uint4*>(dy + size_t(row) * D + c);
      }
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int c = (it * 32 + lane) * V;
        *reinterpret_cast<uint4*>(sdu + warp * D + c) = raw[0][it];
        *reinterpret_cast<uint4*>(sx + warp * D + c) = raw[1][it];
        *reinterpret_cast<uint4*>(sdy + warp * D + c) = raw[2][it];
      }
      // With g = du * (gamma w):  s1 = mean(g),  s2 = mean(g * xhat) = rs * (mean(g x) - mu * s1);
      // dx = dy + du * (gammax + rs * gamma w) - rs * s1 - rs^2 * s2 * (x - mu)      (4 FMA-class ops per element)
      float s1 = 0.f, tx = 0.f;
#pragma unroll 1
      for (int it = 0; it < kIt; ++it) {
        const int c = (it * 32 + lane) * V;
        float duv[V], xv[V];
        Vec<T>::load(sdu + warp * D + c, duv);
        Vec<T>::load(sx + warp * D + c, xv);
#pragma unroll
        for (int i4 = 0; i4 < V; i4 += 4) {
          const float4 gw = lds128_volatile(&sp[4 * D + c + i4]);
          const float gw4[4] = {gw.x, gw.y, gw.z, gw.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float g = duv[i4 + i] * gw4[i];
            s1 += g;
            tx = fmaf(g, xv[i4 + i], tx);
          }
        }
      }
      s1 = warp_sum(s1) * (1.0f / D);
      tx = warp_sum(tx) * (1.0f / D);
      const float s2 = rs * (tx - mu * s1);
      const float cx = rs * rs * s2;                 // coefficient of x
      const float c0 = fmaf(cx, mu, -rs * s1);       // row constant
#pragma unroll 1
      for (int it = 0; it < kIt; ++it) {
        const int c = (it * 32 + lane) * V;
        float o[V], dyv[V], duv[V], xv[V];
        Vec<T>::load(sdy + warp * D + c, dyv);
        Vec<T>::load(sdu + warp * D + c, duv);
        Vec<T>::load(sx + warp * D + c, xv);
#pragma unroll
        for (int i4 = 0; i4 < V; i4 += 4) {
          const float4 gw = lds128_volatile(&sp[4 * D + c + i4]), gx = lds128_volatile(&sp[3 * D + c + i4]);
          const float gw4[4] = {gw.x, gw.y, gw.z, gw.w}, gx4[4] = {gx.x, gx.y, gx.z, gx.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float k1 = fmaf(rs, gw4[i], gx4[i]);
            o[i4 + i] = fmaf(-cx, xv[i4 + i], fmaf(duv[i4 + i], k1, dyv[i4 + i])) + c0;
          }
        }
        Vec<T>::store(dx + size_t(row) * D + c, o);
      }
    }
    __syncthreads();
    // ---------------- phase B
    const int nr = min(R, M - tile * R);
    typedef typename Pair<T>::type P;
    // per column: U = sum du * xhat, Sd = sum du, sum du * x, sum dy; the parameter gradients are linear in them
    for (int r = 0; r < nr; ++r) {
      const float mu = srow[2 * r], rs = srow[2 * r + 1];
      const float2 d2 = Pair<T>::get(reinterpret_cast<const P*>(sdu + r * D)[threadIdx.x]);
      const float2 x2 = Pair<T>::get(reinterpret_cast<const P*>(sx + r * D)[threadIdx.x]);
      const float2 y2 = Pair<T>::get(reinterpret_cast<const P*>(sdy + r * D)[threadIdx.x]);
      const float t0 = d2.x * x2.x, t1 = d2.y * x2.y;
      agx0 += t0; agx1 += t1;
      aw0 = fmaf(rs, fmaf(-mu, d2.x, t0), aw0); aw1 = fmaf(rs, fmaf(-mu, d2.y, t1), aw1);    // U
      ab0 += d2.x; ab1 += d2.y;                                                              // Sd
      ady0 += y2.x; ady1 += y2.y;
    }
    __syncthreads();
  }
  // u = (xhat w + b) gamma + x gammax:  d w = gamma U,  d b = gamma Sd,  d gamma = w U + b Sd
  ag0 = fmaf(pw0, aw0, pb0 * ab0); ag1 = fmaf(pw1, aw1, pb1 * ab1);
  aw0 *= pg0; aw1 *= pg1; ab0 *= pg0; ab1 *= pg1;
  if (dw) { atomicAdd(dw + c2, aw0); atomicAdd(dw + c2 + 1, aw1); }
  if (db) { atomicAdd(db + c2, ab0); atomicAdd(db + c2 + 1, ab1); }
  if (dgamma) { atomicAdd(dgamma + c2, ag0); atomicAdd(dgamma + c2 + 1, ag1); }
  if (dgammax) { atomicAdd(dgammax + c2, agx0); atomicAdd(dgammax + c2 + 1, agx1); }
  if (dycol) { atomicAdd(dycol + c2, ady0); atomicAdd(dycol + c2 + 1, ady1); }
}

// ---------------------------------------------------------------------------------------------------
// Fixed-width fast paths (D known at compile time): persistent warps walk rows with a grid stride and
// keep the affine parameters in registers, so the inner loop is pure 128-bit row traffic + shuffles.
// ---------------------------------------------------------------------------------------------------
constexpr int kFastWarps = 8;

template <typename T, int D, bool MIX>
__global__ void __launch_bounds__(kFastWarps * 32, 2)   // <= 128 registers: two CTAs per SM also for the MIX variant (155 registers unconstrained)
ln_fwd_fast_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ b,
                   const float* __restrict__ gamma, const float* __restrict__ gammax, T* __restrict__ y, int64_t ldy,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, float eps) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  constexpr int IT = D / (32 * V);
  constexpr int RW = 2;  // rows in flight per warp (memory-level parallelism)
  static_assert(D % (32 * V) == 0, "row must split into whole 16-byte vectors per lane");
  __shared__ __align__(16) float sp[MIX ? 3 : 2][D];
  for (int c = threadIdx.x; c < D; c += kFastWarps * 32) {
    if (MIX) { const float g = gamma[c]; sp[0][c] = w[c] * g; sp[1][c] = b[c] * g; sp[2][c] = gammax[c]; }
    else { sp[0][c] = w[c]; sp[1][c] = b[c]; }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // plain LayerNorm: this lane's 2 x D/32 affine parameters live in registers for the whole kernel (the mix variant needs a third
  // vector and re-reads shared memory per row to stay under 128 registers)
  float pw[MIX ? 1 : IT][V], pbv[MIX ? 1 : IT][V];
  if (!MIX) {
#pragma unroll
    for (int it = 0; it < IT; ++it)
#pragma unroll
      for (int i = 0; i < V; ++i) { pw[it][i] = sp[0][(it * 32 + lane) * V + i]; pbv[it][i] = sp[1][(it * 32 + lane) * V + i]; }
  }
  for (int row0 = (blockIdx.x * kFastWarps + warp) * RW; row0 < M; row0 += gridDim.x * kFastWarps * RW) {
    uint4 raw[RW][IT];
#pragma unroll
    for (int r = 0; r < RW; ++r)
      if (row0 + r < M) {
#pragma unroll
        for (int it = 0; it < IT; ++it) raw[r][it] = *reinterpret_cast<const uint4*>(x + int64_t(row0 + r) * ldx + (it * 32 + lane) * V);
      }
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int row = row0 + r;
      if (row >= M) break;
      // single pass with a shift (first element of the row) so E[(x-k)^2] - E[x-k]^2 does not cancel
      float f0[V];
      Vec<T>::unpack(raw[r][0], f0);
      const float k = __shfl_sync(0xffffffffu, f0[0], 0);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        float f[V];
        Vec<T>::unpack(raw[r][it], f);
#pragma unroll
        for (int i = 0; i < V; ++i) { const float d = f[i] - k; s1 += d; s2 = fmaf(d, d, s2); }
      }
      s1 = warp_sum(s1) * (1.0f / D);
      s2 = warp_sum(s2) * (1.0f / D);
      const float mean = k + s1;
      const float rstd = rsqrtf(fmaxf(s2 - s1 * s1, 0.f) + eps);
      if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
      }
      T* yr = y + int64_t(row) * ldy;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int c = (it * 32 + lane) * V;
        float f[V], o[V];
        Vec<T>::unpack(raw[r][it], f);
        if (!MIX) {
#pragma unroll
          for (int i = 0; i < V; ++i) o[i] = fmaf((f[i] - mean) * rstd, pw[it][i], pbv[it][i]);
          Vec<T>::store(yr + c, o);
          continue;
        }
#pragma unroll
        for (int i4 = 0; i4 < V; i4 += 4) {
          const float4 pa = lds128_volatile(&sp[0][c + i4]), pb = lds128_volatile(&sp[1][c + i4]);
          const float a4[4] = {pa.x, pa.y, pa.z, pa.w}, b4[4] = {pb.x, pb.y, pb.z, pb.w};
          float g4[4] = {0.f, 0.f, 0.f, 0.f};
          if (MIX) { const float4 pg = lds128_volatile(&sp[2][c + i4]); g4[0] = pg.x; g4[1] = pg.y; g4[2] = pg.z; g4[3] = pg.w; }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float n = fmaf((f[i4 + i] - mean) * rstd, a4[i], b4[i]);
            if (MIX) n = fmaf(f[i4 + i], g4[i], n);
            o[i4 + i] = n;
          }
        }
        Vec<T>::store(yr + c, o);
      }
    }
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(kFastWarps * 32)
ln_bwd_fast_kernel(const T* __restrict__ g, int64_t ldg, const T* __restrict__ x, int64_t ldx, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ w, const T* __restrict__ dres, int64_t ldr,
                   T* __restrict__ dx, int64_t lddx, int M) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  constexpr int IT = D / (32 * V);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float pw[IT][V];
#pragma unroll
  for (int it = 0; it < IT; ++it)
#pragma unroll
    for (int i = 0; i < V; ++i) pw[it][i] = w[(it * 32 + lane) * V + i];
  for (int row = blockIdx.x * kFastWarps + warp; row < M; row += gridDim.x * kFastWarps) {
    const float mu = mean[row], rs = rstd[row];
    // With gh = g * w:  s1 = mean(gh),  s2 = mean(gh * xhat) = rs * (mean(gh x) - mu s1);
    // dx = dres + rs gh - rs s1 - rs^2 s2 (x - mu)                       (three FMA-class ops per element in the second pass)
    float gh[IT][V], xr[IT][V], o[IT][V];
    float s1 = 0.f, tx = 0.f;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int c = (it * 32 + lane) * V;
      float gv[V];
      Vec<T>::load(g + int64_t(row) * ldg + c, gv);
      Vec<T>::load(x + int64_t(row) * ldx + c, xr[it]);
      if (dres != nullptr) Vec<T>::load(dres + int64_t(row) * ldr + c, o[it]);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) o[it][i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        gh[it][i] = gv[i] * pw[it][i];
        s1 += gh[it][i];
        tx = fmaf(gh[it][i], xr[it][i], tx);
      }
    }
    s1 = warp_sum(s1) * (1.0f / D);
    tx = warp_sum(tx) * (1.0f / D);
    const float s2 = rs * (tx - mu * s1);
    const float cx = rs * rs * s2;
    const float c0 = fmaf(cx, mu, -rs * s1);
#pragma unroll
    for (int it = 0; it < IT; ++it) {
#pragma unroll
      for (int i = 0; i < V; ++i) o[it][i] = fmaf(-cx, xr[it][i], fmaf(gh[it][i], rs, o[it][i])) + c0;
      Vec<T>::store(dx + int64_t(row) * lddx + (it * 32 + lane) * V, o[it]);
    }
  }
}

inline int fast_grid(int M, int rows_per_warp = 1) {
  int g = (M + kFastWarps * rows_per_warp - 1) / (kFastWarps * rows_per_warp);
  const int cap = sm_count() * 6;
  return g > cap ? cap : (g < 1 ? 1 : g);
}

template <typename T, int D>
int ln_fwd_fast(const ngu_ln_desc& d, cudaStream_t s) {
  const T* x = reinterpret_cast<const T*>(d.x);
  T* y = reinterpret_cast<T*>(d.y);
  if (d.gamma != nullptr)
    launch_pdl(ln_fwd_fast_kernel<T, D, true>, dim3(fast_grid(d.M, 2)), dim3(kFastWarps * 32), size_t(0), s, x, d.ldx, d.w, d.b, d.gamma, d.gammax, y, d.ldy, d.mean, d.rstd, d.M, d.eps);
  else
    launch_pdl(ln_fwd_fast_kernel<T, D, false>, dim3(fast_grid(d.M, 2)), dim3(kFastWarps * 32), size_t(0), s, x, d.ldx, d.w, d.b, nullptr, nullptr, y, d.ldy, d.mean, d.rstd, d.M, d.eps);
  return check_launch("ln_fwd");
}
template <typename T, int D>
int ln_bwd_fast(const ngu_ln_bwd_desc& d, cudaStream_t s) {
  launch_pdl(ln_bwd_fast_kernel<T, D>, dim3(fast_grid(d.M)), dim3(kFastWarps * 32), size_t(0), s, 
      reinterpret_cast<const T*>(d.g), d.ldg, reinterpret_cast<const T*>(d.x), d.ldx, d.mean, d.rstd, d.w,
      reinterpret_cast<const T*>(d.dres), d.ldr, reinterpret_cast<T*>(d.dx), d.lddx, d.M);
  return check_launch("ln_bwd");
}

template <typename T>
int ln_fwd_t(const ngu_ln_desc& d, cudaStream_t s) {
  const int grid = (d.M + kLnWarps - 1) / kLnWarps;
  launch_pdl(ln_fwd_kernel<T>, dim3(grid), dim3(kLnWarps * 32), size_t(0), s, reinterpret_cast<const T*>(d.x), d.ldx, d.w, d.b, d.gamma, d.gammax,
                                                   reinterpret_cast<T*>(d.y), d.ldy, d.mean, d.rstd, d.M, d.D, d.eps);
  return check_launch("ln_fwd");
}
template <typename T>
int ln_bwd_t(const ngu_ln_bwd_desc& d, cudaStream_t s) {
  const int grid = (d.M + kLnWarps - 1) / kLnWarps;
  launch_pdl(ln_bwd_kernel<T>, dim3(grid), dim3(kLnWarps * 32), size_t(0), s, reinterpret_cast<const T*>(d.g), d.ldg, reinterpret_cast<const T*>(d.x), d.ldx,
                                                   d.mean, d.rstd, d.w, reinterpret_cast<const T*>(d.dres), d.ldr,
                                                   reinterpret_cast<T*>(d.dx), d.lddx, d.M, d.D);
  return check_launch("ln_bwd");
}
template <typename T, int D>
int mona_pre_bwd_t(const ngu_mona_pre_bwd_desc& d, cudaStream_t s) {
  constexpr int R = D / 64;
  const int smem = (5 * D + 2 * R) * int(sizeof(float)) + 3 * R * D * int(sizeof(T));
  cudaError_t e = cudaFuncSetAttribute(mona_pre_bwd_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_status(e, "mona_pre_bwd attr");
  const int ntiles = (d.M + R - 1) / R;
  int grid = sm_count() * (sizeof(T) == 2 ? 2 : 1);
  if (grid > ntiles) grid = ntiles;
  launch_pdl(mona_pre_bwd_kernel<T, D>, dim3(grid), dim3(D / 2), size_t(smem), s, 
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
  if (d.D == 768) return d.dtype == NGU_F32 ? ln_fwd_fast<float, 768>(d, s) : ln_fwd_fast<bf16, 768>(d, s);
  if (d.D == 1024) return d.dtype == NGU_F32 ? ln_fwd_fast<float, 1024>(d, s) : ln_fwd_fast<bf16, 1024>(d, s);
  return d.dtype == NGU_F32 ? ln_fwd_t<float>(d, s) : ln_fwd_t<bf16>(d, s);
}
int ln_bwd(const ngu_ln_bwd_desc& d, cudaStream_t s) {
  if (int rc = check_ln_shape(d.M, d.D, d.dtype, "ln_bwd")) return rc;
  const int V = d.dtype == NGU_F32 ? 4 : 8;
  if ((d.ldx % V) || (d.ldg % V) || (d.lddx % V) || (d.dres && (d.ldr % V))) { set_last_error("ln_bwd: row strides must keep 16-byte alignment"); return NGU_ERR_ALIGN; }
  if (d.D == 768) return d.dtype == NGU_F32 ? ln_bwd_fast<float, 768>(d, s) : ln_bwd_fast<bf16, 768>(d, s);
  if (d.D == 1024) return d.dtype == NGU_F32 ? ln_bwd_fast<float, 1024>(d, s) : ln_bwd_fast<bf16, 1024>(d, s);
  return d.dtype == NGU_F32 ? ln_bwd_t<float>(d, s) : ln_bwd_t<bf16>(d, s);
}
int mona_pre_bwd(const ngu_mona_pre_bwd_desc& d, cudaStream_t s) {
  if (d.M <= 0) { set_last_error("mona_pre_bwd: empty"); return NGU_ERR_SHAPE; }
  if (d.D == 768) return d.dtype == NGU_F32 ? mona_pre_bwd_t<float, 768>(d, s) : mona_pre_bwd_t<bf16, 768>(d, s);
  if (d.D == 1024) return d.dtype == NGU_F32 ? mona_pre_bwd_t<float, 1024>(d, s) : mona_pre_bwd_t<bf16, 1024>(d, s);
  if (d.D == 256) return d.dtype == NGU_F32 ? mona_pre_bwd_t<float, 256>(d, s) : mona_pre_bwd_t<bf16, 256>(d, s);
  if (d.D == 512) return d.dtype == NGU_F32 ? mona_pre_bwd_t<float, 512>(d, s) : mona_pre_bwd_t<bf16, 512>(d, s);
  set_last_error("mona_pre_bwd: embed dim %d not instantiated (256, 512, 768, 1024)", d.D);
  return NGU_ERR_SHAPE;
}

}  // namespace ngu
