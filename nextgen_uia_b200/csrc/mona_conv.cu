// Mona bottleneck stage, one CTA per image, everything between project1 and project2 in shared
// memory (src/adapters/mona.py:85-93 BaselineMonaOp and :129-147 of BaselineMona.forward):
//     z   = (dw3(h) + dw5(h) + dw7(h)) / 3 + h          depthwise, zero padded, per-branch bias
//     a   = z + P z + bp                                 1x1 projector + residual
//     g   = dropout(gelu(a))                             CLS token (if any) skips the conv: a = h
// The three depthwise kernels are merged into ONE 7x7 stencil (exact: (pad(k3)+pad(k5)+k7)/3 + delta)
// so the stencil reads each neighbour once; backward recomputes z/a from the saved h instead of
// storing them, and produces dh plus all parameter gradients (fp32 atomics, one set per CTA).
//
// Layout: h/g/dh are [B, N, C] token-major (N = has_cls + H*W), C = 64 channels contiguous, so a
// warp reads 32 consecutive channels of one token (coalesced, bank-conflict free in smem).
#include "common.cuh"
#include "kernels.h"
#include "mona_stage.cuh"

namespace ngu {
#ifdef NGU_CONV_PROF
// debug build only: per-phase clock64() stamps of CTA 0 (tools/gpu_conv_phases.py)
__device__ long long g_conv_prof[32];
#define NGU_PROF(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_conv_prof[i] = clock64(); } while (0)
#else
#define NGU_PROF(i) do { } while (0)
#endif
namespace {

using mona_stage::C;           // bottleneck channels (reference default --mona_bottleneck 64)
using namespace mona_stage;
constexpr int kThreads = 256;
constexpr int kChunk = 32;   // tokens per projector chunk

template <typename T> NGU_DEVINL float ldT(const T* p) { return to_f32<T>(*p); }

struct ConvSmem {
  float kc[49][C];     // merged stencil, tap-major
  float bc[C];         // merged bias
  float pt[C][C];      // projector weight transposed: pt[i][o] = P[o][i]
  float bp[C];
  float zc[kChunk][C]; // chunk of z
  float dac[kChunk][C];// chunk of da (backward)
};

NGU_DEVINL void load_weights(ConvSmem& s, const ngu_mona_conv_weights& w) {
  for (int i = threadIdx.x; i < 49 * C; i += kThreads) {
    const int c = i % C, t = i / C;
    const int ky = t / 7, kx = t % 7;
    float v = w.k7[c * 49 + t];
    if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) v += w.k5[c * 25 + (ky - 1) * 5 + (kx - 1)];
    if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) v += w.k3[c * 9 + (ky - 2) * 3 + (kx - 2)];
    v *= (1.0f / 3.0f);
    if (t == 24) v += 1.0f;
    s.kc[t][c] = v;
  }
  for (int i = threadIdx.x; i < C * C; i += kThreads) {
    const int o = i / C, ii = i % C;
    s.pt[ii][o] = w.P[i];
  }
  if (threadIdx.x < C) {
    s.bc[threadIdx.x] = (w.b3[threadIdx.x] + w.b5[threadIdx.x] + w.b7[threadIdx.x]) * (1.0f / 3.0f);
    s.bp[threadIdx.x] = w.bp[threadIdx.x];
  }
}

// z for tokens [p0, p0+kChunk) of the H x W grid from the full h tile (hs: [HW][C] of T)
template <typename T>
NGU_DEVINL void conv_chunk(ConvSmem& s, const T* hs, int p0, int H, int W) {
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  for (int lp = grp; lp < kChunk; lp += kThreads / C) {
    const int p = p0 + lp;
    if (p >= H * W) { s.zc[lp][c] = 0.f; continue; }
    const int y = p / W, x = p % W;
    float acc = s.bc[c];
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
      const int yy = y + ky - 3;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const int xx = x + kx - 3;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(s.kc[ky * 7 + kx][c], to_f32<T>(hs[(yy * W + xx) * C + c]), acc);
      }
    }
    s.zc[lp][c] = acc;
  }
}

// a[lp][o] = z + bp + sum_i P[o][i] z[lp][i]
NGU_DEVINL float proj_out(const ConvSmem& s, int lp, int o) {
  float acc = s.zc[lp][o] + s.bp[o];
#pragma unroll 16
  for (int i = 0; i < C; ++i) acc = fmaf(s.pt[i][o], s.zc[lp][i], acc);
  return acc;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
mona_conv_fwd_kernel(const T* __restrict__ h, T* __restrict__ g, ngu_mona_conv_weights w, int N, int H, int W,
                     int has_cls, float drop_p, uint64_t seed, const uint64_t* seed_ctr) {
  pdl_prologue();
  seed = mix_seed(seed, seed_ctr);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  ConvSmem& s = *reinterpret_cast<ConvSmem*>(smem_dyn);
  T* hs = reinterpret_cast<T*>(smem_dyn + sizeof(ConvSmem));
  const int img = blockIdx.x;
  const T* hb = h + size_t(img) * N * C;
  T* gb = g + size_t(img) * N * C;
  const int HW = H * W;
  load_weights(s, w);
  {
    constexpr int V = Vec<T>::N;
    const T* src = hb + has_cls * C;
    for (int i = threadIdx.x * V; i < HW * C; i += kThreads * V) {
      *reinterpret_cast<uint4*>(hs + i) = *reinterpret_cast<const uint4*>(src + i);
    }
  }
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  if (has_cls && grp == 0) {
    const uint64_t idx = (uint64_t(img) * N) * C + c;
    float v = gelu_t<T>(ldT(hb + c));
    if (drop_p > 0.f) v *= dropout_scale(seed, idx, drop_p);
    gb[c] = from_f32<T>(v);
  }
  __syncthreads();
  for (int p0 = 0; p0 < HW; p0 += kChunk) {
    conv_chunk<T>(s, hs, p0, H, W);
    __syncthreads();
    for (int lp = grp; lp < kChunk; lp += kThreads / C) {
      const int p = p0 + lp;
      if (p >= HW) break;
      float v = gelu_t<T>(proj_out(s, lp, c));
      const uint64_t tok = uint64_t(img) * N + has_cls + p;
      if (drop_p > 0.f) v *= dropout_scale(seed, tok * C + c, drop_p);
      gb[(has_cls + p) * C + c] = from_f32<T>(v);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
mona_conv_bwd_kernel(const T* __restrict__ h, const T* __restrict__ dg, T* __restrict__ dh, ngu_mona_conv_weights w,
                     ngu_mona_conv_grads gr, int N, int H, int W, int has_cls, float drop_p, uint64_t seed, const uint64_t* seed_ctr) {
  pdl_prologue();
  seed = mix_seed(seed, seed_ctr);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  ConvSmem& s = *reinterpret_cast<ConvSmem*>(smem_dyn);
  const int HW = H * W;
  T* hs = reinterpret_cast<T*>(smem_dyn + sizeof(ConvSmem));
  T* dzs = hs + size_t(HW) * C;
  const int img = blockIdx.x;
  const T* hb = h + size_t(img) * N * C;
  const T* dgb = dg + size_t(img) * N * C;
  T* dhb = dh + size_t(img) * N * C;
  load_weights(s, w);
  {
    constexpr int V = Vec<T>::N;
    const T* src = hb + has_cls * C;
    for (int i = threadIdx.x * V; i < HW * C; i += kThreads * V)
      *reinterpret_cast<uint4*>(hs + i) = *reinterpret_cast<const uint4*>(src + i);
  }
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  float db1_acc = 0.f;  // column sum of dh (= d project1.bias), channel c, this thread's tokens
  if (has_cls && grp == 0) {
    const uint64_t idx = (uint64_t(img) * N) * C + c;
    float v = ldT(dgb + c) * gelu_grad_t<T>(ldT(hb + c));
    if (drop_p > 0.f) v *= dropout_scale(seed, idx, drop_p);
    dhb[c] = from_f32<T>(v);
    db1_acc += v;
  }
  __syncthreads();

  // dP partials: thread owns P[o][i] for o = c, i in {grp*16 .. grp*16+15}
  float dP[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dP[i] = 0.f;
  float dbp_acc = 0.f;

  for (int p0 = 0; p0 < HW; p0 += kChunk) {
    conv_chunk<T>(s, hs, p0, H, W);
    __syncthreads();
    // da = dg * mask * gelu'(a)
    for (int lp = grp; lp < kChunk; lp += kThreads / C) {
      const int p = p0 + lp;
      float v = 0.f;
      if (p < HW) {
        const float a = proj_out(s, lp, c);
        v = ldT(dgb + (has_cls + p) * C + c) * gelu_grad_t<T>(a);
        const uint64_t tok = uint64_t(img) * N + has_cls + p;
        if (drop_p > 0.f) v *= dropout_scale(seed, tok * C + c, drop_p);
      }
      s.dac[lp][c] = v;
    }
    __syncthreads();
    // dP[o=c][i] += sum_lp da[lp][o] * z[lp][i];  dbp[o] += sum_lp da[lp][o]
#pragma unroll 4
    for (int lp = 0; lp < kChunk; ++lp) {
      const float d = s.dac[lp][c];
      if (grp == 0) dbp_acc += d;
#pragma unroll
      for (int i = 0; i < 16; ++i) dP[i] = fmaf(d, s.zc[lp][grp * 16 + i], dP[i]);
    }
    // dz[lp][i=c] = da[lp][i] + sum_o P[o][i] da[lp][o]   -> full-size dz tile
    for (int lp = grp; lp < kChunk; lp += kThreads / C) {
      const int p = p0 + lp;
      if (p >= HW) break;
      float acc = s.dac[lp][c];
#pragma unroll 16
      for (int o = 0; o < C; ++o) acc = fmaf(__ldg(w.P + o * C + c), s.dac[lp][o], acc);
      dzs[p * C + c] = from_f32<T>(acc);
    }
    __syncthreads();
  }

  // dh = correlate(dz, flipped stencil):  dh[p] = sum_t kc[t] * dz[p - off_t]
  for (int p = grp; p < HW; p += kThreads / C) {
    const int y = p / W, x = p % W;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
      const int yy = y - (ky - 3);
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const int xx = x - (kx - 3);
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(s.kc[ky * 7 + kx][c], to_f32<T>(dzs[(yy * W + xx) * C + c]), acc);
      }
    }
    dhb[(has_cls + p) * C + c] = from_f32<T>(acc);
    db1_acc += acc;
  }

  // stencil weight grads: dkc[t][c] = sum_p dz[p][c] * h[p + off_t][c]; taps split over the 4 groups
  float dbc_acc = 0.f;
  if (grp == 0)
    for (int p = 0; p < HW; ++p) dbc_acc += to_f32<T>(dzs[p * C + c]);
  for (int t = grp; t < 49; t += kThreads / C) {
    const int ky = t / 7, kx = t % 7;
    float acc = 0.f;
    const int y0 = max(0, 3 - ky), y1 = min(H, H + 3 - ky);
    const int x0 = max(0, 3 - kx), x1 = min(W, W + 3 - kx);
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x)
        acc = fmaf(to_f32<T>(dzs[(y * W + x) * C + c]), to_f32<T>(hs[((y + ky - 3) * W + (x + kx - 3)) * C + c]), acc);
    acc *= (1.0f / 3.0f);
    atomicAdd(gr.dk7 + c * 49 + t, acc);
    if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) atomicAdd(gr.dk5 + c * 25 + (ky - 1) * 5 + (kx - 1), acc);
    if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) atomicAdd(gr.dk3 + c * 9 + (ky - 2) * 3 + (kx - 2), acc);
  }
  if (grp == 0) {
    const float b = dbc_acc * (1.0f / 3.0f);
    atomicAdd(gr.db3 + c, b);
    atomicAdd(gr.db5 + c, b);
    atomicAdd(gr.db7 + c, b);
    atomicAdd(gr.dbp + c, dbp_acc);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) atomicAdd(gr.dP + c * C + grp * 16 + i, dP[i]);
  atomicAdd(gr.db1 + c, db1_acc);
}

// =====================================================================================================
// Register-blocked full-tile kernels (grids up to 16x16 tokens, e.g. ViT-B/16 @ 224 -> 14x14).
// Same math as the generic kernels above; differences are purely about shared-memory traffic:
//   * stencil taps (49) and one projector row/column (64) live in REGISTERS of the thread that owns the channel,
//   * the stencil is evaluated on 8-wide output strips with a 14-wide sliding window (13/49 loads per FMA),
//   * the 1x1 projector reads the token's 64 inputs as 128-bit shared-memory broadcasts.
// Intermediates z / da / dz stay in smem in the activation dtype (two CTAs per SM in bf16).
// =====================================================================================================
constexpr int kStrip = 8;

constexpr int CH = C / 4;  // noise-estimator hidden width (reference: in_features // 4)

struct FastSmem {
  float kc[49][C];   // effective merged stencil: f_c * (w1 k3 + w2 k5 + w3 k7) + delta
  float bc[C];       // w1 b3 + w2 b5 + w3 b7
  float pt[C][C];    // pt[i][o] = P[o][i]
  float bp[C];
  // variant state (baseline: f = 1, w = 1/3)
  float fr[C];       // freq_filter
  float hm[C];       // mean over positions of h (per channel)
  float part[4][C];  // scratch for the per-channel reductions
  float hid[CH];     // ReLU(W1 gap + b1)
  float wbr[4];      // branch weights w1, w2, w3 (softmax) [+ pad]
  float dwbr[4];     // backward: d loss / d w_i
  float dgap[C];     // backward: d loss / d gap_c
};

// Loads the image tile and the projector, evaluates the variant prologue (frequency scale, noise-estimator branch
// weights) and builds the effective merged stencil.  Ends with a __syncthreads().
template <typename T>
NGU_DEVINL void fast_load_common(FastSmem& s, T* hs, const T* hb, const ngu_mona_conv_weights& w, int HW, int has_cls) {
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  for (int i = threadIdx.x; i < C * C; i += kThreads) s.pt[i % C][i / C] = w.P[i];
  if (threadIdx.x < C) {
    s.bp[threadIdx.x] = w.bp[threadIdx.x];
    s.fr[threadIdx.x] = w.freq ? w.freq[threadIdx.x] : 1.0f;
  }
  constexpr int V = Vec<T>::N;
  const T* src = hb + has_cls * C;
  for (int i = threadIdx.x * V; i < HW * C; i += kThreads * V) *reinterpret_cast<uint4*>(hs + i) = *reinterpret_cast<const uint4*>(src + i);
  __syncthreads();
  if (w.ne_w1 != nullptr) {
    // noise estimator: gap_c = f_c * mean_p h[p][c]  ->  hid = relu(W1 gap + b1)  ->  w = softmax(W2 hid + b2)
    float a = 0.f;
    for (int p = grp; p < HW; p += kThreads / C) a += to_f32<T>(hs[p * C + c]);
    s.part[grp][c] = a;
    __syncthreads();
    if (threadIdx.x < C) s.hm[c] = (s.part[0][c] + s.part[1][c] + s.part[2][c] + s.part[3][c]) / float(HW);
    __syncthreads();
    if (threadIdx.x < CH) {
      float acc = w.ne_b1[threadIdx.x];
      for (int i = 0; i < C; ++i) acc = fmaf(w.ne_w1[threadIdx.x * C + i], s.fr[i] * s.hm[i], acc);
      s.hid[threadIdx.x] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float lg[3];
      for (int i = 0; i < 3; ++i) {
        float acc = w.ne_b2[i];
        for (int j = 0; j < CH; ++j) acc = fmaf(w.ne_w2[i * CH + j], s.hid[j], acc);
        lg[i] = acc;
      }
      const float mx = fmaxf(lg[0], fmaxf(lg[1], lg[2]));
      const float e0 = expf(lg[0] - mx), e1 = expf(lg[1] - mx), e2 = expf(lg[2] - mx);
      const float inv = 1.f / (e0 + e1 + e2);
      s.wbr[0] = e0 * inv; s.wbr[1] = e1 * inv; s.wbr[2] = e2 * inv;
    }
  } else if (threadIdx.x == 0) {
    s.wbr[0] = s.wbr[1] = s.wbr[2] = 1.0f / 3.0f;
  }
  __syncthreads();
  const float w1 = s.wbr[0], w2 = s.wbr[1], w3 = s.wbr[2];
  for (int i = threadIdx.x; i < 49 * C; i += kThreads) {
    const int cc = i % C, t = i / C;
    const int ky = t / 7, kx = t % 7;
    float v = w3 * w.k7[cc * 49 + t];
    if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) v = fmaf(w2, w.k5[cc * 25 + (ky - 1) * 5 + (kx - 1)], v);
    if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) v = fmaf(w1, w.k3[cc * 9 + (ky - 2) * 3 + (kx - 2)], v);
    v *= s.fr[cc];
    if (t == 24) v += 1.0f;
    s.kc[t][cc] = v;
  }
  if (threadIdx.x < C) s.bc[threadIdx.x] = w1 * w.b3[threadIdx.x] + w2 * w.b5[threadIdx.x] + w3 * w.b7[threadIdx.x];
  __syncthreads();
}

// out[y][x0 + j] = bias + sum_{ky,kx} k[ky*7 + (FLIP ? 6-kx : kx)] * in[y + (FLIP ? 3-ky : ky-3)][x0 + j + kx - 3]   (channel c)
template <typename T, bool FLIP>
NGU_DEVINL void stencil_strip(const T* in, const float (&k)[49], float bias, int y, int x0, int H, int W, int c, float (&acc)[kStrip]) {
#pragma unroll
  for (int j = 0; j < kStrip; ++j) acc[j] = bias;
#pragma unroll
  for (int ky = 0; ky < 7; ++ky) {
    const int yy = FLIP ? y + 3 - ky : y + ky - 3;
    if (yy < 0 || yy >= H) continue;
    float win[kStrip + 6];
#pragma unroll
    for (int i = 0; i < kStrip + 6; ++i) {
      const int xx = x0 + i - 3;
      win[i] = (xx >= 0 && xx < W) ? to_f32<T>(in[(yy * W + xx) * C + c]) : 0.f;
    }
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const float kv = k[ky * 7 + (FLIP ? 6 - kx : kx)];
#pragma unroll
      for (int j = 0; j < kStrip; ++j) acc[j] = fmaf(kv, win[j + kx], acc[j]);
    }
  }
}

// a = z[p][o] + bp[o] + sum_i pt[i][o] z[p][i], z row read as 128-bit broadcasts
template <typename T>
NGU_DEVINL float proj_token(const T* zrow, const float (&pcol)[C], float bp, int o) {
  float acc[4] = {to_f32<T>(zrow[o]) + bp, 0.f, 0.f, 0.f};  // 4 independent FMA chains
  constexpr int V = Vec<T>::N;
#pragma unroll
  for (int i0 = 0; i0 < C; i0 += V) {
    float zv[V];
    Vec<T>::load(zrow + i0, zv);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i & 3] = fmaf(pcol[i0 + i], zv[i], acc[i & 3]);
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
mona_conv_fwd_fast_kernel(const T* __restrict__ h, T* __restrict__ g, ngu_mona_conv_weights w, int N, int H, int W,
                          int has_cls, float drop_p, uint64_t seed, const uint64_t* seed_ctr) {
  pdl_prologue();
  seed = mix_seed(seed, seed_ctr);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  FastSmem& s = *reinterpret_cast<FastSmem*>(smem_dyn);
  const int HW = H * W;
  T* hs = reinterpret_cast<T*>(smem_dyn + sizeof(FastSmem));
  T* zs = hs + HW * C;
  const int img = blockIdx.x;
  const T* hb = h + size_t(img) * N * C;
  T* gb = g + size_t(img) * N * C;
  fast_load_common<T>(s, hs, hb, w, HW, has_cls);
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  if (has_cls && grp == 0) {
    float v = gelu_t<T>(to_f32<T>(hb[c]));
    if (drop_p > 0.f) v *= dropout_scale(seed, (uint64_t(img) * N) * C + c, drop_p);
    gb[c] = from_f32<T>(v);
  }
  __syncthreads();
  {
    float k[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) k[t] = s.kc[t][c];
    const float bias = s.bc[c];
    const int spr = (W + kStrip - 1) / kStrip;
    for (int st = grp; st < H * spr; st += kThreads / C) {
      const int y = st / spr, x0 = (st % spr) * kStrip;
      float acc[kStrip];
      stencil_strip<T, false>(hs, k, bias, y, x0, H, W, c, acc);
#pragma unroll
      for (int j = 0; j < kStrip; ++j)
        if (x0 + j < W) zs[(y * W + x0 + j) * C + c] = from_f32<T>(acc[j]);
    }
  }
  __syncthreads();
  {
    float pcol[C];
#pragma unroll
    for (int i = 0; i < C; ++i) pcol[i] = s.pt[i][c];
    const float bp = s.bp[c];
    for (int p = grp; p < HW; p += kThreads / C) {
      float v = gelu_t<T>(proj_token<T>(zs + p * C, pcol, bp, c));
      const uint64_t tok = uint64_t(img) * N + has_cls + p;
      if (drop_p > 0.f) v *= dropout_scale(seed, tok * C + c, drop_p);
      gb[(has_cls + p) * C + c] = from_f32<T>(v);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
mona_conv_bwd_fast_kernel(const T* __restrict__ h, const T* __restrict__ dg, T* __restrict__ dh, ngu_mona_conv_weights w,
                          ngu_mona_conv_grads gr, int N, int H, int W, int has_cls, float drop_p, uint64_t seed, const uint64_t* seed_ctr) {
  pdl_prologue();
  seed = mix_seed(seed, seed_ctr);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  FastSmem& s = *reinterpret_cast<FastSmem*>(smem_dyn);
  const int HW = H * W;
  T* hs = reinterpret_cast<T*>(smem_dyn + sizeof(FastSmem));
  T* zs = hs + HW * C;    // z, later dz
  T* das = zs + HW * C;   // da
  const int img = blockIdx.x;
  const T* hb = h + size_t(img) * N * C;
  const T* dgb = dg + size_t(img) * N * C;
  T* dhb = dh + size_t(img) * N * C;
  fast_load_common<T>(s, hs, hb, w, HW, has_cls);
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  float db1_acc = 0.f;
  if (has_cls && grp == 0) {
    float v = to_f32<T>(dgb[c]) * gelu_grad_t<T>(to_f32<T>(hb[c]));
    if (drop_p > 0.f) v *= dropout_scale(seed, (uint64_t(img) * N) * C + c, drop_p);
    dhb[c] = from_f32<T>(v);
    db1_acc += v;
  }
  __syncthreads();
  const int spr = (W + kStrip - 1) / kStrip;
  // ---- phase 1: z = stencil(h)
  {
    float k[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) k[t] = s.kc[t][c];
    const float bias = s.bc[c];
    for (int st = grp; st < H * spr; st += kThreads / C) {
      const int y = st / spr, x0 = (st % spr) * kStrip;
      float acc[kStrip];
      stencil_strip<T, false>(hs, k, bias, y, x0, H, W, c, acc);
#pragma unroll
      for (int j = 0; j < kStrip; ++j)
        if (x0 + j < W) zs[(y * W + x0 + j) * C + c] = from_f32<T>(acc[j]);
    }
  }
  __syncthreads();
  // ---- phase 2: da = dg * mask * gelu'(z + P z + bp)
  {
    float pcol[C];
#pragma unroll
    for (int i = 0; i < C; ++i) pcol[i] = s.pt[i][c];
    const float bp = s.bp[c];
    for (int p = grp; p < HW; p += kThreads / C) {
      const float a = proj_token<T>(zs + p * C, pcol, bp, c);
      float v = to_f32<T>(dgb[(has_cls + p) * C + c]) * gelu_grad_t<T>(a);
      const uint64_t tok = uint64_t(img) * N + has_cls + p;
      if (drop_p > 0.f) v *= dropout_scale(seed, tok * C + c, drop_p);
      das[p * C + c] = from_f32<T>(v);
    }
  }
  __syncthreads();
  // ---- phase 3: dP[o = c][i in 16*grp..+15] = sum_p da[p][o] z[p][i];  dbp[o] = sum_p da[p][o]
  {
    float dP[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dP[i] = 0.f;
    float dbp_acc = 0.f;
    constexpr int V = Vec<T>::N;
    for (int p = 0; p < HW; ++p) {
      const float d = to_f32<T>(das[p * C + c]);
      dbp_acc += d;
#pragma unroll
      for (int i0 = 0; i0 < 16; i0 += V) {
        float zv[V];
        Vec<T>::load(zs + p * C + grp * 16 + i0, zv);
#pragma unroll
        for (int i = 0; i < V; ++i) dP[i0 + i] = fmaf(d, zv[i], dP[i0 + i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) atomicAdd(gr.dP + c * C + grp * 16 + i, dP[i]);
    if (grp == 0) atomicAdd(gr.dbp + c, dbp_acc);
  }
  __syncthreads();
  // ---- phase 4: dz[p][i = c] = da[p][i] + sum_o P[o][i] da[p][o]   (overwrites z)
  {
    float prow[C];
#pragma unroll
    for (int o = 0; o < C; ++o) prow[o] = s.pt[c][o];
    constexpr int V = Vec<T>::N;
    for (int p = grp; p < HW; p += kThreads / C) {
      float acc[4] = {to_f32<T>(das[p * C + c]), 0.f, 0.f, 0.f};
#pragma unroll
      for (int o0 = 0; o0 < C; o0 += V) {
        float dv[V];
        Vec<T>::load(das + p * C + o0, dv);
#pragma unroll
        for (int o = 0; o < V; ++o) acc[o & 3] = fmaf(prow[o0 + o], dv[o], acc[o & 3]);
      }
      zs[p * C + c] = from_f32<T>((acc[0] + acc[1]) + (acc[2] + acc[3]));
    }
  }
  __syncthreads();
  // ---- phase 5: correlation sums G[t][c] = sum_p dz[p][c] h[p + off_t][c]  and  S[c] = sum_p dz[p][c]
  //      (z = f_c * sum_t kw[c][t] h[p+off_t] + h + bias, kw = sum_i w_i k_i  =>  every stencil-side gradient is a
  //      contraction of G / S); accumulated across the four position groups with shared-memory atomics.
  float* Gs = reinterpret_cast<float*>(das);          // [49][C] floats (12.5 KB) — da is dead after phase 4
  float* Ss = s.part[0];                               // [C]
  for (int i = threadIdx.x; i < 49 * C; i += kThreads) Gs[i] = 0.f;
  if (threadIdx.x < C) Ss[threadIdx.x] = 0.f;
  __syncthreads();
  {
    float dk[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) dk[t] = 0.f;
    float dbc_acc = 0.f;
    for (int st = grp; st < H * spr; st += kThreads / C) {
      const int y = st / spr, x0 = (st % spr) * kStrip;
      float dzv[kStrip];
#pragma unroll
      for (int j = 0; j < kStrip; ++j) {
        dzv[j] = (x0 + j < W) ? to_f32<T>(zs[(y * W + x0 + j) * C + c]) : 0.f;
        dbc_acc += dzv[j];
      }
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + ky - 3;
        if (yy < 0 || yy >= H) continue;
        float win[kStrip + 6];
#pragma unroll
        for (int i = 0; i < kStrip + 6; ++i) {
          const int xx = x0 + i - 3;
          win[i] = (xx >= 0 && xx < W) ? to_f32<T>(hs[(yy * W + xx) * C + c]) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          float a = dk[ky * 7 + kx];
#pragma unroll
          for (int j = 0; j < kStrip; ++j) a = fmaf(dzv[j], win[j + kx], a);
          dk[ky * 7 + kx] = a;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 49; ++t) atomicAdd(&Gs[t * C + c], dk[t]);
    atomicAdd(&Ss[c], dbc_acc);
  }
  __syncthreads();
  // ---- phase 6: stencil / bias / frequency / branch-weight gradients from G and S
  const float w1 = s.wbr[0], w2 = s.wbr[1], w3 = s.wbr[2];
  if (threadIdx.x < C) {
    const float f = s.fr[c], Sc = Ss[c];
    float q1 = 0.f, q2 = 0.f, q3 = 0.f;  // q_i = sum_{t in branch i} k_i[c][t] G[t][c]
    for (int t = 0; t < 49; ++t) {
      const int ky = t / 7, kx = t % 7;
      const float g = Gs[t * C + c];
      q3 = fmaf(w.k7[c * 49 + t], g, q3);
      atomicAdd(gr.dk7 + c * 49 + t, w3 * f * g);
      if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) {
        const int u = (ky - 1) * 5 + (kx - 1);
        q2 = fmaf(w.k5[c * 25 + u], g, q2);
        atomicAdd(gr.dk5 + c * 25 + u, w2 * f * g);
      }
      if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) {
        const int u = (ky - 2) * 3 + (kx - 2);
        q1 = fmaf(w.k3[c * 9 + u], g, q1);
        atomicAdd(gr.dk3 + c * 9 + u, w1 * f * g);
      }
    }
    atomicAdd(gr.db3 + c, w1 * Sc);
    atomicAdd(gr.db5 + c, w2 * Sc);
    atomicAdd(gr.db7 + c, w3 * Sc);
    // direct frequency-filter gradient: d z / d f_c = sum_t kw[c][t] h[p + off_t]
    s.part[1][c] = w1 * q1 + w2 * q2 + w3 * q3;
    // branch-weight gradient contributions of this channel
    s.part[2][c] = f * q1 + w.b3[c] * Sc;
    s.part[3][c] = f * q2 + w.b5[c] * Sc;
    s.dgap[c] = f * q3 + w.b7[c] * Sc;  // (temporarily) third branch contribution
  }
  __syncthreads();
  float cst = 0.f;  // constant added to dh at every position of channel c (through the global average pool)
  if (w.ne_w1 != nullptr) {
    if (threadIdx.x < 3) {
      const float* src = threadIdx.x == 0 ? s.part[2] : threadIdx.x == 1 ? s.part[3] : s.dgap;
      float a = 0.f;
      for (int i = 0; i < C; ++i) a += src[i];
      s.dwbr[threadIdx.x] = a;
    }
    __syncthreads();
    // softmax backward -> d logits; then the two 1x1 layers
    const float dot = w1 * s.dwbr[0] + w2 * s.dwbr[1] + w3 * s.dwbr[2];
    const float dl0 = w1 * (s.dwbr[0] - dot), dl1 = w2 * (s.dwbr[1] - dot), dl2 = w3 * (s.dwbr[2] - dot);
    __syncthreads();  // everyone has read dwbr / dgap(temp) before they are overwritten
    if (threadIdx.x < CH) {
      const int j = threadIdx.x;
      const float hj = s.hid[j];
      atomicAdd(gr.dne_w2 + 0 * CH + j, dl0 * hj);
      atomicAdd(gr.dne_w2 + 1 * CH + j, dl1 * hj);
      atomicAdd(gr.dne_w2 + 2 * CH + j, dl2 * hj);
      const float dh_ = hj > 0.f ? (w.ne_w2[0 * CH + j] * dl0 + w.ne_w2[1 * CH + j] * dl1 + w.ne_w2[2 * CH + j] * dl2) : 0.f;
      s.hid[j] = dh_;  // reuse as d hid
      atomicAdd(gr.dne_b1 + j, dh_);
    }
    if (threadIdx.x == 0) { atomicAdd(gr.dne_b2 + 0, dl0); atomicAdd(gr.dne_b2 + 1, dl1); atomicAdd(gr.dne_b2 + 2, dl2); }
    __syncthreads();
    if (threadIdx.x < C) {
      float dg = 0.f;
      const float gap = s.fr[c] * s.hm[c];
      for (int j = 0; j < CH; ++j) {
        dg = fmaf(w.ne_w1[j * C + c], s.hid[j], dg);
        atomicAdd(gr.dne_w1 + j * C + c, s.hid[j] * gap);
      }
      s.dgap[c] = dg;
    }
    __syncthreads();
    cst = s.dgap[c] * s.fr[c] / float(HW);
    if (threadIdx.x < C && gr.dfreq) atomicAdd(gr.dfreq + c, s.part[1][c] + s.dgap[c] * s.hm[c]);
  } else {
    if (threadIdx.x < C && gr.dfreq && w.freq) atomicAdd(gr.dfreq + c, s.part[1][c]);
  }
  // ---- phase 7: dh = transposed effective stencil of dz (+ the pooled-path constant)
  {
    float k[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) k[t] = s.kc[t][c];
    for (int st = grp; st < H * spr; st += kThreads / C) {
      const int y = st / spr, x0 = (st % spr) * kStrip;
      float acc[kStrip];
      stencil_strip<T, true>(zs, k, cst, y, x0, H, W, c, acc);
#pragma unroll
      for (int j = 0; j < kStrip; ++j)
        if (x0 + j < W) { dhb[(has_cls + y * W + x0 + j) * C + c] = from_f32<T>(acc[j]); db1_acc += acc[j]; }
    }
  }
  atomicAdd(gr.db1 + c, db1_acc);
}

// =====================================================================================================
// bf16 product path: same algorithm, with the three projector contractions (z P^T, da^T z, da P) on the tensor
// cores (mma.sync m16n8k16, bf16 in / fp32 accumulate, operands via ldmatrix from 128-byte-swizzled smem tiles).
// The depthwise stencil and its correlation sums stay on CUDA cores (no contraction over channels to exploit).
// Tiles hs / zs / das are [HWp][64] bf16 with HWp = HW rounded up to 16 (pad rows zero); element (p, c) lives at
// p*64 + (((c >> 3) ^ (p & 7)) << 3) + (c & 7)   (16-byte chunks XOR-swizzled by the row, so ldmatrix is conflict free).
// =====================================================================================================
struct TcSmem {
  float kc[49][C];
  float bc[C];
  float bp[C];
  float fr[C];
  float hm[C];
  float part[4][C];
  float hid[CH];
  float wbr[4];
  float dwbr[4];
  float dgap[C];
};

// image tile, projector (bf16, swizzled), variant prologue, effective stencil.  Ends with __syncthreads().
NGU_DEVINL void tc_load_common(TcSmem& s, bf16* hs, bf16* pbs, const bf16* hb, const ngu_mona_conv_weights& w, int HW, int HWp, int has_cls) {
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  for (int i = threadIdx.x; i < C * C; i += kThreads) pbs[swz(i / C, i % C)] = __float2bfloat16_rn(w.P[i]);
  if (threadIdx.x < C) {
    s.bp[threadIdx.x] = w.bp[threadIdx.x];
    s.fr[threadIdx.x] = w.freq ? w.freq[threadIdx.x] : 1.0f;
  }
  const bf16* src = hb + has_cls * C;
  for (int i = threadIdx.x; i < HWp * 8; i += kThreads) {  // one 16-byte chunk per thread-iteration
    const int p = i >> 3, ch = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (p < HW) v = *reinterpret_cast<const uint4*>(src + p * C + ch * 8);
    *reinterpret_cast<uint4*>(hs + p * C + ch * 8) = v;   // h stays LINEAR: only the streaming stencils read it
  }
  __syncthreads();
  if (w.ne_w1 != nullptr) {
    float a = 0.f;
    for (int p = grp; p < HW; p += kThreads / C) a += __bfloat162float(hs[p * C + c]);
    s.part[grp][c] = a;
    __syncthreads();
    if (threadIdx.x < C) s.hm[c] = (s.part[0][c] + s.part[1][c] + s.part[2][c] + s.part[3][c]) / float(HW);
    __syncthreads();
    if (threadIdx.x < CH) {
      float acc = w.ne_b1[threadIdx.x];
      for (int i = 0; i < C; ++i) acc = fmaf(w.ne_w1[threadIdx.x * C + i], s.fr[i] * s.hm[i], acc);
      s.hid[threadIdx.x] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float lg[3];
      for (int i = 0; i < 3; ++i) {
        float acc = w.ne_b2[i];
        for (int j = 0; j < CH; ++j) acc = fmaf(w.ne_w2[i * CH + j], s.hid[j], acc);
        lg[i] = acc;
      }
      const float mx = fmaxf(lg[0], fmaxf(lg[1], lg[2]));
      const float e0 = expf(lg[0] - mx), e1 = expf(lg[1] - mx), e2 = expf(lg[2] - mx);
      const float inv = 1.f / (e0 + e1 + e2);
      s.wbr[0] = e0 * inv; s.wbr[1] = e1 * inv; s.wbr[2] = e2 * inv;
    }
  } else if (threadIdx.x == 0) {
    s.wbr[0] = s.wbr[1] = s.wbr[2] = 1.0f / 3.0f;
  }
  __syncthreads();
  const float w1 = s.wbr[0], w2 = s.wbr[1], w3 = s.wbr[2];
  for (int i = threadIdx.x; i < 49 * C; i += kThreads) {
    const int cc = i % C, t = i / C;
    const int ky = t / 7, kx = t % 7;
    float v = w3 * w.k7[cc * 49 + t];
    if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) v = fmaf(w2, w.k5[cc * 25 + (ky - 1) * 5 + (kx - 1)], v);
    if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) v = fmaf(w1, w.k3[cc * 9 + (ky - 2) * 3 + (kx - 2)], v);
    v *= s.fr[cc];
    if (t == 24) v += 1.0f;
    s.kc[t][cc] = v;
  }
  if (threadIdx.x < C) s.bc[threadIdx.x] = w1 * w.b3[threadIdx.x] + w2 * w.b5[threadIdx.x] + w3 * w.b7[threadIdx.x];
  __syncthreads();
}

// z = stencil(h): h linear, z written into the swizzled tile (ldmatrix operand of the projector MMAs)
NGU_DEVINL void conv_phase_stream(const TcSmem& s, const bf16* hs, bf16* zs, int H, int W, int c, int grp) {
  float k[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) k[t] = s.kc[t][c];
  for (int x0 = grp * kSW; x0 < W; x0 += (kThreads / C) * kSW) {
    stencil_stream<false>(hs, k, s.bc[c], x0, H, W, c, [&](int y, const float (&a)[kSW]) {
#pragma unroll
      for (int j = 0; j < kSW; ++j)
        if (x0 + j < W) zs[swz(y * W + x0 + j, c)] = __float2bfloat16_rn(a[j]);
    });
  }
}

__global__ void __launch_bounds__(kThreads, 2)
mona_conv_fwd_tc_kernel(const bf16* __restrict__ h, bf16* __restrict__ g, ngu_mona_conv_weights w, int N, int H, int W,
                        int has_cls, float drop_p, uint64_t seed, const uint64_t* seed_ctr) {
  pdl_prologue();
  seed = mix_seed(seed, seed_ctr);
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(smem_dyn);
  const int HW = H * W, HWp = (HW + 15) & ~15;
  bf16* pbs = reinterpret_cast<bf16*>(smem_dyn + ((sizeof(TcSmem) + 1023) & ~size_t(1023)));
  bf16* hs = pbs + C * C;
  bf16* zs = hs + HWp * C;
  const int img = blockIdx.x;
  const bf16* hb = h + size_t(img) * N * C;
  bf16* gb = g + size_t(img) * N * C;
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  for (int i = threadIdx.x; i < (HWp - HW) * 8; i += kThreads)  // zero the pad rows of z
    *reinterpret_cast<uint4*>(zs + (HW + (i >> 3)) * C + ((i & 7) << 3)) = make_uint4(0, 0, 0, 0);
  NGU_PROF(16);
  tc_load_common(s, hs, pbs, hb, w, HW, HWp, has_cls);
  NGU_PROF(17);
  if (has_cls && grp == 0) {
    float v = gelu_erf(__bfloat162float(hb[c]));
    if (drop_p > 0.f) v *= dropout_scale(seed, (uint64_t(img) * N) * C + c, drop_p);
    gb[c] = __float2bfloat16_rn(v);
  }
  conv_phase_stream(s, hs, zs, H, W, c, grp);
  __syncthreads();
  NGU_PROF(18);
  // a = z + bp + z P^T  ->  g = dropout(gelu(a))
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  const uint32_t zs_u = smem_u32(zs), pb_u = smem_u32(pbs);
  for (int rt = warp; rt < HWp / 16; rt += kThreads / 32) {
    float acc[8][4];
    proj_mma<false>(acc, zs_u, rt, pb_u, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r = rt * 16 + gq + hh * 8, cc = nt * 8 + 2 * tq;
        if (r < HW) {
          const float2 zz = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(zs + swz(r, cc)));
          float v0 = gelu_erf(acc[nt][2 * hh] + zz.x + s.bp[cc]);
          float v1 = gelu_erf(acc[nt][2 * hh + 1] + zz.y + s.bp[cc + 1]);
          if (drop_p > 0.f) {
            const uint64_t e = (uint64_t(img) * N + has_cls + r) * C + cc;
            v0 *= dropout_scale(seed, e, drop_p);
            v1 *= dropout_scale(seed, e + 1, drop_p);
          }
          *reinterpret_cast<uint32_t*>(gb + (has_cls + r) * C + cc) = pack_bf16x2(v0, v1);
        }
      }
  }
  NGU_PROF(19);
}

__global__ void __launch_bounds__(kThreads, 2)
mona_conv_bwd_tc_kernel(const bf16* __restrict__ h, const bf16* __restrict__ dg, bf16* __restrict__ dh, ngu_mona_conv_weights w,
                        ngu_mona_conv_grads gr, int N, int H, int W, int has_cls, float drop_p, uint64_t seed, const uint64_t* seed_ctr) {
  pdl_prologue();
  seed = mix_seed(seed, seed_ctr);
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(smem_dyn);
  const int HW = H * W, HWp = (HW + 15) & ~15;
  bf16* pbs = reinterpret_cast<bf16*>(smem_dyn + ((sizeof(TcSmem) + 1023) & ~size_t(1023)));
  bf16* hs = pbs + C * C;
  bf16* zs = hs + HWp * C;    // z, later dz
  bf16* das = zs + HWp * C;   // da; later reused as the fp32 [49][C] correlation buffer
  const int img = blockIdx.x;
  const bf16* hb = h + size_t(img) * N * C;
  const bf16* dgb = dg + size_t(img) * N * C;
  bf16* dhb = dh + size_t(img) * N * C;
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  for (int i = threadIdx.x; i < (HWp - HW) * 8; i += kThreads) {
    *reinterpret_cast<uint4*>(zs + (HW + (i >> 3)) * C + ((i & 7) << 3)) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(das + (HW + (i >> 3)) * C + ((i & 7) << 3)) = make_uint4(0, 0, 0, 0);
  }
  NGU_PROF(0);
  tc_load_common(s, hs, pbs, hb, w, HW, HWp, has_cls);
  NGU_PROF(1);
  float db1_acc = 0.f;
  if (has_cls && grp == 0) {
    float v = __bfloat162float(dgb[c]) * gelu_erf_grad(__bfloat162float(hb[c]));
    if (drop_p > 0.f) v *= dropout_scale(seed, (uint64_t(img) * N) * C + c, drop_p);
    dhb[c] = __float2bfloat16_rn(v);
    db1_acc += v;
  }
  // ---- phase 1: z = stencil(h)
  conv_phase_stream(s, hs, zs, H, W, c, grp);
  __syncthreads();
  NGU_PROF(2);
  const uint32_t zs_u = smem_u32(zs), da_u = smem_u32(das), pb_u = smem_u32(pbs);
  // ---- phase 2: da = dg * mask * gelu'(z + bp + z P^T)
  for (int rt = warp; rt < HWp / 16; rt += kThreads / 32) {
    float acc[8][4];
    proj_mma<false>(acc, zs_u, rt, pb_u, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r = rt * 16 + gq + hh * 8, cc = nt * 8 + 2 * tq;
        float v0 = 0.f, v1 = 0.f;
        if (r < HW) {
          const float2 zz = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(zs + swz(r, cc)));
          const float2 gg = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dgb + (has_cls + r) * C + cc));
          v0 = gg.x * gelu_erf_grad(acc[nt][2 * hh] + zz.x + s.bp[cc]);
          v1 = gg.y * gelu_erf_grad(acc[nt][2 * hh + 1] + zz.y + s.bp[cc + 1]);
          if (drop_p > 0.f) {
            const uint64_t e = (uint64_t(img) * N + has_cls + r) * C + cc;
            v0 *= dropout_scale(seed, e, drop_p);
            v1 *= dropout_scale(seed, e + 1, drop_p);
          }
        }
        *reinterpret_cast<uint32_t*>(das + swz(r, cc)) = pack_bf16x2(v0, v1);
      }
  }
  __syncthreads();
  NGU_PROF(3);
  // ---- phase 3: dP[o][i] += sum_p da[p][o] z[p][i]  (A = da^T, B = z, both read transposed from k-major tiles);  dbp
  {
    const int mt = warp & 3, nh = warp >> 2;
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    for (int kt = 0; kt < HWp / 16; ++kt) {
      uint32_t a[4];
      ldsm_x4_t(a, tile_addr(da_u, kt * 16 + (lane >> 4) * 8 + (lane & 7), mt * 2 + ((lane >> 3) & 1)));
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        ldsm_x4_t(b, tile_addr(zs_u, kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), nh * 4 + np * 2 + (lane >> 4)));
        mma16816(acc[2 * np], a, b[0], b[1]);
        mma16816(acc[2 * np + 1], a, b[2], b[3]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int o = mt * 16 + gq + hh * 8, i = (nh * 4 + nt) * 8 + 2 * tq;
        atomicAdd(gr.dP + o * C + i, acc[nt][2 * hh]);
        atomicAdd(gr.dP + o * C + i + 1, acc[nt][2 * hh + 1]);
      }
    // dbp[o] = sum_p da[p][o]
    float a = 0.f;
    for (int p = grp; p < HW; p += kThreads / C) a += __bfloat162float(das[swz(p, c)]);
    atomicAdd(gr.dbp + c, a);
  }
  __syncthreads();
  NGU_PROF(4);
  // ---- phase 4: dz = da + da P   (overwrites z)
  for (int rt = warp; rt < HWp / 16; rt += kThreads / 32) {
    float acc[8][4];
    proj_mma<true>(acc, da_u, rt, pb_u, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r = rt * 16 + gq + hh * 8, cc = nt * 8 + 2 * tq;
        const float2 dd = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(das + swz(r, cc)));
        // dz is only read by the streaming stencils from here on: store it LINEAR
        *reinterpret_cast<uint32_t*>(zs + r * C + cc) = pack_bf16x2(acc[nt][2 * hh] + dd.x, acc[nt][2 * hh + 1] + dd.y);
      }
  }
  __syncthreads();
  NGU_PROF(5);
  // ---- phase 5: correlation sums G[t][c] = sum_p dz[p][c] h[p + off_t][c], S[c] = sum_p dz[p][c]
  //      streamed over the rows of h with the 7 dz rows it pairs with held in a rolling register window
  float* Gs = reinterpret_cast<float*>(das);
  float* Ss = s.part[0];
  for (int i = threadIdx.x; i < 49 * C; i += kThreads) Gs[i] = 0.f;
  if (threadIdx.x < C) Ss[threadIdx.x] = 0.f;
  __syncthreads();
  for (int x0 = grp * kSW; x0 < W; x0 += (kThreads / C) * kSW) {
    float G[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) G[t] = 0.f;
    float dsum = 0.f;
    float dzb[7][kSW];  // slot s_ <-> dz row yy - 3 + s_
    auto load_dz = [&](int y, float (&dst)[kSW]) {
#pragma unroll
      for (int j = 0; j < kSW; ++j) {
        dst[j] = (unsigned(y) < unsigned(H) && x0 + j < W) ? __bfloat162float(zs[(y * W + x0 + j) * C + c]) : 0.f;
        dsum += dst[j];
      }
    };
#pragma unroll
    for (int s_ = 0; s_ < 7; ++s_) load_dz(s_ - 3, dzb[s_]);
    for (int yy = 0; yy < H; ++yy) {
      float win[kSW + 6];
      const bf16* rowp = hs + (yy * W) * C + c;
#pragma unroll
      for (int i = 0; i < kSW + 6; ++i) {
        const int xx = x0 + i - 3;
        win[i] = (unsigned(xx) < unsigned(W)) ? __bfloat162float(rowp[xx * C]) : 0.f;
      }
#pragma unroll
      for (int s_ = 0; s_ < 7; ++s_) {
        const int ky = 6 - s_;  // h row yy = dz row (yy - 3 + s_) + ky - 3
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          float a = G[ky * 7 + kx];
#pragma unroll
          for (int j = 0; j < kSW; ++j) a = fmaf(dzb[s_][j], win[j + kx], a);
          G[ky * 7 + kx] = a;
        }
      }
#pragma unroll
      for (int s_ = 0; s_ < 6; ++s_)
#pragma unroll
        for (int j = 0; j < kSW; ++j) dzb[s_][j] = dzb[s_ + 1][j];
      load_dz(yy + 4, dzb[6]);
    }
#pragma unroll
    for (int t = 0; t < 49; ++t) atomicAdd(&Gs[t * C + c], G[t]);
    atomicAdd(&Ss[c], dsum);
  }
  __syncthreads();
  NGU_PROF(6);
  // ---- phase 6: stencil / bias / frequency / branch-weight gradients from G and S (same as the generic fast kernel)
  const float w1 = s.wbr[0], w2 = s.wbr[1], w3 = s.wbr[2];
  if (threadIdx.x < C) {
    const float f = s.fr[c], Sc = Ss[c];
    float q1 = 0.f, q2 = 0.f, q3 = 0.f;
    for (int t = 0; t < 49; ++t) {
      const int ky = t / 7, kx = t % 7;
      const float g_ = Gs[t * C + c];
      q3 = fmaf(w.k7[c * 49 + t], g_, q3);
      atomicAdd(gr.dk7 + c * 49 + t, w3 * f * g_);
      if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) {
        const int u = (ky - 1) * 5 + (kx - 1);
        q2 = fmaf(w.k5[c * 25 + u], g_, q2);
        atomicAdd(gr.dk5 + c * 25 + u, w2 * f * g_);
      }
      if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) {
        const int u = (ky - 2) * 3 + (kx - 2);
        q1 = fmaf(w.k3[c * 9 + u], g_, q1);
        atomicAdd(gr.dk3 + c * 9 + u, w1 * f * g_);
      }
    }
    atomicAdd(gr.db3 + c, w1 * Sc);
    atomicAdd(gr.db5 + c, w2 * Sc);
    atomicAdd(gr.db7 + c, w3 * Sc);
    s.part[1][c] = w1 * q1 + w2 * q2 + w3 * q3;
    s.part[2][c] = f * q1 + w.b3[c] * Sc;
    s.part[3][c] = f * q2 + w.b5[c] * Sc;
    s.dgap[c] = f * q3 + w.b7[c] * Sc;
  }
  __syncthreads();
  float cst = 0.f;
  if (w.ne_w1 != nullptr) {
    if (threadIdx.x < 3) {
      const float* src = threadIdx.x == 0 ? s.part[2] : threadIdx.x == 1 ? s.part[3] : s.dgap;
      float a = 0.f;
      for (int i = 0; i < C; ++i) a += src[i];
      s.dwbr[threadIdx.x] = a;
    }
    __syncthreads();
    const float dot = w1 * s.dwbr[0] + w2 * s.dwbr[1] + w3 * s.dwbr[2];
    const float dl0 = w1 * (s.dwbr[0] - dot), dl1 = w2 * (s.dwbr[1] - dot), dl2 = w3 * (s.dwbr[2] - dot);
    __syncthreads();
    if (threadIdx.x < CH) {
      const int j = threadIdx.x;
      const float hj = s.hid[j];
      atomicAdd(gr.dne_w2 + 0 * CH + j, dl0 * hj);
      atomicAdd(gr.dne_w2 + 1 * CH + j, dl1 * hj);
      atomicAdd(gr.dne_w2 + 2 * CH + j, dl2 * hj);
      const float dh_ = hj > 0.f ? (w.ne_w2[0 * CH + j] * dl0 + w.ne_w2[1 * CH + j] * dl1 + w.ne_w2[2 * CH + j] * dl2) : 0.f;
      s.hid[j] = dh_;
      atomicAdd(gr.dne_b1 + j, dh_);
    }
    if (threadIdx.x == 0) { atomicAdd(gr.dne_b2 + 0, dl0); atomicAdd(gr.dne_b2 + 1, dl1); atomicAdd(gr.dne_b2 + 2, dl2); }
    __syncthreads();
    if (threadIdx.x < C) {
      float dgv = 0.f;
      const float gap = s.fr[c] * s.hm[c];
      for (int j = 0; j < CH; ++j) {
        dgv = fmaf(w.ne_w1[j * C + c], s.hid[j], dgv);
        atomicAdd(gr.dne_w1 + j * C + c, s.hid[j] * gap);
      }
      s.dgap[c] = dgv;
    }
    __syncthreads();
    cst = s.dgap[c] * s.fr[c] / float(HW);
    if (threadIdx.x < C && gr.dfreq) atomicAdd(gr.dfreq + c, s.part[1][c] + s.dgap[c] * s.hm[c]);
  } else {
    if (threadIdx.x < C && gr.dfreq && w.freq) atomicAdd(gr.dfreq + c, s.part[1][c]);
  }
  // ---- phase 7: dh = transposed effective stencil of dz (+ pooled-path constant)
  NGU_PROF(7);
  {
    float k[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) k[t] = s.kc[t][c];
    for (int x0 = grp * kSW; x0 < W; x0 += (kThreads / C) * kSW) {
      stencil_stream<true>(zs, k, cst, x0, H, W, c, [&](int y, const float (&a)[kSW]) {
#pragma unroll
        for (int j = 0; j < kSW; ++j)
          if (x0 + j < W) { dhb[(has_cls + y * W + x0 + j) * C + c] = __float2bfloat16_rn(a[j]); db1_acc += a[j]; }
      });
    }
  }
  atomicAdd(gr.db1 + c, db1_acc);
  NGU_PROF(8);
}

size_t tc_smem_bytes(int HW, bool bwd) {
  const size_t HWp = (size_t(HW) + 15) & ~size_t(15);
  const size_t tile = HWp * C * 2;
  const size_t da = tile > size_t(49) * C * sizeof(float) ? tile : size_t(49) * C * sizeof(float);
  return ((sizeof(TcSmem) + 1023) & ~size_t(1023)) + size_t(C) * C * 2 + 2 * tile + (bwd ? da : 0);
}

// h tile + z tile (+ da tile in backward; the da region is reused as the [49][C] fp32 correlation buffer, so it
// is at least that large)
template <typename T>
size_t fast_smem_bytes(int HW, bool bwd) {
  const size_t tile = size_t(HW) * C * sizeof(T);
  const size_t da = tile > size_t(49) * C * sizeof(float) ? tile : size_t(49) * C * sizeof(float);
  return sizeof(FastSmem) + 2 * tile + (bwd ? da : 0);
}

template <typename T>
bool fast_ok(const ngu_mona_conv_desc& d, bool bwd) {
  const int HW = d.H * d.W;
  const size_t smem = fast_smem_bytes<T>(HW, bwd);
  return d.W <= 16 && d.H <= 16 && smem <= 227 * 1024;
}

template <typename T>
int conv_smem_bytes(int HW, bool bwd) { return int(sizeof(ConvSmem)) + HW * C * int(sizeof(T)) * (bwd ? 2 : 1); }

template <typename T>
int launch_fwd(const ngu_mona_conv_desc& d, cudaStream_t st) {
  if (sizeof(T) == 2 && fast_ok<T>(d, false) && !d.force_simt) {
    const int smf = int(tc_smem_bytes(d.H * d.W, false));
    cudaError_t e = cudaFuncSetAttribute(mona_conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smf);
    if (e != cudaSuccess) return cuda_status(e, "mona_conv_fwd attr");
    launch_pdl(mona_conv_fwd_tc_kernel, dim3(d.B), dim3(kThreads), size_t(smf), st, reinterpret_cast<const bf16*>(d.h), reinterpret_cast<bf16*>(d.g), d.w, d.N, d.H, d.W,
                                                          d.has_cls, d.drop_p, d.seed, seed_counter());
    return check_launch("mona_conv_fwd");
  }
  if (fast_ok<T>(d, false)) {
    const int smf = int(fast_smem_bytes<T>(d.H * d.W, false));
    cudaError_t e = cudaFuncSetAttribute(mona_conv_fwd_fast_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf);
    if (e != cudaSuccess) return cuda_status(e, "mona_conv_fwd attr");
    launch_pdl(mona_conv_fwd_fast_kernel<T>, dim3(d.B), dim3(kThreads), size_t(smf), st, reinterpret_cast<const T*>(d.h), reinterpret_cast<T*>(d.g), d.w, d.N, d.H,
                                                              d.W, d.has_cls, d.drop_p, d.seed, seed_counter());
    return check_launch("mona_conv_fwd");
  }
  const int smem = conv_smem_bytes<T>(d.H * d.W, false);
  if (smem > 227 * 1024) { set_last_error("mona_conv_fwd: %dx%d grid needs %d B smem", d.H, d.W, smem); return NGU_ERR_SHAPE; }
  cudaError_t e = cudaFuncSetAttribute(mona_conv_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_status(e, "mona_conv_fwd attr");
  launch_pdl(mona_conv_fwd_kernel<T>, dim3(d.B), dim3(kThreads), size_t(smem), st, reinterpret_cast<const T*>(d.h), reinterpret_cast<T*>(d.g), d.w, d.N, d.H,
                                                        d.W, d.has_cls, d.drop_p, d.seed, seed_counter());
  return check_launch("mona_conv_fwd");
}
template <typename T>
int launch_bwd(const ngu_mona_conv_desc& d, cudaStream_t st) {
  if (sizeof(T) == 2 && fast_ok<T>(d, true) && !d.force_simt) {
    const int smf = int(tc_smem_bytes(d.H * d.W, true));
    cudaError_t e = cudaFuncSetAttribute(mona_conv_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smf);
    if (e != cudaSuccess) return cuda_status(e, "mona_conv_bwd attr");
    launch_pdl(mona_conv_bwd_tc_kernel, dim3(d.B), dim3(kThreads), size_t(smf), st, reinterpret_cast<const bf16*>(d.h), reinterpret_cast<const bf16*>(d.dg),
                                                          reinterpret_cast<bf16*>(d.dh), d.w, d.gr, d.N, d.H, d.W, d.has_cls, d.drop_p, d.seed, seed_counter());
    return check_launch("mona_conv_bwd");
  }
  if (fast_ok<T>(d, true)) {
    const int smf = int(fast_smem_bytes<T>(d.H * d.W, true));
    cudaError_t e = cudaFuncSetAttribute(mona_conv_bwd_fast_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf);
    if (e != cudaSuccess) return cuda_status(e, "mona_conv_bwd attr");
    launch_pdl(mona_conv_bwd_fast_kernel<T>, dim3(d.B), dim3(kThreads), size_t(smf), st, reinterpret_cast<const T*>(d.h), reinterpret_cast<const T*>(d.dg),
                                                              reinterpret_cast<T*>(d.dh), d.w, d.gr, d.N, d.H, d.W, d.has_cls,
                                                              d.drop_p, d.seed, seed_counter());
    return check_launch("mona_conv_bwd");
  }
  const int smem = conv_smem_bytes<T>(d.H * d.W, true);
  if (smem > 227 * 1024) { set_last_error("mona_conv_bwd: %dx%d grid needs %d B smem", d.H, d.W, smem); return NGU_ERR_SHAPE; }
  cudaError_t e = cudaFuncSetAttribute(mona_conv_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_status(e, "mona_conv_bwd attr");
  launch_pdl(mona_conv_bwd_kernel<T>, dim3(d.B), dim3(kThreads), size_t(smem), st, reinterpret_cast<const T*>(d.h), reinterpret_cast<const T*>(d.dg),
                                                        reinterpret_cast<T*>(d.dh), d.w, d.gr, d.N, d.H, d.W, d.has_cls,
                                                        d.drop_p, d.seed, seed_counter());
  return check_launch("mona_conv_bwd");
}

int validate(const ngu_mona_conv_desc& d, const char* what) {
  if (d.C != C) { set_last_error("%s: bottleneck %d not instantiated (only %d)", what, d.C, C); return NGU_ERR_SHAPE; }
  if (d.B <= 0 || d.H <= 0 || d.W <= 0 || d.N != d.H * d.W + (d.has_cls ? 1 : 0)) {
    set_last_error("%s: bad shape B=%d N=%d H=%d W=%d has_cls=%d", what, d.B, d.N, d.H, d.W, d.has_cls);
    return NGU_ERR_SHAPE;
  }
  if (d.drop_p < 0.f || d.drop_p >= 1.f) { set_last_error("%s: dropout p=%f out of range", what, d.drop_p); return NGU_ERR_ARG; }
  return NGU_OK;
}

int validate_variant(const ngu_mona_conv_desc& d, const char* what, bool bwd) {
  const bool noise = d.w.ne_w1 != nullptr;
  if (noise && !(d.w.ne_b1 && d.w.ne_w2 && d.w.ne_b2)) { set_last_error("%s: noise estimator needs w1, b1, w2, b2", what); return NGU_ERR_ARG; }
  if ((noise || d.w.freq) && !(d.H <= 16 && d.W <= 16)) {
    set_last_error("%s: Mona variants (frequency filter / noise estimator) are built for grids up to 16x16 (got %dx%d)", what, d.H, d.W);
    return NGU_ERR_SHAPE;
  }
  if (bwd && noise && !(d.gr.dne_w1 && d.gr.dne_b1 && d.gr.dne_w2 && d.gr.dne_b2)) { set_last_error("%s: missing noise-estimator gradient buffers", what); return NGU_ERR_ARG; }
  if (bwd && d.w.freq && !d.gr.dfreq) { set_last_error("%s: missing freq_filter gradient buffer", what); return NGU_ERR_ARG; }
  return NGU_OK;
}

}  // namespace

int mona_conv_fwd(const ngu_mona_conv_desc& d, cudaStream_t st) {
  if (int rc = validate(d, "mona_conv_fwd")) return rc;
  if (int rc = validate_variant(d, "mona_conv_fwd", false)) return rc;
  return d.dtype == NGU_F32 ? launch_fwd<float>(d, st) : launch_fwd<bf16>(d, st);
}
int mona_conv_bwd(const ngu_mona_conv_desc& d, cudaStream_t st) {
  if (int rc = validate(d, "mona_conv_bwd")) return rc;
  if (int rc = validate_variant(d, "mona_conv_bwd", true)) return rc;
  return d.dtype == NGU_F32 ? launch_bwd<float>(d, st) : launch_bwd<bf16>(d, st);
}

#ifdef NGU_CONV_PROF
extern "C" int ngu_debug_conv_prof(long long* out) { return cudaMemcpyFromSymbol(out, g_conv_prof, sizeof(g_conv_prof)) == cudaSuccess ? 0 : -4; }
#endif
}  // namespace ngu
