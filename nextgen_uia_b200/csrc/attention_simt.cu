// CUDA-core scaled-dot-product attention, forward and backward, generic over the activation dtype.
// This is the fp32 check mode of the attention core (and the on-device cross-check for the tcgen05
// kernel in attention_tc.cu).  Semantics = F.scaled_dot_product_attention(q, k, v) as called by timm
// Attention (pinned dep) and src/adapters/lora.py:188-190: softmax(q k^T / sqrt(dh)) v, optional
// causal mask (CLIP text tower, src/third_party/openai_clip/model.py:361-374), no dropout.
// One CTA per (batch, head); K/V (and Q/dO in backward) staged in shared memory.
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

constexpr int kWarps = 8;
constexpr int DH = 64;

template <typename T> struct Pad;
template <> struct Pad<float> { static constexpr int LD = DH + 1; };
template <> struct Pad<bf16> { static constexpr int LD = DH + 2; };

template <typename T>
NGU_DEVINL void load_tile(T* dst, const T* src, int rows, int64_t ts) {
  constexpr int LD = Pad<T>::LD;
  for (int i = threadIdx.x; i < rows * DH; i += kWarps * 32) {
    const int r = i / DH, c = i % DH;
    dst[r * LD + c] = src[int64_t(r) * ts + c];
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
attn_fwd_simt_kernel(ngu_attn_desc d) {
  pdl_prologue();
  constexpr int LD = Pad<T>::LD;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const int b = blockIdx.x / d.H, hd = blockIdx.x % d.H;
  const int N = d.N;
  int S = d.S;
  const int Sfull = d.S;
  if (d.kv_len) {   // key-padding mask: only the first kv_len[b] keys take part
    const int l = d.kv_len[b];
    S = l < 1 ? 1 : (l > Sfull ? Sfull : l);
  }
  T* Ks = reinterpret_cast<T*>(smem_dyn);
  T* Vs = Ks + Sfull * LD;
  float* pbuf = reinterpret_cast<float*>(Vs + Sfull * LD);  // [kWarps][S]
  float* qbuf = pbuf + kWarps * Sfull;                      // [kWarps][DH]
  const T* q = reinterpret_cast<const T*>(d.q) + int64_t(b) * d.q_bs + hd * DH;
  const T* k = reinterpret_cast<const T*>(d.k) + int64_t(b) * d.k_bs + hd * DH;
  const T* v = reinterpret_cast<const T*>(d.v) + int64_t(b) * d.v_bs + hd * DH;
  T* o = reinterpret_cast<T*>(d.o) + int64_t(b) * d.o_bs + hd * DH;
  load_tile<T>(Ks, k, S, d.k_ts);
  load_tile<T>(Vs, v, S, d.v_ts);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pw = pbuf + warp * Sfull;
  float* qw = qbuf + warp * DH;
  for (int i = warp; i < N; i += kWarps) {
    qw[lane] = to_f32<T>(q[int64_t(i) * d.q_ts + lane]);
    qw[lane + 32] = to_f32<T>(q[int64_t(i) * d.q_ts + lane + 32]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      float s = 0.f;
#pragma unroll 16
      for (int c = 0; c < DH; ++c) s = fmaf(qw[c], to_f32<T>(Ks[j * LD + c]), s);
      s *= d.scale;
      if (d.causal && j > i) s = -INFINITY;
      pw[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float p = expf(pw[j] - mx);
      pw[j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float p = pw[j];
      o0 = fmaf(p, to_f32<T>(Vs[j * LD + lane]), o0);
      o1 = fmaf(p, to_f32<T>(Vs[j * LD + lane + 32]), o1);
    }
    const float inv = 1.f / sum;
    o[int64_t(i) * d.o_ts + lane] = from_f32<T>(o0 * inv);
    o[int64_t(i) * d.o_ts + lane + 32] = from_f32<T>(o1 * inv);
    if (lane == 0 && d.lse) d.lse[(int64_t(b) * d.H + hd) * N + i] = mx + logf(sum);
    __syncwarp();
  }
}

// Backward: pass A (warp per query row) -> dq; pass B (warp per key row) -> dk, dv.  Probabilities
// are recomputed from the saved log-sum-exp.  When Q, dO, K and V do not fit in shared memory together (N = 577 of
// ViT-L/14@336, N = 485 of ViT-B/16@352) the kernel runs in two phases that share one pair of tiles: K/V resident for pass A
// (the warp's q / dO row comes from global memory), then Q/dO resident for pass B (k / v row from global memory).
template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
attn_bwd_simt_kernel(ngu_attn_desc d, int two_phase) {
  pdl_prologue();
  constexpr int LD = Pad<T>::LD;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const int b = blockIdx.x / d.H, hd = blockIdx.x % d.H;
  const int N = d.N, S = d.S;
  // key-padding mask: keys >= kv_len[b] get probability 0 (pass A stops there, pass B writes zero dk / dv rows)
  int Sk = S;
  if (d.kv_len) { const int l = d.kv_len[b]; Sk = l < 1 ? 1 : (l > S ? S : l); }
  const int Lr = N > S ? N : S;
  T* Qs = reinterpret_cast<T*>(smem_dyn);
  T* dOs = Qs + (two_phase ? Lr : N) * LD;
  T* Ks = two_phase ? Qs : dOs + N * LD;
  T* Vs = two_phase ? dOs : Ks + S * LD;
  float* delta = reinterpret_cast<float*>((two_phase ? dOs + Lr * LD : Vs + S * LD));  // [N]
  float* lse_s = delta + N;                              // [N]
  const int L = N > S ? N : S;
  float* buf0 = lse_s + N;                               // [kWarps][L]
  float* buf1 = buf0 + kWarps * L;                       // [kWarps][L]
  float* rowb = buf1 + kWarps * L;                       // [kWarps][2*DH]
  const T* q = reinterpret_cast<const T*>(d.q) + int64_t(b) * d.q_bs + hd * DH;
  const T* k = reinterpret_cast<const T*>(d.k) + int64_t(b) * d.k_bs + hd * DH;
  const T* v = reinterpret_cast<const T*>(d.v) + int64_t(b) * d.v_bs + hd * DH;
  const T* o = reinterpret_cast<const T*>(d.o) + int64_t(b) * d.o_bs + hd * DH;
  const T* dO = reinterpret_cast<const T*>(d.d_o) + int64_t(b) * d.o_bs + hd * DH;
  T* dq = reinterpret_cast<T*>(d.dq) + int64_t(b) * d.q_bs + hd * DH;
  T* dk = reinterpret_cast<T*>(d.dk) + int64_t(b) * d.k_bs + hd * DH;
  T* dv = reinterpret_cast<T*>(d.dv) + int64_t(b) * d.v_bs + hd * DH;
  if (!two_phase) {
    load_tile<T>(Qs, q, N, d.q_ts);
    load_tile<T>(dOs, dO, N, d.o_ts);
  }
  load_tile<T>(Ks, k, S, d.k_ts);
  load_tile<T>(Vs, v, S, d.v_ts);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < N; i += kWarps) {
    float t = to_f32<T>(o[int64_t(i) * d.o_ts + lane]) * to_f32<T>(dO[int64_t(i) * d.o_ts + lane]) +
              to_f32<T>(o[int64_t(i) * d.o_ts + lane + 32]) * to_f32<T>(dO[int64_t(i) * d.o_ts + lane + 32]);
    t = warp_sum(t);
    if (lane == 0) { delta[i] = t; lse_s[i] = d.lse[(int64_t(b) * d.H + hd) * N + i]; }
  }
  __syncthreads();
  float* b0 = buf0 + warp * L;
  float* b1 = buf1 + warp * L;
  float* rw = rowb + warp * 2 * DH;
  // pass A: dq_i = scale * sum_j ds_ij k_j
  for (int i = warp; i < N; i += kWarps) {
    if (two_phase) {
      rw[lane] = to_f32<T>(q[int64_t(i) * d.q_ts + lane]); rw[lane + 32] = to_f32<T>(q[int64_t(i) * d.q_ts + lane + 32]);
      rw[DH + lane] = to_f32<T>(dO[int64_t(i) * d.o_ts + lane]); rw[DH + lane + 32] = to_f32<T>(dO[int64_t(i) * d.o_ts + lane + 32]);
    } else {
      rw[lane] = to_f32<T>(Qs[i * LD + lane]); rw[lane + 32] = to_f32<T>(Qs[i * LD + lane + 32]);
      rw[DH + lane] = to_f32<T>(dOs[i * LD + lane]); rw[DH + lane + 32] = to_f32<T>(dOs[i * LD + lane + 32]);
    }
    __syncwarp();
    const float li = lse_s[i], di = delta[i];
    for (int j = lane; j < Sk; j += 32) {
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int c = 0; c < DH; ++c) {
        s = fmaf(rw[c], to_f32<T>(Ks[j * LD + c]), s);
        dp = fmaf(rw[DH + c], to_f32<T>(Vs[j * LD + c]), dp);
      }
      float p = expf(s * d.scale - li);
      if (d.causal && j > i) p = 0.f;
      b0[j] = p * (dp - di) * d.scale;
    }
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < Sk; ++j) {
      const float ds = b0[j];
      a0 = fmaf(ds, to_f32<T>(Ks[j * LD + lane]), a0);
      a1 = fmaf(ds, to_f32<T>(Ks[j * LD + lane + 32]), a1);
    }
    dq[int64_t(i) * d.q_ts + lane] = from_f32<T>(a0);
    dq[int64_t(i) * d.q_ts + lane + 32] = from_f32<T>(a1);
    __syncwarp();
  }
  // pass B: dv_j = sum_i p_ij dO_i ; dk_j = scale * sum_i ds_ij q_i
  if (two_phase) {   // the K/V tiles make room for Q/dO
    __syncthreads();
    load_tile<T>(Qs, q, N, d.q_ts);
    load_tile<T>(dOs, dO, N, d.o_ts);
    __syncthreads();
  }
  for (int j = warp; j < S; j += kWarps) {
    if (j >= Sk) {   // masked key: no query attends to it
      dv[int64_t(j) * d.v_ts + lane] = from_f32<T>(0.f); dv[int64_t(j) * d.v_ts + lane + 32] = from_f32<T>(0.f);
      dk[int64_t(j) * d.k_ts + lane] = from_f32<T>(0.f); dk[int64_t(j) * d.k_ts + lane + 32] = from_f32<T>(0.f);
      continue;
    }
    if (two_phase) {
      rw[lane] = to_f32<T>(k[int64_t(j) * d.k_ts + lane]); rw[lane + 32] = to_f32<T>(k[int64_t(j) * d.k_ts + lane + 32]);
      rw[DH + lane] = to_f32<T>(v[int64_t(j) * d.v_ts + lane]); rw[DH + lane + 32] = to_f32<T>(v[int64_t(j) * d.v_ts + lane + 32]);
    } else {
      rw[lane] = to_f32<T>(Ks[j * LD + lane]); rw[lane + 32] = to_f32<T>(Ks[j * LD + lane + 32]);
      rw[DH + lane] = to_f32<T>(Vs[j * LD + lane]); rw[DH + lane + 32] = to_f32<T>(Vs[j * LD + lane + 32]);
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int c = 0; c < DH; ++c) {
        s = fmaf(rw[c], to_f32<T>(Qs[i * LD + c]), s);
        dp = fmaf(rw[DH + c], to_f32<T>(dOs[i * LD + c]), dp);
      }
      float p = expf(s * d.scale - lse_s[i]);
      if (d.causal && j > i) p = 0.f;
      b0[i] = p;
      b1[i] = p * (dp - delta[i]) * d.scale;
    }
    __syncwarp();
    float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
    for (int i = 0; i < N; ++i) {
      const float p = b0[i], ds = b1[i];
      v0 = fmaf(p, to_f32<T>(dOs[i * LD + lane]), v0);
      v1 = fmaf(p, to_f32<T>(dOs[i * LD + lane + 32]), v1);
      k0 = fmaf(ds, to_f32<T>(Qs[i * LD + lane]), k0);
      k1 = fmaf(ds, to_f32<T>(Qs[i * LD + lane + 32]), k1);
    }
    dv[int64_t(j) * d.v_ts + lane] = from_f32<T>(v0);
    dv[int64_t(j) * d.v_ts + lane + 32] = from_f32<T>(v1);
    dk[int64_t(j) * d.k_ts + lane] = from_f32<T>(k0);
    dk[int64_t(j) * d.k_ts + lane + 32] = from_f32<T>(k1);
    __syncwarp();
  }
}

template <typename T>
int launch_fwd(const ngu_attn_desc& d, cudaStream_t st) {
  constexpr int LD = Pad<T>::LD;
  const int smem = 2 * d.S * LD * int(sizeof(T)) + kWarps * d.S * 4 + kWarps * DH * 4;
  if (smem > 227 * 1024) { set_last_error("attn_fwd(simt): S=%d needs %d B smem", d.S, smem); return NGU_ERR_SHAPE; }
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_status(e, "attn_fwd attr");
  launch_pdl(attn_fwd_simt_kernel<T>, dim3(d.B * d.H), dim3(kWarps * 32), size_t(smem), st, d);
  return check_launch("attn_fwd_simt");
}
template <typename T>
int launch_bwd(const ngu_attn_desc& d, cudaStream_t st) {
  constexpr int LD = Pad<T>::LD;
  const int L = d.N > d.S ? d.N : d.S;
  const int rest = 2 * d.N * 4 + 2 * kWarps * L * 4 + kWarps * 2 * DH * 4;
  int smem = (2 * d.N + 2 * d.S) * LD * int(sizeof(T)) + rest;
  int two_phase = 0;
  if (smem > 227 * 1024) {   // Q/dO and K/V take turns in one pair of tiles
    two_phase = 1;
    smem = 2 * L * LD * int(sizeof(T)) + rest;
  }
  if (smem > 227 * 1024) { set_last_error("attn_bwd(simt): N=%d S=%d needs %d B smem", d.N, d.S, smem); return NGU_ERR_SHAPE; }
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_status(e, "attn_bwd attr");
  launch_pdl(attn_bwd_simt_kernel<T>, dim3(d.B * d.H), dim3(kWarps * 32), size_t(smem), st, d, two_phase);
  return check_launch("attn_bwd_simt");
}

}  // namespace

int attn_validate(const ngu_attn_desc& d, const char* what, bool bwd) {
  if (d.dh != DH) { set_last_error("%s: head dim %d not instantiated (only 64)", what, d.dh); return NGU_ERR_SHAPE; }
  if (d.B <= 0 || d.H <= 0 || d.N <= 0 || d.S <= 0) { set_last_error("%s: empty problem", what); return NGU_ERR_SHAPE; }
  if (!d.q || !d.k || !d.v || !d.o) { set_last_error("%s: null q/k/v/o", what); return NGU_ERR_ARG; }
  if (bwd && (!d.d_o || !d.dq || !d.dk || !d.dv || !d.lse)) { set_last_error("%s: backward needs d_o, dq, dk, dv, lse", what); return NGU_ERR_ARG; }
  return NGU_OK;
}

int attn_fwd_simt(const ngu_attn_desc& d, cudaStream_t st) {
  if (int rc = attn_validate(d, "attn_fwd", false)) return rc;
  return d.dtype == NGU_F32 ? launch_fwd<float>(d, st) : launch_fwd<bf16>(d, st);
}
int attn_bwd_simt(const ngu_attn_desc& d, cudaStream_t st) {
  if (int rc = attn_validate(d, "attn_bwd", true)) return rc;
  return d.dtype == NGU_F32 ? launch_bwd<float>(d, st) : launch_bwd<bf16>(d, st);
}

}  // namespace ngu
