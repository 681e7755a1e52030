// Small data-movement kernels around the block: patchify (im2col of the stride-16 patch-embed conv,
// timm PatchEmbed / src/third_party/openai_clip/model.py:221,234), token assembly (+cls, +pos),
// fp32 -> bf16 casts with optional transpose for the trainable adapter matrices, and the SIMT
// weight-gradient reduction used by the fp32 check mode.
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

// images [B,3,R,R] fp32 (NCHW) -> patches [B*G*G, 3*P*P] (T), column order (c, py, px) = conv weight flattening.
// Thread per 8 consecutive pixels of one image row: two 16-byte loads, one contiguous 16/32-byte store (P % 8 == 0);
// 32-bit index math.
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int R, int P) {
  pdl_prologue();
  const unsigned G = R / P, K = 3 * P * P, R8 = R / 8;
  const unsigned total8 = unsigned(B) * 3u * unsigned(R) * R8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += gridDim.x * blockDim.x) {
    const unsigned x8 = i % R8, rowi = i / R8;            // rowi = (b*3 + c)*R + yy
    const unsigned yy = rowi % unsigned(R), bc = rowi / unsigned(R);
    const unsigned c = bc % 3u, b = bc / 3u;
    const unsigned xx = x8 * 8;
    const float4 v0 = *reinterpret_cast<const float4*>(img + size_t(i) * 8);
    const float4 v1 = *reinterpret_cast<const float4*>(img + size_t(i) * 8 + 4);
    const unsigned gx = xx / unsigned(P), px = xx % unsigned(P), gy = yy / unsigned(P), py = yy % unsigned(P);
    T* dst = out + (size_t(b) * G * G + size_t(gy) * G + gx) * K + c * P * P + py * P + px;
    const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    if (sizeof(T) == 2) {
      uint4 u;
      u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(dst) = u;
    } else {
      reinterpret_cast<float4*>(dst)[0] = v0;
      reinterpret_cast<float4*>(dst)[1] = v1;
    }
  }
}

// Any patch size (ViT-L/14): thread per pixel; output rows have pitch Kp = ceil8(3*P*P), the tail stays as the caller
// initialised it (zero) so the row is a legal K operand of the GEMM.
template <typename T>
__global__ void patchify_generic_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int R, int P, int Kp) {
  pdl_prologue();
  const unsigned G = R / P;
  const unsigned total = unsigned(B) * 3u * unsigned(R) * unsigned(R);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned xx = i % unsigned(R), rowi = i / unsigned(R);
    const unsigned yy = rowi % unsigned(R), bc = rowi / unsigned(R);
    const unsigned c = bc % 3u, b = bc / 3u;
    const unsigned gx = xx / unsigned(P), px = xx % unsigned(P), gy = yy / unsigned(P), py = yy % unsigned(P);
    if (gx >= G || gy >= G) continue;                      // pixels beyond the last whole patch are dropped (conv stride = P)
    out[(size_t(b) * G * G + size_t(gy) * G + gx) * Kp + c * P * P + py * P + px] = from_f32<T>(img[i]);
  }
}

// x0[b,0,:] = cls + pos[0];  x0[b,1+p,:] = patch[b*np+p,:] + pos[1+p]
template <typename T>
__global__ void assemble_kernel(const T* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                                T* __restrict__ out, int B, int np, int D) {
  pdl_prologue();
  // one 16-byte vector of a token row per thread; 32-bit index math (the scalar 64-bit div/mod version ran 8x off the
  // HBM roofline)
  constexpr int V = Vec<T>::N;
  const int vpr = D / V;                                 // vectors per row
  const unsigned total = unsigned(B) * unsigned(np + 1) * unsigned(vpr);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned row = i / unsigned(vpr);
    const int c = int(i - row * unsigned(vpr)) * V;
    const int b = int(row / unsigned(np + 1));
    const int n = int(row - unsigned(b) * unsigned(np + 1));
    float v[V], a[V];
#pragma unroll
    for (int e = 0; e < V; e += 4) {
      const float4 q = *reinterpret_cast<const float4*>(pos + size_t(n) * D + c + e);
      v[e] = q.x; v[e + 1] = q.y; v[e + 2] = q.z; v[e + 3] = q.w;
    }
    if (n == 0) {
#pragma unroll
      for (int e = 0; e < V; ++e) a[e] = cls[c + e];
    } else {
      Vec<T>::load(patch + (size_t(b) * np + (n - 1)) * D + c, a);
    }
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] += a[e];
    Vec<T>::store(out + size_t(row) * D + c, v);
  }
}

// BERT input embeddings: out[b,s,:] = word[ids[b,s]] + pos[s] + type0   (HF BertEmbeddings before its LayerNorm)
template <typename T>
__global__ void embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
                             const float* __restrict__ type0, T* __restrict__ out, int B, int S, int D, int vocab) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  const int vpr = D / V;
  const unsigned total = unsigned(B) * unsigned(S) * unsigned(vpr);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned tok = i / unsigned(vpr);
    const int c = int(i - tok * unsigned(vpr)) * V;
    const int sidx = int(tok % unsigned(S));
    int64_t id = ids[tok];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    float v[V];
#pragma unroll
    for (int e = 0; e < V; e += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(word + size_t(id) * D + c + e);
      const float4 p4 = *reinterpret_cast<const float4*>(pos + size_t(sidx) * D + c + e);
      const float4 t4 = *reinterpret_cast<const float4*>(type0 + c + e);
      v[e] = w4.x + p4.x + t4.x; v[e + 1] = w4.y + p4.y + t4.y; v[e + 2] = w4.z + p4.z + t4.z; v[e + 3] = w4.w + p4.w + t4.w;
    }
    Vec<T>::store(out + size_t(tok) * D + c, v);
  }
}

// out (T) [cols, rows] or [rows, cols] <- in fp32 [rows, cols] * scale
template <typename T>
__global__ void cast_kernel(const float* __restrict__ in, T* __restrict__ out, int rows, int cols, int transpose, float scale) {
  pdl_prologue();
  const size_t total = size_t(rows) * cols;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int r = int(i / cols), c = int(i % cols);
    const float v = in[i] * scale;
    if (transpose) out[size_t(c) * rows + r] = from_f32<T>(v);
    else out[i] = from_f32<T>(v);
  }
}

// table of casts in one launch: blockIdx.y = item, blockIdx.x strides over its elements
template <typename T>
__global__ void cast_batch_kernel(const ngu_cast_item* __restrict__ items) {
  pdl_prologue();
  const ngu_cast_item it = items[blockIdx.y];
  const unsigned total = unsigned(it.rows) * unsigned(it.cols);
  T* out = reinterpret_cast<T*>(it.out);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned r = i / unsigned(it.cols), c = i - r * unsigned(it.cols);
    const float v = it.in[i] * it.scale;
    if (it.transpose) out[size_t(c) * it.rows + r] = from_f32<T>(v);
    else out[i] = from_f32<T>(v);
  }
}

// D[Mo,No] += X^T Y,  X [T,Mo], Y [T,No] row-major, reduction over tokens split across blockIdx.z
constexpr int WT = 64, WK = 16;
template <typename T>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const T* __restrict__ X, int ldx, const T* __restrict__ Y, int ldy,
                                                         float* __restrict__ D, int ldd, int Tn, int Mo, int No, int tchunk) {
  pdl_prologue();
  __shared__ float sX[WK][WT + 1];
  __shared__ float sY[WK][WT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * WT, n0 = blockIdx.x * WT;
  const int t0 = blockIdx.z * tchunk, t1 = min(Tn, t0 + tchunk);
  float acc[4][4] = {};
  for (int tb = t0; tb < t1; tb += WK) {
    for (int i = threadIdx.x; i < WK * WT; i += 256) {
      const int r = i / WT, c = i % WT;
      const int t = tb + r;
      sX[r][c] = (t < t1 && m0 + c < Mo) ? to_f32<T>(X[size_t(t) * ldx + m0 + c]) : 0.f;
      sY[r][c] = (t < t1 && n0 + c < No) ? to_f32<T>(Y[size_t(t) * ldy + n0 + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sX[k][ty * 4 + i]; b[i] = sY[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < Mo && n < No) atomicAdd(D + size_t(m) * ldd + n, acc[i][j]);
    }
}

// column sums of a [T, C] activation into fp32 (bias gradients): CTA = 8 warps over a 32-vector-wide column tile,
// each warp walks rows with 16-byte loads, cross-warp reduce in smem, one atomic per column per CTA.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, int ldx, float* __restrict__ out, int Tn, int Cn, int tchunk) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * 32 + lane) * V;
  const int t0 = blockIdx.y * tchunk, t1 = min(Tn, t0 + tchunk);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (c0 + V <= Cn) {
    for (int t = t0 + warp; t < t1; t += 8) {
      float v[V];
      Vec<T>::load(X + size_t(t) * ldx + c0, v);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += v[i];
    }
  } else {
    for (int t = t0 + warp; t < t1; t += 8)
      for (int i = 0; i < V; ++i)
        if (c0 + i < Cn) acc[i] += to_f32<T>(X[size_t(t) * ldx + c0 + i]);
  }
  __shared__ float red[8][32 * V + 1];
#pragma unroll
  for (int i = 0; i < V; ++i) red[warp][lane * V + i] = acc[i];
  __syncthreads();
  for (int j = threadIdx.x; j < 32 * V; j += 256) {
    const int c = blockIdx.x * 32 * V + j;
    if (c < Cn) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][j];
      atomicAdd(out + c, s);
    }
  }
}

// out = (accumulate ? out : 0) + x * mask(seed, i) / (1-p)   (LoRA input dropout, src/adapters/lora.py:82-83)
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ out, size_t n, float p, uint64_t seed, int accumulate,
                               const uint64_t* seed_ctr) {
  pdl_prologue();
  seed = mix_seed(seed, seed_ctr);
  // one 16-byte vector per thread and iteration, one hash per four elements
  constexpr int V = Vec<T>::N;
  const uint32_t thr = dropout_threshold(p);
  const float ks = 1.0f / (1.0f - p);
  const size_t nv = n / V;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += size_t(gridDim.x) * blockDim.x) {
    float v[V], o[V];
    Vec<T>::load(x + i * V, v);
    if (accumulate) Vec<T>::load(out + i * V, o);
#pragma unroll
    for (int g = 0; g < V / 4; ++g) {
      const uint64_t bits = dropout_bits(seed, (i * V) / 4 + g);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float r = v[4 * g + e] * dropout_pick(bits, e, thr, ks);
        v[4 * g + e] = accumulate ? r + o[4 * g + e] : r;
      }
    }
    Vec<T>::store(out + i * V, v);
  }
  for (size_t i = nv * V + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    float v = to_f32<T>(x[i]) * dropout_scale(seed, i, p);
    if (accumulate) v += to_f32<T>(out[i]);
    out[i] = from_f32<T>(v);
  }
}

// Key-padding lengths of a right-padded token batch (open_clip HFTextEncoder.forward: attn_mask = (x != pad_token_id)):
// one warp per sequence; flag |= 1 when the valid tokens are not a non-empty prefix (the attention kernels only take lengths).
__global__ void kv_len_kernel(const int64_t* __restrict__ ids, int64_t pad, int* __restrict__ out, int* __restrict__ flag, int B, int S) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  int cnt = 0, last = -1;
  for (int j = lane; j < S; j += 32)
    if (ids[size_t(row) * S + j] != pad) { ++cnt; last = j; }
  cnt = warp_sum(cnt);
  last = warp_max(last);
  if (lane == 0) {
    out[row] = cnt < 1 ? 1 : cnt;
    if (cnt < 1 || last != cnt - 1) atomicOr(flag, 1);
  }
}

// Zero-shot prompt-ensemble scorer (src/models/biomedclip/zero_shot.py:176-228): the mean over a class's prompts of
// 100 * Ihat . That_p equals 100 * Ihat . mean_p(That_p), so scoring needs one prototype per class.
//   prototypes: proto[c] = mean over prompts p of class c of  t_p / ||t_p||          (one block per class)
//   scores    : logits[b][c] = scale * <f_b / ||f_b||, proto[c]>,  pred[b] = argmax_c  (one warp per image)
template <typename T>
__global__ void __launch_bounds__(256) zs_proto_kernel(const T* __restrict__ tf, const int* __restrict__ cls, float* __restrict__ proto, int P, int E) {
  pdl_prologue();
  __shared__ float red[8];
  __shared__ float inv_s;
  const int c = blockIdx.x;
  int cnt = 0;
  for (int e = threadIdx.x; e < E; e += 256) proto[size_t(c) * E + e] = 0.f;
  for (int pi = 0; pi < P; ++pi) {
    if (cls[pi] != c) continue;      // block-uniform
    ++cnt;
    float s = 0.f;
    for (int e = threadIdx.x; e < E; e += 256) { const float v = to_f32<T>(tf[size_t(pi) * E + e]); s = fmaf(v, v, s); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; inv_s = rsqrtf(fmaxf(t, 1e-24f)); }
    __syncthreads();
    const float inv = inv_s;
    for (int e = threadIdx.x; e < E; e += 256) proto[size_t(c) * E + e] += to_f32<T>(tf[size_t(pi) * E + e]) * inv;
    __syncthreads();
  }
  const float k = cnt > 0 ? 1.f / float(cnt) : 0.f;
  for (int e = threadIdx.x; e < E; e += 256) proto[size_t(c) * E + e] *= k;
}

template <typename T>
__global__ void __launch_bounds__(128) zs_score_kernel(const T* __restrict__ f, const float* __restrict__ proto, float* __restrict__ logits,
                                                       int* __restrict__ pred, int B, int E, int C, float scale) {
  pdl_prologue();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  float ss = 0.f;
  for (int e = lane; e < E; e += 32) { const float v = to_f32<T>(f[size_t(b) * E + e]); ss = fmaf(v, v, ss); }
  const float inv = rsqrtf(fmaxf(warp_sum(ss), 1e-24f));
  float best = -INFINITY;
  int arg = 0;
  for (int c = 0; c < C; ++c) {
    float d = 0.f;
    for (int e = lane; e < E; e += 32) d = fmaf(to_f32<T>(f[size_t(b) * E + e]), proto[size_t(c) * E + e], d);
    d = warp_sum(d) * inv * scale;
    if (lane == 0) logits[size_t(b) * C + c] = d;
    if (d > best) { best = d; arg = c; }       // first maximum wins, like torch.argmax
  }
  if (lane == 0) pred[b] = arg;
}

int grid_for(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = size_t(sm_count()) * 16;
  return int(g > cap ? cap : (g == 0 ? 1 : g));
}

}  // namespace

int patchify(const float* img, void* out, int B, int R, int P, int dtype, cudaStream_t st) {
  if (B <= 0 || R <= 0 || P <= 0 || R % P) { set_last_error("patchify: bad shape B=%d R=%d P=%d", B, R, P); return NGU_ERR_SHAPE; }
  if (B <= 0 || P <= 0 || R < P) { set_last_error("patchify: bad shape B=%d R=%d P=%d", B, R, P); return NGU_ERR_SHAPE; }
  if (size_t(B) * 3 * R * R >= (size_t(1) << 32)) { set_last_error("patchify: more than 2^32 pixels"); return NGU_ERR_SHAPE; }
  if (P % 8 || R % P) {
    const int Kp = (3 * P * P + 7) & ~7;
    const size_t tot = size_t(B) * 3 * R * R;
    if (dtype == NGU_F32) launch_pdl(patchify_generic_kernel<float>, dim3(grid_for(tot)), dim3(256), size_t(0), st, img, reinterpret_cast<float*>(out), B, R, P, Kp);
    else launch_pdl(patchify_generic_kernel<bf16>, dim3(grid_for(tot)), dim3(256), size_t(0), st, img, reinterpret_cast<bf16*>(out), B, R, P, Kp);
    return check_launch("patchify");
  }
  const size_t total = size_t(B) * 3 * R * R / 8;
  if (dtype == NGU_F32) launch_pdl(patchify_kernel<float>, dim3(grid_for(total)), dim3(256), size_t(0), st, img, reinterpret_cast<float*>(out), B, R, P);
  else launch_pdl(patchify_kernel<bf16>, dim3(grid_for(total)), dim3(256), size_t(0), st, img, reinterpret_cast<bf16*>(out), B, R, P);
  return check_launch("patchify");
}
int assemble_tokens(const void* patch, const float* cls, const float* pos, void* out, int B, int np, int D, int dtype, cudaStream_t st) {
  if (B <= 0 || np <= 0 || D <= 0) { set_last_error("assemble_tokens: empty"); return NGU_ERR_SHAPE; }
  if (D % 8) { set_last_error("assemble_tokens: D must be a multiple of 8"); return NGU_ERR_SHAPE; }
  const size_t total = size_t(B) * (np + 1) * D / (dtype == NGU_F32 ? 4 : 8);
  if (dtype == NGU_F32) launch_pdl(assemble_kernel<float>, dim3(grid_for(total)), dim3(256), size_t(0), st, reinterpret_cast<const float*>(patch), cls, pos, reinterpret_cast<float*>(out), B, np, D);
  else launch_pdl(assemble_kernel<bf16>, dim3(grid_for(total)), dim3(256), size_t(0), st, reinterpret_cast<const bf16*>(patch), cls, pos, reinterpret_cast<bf16*>(out), B, np, D);
  return check_launch("assemble_tokens");
}
int embed_tokens(const int64_t* ids, const float* word, const float* pos, const float* type0, void* out, int B, int S, int D,
                 int vocab, int dtype, cudaStream_t st) {
  if (B <= 0 || S <= 0 || D <= 0 || vocab <= 0) { set_last_error("embed_tokens: empty"); return NGU_ERR_SHAPE; }
  if (D % 8) { set_last_error("embed_tokens: D must be a multiple of 8"); return NGU_ERR_SHAPE; }
  const size_t total = size_t(B) * S * D / (dtype == NGU_F32 ? 4 : 8);
  if (dtype == NGU_F32) launch_pdl(embed_kernel<float>, dim3(grid_for(total)), dim3(256), size_t(0), st, ids, word, pos, type0, reinterpret_cast<float*>(out), B, S, D, vocab);
  else launch_pdl(embed_kernel<bf16>, dim3(grid_for(total)), dim3(256), size_t(0), st, ids, word, pos, type0, reinterpret_cast<bf16*>(out), B, S, D, vocab);
  return check_launch("embed_tokens");
}
int cast_f32(const float* in, void* out, int rows, int cols, int transpose, float scale, int dtype, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) { set_last_error("cast: empty"); return NGU_ERR_SHAPE; }
  const size_t total = size_t(rows) * cols;
  if (dtype == NGU_F32) launch_pdl(cast_kernel<float>, dim3(grid_for(total)), dim3(256), size_t(0), st, in, reinterpret_cast<float*>(out), rows, cols, transpose, scale);
  else launch_pdl(cast_kernel<bf16>, dim3(grid_for(total)), dim3(256), size_t(0), st, in, reinterpret_cast<bf16*>(out), rows, cols, transpose, scale);
  return check_launch("cast");
}
int cast_f32_batch(const ngu_cast_item* items, int n, int dtype, cudaStream_t st) {
  if (n <= 0 || items == nullptr) { set_last_error("cast_batch: empty table"); return NGU_ERR_SHAPE; }
  if (n > 65535) { set_last_error("cast_batch: at most 65535 items per launch"); return NGU_ERR_ARG; }
  const dim3 grid(48, n);
  if (dtype == NGU_F32) launch_pdl(cast_batch_kernel<float>, dim3(grid), dim3(256), size_t(0), st, items);
  else launch_pdl(cast_batch_kernel<bf16>, dim3(grid), dim3(256), size_t(0), st, items);
  return check_launch("cast_batch");
}
int dropout(const void* x, void* out, size_t n, float p, uint64_t seed, int accumulate, int dtype, cudaStream_t st) {
  if (n == 0 || p < 0.f || p >= 1.f) { set_last_error("dropout: bad n/p"); return NGU_ERR_ARG; }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) { set_last_error("dropout: pointers must be 16-byte aligned"); return NGU_ERR_ALIGN; }
  if (dtype == NGU_F32) launch_pdl(dropout_kernel<float>, dim3(grid_for(n / 4 + 1)), dim3(256), size_t(0), st, reinterpret_cast<const float*>(x), reinterpret_cast<float*>(out), n, p, seed, accumulate, seed_counter());
  else launch_pdl(dropout_kernel<bf16>, dim3(grid_for(n / 8 + 1)), dim3(256), size_t(0), st, reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(out), n, p, seed, accumulate, seed_counter());
  return check_launch("dropout");
}
int wgrad_simt(const void* X, int ldx, const void* Y, int ldy, float* D, int ldd, int Tn, int Mo, int No, int dtype, cudaStream_t st) {
  if (Tn <= 0 || Mo <= 0 || No <= 0) { set_last_error("wgrad: empty"); return NGU_ERR_SHAPE; }
  int splits = (Tn + 511) / 512;
  if (splits > 512) splits = 512;
  int tchunk = (Tn + splits - 1) / splits;
  tchunk = (tchunk + WK - 1) / WK * WK;
  splits = (Tn + tchunk - 1) / tchunk;
  dim3 grid((No + WT - 1) / WT, (Mo + WT - 1) / WT, splits);
  if (dtype == NGU_F32) launch_pdl(wgrad_simt_kernel<float>, dim3(grid), dim3(256), size_t(0), st, reinterpret_cast<const float*>(X), ldx, reinterpret_cast<const float*>(Y), ldy, D, ldd, Tn, Mo, No, tchunk);
  else launch_pdl(wgrad_simt_kernel<bf16>, dim3(grid), dim3(256), size_t(0), st, reinterpret_cast<const bf16*>(X), ldx, reinterpret_cast<const bf16*>(Y), ldy, D, ldd, Tn, Mo, No, tchunk);
  return check_launch("wgrad_simt");
}
int colsum(const void* X, int ldx, float* out, int Tn, int Cn, int dtype, cudaStream_t st) {
  if (Tn <= 0 || Cn <= 0) { set_last_error("colsum: empty"); return NGU_ERR_SHAPE; }
  const int V = dtype == NGU_F32 ? 4 : 8;
  if ((ldx % V) || (reinterpret_cast<uintptr_t>(X) & 15)) { set_last_error("colsum: rows must be 16-byte aligned"); return NGU_ERR_ALIGN; }
  const int ctile = 32 * V;
  const int gx = (Cn + ctile - 1) / ctile;
  int splits = (sm_count() * 4 + gx - 1) / gx;
  if (splits > (Tn + 63) / 64) splits = (Tn + 63) / 64;
  if (splits < 1) splits = 1;
  const int tchunk = (Tn + splits - 1) / splits;
  dim3 grid(gx, (Tn + tchunk - 1) / tchunk);
  if (dtype == NGU_F32) launch_pdl(colsum_kernel<float>, dim3(grid), dim3(256), size_t(0), st, reinterpret_cast<const float*>(X), ldx, out, Tn, Cn, tchunk);
  else launch_pdl(colsum_kernel<bf16>, dim3(grid), dim3(256), size_t(0), st, reinterpret_cast<const bf16*>(X), ldx, out, Tn, Cn, tchunk);
  return check_launch("colsum");
}

int kv_len(const int64_t* ids, int64_t pad, int* out, int* flag, int B, int S, cudaStream_t st) {
  if (B <= 0 || S <= 0 || !ids || !out || !flag) { set_last_error("kv_len: bad arguments"); return NGU_ERR_ARG; }
  launch_pdl(kv_len_kernel, dim3((B + 3) / 4), dim3(128), size_t(0), st, ids, pad, out, flag, B, S);
  return check_launch("kv_len");
}

int zero_shot_prototypes(const void* tf, const int* cls, float* proto, int P, int E, int C, int dtype, cudaStream_t st) {
  if (P <= 0 || E <= 0 || C <= 0 || !tf || !cls || !proto) { set_last_error("zero_shot_prototypes: bad arguments"); return NGU_ERR_ARG; }
  if (dtype == NGU_F32) launch_pdl(zs_proto_kernel<float>, dim3(C), dim3(256), size_t(0), st, reinterpret_cast<const float*>(tf), cls, proto, P, E);
  else launch_pdl(zs_proto_kernel<bf16>, dim3(C), dim3(256), size_t(0), st, reinterpret_cast<const bf16*>(tf), cls, proto, P, E);
  return check_launch("zero_shot_prototypes");
}
int zero_shot_score(const void* f, const float* proto, float* logits, int* pred, int B, int E, int C, float scale, int dtype, cudaStream_t st) {
  if (B <= 0 || E <= 0 || C <= 0 || !f || !proto || !logits || !pred) { set_last_error("zero_shot_score: bad arguments"); return NGU_ERR_ARG; }
  if (dtype == NGU_F32) launch_pdl(zs_score_kernel<float>, dim3((B + 3) / 4), dim3(128), size_t(0), st, reinterpret_cast<const float*>(f), proto, logits, pred, B, E, C, scale);
  else launch_pdl(zs_score_kernel<bf16>, dim3((B + 3) / 4), dim3(128), size_t(0), st, reinterpret_cast<const bf16*>(f), proto, logits, pred, B, E, C, scale);
  return check_launch("zero_shot_score");
}

}  // namespace ngu
