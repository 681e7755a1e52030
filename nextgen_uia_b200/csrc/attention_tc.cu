// tcgen05 attention core for sm_100a: softmax(q k^T * scale) v, forward and backward, head dim 64,
// sequence length N <= 256 (ViT-B/16 @ 224: N = 197; BERT text tower: N = 77), packed timm layout
// qkv [B*N, 3*H*64].  Replaces F.scaled_dot_product_attention (timm Attention, pinned dep;
// src/adapters/lora.py:188-190) on the bf16 product path.
//
// One CTA per (batch, head).  Q/K/V(/dO) tiles are TMA-loaded once into 128-byte-swizzled smem as
// [rows, 64] tiles; the same tile serves as a K-major operand (rows = M or N, head dim = K) and as an
// MN-major operand (rows = K, head dim = N), so no transposes are ever materialised.
//
// Forward  (per 128-row query tile t):   S_t = Q_t K^T        (SS MMA, fp32 in TMEM, N = padded kv len)
//     one thread per query row: max / exp2 / sum straight out of TMEM, P written back to TMEM as bf16
//     (aliasing S),                       O_t = P_t V         (TS MMA: A from TMEM, B = V MN-major)
// Backward (per kv tile j, query tile i), "transposed" formulation so kv rows sit on TMEM lanes:
//     ST = K_j Q_i^T, dPT = V_j dO_i^T    (SS)   -> PT = exp2(ST*c - lse_i), dST = PT*(dPT - delta_i)*scale
//     dV_j += PT dO_i, dK_j += dST Q_i    (TS, A = PT / dST bf16 in TMEM, B MN-major)
//     dQ_i += dS K_j                      (SS, A = dST staged in smem as an MN-major operand, B = K_j MN-major)
//     dV/dK/dQ accumulate in TMEM across the loop and are written once (no atomics).
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

constexpr int DH = 64;
constexpr int TILE = 128;
constexpr int kTileBytes = TILE * DH * 2;  // 16 KB: 128 rows x 128 B
constexpr float kLog2e = 1.4426950408889634f;

struct AttnTcParams {
  CUtensorMap tmQKV;  // [B*N rows, 3*H*64 cols] bf16, box 64 cols x 128 rows, SWIZZLE_128B
  CUtensorMap tmDO;   // [B*N rows, H*64 cols]   bf16, same box (backward)
  bf16* o;            // [B*N, H*64]
  const bf16* o_in;   // backward: forward output
  const bf16* d_o;    // backward
  float* lse;         // [B, H, N]
  bf16* dqkv;         // [B*N, 3*H*64]
  int B, H, N;
  float scale;
};

NGU_DEVINL uint64_t desc_kmajor(uint32_t addr) { return make_smem_desc_sw128(addr, 16, 1024); }
// MN-major: K index = 128-byte row; SBO = 8 rows; LBO = stride between 64-element MN chunks
NGU_DEVINL uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo) { return make_smem_desc_sw128(addr, lbo, 1024); }

NGU_DEVINL void st_row_bf16(bf16* dst, const float (&v)[64]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
    u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
    u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
    u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
    reinterpret_cast<uint4*>(dst)[j] = u;
  }
}

// =====================================================================================================
// forward
// =====================================================================================================
constexpr int kFwdThreads = 32 * 5;  // 4 softmax warps (one query row per thread) + 1 control warp
constexpr int kFwdSmem = 5 * kTileBytes + 1024 + 1024;  // Q tile + K (2 tiles) + V (2 tiles)

// One CTA per (batch, head, 128-row query tile); 256 TMEM columns and ~82 KB smem so two CTAs share an SM and
// overlap each other's load / MMA / softmax phases.
__global__ void __launch_bounds__(kFwdThreads, 2) attn_fwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + kTileBytes, sV = base + 3 * kTileBytes;
  const uint32_t sBar = base + 5 * kTileBytes;
  const uint32_t bar_kv = sBar, bar_q = sBar + 8, bar_s = sBar + 16, bar_p = sBar + 24, bar_o = sBar + 32;
  const uint32_t sTmem = sBar + 40;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, D = p.H * DH;
  const int ntiles = (N + TILE - 1) / TILE;      // 1 or 2
  const int t = blockIdx.x % ntiles;             // query tile of this CTA
  const int bh = blockIdx.x / ntiles;
  const int b = bh / p.H, h = bh % p.H;
  const int npad = (N + 15) & ~15;               // MMA N extent over the kv axis
  const int row0 = b * N;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV);
    mbar_init(bar_kv, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(sTmem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, kTileBytes);
      tma_load_2d(sQ, &p.tmQKV, bar_q, h * DH, row0 + t * TILE);
      mbar_arrive_expect_tx(bar_kv, 2 * ntiles * kTileBytes);
      for (int u = 0; u < ntiles; ++u) {
        tma_load_2d(sK + u * kTileBytes, &p.tmQKV, bar_kv, D + h * DH, row0 + u * TILE);
        tma_load_2d(sV + u * kTileBytes, &p.tmQKV, bar_kv, 2 * D + h * DH, row0 + u * TILE);
      }
      // ---- S = Q K^T
      const uint32_t idesc_s = make_idesc_bf16(TILE, npad);
      mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_ss(tmem, desc_kmajor(sQ + k * 32), desc_kmajor(sK + k * 32), idesc_s, k != 0);
      umma_commit(bar_s);
      // ---- O = P V   (A = P in TMEM at columns [0, npad/2), D = O at column 128)
      constexpr uint32_t idesc_o = make_idesc_bf16(TILE, DH, 0, 1);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      for (int j = 0; j < npad / 16; ++j) umma_ts(tmem + 128, tmem + j * 8, desc_mnmajor(sV + j * 2048, 0), idesc_o, j != 0);
      umma_commit(bar_o);
    }
  } else {
    const int q = warp;
    const int r = t * TILE + q * 32 + lane;  // query row within the sequence
    const uint32_t trow = tmem + (uint32_t(q * 32) << 16);
    const float c = p.scale * kLog2e;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    const int nchunks = (npad + 31) / 32;
    float mx = -INFINITY;
    for (int ch = 0; ch < nchunks; ++ch) {
      uint32_t v[32];
      tmem_ld32(trow + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (ch * 32 + i < N) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    float sum = 0.f;
    const float mc = mx * c;
    for (int ch = 0; ch < nchunks; ++ch) {
      uint32_t v[32];
      tmem_ld32(trow + ch * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c0 = ch * 32 + 2 * i;
        float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, -mc));
        float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
        p0 = (c0 < N) ? p0 : 0.f;
        p1 = (c0 + 1 < N) ? p1 : 0.f;
        // the PV MMA sees bf16 probabilities: accumulate the same rounded values into the row sum
        const uint32_t w = pack_bf16x2(p0, p1);
        const float2 rr = unpack_bf16x2(w);
        sum += rr.x + rr.y;
        pk[i] = w;
      }
      tmem_st16(trow + ch * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    uint32_t ov[64];
    {
      uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&ov[0]);
      uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&ov[32]);
      tmem_ld32(trow + 128, lo);
      tmem_ld32(trow + 160, hi);
      tmem_ld_wait();
    }
    if (r < N) {
      const float inv = 1.f / sum;
      float of[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) of[i] = __uint_as_float(ov[i]) * inv;
      st_row_bf16(p.o + size_t(row0 + r) * D + h * DH, of);
      if (p.lse) p.lse[(size_t(b) * p.H + h) * N + r] = mx * p.scale + logf(sum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// =====================================================================================================
// backward
// =====================================================================================================
constexpr int kBwdThreads = 32 * 9;  // 8 compute warps (kv row = lane quarter, query columns split in halves) + 1 control warp
constexpr int kBwdSmem = 10 * kTileBytes + 2048 + 1024 + 1024;

__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + 2 * kTileBytes, sV = base + 4 * kTileBytes, sDO = base + 6 * kTileBytes;
  const uint32_t sDS = base + 8 * kTileBytes;                 // [2 q-chunks of 64][128 kv rows][128 B]
  const uint32_t sStat = base + 10 * kTileBytes;              // lse2[256], delta[256] (fp32)
  const uint32_t sBar = sStat + 2048;
  const uint32_t bar_load = sBar, bar_s = sBar + 8, bar_p = sBar + 16, bar_acc = sBar + 24;
  const uint32_t sTmem = sBar + 32;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* lse2 = reinterpret_cast<float*>(gen + 10 * kTileBytes);
  float* delta = lse2 + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int N = p.N, D = p.H * DH;
  const int ntiles = (N + TILE - 1) / TILE;
  const int row0 = b * N;
  // TMEM columns
  constexpr uint32_t cST = 0, cDPT = 128, cDV = 256, cDK = 320, cDQ = 384;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmDO);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(sTmem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_load, 4 * ntiles * kTileBytes);
      for (int t = 0; t < ntiles; ++t) {
        tma_load_2d(sQ + t * kTileBytes, &p.tmQKV, bar_load, h * DH, row0 + t * TILE);
        tma_load_2d(sK + t * kTileBytes, &p.tmQKV, bar_load, D + h * DH, row0 + t * TILE);
        tma_load_2d(sV + t * kTileBytes, &p.tmQKV, bar_load, 2 * D + h * DH, row0 + t * TILE);
        tma_load_2d(sDO + t * kTileBytes, &p.tmDO, bar_load, h * DH, row0 + t * TILE);
      }
      mbar_wait(bar_load, 0);
      constexpr uint32_t idesc_st = make_idesc_bf16(TILE, TILE);          // ST / dPT: K-major x K-major
      constexpr uint32_t idesc_ts = make_idesc_bf16(TILE, DH, 0, 1);      // dV / dK: A in TMEM, B MN-major
      constexpr uint32_t idesc_dq = make_idesc_bf16(TILE, DH, 1, 1);      // dQ: A MN-major (smem), B MN-major
      uint32_t ph = 0;
      for (int j = 0; j < ntiles; ++j) {
        for (int i = 0; i < ntiles; ++i) {
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_ss(tmem + cST, desc_kmajor(sK + j * kTileBytes + k * 32), desc_kmajor(sQ + i * kTileBytes + k * 32), idesc_st, k != 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_ss(tmem + cDPT, desc_kmajor(sV + j * kTileBytes + k * 32), desc_kmajor(sDO + i * kTileBytes + k * 32), idesc_st, k != 0);
          umma_commit(bar_s);
          mbar_wait(bar_p, ph);
          tc_fence_after();
#pragma unroll
          for (int s = 0; s < TILE / 16; ++s) {
            umma_ts(tmem + cDV, tmem + cST + s * 8, desc_mnmajor(sDO + i * kTileBytes + s * 2048, 0), idesc_ts, (i | s) != 0);
            umma_ts(tmem + cDK, tmem + cDPT + s * 8, desc_mnmajor(sQ + i * kTileBytes + s * 2048, 0), idesc_ts, (i | s) != 0);
            umma_ss(tmem + cDQ + i * 64, desc_mnmajor(sDS + s * 2048, kTileBytes), desc_mnmajor(sK + j * kTileBytes + s * 2048, 0),
                    idesc_dq, (j | s) != 0);
          }
          ph ^= 1u;
        }
        umma_commit(bar_acc);  // dV_j / dK_j (and, after the last j, dQ) complete
      }
    }
  } else {
    const int qd = warp & 3, hf = warp >> 2;     // TMEM lane quarter, query-column half
    const int t = qd * 32 + lane;                // kv row within the tile (= TMEM lane)
    // ---- prologue: lse (log2 domain) and delta = rowsum(dO * O) for every query row (one row per thread)
    {
      const int r = threadIdx.x;
      if (r < ntiles * TILE) {
        float l2 = 0.f, dl = 0.f;
        if (r < N) {
          l2 = p.lse[(size_t(b) * p.H + h) * N + r] * kLog2e;
          const uint4* po = reinterpret_cast<const uint4*>(p.o_in + size_t(row0 + r) * D + h * DH);
          const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + size_t(row0 + r) * D + h * DH);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 a = __ldg(po + j), g = __ldg(pd + j);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(gw[e]);
              dl = fmaf(x.x, y.x, dl);
              dl = fmaf(x.y, y.y, dl);
            }
          }
        }
        lse2[r] = l2;
        delta[r] = dl;
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t trow = tmem + (uint32_t(qd * 32) << 16);
    const float c = p.scale * kLog2e;
    uint32_t ph = 0, acc_ph = 0;
    for (int j = 0; j < ntiles; ++j) {
      const int kv = j * TILE + t;
      const bool kv_ok = kv < N;
      const bool warp_live = j * TILE + qd * 32 < N;   // any valid kv row in this warp (warp-uniform)
      for (int i = 0; i < ntiles; ++i) {
        mbar_wait(bar_s, ph);
        tc_fence_after();
        for (int cc = 0; cc < 2; ++cc) {  // this warp's two 32-column chunks of the 128 query columns
          const int ch = hf * 2 + cc;
          const bool chunk_live = warp_live && (i * TILE + ch * 32 < N);
          uint32_t pp[16], ds[16];
          if (chunk_live) {
            uint32_t sv[32], dv[32];
            tmem_ld32(trow + cST + ch * 32, sv);
            tmem_ld32(trow + cDPT + ch * 32, dv);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int q0 = i * TILE + ch * 32 + 2 * e;
              float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), c, -lse2[q0]));
              float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), c, -lse2[q0 + 1]));
              p0 = (kv_ok && q0 < N) ? p0 : 0.f;
              p1 = (kv_ok && q0 + 1 < N) ? p1 : 0.f;
              const float d0 = p0 * (__uint_as_float(dv[2 * e]) - delta[q0]) * p.scale;
              const float d1 = p1 * (__uint_as_float(dv[2 * e + 1]) - delta[q0 + 1]) * p.scale;
              pp[e] = pack_bf16x2(p0, p1);
              ds[e] = pack_bf16x2(d0, d1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) { pp[e] = 0u; ds[e] = 0u; }
          }
          // PT / dST (bf16) alias the ST / dPT columns.  Chunk ch writes columns [16ch, 16ch+16): for hf = 1 those
          // are columns 32..63 = fp32 chunk 1 of the OTHER warp half -> order the two halves with a named barrier.
          if (cc == 0 && hf == 1) asm volatile("bar.sync 2, 256;" ::: "memory");
          tmem_st16(trow + cST + ch * 16, pp);
          tmem_st16(trow + cDPT + ch * 16, ds);
          if (cc == 1 && hf == 0) { tmem_st_wait(); asm volatile("bar.arrive 2, 256;" ::: "memory"); }
          // dS^T row of this kv index into the MN-major smem operand: 32 q values = 4 x 16 B
          const uint32_t rbase = sDS + (ch >> 1) * kTileBytes + (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
          for (int piece = 0; piece < 4; ++piece) {
            const uint32_t idx = uint32_t((ch & 1) * 4 + piece);
            const uint32_t a = rbase + ((idx ^ uint32_t(t & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(ds[4 * piece]), "r"(ds[4 * piece + 1]),
                         "r"(ds[4 * piece + 2]), "r"(ds[4 * piece + 3])
                         : "memory");
          }
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_p);
        ph ^= 1u;
      }
      // ---- dV_j, dK_j complete: each warp half writes 32 of the 64 head-dim columns
      mbar_wait(bar_acc, acc_ph);
      acc_ph ^= 1u;
      tc_fence_after();
      uint32_t v[32];
      auto store32 = [&](bf16* dst) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[8 * jj + 0]), __uint_as_float(v[8 * jj + 1]));
          u.y = pack_bf16x2(__uint_as_float(v[8 * jj + 2]), __uint_as_float(v[8 * jj + 3]));
          u.z = pack_bf16x2(__uint_as_float(v[8 * jj + 4]), __uint_as_float(v[8 * jj + 5]));
          u.w = pack_bf16x2(__uint_as_float(v[8 * jj + 6]), __uint_as_float(v[8 * jj + 7]));
          reinterpret_cast<uint4*>(dst)[jj] = u;
        }
      };
      tmem_ld32(trow + cDV + hf * 32, v);
      tmem_ld_wait();
      if (kv_ok) store32(p.dqkv + size_t(row0 + kv) * 3 * D + 2 * D + h * DH + hf * 32);
      tmem_ld32(trow + cDK + hf * 32, v);
      tmem_ld_wait();
      if (kv_ok) store32(p.dqkv + size_t(row0 + kv) * 3 * D + D + h * DH + hf * 32);
      if (j == ntiles - 1) {
        for (int i = 0; i < ntiles; ++i) {
          const int qr = i * TILE + t;
          tmem_ld32(trow + cDQ + i * 64 + hf * 32, v);
          tmem_ld_wait();
          if (qr < N) store32(p.dqkv + size_t(row0 + qr) * 3 * D + h * DH + hf * 32);
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// packed-layout check: q/k/v views of one [B*N, 3D] buffer, o [B*N, D]
bool is_packed(const ngu_attn_desc& d) {
  const int64_t D = int64_t(d.H) * d.dh;
  const char* q = reinterpret_cast<const char*>(d.q);
  return d.N == d.S && !d.causal && d.q_ts == 3 * D && d.k_ts == 3 * D && d.v_ts == 3 * D && d.o_ts == D &&
         d.q_bs == int64_t(d.N) * 3 * D && d.k_bs == d.q_bs && d.v_bs == d.q_bs && d.o_bs == int64_t(d.N) * D &&
         reinterpret_cast<const char*>(d.k) == q + D * 2 && reinterpret_cast<const char*>(d.v) == q + 4 * D;
}

int fill_params(const ngu_attn_desc& d, AttnTcParams& p, bool bwd) {
  memset(&p, 0, sizeof(p));
  const int D = d.H * d.dh;
  int rc;
  if ((rc = make_tmap_2d_bf16(&p.tmQKV, d.q, uint64_t(d.B) * d.N, 3 * D, 3 * D, TILE, DH, true))) return rc;
  if (bwd) {
    if ((rc = make_tmap_2d_bf16(&p.tmDO, d.d_o, uint64_t(d.B) * d.N, D, D, TILE, DH, true))) return rc;
  } else {
    p.tmDO = p.tmQKV;
  }
  p.o = reinterpret_cast<bf16*>(d.o);
  p.o_in = reinterpret_cast<const bf16*>(d.o);
  p.d_o = reinterpret_cast<const bf16*>(d.d_o);
  p.lse = d.lse;
  p.dqkv = reinterpret_cast<bf16*>(d.dq);
  p.B = d.B; p.H = d.H; p.N = d.N;
  p.scale = d.scale;
  return NGU_OK;
}

}  // namespace

bool attn_tc_supported(const ngu_attn_desc& d, bool bwd) {
  if (d.dtype != NGU_BF16 || d.dh != DH || d.N > 2 * TILE || !is_packed(d)) return false;
  if (bwd) {
    const int64_t D = int64_t(d.H) * d.dh;
    const char* dq = reinterpret_cast<const char*>(d.dq);
    if (reinterpret_cast<const char*>(d.dk) != dq + D * 2 || reinterpret_cast<const char*>(d.dv) != dq + D * 4) return false;
  }
  return true;
}

int attn_fwd_tc(const ngu_attn_desc& d, cudaStream_t st) {
  AttnTcParams p;
  if (int rc = fill_params(d, p, false)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return cuda_status(e, "attn_fwd_tc attr");
    attr = true;
  }
  attn_fwd_tc_kernel<<<d.B * d.H * ((d.N + TILE - 1) / TILE), kFwdThreads, kFwdSmem, st>>>(p);
  return check_launch("attn_fwd_tc");
}

int attn_bwd_tc(const ngu_attn_desc& d, cudaStream_t st) {
  AttnTcParams p;
  if (int rc = fill_params(d, p, true)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    if (e != cudaSuccess) return cuda_status(e, "attn_bwd_tc attr");
    attr = true;
  }
  attn_bwd_tc_kernel<<<d.B * d.H, kBwdThreads, kBwdSmem, st>>>(p);
  return check_launch("attn_bwd_tc");
}

}  // namespace ngu
