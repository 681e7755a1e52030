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
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

constexpr int DH = 64;
constexpr int TILE = 128;
constexpr int kTileBytes = TILE * DH * 2;  // 16 KB: 128 rows x 128 B
constexpr float kLog2e = 1.4426950408889634f;

#ifdef NGU_ATTN_TRACE
// Per-warp private event slots (no atomics: a returning atomic would stall the traced warp for ~1 us per event).
// Layout: [warp][256 events][code, arg, clock]
__device__ unsigned long long g_attn_trace[32 * 256 * 3];
NGU_DEVINL void trace_ev(uint32_t& n, int code, unsigned a) {
  if (blockIdx.x != 0 || n >= 256) return;
  unsigned long long* e = g_attn_trace + (size_t(threadIdx.x >> 5) * 256 + n) * 3;
  e[0] = code; e[1] = a; e[2] = clock64();
  ++n;
}
#define TRACE_DECL uint32_t tr_n = 0
#define TRACE(code, a) do { if ((threadIdx.x & 31) == 0) trace_ev(tr_n, code, a); } while (0)
#else
#define TRACE_DECL
#define TRACE(code, a) do {} while (0)
#endif

struct AttnTcParams {
  CUtensorMap tmQKV;  // [B*N rows, 3*H*64 cols] bf16, box 64 cols x 128 rows, SWIZZLE_128B
  CUtensorMap tmDO;   // [B*N rows, H*64 cols]   bf16, same box (backward)
  CUtensorMap tmQKV1, tmDO1;  // backward, N > 128: boxes of ceil16(N - 128) rows for the second tile
  CUtensorMap tmO3;           // forward output as [B, N, H*64], box 1 x 128 x 64: rows past N are clipped by the TMA store
  bf16* o;            // [B*N, H*64]
  const bf16* o_in;   // backward: forward output
  const bf16* d_o;    // backward
  float* lse;         // [B, H, N]
  bf16* dqkv;         // [B*N, 3*H*64]
  int B, H, N;
  float scale;
  const int* kv_len;  // forward: per-batch number of valid keys (NULL = N)
  int causal;         // forward (per-tile kernel): key j is visible to query i iff j <= i (CLIP text tower, model.py:344-350)
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
constexpr int kFwdSoftmaxWarps = 8;
constexpr int kFwdThreads = 32 * (kFwdSoftmaxWarps + 1);  // 8 softmax warps (two threads per query row) + 1 control warp
constexpr int kFwdSmem = 5 * kTileBytes + 2048 + 1024 + 1024;  // Q tile + K (2 tiles) + V (2 tiles) + pair-exchange floats

// One CTA per (batch, head, 128-row query tile); 256 TMEM columns and ~83 KB smem so two CTAs share an SM and
// overlap each other's load / MMA / softmax phases.  Softmax: warps w and w+4 share TMEM lane quarter w&3 and split
// the kv columns of a row (32-column chunks [0, nc0) and [nc0, nch)); row max and row sum are exchanged through smem.
// P (bf16) aliases the S columns, so every thread keeps its P values in registers until BOTH threads of the row have
// finished reading S.
__global__ void __launch_bounds__(kFwdThreads, 2) attn_fwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + kTileBytes, sV = base + 3 * kTileBytes;
  const uint32_t sRed = base + 5 * kTileBytes;       // max [2 halves][128 rows], sum [2][128] fp32
  const uint32_t sBar = sRed + 2048;
  const uint32_t bar_kv = sBar, bar_q = sBar + 8, bar_s = sBar + 16, bar_p = sBar + 24, bar_o = sBar + 32;
  const uint32_t sTmem = sBar + 40;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TRACE_DECL;
  const int N = p.N, D = p.H * DH;
  const int ntiles = (N + TILE - 1) / TILE;      // 1 or 2
  const int t = blockIdx.x % ntiles;             // query tile of this CTA
  const int bh = blockIdx.x / ntiles;
  const int b = bh / p.H, h = bh % p.H;
  const int npad = (N + 15) & ~15;               // MMA N extent over the kv axis
  const int row0 = b * N;
  int Lk = p.kv_len ? __ldg(p.kv_len + b) : N;   // valid keys of this sequence (key-padding mask = suffix of the row)
  Lk = Lk < 1 ? 1 : (Lk > N ? N : Lk);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV);
    mbar_init(bar_kv, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 32 * kFwdSoftmaxWarps);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == kFwdSoftmaxWarps) {
    tmem_alloc(sTmem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));

  if (warp == kFwdSoftmaxWarps) {
    // whole warp walks (warp-uniform values stay in uniform registers); one elected lane issues
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, kTileBytes);
      tma_load_2d(sQ, &p.tmQKV, bar_q, h * DH, row0 + t * TILE);
      mbar_arrive_expect_tx(bar_kv, 2 * ntiles * kTileBytes);
      for (int u = 0; u < ntiles; ++u) {
        tma_load_2d(sK + u * kTileBytes, &p.tmQKV, bar_kv, D + h * DH, row0 + u * TILE);
        tma_load_2d(sV + u * kTileBytes, &p.tmQKV, bar_kv, 2 * D + h * DH, row0 + u * TILE);
      }
    }
    __syncwarp();
    // ---- S = Q K^T
    const uint32_t idesc_s = make_idesc_bf16(TILE, npad);
    const uint64_t dq = desc_kmajor(sQ), dk = desc_kmajor(sK), dv = desc_mnmajor(sV, 0);
    mbar_wait(bar_q, 0);
    mbar_wait(bar_kv, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_ss(tmem, dq + uint64_t(k * 2), dk + uint64_t(k * 2), idesc_s, k != 0);
      umma_commit(bar_s);
    }
    __syncwarp();
    // ---- O = P V   (A = P in TMEM at columns [0, npad/2), D = O at column 128)
    constexpr uint32_t idesc_o = make_idesc_bf16(TILE, DH, 0, 1);
    const int nsl = npad / 16;
    mbar_wait(bar_p, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int j = 0; j < 16; ++j) umma_ts_if(j < nsl, tmem + 128, tmem + j * 8, dv + uint64_t(j * 128), idesc_o, j != 0);
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    const int q = warp & 3, hf = warp >> 2;
    const int rt = q * 32 + lane;              // row within the tile (= TMEM lane)
    const int r = t * TILE + rt;               // query row within the sequence
    const bool live = t * TILE + q * 32 < N;   // warp-uniform: any valid query row in this warp (same for both halves)
    const uint32_t trow = tmem + (uint32_t(q * 32) << 16);
    const float c = p.scale * kLog2e;
    const int nch = (npad + 31) / 32;          // 32-column chunks of the kv axis (<= 8)
    const int nc0 = (nch + 1) / 2;             // half 0: chunks [0, nc0), half 1: [nc0, nch)
    const int cb = hf ? nc0 : 0, ce = hf ? nch : nc0;
    const uint32_t my_red = sRed + 4u * uint32_t(hf * 128 + rt), other_red = sRed + 4u * uint32_t((hf ^ 1) * 128 + rt);
    float sum = 0.f, mx = -INFINITY;
    if (p.causal) Lk = min(Lk, min(r, N - 1) + 1);   // per-row key limit: the causal mask is one more upper bound on the key index
    if (live) {
      mbar_wait(bar_s, 0);
      tc_fence_after();
      // ---- pass 1: row max over this thread's columns
      for (int ch = cb; ch < ce; ++ch) {
        uint32_t v[32];
        tmem_ld32(trow + ch * 32, v);
        tmem_ld_wait();
        if (ch * 32 + 32 <= Lk) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ch * 32 + i < Lk) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_red), "f"(mx) : "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      float omx;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(omx) : "r"(other_red));
      mx = fmaxf(mx, omx);
      const float mc = mx * c;
      // ---- pass 2: P = exp2(S*c - max*c) kept in registers (bf16 pairs), row sum of this thread's columns
      uint32_t pk[4][16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ch = cb + k;
        if (ch < ce) {
          uint32_t v[32];
          tmem_ld32(trow + ch * 32, v);
          tmem_ld_wait();
          const bool full = ch * 32 + 32 <= Lk;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, -mc));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
            if (!full) {
              p0 = (ch * 32 + 2 * i < Lk) ? p0 : 0.f;
              p1 = (ch * 32 + 2 * i + 1 < Lk) ? p1 : 0.f;
            }
            sum += p0 + p1;
            pk[k][i] = pack_bf16x2(p0, p1);
          }
        }
      }
      // both threads of the row are done reading S (and publish their partial sums) before P overwrites it
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_red + 1024u), "f"(sum) : "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      float osum;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(osum) : "r"(other_red + 1024u));
      sum += osum;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (cb + k < ce) tmem_st16(trow + (cb + k) * 16, pk[k]);
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(bar_p);
    if (live) {
      mbar_wait(bar_o, 0);
      tc_fence_after();
      uint32_t ov[32];
      tmem_ld32(trow + 128 + hf * 32, ov);
      tmem_ld_wait();
      // O tile staged in the (now dead) Q tile, 128B-swizzled rows of 64 head-dim values, and written with ONE TMA store:
      // per-lane row stores (32 distinct lines per instruction) made the LSU the bottleneck of this phase.
      {
        const float inv = rcp_approx(sum);
        const uint32_t rb = sQ + uint32_t(rt) * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t a = rb + ((uint32_t(hf * 4 + j) ^ uint32_t(rt & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                       "r"(pack_bf16x2(__uint_as_float(ov[8 * j + 0]) * inv, __uint_as_float(ov[8 * j + 1]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(ov[8 * j + 2]) * inv, __uint_as_float(ov[8 * j + 3]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(ov[8 * j + 4]) * inv, __uint_as_float(ov[8 * j + 5]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(ov[8 * j + 6]) * inv, __uint_as_float(ov[8 * j + 7]) * inv))
                       : "memory");
        }
        if (r < N && p.lse && hf == 0) p.lse[(size_t(b) * p.H + h) * N + r] = fmaf(mx, p.scale, 0.6931471805599453f * lg2_approx(sum));
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 9, %0;" ::"r"(32 * kFwdSoftmaxWarps) : "memory");   // all softmax warps (dead ones included)
    if (warp == 0 && lane == 0) {
      tma_store_3d(&p.tmO3, sQ, h * DH, t * TILE, b);
      tma_store_commit();
      tma_store_wait_read<0>();   // smem must outlive the store's read; global visibility comes with kernel completion
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kFwdSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}


// ---- persistent forward ---------------------------------------------------------------------------------
// One CTA per SM walks (batch, head) items; a "unit" is one 128-row query tile of an item.  Roles:
//   warps 0-7   softmax group 0 (units 0, 2, 4, ...; TMEM slot 0)     warps 8-15  softmax group 1 (odd units; slot 1)
//               two threads per query row (warps w and w+4 of a group share TMEM lane quarter w&3)
//   warp 16     MMA issuer: event loop over both groups (S = Q K^T as soon as the slot is drained and the tiles have
//               landed, O = P V as soon as the group has written P)
//   warp 17     TMA loader: K/V of item n+1 (double-buffered) and the Q tiles of the next units (4-deep ring)
// Each group runs its own serial chain (S -> softmax -> P -> O -> store) and the two chains interleave on the SM, as
// two co-resident CTAs would, but K/V are fetched once per item and nothing is re-initialised between units.
constexpr int kFwdGroupWarps = 8;
constexpr int kFwdPThreads = 32 * (2 * kFwdGroupWarps + 2);
constexpr int kFwdPSmem = 12 * kTileBytes + 2 * 2048 + 1024 + 1024;   // K/V 2 x 4 tiles, Q 4 tiles, exchange, barriers

__global__ void __launch_bounds__(kFwdPThreads, 1) attn_fwd_persistent_kernel(const __grid_constant__ AttnTcParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sK = [&](int buf, int u) { return base + uint32_t(buf * 4 + u) * kTileBytes; };
  auto sV = [&](int buf, int u) { return base + uint32_t(buf * 4 + 2 + u) * kTileBytes; };
  auto sQ = [&](int slot) { return base + uint32_t(8 + slot) * kTileBytes; };
  const uint32_t sRed = base + 12 * kTileBytes;       // per group: max [2 halves][128 rows], sum [2][128] fp32
  const uint32_t sBar = sRed + 2 * 2048;
  auto bar_kv = [&](int b) { return sBar + 8u * b; };            // 2: K/V of item parity b landed
  auto bar_kvfree = [&](int b) { return sBar + 16u + 8u * b; };  // 2: all PV MMAs of that item complete
  auto bar_q = [&](int s) { return sBar + 32u + 8u * s; };       // 4: Q tile of unit (k & 3) landed
  auto bar_qfree = [&](int s) { return sBar + 64u + 8u * s; };   // 4: S MMAs of that unit complete
  auto bar_s = [&](int g) { return sBar + 96u + 8u * g; };       // 2: S of group g's unit complete
  auto bar_p = [&](int g) { return sBar + 112u + 8u * g; };      // 2: P written (256 arrivals)
  auto bar_o = [&](int g) { return sBar + 128u + 8u * g; };      // 2: O complete
  auto bar_drained = [&](int g) { return sBar + 144u + 8u * g; };  // 2: O read out (256 arrivals)
  const uint32_t sTmem = sBar + 160u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TRACE_DECL;
  const int N = p.N, D = p.H * DH;
  const int ntiles = (N + TILE - 1) / TILE;      // 1 or 2
  const int npad = (N + 15) & ~15;               // MMA N extent over the kv axis
  const int rows1 = ntiles == 2 ? ((N - TILE + 15) & ~15) : 0;
  const int items = p.B * p.H;
  const int n_local = (items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int n_units = n_local * ntiles;
  constexpr int kGroup = 32 * kFwdGroupWarps;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_kv(b), 1); mbar_init(bar_kvfree(b), 1);
      mbar_init(bar_s(b), 1); mbar_init(bar_p(b), kGroup); mbar_init(bar_o(b), 1); mbar_init(bar_drained(b), kGroup);
    }
    for (int s = 0; s < 4; ++s) { mbar_init(bar_q(s), 1); mbar_init(bar_qfree(s), 1); }
    fence_mbar_init();
  }
  if (warp == 2 * kFwdGroupWarps) {
    tmem_alloc(sTmem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));

  if (warp == 2 * kFwdGroupWarps + 1) {
    // ================================ TMA loader ================================
    if (lane == 0) {
      for (int k = 0; k < n_units; ++k) {
        const int n = ntiles == 2 ? k >> 1 : k, t = ntiles == 2 ? k & 1 : 0;
        const int item = int(blockIdx.x) + n * int(gridDim.x);
        const int b = item / p.H, h = item - b * p.H;
        const int row0 = b * N;
        if (t == 0) {
          const int kb = n & 1;
          if (n >= 2) mbar_wait(bar_kvfree(kb), uint32_t((n >> 1) - 1) & 1u);
          mbar_arrive_expect_tx(bar_kv(kb), 2 * kTileBytes + 2 * rows1 * 128);
          tma_load_2d(sK(kb, 0), &p.tmQKV, bar_kv(kb), D + h * DH, row0);
          tma_load_2d(sV(kb, 0), &p.tmQKV, bar_kv(kb), 2 * D + h * DH, row0);
          if (ntiles == 2) {
            tma_load_2d(sK(kb, 1), &p.tmQKV1, bar_kv(kb), D + h * DH, row0 + TILE);
            tma_load_2d(sV(kb, 1), &p.tmQKV1, bar_kv(kb), 2 * D + h * DH, row0 + TILE);
          }
        }
        const int qs = k & 3;
        if (k >= 4) mbar_wait(bar_qfree(qs), uint32_t((k >> 2) - 1) & 1u);
        mbar_arrive_expect_tx(bar_q(qs), t ? rows1 * 128 : kTileBytes);
        tma_load_2d(sQ(qs), t ? &p.tmQKV1 : &p.tmQKV, bar_q(qs), h * DH, row0 + t * TILE);
      }
    }
  } else if (warp == 2 * kFwdGroupWarps) {
    // ================================ MMA issuer ================================
    // Per group: state 0 = S part A of the next unit not issued, 1 = part B (and the commit) pending, 2 = waiting for P.
    // TMEM slot layout: S in columns [0, npad), P (bf16) aliases [0, npad/2), O in [192, 256).  S therefore overlaps the
    // previous unit's O only in columns >= 192: part A (kv columns < 192) is issued right after the previous PV, part B
    // (kv columns 192..npad, at most 64) once the group has read O out.  Whole warp walks; one lane issues.
    const int na = npad < 192 ? npad : 192, nb = npad - na;
    const uint32_t idesc_a = make_idesc_bf16(TILE, na), idesc_b = make_idesc_bf16(TILE, nb > 0 ? nb : 16);
    constexpr uint32_t idesc_o = make_idesc_bf16(TILE, DH, 0, 1);
    const int nsl = npad / 16;
    int unit[2] = {0, 1};
    int state[2] = {0, 0};
    int remaining = n_units;      // units whose PV has not been issued yet
    int pv_issued[2] = {0, 0};    // per K/V buffer: PV MMAs issued for the item it holds (the groups run at their own pace)
    while (remaining > 0) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int k = unit[g];
        if (k >= n_units) continue;
        const int n = ntiles == 2 ? k >> 1 : k;
        const int kb = n & 1, qs = k & 3;
        const uint32_t slot = tmem + uint32_t(g) * 256u;
        if (state[g] == 0) {
          const bool ready = mbar_try_wait(bar_kv(kb), (n >> 1) & 1) && mbar_try_wait(bar_q(qs), (k >> 2) & 1);
          if (__all_sync(0xffffffffu, ready)) {
            tc_fence_after();
            const uint64_t dq = desc_kmajor(sQ(qs)), dk = desc_kmajor(sK(kb, 0));
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < DH / 16; ++kk) umma_ss(slot, dq + uint64_t(kk * 2), dk + uint64_t(kk * 2), idesc_a, kk != 0);
              if (nb == 0) { umma_commit(bar_s(g)); umma_commit(bar_qfree(qs)); }
            }
            __syncwarp();
            TRACE(60 + g, k);
            state[g] = nb == 0 ? 2 : 1;
          }
        } else if (state[g] == 1) {
          const bool ready = k < 2 || mbar_try_wait(bar_drained(g), uint32_t((k >> 1) - 1) & 1u);
          if (__all_sync(0xffffffffu, ready)) {
            tc_fence_after();
            // kv rows 192.. : K tile 1, rows 64.. (8 KB in), S columns 192..
            const uint64_t dq = desc_kmajor(sQ(qs)), dk = desc_kmajor(sK(kb, 1) + 8192);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < DH / 16; ++kk) umma_ss(slot + 192, dq + uint64_t(kk * 2), dk + uint64_t(kk * 2), idesc_b, kk != 0);
              umma_commit(bar_s(g));
              umma_commit(bar_qfree(qs));
            }
            __syncwarp();
            TRACE(64 + g, k);
            state[g] = 2;
          }
        } else {
          if (__all_sync(0xffffffffu, mbar_try_wait(bar_p(g), (k >> 1) & 1))) {
            tc_fence_after();
            const uint64_t dv = desc_mnmajor(sV(kb, 0), 0);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 16; ++j) umma_ts_if(j < nsl, slot + 192, slot + j * 8, dv + uint64_t(j * 128), idesc_o, j != 0);
              umma_commit(bar_o(g));
              if (pv_issued[kb] + 1 == ntiles) umma_commit(bar_kvfree(kb));   // the item's last PV (of either group)
            }
            __syncwarp();
            TRACE(62 + g, k);
            pv_issued[kb] = (pv_issued[kb] + 1 == ntiles) ? 0 : pv_issued[kb] + 1;
            state[g] = 0;
            unit[g] = k + 2;
            --remaining;
          }
        }
      }
    }
  } else {
    // ================================ softmax groups ================================
    const int g = warp / kFwdGroupWarps, w = warp % kFwdGroupWarps;
    const int q = w & 3, hf = w >> 2;
    const int rt = q * 32 + lane;              // row within the tile (= TMEM lane)
    const uint32_t trow = tmem + uint32_t(g) * 256u + (uint32_t(q * 32) << 16);
    const float c = p.scale * kLog2e;
    const int nch = (npad + 31) / 32;          // 32-column chunks of the kv axis (<= 8)
    const int nc0 = (nch + 1) / 2;             // half 0: chunks [0, nc0), half 1: [nc0, nch)
    const int cb = hf ? nc0 : 0, ce = hf ? nch : nc0;
    const uint32_t red = sRed + uint32_t(g) * 2048u;
    const uint32_t my_red = red + 4u * uint32_t(hf * 128 + rt), other_red = red + 4u * uint32_t((hf ^ 1) * 128 + rt);
    const int pair_bar = 1 + g * 4 + q;
    uint32_t cnt = 0;
    for (int k = g; k < n_units; k += 2, ++cnt) {
      const int n = ntiles == 2 ? k >> 1 : k, t = ntiles == 2 ? k & 1 : 0;
      const int item = int(blockIdx.x) + n * int(gridDim.x);
      const int b = item / p.H, h = item - b * p.H;
      const int row0 = b * N;
      const int r = t * TILE + rt;               // query row within the sequence
      const bool live = t * TILE + q * 32 < N;   // warp-uniform: any valid query row in this warp (same for both halves)
      float sum = 0.f, mx = -INFINITY;
      // dead warps (all rows past the sequence end) skip the math but keep in step with the barriers' phases
      if (w == 0) TRACE(40 + g, k);
      mbar_wait(bar_s(g), cnt & 1u);
      if (w == 0) TRACE(42 + g, k);
      tc_fence_after();
      if (live) {
        // ---- pass 1: row max over this thread's columns (four independent running maxima)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int ch = cb; ch < ce; ++ch) {
          uint32_t v[32];
          tmem_ld32(trow + ch * 32, v);
          tmem_ld_wait();
          if (ch * 32 + 32 <= N) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
#pragma unroll
              for (int e = 0; e < 4; ++e) m4[e] = fmaxf(m4[e], fmaxf(__uint_as_float(v[i + 2 * e]), __uint_as_float(v[i + 2 * e + 1])));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (ch * 32 + i < N) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
          }
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        if (w == 0) TRACE(44 + g, k);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_red), "f"(mx) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        float omx;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(omx) : "r"(other_red));
        mx = fmaxf(mx, omx);
        const float mc = mx * c;
        // ---- pass 2: P = exp2(S*c - max*c) kept in registers (bf16 pairs), row sum of this thread's columns
        uint32_t pk[4][16];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int ch = cb + kk;
          if (ch < ce) {
            uint32_t v[32];
            tmem_ld32(trow + ch * 32, v);
            tmem_ld_wait();
            const bool full = ch * 32 + 32 <= N;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, -mc));
              float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
              if (!full) {
                p0 = (ch * 32 + 2 * i < N) ? p0 : 0.f;
                p1 = (ch * 32 + 2 * i + 1 < N) ? p1 : 0.f;
              }
              sum += p0 + p1;
              pk[kk][i] = pack_bf16x2(p0, p1);
            }
          }
        }
        // both threads of the row are done reading S (and publish their partial sums) before P overwrites it
        if (w == 0) TRACE(46 + g, k);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_red + 1024u), "f"(sum) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        float osum;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(osum) : "r"(other_red + 1024u));
        sum += osum;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (cb + kk < ce) tmem_st16(trow + (cb + kk) * 16, pk[kk]);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_p(g));
      if (w == 0) TRACE(48 + g, k);
      mbar_wait(bar_o(g), cnt & 1u);
      if (w == 0) TRACE(50 + g, k);
      tc_fence_after();
      if (live) {
        uint32_t ov[32];
        tmem_ld32(trow + 192 + hf * 32, ov);
        tmem_ld_wait();
        if (r < N) {
          const float inv = rcp_approx(sum);
          bf16* dst = p.o + size_t(row0 + r) * D + h * DH + hf * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[8 * j + 0]) * inv, __uint_as_float(ov[8 * j + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(ov[8 * j + 2]) * inv, __uint_as_float(ov[8 * j + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(ov[8 * j + 4]) * inv, __uint_as_float(ov[8 * j + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(ov[8 * j + 6]) * inv, __uint_as_float(ov[8 * j + 7]) * inv);
            reinterpret_cast<uint4*>(dst)[j] = u;
          }
          if (p.lse && hf == 0) p.lse[(size_t(b) * p.H + h) * N + r] = fmaf(mx, p.scale, 0.6931471805599453f * lg2_approx(sum));
        }
      }
      tc_fence_before();
      mbar_arrive(bar_drained(g));
      if (w == 0) TRACE(52 + g, k);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2 * kFwdGroupWarps) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---- pipelined forward ------------------------------------------------------------------------------------
// One CTA per SM walks (batch, head) items; a "unit" is one 128-row query tile of an item.  The serial chain of the per-tile
// kernel (load -> S -> softmax -> P -> O -> store, 13 k cycles per item and SM with every pipe under 30 % busy) is cut into two
// roles that run one unit apart, with S double-buffered in TMEM so the tensor pipe works a unit ahead of the softmax:
//   warps 0-15   softmax: FOUR threads per query row (warps w, w+4, w+8, w+12 share TMEM lane quarter w&3 and split the kv axis
//                in runs of 8-column chunks).  A thread reads its run of S ONCE (<= 64 fp32 registers), exchanges the row max
//                through smem, and writes P = exp2(..) back as bf16 over the S columns; partial row sums go to smem.
//   warps 16-19  epilogue: one thread per query row reads O, applies 1/rowsum, stages the tile in the unit's (dead) Q slot for a
//                TMA store and writes the log-sum-exp.  Warp 19 (lane quarter 3, idle on a short second tile) is also the control warp: its elected lane issues O(u) = P(u) V when
//                P(u) is written, S(u+2) when O(u) has left the buffer, and the TMA loads (Q three units ahead, K/V of item
//                m + kvbufs when item m's last O is complete).  One thread issuing stores and loads orders them without barriers.
// 20 warps = 5 per scheduler: 96 registers per thread, so the S run stays in registers (22 warps would cap at 80 and spill it).
// TMEM: two buffers of 256 columns: S in [0, npad), P (bf16) aliases [0, npad/2), O in [128, 192) (over dead S columns).
// Key padding (kv_len) and the causal mask are one upper bound on the key index per row, applied as -inf before the row max.
constexpr int kFwd3SoftmaxWarps = 16;
constexpr int kFwd3Threads = 32 * (kFwd3SoftmaxWarps + 4);
constexpr int kFwd3Smem = 12 * kTileBytes + 2 * 2 * 2048 + 2 * 512 + 1024 + 1024;   // K/V 8 tiles, Q 4 tiles, max/sum tables, barriers, alignment

// kRun: 8-column chunks of S one softmax thread may hold (7: N <= 224, 56 registers; 8: N <= 256)
template <int kRun>
__global__ void __launch_bounds__(kFwd3Threads, 1) attn_fwd_pipe_kernel(const __grid_constant__ AttnTcParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TRACE_DECL;
  const int N = p.N, D = p.H * DH;
  const int ntiles = (N + TILE - 1) / TILE;      // 1 or 2
  const int kvbufs = ntiles == 2 ? 2 : 4;        // K/V buffers in the 8-tile region (one tile each of K and V when N <= 128)
  auto sK = [&](int j) { return base + uint32_t(j * 2 * ntiles) * kTileBytes; };
  auto sV = [&](int j) { return base + uint32_t(j * 2 * ntiles + ntiles) * kTileBytes; };
  auto sQ = [&](int slot) { return base + uint32_t(8 + slot) * kTileBytes; };
  const uint32_t sMax = base + 12 * kTileBytes;       // [buffer][4 column runs][128 rows] partial row maxima
  const uint32_t sSum = sMax + 2 * 2048;              // [buffer][4][128] partial row sums
  const uint32_t sMxF = sSum + 2 * 2048;              // [buffer][128] final row max (for the log-sum-exp)
  const uint32_t sBar = sMxF + 2 * 512;
  auto bar_kv = [&](int j) { return sBar + 8u * j; };            // 4: K/V of the item in buffer j landed
  auto bar_q = [&](int s) { return sBar + 32u + 8u * s; };       // 4: Q tile of unit (k & 3) landed
  auto bar_s = [&](int g) { return sBar + 64u + 8u * g; };       // 2: S of the buffer's unit complete
  auto bar_p = [&](int g) { return sBar + 80u + 8u * g; };       // 2: P written (one arrival per softmax warp)
  auto bar_o = [&](int g) { return sBar + 96u + 8u * g; };       // 2: O complete
  auto bar_drained = [&](int g) { return sBar + 112u + 8u * g; };  // 2: O read out (one arrival per epilogue warp)
  const uint32_t sTmem = sBar + 128u;

  const int npad = (N + 15) & ~15;               // MMA N extent over the kv axis
  const int rows1 = ntiles == 2 ? ((N - TILE + 15) & ~15) : 0;
  const int items = p.B * p.H;
  const int n_local = (items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int n_units = n_local * ntiles;
  constexpr int kCtlWarp = kFwd3SoftmaxWarps + 3;   // the epilogue warp of lane quarter 3: idle on the short second tile of N = 197

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmQKV1);
    tma_prefetch_desc(&p.tmO3);
    for (int j = 0; j < 4; ++j) { mbar_init(bar_kv(j), 1); mbar_init(bar_q(j), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_s(b), 1); mbar_init(bar_p(b), kFwd3SoftmaxWarps); mbar_init(bar_o(b), 1); mbar_init(bar_drained(b), 4);
    }
    fence_mbar_init();
  }
  if (warp == kCtlWarp) {
    tmem_alloc(sTmem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));

  if (warp < kFwd3SoftmaxWarps) {
    // ================================ softmax ================================
    const int q = warp & 3, hf = warp >> 2;
    const int rt = q * 32 + lane;              // row within the tile (= TMEM lane)
    const float c = p.scale * kLog2e;
    const int nch = npad / 8;                  // 8-column chunks of the kv axis (<= 32), split into four runs of <= kRun
    const int cbase = nch >> 2, crem = nch & 3;
    const int cnt = cbase + (hf < crem ? 1 : 0);
    const int c0 = hf * cbase + (hf < crem ? hf : crem);
    // The S read-out (TMEM reads run at 16 B/clk per scheduler: ~1600 cycles per unit) is a phase of its own.  Refilling the
    // registers of each finished chunk of P(k) with the same chunk of S(k+1) to hide it under the exp2 phase was measured SLOWER
    // (139 us against 107 us): LDTM and MUFU instructions leave through the same dispatch queue, and a queue full of LDTMs waiting
    // for the read port holds the MUFUs back.
    for (int k = 0; k < n_units; ++k) {
      const int n = ntiles == 2 ? k >> 1 : k, t = ntiles == 2 ? k & 1 : 0;
      const int buf = k & 1;
      const uint32_t trow = tmem + uint32_t(buf) * 256u + (uint32_t(q * 32) << 16);
      const bool live = t * TILE + q * 32 < N;   // warp-uniform: any valid query row in this warp
      int Lrow = N, Lwarp = N;                   // this row's key limit / the smallest limit in this warp
      if (p.kv_len != nullptr || p.causal) {
        const int item = int(blockIdx.x) + n * int(gridDim.x);
        int Lk = p.kv_len ? __ldg(p.kv_len + item / p.H) : N;   // valid keys of this sequence (key-padding mask = suffix of the row)
        Lk = Lk < 1 ? 1 : (Lk > N ? N : Lk);
        Lrow = p.causal ? min(Lk, min(t * TILE + rt, N - 1) + 1) : Lk;
        Lwarp = p.causal ? min(Lk, min(t * TILE + q * 32, N - 1) + 1) : Lk;
      }
      // dead warps (all rows past the sequence end) skip the math but keep in step with the barriers' phases
      if (warp == 0) TRACE(40, k);
      mbar_wait(bar_s(buf), uint32_t(k >> 1) & 1u);
      tc_fence_after();
      if (warp == 0) TRACE(42, k);
      if (live) {
        uint32_t v[kRun][8];                     // this thread's run of S, read once
        const uint32_t sbase = trow + uint32_t(c0) * 8u, pbase = trow + uint32_t(c0) * 4u;
#pragma unroll
        for (int kk = 0; kk < kRun; ++kk)
          if (kk < cnt) tmem_ld8(sbase + kk * 8, v[kk]);
        tmem_ld_wait();
        // ---- row max over this thread's columns; masked keys become -inf (and exp2 of them 0)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int kk = 0; kk < kRun; ++kk) {
          if (kk < cnt) {
            const int col0 = (c0 + kk) * 8;
            if (col0 + 8 > Lwarp) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[kk][i] = (col0 + i < Lrow) ? v[kk][i] : 0xff800000u;
            }
#pragma unroll
            for (int i = 0; i < 8; i += 4)
              m4[(2 * kk + (i >> 2)) & 3] = fmaxf(fmaxf(m4[(2 * kk + (i >> 2)) & 3], fmaxf(__uint_as_float(v[kk][i]), __uint_as_float(v[kk][i + 1]))),
                                                  fmaxf(__uint_as_float(v[kk][i + 2]), __uint_as_float(v[kk][i + 3])));
          }
        }
        float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const uint32_t mrow = sMax + uint32_t(buf) * 2048u + 4u * uint32_t(rt);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(mrow + 512u * uint32_t(hf)), "f"(mx) : "memory");
        if (warp == 0) TRACE(44, k);
        named_bar_sync(1 + q, 128);
        if (warp == 0) TRACE(46, k);
        {
          float o0, o1, o2, o3;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o0) : "r"(mrow));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o1) : "r"(mrow + 512u));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o2) : "r"(mrow + 1024u));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o3) : "r"(mrow + 1536u));
          mx = fmaxf(fmaxf(o0, o1), fmaxf(o2, o3));
        }
        const float mc = mx * c;
        // ---- P = exp2(S*c - max*c) as bf16 pairs over this thread's own S registers; every thread of the row has finished
        //      reading S (it sits in registers since before the barrier), so P may overwrite it in TMEM right away
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int kk = 0; kk < kRun; ++kk) {
          if (kk < cnt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[kk][2 * i]), c, -mc));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[kk][2 * i + 1]), c, -mc));
              s0 += p0;
              s1 += p1;
              v[kk][i] = pack_bf16x2(p0, p1);
            }
            tmem_st4(pbase + kk * 4, v[kk]);
          }
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sSum + uint32_t(buf) * 2048u + 512u * uint32_t(hf) + 4u * uint32_t(rt)), "f"(s0 + s1) : "memory");
        if (hf == 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sMxF + uint32_t(buf) * 512u + 4u * uint32_t(rt)), "f"(mx) : "memory");
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p(buf));
      if (warp == 0) TRACE(48, k);
      if (warp == 15) TRACE(49, k);
      if (warp == 12) TRACE(47, k);
    }
  } else {
    // ================================ epilogue + control ================================
    const int q = warp & 3;
    const int rt = q * 32 + lane;
    const bool ctl = warp == kCtlWarp;
    // The control warp walks its role code as a whole (warp-uniform values stay in uniform registers, where the MMA / TMA
    // instructions take their operands); the elected lane (the same one every time) issues, so program order is issue order.
    const uint32_t idesc_s = make_idesc_bf16(TILE, npad);
    constexpr uint32_t idesc_o = make_idesc_bf16(TILE, DH, 0, 1);
    const int nsl = npad / 16;
    auto item_of = [&](int n) { return int(blockIdx.x) + n * int(gridDim.x); };
    auto load_kv = [&](int n) {        // issuer only
      const int item = item_of(n), b = item / p.H, h = item - b * p.H, row0 = b * N, j = n % kvbufs;
      mbar_arrive_expect_tx(bar_kv(j), 2 * kTileBytes + 2 * rows1 * 128);
      tma_load_2d(sK(j), &p.tmQKV, bar_kv(j), D + h * DH, row0);
      tma_load_2d(sV(j), &p.tmQKV, bar_kv(j), 2 * D + h * DH, row0);
      if (ntiles == 2) {
        tma_load_2d(sK(j) + kTileBytes, &p.tmQKV1, bar_kv(j), D + h * DH, row0 + TILE);
        tma_load_2d(sV(j) + kTileBytes, &p.tmQKV1, bar_kv(j), 2 * D + h * DH, row0 + TILE);
      }
    };
    auto load_q = [&](int k) {         // issuer only
      const int n = ntiles == 2 ? k >> 1 : k, t = ntiles == 2 ? k & 1 : 0;
      const int item = item_of(n), b = item / p.H, h = item - b * p.H, qs = k & 3;
      mbar_arrive_expect_tx(bar_q(qs), t ? rows1 * 128 : kTileBytes);
      tma_load_2d(sQ(qs), t ? &p.tmQKV1 : &p.tmQKV, bar_q(qs), h * DH, b * N + t * TILE);
    };
    auto issue_s = [&](int k) {        // control warp (all lanes wait, the issuer issues)
      const int n = ntiles == 2 ? k >> 1 : k;
      const int j = n % kvbufs, qs = k & 3;
      mbar_wait(bar_kv(j), uint32_t(n / kvbufs) & 1u);
      mbar_wait(bar_q(qs), uint32_t(k >> 2) & 1u);
      if (k >= 2) mbar_wait(bar_drained(k & 1), uint32_t((k >> 1) - 1) & 1u);   // O(k-2) has left this buffer
      TRACE(59, k);
      tc_fence_after();
      const uint64_t dq = desc_kmajor(sQ(qs)), dk = desc_kmajor(sK(j));
      const uint32_t slot = tmem + uint32_t(k & 1) * 256u;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) umma_ss(slot, dq + uint64_t(kk * 2), dk + uint64_t(kk * 2), idesc_s, kk != 0);
        umma_commit(bar_s(k & 1));
      }
      __syncwarp();
      TRACE(60, k);
    };
    if (ctl) {
      if (elect_one()) {
        for (int n = 0; n < kvbufs && n < n_local; ++n) load_kv(n);
        for (int k = 0; k < 3 && k < n_units; ++k) load_q(k);
      }
      __syncwarp();
      if (n_units > 0) issue_s(0);
      if (n_units > 1) issue_s(1);
    }
    for (int k = 0; k < n_units; ++k) {
      const int n = ntiles == 2 ? k >> 1 : k, t = ntiles == 2 ? k & 1 : 0;
      const int item = item_of(n);
      const int b = item / p.H, h = item - b * p.H;
      const int buf = k & 1, qs = k & 3;
      const int r = t * TILE + rt;
      const bool live = t * TILE + q * 32 < N;
      const uint32_t slot = tmem + uint32_t(buf) * 256u;
      const uint32_t orow = slot + (uint32_t(q * 32) << 16) + 128u;
      if (ctl) {
        // ---- O(k) = P(k) V.  This wait spans most of a unit: back off between probes so the spinning warp does not take issue
        //      slots from the four softmax warps on its scheduler
        while (!mbar_try_wait(bar_p(buf), uint32_t(k >> 1) & 1u)) __nanosleep(32);
        TRACE(61, k);
        tc_fence_after();
        const uint64_t dv = desc_mnmajor(sV(n % kvbufs), 0);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 16; ++j) umma_ts_if(j < nsl, slot + 128, slot + j * 8, dv + uint64_t(j * 128), idesc_o, j != 0);
          umma_commit(bar_o(buf));
          // while P V runs: Q of unit k + 3 goes into the slot O(k - 1) was staged in; that store (issued a unit ago) must have
          // read it
          if (k + 3 < n_units) {
            tma_store_wait_read<0>();
            load_q(k + 3);
          }
        }
        __syncwarp();
        TRACE(62, k);
      }
      if (q == 0) TRACE(50, k);
      mbar_wait(bar_o(buf), uint32_t(k >> 1) & 1u);
      if (q == 0) TRACE(51, k);
      tc_fence_after();
      float sum = 1.f, mx = 0.f;
      uint32_t oa[32], pa[16];
      if (live) {
        tmem_ld32(orow, oa);
        const uint32_t srow = sSum + uint32_t(buf) * 2048u + 4u * uint32_t(rt);
        float a0, a1, a2, a3;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a0) : "r"(srow));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a1) : "r"(srow + 512u));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a2) : "r"(srow + 1024u));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a3) : "r"(srow + 1536u));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mx) : "r"(sMxF + uint32_t(buf) * 512u + 4u * uint32_t(rt)));
        sum = (a0 + a1) + (a2 + a3);
        tmem_ld_wait();
        const float inv0 = rcp_approx(sum);
#pragma unroll
        for (int j = 0; j < 16; ++j) pa[j] = pack_bf16x2(__uint_as_float(oa[2 * j]) * inv0, __uint_as_float(oa[2 * j + 1]) * inv0);
        tmem_ld32(orow + 32, oa);
        tmem_ld_wait();
      }
      // O and the row statistics of this buffer are in registers: the buffer may take S of unit k + 2
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_drained(buf));
      if (q == 0) TRACE(52, k);
      if (live) {
        const float inv = rcp_approx(sum);
        const uint32_t rb = sQ(qs) + uint32_t(rt) * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t a = rb + ((uint32_t(j) ^ uint32_t(rt & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pa[4 * j]), "r"(pa[4 * j + 1]), "r"(pa[4 * j + 2]), "r"(pa[4 * j + 3]) : "memory");
        }
#pragma unroll
        for (int j = 4; j < 8; ++j) {
          const uint32_t* ov = oa + 8 * (j - 4);
          const uint32_t a = rb + ((uint32_t(j) ^ uint32_t(rt & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                       "r"(pack_bf16x2(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv))
                       : "memory");
        }
        if (r < N && p.lse) p.lse[(size_t(b) * p.H + h) * N + r] = fmaf(mx, p.scale, 0.6931471805599453f * lg2_approx(sum));
      }
      fence_proxy_async_smem();
      named_bar_sync(10, 128);
      if (q == 0) TRACE(53, k);
      if (ctl) {
        if (elect_one()) {
          tma_store_3d(&p.tmO3, sQ(qs), h * DH, t * TILE, b);
          tma_store_commit();
          TRACE(54, k);
          // O(k) is complete, so every MMA that read this item's K/V is: its buffer takes item n + kvbufs
          if (t == ntiles - 1 && n + kvbufs < n_local) load_kv(n + kvbufs);
        }
        __syncwarp();
        TRACE(55, k);
        if (k + 2 < n_units) issue_s(k + 2);
      }
    }
    if (ctl) {
      if (elect_one()) tma_store_wait_read<0>();
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kCtlWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =====================================================================================================
// backward
// =====================================================================================================
// Persistent, software-pipelined backward.  One CTA per SM walks (batch, head) items.  Roles:
//   warps 0-15  compute  (kv row = TMEM lane = 32*(warp&3)+lane; warp>>2 picks 16 of the step's 64 query columns)
//   warp  16    MMA issuer (whole warp walks the loop, one elected lane issues)
//   warp  17    TMA loader (one lane): fetches tile groups of the NEXT item as soon as the MMAs that read them finish
//   warps 18-19 statistics: lse*log2e and scale*rowsum(dO * O) of the next item into a double-buffered smem table
// An item is cut into "steps" of (kv tile j, query tile i, 64-column half).  S^T / dP^T of step s+1 are issued into
// the other TMEM buffer before the issuer waits for P^T of step s, the dV/dK/dQ accumulators of a kv tile are drained
// by the compute warps AFTER they have produced the next step's P^T, and the query-tile order alternates between
// items (j=0: i=0,1  j=1: i=1,0 | next item j=0: i=1,0  j=1: i=0,1) so that every tile group of the next item is free
// at least one pair before it is needed.  Query extents are trimmed to the padded live length (N = 197: the second
// query tile costs 64 + 16 columns instead of 128) and the kv extent of dQ to ceil16(live kv rows).
constexpr int kBwdComputeWarps = 16;
constexpr int kBwdThreads = 32 * (kBwdComputeWarps + 4);
// The issuer paces the kernel; it sits on scheduler 3, whose compute warps (kv rows 96..127 of a tile) idle on the short second kv
// tile of N = 197.  Loader on scheduler 2, the two statistics warps on schedulers 0 and 1.
constexpr int kBwdIssueWarp = kBwdComputeWarps + 3, kBwdLoadWarp = kBwdComputeWarps + 2, kBwdStatWarp0 = kBwdComputeWarps;
constexpr int kBwdSmem = 13 * kTileBytes + 2 * 2048 + 1024 + 1024;   // + one tile to transpose the accumulator read-out

struct BwdStep {
  int j, i, ii, half, nq;  // ii: position of the pair within its kv tile; nq: live query columns padded to 16 (0 = dead)
};
// ntiles is 1 or 2: no divisions on the issuing thread's critical path
NGU_DEVINL BwdStep bwd_step(int r, int par, int ntiles, int N) {
  BwdStep s;
  s.half = r & 1;
  s.ii = ntiles == 2 ? (r >> 1) & 1 : 0;
  s.j = ntiles == 2 ? r >> 2 : 0;
  s.i = ntiles == 2 ? (s.ii ^ s.j ^ par) : 0;
  int live = N - (s.i * TILE + s.half * 64);
  live = live < 0 ? 0 : (live > 64 ? 64 : live);
  s.nq = (live + 15) & ~15;
  return s;
}

// ---- MMA issuer of the backward kernel.  This single warp paces the whole CTA, so its per-step instruction count is
// what matters: the step sequence of an item is unrolled at compile time (kv tile, query tile, half and all smem
// descriptor offsets are constants; only the live query width of the last tile is a run-time value).
template <int NT>
struct BwdIssuer {
  static constexpr int kRaw = 2 * NT * NT;
  __host__ __device__ static constexpr int sj(int r) { return NT == 2 ? r >> 2 : 0; }
  __host__ __device__ static constexpr int sii(int r) { return NT == 2 ? (r >> 1) & 1 : 0; }
  __host__ __device__ static constexpr int si(int r, int par) { return NT == 2 ? (sii(r) ^ sj(r) ^ par) : 0; }
  __host__ __device__ static constexpr int sh(int r) { return r & 1; }

  uint32_t tmem, sBar;
  int n_local;
  int nq[2][2];     // [query tile][half]: live columns padded to 16 (0 = dead step)
  int nkvs[2];      // [kv tile]: ceil16(live kv rows) / 16
  uint64_t dQ0, dK0, dV0, dDO0, mQ0, mK0, mDO0, mDS0;
  uint32_t sc_s = 0, sc_m = 0, pc = 0, jc = 0;
  TRACE_DECL;

  NGU_DEVINL uint32_t bar_full(int g) const { return sBar + 8u * g; }
  NGU_DEVINL uint32_t bar_free(int g) const { return sBar + 32u + 8u * g; }
  NGU_DEVINL uint32_t bar_s(uint32_t b) const { return sBar + 64u + 8u * b; }
  NGU_DEVINL uint32_t bar_p(uint32_t b) const { return sBar + 80u + 8u * b; }
  NGU_DEVINL uint32_t bar_dq(uint32_t b) const { return sBar + 96u + 8u * b; }
  NGU_DEVINL uint32_t bar_acc() const { return sBar + 144u; }
  NGU_DEVINL uint32_t bar_drained() const { return sBar + 152u; }

  // S^T = K_j Q_{i,half}^T and dP^T = V_j dO_{i,half}^T into the next TMEM buffer
  template <int R, int PAR>
  NGU_DEVINL void issue_s(int n) {
    constexpr int j = sj(R), i = si(R, PAR), half = sh(R);
    const uint32_t ph = uint32_t(n) & 1u;
    if constexpr (R == 0) { mbar_wait(bar_full(0), ph); mbar_wait(bar_full(2 + i), ph); }   // first use of K0/V0, Q_i/dO_i
    if constexpr (NT == 2 && R == 2) mbar_wait(bar_full(2 + i), ph);                        // first use of the other Q/dO
    if constexpr (NT == 2 && R == 4) mbar_wait(bar_full(1), ph);                            // first use of K1/V1
    TRACE(10, sc_s);
    tc_fence_after();
    const uint32_t cb = tmem + (sc_s & 1u) * 128u;
    const uint32_t idesc_st = make_idesc_bf16(TILE, nq[i][half]);
    constexpr uint64_t qo = uint64_t((i * kTileBytes + half * 8192) >> 4), ko = uint64_t((j * kTileBytes) >> 4);
    if (elect_one()) {
      // the two accumulation chains are interleaved: consecutive MMAs into the same accumulator are latency-bound
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) {
        umma_ss(cb, dK0 + (ko + 2 * k), dQ0 + (qo + 2 * k), idesc_st, k != 0);
        umma_ss(cb + 64, dV0 + (ko + 2 * k), dDO0 + (qo + 2 * k), idesc_st, k != 0);
      }
      umma_commit(bar_s(sc_s & 1u));
    }
    __syncwarp();
    TRACE(11, sc_s);
    ++sc_s;
  }

  // S^T / dP^T of the first live step at or after R.  Crossing into the next item happens BEFORE this step's P^T is
  // awaited when NT == 2 (its tiles were released at least one pair ago) and AFTER this step's MMAs when NT == 1
  // (its tiles are released by exactly those MMAs).
  template <int R, int PAR>
  NGU_DEVINL void lookahead(int n, bool before) {
    if constexpr (R < kRaw) {
      if (nq[si(R, PAR)][sh(R)] > 0) {
        if (before) issue_s<R, PAR>(n);
      } else {
        lookahead<R + 1, PAR>(n, before);
      }
    } else {
      if (((NT == 2) == before) && n + 1 < n_local) issue_s<0, PAR ^ 1>(n + 1);
    }
  }

  // dV_j += P^T dO, dK_j += dS^T Q (A from TMEM), and after the pair's last half dQ_i += dS K_j (A = dS^T in smem)
  template <int R, int PAR>
  NGU_DEVINL void mma(int n) {
    constexpr int j = sj(R), ii = sii(R), i = si(R, PAR), half = sh(R);
    constexpr bool first_of_j = ii == 0 && half == 0;
    constexpr uint32_t idesc_ts = make_idesc_bf16(TILE, DH, 0, 1);
    constexpr uint32_t idesc_dq = make_idesc_bf16(TILE, DH, 1, 1);
    constexpr uint64_t qo = uint64_t((i * kTileBytes + half * 8192) >> 4), ko = uint64_t((j * kTileBytes) >> 4);
    const uint32_t cb = tmem + (sc_m & 1u) * 128u;
    const int nks = nq[i][half] >> 4;
    const bool last_half = half == 1 || nq[i][1] == 0;
    TRACE(20, sc_m);
    mbar_wait(bar_p(sc_m & 1u), (sc_m >> 1) & 1u);
    TRACE(21, sc_m);
    // the accumulators of the previous kv tile (and dQ of the previous item) must have been read out
    if constexpr (first_of_j) { if (jc > 0) mbar_wait(bar_drained(), (jc - 1u) & 1u); }
    tc_fence_after();
    if (elect_one()) {
      const uint64_t dsd = mDS0 + uint64_t(((pc & 1u) * 2 * kTileBytes) >> 4);
      const int nk = last_half ? nkvs[j] : 0;
      // three accumulation chains (dV, dK, dQ) interleaved: consecutive MMAs into one accumulator are latency-bound
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // P^T / dS^T (bf16) of query slice k sit in columns [16k, 16k + 8) of the S^T / dP^T block
        umma_ts_if(k < nks, tmem + 256, cb + 16 * k, mDO0 + (qo + 128 * k), idesc_ts, (first_of_j && k == 0) ? 0u : 1u);
        umma_ts_if(k < nks, tmem + 320, cb + 64 + 16 * k, mQ0 + (qo + 128 * k), idesc_ts, (first_of_j && k == 0) ? 0u : 1u);
        umma_ss_if(2 * k < nk, tmem + 384 + i * 64, dsd + 128 * (2 * k), mK0 + (ko + 128 * (2 * k)), idesc_dq, (j | k) != 0 ? 1u : 0u);
        umma_ss_if(2 * k + 1 < nk, tmem + 384 + i * 64, dsd + 128 * (2 * k + 1), mK0 + (ko + 128 * (2 * k + 1)), idesc_dq, 1u);
      }
      if (last_half) {
        umma_commit(bar_dq(pc & 1u));
        if constexpr (j == NT - 1) umma_commit(bar_free(2 + i));   // last kv tile: Q_i / dO_i are done
        if constexpr (ii == NT - 1) {                              // last pair of this kv tile
          umma_commit(bar_free(j));
          umma_commit(bar_acc());
        }
      }
    }
    __syncwarp();
    TRACE(22, sc_m);
    ++sc_m;
    if (last_half) {
      ++pc;
      if constexpr (ii == NT - 1) ++jc;
    }
  }

  template <int R, int PAR>
  NGU_DEVINL void steps(int n) {
    if constexpr (R < kRaw) {
      if (nq[si(R, PAR)][sh(R)] > 0) {
        lookahead<R + 1, PAR>(n, true);
        mma<R, PAR>(n);
        lookahead<R + 1, PAR>(n, false);
      }
      steps<R + 1, PAR>(n);
    }
  }

  NGU_DEVINL void run() {
    if (n_local > 0) issue_s<0, 0>(0);
    for (int n = 0; n < n_local; n += 2) {
      steps<0, 0>(n);
      if (n + 1 < n_local) steps<0, 1>(n + 1);
    }
  }
};


template <int NT>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + 2 * kTileBytes, sV = base + 4 * kTileBytes, sDO = base + 6 * kTileBytes;
  const uint32_t sDS = base + 8 * kTileBytes;   // 2 buffers x [2 q-chunks of 64][128 kv rows][128 B]
  const uint32_t sOut = base + 12 * kTileBytes;   // [128 rows][128 B] staging tile of the coalesced accumulator stores
  const uint32_t sStat = base + 13 * kTileBytes;  // [2 buffers][lse2[256], delta[256]] fp32
  const uint32_t sBar = sStat + 2 * 2048;
  // tile groups: 0 = {K0,V0}  1 = {K1,V1}  2 = {Q0,dO0}  3 = {Q1,dO1}
  auto bar_full = [&](int g) { return sBar + 8u * g; };           // 4: tile group landed
  auto bar_free = [&](int g) { return sBar + 32u + 8u * g; };     // 4: tile group no longer read by any MMA
  auto bar_s = [&](int b) { return sBar + 64u + 8u * b; };        // 2: S^T / dP^T of the buffer complete
  auto bar_p = [&](int b) { return sBar + 80u + 8u * b; };        // 2: P^T / dS^T of the buffer written (512 arrivals)
  auto bar_dq = [&](int b) { return sBar + 96u + 8u * b; };       // 2: dQ MMAs that read dS smem buffer b complete
  auto bar_stat = [&](int b) { return sBar + 112u + 8u * b; };    // 2: lse / delta table b filled (64 arrivals)
  auto bar_statfree = [&](int b) { return sBar + 128u + 8u * b; };  // 2: compute warps done reading table b (512 arrivals)
  const uint32_t bar_acc = sBar + 144u;                           // dV_j / dK_j (and dQ after the last j) complete
  const uint32_t bar_drained = sBar + 152u;                       // accumulators read out (512 arrivals)
  const uint32_t sTmem = sBar + 160u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TRACE_DECL;
  const int N = p.N, D = p.H * DH;
  constexpr int ntiles = NT;
  constexpr int nraw = ntiles * ntiles * 2;
  const int items = p.B * p.H;
  const int n_local = (items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int rows1 = ntiles == 2 ? ((N - TILE + 15) & ~15) : 0;   // rows of the second tile that are fetched
  // TMEM columns: buffer b holds S^T at b*128 (64 fp32 cols) and dP^T at b*128 + 64
  constexpr uint32_t cDV = 256, cDK = 320, cDQ = 384;
  constexpr int kCompute = 32 * kBwdComputeWarps;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmDO);
    for (int g = 0; g < 4; ++g) { mbar_init(bar_full(g), 1); mbar_init(bar_free(g), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_s(b), 1);
      mbar_init(bar_p(b), kCompute);
      mbar_init(bar_dq(b), 1);
      mbar_init(bar_stat(b), 64);
      mbar_init(bar_statfree(b), kCompute);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_drained, kCompute);
    fence_mbar_init();
  }
  if (warp == kBwdIssueWarp) {
    tmem_alloc(sTmem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));

  if (warp == kBwdLoadWarp) {
    // ================================ TMA loader ================================
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int item = int(blockIdx.x) + n * int(gridDim.x);
        const int b = item / p.H, h = item - b * p.H;
        const int row0 = b * N;
        const uint32_t fph = uint32_t(n - 1) & 1u;
        auto load_group = [&](int g) {
          if (n > 0) mbar_wait(bar_free(g), fph);
          const int t = g & 1;                       // tile index within the sequence
          const uint32_t off = t * kTileBytes;
          mbar_arrive_expect_tx(bar_full(g), t ? 2 * rows1 * 128 : 2 * kTileBytes);
          const CUtensorMap* mq = t ? &p.tmQKV1 : &p.tmQKV;
          if (g < 2) {
            tma_load_2d(sK + off, mq, bar_full(g), D + h * DH, row0 + t * TILE);
            tma_load_2d(sV + off, mq, bar_full(g), 2 * D + h * DH, row0 + t * TILE);
          } else {
            tma_load_2d(sQ + off, mq, bar_full(g), h * DH, row0 + t * TILE);
            tma_load_2d(sDO + off, t ? &p.tmDO1 : &p.tmDO, bar_full(g), h * DH, row0 + t * TILE);
          }
        };
        // in the order the item needs them (which is also the order the previous item releases them)
        load_group(0);
        if (ntiles == 2) {
          const int first_i = n & 1;
          load_group(2 + first_i);
          load_group(2 + (first_i ^ 1));
          load_group(1);
        } else {
          load_group(2);
        }
      }
    }
  } else if (warp == kBwdStatWarp0 || warp == kBwdStatWarp0 + 1) {
    // ================================ statistics ================================
    const int l64 = (warp - kBwdStatWarp0) * 32 + lane;
    for (int n = 0; n < n_local; ++n) {
      const int item = int(blockIdx.x) + n * int(gridDim.x);
      const int b = item / p.H, h = item - b * p.H;
      const int row0 = b * N;
      // table n&1 was last read by item n-2
      if (n >= 2) mbar_wait(bar_statfree(n & 1), uint32_t((n >> 1) - 1) & 1u);
      const uint32_t tab = sStat + (n & 1) * 2048;
      for (int r = l64; r < ntiles * TILE; r += 64) {
        // rows past the sequence end: lse = +inf makes every probability of that query column exactly 0
        float l2 = INFINITY, dl = 0.f;
        if (r < N) {
          l2 = p.lse[(size_t(b) * p.H + h) * N + r] * kLog2e;
          const uint4* po = reinterpret_cast<const uint4*>(p.o_in + size_t(row0 + r) * D + h * DH);
          const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + size_t(row0 + r) * D + h * DH);
          uint4 a[8], g[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { a[j] = __ldg(po + j); g[j] = __ldg(pd + j); }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t aw[4] = {a[j].x, a[j].y, a[j].z, a[j].w}, gw[4] = {g[j].x, g[j].y, g[j].z, g[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(gw[e]);
              dl = fmaf(x.x, y.x, dl);
              dl = fmaf(x.y, y.y, dl);
            }
          }
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(tab + 4u * r), "f"(l2) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(tab + 1024u + 4u * r), "f"(dl * p.scale) : "memory");
      }
      mbar_arrive(bar_stat(n & 1));
    }
  } else if (warp == kBwdIssueWarp) {
    // ================================ MMA issuer ================================
    BwdIssuer<NT> is;
    is.tmem = tmem;
    is.sBar = sBar;
    is.n_local = n_local;
    for (int i = 0; i < 2; ++i)
      for (int hh = 0; hh < 2; ++hh) {
        int live = N - (i * TILE + hh * 64);
        live = live < 0 ? 0 : (live > 64 ? 64 : live);
        is.nq[i][hh] = (live + 15) & ~15;
      }
    for (int j = 0; j < 2; ++j) {
      int nkv = N - j * TILE;
      nkv = nkv < 0 ? 0 : (nkv > TILE ? TILE : nkv);
      is.nkvs[j] = (nkv + 15) >> 4;
    }
    is.dQ0 = desc_kmajor(sQ); is.dK0 = desc_kmajor(sK); is.dV0 = desc_kmajor(sV); is.dDO0 = desc_kmajor(sDO);
    is.mQ0 = desc_mnmajor(sQ, 0); is.mK0 = desc_mnmajor(sK, 0); is.mDO0 = desc_mnmajor(sDO, 0);
    is.mDS0 = desc_mnmajor(sDS, kTileBytes);
    is.run();
  } else {
    // ================================ compute ================================
    const int qd = warp & 3, hq = warp >> 2;     // TMEM lane quarter, 16-column slice of the step's 64 query columns
    const int t = qd * 32 + lane;                // kv row within the tile (= TMEM lane)
    const uint32_t trow = tmem + (uint32_t(qd * 32) << 16);
    const float c = p.scale * kLog2e;
    uint32_t sc = 0, pc = 0, accn = 0;
    // deferred read-out of dV_j / dK_j (/ dQ): done after the NEXT step's P^T so the issuer never waits for it
    bool pend = false;
    int pend_row0 = 0, pend_h = 0, pend_j = 0;
    // Accumulator read-out.  Each thread holds 16 columns of one row; storing them directly costs 32 LSU wavefronts per
    // instruction (one 16/32-byte piece of 32 different rows).  The pieces are transposed through a 128B-swizzled smem
    // tile instead and written as whole 128-byte rows (4 rows per instruction).
    auto pack16 = [&](const uint32_t (&v)[16], uint4& u0, uint4& u1) {
      u0.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));
      u0.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
      u0.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));
      u0.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
      u1.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9]));
      u1.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
      u1.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13]));
      u1.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
    };
    const int ctid = threadIdx.x;                 // compute threads are 0 .. kCompute-1
    // tile [128 x 64] (this thread's 2 pieces at row t) -> global rows `grow0 + r` (r < nrows), columns gcol .. gcol+63
    auto flush_tile = [&](const uint4& u0, const uint4& u1, int grow0, int nrows, int gcol) {
      asm volatile("bar.sync 10, %0;" ::"r"(kCompute) : "memory");          // previous tile fully copied out
      const uint32_t rb = sOut + uint32_t(t) * 128u;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rb + ((uint32_t(hq * 2) ^ uint32_t(t & 7)) << 4)),
                   "r"(u0.x), "r"(u0.y), "r"(u0.z), "r"(u0.w) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rb + ((uint32_t(hq * 2 + 1) ^ uint32_t(t & 7)) << 4)),
                   "r"(u1.x), "r"(u1.y), "r"(u1.z), "r"(u1.w) : "memory");
      asm volatile("bar.sync 10, %0;" ::"r"(kCompute) : "memory");
#pragma unroll
      for (int k = 0; k < (TILE * 8) / kCompute; ++k) {
        const int idx = k * kCompute + ctid;
        const int r = idx >> 3, pc8 = idx & 7;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(sOut + uint32_t(r) * 128u + ((uint32_t(pc8) ^ uint32_t(r & 7)) << 4)));
        if (r < nrows) *reinterpret_cast<uint4*>(p.dqkv + size_t(grow0 + r) * 3 * D + gcol + pc8 * 8) = v;
      }
    };
    auto drain = [&]() {
      mbar_wait(bar_acc, accn & 1u);
      ++accn;
      tc_fence_after();
      const bool with_dq = pend_j == ntiles - 1;
      uint4 v0, v1, k0, k1, q0, q1, q2, q3;
      {
        uint32_t a[16], bq[16];
        tmem_ld16(trow + cDV + hq * 16, a);
        tmem_ld16(trow + cDK + hq * 16, bq);
        tmem_ld_wait();
        pack16(a, v0, v1);
        pack16(bq, k0, k1);
        if (with_dq) {
          tmem_ld16(trow + cDQ + hq * 16, a);
          if (ntiles == 2) tmem_ld16(trow + cDQ + 64 + hq * 16, bq);
          tmem_ld_wait();
          pack16(a, q0, q1);
          if (ntiles == 2) pack16(bq, q2, q3);
        }
      }
      // TMEM is free again: let the issuer go on before the (slower) global stores
      tc_fence_before();
      mbar_arrive(bar_drained);
      int nkv = N - pend_j * TILE;
      nkv = nkv > TILE ? TILE : nkv;
      const int grow = pend_row0 + pend_j * TILE;
      flush_tile(v0, v1, grow, nkv, 2 * D + pend_h * DH);
      flush_tile(k0, k1, grow, nkv, D + pend_h * DH);
      if (with_dq) {
        flush_tile(q0, q1, pend_row0, N > TILE ? TILE : N, pend_h * DH);
        if (ntiles == 2) flush_tile(q2, q3, pend_row0 + TILE, N - TILE, pend_h * DH);
      }
      pend = false;
    };
    for (int n = 0; n < n_local; ++n) {
      const int item = int(blockIdx.x) + n * int(gridDim.x);
      const int b = item / p.H, h = item - b * p.H;
      const int row0 = b * N;
      const uint32_t tab = sStat + (n & 1) * 2048;
      int Lk = p.kv_len ? __ldg(p.kv_len + b) : N;       // key-padding mask: kv rows >= Lk get P = dS = 0 (dK = dV = 0 there)
      Lk = Lk < 1 ? 1 : (Lk > N ? N : Lk);
      mbar_wait(bar_stat(n & 1), (n >> 1) & 1);
      for (int r = 0; r < nraw; ++r) {
        const BwdStep s = bwd_step(r, n & 1, ntiles, N);
        if (s.nq == 0) continue;
        const int kvb = s.j * TILE + qd * 32;              // first kv row of this warp
        const bool warp_live = kvb < N;                    // any valid kv row in this warp (warp-uniform)
        const bool partial = kvb + 32 > Lk;                // some rows of the warp are past the (valid) end
        const uint32_t cb = trow + (sc & 1u) * 128u;
        if (warp == 0) TRACE(30, sc);
        mbar_wait(bar_s(sc & 1u), (sc >> 1) & 1u);
        if (warp == 0) TRACE(31, sc);
        tc_fence_after();
        // the dS smem buffer of this pair was last read by the dQ MMAs of pair pc-2
        if (s.half == 0 && pc >= 2) mbar_wait(bar_dq(pc & 1u), ((pc >> 1) - 1u) & 1u);
        if (warp_live && hq * 16 < s.nq) {
          uint32_t sv[16], dv[16], pp[8], ds[8];
          tmem_ld16(cb + hq * 16, sv);
          tmem_ld16(cb + 64 + hq * 16, dv);
          const uint32_t qa = tab + 4u * uint32_t(s.i * TILE + s.half * 64 + hq * 16);
          float l2[16], dl[16];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l2[4 * e]), "=f"(l2[4 * e + 1]), "=f"(l2[4 * e + 2]), "=f"(l2[4 * e + 3]) : "r"(qa + 16u * e));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dl[4 * e]), "=f"(dl[4 * e + 1]), "=f"(dl[4 * e + 2]), "=f"(dl[4 * e + 3]) : "r"(qa + 1024u + 16u * e));
          }
          tmem_ld_wait();
          const bool row_ok = !partial || (kvb + lane < Lk);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            // query columns past the end have lse = +inf -> p = 0 exactly; kv rows past the end are zeroed below
            float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), c, -l2[2 * e]));
            float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), c, -l2[2 * e + 1]));
            float d0 = p0 * fmaf(__uint_as_float(dv[2 * e]), p.scale, -dl[2 * e]);
            float d1 = p1 * fmaf(__uint_as_float(dv[2 * e + 1]), p.scale, -dl[2 * e + 1]);
            pp[e] = pack_bf16x2(p0, p1);
            ds[e] = pack_bf16x2(d0, d1);
          }
          if (partial && !row_ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { pp[e] = 0u; ds[e] = 0u; }
          }
          // P^T / dS^T (bf16) go back into the first 8 of the 16 columns this warp alone has just read
          tmem_st8(cb + hq * 16, pp);
          tmem_st8(cb + 64 + hq * 16, ds);
          // dS^T row of this kv index into the MN-major smem operand of dQ: 16 q values = 2 x 16 B
          const uint32_t rbase = sDS + (pc & 1u) * 2 * kTileBytes + s.half * kTileBytes + (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
          for (int piece = 0; piece < 2; ++piece) {
            const uint32_t idx = uint32_t(hq * 2 + piece);
            const uint32_t a = rbase + ((idx ^ uint32_t(t & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(ds[4 * piece]), "r"(ds[4 * piece + 1]),
                         "r"(ds[4 * piece + 2]), "r"(ds[4 * piece + 3])
                         : "memory");
          }
          tmem_st_wait();
          fence_proxy_async_smem();
        }
        tc_fence_before();
        mbar_arrive(bar_p(sc & 1u));
        if (warp == 0) TRACE(32, sc);
        ++sc;
        if (pend) { drain(); if (warp == 0) TRACE(33, sc); }
        const bool last_half = s.half == 1 || N <= s.i * TILE + 64;
        if (last_half) {
          ++pc;
          if (s.ii == ntiles - 1) { pend = true; pend_row0 = row0; pend_h = h; pend_j = s.j; }
        }
      }
      mbar_arrive(bar_statfree(n & 1));
    }
    if (pend) drain();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBwdIssueWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// packed-layout check: q/k/v views of one [B*N, 3D] buffer, o [B*N, D]
bool is_packed(const ngu_attn_desc& d) {
  const int64_t D = int64_t(d.H) * d.dh;
  const char* q = reinterpret_cast<const char*>(d.q);
  return d.N == d.S && d.q_ts == 3 * D && d.k_ts == 3 * D && d.v_ts == 3 * D && d.o_ts == D &&
         d.q_bs == int64_t(d.N) * 3 * D && d.k_bs == d.q_bs && d.v_bs == d.q_bs && d.o_bs == int64_t(d.N) * D &&
         reinterpret_cast<const char*>(d.k) == q + D * 2 && reinterpret_cast<const char*>(d.v) == q + 4 * D;
}

int fill_params(const ngu_attn_desc& d, AttnTcParams& p, bool bwd) {
  memset(&p, 0, sizeof(p));
  const int D = d.H * d.dh;
  int rc;
  if ((rc = make_tmap_2d_bf16(&p.tmQKV, d.q, uint64_t(d.B) * d.N, 3 * D, 3 * D, TILE, DH, true))) return rc;
  const int rows1 = d.N > TILE ? ((d.N - TILE + 15) & ~15) : TILE;
  if ((rc = make_tmap_2d_bf16(&p.tmQKV1, d.q, uint64_t(d.B) * d.N, 3 * D, 3 * D, rows1, DH, true))) return rc;
  if (bwd) {
    if ((rc = make_tmap_2d_bf16(&p.tmDO, d.d_o, uint64_t(d.B) * d.N, D, D, TILE, DH, true))) return rc;
    if ((rc = make_tmap_2d_bf16(&p.tmDO1, d.d_o, uint64_t(d.B) * d.N, D, D, rows1, DH, true))) return rc;
  } else {
    p.tmDO = p.tmQKV;
  }
  if (!bwd) {
    if ((rc = make_tmap_3d_bf16(&p.tmO3, d.o, d.B, d.N, D, D, uint64_t(d.N) * D, TILE, DH, 1))) return rc;
  }
  p.o = reinterpret_cast<bf16*>(d.o);
  p.o_in = reinterpret_cast<const bf16*>(d.o);
  p.d_o = reinterpret_cast<const bf16*>(d.d_o);
  p.lse = d.lse;
  p.dqkv = reinterpret_cast<bf16*>(d.dq);
  p.B = d.B; p.H = d.H; p.N = d.N;
  p.scale = d.scale;
  p.kv_len = d.kv_len;
  p.causal = d.causal;
  return NGU_OK;
}

}  // namespace

#ifdef NGU_ATTN_TRACE
extern "C" int ngu_debug_attn_trace(void* out, int reset) {
  if (reset) { void* sym; cudaGetSymbolAddress(&sym, g_attn_trace); return cudaMemset(sym, 0, sizeof(g_attn_trace)) != cudaSuccess; }
  return cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(g_attn_trace)) != cudaSuccess;
}
#endif

bool attn_tc_supported(const ngu_attn_desc& d, bool bwd) {
  if (d.dtype != NGU_BF16 || d.dh != DH || d.N > 2 * TILE || !is_packed(d)) return false;
  if (d.causal && bwd) return false;    // causal towers are frozen on this path (CLIP text): forward only on tensor cores
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
  // Default: the pipelined persistent kernel (attn_fwd_pipe_kernel).  NGU_ATTN_FWD=0 / desc.impl = 2: one CTA per (batch, head, query
  // tile), two CTAs per SM (137 us at the ViT-B/16 shape); NGU_ATTN_FWD=1: the older two-group persistent kernel (no masks).
  static const int mode = [] { const char* e = getenv("NGU_ATTN_FWD"); return e ? atoi(e) : 3; }();
  if (mode == 3 && d.impl != 2) {
    const int items = d.B * d.H;
    const dim3 grid(items < sm_count() ? items : sm_count());
    static bool attr3 = false;
    if (!attr3) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_pipe_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd3Smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_pipe_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd3Smem);
      if (e != cudaSuccess) return cuda_status(e, "attn_fwd_pipe attr");
      attr3 = true;
    }
    if (d.N <= 224) launch_pdl(attn_fwd_pipe_kernel<7>, grid, dim3(kFwd3Threads), size_t(kFwd3Smem), st, p);
    else launch_pdl(attn_fwd_pipe_kernel<8>, grid, dim3(kFwd3Threads), size_t(kFwd3Smem), st, p);
    return check_launch("attn_fwd_pipe");
  }
  if (mode != 1 || d.impl == 2 || d.kv_len != nullptr || d.causal) {   // key padding / causal: the old persistent kernel does not mask
    launch_pdl(attn_fwd_tc_kernel, dim3(d.B * d.H * ((d.N + TILE - 1) / TILE)), dim3(kFwdThreads), size_t(kFwdSmem), st, p);
    return check_launch("attn_fwd_tc");
  }
  static bool attr2 = false;
  if (!attr2) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdPSmem);
    if (e != cudaSuccess) return cuda_status(e, "attn_fwd_persistent attr");
    attr2 = true;
  }
  const int items = d.B * d.H;
  launch_pdl(attn_fwd_persistent_kernel, dim3(items < sm_count() ? items : sm_count()), dim3(kFwdPThreads), size_t(kFwdPSmem), st, p);
  return check_launch("attn_fwd_persistent");
}

int attn_bwd_tc(const ngu_attn_desc& d, cudaStream_t st) {
  AttnTcParams p;
  if (int rc = fill_params(d, p, true)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    if (e != cudaSuccess) return cuda_status(e, "attn_bwd_tc attr");
    attr = true;
  }
  const int items = d.B * d.H;
  const int grid = items < sm_count() ? items : sm_count();
  if (d.N > TILE) launch_pdl(attn_bwd_tc_kernel<2>, dim3(grid), dim3(kBwdThreads), size_t(kBwdSmem), st, p);
  else launch_pdl(attn_bwd_tc_kernel<1>, dim3(grid), dim3(kBwdThreads), size_t(kBwdSmem), st, p);
  return check_launch("attn_bwd_tc");
}

}  // namespace ngu
