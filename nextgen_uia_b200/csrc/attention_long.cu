// tcgen05 attention for LONG sequences (256 < N <= 1024), head dim 64, packed timm / in-proj layout qkv [B*N, 3*H*64]:
// ViT-L/14 @ 336 (N = 577, BASELINE.json configs[3]) and ViT-B/16 @ 352 (N = 485, configs[4]) — the attention core of
// src/third_party/openai_clip/model.py:195-197 and of timm Attention at those geometries.  attention_tc.cu keeps the whole
// key axis of a sequence in one TMEM tile (N <= 256); here the key axis is walked in tiles with an online softmax.
//
//   forward   one CTA per (batch, head, 128-row query tile), 2 CTAs / SM.  Per 128-key tile j:
//               S = Q K_j^T (SS MMA -> TMEM)  ->  two threads per query row: running max / sum, P = exp2(S c - m) written
//               back over S as bf16  ->  O_j = P V_j (TS MMA, V as MN-major B)  ->  O = O * alpha + O_j in registers.
//   backward  FlashAttention-2 split, no atomics:
//     delta   delta[b,h,n] = sum_d dO * O
//     dQ      one CTA per (batch, head, 128-row query tile), 64-key tiles:  S = Q K_j^T, dP = dO V_j^T  ->
//               dS = P (dP - delta) scale (bf16, over S)  ->  dQ += dS K_j (TS MMA, accumulates in TMEM across j)
//     dK/dV   one CTA per (batch, head, 128-key tile), 64-query tiles, transposed so key rows sit on TMEM lanes:
//               S^T = K_j Q_i^T, dP^T = V_j dO_i^T  ->  P^T, dS^T (bf16)  ->  dV += P^T dO_i, dK += dS^T Q_i (TS MMAs)
//   Every Q/K/V/dO tile is one TMA box of the packed activations (128-byte swizzle); the same tile serves as K-major
//   and as MN-major operand, so nothing is transposed in memory.  Rows past the end of a sequence inside a box belong
//   to the next sequence (or are zero-filled past the tensor): they are masked, never multiplied.
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

constexpr int DH = 64;
constexpr int TILE = 128;
constexpr int KT = 64;                       // key / query tile of the backward kernels
constexpr int kTileB = TILE * DH * 2;        // 16 KB
constexpr int kHalfB = KT * DH * 2;          // 8 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kThreads = 32 * 9;             // 8 compute warps (two threads per TMEM lane) + 1 control warp

struct LongParams {
  CUtensorMap tmQKV128, tmQKV64;   // [B*N, 3*H*64], boxes 128 x 64 and 64 x 64
  CUtensorMap tmDO128, tmDO64;     // [B*N, H*64]
  CUtensorMap tmO3;                // forward output [B, N, H*64], box 1 x 128 x 64 (rows past N clipped)
  CUtensorMap tmDQKV3;             // backward output [B, N, 3*H*64], box 1 x 128 x 64
  const bf16* o; const bf16* d_o;
  float* lse; float* delta;
  int B, H, N;
  float scale;
  const int* kv_len;
};

NGU_DEVINL uint64_t desc_k(uint32_t addr) { return make_smem_desc_sw128(addr, 16, 1024); }   // K-major [rows, 64]
NGU_DEVINL uint64_t desc_mn(uint32_t addr) { return make_smem_desc_sw128(addr, 0, 1024); }   // MN-major: K index = 128-byte row

// 32 fp32 values of one row -> bf16 -> 64 bytes at 16-byte pieces [piece0, piece0 + 4) of a 128-byte-swizzled [128, 64] tile row
NGU_DEVINL void stage_row_half(uint32_t tile, int row, int piece0, const uint32_t (&v)[32], float mul) {
  const uint32_t rb = tile + uint32_t(row) * 128u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = rb + ((uint32_t(piece0 + j) ^ uint32_t(row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                 "r"(pack_bf16x2(__uint_as_float(v[8 * j + 0]) * mul, __uint_as_float(v[8 * j + 1]) * mul)),
                 "r"(pack_bf16x2(__uint_as_float(v[8 * j + 2]) * mul, __uint_as_float(v[8 * j + 3]) * mul)),
                 "r"(pack_bf16x2(__uint_as_float(v[8 * j + 4]) * mul, __uint_as_float(v[8 * j + 5]) * mul)),
                 "r"(pack_bf16x2(__uint_as_float(v[8 * j + 6]) * mul, __uint_as_float(v[8 * j + 7]) * mul))
                 : "memory");
  }
}

// =====================================================================================================
// forward
// =====================================================================================================
constexpr int kFwdSmem = 5 * kTileB + 2048 + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 2) attn_fwd_long_kernel(const __grid_constant__ LongParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + kTileB, sV = base + 3 * kTileB;   // K, V: two slots each
  const uint32_t sRed = base + 5 * kTileB;                                 // [2 halves][128 rows] fp32
  const uint32_t sBar = sRed + 2048;
  const uint32_t bar_q = sBar, bar_s = sBar + 8, bar_p = sBar + 16, bar_o = sBar + 24;
  auto bar_kv = [&](int s) { return sBar + 32u + 8u * s; };
  const uint32_t sTmem = sBar + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, D = p.H * DH;
  const int ntq = (N + TILE - 1) / TILE;
  const int t = blockIdx.x % ntq, bh = blockIdx.x / ntq;
  const int b = bh / p.H, h = bh % p.H;
  const int row0 = b * N;
  int Lk = p.kv_len ? __ldg(p.kv_len + b) : N;
  Lk = Lk < 1 ? 1 : (Lk > N ? N : Lk);
  const int nkv = (Lk + TILE - 1) / TILE;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV128);
    mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1);
    mbar_init(bar_kv(0), 1); mbar_init(bar_kv(1), 1);
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(sTmem, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));

  if (warp == 8) {
    // ================================ control warp: TMA + MMA issue ================================
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, kTileB);
      tma_load_2d(sQ, &p.tmQKV128, bar_q, h * DH, row0 + t * TILE);
      mbar_arrive_expect_tx(bar_kv(0), 2 * kTileB);
      tma_load_2d(sK, &p.tmQKV128, bar_kv(0), D + h * DH, row0);
      tma_load_2d(sV, &p.tmQKV128, bar_kv(0), 2 * D + h * DH, row0);
    }
    __syncwarp();
    constexpr uint32_t idesc_o = make_idesc_bf16(TILE, DH, 0, 1);
    const uint64_t dq = desc_k(sQ);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      if (j + 1 < nkv && elect_one()) {      // slot s^1 was released by the wait on bar_o at the end of iteration j-1
        mbar_arrive_expect_tx(bar_kv(s ^ 1), 2 * kTileB);
        tma_load_2d(sK + (s ^ 1) * kTileB, &p.tmQKV128, bar_kv(s ^ 1), D + h * DH, row0 + (j + 1) * TILE);
        tma_load_2d(sV + (s ^ 1) * kTileB, &p.tmQKV128, bar_kv(s ^ 1), 2 * D + h * DH, row0 + (j + 1) * TILE);
      }
      __syncwarp();
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv(s), (j >> 1) & 1);
      tc_fence_after();
      int cols = N - j * TILE; cols = cols > TILE ? TILE : cols;
      const int npad = (cols + 15) & ~15;
      const uint32_t idesc_s = make_idesc_bf16(TILE, npad);
      const uint64_t dk = desc_k(sK + s * kTileB), dv = desc_mn(sV + s * kTileB);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ss(tmem, dq + uint64_t(k * 2), dk + uint64_t(k * 2), idesc_s, k != 0);
        umma_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_p, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const int nsl = npad / 16;
        for (int k = 0; k < nsl; ++k) umma_ts(tmem + 128, tmem + k * 8, dv + uint64_t(k * 128), idesc_o, k != 0);
        umma_commit(bar_o);
      }
      __syncwarp();
      mbar_wait(bar_o, j & 1);      // P / S columns and the K/V slot may be overwritten
    }
  } else {
    // ================================ softmax: two threads per query row ================================
    const int q = warp & 3, hf = warp >> 2;
    const int rt = q * 32 + lane;
    const int r = t * TILE + rt;
    const bool live = t * TILE + q * 32 < N;
    const uint32_t trow = tmem + (uint32_t(q * 32) << 16);
    const float c = p.scale * kLog2e;
    const uint32_t my_red = sRed + 4u * uint32_t(hf * 128 + rt), other_red = sRed + 4u * uint32_t((hf ^ 1) * 128 + rt);
    float o_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      float alpha = 0.f;
      if (live) {
        const int cbase = j * TILE + hf * 64;
        float mx = -INFINITY;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld32(trow + hf * 64 + ch * 32, v);
          tmem_ld_wait();
          if (cbase + ch * 32 + 32 <= Lk) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cbase + ch * 32 + i < Lk) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_red), "f"(mx) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        float omx;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(omx) : "r"(other_red));
        const float m_new = fmaxf(m_run, fmaxf(mx, omx));       // finite: tile j has at least one valid key (j*128 < Lk)
        alpha = (m_run == -INFINITY) ? 0.f : ex2_approx((m_run - m_new) * c);
        const float mc = m_new * c;
        uint32_t pk[2][16];
        float lsum = 0.f;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld32(trow + hf * 64 + ch * 32, v);
          tmem_ld_wait();
          const bool full = cbase + ch * 32 + 32 <= Lk;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, -mc));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
            if (!full) {
              p0 = (cbase + ch * 32 + 2 * i < Lk) ? p0 : 0.f;
              p1 = (cbase + ch * 32 + 2 * i + 1 < Lk) ? p1 : 0.f;
            }
            lsum += p0 + p1;
            pk[ch][i] = pack_bf16x2(p0, p1);
          }
        }
        l_run = fmaf(l_run, alpha, lsum);
        m_run = m_new;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // both threads of the row are done reading S
        tmem_st16(trow + (hf * 2 + 0) * 16, pk[0]);
        tmem_st16(trow + (hf * 2 + 1) * 16, pk[1]);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, j & 1);
      tc_fence_after();
      if (live) {
        uint32_t ov[32];
        tmem_ld32(trow + 128 + hf * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(ov[i]));
      }
    }
    if (live) {
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_red), "f"(l_run) : "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      float ol;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ol) : "r"(other_red));
      const float l = l_run + ol;
      const float inv = rcp_approx(l);
      uint32_t ou[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) ou[i] = __float_as_uint(o_acc[i]);
      stage_row_half(sQ, rt, hf * 4, ou, inv);       // Q tile is dead: every S MMA has completed
      if (r < N && p.lse && hf == 0) p.lse[(size_t(b) * p.H + h) * N + r] = fmaf(m_run, p.scale, 0.6931471805599453f * lg2_approx(l));
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 9, 256;" ::: "memory");
    if (warp == 0 && lane == 0) {
      tma_store_3d(&p.tmO3, sQ, h * DH, t * TILE, b);
      tma_store_commit();
      tma_store_wait_read<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// =====================================================================================================
// backward: delta
// =====================================================================================================
__global__ void __launch_bounds__(128) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta,
                                                         int B, int H, int N) {
  pdl_prologue();
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(B) * N * H) return;
  const int64_t row = idx / H;
  const int h = int(idx % H);
  const uint4* po = reinterpret_cast<const uint4*>(o + (row * H + h) * DH);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + (row * H + h) * DH);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 a = __ldg(po + j), g = __ldg(pd + j);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16x2(aw[i]), y = unpack_bf16x2(gw[i]);
      s = fmaf(x.x, y.x, s);
      s = fmaf(x.y, y.y, s);
    }
  }
  const int b = int(row / N), n = int(row % N);
  delta[(size_t(b) * H + h) * N + n] = s;
}

// =====================================================================================================
// backward: dQ
// =====================================================================================================
constexpr int kDqSmem = 2 * kTileB + 4 * kHalfB + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 2) attn_bwd_long_dq_kernel(const __grid_constant__ LongParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sDO = base + kTileB, sK = base + 2 * kTileB, sV = sK + 2 * kHalfB;   // K, V: two 64-row slots each
  const uint32_t sBar = sV + 2 * kHalfB;
  const uint32_t bar_q = sBar, bar_s = sBar + 8, bar_p = sBar + 16, bar_o = sBar + 24;
  auto bar_kv = [&](int s) { return sBar + 32u + 8u * s; };
  const uint32_t sTmem = sBar + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, D = p.H * DH;
  const int ntq = (N + TILE - 1) / TILE;
  const int t = blockIdx.x % ntq, bh = blockIdx.x / ntq;
  const int b = bh / p.H, h = bh % p.H;
  const int row0 = b * N;
  int Lk = p.kv_len ? __ldg(p.kv_len + b) : N;
  Lk = Lk < 1 ? 1 : (Lk > N ? N : Lk);
  const int nkv = (Lk + KT - 1) / KT;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV128); tma_prefetch_desc(&p.tmQKV64);
    mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1);
    mbar_init(bar_kv(0), 1); mbar_init(bar_kv(1), 1);
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(sTmem, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));
  // TMEM columns: S [0, 64)   dP [64, 128)   dQ [128, 192);  dS (bf16) is written over S at [0, 32)

  if (warp == 8) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, 2 * kTileB);
      tma_load_2d(sQ, &p.tmQKV128, bar_q, h * DH, row0 + t * TILE);
      tma_load_2d(sDO, &p.tmDO128, bar_q, h * DH, row0 + t * TILE);
      mbar_arrive_expect_tx(bar_kv(0), 2 * kHalfB);
      tma_load_2d(sK, &p.tmQKV64, bar_kv(0), D + h * DH, row0);
      tma_load_2d(sV, &p.tmQKV64, bar_kv(0), 2 * D + h * DH, row0);
    }
    __syncwarp();
    constexpr uint32_t idesc_dq = make_idesc_bf16(TILE, DH, 0, 1);
    const uint64_t dq = desc_k(sQ), ddo = desc_k(sDO);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      if (j + 1 < nkv && elect_one()) {
        mbar_arrive_expect_tx(bar_kv(s ^ 1), 2 * kHalfB);
        tma_load_2d(sK + (s ^ 1) * kHalfB, &p.tmQKV64, bar_kv(s ^ 1), D + h * DH, row0 + (j + 1) * KT);
        tma_load_2d(sV + (s ^ 1) * kHalfB, &p.tmQKV64, bar_kv(s ^ 1), 2 * D + h * DH, row0 + (j + 1) * KT);
      }
      __syncwarp();
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv(s), (j >> 1) & 1);
      tc_fence_after();
      int cols = N - j * KT; cols = cols > KT ? KT : cols;
      const int npad = (cols + 15) & ~15;
      const uint32_t idesc_s = make_idesc_bf16(TILE, npad);
      const uint64_t dk = desc_k(sK + s * kHalfB), dv = desc_k(sV + s * kHalfB), dkmn = desc_mn(sK + s * kHalfB);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ss(tmem, dq + uint64_t(k * 2), dk + uint64_t(k * 2), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ss(tmem + 64, ddo + uint64_t(k * 2), dv + uint64_t(k * 2), idesc_s, k != 0);
        umma_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_p, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const int nsl = npad / 16;
        for (int k = 0; k < nsl; ++k) umma_ts(tmem + 128, tmem + k * 8, dkmn + uint64_t(k * 128), idesc_dq, (j | k) != 0);
        umma_commit(bar_o);
      }
      __syncwarp();
      mbar_wait(bar_o, j & 1);
    }
  } else {
    const int q = warp & 3, hf = warp >> 2;
    const int rt = q * 32 + lane;
    const int r = t * TILE + rt;
    const bool live = t * TILE + q * 32 < N;
    const uint32_t trow = tmem + (uint32_t(q * 32) << 16);
    const float c = p.scale * kLog2e;
    float lse2 = INFINITY, dl = 0.f;
    if (r < N) {
      lse2 = p.lse[(size_t(b) * p.H + h) * N + r] * kLog2e;
      dl = p.delta[(size_t(b) * p.H + h) * N + r];
    }
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      if (live) {
        const int cbase = j * KT + hf * 32;
        uint32_t sv[32], dv[32], pk[16];
        tmem_ld32(trow + hf * 32, sv);
        tmem_ld32(trow + 64 + hf * 32, dv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * i]), c, -lse2));
          float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 1]), c, -lse2));
          p0 = (cbase + 2 * i < Lk) ? p0 : 0.f;
          p1 = (cbase + 2 * i + 1 < Lk) ? p1 : 0.f;
          const float d0 = p0 * (__uint_as_float(dv[2 * i]) - dl) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv[2 * i + 1]) - dl) * p.scale;
          pk[i] = pack_bf16x2((cbase + 2 * i < Lk) ? d0 : 0.f, (cbase + 2 * i + 1 < Lk) ? d1 : 0.f);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // both threads of the row are done reading S / dP
        tmem_st16(trow + hf * 16, pk);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, j & 1);
      tc_fence_after();
    }
    if (live) {
      uint32_t ov[32];
      tmem_ld32(trow + 128 + hf * 32, ov);
      tmem_ld_wait();
      stage_row_half(sQ, rt, hf * 4, ov, 1.0f);
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 9, 256;" ::: "memory");
    if (warp == 0 && lane == 0) {
      tma_store_3d(&p.tmDQKV3, sQ, h * DH, t * TILE, b);
      tma_store_commit();
      tma_store_wait_read<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// =====================================================================================================
// backward: dK / dV
// =====================================================================================================
constexpr int kDkvSmem = 2 * kTileB + 4 * kHalfB + 1024 + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 2) attn_bwd_long_dkv_kernel(const __grid_constant__ LongParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = base + kTileB, sQ = base + 2 * kTileB, sDO = sQ + 2 * kHalfB;   // Q, dO: two 64-row slots each
  const uint32_t sTab = sDO + 2 * kHalfB;                                                       // [2 slots][lse2[64], delta[64]] fp32
  const uint32_t sBar = sTab + 1024;
  const uint32_t bar_kv = sBar, bar_s = sBar + 8, bar_p = sBar + 16, bar_o = sBar + 24;
  auto bar_q = [&](int s) { return sBar + 32u + 8u * s; };
  const uint32_t sTmem = sBar + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, D = p.H * DH;
  const int ntk = (N + TILE - 1) / TILE;
  const int jt = blockIdx.x % ntk, bh = blockIdx.x / ntk;
  const int b = bh / p.H, h = bh % p.H;
  const int row0 = b * N;
  int Lk = p.kv_len ? __ldg(p.kv_len + b) : N;
  Lk = Lk < 1 ? 1 : (Lk > N ? N : Lk);
  const int nq = (N + KT - 1) / KT;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQKV128); tma_prefetch_desc(&p.tmQKV64); tma_prefetch_desc(&p.tmDO64);
    mbar_init(bar_kv, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1);
    mbar_init(bar_q(0), 1); mbar_init(bar_q(1), 1);
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(sTmem, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));
  // TMEM columns: S^T [0, 64)   dP^T [64, 128)   dV [128, 192)   dK [192, 256);  P^T (bf16) over S^T at [0, 32), dS^T over dP^T at [64, 96)

  if (warp == 8) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_kv, 2 * kTileB);
      tma_load_2d(sK, &p.tmQKV128, bar_kv, D + h * DH, row0 + jt * TILE);
      tma_load_2d(sV, &p.tmQKV128, bar_kv, 2 * D + h * DH, row0 + jt * TILE);
      mbar_arrive_expect_tx(bar_q(0), 2 * kHalfB);
      tma_load_2d(sQ, &p.tmQKV64, bar_q(0), h * DH, row0);
      tma_load_2d(sDO, &p.tmDO64, bar_q(0), h * DH, row0);
    }
    __syncwarp();
    constexpr uint32_t idesc_acc = make_idesc_bf16(TILE, DH, 0, 1);
    const uint64_t dk = desc_k(sK), dv = desc_k(sV);
    for (int i = 0; i < nq; ++i) {
      const int s = i & 1;
      if (i + 1 < nq && elect_one()) {
        mbar_arrive_expect_tx(bar_q(s ^ 1), 2 * kHalfB);
        tma_load_2d(sQ + (s ^ 1) * kHalfB, &p.tmQKV64, bar_q(s ^ 1), h * DH, row0 + (i + 1) * KT);
        tma_load_2d(sDO + (s ^ 1) * kHalfB, &p.tmDO64, bar_q(s ^ 1), h * DH, row0 + (i + 1) * KT);
      }
      __syncwarp();
      if (i == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q(s), (i >> 1) & 1);
      tc_fence_after();
      int cols = N - i * KT; cols = cols > KT ? KT : cols;
      const int npad = (cols + 15) & ~15;
      const uint32_t idesc_s = make_idesc_bf16(TILE, npad);
      const uint64_t dqk = desc_k(sQ + s * kHalfB), ddok = desc_k(sDO + s * kHalfB);
      const uint64_t dqmn = desc_mn(sQ + s * kHalfB), ddomn = desc_mn(sDO + s * kHalfB);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ss(tmem, dk + uint64_t(k * 2), dqk + uint64_t(k * 2), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ss(tmem + 64, dv + uint64_t(k * 2), ddok + uint64_t(k * 2), idesc_s, k != 0);
        umma_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_p, i & 1);
      tc_fence_after();
      if (elect_one()) {
        const int nsl = npad / 16;
        for (int k = 0; k < nsl; ++k) umma_ts(tmem + 128, tmem + k * 8, ddomn + uint64_t(k * 128), idesc_acc, (i | k) != 0);
        for (int k = 0; k < nsl; ++k) umma_ts(tmem + 192, tmem + 64 + k * 8, dqmn + uint64_t(k * 128), idesc_acc, (i | k) != 0);
        umma_commit(bar_o);
      }
      __syncwarp();
      mbar_wait(bar_o, i & 1);
    }
  } else {
    const int q = warp & 3, hf = warp >> 2;
    const int rt = q * 32 + lane;
    const int g = jt * TILE + rt;                 // key row within the sequence
    const bool rowvalid = g < Lk;
    const uint32_t trow = tmem + (uint32_t(q * 32) << 16);
    const float c = p.scale * kLog2e;
    const int ct = threadIdx.x;                   // 0 .. 255
    for (int i = 0; i < nq; ++i) {
      const int s = i & 1;
      // lse / delta of this query tile -> smem table (slot s was last read two iterations ago)
      if (ct < 2 * KT) {
        const int qi = i * KT + (ct & (KT - 1));
        float v = (ct < KT) ? INFINITY : 0.f;
        if (qi < N) v = (ct < KT) ? p.lse[(size_t(b) * p.H + h) * N + qi] * kLog2e : p.delta[(size_t(b) * p.H + h) * N + qi];
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sTab + uint32_t(s) * 512u + 4u * uint32_t(ct)), "f"(v) : "memory");
      }
      asm volatile("bar.sync 9, 256;" ::: "memory");
      mbar_wait(bar_s, i & 1);
      tc_fence_after();
      {
        uint32_t sv[32], dv[32], pk[16], dk_[16];
        tmem_ld32(trow + hf * 32, sv);
        tmem_ld32(trow + 64 + hf * 32, dv);
        tmem_ld_wait();
        const uint32_t tb = sTab + uint32_t(s) * 512u + 4u * uint32_t(hf * 32);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          float l0, l1, e0, e1;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(l0), "=f"(l1) : "r"(tb + 8u * u));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e0), "=f"(e1) : "r"(tb + 256u + 8u * u));
          float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * u]), c, -l0));       // lse2 = +inf past the sequence end -> 0
          float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * u + 1]), c, -l1));
          p0 = rowvalid ? p0 : 0.f;
          p1 = rowvalid ? p1 : 0.f;
          const float d0 = p0 * (__uint_as_float(dv[2 * u]) - e0) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv[2 * u + 1]) - e1) * p.scale;
          pk[u] = pack_bf16x2(p0, p1);
          dk_[u] = pack_bf16x2(p0 != 0.f ? d0 : 0.f, p1 != 0.f ? d1 : 0.f);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // both threads of the row are done reading S^T / dP^T
        tmem_st16(trow + hf * 16, pk);
        tmem_st16(trow + 64 + hf * 16, dk_);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, i & 1);
      tc_fence_after();
    }
    {
      uint32_t ov[32];
      tmem_ld32(trow + 128 + hf * 32, ov);
      tmem_ld_wait();
      stage_row_half(sV, rt, hf * 4, ov, 1.0f);       // K / V tiles are dead: every MMA has completed
      tmem_ld32(trow + 192 + hf * 32, ov);
      tmem_ld_wait();
      stage_row_half(sK, rt, hf * 4, ov, 1.0f);
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 9, 256;" ::: "memory");
    if (warp == 0 && lane == 0) {
      tma_store_3d(&p.tmDQKV3, sK, D + h * DH, jt * TILE, b);
      tma_store_3d(&p.tmDQKV3, sV, 2 * D + h * DH, jt * TILE, b);
      tma_store_commit();
      tma_store_wait_read<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int fill(const ngu_attn_desc& d, LongParams& p, bool bwd) {
  memset(&p, 0, sizeof(p));
  const int D = d.H * d.dh;
  const uint64_t rows = uint64_t(d.B) * d.N;
  int rc;
  if ((rc = make_tmap_2d_bf16(&p.tmQKV128, d.q, rows, 3 * D, 3 * D, TILE, DH, 1))) return rc;
  if ((rc = make_tmap_2d_bf16(&p.tmQKV64, d.q, rows, 3 * D, 3 * D, KT, DH, 1))) return rc;
  if (bwd) {
    if ((rc = make_tmap_2d_bf16(&p.tmDO128, d.d_o, rows, D, D, TILE, DH, 1))) return rc;
    if ((rc = make_tmap_2d_bf16(&p.tmDO64, d.d_o, rows, D, D, KT, DH, 1))) return rc;
    if ((rc = make_tmap_3d_bf16(&p.tmDQKV3, d.dq, d.B, d.N, 3 * D, 3 * D, uint64_t(d.N) * 3 * D, TILE, DH, 1))) return rc;
  } else {
    if ((rc = make_tmap_3d_bf16(&p.tmO3, d.o, d.B, d.N, D, D, uint64_t(d.N) * D, TILE, DH, 1))) return rc;
  }
  p.o = reinterpret_cast<const bf16*>(d.o);
  p.d_o = reinterpret_cast<const bf16*>(d.d_o);
  p.lse = d.lse;
  p.delta = d.ws;
  p.B = d.B; p.H = d.H; p.N = d.N;
  p.scale = d.scale;
  p.kv_len = d.kv_len;
  return NGU_OK;
}

}  // namespace

bool attn_long_supported(const ngu_attn_desc& d, bool bwd) {
  const int64_t D = int64_t(d.H) * d.dh;
  const char* q = reinterpret_cast<const char*>(d.q);
  const bool packed = d.N == d.S && !d.causal && d.q_ts == 3 * D && d.k_ts == 3 * D && d.v_ts == 3 * D && d.o_ts == D &&
                      d.q_bs == int64_t(d.N) * 3 * D && d.k_bs == d.q_bs && d.v_bs == d.q_bs && d.o_bs == int64_t(d.N) * D &&
                      reinterpret_cast<const char*>(d.k) == q + D * 2 && reinterpret_cast<const char*>(d.v) == q + 4 * D;
  if (d.dtype != NGU_BF16 || d.dh != DH || d.N <= 2 * TILE || d.N > 1024 || !packed) return false;
  if (bwd) {
    const char* dq = reinterpret_cast<const char*>(d.dq);
    if (reinterpret_cast<const char*>(d.dk) != dq + D * 2 || reinterpret_cast<const char*>(d.dv) != dq + D * 4 || d.ws == nullptr) return false;
  }
  return true;
}

int attn_fwd_long(const ngu_attn_desc& d, cudaStream_t st) {
  LongParams p;
  if (int rc = fill(d, p, false)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return cuda_status(e, "attn_fwd_long attr");
    attr = true;
  }
  const int ntq = (d.N + TILE - 1) / TILE;
  launch_pdl(attn_fwd_long_kernel, dim3(d.B * d.H * ntq), dim3(kThreads), size_t(kFwdSmem), st, p);
  return check_launch("attn_fwd_long");
}

int attn_bwd_long(const ngu_attn_desc& d, cudaStream_t st) {
  LongParams p;
  if (int rc = fill(d, p, true)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_long_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_long_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem);
    if (e != cudaSuccess) return cuda_status(e, "attn_bwd_long attr");
    attr = true;
  }
  const int64_t tot = int64_t(d.B) * d.N * d.H;
  launch_pdl(attn_delta_kernel, dim3(unsigned((tot + 127) / 128)), dim3(128), size_t(0), st, p.o, p.d_o, p.delta, d.B, d.H, d.N);
  if (int rc = check_launch("attn_delta")) return rc;
  const int nt = (d.N + TILE - 1) / TILE;
  launch_pdl(attn_bwd_long_dq_kernel, dim3(d.B * d.H * nt), dim3(kThreads), size_t(kDqSmem), st, p);
  if (int rc = check_launch("attn_bwd_long_dq")) return rc;
  launch_pdl(attn_bwd_long_dkv_kernel, dim3(d.B * d.H * nt), dim3(kThreads), size_t(kDkvSmem), st, p);
  return check_launch("attn_bwd_long_dkv");
}

}  // namespace ngu
