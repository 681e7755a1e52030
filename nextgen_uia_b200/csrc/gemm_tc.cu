// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = epi( A[M,K] * B[N,K]^T (+ A2 * B2^T) )
//
// Replaces the cuBLASLt calls behind the reference's nn.Linear layers on the hot path
// (timm Block qkv/proj/fc1/fc2, reference call sites src/adapters/lora.py:78-90 and the
// ResidualAttentionBlock at src/third_party/openai_clip/model.py:177-202) and their dgrads.
//
//   * A and B are bf16, K-major (row-major [rows, K]); frozen weights are kept in both [N,K] and
//     [K,N] copies by the host so forward and dgrad are both "NT" problems.
//   * warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer (+ TMEM allocator),
//     warps 2..5 = epilogue (TMEM -> registers -> swizzled smem slab -> TMA store).
//   * accumulators: 2 x BLOCK_N fp32 columns of TMEM (double buffered so the epilogue of tile i
//     overlaps the MMAs of tile i+1); smem ring of kStages x (A 128x64 + B BLOCK_Nx64) bf16 tiles
//     in the 128-byte-swizzle K-major UMMA layout that TMA writes directly.
//   * optional second operand pair (A2 [M,K2], B2 [N,K2]) is accumulated into the same TMEM tile
//     as extra K blocks: this is how LoRA's s*(x A^T) B^T rides on the base projection.
//   * epilogue: + bias, erf-GELU / QuickGELU (optionally also storing the pre-activation for
//     backward), + residual, or * act'(pre) for the backward through the activation.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace ngu {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumEpiWarps = 16;
constexpr int kGemmThreads = 32 * (2 + kNumEpiWarps);
constexpr int kChunkCols = 32;              // columns per epilogue work item
constexpr int kSlabBytes = 32 * kChunkCols * 2;  // 32 rows x 32 bf16 (64-byte rows, SWIZZLE_64B)

// kAuxTma: number of elementwise [M,N] operands (0, 1 = residual, 2 = Mona dx) that reach the epilogue through per-warp TMA
// rings of 32x32 slabs instead of per-lane global loads.  For the short-K launches (Mona project2 K = 64, Mona dx K = 128) the
// epilogue IS the kernel: per-lane loads of 64-byte row pieces cost 32 LSU wavefronts per instruction and bounded those
// launches at ~2.5 TB/s; the TMA engine writes the slabs without touching the LSU and keeps kAuxDepth slabs per warp in flight.
// kPreU8: the one-byte activation derivative (save_pre == 2) leaves through its own per-warp 32 x 32-byte slab and a TMA store:
// per-lane 16-byte stores (64 partial-sector requests per chunk) cost fc1 ~70 us of 280.
constexpr int kPreSlabBytes = 32 * kChunkCols;
template <int BLOCK_N, bool kPair = false, int kAuxTma = 0, bool kPreU8 = false>
struct GemmCfg {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = (kPair ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;  // pair mode: each CTA holds half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = kNumEpiWarps * (kSlabBytes + (kPreU8 ? kPreSlabBytes : 0));  // one 32x32 slab per epilogue warp (+ derivative bytes)
  static constexpr int kAuxDepth = 2;   // slabs in flight per epilogue warp (3 measured no faster for the short-K launches)
  static constexpr int kAuxBytes = kNumEpiWarps * kAuxDepth * kAuxTma * kSlabBytes;
  static constexpr int kBarBytes = 1024;  // mbarriers + tmem ptr
  static constexpr int kSmemBudget = 227 * 1024 - 1024 /*align slack*/;
  static constexpr int kStagesRaw = (kSmemBudget - kEpiBytes - kAuxBytes - kBarBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kAuxBytes + kBarBytes + 1024;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
};

struct GemmKernelParams {
  CUtensorMap tmA, tmB, tmA2, tmB2, tmC;
  CUtensorMap tmBh, tmB2h;  // B with a half-height box (cluster multicast: each CTA fetches BLOCK_N/2 rows)
  CUtensorMap tmAux, tmAux2;  // kAuxTma: 32 x 32 boxes (SWIZZLE_64B) of the elementwise operands
  CUtensorMap tmPre;          // kPreU8: the [M, N] byte matrix seen as [M, N/2] bf16, boxes of 32 rows x 16 (= 32 bytes)
  void* pre;   // [M, ldpre] pre-activation output (save_pre)
  int ldpre;
  const float* bias;  // [N] fp32 or nullptr
  const bf16* aux;    // [M, ldaux] residual / saved pre-activation, or nullptr
  int ldaux;
  const bf16* aux2;   // [M, ldaux2] second elementwise operand (NGU_AUX_MONA_DX: x), or nullptr
  int ldaux2;
  const float* rowab; // [M, 2] per-row (alpha_r, beta_r) of NGU_AUX_MONA_DX
  float* cf32;        // fp32 output [M, ldcf] (c_dtype == NGU_F32: plain alpha * acc epilogue, InfoNCE logits / feature grads)
  int ldcf;
  int M, N, K, K2;
  int act;       // NGU_ACT_*
  int aux_mode;  // NGU_AUX_*
  int save_pre;  // 1: also store act'(acc + bias) (bf16) in `pre`; 2: as one byte per element (pack_dact4)
  int aux_u8;    // NGU_AUX_DACT operand is the one-byte form
  float alpha;   // scale on the accumulator before bias
  int prefetch;  // L2 prefetch distance for A in k-blocks (0 = off)
};

// Epilogue math for one 32-column chunk of one row: v = raw fp32 accumulators, ax = aux row chunk (bf16x2 words),
// bias_l = this lane's bias[chunk column `lane`].  Compile-time ACT/AUX so the hot loop stays small (the run-time
// switch happens once per chunk, warp-uniform).
// 8-bit activation derivative (save_pre == 2 / NGU_AUX_DACT_U8): act'(pre) of GELU / QuickGELU lies in [-0.17, 1.13]; stored as
// q = round(d * 170 + 43) in one byte (step 1/170 = 0.0059, |error| <= 0.003: the size of a bf16 rounding of a value near 1),
// which halves the bytes fc1 writes for backward and the bytes the dGELU-multiply dgrad reads.
constexpr float kDactScale = 170.0f, kDactZero = 43.0f, kDactOff = kDactZero / kDactScale;
NGU_DEVINL uint32_t pack_dact4(const float (&d)[4]) {
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float qf = fminf(fmaxf(fmaf(d[i], kDactScale, kDactZero + 0.5f), 0.f), 255.f);
    w |= uint32_t(int(qf)) << (8 * i);
  }
  return w;
}
// The same quantisation through the 2^23 magic add: fma rounds d * 170 + 43 to the nearest integer, which then sits in the low
// mantissa byte (d in [-0.25, 1.24] covers every finite input of both activations); three PRMTs gather four bytes.
NGU_DEVINL uint32_t quant_dact4(float2 d01, float2 d23) {
  const float2 k = make_float2(kDactScale, kDactScale), m = make_float2(8388608.0f + kDactZero, 8388608.0f + kDactZero);
  const float2 a = ffma2(d01, k, m), b = ffma2(d23, k, m);
  const uint32_t lo = __byte_perm(__float_as_uint(a.x), __float_as_uint(a.y), 0x0040);
  const uint32_t hi = __byte_perm(__float_as_uint(b.x), __float_as_uint(b.y), 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}

// Packed-pair forms of the two heaviest epilogues of the block.  With 128 x 256 accumulators per tile and a K = 768 mainloop of
// ~6100 clk, the scalar GELU + derivative epilogue (~22 issue slots per element, 5900 clk per tile over four schedulers) set the
// pace of fc1; on register pairs it needs ~10 (FMA pipe: 14 clk per 32 elements) and the tensor pipe is the bound again.
//   fc1 forward (bf16):  y = gelu(acc * alpha + bias), derivative as one byte.  bias4 = this chunk's 32 biases or nullptr.
template <bool SAVE>
NGU_DEVINL void epi_chunk_gelu_x2(const uint32_t (&v)[32], const float4* bias4, float alpha, uint32_t (&outp)[16], uint32_t (&prep)[16]) {
  const float2 al = make_float2(alpha, alpha);
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b = bias4 != nullptr ? __ldg(bias4 + j4) : make_float4(0.f, 0.f, 0.f, 0.f);   // warp-uniform address: one L1 broadcast
    const float2 x01 = ffma2(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), al, make_float2(b.x, b.y));
    const float2 x23 = ffma2(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), al, make_float2(b.z, b.w));
    float2 y01, y23, d01, d23;
    gelu_pair<SAVE>(x01, y01, d01);
    gelu_pair<SAVE>(x23, y23, d23);
    outp[2 * j4] = pack_bf16x2(y01.x, y01.y);
    outp[2 * j4 + 1] = pack_bf16x2(y23.x, y23.y);
    if (SAVE) prep[j4] = quant_dact4(d01, d23);
  }
}
//   fc2 dgrad (bf16):  out = acc * alpha * (q - 43) / 170 with q the saved derivative byte (no bias on this path)
NGU_DEVINL void epi_chunk_dact_u8_x2(const uint32_t (&v)[32], const uint4 (&ax)[4], float alpha, uint32_t (&outp)[16]) {
  const uint32_t* axw = reinterpret_cast<const uint32_t*>(ax);
  const float2 al = make_float2(alpha / kDactScale, alpha / kDactScale);
  const float2 nz = make_float2(-(8388608.0f + kDactZero), -(8388608.0f + kDactZero));
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const uint32_t w = axw[j4];
    // byte -> 2^23 + q (PRMT into the mantissa of 0x4B000000), minus (2^23 + 43): exact
    const float2 e01 = fadd2(make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651))), nz);
    const float2 e23 = fadd2(make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653))), nz);
    const float2 o01 = fmul2(fmul2(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), al), e01);
    const float2 o23 = fmul2(fmul2(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), al), e23);
    outp[2 * j4] = pack_bf16x2(o01.x, o01.y);
    outp[2 * j4 + 1] = pack_bf16x2(o23.x, o23.y);
  }
}

template <int ACT, int AUX, bool U8 = false>
NGU_DEVINL void epi_chunk(const uint32_t (&v)[32], const uint4 (&ax)[4], float bias_l, float alpha, bool save,
                          uint32_t (&outp)[16], uint32_t (&prep)[16]) {
  const uint32_t* axw = reinterpret_cast<const uint32_t*>(ax);
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    float x[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float b = __shfl_sync(0xffffffffu, bias_l, 4 * j4 + i);   // element j of the row needs lane j's bias
      x[i] = fmaf(__uint_as_float(v[4 * j4 + i]), alpha, b);
      d[i] = x[i];
    }
    float a[4];
    if (U8 && AUX == NGU_AUX_DACT) {
      const uint32_t w = axw[j4];   // 4 bytes = 4 derivatives
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = fmaf(float((w >> (8 * i)) & 0xffu), 1.0f / kDactScale, -kDactOff);
    } else {
      const float2 a01 = unpack_bf16x2(axw[2 * j4]);
      const float2 a23 = unpack_bf16x2(axw[2 * j4 + 1]);
      a[0] = a01.x; a[1] = a01.y; a[2] = a23.x; a[3] = a23.y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (AUX == NGU_AUX_DACT) {
        x[i] *= a[i];  // aux holds act'(pre) saved by the forward epilogue
      } else {
        if (ACT == NGU_ACT_GELU) {
          if (save) gelu_and_grad(x[i], x[i], d[i]); else x[i] = gelu_fast(x[i]);
        } else if (ACT == NGU_ACT_QUICKGELU) {
          if (save) quick_gelu_and_grad(x[i], x[i], d[i]); else x[i] = quick_gelu(x[i]);
        }
        if (AUX == NGU_AUX_RESIDUAL) x[i] += a[i];
      }
    }
    if (U8 && AUX != NGU_AUX_DACT) {
      prep[j4] = pack_dact4(d);
    } else {
      prep[2 * j4] = pack_bf16x2(d[0], d[1]);
      prep[2 * j4 + 1] = pack_bf16x2(d[2], d[3]);
    }
    outp[2 * j4] = pack_bf16x2(x[0], x[1]);
    outp[2 * j4 + 1] = pack_bf16x2(x[2], x[3]);
  }
}

// NGU_AUX_MONA_DX epilogue: out = acc + aux + beta_r * aux2 + alpha_r  (backward of the Mona input mix + residual; the
// LayerNorm-backward row terms arrive as two per-row scalars, see mona_fused.cu)
NGU_DEVINL void epi_chunk_dx(const uint32_t (&v)[32], const uint4 (&ax)[4], const uint4 (&bx)[4], float alpha_r, float beta_r,
                             uint32_t (&outp)[16]) {
  const uint32_t* axw = reinterpret_cast<const uint32_t*>(ax);
  const uint32_t* bxw = reinterpret_cast<const uint32_t*>(bx);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 a = unpack_bf16x2(axw[j]);
    const float2 b = unpack_bf16x2(bxw[j]);
    const float o0 = fmaf(beta_r, b.x, __uint_as_float(v[2 * j]) + alpha_r) + a.x;
    const float o1 = fmaf(beta_r, b.y, __uint_as_float(v[2 * j + 1]) + alpha_r) + a.y;
    outp[j] = pack_bf16x2(o0, o1);
  }
}

template <int BLOCK_N, int kCluster, bool kPair, bool kDX = false, int kAuxTma = 0, bool kPreU8 = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ GemmKernelParams p) {
  pdl_prologue();
  static_assert(!kPair || kCluster == 2, "pair mode is a 2-CTA cluster");
  static_assert(kAuxTma == 0 || kAuxTma == (kDX ? 2 : 1), "TMA aux rings: one operand (residual) or two (Mona dx)");
  using Cfg = GemmCfg<BLOCK_N, kPair, kAuxTma, kPreU8>;
  constexpr int kAuxDepth = Cfg::kAuxDepth;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA0 = smem_base;
  const uint32_t sB0 = smem_base + Cfg::kStages * Cfg::kABytes;
  const uint32_t sEpi = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t sPre = sEpi + kNumEpiWarps * kSlabBytes;   // kPreU8: [epilogue warp] 1 KB slabs
  const uint32_t sAux = sEpi + Cfg::kEpiBytes;        // [epilogue warp][slot][operand] 2 KB slabs
  const uint32_t sBar = sAux + Cfg::kAuxBytes;
  // barrier slots (8 bytes each)
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return sBar + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return sBar + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t sTmemPtr = sBar + 8u * (2 * Cfg::kStages + 4);
  auto aux_bar = [&](int ew, int slot) { return sBar + 8u * (2 * Cfg::kStages + 6 + ew * kAuxDepth + slot); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Tiles are handed out per CLUSTER: the kCluster CTAs of a cluster take vertically adjacent 128-row tiles of the
  // same N tile, so the B (weight) tile is common and is fetched once from L2 with TMA multicast (each CTA loads
  // 1/kCluster of it into every CTA's smem).  "Virtual" tile index vt = (pair index) * kCluster + rank.
  const int rank = (kCluster > 1) ? int(cluster_ctarank()) : 0;
  const int m_tiles_real = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int m_tiles = (m_tiles_real + kCluster - 1) / kCluster * kCluster;  // padded so every CTA of a cluster has a tile
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int tile_first = (int(blockIdx.x) / kCluster) * kCluster;          // same for all CTAs of the cluster
  const int tile_stride = int(gridDim.x);
  auto tile_m0 = [&](int t) { return (((t / kCluster) / n_tiles) * kCluster + rank) * BLOCK_M; };
  auto tile_n0 = [&](int t) { return ((t / kCluster) % n_tiles) * BLOCK_N; };
  const int kb1 = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb2 = (p.K2 + BLOCK_K - 1) / BLOCK_K;
  const int num_kb = kb1 + kb2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmC);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kPair ? 1 : kCluster);  // released by the MMA warp(s) of every CTA that reads this slot
    }
    if (kAuxTma > 0)
      for (int w = 0; w < kNumEpiWarps; ++w)
        for (int sl = 0; sl < kAuxDepth; ++sl) mbar_init(aux_bar(w, sl), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), ((BLOCK_N / kChunkCols >= 4) ? kNumEpiWarps : 4 * (BLOCK_N / kChunkCols)) * (kPair ? 2 : 1));  // pair: both CTAs' epilogues
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) { tmem_alloc_2cta(sTmemPtr, Cfg::kTmemCols); tmem_relinquish_2cta(); }
    else { tmem_alloc(sTmemPtr, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // peers' mbarriers must be initialised before any multicast signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmemPtr));

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp walks the loop (warp-uniform values live in uniform registers, where the TMA instructions take their
    // operands); one elected lane issues.  This warp shares its scheduler with four epilogue warps: every instruction
    // it does not execute is a k-block that arrives earlier.
    {
      int s = 0;
      uint32_t ph = 0;
      constexpr int kBRows = BLOCK_N / kCluster;            // rows of the B tile this CTA fetches (and multicasts)
      constexpr uint16_t kMask = uint16_t((1u << kCluster) - 1u);
      for (int t = tile_first; t < num_tiles; t += tile_stride) {
        const int m0 = tile_m0(t);
        const int n0 = tile_n0(t);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const bool main_k = kb < kb1;
          const int kc = (main_k ? kb : kb - kb1) * BLOCK_K;
          const CUtensorMap* ta = main_k ? &p.tmA : &p.tmA2;
          const CUtensorMap* tb = main_k ? (kCluster > 1 ? &p.tmBh : &p.tmB) : (kCluster > 1 ? &p.tmB2h : &p.tmB2);
          const uint32_t da = sA0 + s * Cfg::kABytes, db = sB0 + s * Cfg::kBBytes, fb = full_bar(s);
          if (elect_one()) {
            if (kPair) {
              // both CTAs' loads complete on the LEADER's full barrier (it feeds the single issuing MMA thread)
              const uint32_t lead_full = mapa_cluster(fb, 0);
              if (rank == 0) mbar_arrive_expect_tx(fb, 2 * Cfg::kStageBytes);
              tma_load_2d_2cta(da, ta, lead_full, kc, m0, kEvictNormal);
              tma_load_2d_2cta(db, tb, lead_full, kc, n0 + rank * kBRows, kEvictLast);
            } else {
              mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
              tma_load_2d(da, ta, fb, kc, m0, kEvictNormal);
              // A streams from HBM: warm L2 for the block kPrefetch steps ahead (next tile's rows once this tile's K is done)
              if (main_k && p.prefetch > 0) {
                int pk = kb + p.prefetch, pt = t;
                if (pk >= kb1) { pk -= kb1; pt += tile_stride; }
                if (pt < num_tiles && pk < kb1) tma_prefetch_l2_2d(&p.tmA, pk * BLOCK_K, tile_m0(pt));
              }
              if (kCluster > 1)
                tma_load_2d_mcast(db + rank * kBRows * BLOCK_K * 2, tb, fb, kc, n0 + rank * kBRows, kMask, kEvictLast);
              else
                tma_load_2d(db, tb, fb, kc, n0, kEvictLast);
            }
          }
          __syncwarp();
          if (++s == Cfg::kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp walks the loop (so every value is warp-uniform and lives in uniform registers); one elected
    // lane issues the tcgen05.mma / commit instructions.  Descriptors differ only in their 14-bit address field:
    // desc = base + (byte offset >> 4).
    constexpr uint32_t idesc = make_idesc_bf16(kPair ? 2 * BLOCK_M : BLOCK_M, BLOCK_N);
    const bool issuer = !kPair || rank == 0;   // pair mode: only the leader CTA issues (on behalf of both)
    const uint64_t a_base = make_smem_desc_sw128(sA0, 16, 1024);
    const uint64_t b_base = make_smem_desc_sw128(sB0, 16, 1024);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int t = tile_first; issuer && t < num_tiles; t += tile_stride) {
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + uint32_t(acc * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        // number of valid 16-wide K slices in this block (zero-filled tails are skipped)
        int kext = (kb < kb1) ? (p.K - kb * BLOCK_K) : (p.K2 - (kb - kb1) * BLOCK_K);
        kext = kext > BLOCK_K ? BLOCK_K : kext;
        const uint64_t ad = a_base + uint64_t((s * Cfg::kABytes) >> 4);
        const uint64_t bd = b_base + uint64_t((s * Cfg::kBBytes) >> 4);
        if (elect_one()) {
          const int nk = (kext + UMMA_K - 1) / UMMA_K;
          if (kext == BLOCK_K) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              if (kPair) umma_ss_2cta(d_tmem, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_ss(d_tmem, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            }
          } else {
            for (int k = 0; k < nk; ++k) {
              if (kPair) umma_ss_2cta(d_tmem, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_ss(d_tmem, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          // frees the smem slot once these MMAs have read it (in every CTA whose producer writes into it)
          if (kPair) umma_commit_2cta_mcast(empty_bar(s), 3);
          else if (kCluster > 1) umma_commit_mcast(empty_bar(s), uint16_t((1u << kCluster) - 1u));
          else umma_commit(empty_bar(s));
          if (kb == num_kb - 1) {
            if (kPair) umma_commit_2cta_mcast(tfull_bar(acc), 3); else umma_commit(tfull_bar(acc));
          }
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
    }
  } else {
    // ================================ epilogue ================================
    // 16 warps: warp w owns TMEM lane quarter (w & 3) and BLOCK_N/128 of the tile's 32-column chunks, so four warps
    // share each SM sub-partition and hide each other's TMEM / global-load / MUFU / dependent-issue latencies (the
    // GELU + saved-derivative epilogue needs ~23 issue slots per element; with two warps per scheduler it, not the
    // tensor pipe, set the pace).
    const int ew = warp - 2;
    const int q = warp & 3;
    const int g = ew >> 2;
    constexpr int kChunks = BLOCK_N / kChunkCols;
    constexpr int kPerWarp = kChunks >= 4 ? kChunks / 4 : 1;
    const int c_begin = g * kPerWarp;
    const bool active = g < kChunks;
    const uint32_t slab = sEpi + ew * kSlabBytes;
    const bool use_aux = p.aux_mode != NGU_AUX_NONE;
    const bf16* pre_out = reinterpret_cast<const bf16*>(p.pre);

    const bool aux_u8 = p.aux_mode == NGU_AUX_DACT && p.aux_u8;
    auto load_aux = [&](int t, int c, uint4 (&dst)[4]) {
      const int m0 = tile_m0(t), nc = tile_n0(t) + c * kChunkCols;
      const int row = m0 + q * 32 + lane;
      const bool ok = use_aux && t < num_tiles && nc < p.N && row < p.M;
      if (aux_u8) {   // one byte per element: 32 bytes of this lane's row
        const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.aux) + size_t(ok ? row : 0) * p.ldaux + (ok ? nc : 0));
#pragma unroll
        for (int j = 0; j < 2; ++j) dst[j] = (ok && nc + j * 16 < p.N) ? __ldg(ap + j) : make_uint4(0, 0, 0, 0);
        return;
      }
      const uint4* ap = reinterpret_cast<const uint4*>(p.aux + size_t(ok ? row : 0) * p.ldaux + (ok ? nc : 0));
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = (ok && nc + j * 8 < p.N) ? __ldg(ap + j) : make_uint4(0, 0, 0, 0);
    };
    auto load_aux2 = [&](int t, int c, uint4 (&dst)[4]) {
      const int m0 = tile_m0(t), nc = tile_n0(t) + c * kChunkCols;
      const int row = m0 + q * 32 + lane;
      const bool ok = t < num_tiles && nc < p.N && row < p.M;
      const uint4* ap = reinterpret_cast<const uint4*>(p.aux2 + size_t(ok ? row : 0) * p.ldaux2 + (ok ? nc : 0));
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = (ok && nc + j * 8 < p.N) ? __ldg(ap + j) : make_uint4(0, 0, 0, 0);
    };

    // TMA aux ring: work item n of this warp = (tile tile_first + (n / kPerWarp) * tile_stride, chunk c_begin + n % kPerWarp)
    auto aux_issue = [&](int n) {
      const int t = tile_first + (n / kPerWarp) * tile_stride, c = c_begin + n % kPerWarp;
      if (t >= num_tiles) return;
      const int nc = tile_n0(t) + c * kChunkCols;
      if (nc >= p.N) return;
      const int sl = n % kAuxDepth;
      const uint32_t dst = sAux + uint32_t((ew * kAuxDepth + sl) * kAuxTma) * kSlabBytes, bar = aux_bar(ew, sl);
      if (aux_u8) {   // derivative bytes: 32 rows x 32 B, no swizzle (the byte matrix is mapped as [M, N/2] bf16)
        mbar_arrive_expect_tx(bar, kPreSlabBytes);
        tma_load_2d(dst, &p.tmAux, bar, nc >> 1, tile_m0(t) + q * 32, kEvictFirst);
        return;
      }
      mbar_arrive_expect_tx(bar, kAuxTma * kSlabBytes);
      tma_load_2d(dst, &p.tmAux, bar, nc, tile_m0(t) + q * 32, kEvictFirst);
      if (kAuxTma == 2) tma_load_2d(dst + kSlabBytes, &p.tmAux2, bar, nc, tile_m0(t) + q * 32, kEvictFirst);
    };
    if (active) {
      int acc = 0;
      uint32_t acc_ph = 0;
      int item = 0;
      uint32_t aux_ph = 0;       // bit sl = parity of the next completion of ring slot sl (advances only for armed = live items)
      uint4 axn[4];
      uint4 bxn[kDX ? 4 : 1];
      if (kAuxTma > 0) {
        if (lane == 0)
          for (int n = 0; n < kAuxDepth; ++n) aux_issue(n);
        __syncwarp();
      } else {
        load_aux(tile_first, c_begin, axn);
        if (kDX) load_aux2(tile_first, c_begin, reinterpret_cast<uint4 (&)[4]>(bxn));
      }
      for (int t = tile_first; t < num_tiles; t += tile_stride) {
        const int m0 = tile_m0(t);
        const int n0 = tile_n0(t);
        float2 rab = make_float2(0.f, 0.f);
        if (kDX) { const int row = m0 + q * 32 + lane; if (row < p.M) rab = __ldg(reinterpret_cast<const float2*>(p.rowab) + row); }
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BLOCK_N);
#pragma unroll 1
        for (int ci = 0; ci < kPerWarp; ++ci) {
          const int c = c_begin + ci;
          const int nc = n0 + c * kChunkCols;
          const bool live = nc < p.N;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(t_addr + c * kChunkCols, v);
          float bias_l = 0.f;
          if (p.bias != nullptr && live) bias_l = (nc + lane < p.N) ? __ldg(p.bias + nc + lane) : 0.f;
          // aux chunk was prefetched one work item ago; start fetching the next one now
          uint4 ax[4];
          uint4 bx[kDX ? 4 : 1];
          if (kAuxTma > 0) {
            if (live) {
              // this item's slabs have landed (TMA zero-fills rows / columns past the matrix): lane = row, 64-byte rows, SWIZZLE_64B
              const int sl = item % kAuxDepth;
              mbar_wait(aux_bar(ew, sl), (aux_ph >> sl) & 1u);
              aux_ph ^= 1u << sl;
              const uint32_t src = sAux + uint32_t((ew * kAuxDepth + sl) * kAuxTma) * kSlabBytes + lane * (aux_u8 ? 32 : 64);
              const uint32_t sw = uint32_t(lane >> 1) & 3u;
              if (aux_u8) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ax[j].x), "=r"(ax[j].y), "=r"(ax[j].z), "=r"(ax[j].w) : "r"(src + 16u * j));
              } else
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t a = src + ((uint32_t(j) ^ sw) << 4);
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ax[j].x), "=r"(ax[j].y), "=r"(ax[j].z), "=r"(ax[j].w) : "r"(a));
                if (kDX) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bx[j].x), "=r"(bx[j].y), "=r"(bx[j].z), "=r"(bx[j].w) : "r"(a + kSlabBytes));
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) ax[j] = axn[j];
            if (kDX) {
#pragma unroll
              for (int j = 0; j < 4; ++j) bx[j] = bxn[j];
            }
            if (use_aux) {
              if (ci + 1 < kPerWarp) load_aux(t, c + 1, axn);
              else load_aux(t + tile_stride, c_begin, axn);
            }
            if (kDX) {
              if (ci + 1 < kPerWarp) load_aux2(t, c + 1, reinterpret_cast<uint4 (&)[4]>(bxn));
              else load_aux2(t + tile_stride, c_begin, reinterpret_cast<uint4 (&)[4]>(bxn));
            }
          }
          tmem_ld_wait();
          if (ci == kPerWarp - 1) {
            // all TMEM reads of this accumulator by this warp are done: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (kPair && rank != 0) mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0)); else mbar_arrive(tempty_bar(acc));
            }
          }
          if (!live) { ++item; continue; }
          if (p.cf32 != nullptr) {
            // fp32 output: every lane stores its row's 32 accumulators (scaled) straight to global memory
            const int row = m0 + q * 32 + lane;
            if (row < p.M) {
              float* crow = p.cf32 + size_t(row) * p.ldcf + nc;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (nc + 4 * j < p.N)
                  *reinterpret_cast<float4*>(crow + 4 * j) = make_float4(__uint_as_float(v[4 * j]) * p.alpha, __uint_as_float(v[4 * j + 1]) * p.alpha,
                                                                        __uint_as_float(v[4 * j + 2]) * p.alpha, __uint_as_float(v[4 * j + 3]) * p.alpha);
            }
            ++item;
            continue;
          }

          uint32_t outp[16];
          uint32_t prep[16];
          if (kDX) {
            epi_chunk_dx(v, ax, reinterpret_cast<const uint4 (&)[4]>(bx), rab.x, rab.y, outp);
          } else {
            const bool sv = p.save_pre != 0;
            const int mode = (p.aux_mode == NGU_AUX_DACT) ? 100 : p.act * 3 + p.aux_mode;  // warp-uniform
            switch (mode) {
              case 100:
                if (aux_u8 && p.bias == nullptr) epi_chunk_dact_u8_x2(v, ax, p.alpha, outp);
                else if (aux_u8) epi_chunk<NGU_ACT_NONE, NGU_AUX_DACT, true>(v, ax, bias_l, p.alpha, false, outp, prep);
                else epi_chunk<NGU_ACT_NONE, NGU_AUX_DACT>(v, ax, bias_l, p.alpha, false, outp, prep);
                break;
              case NGU_ACT_NONE * 3 + NGU_AUX_RESIDUAL: epi_chunk<NGU_ACT_NONE, NGU_AUX_RESIDUAL>(v, ax, bias_l, p.alpha, sv, outp, prep); break;
              case NGU_ACT_GELU * 3 + NGU_AUX_NONE:
                if (p.save_pre != 1 && nc + kChunkCols <= p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
                  const float4* b4 = p.bias != nullptr ? reinterpret_cast<const float4*>(p.bias + nc) : nullptr;
                  if (p.save_pre == 2) epi_chunk_gelu_x2<true>(v, b4, p.alpha, outp, prep);
                  else epi_chunk_gelu_x2<false>(v, b4, p.alpha, outp, prep);
                } else if (p.save_pre == 2) epi_chunk<NGU_ACT_GELU, NGU_AUX_NONE, true>(v, ax, bias_l, p.alpha, true, outp, prep);
                else epi_chunk<NGU_ACT_GELU, NGU_AUX_NONE>(v, ax, bias_l, p.alpha, sv, outp, prep);
                break;
              case NGU_ACT_GELU * 3 + NGU_AUX_RESIDUAL: epi_chunk<NGU_ACT_GELU, NGU_AUX_RESIDUAL>(v, ax, bias_l, p.alpha, sv, outp, prep); break;
              case NGU_ACT_QUICKGELU * 3 + NGU_AUX_NONE:
                if (p.save_pre == 2) epi_chunk<NGU_ACT_QUICKGELU, NGU_AUX_NONE, true>(v, ax, bias_l, p.alpha, true, outp, prep);
                else epi_chunk<NGU_ACT_QUICKGELU, NGU_AUX_NONE>(v, ax, bias_l, p.alpha, sv, outp, prep);
                break;
              case NGU_ACT_QUICKGELU * 3 + NGU_AUX_RESIDUAL: epi_chunk<NGU_ACT_QUICKGELU, NGU_AUX_RESIDUAL>(v, ax, bias_l, p.alpha, sv, outp, prep); break;
              default: epi_chunk<NGU_ACT_NONE, NGU_AUX_NONE>(v, ax, bias_l, p.alpha, sv, outp, prep); break;
            }
          }
          if (kAuxTma > 0) {
            // the slabs of this item were consumed into `outp`: refill the slot with the item kAuxDepth ahead
            __syncwarp();
            if (lane == 0) aux_issue(item + kAuxDepth);
          }
          ++item;
          if (lane == 0) tma_store_wait_read<0>();   // previous TMA store has finished reading this warp's slab
          __syncwarp();
          // slab rows are 64 bytes; SWIZZLE_64B: 16-byte piece index ^= (row >> 1) & 3
          const uint32_t rbase = slab + lane * 64;
          const uint32_t rsw = uint32_t(lane >> 1) & 3u;
          if (kPreU8 && p.save_pre == 2) {
            // 8-bit derivative: 32 rows x 32 bytes staged in this warp's byte slab, stored by TMA together with the C chunk
            const uint32_t pa = sPre + ew * kPreSlabBytes + lane * 32;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa), "r"(prep[0]), "r"(prep[1]), "r"(prep[2]), "r"(prep[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa + 16), "r"(prep[4]), "r"(prep[5]), "r"(prep[6]), "r"(prep[7]) : "memory");
          } else if (!kDX && p.save_pre == 2) {
            // 8-bit derivative: this lane's row chunk is 32 bytes = one full DRAM sector, stored directly
            const int row = m0 + q * 32 + lane;
            if (row < p.M) {
              uint4* pb = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.pre) + size_t(row) * p.ldpre + nc);
              if (nc < p.N) pb[0] = make_uint4(prep[0], prep[1], prep[2], prep[3]);
              if (nc + 16 < p.N) pb[1] = make_uint4(prep[4], prep[5], prep[6], prep[7]);
            }
          } else if (!kDX && p.save_pre) {
            // activation derivative (saved for backward): transpose through the slab, then coalesced 64-byte row pieces
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t a = rbase + ((uint32_t(j) ^ rsw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(prep[4 * j]), "r"(prep[4 * j + 1]),
                           "r"(prep[4 * j + 2]), "r"(prep[4 * j + 3]) : "memory");
            }
            __syncwarp();
            const int piece = lane & 3, rsub = lane >> 2;
            bf16* pbase = const_cast<bf16*>(pre_out) + size_t(m0 + q * 32) * p.ldpre + nc + piece * 8;
            uint4 val[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = u * 8 + rsub;
              const uint32_t a = slab + r * 64 + ((uint32_t(piece) ^ (uint32_t(r >> 1) & 3u)) << 4);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val[u].x), "=r"(val[u].y), "=r"(val[u].z), "=r"(val[u].w) : "r"(a));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = u * 8 + rsub;
              if (m0 + q * 32 + r < p.M && nc + piece * 8 < p.N) *reinterpret_cast<uint4*>(pbase + size_t(r) * p.ldpre) = val[u];
            }
            __syncwarp();
          }
          // registers -> swizzled slab -> TMA store (per-warp 32 x 32 box)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t a = rbase + ((uint32_t(j) ^ rsw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(outp[4 * j]),
                         "r"(outp[4 * j + 1]), "r"(outp[4 * j + 2]), "r"(outp[4 * j + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.tmC, slab, nc, m0 + q * 32);
            if (kPreU8 && p.save_pre == 2) tma_store_2d(&p.tmPre, sPre + ew * kPreSlabBytes, nc >> 1, m0 + q * 32);
            tma_store_commit();
          }
        }
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
      if (lane == 0) tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into it / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int kCluster, bool kPair = false, bool kDX = false, int kAuxTma = 0, bool kPreU8 = false>
int launch_gemm_tc(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, kPair, kAuxTma, kPreU8>;
  static_assert(Cfg::kStages >= 2, "operand ring too shallow");
  GemmKernelParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_tmap_2d_bf16(&p.tmA, a.A, a.M, a.K, a.lda, BLOCK_M, BLOCK_K, true))) return rc;
  if ((rc = make_tmap_2d_bf16(&p.tmB, a.B, a.N, a.K, a.ldb, BLOCK_N, BLOCK_K, true))) return rc;
  if ((rc = make_tmap_2d_bf16(&p.tmBh, a.B, a.N, a.K, a.ldb, BLOCK_N / 2, BLOCK_K, true))) return rc;
  if (a.K2 > 0) {
    if ((rc = make_tmap_2d_bf16(&p.tmA2, a.A2, a.M, a.K2, a.lda2, BLOCK_M, BLOCK_K, true))) return rc;
    if ((rc = make_tmap_2d_bf16(&p.tmB2, a.B2, a.N, a.K2, a.ldb2, BLOCK_N, BLOCK_K, true))) return rc;
    if ((rc = make_tmap_2d_bf16(&p.tmB2h, a.B2, a.N, a.K2, a.ldb2, BLOCK_N / 2, BLOCK_K, true))) return rc;
  } else {
    p.tmA2 = p.tmA;
    p.tmB2 = p.tmB;
    p.tmB2h = p.tmBh;
  }
  if (a.c_dtype == NGU_F32) {
    p.tmC = p.tmA;   // unused
    p.cf32 = reinterpret_cast<float*>(a.C);
    p.ldcf = a.ldc;
  } else if ((rc = make_tmap_2d_bf16(&p.tmC, a.C, a.M, a.N, a.ldc, 32, kChunkCols, 2))) return rc;
  if (kAuxTma > 0 && a.aux_mode == NGU_AUX_DACT_U8) {
    if ((rc = make_tmap_2d_bf16(&p.tmAux, a.aux, a.M, a.N / 2, a.ldaux / 2, 32, kChunkCols / 2, 0))) return rc;
    p.tmAux2 = p.tmAux;
  } else if (kAuxTma > 0) {
    if ((rc = make_tmap_2d_bf16(&p.tmAux, a.aux, a.M, a.N, a.ldaux, 32, kChunkCols, 2))) return rc;
    if (kAuxTma == 2) { if ((rc = make_tmap_2d_bf16(&p.tmAux2, a.aux2, a.M, a.N, a.ldaux2, 32, kChunkCols, 2))) return rc; }
    else p.tmAux2 = p.tmAux;
  } else {
    p.tmAux = p.tmA; p.tmAux2 = p.tmA;
  }
  if (kPreU8) {
    if ((rc = make_tmap_2d_bf16(&p.tmPre, a.Pre, a.M, a.N / 2, a.ldpre / 2, 32, kChunkCols / 2, 0))) return rc;
  } else {
    p.tmPre = p.tmA;
  }
  p.pre = a.Pre;
  p.ldpre = a.ldpre;
  p.bias = a.bias;
  p.aux = reinterpret_cast<const bf16*>(a.aux);
  p.ldaux = a.ldaux;
  p.aux2 = reinterpret_cast<const bf16*>(a.aux2);
  p.ldaux2 = a.ldaux2;
  p.rowab = a.rowab;
  p.M = a.M; p.N = a.N; p.K = a.K; p.K2 = a.K2;
  p.act = a.act; p.aux_mode = a.aux_mode; p.save_pre = a.save_pre;
  p.aux_u8 = 0;
  if (a.aux_mode == NGU_AUX_DACT_U8) { p.aux_mode = NGU_AUX_DACT; p.aux_u8 = 1; }
  p.alpha = a.alpha;
  {
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("NGU_GEMM_PREFETCH"); pf = e ? atoi(e) : 0; }  // default off: on B200 the extra TMA issue costs more than the L2 warm-up saves
    p.prefetch = pf;
  }

  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, kCluster, kPair, kDX, kAuxTma, kPreU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return cuda_status(e, "gemm_tc smem attribute");
    attr_done = true;
  }
  const int m_tiles = ((a.M + BLOCK_M - 1) / BLOCK_M + kCluster - 1) / kCluster * kCluster;
  const int n_tiles = (a.N + BLOCK_N - 1) / BLOCK_N;
  int grid = m_tiles * n_tiles;
  const int sms = sm_count() / kCluster * kCluster;
  if (grid > sms) grid = sms;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BLOCK_N, kCluster, kPair, kDX, kAuxTma, kPreU8>, p);
  count_launch(1);
  if (e != cudaSuccess) return cuda_status(e, "gemm_tc launch");
  return cuda_status(cudaGetLastError(), "gemm_tc");
}

}  // namespace

int gemm_tc(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) { set_last_error("gemm_tc: empty problem M=%d N=%d K=%d", a.M, a.N, a.K); return NGU_ERR_SHAPE; }
  if (a.c_dtype == NGU_F32 && (a.bias != nullptr || a.act != NGU_ACT_NONE || a.aux_mode != NGU_AUX_NONE || a.save_pre || (a.N % 4) || (a.ldc % 4) ||
                               (reinterpret_cast<uintptr_t>(a.C) & 15))) {
    set_last_error("gemm_tc: fp32 output supports the plain alpha * acc epilogue only (N, ldc multiples of 4, C 16-byte aligned)");
    return NGU_ERR_ARG;
  }
  if ((a.K % 8) || (a.lda % 8) || (a.ldb % 8) || ((a.ldc % 8) && a.c_dtype != NGU_F32) || (a.K2 % 8)) {
    set_last_error("gemm_tc: K, K2 and leading dimensions must be multiples of 8 (16-byte rows)");
    return NGU_ERR_ALIGN;
  }
  if (a.aux_mode == NGU_AUX_DACT_U8 && (a.aux == nullptr || (a.ldaux % 16) || (a.N % 16) || (reinterpret_cast<uintptr_t>(a.aux) & 15))) {
    set_last_error("gemm_tc: NGU_AUX_DACT_U8 needs a 16-byte aligned aux pointer, ldaux %% 16 == 0 and N %% 16 == 0");
    return NGU_ERR_ALIGN;
  }
  if (a.save_pre == 2 && (a.act == NGU_ACT_NONE || a.aux_mode != NGU_AUX_NONE || (a.ldpre % 16) || (a.N % 16) || (reinterpret_cast<uintptr_t>(a.Pre) & 15))) {
    set_last_error("gemm_tc: save_pre = 2 (one-byte derivative) needs an activation, no aux operand, ldpre %% 16 == 0, N %% 16 == 0, aligned Pre");
    return NGU_ERR_ARG;
  }
  if (a.aux_mode != NGU_AUX_NONE && a.aux_mode != NGU_AUX_DACT_U8 && (a.aux == nullptr || (a.ldaux % 8) || (a.N % 8))) {
    set_last_error("gemm_tc: aux operand needs a pointer, ldaux %% 8 == 0 and N %% 8 == 0");
    return NGU_ERR_ALIGN;
  }
  if (a.save_pre && (a.Pre == nullptr || (a.ldpre % 8) || (a.N % 8))) {
    set_last_error("gemm_tc: save_pre needs a Pre pointer, ldpre %% 8 == 0 and N %% 8 == 0");
    return NGU_ERR_ALIGN;
  }
  if (a.aux_mode == NGU_AUX_MONA_DX) {
    if (a.aux2 == nullptr || a.rowab == nullptr || (a.ldaux2 % 8) || a.act != NGU_ACT_NONE || a.save_pre || a.bias != nullptr || a.alpha != 1.0f) {
      set_last_error("gemm_tc: NGU_AUX_MONA_DX needs aux, aux2 (ldaux2 %% 8 == 0), rowab and no bias / activation / save_pre / alpha");
      return NGU_ERR_ARG;
    }
    // short K (128), two streamed elementwise operands: the epilogue IS the kernel -> independent-CTA multicast variant with both
    // operands staged through the per-warp TMA rings (NGU_GEMM_AUXTMA=0: per-lane loads, A/B switch)
    static const int auxtma = [] { const char* e = getenv("NGU_GEMM_AUXTMA"); return e ? atoi(e) : 1; }();
    if (a.M <= BLOCK_M) return launch_gemm_tc<256, 1, false, true>(a, stream);
    if (auxtma && a.K + a.K2 <= 256) return launch_gemm_tc<128, 2, false, true, 2>(a, stream);
    return launch_gemm_tc<256, 2, false, true>(a, stream);
  }
  if (a.aux_mode == NGU_AUX_RESIDUAL && a.act == NGU_ACT_NONE && !a.save_pre && a.c_dtype != NGU_F32 && a.block_n == 0 && a.M > BLOCK_M &&
      a.N > 128 && a.K + a.K2 <= 256) {
    static const int auxtma = [] { const char* e = getenv("NGU_GEMM_AUXTMA"); return e ? atoi(e) : 1; }();
    // Mona project2 + residual (K = 64): the epilogue is the kernel; NGU_GEMM_AUXTMA=2 selects the 128-column tile (the Mona dx shape)
    if (auxtma == 2) return launch_gemm_tc<128, 2, false, false, 1>(a, stream);
    if (auxtma) return launch_gemm_tc<256, 2, false, false, 1>(a, stream);
  }
  if (a.save_pre == 2 && a.block_n == 0 && a.M > BLOCK_M && a.N > 128) {
    // fc1 forward: bias + activation + one-byte derivative, both outputs through TMA stores (NGU_GEMM_PREU8: 0 off, 1 multicast cluster, 2 CTA pair)
    static const int preu8 = [] { const char* e = getenv("NGU_GEMM_PREU8"); return e ? atoi(e) : 1; }();
    if (preu8 == 1) return launch_gemm_tc<256, 2, false, false, 0, true>(a, stream);
    if (preu8 == 2) return launch_gemm_tc<256, 2, true, false, 0, true>(a, stream);
  }
  if (a.block_n == 0 && a.M > BLOCK_M && a.N > 128 && a.c_dtype != NGU_F32 && a.act == NGU_ACT_NONE && !a.save_pre &&
      ((a.aux_mode == NGU_AUX_DACT_U8 && a.bias == nullptr) || a.aux_mode == NGU_AUX_RESIDUAL)) {
    // elementwise operand (residual, or the derivative bytes of the fc2 dgrad) through the per-warp TMA rings of the CTA-pair variant
    // (NGU_GEMM_AUXPAIR: bit 0 residual, bit 1 derivative bytes)
    static const int auxpair = [] { const char* e = getenv("NGU_GEMM_AUXPAIR"); return e ? atoi(e) : 3; }();
    if (auxpair & (a.aux_mode == NGU_AUX_RESIDUAL ? 1 : 2)) return launch_gemm_tc<256, 2, true, false, 1>(a, stream);
  }
  int bn = a.block_n;
  if (bn == 0) bn = (a.N > 128) ? 256 : (a.N > 64 ? 128 : 64);
  // block_n + 2000 forces the CTA-pair (cta_group::2) variant, block_n + 1000 the single-CTA (no multicast) one: tests / tuning
  if (a.block_n >= 2000) {
    if (a.M <= BLOCK_M) { set_last_error("gemm_tc: pair mode needs M > 128"); return NGU_ERR_ARG; }
    switch (a.block_n - 2000) {
      case 0: case 256: return launch_gemm_tc<256, 2, true>(a, stream);
      case 128: return launch_gemm_tc<128, 2, true>(a, stream);
      default: set_last_error("gemm_tc: pair mode supports block_n 128/256"); return NGU_ERR_ARG;
    }
  }
  // Auto: the CTA-pair variant (256x256 tile per SM pair, half the smem traffic per MMA, 6-deep pipeline) wins when the
  // mainloop dominates; epilogue-heavy launches (GELU + saved derivative, derivative multiply) with a short K stay on the
  // independent-CTA multicast variant, where one CTA's epilogue never stalls its peer's accumulator.
  // NGU_GEMM_PAIR: 0 never, 1 heuristic (default), 2 whenever legal.
  static const int pair_mode = [] { const char* e = getenv("NGU_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  if (a.block_n == 0 && bn == 256 && a.M > BLOCK_M && pair_mode > 0) {
    const bool heavy_epi = a.act != NGU_ACT_NONE || a.save_pre || a.aux_mode == NGU_AUX_DACT || a.aux_mode == NGU_AUX_DACT_U8;
    if (pair_mode >= 2 || !heavy_epi || a.K + a.K2 >= 2048) return launch_gemm_tc<256, 2, true>(a, stream);
  }
  const bool no_cluster = a.block_n >= 1000;
  if (no_cluster) bn = a.block_n - 1000 ? a.block_n - 1000 : ((a.N > 128) ? 256 : (a.N > 64 ? 128 : 64));
  const bool cluster = !no_cluster && a.M > BLOCK_M;
  switch (bn) {
    case 256: return cluster ? launch_gemm_tc<256, 2>(a, stream) : launch_gemm_tc<256, 1>(a, stream);
    case 128: return cluster ? launch_gemm_tc<128, 2>(a, stream) : launch_gemm_tc<128, 1>(a, stream);
    case 64: return cluster ? launch_gemm_tc<64, 2>(a, stream) : launch_gemm_tc<64, 1>(a, stream);
    default: set_last_error("gemm_tc: unsupported block_n %d", bn); return NGU_ERR_ARG;
  }
}

}  // namespace ngu
