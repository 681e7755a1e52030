// Fused Mona adapter kernels (bf16 product path) — src/adapters/mona.py:115-151 with the stage of :85-93 inside.
//
// Forward (mona_fwd_stage_kernel, persistent, one CTA per SM, one image at a time, two images in flight):
//   warp 0      TMA producer: 128x64 tiles of x straight from the [B,N,D] activations (3-D map, rows past the
//               sequence end zero-filled) + the matching 128x64 k-block of Wab = [W1*(ln_w*gamma) ; W1*gammax]
//   warp 1      tcgen05.mma issuer: D[128 tokens x 128] += x_tile * Wab_kblock^T  (two token tiles per image), TMEM
//   warps 2..5  LayerNorm statistics of the SAME shared-memory tiles (shifted sum / sum of squares per token row), then
//               the epilogue  h = rstd*Da + Db - mean*rstd*ca + cb  -> bf16 image tile in smem (+ h, hA to HBM)
//   warps 6..13 the bottleneck stage of the previous image out of smem: merged 7x7 depthwise stencil (CUDA cores),
//               1x1 projector on mma.sync, GELU, dropout -> g
//   so x is read from HBM exactly once and the LayerNorm output / pre-scaled u never exist in memory.
// Backward (mona_bwd_stage_kernel): the stage backward per image (recompute z / a from h), emitting dhcat = [dh | dh*rstd]
//   and the two per-row scalars that carry the LayerNorm-backward row terms into the dx GEMM epilogue
//   (NGU_AUX_MONA_DX in gemm_tc.cu); all stage reductions go to a small fp32 workspace and mona_finish_kernel turns the
//   workspace (+ G = x^T dhcat from wgrad_tc.cu) into every parameter gradient.
#include "common.cuh"
#include "kernels.h"
#include "mona_stage.cuh"

namespace ngu {
#ifdef NGU_CONV_PROF
// debug build only: clock64() stamps of CTA 0 (tools/gpu_conv_phases.py)
__device__ long long g_fused_prof[64];
#define NGU_FPROF(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_fused_prof[i] = clock64(); } while (0)
#else
#define NGU_FPROF(i) do { } while (0)
#endif
namespace {
using namespace mona_stage;

// ---------------------------------------------------------------------------------------------------------------
// workspace layout (floats): G [D][128] | Gacc [49][64] | Sz [64] | Sdh [64] | Tmu [64]
// ---------------------------------------------------------------------------------------------------------------
NGU_DEVINL size_t ws_gacc(int D) { return size_t(D) * 128; }
__host__ __device__ inline size_t ws_floats(int D) { return size_t(D) * 128 + 49 * C + 3 * C; }

// ---------------------------------------------------------------------------------------------------------------
// prep: derived operands from the fp32 parameters.  grid (64, n): block (c, item) owns project1 row c / project2 column c.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mona_prep_kernel(const ngu_mona_prep_item* __restrict__ items, int D) {
  pdl_prologue();
  const ngu_mona_prep_item& it = items[blockIdx.y];
  const ngu_mona_params& p = it.p;
  const ngu_mona_derived& d = it.d;
  const int c = blockIdx.x;
  bf16* wab = reinterpret_cast<bf16*>(d.wab);
  bf16* wcat = reinterpret_cast<bf16*>(d.wcat_t);
  bf16* w2 = reinterpret_cast<bf16*>(d.w2);
  bf16* w2t = reinterpret_cast<bf16*>(d.w2_t);
  float sa = 0.f, sb = 0.f;
  for (int k = threadIdx.x; k < D; k += 256) {
    const float w1 = p.w1[size_t(c) * D + k];
    const float gm = p.gamma[k];
    const bf16 a = __float2bfloat16_rn(w1 * p.ln_w[k] * gm);
    const bf16 b = __float2bfloat16_rn(w1 * p.gammax[k]);
    wab[size_t(c) * D + k] = a;
    wab[size_t(C + c) * D + k] = b;
    wcat[size_t(k) * 128 + c] = b;
    wcat[size_t(k) * 128 + C + c] = a;
    sa += __bfloat162float(a);            // ca from the ROUNDED operand: the mean term then cancels exactly what the MMA summed
    sb = fmaf(w1, p.ln_b[k] * gm, sb);
    const bf16 v2 = __float2bfloat16_rn(p.w2[size_t(k) * C + c]);
    w2[size_t(k) * C + c] = v2;
    w2t[size_t(c) * D + k] = v2;
  }
  __shared__ float red[2][8];
  sa = warp_sum(sa); sb = warp_sum(sb);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sa; red[1][threadIdx.x >> 5] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
    for (int i = 0; i < 8; ++i) { ta += red[0][i]; tb += red[1][i]; }
    d.ca[c] = ta;
    d.cb[c] = p.b1[c] + tb;
  }
  if (c == 0) {
    const ngu_mona_conv_weights& w = p.conv;
    for (int i = threadIdx.x; i < 49 * C; i += 256) {
      const int cc = i % C, t = i / C;
      const int ky = t / 7, kx = t % 7;
      float v = w.k7[cc * 49 + t];
      if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) v += w.k5[cc * 25 + (ky - 1) * 5 + (kx - 1)];
      if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) v += w.k3[cc * 9 + (ky - 2) * 3 + (kx - 2)];
      v *= (1.0f / 3.0f) * (w.freq ? w.freq[cc] : 1.0f);
      if (t == 24) v += 1.0f;
      d.kc[i] = v;
    }
    bf16* pb = reinterpret_cast<bf16*>(d.pb);
    for (int i = threadIdx.x; i < C * C; i += 256) pb[swz(i / C, i % C)] = __float2bfloat16_rn(w.P[i]);
    if (threadIdx.x < C) {
      d.bc[threadIdx.x] = (w.b3[threadIdx.x] + w.b5[threadIdx.x] + w.b7[threadIdx.x]) * (1.0f / 3.0f);
      d.bp[threadIdx.x] = w.bp[threadIdx.x];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward stage
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 32 * 14;
constexpr int kBox = 128 * 64 * 2;          // one 128-row x 64-column bf16 tile (128-byte-swizzled)
constexpr int kMaxRingStages = 4;
constexpr int kConvThreads = 256;

struct FwdSmall {
  float kc[49][C];
  float bc[C], bp[C], ca[C], cb[C];
  float clsh[2][C];
};

struct MonaFwdParams {
  CUtensorMap tmX;   // [B, N, D] bf16, box [1, 128, 64]
  CUtensorMap tmX1;  // same tensor, box [1, r1, 64]: second token tile (rows 128 .. 128 + r1), r1 = ceil8(N - 128)
  CUtensorMap tmW;   // [128, D] bf16, box [128, 64]
  int r1, stages;    // rows of the second token tile's box (0 if N <= 128); ring depth
  ngu_mona_derived d;
  bf16* h; bf16* hA; bf16* g;
  float* mean; float* rstd;
  int B, N, H, W, D, has_cls;
  float eps, drop_p;
  uint64_t seed;
  const uint64_t* seed_ctr;
};

NGU_DEVINL void lds128(uint4& v, uint32_t addr) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
}
NGU_DEVINL void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// shifted sums of one 128-byte row (64 bf16): S1 += x - s0, S2 += (x - s0)^2, as four independent packed (f32x2) chains
struct RowAcc { float2 s1[2], s2[2]; };
NGU_DEVINL void row_stats(uint32_t row_addr, uint32_t rot, float2 ms0, RowAcc& A) {
  uint4 ch[8];
  // the order of the pieces of a row is irrelevant for the sums; the rotation by the row index makes the 8 lanes of a
  // quarter-warp read 8 different 16-byte bank groups (row pitch = 128 B: unrotated, every lane would hit the same one)
#pragma unroll
  for (int j = 0; j < 8; ++j) lds128(ch[j], row_addr + ((uint32_t(j) ^ rot) << 4));
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t w[4] = {ch[j].x, ch[j].y, ch[j].z, ch[j].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 d = fadd2(bf16pair_to_float2(w[i]), ms0);
      A.s1[i & 1] = fadd2(A.s1[i & 1], d);
      A.s2[i & 1] = ffma2(d, d, A.s2[i & 1]);
    }
  }
}

__global__ void __launch_bounds__(kFwdThreads, 1) mona_fwd_stage_kernel(const __grid_constant__ MonaFwdParams p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gb = smem_raw + (base - smem_u32(smem_raw));
  const int HW = p.H * p.W, HWp = (HW + 15) & ~15;
  const uint32_t tileB = uint32_t(HWp) * 128u;
  const int r1 = p.r1;
  const uint32_t stageB = uint32_t(2 * kBox + r1 * 128);      // x tile 0 | x tile 1 (r1 rows) | Wab k-block
  const int nst = p.stages;
  const uint32_t oPb = uint32_t(nst) * stageB;
  const uint32_t oZs = oPb + C * C * 2;
  const uint32_t oHs = oZs + tileB;
  const uint32_t oSmall = oHs + 2 * tileB;
  const uint32_t oBar = oSmall + uint32_t((sizeof(FwdSmall) + 15) & ~size_t(15));
  bf16* pbs = reinterpret_cast<bf16*>(gb + oPb);
  bf16* zs = reinterpret_cast<bf16*>(gb + oZs);
  FwdSmall& sm = *reinterpret_cast<FwdSmall*>(gb + oSmall);
  const uint32_t sBar = base + oBar;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (kMaxRingStages + s); };
  auto tfull_bar = [&](int a) { return sBar + 8u * (2 * kMaxRingStages + a); };
  auto tempty_bar = [&](int a) { return sBar + 8u * (2 * kMaxRingStages + 2 + a); };
  auto hsfull_bar = [&](int b) { return sBar + 8u * (2 * kMaxRingStages + 4 + b); };
  auto hsempty_bar = [&](int b) { return sBar + 8u * (2 * kMaxRingStages + 6 + b); };
  const uint32_t sTmemPtr = sBar + 8u * (2 * kMaxRingStages + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.D / 64;
  const bool tile1 = p.N > 128;
  const int N = p.N;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW);
    for (int s = 0; s < nst; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 5); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4);
      mbar_init(hsfull_bar(a), 4); mbar_init(hsempty_bar(a), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(sTmemPtr, 512); tmem_relinquish(); }
  // stage constants (once per CTA)
  for (int i = threadIdx.x; i < 49 * C; i += kFwdThreads) (&sm.kc[0][0])[i] = p.d.kc[i];
  if (threadIdx.x < C) {
    sm.bc[threadIdx.x] = p.d.bc[threadIdx.x]; sm.bp[threadIdx.x] = p.d.bp[threadIdx.x];
    sm.ca[threadIdx.x] = p.d.ca[threadIdx.x]; sm.cb[threadIdx.x] = p.d.cb[threadIdx.x];
  }
  for (int i = threadIdx.x; i < C * C / 8; i += kFwdThreads)
    reinterpret_cast<uint4*>(pbs)[i] = reinterpret_cast<const uint4*>(p.d.pb)[i];
  for (int i = threadIdx.x; i < (HWp - HW) * 8; i += kFwdThreads)   // pad rows of z stay zero
    *reinterpret_cast<uint4*>(zs + (HW + (i >> 3)) * C + ((i & 7) << 3)) = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmemPtr));
  if (warp == 0) NGU_FPROF(0);

  if (warp == 0) {
    // ================================ TMA producer ================================
    int s = 0; uint32_t ph = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = base + s * stageB;
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), stageB);
          tma_load_3d(st, &p.tmX, full_bar(s), kb * 64, 0, img, kEvictFirst);
          if (tile1) tma_load_3d(st + kBox, &p.tmX1, full_bar(s), kb * 64, 128, img, kEvictFirst);
          tma_load_2d(st + kBox + r1 * 128, &p.tmW, full_bar(s), kb * 64, 0, kEvictLast);
        }
        __syncwarp();
        if (++s == nst) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = make_idesc_bf16(128, 128);
    const uint64_t a_base = make_smem_desc_sw128(base, 16, 1024);
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + uint32_t(acc * 256);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t a0 = a_base + uint64_t((s * stageB) >> 4);
        const uint64_t a1 = a0 + uint64_t(kBox >> 4);
        const uint64_t bd = a0 + uint64_t((kBox + r1 * 128) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(d0, a0 + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          if (tile1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(d0 + 128, a1 + uint64_t(k * 2), bd + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          if (kb == nkb - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++s == nst) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 6) {
    // ================================ statistics + projection epilogue ================================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;            // token row inside a 128-row tile
    const float invD = 1.0f / float(p.D);
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x, ++it) {
      const int acc = it & 1;
      float s0a = 0.f, s0b = 0.f;
      RowAcc Aa, Ab;
      Aa.s1[0] = Aa.s1[1] = Aa.s2[0] = Aa.s2[1] = Ab.s1[0] = Ab.s1[1] = Ab.s2[0] = Ab.s2[1] = make_float2(0.f, 0.f);
      const bool row1 = tile1 && r < r1;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        const uint32_t ra = base + s * stageB + uint32_t(r) * 128u;
        if (kb == 0) {
          // first logical element of the row lives in physical 16-byte piece (r & 7)
          uint32_t w0, w1;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(ra + (uint32_t(r & 7) << 4)));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(ra + (row1 ? kBox : 0) + (uint32_t(r & 7) << 4)));
          s0a = unpack_bf16x2(w0).x;
          s0b = unpack_bf16x2(w1).x;
        }
        row_stats(ra, uint32_t(r & 7), make_float2(-s0a, -s0a), Aa);
        if (row1) row_stats(ra + kBox, uint32_t(r & 7), make_float2(-s0b, -s0b), Ab);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
        if (++s == nst) { s = 0; ph ^= 1u; }
      }
      if (warp == 2) NGU_FPROF(24 + it * 3);
      float mu[2], rs[2];
      {
        const float S1a = (Aa.s1[0].x + Aa.s1[0].y) + (Aa.s1[1].x + Aa.s1[1].y), S2a = (Aa.s2[0].x + Aa.s2[0].y) + (Aa.s2[1].x + Aa.s2[1].y);
        const float S1b = (Ab.s1[0].x + Ab.s1[0].y) + (Ab.s1[1].x + Ab.s1[1].y), S2b = (Ab.s2[0].x + Ab.s2[0].y) + (Ab.s2[1].x + Ab.s2[1].y);
        const float ma = S1a * invD, mb = S1b * invD;
        mu[0] = s0a + ma; mu[1] = s0b + mb;
        rs[0] = rsqrtf(fmaxf(S2a * invD - ma * ma, 0.f) + p.eps);
        rs[1] = rsqrtf(fmaxf(S2b * invD - mb * mb, 0.f) + p.eps);
      }
      mbar_wait(tfull_bar(acc), (it >> 1) & 1);
      tc_fence_after();
      mbar_wait(hsempty_bar(acc), ((it >> 1) & 1) ^ 1u);
      if (warp == 2) NGU_FPROF(25 + it * 3);
      bf16* hs = reinterpret_cast<bf16*>(gb + oHs + acc * tileB);
#pragma unroll 1
      for (int tt = 0; tt < (tile1 ? 2 : 1); ++tt) {
        const int t = tt * 128 + r;
        const bool valid = t < N;
        const float m = mu[tt], rstd = rs[tt], mr = m * rstd;
        if (valid) { p.mean[size_t(img) * N + t] = m; p.rstd[size_t(img) * N + t] = rstd; }
        const uint32_t ta = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * 256 + tt * 128);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t va[32], vb[32];
          tmem_ld32(ta + half * 32, va);
          tmem_ld32(ta + 64 + half * 32, vb);
          tmem_ld_wait();
          uint32_t hp[16], ap[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c0 = half * 32 + 2 * j;
            const float ca0 = sm.ca[c0], ca1 = sm.ca[c0 + 1];
            const float da0 = __uint_as_float(va[2 * j]), da1 = __uint_as_float(va[2 * j + 1]);
            const float h0 = fmaf(rstd, da0, __uint_as_float(vb[2 * j])) + fmaf(-mr, ca0, sm.cb[c0]);
            const float h1 = fmaf(rstd, da1, __uint_as_float(vb[2 * j + 1])) + fmaf(-mr, ca1, sm.cb[c0 + 1]);
            hp[j] = pack_bf16x2(h0, h1);
            ap[j] = pack_bf16x2(rstd * fmaf(-m, ca0, da0), rstd * fmaf(-m, ca1, da1));
          }
          if (valid) {
            uint4* hg = reinterpret_cast<uint4*>(p.h + (size_t(img) * N + t) * C + half * 32);
            uint4* ag = reinterpret_cast<uint4*>(p.hA + (size_t(img) * N + t) * C + half * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              hg[j] = make_uint4(hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
              ag[j] = make_uint4(ap[4 * j], ap[4 * j + 1], ap[4 * j + 2], ap[4 * j + 3]);
            }
            if (p.has_cls && t == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 f = unpack_bf16x2(hp[j]);
                sm.clsh[acc][half * 32 + 2 * j] = f.x;
                sm.clsh[acc][half * 32 + 2 * j + 1] = f.y;
              }
            } else {
              const uint32_t ha = smem_u32(hs) + uint32_t(t - p.has_cls) * 128u + uint32_t(half) * 64u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int jj = (j + lane) & 3;     // rotate the piece order so 4 neighbouring rows hit different banks
                sts128(ha + jj * 16, hp[4 * jj], hp[4 * jj + 1], hp[4 * jj + 2], hp[4 * jj + 3]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(tempty_bar(acc)); mbar_arrive(hsfull_bar(acc)); }
      if (warp == 2) NGU_FPROF(26 + it * 3);
    }
  } else {
    // ================================ bottleneck stage (conv warps) ================================
    const int ct = threadIdx.x - 6 * 32;
    const int c = ct & (C - 1), grp = ct >> 6, cw = ct >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const uint32_t zs_u = smem_u32(zs), pb_u = smem_u32(pbs);
    float k[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) k[t] = sm.kc[t][c];
    const float bias = sm.bc[c];
    const float drop_p = p.drop_p;
    const uint64_t seed = mix_seed(p.seed, p.seed_ctr);
    int it = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x, ++it) {
      const int buf = it & 1;
      const bf16* hs = reinterpret_cast<const bf16*>(gb + oHs + buf * tileB);
      bf16* gbp = p.g + size_t(img) * N * C;
      if (cw == 0) NGU_FPROF(8 + it * 4);
      mbar_wait(hsfull_bar(buf), (it >> 1) & 1);
      if (cw == 0) NGU_FPROF(9 + it * 4);
      if (p.has_cls && grp == 0) {
        float v = gelu_erf(sm.clsh[buf][c]);
        if (drop_p > 0.f) v *= dropout_scale(seed, (uint64_t(img) * N) * C + c, drop_p);
        gbp[c] = __float2bfloat16_rn(v);
      }
      for (int x0 = grp * kSW; x0 < p.W; x0 += (kConvThreads / C) * kSW) {
        stencil_stream_x2<false>(hs, k, bias, x0, p.H, p.W, c, [&](int y, const float (&a)[kSW]) {
#pragma unroll
          for (int j = 0; j < kSW; ++j)
            if (x0 + j < p.W) zs[swz(y * p.W + x0 + j, c)] = __float2bfloat16_rn(a[j]);
        });
      }
      named_bar_sync(1, kConvThreads);
      if (cw == 0) NGU_FPROF(10 + it * 4);
      if (ct == 0) mbar_arrive(hsempty_bar(buf));   // the image tile (and clsh) may be overwritten by the next-but-one image
      for (int rt = cw; rt < HWp / 16; rt += kConvThreads / 32) {
        float acc[8][4];
        proj_mma<false>(acc, zs_u, rt, pb_u, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int r = rt * 16 + gq + hh * 8, cc = nt * 8 + 2 * tq;
            if (r < HW) {
              const float2 zz = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(zs + swz(r, cc)));
              float v0 = gelu_erf(acc[nt][2 * hh] + zz.x + sm.bp[cc]);
              float v1 = gelu_erf(acc[nt][2 * hh + 1] + zz.y + sm.bp[cc + 1]);
              if (drop_p > 0.f) {
                const uint64_t e = (uint64_t(img) * N + p.has_cls + r) * C + cc;
                v0 *= dropout_scale(seed, e, drop_p);
                v1 *= dropout_scale(seed, e + 1, drop_p);
              }
              *reinterpret_cast<uint32_t*>(gbp + (p.has_cls + r) * C + cc) = pack_bf16x2(v0, v1);
            }
          }
      }
      named_bar_sync(1, kConvThreads);   // z is rewritten by the next image's stencil
      if (cw == 0) NGU_FPROF(11 + it * 4);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) NGU_FPROF(1);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

size_t fwd_fixed_smem_bytes(int HW) {
  const size_t HWp = (size_t(HW) + 15) & ~size_t(15);
  return 1024 + C * C * 2 + 3 * HWp * 128 + ((sizeof(FwdSmall) + 15) & ~size_t(15)) + 8 * (2 * kMaxRingStages + 8) + 16;
}

// ---------------------------------------------------------------------------------------------------------------
// backward stage: one image per CTA iteration, 2 CTAs per SM
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 256;
constexpr int kMaxTokens = 272;

struct BwdSmall {
  float kc[49][C];
  float bc[C], bp[C], ca[C];
  float S[C];
  float clsdh[C];
  float mr[kMaxTokens];
};

struct MonaBwdParams {
  ngu_mona_derived d;
  const bf16* h; const bf16* hA; const bf16* dg;
  const float* mean; const float* rstd;
  bf16* dhcat; float* rowab;
  float* ws; float* dP; float* dbp;
  int B, N, H, W, D, has_cls;
  float drop_p; uint64_t seed;
  const uint64_t* seed_ctr;
};

__global__ void __launch_bounds__(kBwdThreads, 2) mona_bwd_stage_kernel(const __grid_constant__ MonaBwdParams p) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  const int HW = p.H * p.W, HWp = (HW + 15) & ~15, N = p.N;
  BwdSmall& s = *reinterpret_cast<BwdSmall*>(smem_dyn);
  bf16* pbs = reinterpret_cast<bf16*>(smem_dyn + ((sizeof(BwdSmall) + 1023) & ~size_t(1023)));
  bf16* hs = pbs + C * C;      // h tile (linear); later the dh tile
  bf16* zs = hs + HWp * C;     // z (swizzled), later dz (linear)
  bf16* das = zs + HWp * C;    // da (swizzled); later the fp32 [49][C] correlation buffer
  const int c = threadIdx.x & (C - 1), grp = threadIdx.x >> 6;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  float* Gacc = p.ws + size_t(p.D) * 128;
  float* Sz = Gacc + 49 * C;
  float* Sdh = Sz + C;
  float* Tmu = Sdh + C;

  for (int i = threadIdx.x; i < 49 * C; i += kBwdThreads) (&s.kc[0][0])[i] = p.d.kc[i];
  if (threadIdx.x < C) { s.bc[threadIdx.x] = p.d.bc[threadIdx.x]; s.bp[threadIdx.x] = p.d.bp[threadIdx.x]; s.ca[threadIdx.x] = p.d.ca[threadIdx.x]; }
  for (int i = threadIdx.x; i < C * C / 8; i += kBwdThreads) reinterpret_cast<uint4*>(pbs)[i] = reinterpret_cast<const uint4*>(p.d.pb)[i];
  for (int i = threadIdx.x; i < (HWp - HW) * 8; i += kBwdThreads) {
    const int off = (HW + (i >> 3)) * C + ((i & 7) << 3);
    *reinterpret_cast<uint4*>(hs + off) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(zs + off) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(das + off) = make_uint4(0, 0, 0, 0);
  }
  const uint32_t zs_u = smem_u32(zs), da_u = smem_u32(das), pb_u = smem_u32(pbs);
  const float invD = 1.0f / float(p.D);
  const float drop_p = p.drop_p;
  const uint64_t seed = mix_seed(p.seed, p.seed_ctr);
  float sdh_acc = 0.f, tmu_acc = 0.f;   // channel c, this thread's tokens, all images of this CTA
  float dP_acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) dP_acc[a][b] = 0.f;
  float dbp_acc = 0.f;

  for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
    const bf16* hb = p.h + size_t(img) * N * C;
    const bf16* dgb = p.dg + size_t(img) * N * C;
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(32);
    __syncthreads();   // previous image's row phase has finished with hs / s.mr / s.clsdh; constants are loaded
    {
      const bf16* src = hb + p.has_cls * C;
      for (int i = threadIdx.x; i < HW * 8; i += kBwdThreads) reinterpret_cast<uint4*>(hs)[i] = reinterpret_cast<const uint4*>(src)[i];
      for (int t = threadIdx.x; t < N; t += kBwdThreads) s.mr[t] = p.mean[size_t(img) * N + t] * p.rstd[size_t(img) * N + t];
    }
    if (p.has_cls && grp == 0) {
      float v = __bfloat162float(dgb[c]) * gelu_erf_grad(__bfloat162float(hb[c]));
      if (drop_p > 0.f) v *= dropout_scale(seed, (uint64_t(img) * N) * C + c, drop_p);
      v = __bfloat162float(__float2bfloat16_rn(v));
      s.clsdh[c] = v;
    }
    __syncthreads();
    if (p.has_cls && grp == 0) { sdh_acc += s.clsdh[c]; tmu_acc = fmaf(s.clsdh[c], s.mr[0], tmu_acc); }
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(33);
    // ---- phase 1: z = stencil(h)
    {
      float k[49];
#pragma unroll
      for (int t = 0; t < 49; ++t) k[t] = s.kc[t][c];
      for (int x0 = grp * kSW; x0 < p.W; x0 += (kBwdThreads / C) * kSW) {
        stencil_stream_x2<false>(hs, k, s.bc[c], x0, p.H, p.W, c, [&](int y, const float (&a)[kSW]) {
#pragma unroll
          for (int j = 0; j < kSW; ++j)
            if (x0 + j < p.W) zs[swz(y * p.W + x0 + j, c)] = __float2bfloat16_rn(a[j]);
        });
      }
    }
    __syncthreads();
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(34);
    // ---- phase 2: da = dg * mask * gelu'(z + bp + z P^T)
    for (int rt = warp; rt < HWp / 16; rt += kBwdThreads / 32) {
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
            const float2 gg = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dgb + (p.has_cls + r) * C + cc));
            v0 = gg.x * gelu_erf_grad(acc[nt][2 * hh] + zz.x + s.bp[cc]);
            v1 = gg.y * gelu_erf_grad(acc[nt][2 * hh + 1] + zz.y + s.bp[cc + 1]);
            if (drop_p > 0.f) {
              const uint64_t e = (uint64_t(img) * N + p.has_cls + r) * C + cc;
              v0 *= dropout_scale(seed, e, drop_p);
              v1 *= dropout_scale(seed, e + 1, drop_p);
            }
          }
          *reinterpret_cast<uint32_t*>(das + swz(r, cc)) = pack_bf16x2(v0, v1);
        }
    }
    __syncthreads();
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(35);
    // ---- phase 3: dP[o][i] += sum_p da[p][o] z[p][i] (kept in registers across the images of this CTA);  dbp
    {
      const int mt = warp & 3, nh = warp >> 2;
      for (int kt = 0; kt < HWp / 16; ++kt) {
        uint32_t a[4];
        ldsm_x4_t(a, tile_addr(da_u, kt * 16 + (lane >> 4) * 8 + (lane & 7), mt * 2 + ((lane >> 3) & 1)));
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b[4];
          ldsm_x4_t(b, tile_addr(zs_u, kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), nh * 4 + np * 2 + (lane >> 4)));
          mma16816(dP_acc[2 * np], a, b[0], b[1]);
          mma16816(dP_acc[2 * np + 1], a, b[2], b[3]);
        }
      }
      for (int q = grp; q < HW; q += kBwdThreads / C) dbp_acc += __bfloat162float(das[swz(q, c)]);
    }
    __syncthreads();
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(36);
    // ---- phase 4: dz = da + da P   (overwrites z, LINEAR: only the streaming stencils read it from here on)
    for (int rt = warp; rt < HWp / 16; rt += kBwdThreads / 32) {
      float acc[8][4];
      proj_mma<true>(acc, da_u, rt, pb_u, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int r = rt * 16 + gq + hh * 8, cc = nt * 8 + 2 * tq;
          const float2 dd = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(das + swz(r, cc)));
          *reinterpret_cast<uint32_t*>(zs + r * C + cc) = pack_bf16x2(acc[nt][2 * hh] + dd.x, acc[nt][2 * hh + 1] + dd.y);
        }
    }
    __syncthreads();
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(37);
    // ---- phase 5: correlation sums G[t][c] = sum_p dz[p][c] h[p + off_t][c], S[c] = sum_p dz[p][c]
    float* Gs = reinterpret_cast<float*>(das);
    for (int i = threadIdx.x; i < 49 * C; i += kBwdThreads) Gs[i] = 0.f;
    if (threadIdx.x < C) s.S[threadIdx.x] = 0.f;
    __syncthreads();
    for (int x0 = grp * kSW; x0 < p.W; x0 += (kBwdThreads / C) * kSW) {
      float G[49];
#pragma unroll
      for (int t = 0; t < 49; ++t) G[t] = 0.f;
      float dsum = 0.f;
      float dzb[7][kSW];   // slot s_ <-> dz row yy - 3 + s_
      auto load_dz = [&](int y, float (&dst)[kSW]) {
#pragma unroll
        for (int j = 0; j < kSW; ++j) {
          dst[j] = (unsigned(y) < unsigned(p.H) && x0 + j < p.W) ? __bfloat162float(zs[(y * p.W + x0 + j) * C + c]) : 0.f;
          dsum += dst[j];
        }
      };
#pragma unroll
      for (int s_ = 0; s_ < 7; ++s_) load_dz(s_ - 3, dzb[s_]);
      for (int yy = 0; yy < p.H; ++yy) {
        float win[kSW + 6];
        const bf16* rowp = hs + (yy * p.W) * C + c;
#pragma unroll
        for (int i = 0; i < kSW + 6; ++i) {
          const int xx = x0 + i - 3;
          win[i] = (unsigned(xx) < unsigned(p.W)) ? __bfloat162float(rowp[xx * C]) : 0.f;
        }
#pragma unroll
        for (int s_ = 0; s_ < 7; ++s_) {
          const int ky = 6 - s_;   // h row yy = dz row (yy - 3 + s_) + ky - 3
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
      atomicAdd(&s.S[c], dsum);
    }
    __syncthreads();
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(38);
    // ---- phase 6: flush the per-image correlation sums (mona_finish_kernel turns them into conv / bias / freq grads)
    for (int i = threadIdx.x; i < 49 * C; i += kBwdThreads) atomicAdd(Gacc + i, Gs[i]);
    if (threadIdx.x < C) atomicAdd(Sz + threadIdx.x, s.S[threadIdx.x]);
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(39);
    // ---- phase 7: dh = transposed stencil of dz -> bf16 tile in the (dead) h buffer
    {
      float k[49];
#pragma unroll
      for (int t = 0; t < 49; ++t) k[t] = s.kc[t][c];
      for (int x0 = grp * kSW; x0 < p.W; x0 += (kBwdThreads / C) * kSW) {
        stencil_stream_x2<true>(zs, k, 0.f, x0, p.H, p.W, c, [&](int y, const float (&a)[kSW]) {
#pragma unroll
          for (int j = 0; j < kSW; ++j)
            if (x0 + j < p.W) {
              const int q = y * p.W + x0 + j;
              const bf16 v = __float2bfloat16_rn(a[j]);
              hs[q * C + c] = v;
              const float vf = __bfloat162float(v);
              sdh_acc += vf;
              tmu_acc = fmaf(vf, s.mr[p.has_cls + q], tmu_acc);
            }
        });
      }
    }
    __syncthreads();
    if (warp == 0 && img == int(blockIdx.x)) NGU_FPROF(40);
    // ---- phase 8: per-token row terms + dhcat = [dh | dh * rstd]
    for (int t = threadIdx.x; t < N; t += kBwdThreads) {
      const size_t row = size_t(img) * N + t;
      const float mu = p.mean[row], rstd = p.rstd[row];
      const uint4* ag = reinterpret_cast<const uint4*>(p.hA + row * C);
      uint4* og = reinterpret_cast<uint4*>(p.dhcat + row * 128);
      const bool is_cls = p.has_cls && t == 0;
      const uint32_t ra = smem_u32(hs) + uint32_t(t - p.has_cls) * 128u;
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int jj = (j + t) & 7;       // rotated piece order: neighbouring rows read different banks
        uint4 dv;
        if (is_cls) {
          dv.x = pack_bf16x2(s.clsdh[jj * 8 + 0], s.clsdh[jj * 8 + 1]); dv.y = pack_bf16x2(s.clsdh[jj * 8 + 2], s.clsdh[jj * 8 + 3]);
          dv.z = pack_bf16x2(s.clsdh[jj * 8 + 4], s.clsdh[jj * 8 + 5]); dv.w = pack_bf16x2(s.clsdh[jj * 8 + 6], s.clsdh[jj * 8 + 7]);
        } else {
          lds128(dv, ra + jj * 16);
        }
        const uint4 av = __ldg(ag + jj);
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, aw[4] = {av.x, av.y, av.z, av.w};
        uint32_t sw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 d2 = unpack_bf16x2(dw[i]), a2 = unpack_bf16x2(aw[i]);
          m1 = fmaf(d2.x, s.ca[jj * 8 + 2 * i], m1); m1 = fmaf(d2.y, s.ca[jj * 8 + 2 * i + 1], m1);
          m2 = fmaf(d2.x, a2.x, m2); m2 = fmaf(d2.y, a2.y, m2);
          sw[i] = pack_bf16x2(d2.x * rstd, d2.y * rstd);
        }
        og[jj] = dv;
        og[8 + jj] = make_uint4(sw[0], sw[1], sw[2], sw[3]);
      }
      m1 *= invD; m2 *= invD;
      // dx = dy + GEMM + beta * x + alpha:  -rstd*m1 - rstd*m2*xhat  with xhat = (x - mu) * rstd
      const float beta = -rstd * rstd * m2;
      const float alpha = -rstd * m1 - beta * mu;
      reinterpret_cast<float2*>(p.rowab)[row] = make_float2(alpha, beta);
    }
  }
  if (warp == 0) NGU_FPROF(41);
  // ---- once per CTA: dP, dbp, column sums of dh
  {
    const int mt = warp & 3, nh = warp >> 2;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int o = mt * 16 + gq + hh * 8, i = (nh * 4 + nt) * 8 + 2 * tq;
        atomicAdd(p.dP + o * C + i, dP_acc[nt][2 * hh]);
        atomicAdd(p.dP + o * C + i + 1, dP_acc[nt][2 * hh + 1]);
      }
    atomicAdd(p.dbp + c, dbp_acc);
    atomicAdd(Sdh + c, sdh_acc);
    atomicAdd(Tmu + c, tmu_acc);
  }
  if (warp == 0) NGU_FPROF(42);
}

size_t bwd_smem_bytes(int HW) {
  const size_t HWp = (size_t(HW) + 15) & ~size_t(15);
  const size_t tile = HWp * C * 2;
  const size_t da = tile > size_t(49) * C * sizeof(float) ? tile : size_t(49) * C * sizeof(float);
  return ((sizeof(BwdSmall) + 1023) & ~size_t(1023)) + size_t(C) * C * 2 + 2 * tile + da;
}

// ---------------------------------------------------------------------------------------------------------------
// finish: workspace -> parameter gradients.  blocks [0, D/16): 16 columns k each (16 threads per k, 4 channels each);
// last block: stage (conv) gradients.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mona_finish_kernel(ngu_mona_params p, ngu_mona_grads g, const float* __restrict__ ws, int D) {
  pdl_prologue();
  const float* Gacc = ws + size_t(D) * 128;
  const float* Sz = Gacc + 49 * C;
  const float* Sdh = Sz + C;
  const float* Tmu = Sdh + C;
  if (int(blockIdx.x) < D / 16) {
    // 16 columns k per block, 16 threads per k (4 bottleneck channels each): short dependent chains, 49 blocks at D = 768
    __shared__ float red[3][16][17];
    const int kl = threadIdx.x & 15, cq = threadIdx.x >> 4;
    const int k = blockIdx.x * 16 + kl;
    const float wk = p.ln_w[k], bk = p.ln_b[k], gk = p.gamma[k], gxk = p.gammax[k];
    const float wg = wk * gk, bg = bk * gk;
    const float4 gx4 = *reinterpret_cast<const float4*>(ws + size_t(k) * 128 + cq * 4);
    const float4 gs4 = *reinterpret_cast<const float4*>(ws + size_t(k) * 128 + C + cq * 4);
    const float gxv[4] = {gx4.x, gx4.y, gx4.z, gx4.w}, gsv[4] = {gs4.x, gs4.y, gs4.z, gs4.w};
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const int cc = cq * 4 + ci;
      const float w1 = p.w1[size_t(cc) * D + k];
      const float gx = gxv[ci], gxs = gsv[ci] - Tmu[cc];
      const float sd = Sdh[cc];
      // dW1[c][k] = sum_r dh[r][c] u[r][k],  u = x (rstd*w*gamma + gammax) + (b*gamma - mean*rstd*w*gamma)
      g.dw1[size_t(cc) * D + k] += wg * gxs + gxk * gx + bg * sd;
      s0 = fmaf(sd, w1, s0);
      s1 = fmaf(w1, gx, s1);
      s2 = fmaf(w1, gxs, s2);
    }
    red[0][cq][kl] = s0; red[1][cq][kl] = s1; red[2][cq][kl] = s2;
    __syncthreads();
    if (cq == 0) {
      s0 = s1 = s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { s0 += red[0][i][kl]; s1 += red[1][i][kl]; s2 += red[2][i][kl]; }
      // s0 = sum_r du[r][k], s1 = sum_r du[r][k] x[r][k], s2 = sum_r du[r][k] xhat[r][k]
      g.dgammax[k] += s1;
      g.dln_b[k] += gk * s0;
      g.dln_w[k] += gk * s2;
      g.dgamma[k] += wk * s2 + bk * s0;
    }
  } else {
    const ngu_mona_conv_weights& w = p.conv;
    for (int i = threadIdx.x; i < 49 * C; i += 256) {
      const int cc = i % C, t = i / C;
      const int ky = t / 7, kx = t % 7;
      const float f = w.freq ? w.freq[cc] : 1.0f;
      const float v = Gacc[i] * f * (1.0f / 3.0f);
      g.dk7[cc * 49 + t] += v;
      if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) g.dk5[cc * 25 + (ky - 1) * 5 + (kx - 1)] += v;
      if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) g.dk3[cc * 9 + (ky - 2) * 3 + (kx - 2)] += v;
    }
    if (threadIdx.x < C) {
      const int cc = threadIdx.x;
      const float b = Sz[cc] * (1.0f / 3.0f);
      g.db3[cc] += b; g.db5[cc] += b; g.db7[cc] += b;
      g.db1[cc] += Sdh[cc];
      if (w.freq != nullptr && g.dfreq != nullptr) {
        float q = 0.f;
        for (int t = 0; t < 49; ++t) {
          const int ky = t / 7, kx = t % 7;
          float kw = w.k7[cc * 49 + t];
          if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) kw += w.k5[cc * 25 + (ky - 1) * 5 + (kx - 1)];
          if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) kw += w.k3[cc * 9 + (ky - 2) * 3 + (kx - 2)];
          q = fmaf(kw, Gacc[t * C + cc], q);
        }
        g.dfreq[cc] += q * (1.0f / 3.0f);
      }
    }
  }
}

int validate_stage(const ngu_mona_stage_desc& d, const char* what) {
  if (d.B <= 0 || d.H <= 0 || d.W <= 0 || d.N != d.H * d.W + (d.has_cls ? 1 : 0)) {
    set_last_error("%s: bad shape B=%d N=%d H=%d W=%d has_cls=%d", what, d.B, d.N, d.H, d.W, d.has_cls);
    return NGU_ERR_SHAPE;
  }
  if (d.H > 16 || d.W > 16 || d.N > 256 || (d.D % 64) != 0 || d.D < 64) {
    set_last_error("%s: fused Mona path is built for grids up to 16x16 (N <= 256) and D %% 64 == 0 (got %dx%d, D=%d)", what, d.H, d.W, d.D);
    return NGU_ERR_SHAPE;
  }
  if (d.drop_p < 0.f || d.drop_p >= 1.f) { set_last_error("%s: dropout p=%f out of range", what, d.drop_p); return NGU_ERR_ARG; }
  return NGU_OK;
}

}  // namespace

int64_t mona_ws_floats(int D) { return int64_t(ws_floats(D)); }

int mona_prep(const ngu_mona_prep_item* items, int n, int D, cudaStream_t st) {
  if (n <= 0 || D <= 0 || items == nullptr) { set_last_error("mona_prep: bad arguments"); return NGU_ERR_ARG; }
  launch_pdl(mona_prep_kernel, dim3(dim3(C, n)), dim3(256), size_t(0), st, items, D);
  return check_launch("mona_prep");
}

int mona_fwd_stage(const ngu_mona_stage_desc& d, cudaStream_t st) {
  if (int rc = validate_stage(d, "mona_fwd_stage")) return rc;
  MonaFwdParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_tmap_3d_bf16(&p.tmX, d.x, d.B, d.N, d.D, d.D, uint64_t(d.N) * d.D, 128, 64, 1))) return rc;
  if ((rc = make_tmap_2d_bf16(&p.tmW, d.d.wab, 128, d.D, d.D, 128, 64, 1))) return rc;
  p.r1 = d.N > 128 ? ((d.N - 128 + 7) & ~7) : 0;
  if (p.r1 > 0) { if ((rc = make_tmap_3d_bf16(&p.tmX1, d.x, d.B, d.N, d.D, d.D, uint64_t(d.N) * d.D, p.r1, 64, 1))) return rc; }
  else p.tmX1 = p.tmX;
  const size_t stage_bytes = size_t(2 * kBox + p.r1 * 128), fixed = fwd_fixed_smem_bytes(d.H * d.W);
  p.stages = int((size_t(227 * 1024) - fixed) / stage_bytes);
  if (p.stages > kMaxRingStages) p.stages = kMaxRingStages;
  if (p.stages < 2) { set_last_error("mona_fwd_stage: grid %dx%d leaves no room for the operand ring", d.H, d.W); return NGU_ERR_SHAPE; }
  p.d = d.d;
  p.h = reinterpret_cast<bf16*>(d.h); p.hA = reinterpret_cast<bf16*>(d.hA); p.g = reinterpret_cast<bf16*>(d.g);
  p.mean = d.mean; p.rstd = d.rstd;
  p.B = d.B; p.N = d.N; p.H = d.H; p.W = d.W; p.D = d.D; p.has_cls = d.has_cls;
  p.eps = d.eps; p.drop_p = d.drop_p; p.seed = d.seed; p.seed_ctr = seed_counter();
  const int smem = int(fixed + size_t(p.stages) * stage_bytes);
  cudaError_t e = cudaFuncSetAttribute(mona_fwd_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_status(e, "mona_fwd_stage attr");
  int grid = sm_count();
  if (grid > d.B) grid = d.B;
  launch_pdl(mona_fwd_stage_kernel, dim3(grid), dim3(kFwdThreads), size_t(smem), st, p);
  return check_launch("mona_fwd_stage");
}

int mona_bwd_stage(const ngu_mona_stage_desc& d, cudaStream_t st) {
  if (int rc = validate_stage(d, "mona_bwd_stage")) return rc;
  if (!d.ws || !d.dP || !d.dbp || !d.dhcat || !d.rowab) { set_last_error("mona_bwd_stage: missing output / workspace pointer"); return NGU_ERR_ARG; }
  cudaError_t e = cudaMemsetAsync(d.ws, 0, ws_floats(d.D) * sizeof(float), st);
  if (e != cudaSuccess) return cuda_status(e, "mona_bwd_stage memset");
  MonaBwdParams p;
  memset(&p, 0, sizeof(p));
  p.d = d.d;
  p.h = reinterpret_cast<const bf16*>(d.h); p.hA = reinterpret_cast<const bf16*>(d.hA); p.dg = reinterpret_cast<const bf16*>(d.dg);
  p.mean = d.mean; p.rstd = d.rstd;
  p.dhcat = reinterpret_cast<bf16*>(d.dhcat); p.rowab = d.rowab;
  p.ws = d.ws; p.dP = d.dP; p.dbp = d.dbp;
  p.B = d.B; p.N = d.N; p.H = d.H; p.W = d.W; p.D = d.D; p.has_cls = d.has_cls;
  p.drop_p = d.drop_p; p.seed = d.seed; p.seed_ctr = seed_counter();
  const int smem = int(bwd_smem_bytes(d.H * d.W));
  e = cudaFuncSetAttribute(mona_bwd_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_status(e, "mona_bwd_stage attr");
  int grid = 2 * sm_count();
  if (grid > d.B) grid = d.B;
  launch_pdl(mona_bwd_stage_kernel, dim3(grid), dim3(kBwdThreads), size_t(smem), st, p);
  return check_launch("mona_bwd_stage");
}

int mona_finish(const ngu_mona_params& p, const ngu_mona_grads& g, const float* ws, int D, cudaStream_t st) {
  if ((D % 64) != 0 || ws == nullptr) { set_last_error("mona_finish: bad arguments"); return NGU_ERR_ARG; }
  launch_pdl(mona_finish_kernel, dim3(D / 16 + 1), dim3(256), size_t(0), st, p, g, ws, D);
  return check_launch("mona_finish");
}

#ifdef NGU_CONV_PROF
extern "C" int ngu_debug_fused_prof(long long* out) { return cudaMemcpyFromSymbol(out, g_fused_prof, sizeof(g_fused_prof)) == cudaSuccess ? 0 : -4; }
#endif
}  // namespace ngu
