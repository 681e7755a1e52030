// Device helpers of the Mona bottleneck stage shared by mona_conv.cu (stand-alone stage kernels) and mona_fused.cu
// (fused adapter kernels): the 128-byte-swizzled bf16 tile addressing, ldmatrix / mma.sync m16n8k16 wrappers, the 64x64
// projector contraction on a 16-row tile and the streaming depthwise stencil (src/adapters/mona.py:85-93).
#pragma once
#include "common.cuh"

namespace ngu {
namespace mona_stage {

constexpr int C = 64;  // bottleneck channels (reference default --mona_bottleneck 64)

NGU_DEVINL int swz(int p, int c) { return p * C + ((((c >> 3) ^ (p & 7))) << 3) + (c & 7); }
NGU_DEVINL uint32_t tile_addr(uint32_t base, int row, int chunk) { return base + uint32_t(row) * 128u + (uint32_t((chunk ^ (row & 7))) << 4); }
NGU_DEVINL void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
NGU_DEVINL void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
NGU_DEVINL void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[nt][.] (16 rows x 64 cols) = A[rt*16 .. +16][0..64) * Bm, A from a swizzled tile; Bm(k, n) = Pb[n][k] (TRANS_B = false,
// i.e. A P^T) or Pb[k][n] (TRANS_B = true, i.e. A P), Pb = projector weight [o][i] as a swizzled bf16 tile.
template <bool TRANS_B>
NGU_DEVINL void proj_mma(float (&acc)[8][4], uint32_t tileA, int rt, uint32_t pb, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    ldsm_x4(a, tile_addr(tileA, rt * 16 + (lane & 15), kk * 2 + (lane >> 4)));
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      if (!TRANS_B) ldsm_x4(b, tile_addr(pb, (2 * np + (lane >> 4)) * 8 + (lane & 7), kk * 2 + ((lane >> 3) & 1)));
      else ldsm_x4_t(b, tile_addr(pb, kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), 2 * np + (lane >> 4)));
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// Streaming depthwise stencil on a LINEAR bf16 tile [HW][C]: one thread owns channel c and a 4-wide column strip and
// walks the rows once, keeping the 7 output rows that the current input row touches in a rolling register window
// (10 loads per 196 FMAs instead of 13 per 49).  emit(y, acc[4]) receives each finished output row.
//   FLIP = false: out[y][x] = bias + sum k[ky][kx]   in[y+ky-3][x+kx-3]
//   FLIP = true : out[y][x] = bias + sum k[ky][kx]   in[y-(ky-3)][x-(kx-3)]     (transposed stencil)
constexpr int kSW = 4;  // strip width
template <bool FLIP, typename Emit>
NGU_DEVINL void stencil_stream(const bf16* in, const float (&k)[49], float bias, int x0, int H, int W, int c, Emit emit) {
  float acc[7][kSW];
#pragma unroll
  for (int s_ = 0; s_ < 7; ++s_)
#pragma unroll
    for (int j = 0; j < kSW; ++j) acc[s_][j] = bias;
  for (int yy = 0; yy < H + 3; ++yy) {
    if (yy < H) {
      float win[kSW + 6];
      const bf16* rowp = in + (yy * W) * C + c;
#pragma unroll
      for (int i = 0; i < kSW + 6; ++i) {
        const int xx = x0 + i - 3;
        win[i] = (unsigned(xx) < unsigned(W)) ? __bfloat162float(rowp[xx * C]) : 0.f;
      }
#pragma unroll
      for (int s_ = 0; s_ < 7; ++s_) {
        const int ky = FLIP ? s_ : 6 - s_;   // slot s_ <-> output row yy - 3 + s_
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float kv = k[ky * 7 + (FLIP ? 6 - kx : kx)];
#pragma unroll
          for (int j = 0; j < kSW; ++j) acc[s_][j] = fmaf(kv, win[j + kx], acc[s_][j]);
        }
      }
    }
    const int yo = yy - 3;
    if (yo >= 0) emit(yo, acc[0]);
#pragma unroll
    for (int s_ = 0; s_ < 6; ++s_)
#pragma unroll
      for (int j = 0; j < kSW; ++j) acc[s_][j] = acc[s_ + 1][j];
#pragma unroll
    for (int j = 0; j < kSW; ++j) acc[6][j] = bias;
  }
}

// ---- packed-fp32 variant (sm_100 FFMA2: two fp32 FMAs per instruction) -----------------------------------------
// Same ownership as stencil_stream (one channel, 4-wide column strip, taps in registers); the two FMAs of an FFMA2 are two
// ADJACENT output columns: the tap is a scalar operand (SASS broadcasts a 32-bit register: FFMA2 Rd, Rk.F32, Rwin.F32x2, Racc),
// the window operand is the pair (win[i], win[i+1]).  Pairs starting at even and at odd window positions are kept as two
// register arrays so no operand needs re-packing inside the tap loop.
NGU_DEVINL float2 bf16pair_to_float2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

template <bool FLIP, typename Emit>
NGU_DEVINL void stencil_stream_x2(const bf16* in, const float (&k)[49], float bias, int x0, int H, int W, int c, Emit emit) {
  static_assert(kSW == 4, "two column pairs per strip");
  float2 acc[7][2];
#pragma unroll
  for (int s_ = 0; s_ < 7; ++s_) acc[s_][0] = acc[s_][1] = make_float2(bias, bias);
  for (int yy = 0; yy < H + 3; ++yy) {
    if (yy < H) {
      float win[kSW + 6];
      const bf16* rowp = in + (yy * W) * C + c;
#pragma unroll
      for (int i = 0; i < kSW + 6; ++i) {
        const int xx = x0 + i - 3;
        win[i] = (unsigned(xx) < unsigned(W)) ? __bfloat162float(rowp[xx * C]) : 0.f;
      }
      float2 pe[5], po[4];   // pairs (win[i], win[i+1]) for even / odd i
#pragma unroll
      for (int m = 0; m < 5; ++m) pe[m] = make_float2(win[2 * m], win[2 * m + 1]);
#pragma unroll
      for (int m = 0; m < 4; ++m) po[m] = make_float2(win[2 * m + 1], win[2 * m + 2]);
#pragma unroll
      for (int s_ = 0; s_ < 7; ++s_) {
        const int ky = FLIP ? s_ : 6 - s_;   // slot s_ <-> output row yy - 3 + s_
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float kv = k[ky * 7 + (FLIP ? 6 - kx : kx)];
          const float2 kk = make_float2(kv, kv);
          // output pair j covers columns (2j, 2j+1): window positions 2j + kx, 2j + kx + 1
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int i = 2 * j + kx;
            acc[s_][j] = ffma2(kk, (i & 1) ? po[i >> 1] : pe[i >> 1], acc[s_][j]);
          }
        }
      }
    }
    const int yo = yy - 3;
    if (yo >= 0) {
      const float a[kSW] = {acc[0][0].x, acc[0][0].y, acc[0][1].x, acc[0][1].y};
      emit(yo, a);
    }
#pragma unroll
    for (int s_ = 0; s_ < 6; ++s_) { acc[s_][0] = acc[s_ + 1][0]; acc[s_][1] = acc[s_ + 1][1]; }
    acc[6][0] = acc[6][1] = make_float2(bias, bias);
  }
}

}  // namespace mona_stage
}  // namespace ngu
