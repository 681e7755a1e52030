// tcgen05 weight-gradient reduction for the trainable adapter projections:
//     D[Mo, 64] += X[T, Mo]^T · Y[T, 64]        (bf16 in, fp32 accumulate, T = all tokens of the batch)
// Mona project1 / project2 (autograd of src/adapters/mona.py:127,148) and the LoRA A / B factors (lora.py:86,
// rank zero-padded to 64 by the host).  Both operands are consumed straight from their row-major
// activations as MN-major UMMA operands (token index = K), so no transpose is materialised:
//   A = X^T tile [128 x 16 tokens]  : 64-column TMA boxes of X, two boxes per 128-row M tile (LBO = box stride)
//   B = Y^T tile [ 64 x 16 tokens]  : one 64-column box of Y
// Each CTA owns a group of up to 6 M tiles (6 x 64 TMEM columns) and a contiguous token range; the fp32
// partial is added to D with vectorised red.global.add (split-K across ~148 CTAs).
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

constexpr int KB = 64;                 // tokens per pipeline stage
constexpr int kBoxBytes = KB * 128;    // 64 token rows x 64 bf16
constexpr int kStages = 2;
constexpr int kThreads = 32 * 6;
// NB = number of 64-column boxes of Y (No = 64 * NB): NB = 1 -> up to 6 M tiles per CTA (6 x 64 TMEM columns),
// NB = 2 -> up to 3 (3 x 128): Mona's G = x^T [dh | dh*rstd].
template <int NB> struct WgCfg {
  static constexpr int kMaxG = NB == 1 ? 6 : 3;
  static constexpr int kStageBytes = (2 * kMaxG + NB) * kBoxBytes;
  static constexpr int kSmem = kStages * kStageBytes + 1024 + 1024;
};

struct WgradParams {
  CUtensorMap tmX, tmY;  // boxes 64 cols x 64 rows, SWIZZLE_128B
  float* D;
  int ldd, T, Mo, G, splits, kb_per_split;
};

template <int NB>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  pdl_prologue();
  constexpr int kMaxG = WgCfg<NB>::kMaxG;
  constexpr int kStageBytes = WgCfg<NB>::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (kStages + s); };
  const uint32_t done_bar = sBar + 8u * 2 * kStages;
  const uint32_t sTmem = done_bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
  const int total_kb = (p.T + KB - 1) / KB;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(total_kb, kb0 + p.kb_per_split);
  const int nkb = max(0, kb1 - kb0);
  const int G = p.G;
  const int col0 = group * G * 128;  // first X column (= D row) of this CTA

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmY);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(sTmem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));

  if (nkb > 0) {
    if (warp == 0) {
      // whole warp walks, one elected lane issues (uniform registers for the TMA operands)
      int s = 0; uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = base + s * kStageBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), (2 * G + NB) * kBoxBytes);
          for (int c = 0; c < 2 * G; ++c) tma_load_2d(st + c * kBoxBytes, &p.tmX, full_bar(s), col0 + c * 64, kb * KB, kEvictFirst);
#pragma unroll
          for (int c = 0; c < NB; ++c) tma_load_2d(st + (2 * kMaxG + c) * kBoxBytes, &p.tmY, full_bar(s), c * 64, kb * KB, kEvictFirst);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    } else if (warp == 1) {
      // whole warp walks the loop (uniform registers, no per-lane waterfall around tcgen05.mma); one elected lane issues
      constexpr uint32_t idesc = make_idesc_bf16(128, 64 * NB, 1, 1);
      const uint64_t a0 = make_smem_desc_sw128(base, kBoxBytes, 1024);
      const uint64_t b0 = make_smem_desc_sw128(base + 2 * kMaxG * kBoxBytes, NB == 1 ? 0 : kBoxBytes, 1024);
      int s = 0; uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t so = uint64_t((s * kStageBytes) >> 4);
        if (elect_one()) {
          for (int m = 0; m < G; ++m) {
#pragma unroll
            for (int k = 0; k < KB / 16; ++k)
              umma_ss(tmem + m * 64 * NB, a0 + so + uint64_t(((2 * m) * kBoxBytes + k * 2048) >> 4), b0 + so + uint64_t((k * 2048) >> 4), idesc,
                      (kb != kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(done_bar);
      __syncwarp();
    } else {
      const int q = warp & 3;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      for (int mh = 0; mh < G * NB; ++mh) {
        const int m = mh / NB, hb = mh % NB;     // M tile, 64-column half of its accumulator
        uint32_t v[64];
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
        const uint32_t ta = tmem + (uint32_t(q * 32) << 16) + m * 64 * NB + hb * 64;
        tmem_ld32(ta, lo);
        tmem_ld32(ta + 32, hi);
        tmem_ld_wait();
        // Transpose the warp's 32 x 64 fp32 block through smem (the pipeline stages are idle now) so that one red
        // instruction covers two whole 256-byte rows (4 LSU wavefronts) instead of a 16-byte piece of 32 rows (32).
        const uint32_t scr = base + uint32_t(warp - 2) * (32 * 272);          // row pitch 272 B: conflict-free both ways
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(scr + lane * 272 + j * 16), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                       "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
        __syncwarp();
        const int rbase = col0 + m * 128 + q * 32;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int r = 2 * k + (lane >> 4), pc = lane & 15;
          float4 f;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(scr + r * 272 + pc * 16));
          if (rbase + r < p.Mo)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.D + size_t(rbase + r) * p.ldd + hb * 64 + 4 * pc), "f"(f.x), "f"(f.y),
                         "f"(f.z), "f"(f.w) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace

bool wgrad_tc_supported(int ldx, int ldy, int ldd, int Mo, int No, int dtype, const void* X, const void* Y, const float* D) {
  return dtype == NGU_BF16 && (No == 64 || No == 128) && Mo % 128 == 0 && (ldx % 8) == 0 && (ldy % 8) == 0 && (ldd % 4) == 0 &&
         (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0;
}

template <int NB>
static int wgrad_tc_launch(const void* X, int ldx, const void* Y, int ldy, float* D, int ldd, int T, int Mo, cudaStream_t st) {
  constexpr int kMaxG = WgCfg<NB>::kMaxG;
  constexpr int kSmem = WgCfg<NB>::kSmem;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_tmap_2d_bf16(&p.tmX, X, T, Mo, ldx, KB, 64, true))) return rc;
  if ((rc = make_tmap_2d_bf16(&p.tmY, Y, T, 64 * NB, ldy, KB, 64, true))) return rc;
  const int tiles = Mo / 128;
  int G = 1;
  for (int g = kMaxG; g >= 1; --g) if (tiles % g == 0) { G = g; break; }
  const int groups = tiles / G;
  const int total_kb = (T + KB - 1) / KB;
  int splits = sm_count() / groups;
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  p.D = D; p.ldd = ldd; p.T = T; p.Mo = Mo; p.G = G; p.splits = splits;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return cuda_status(e, "wgrad_tc attr");
    attr = true;
  }
  launch_pdl(wgrad_tc_kernel<NB>, dim3(groups * splits), dim3(kThreads), size_t(kSmem), st, p);
  return check_launch("wgrad_tc");
}

int wgrad_tc(const void* X, int ldx, const void* Y, int ldy, float* D, int ldd, int T, int Mo, int No, cudaStream_t st) {
  return No == 128 ? wgrad_tc_launch<2>(X, ldx, Y, ldy, D, ldd, T, Mo, st) : wgrad_tc_launch<1>(X, ldx, Y, ldy, D, ldd, T, Mo, st);
}

}  // namespace ngu
