// Shared device/host helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX
// wrappers, UMMA descriptors, small math (erf-GELU), warp reductions, error plumbing.
// Everything here is written against the PTX ISA for sm_100a (CUDA 12.9); no CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define NGU_DEVINL __device__ __forceinline__

namespace ngu {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------
// error codes of the C-ABI (include/ngu_b200.h)
// ----------------------------------------------------------------------------------------
enum : int {
  NGU_OK = 0,
  NGU_ERR_SHAPE = -1,
  NGU_ERR_ALIGN = -2,
  NGU_ERR_DTYPE = -3,
  NGU_ERR_CUDA = -4,
  NGU_ERR_ARG = -5,
};

void set_last_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);
int check_launch(const char* what);

// ----------------------------------------------------------------------------------------
// generic device helpers
// ----------------------------------------------------------------------------------------
NGU_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
NGU_DEVINL uint32_t lane_id() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
  return r;
}
NGU_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

template <typename T>
NGU_DEVINL T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
NGU_DEVINL T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of this library starts with pdl_prologue() and is launched through
// launch_pdl() / with the programmatic-stream-serialization attribute, so the launch of kernel i+1 (grid setup, CTA
// scheduling) overlaps the tail of kernel i instead of following its completion.  griddepcontrol.wait returns once ALL
// prerequisite grids have completed and flushed, so no global memory is touched before the data it depends on exists;
// launch_dependents is issued right after so the next grid can be scheduled as soon as SM resources free up.
// Both are no-ops when the kernel was launched without the attribute (the default: NGU_PDL=1 enables it; on the
// benchmark step it measured ~3 % SLOWER than plain stream order under a CUDA graph, profiles/r2_notes.md).
// ----------------------------------------------------------------------------------------
NGU_DEVINL void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifndef NGU_PDL_NO_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
NGU_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
NGU_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
NGU_DEVINL void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
NGU_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
NGU_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
NGU_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin on an mbarrier phase (try_wait suspends in hardware between probes).  A spin-count watchdog traps
// instead of wedging the GPU if a barrier protocol bug ever leaves a phase incomplete.
NGU_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
      printf("ngu: mbarrier watchdog block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2-D tiles
// ----------------------------------------------------------------------------------------
NGU_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// L2 cache-policy constants (same encodings CUTLASS uses for TMA::CacheHintSm90)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

NGU_DEVINL void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                            uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// multicast variant: the box lands at the same smem offset of every CTA in cta_mask and completes tx bytes on the
// mbarrier at the same offset in each of them
NGU_DEVINL void tma_load_2d_mcast(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t cta_mask,
                                  uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5, %6;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask), "l"(hint)
      : "memory");
}
NGU_DEVINL void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}
// pull a tile into L2 ahead of the TMA load that will consume it (no smem, no barrier)
NGU_DEVINL void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
NGU_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
NGU_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
NGU_DEVINL void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
NGU_DEVINL void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
NGU_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
NGU_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
NGU_DEVINL void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
NGU_DEVINL void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp, .sync.aligned
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
NGU_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
NGU_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
NGU_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
NGU_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; one thread issues on behalf of the CTA.
NGU_DEVINL void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
NGU_DEVINL void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Predicated forms: the instruction is skipped when `enable` is 0 (branch-free issue loops with run-time trip counts)
NGU_DEVINL void umma_ss_if(uint32_t enable, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(enable)
      : "memory");
}
NGU_DEVINL void umma_ts_if(uint32_t enable, uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(enable)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05 ops of this thread have completed.
NGU_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// same, arriving on the mbarrier at this smem offset in every CTA of cta_mask
NGU_DEVINL void umma_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}

// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster issue one 256-row MMA; the leader (rank 0) issues
NGU_DEVINL void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
NGU_DEVINL void tmem_relinquish_2cta() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
NGU_DEVINL void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
NGU_DEVINL void umma_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
NGU_DEVINL void umma_commit_2cta_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
// shared::cluster address of `local_addr` in the CTA with rank `rank`
NGU_DEVINL uint32_t mapa_cluster(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
NGU_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on an mbarrier of (possibly) the peer CTA of the pair
NGU_DEVINL void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// TMEM -> registers, 32 lanes x 32 columns of 32 bit (thread i of the warp gets lane base+i).
NGU_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
NGU_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
NGU_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
NGU_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM (32 lanes x N columns)
NGU_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15])
      : "memory");
}
NGU_DEVINL void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
NGU_DEVINL void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
NGU_DEVINL void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
NGU_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------
// UMMA descriptors (bit layouts: PTX ISA "tcgen05 matrix / instruction descriptor")
// ----------------------------------------------------------------------------------------
// Instruction descriptor, kind::f16, BF16 x BF16 -> FP32.
//   [4,6) D format (1 = f32) | [7,10) A format (1 = bf16) | [10,13) B format (1 = bf16)
//   [15] A major (0 = K) | [16] B major (0 = K) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0,
                                                       int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) |
         (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// Shared-memory matrix descriptor, 128-byte swizzle.
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell) | [61,64) layout type (2 = SWIZZLE_128B)
// K-major tiles (rows of 64 bf16 = 128 B, 8-row swizzle atoms of 1024 B): SBO = 1024, LBO unused.
// MN-major tiles (each K index is one 128-B row of 64 MN elements): SBO = byte step between
//   groups of 8 K rows, LBO = byte step between 64-element MN chunks.
NGU_DEVINL uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// ----------------------------------------------------------------------------------------
// math
// ----------------------------------------------------------------------------------------
// Phi(x) = 0.5*erfc(-x/sqrt2) evaluated as 2^p(|x|) with a degree-7 polynomial fitted to
// log2(0.5*erfc(a/sqrt2)) on a in [0,6] (|Phi err| <= 2e-6, |gelu err| <= 7e-7): one MUFU.EX2.
NGU_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
NGU_DEVINL float norm_cdf(float x) {
  const float a = fminf(fabsf(x), 6.0f);
  float p = fmaf(-1.889626219e-06f, a, 6.268139987e-05f);
  p = fmaf(p, a, -9.388679173e-04f);
  p = fmaf(p, a, 8.539461531e-03f);
  p = fmaf(p, a, -5.402068794e-02f);
  p = fmaf(p, a, -4.584097862e-01f);
  p = fmaf(p, a, -1.151269197e+00f);
  p = fmaf(p, a, -9.999943376e-01f);
  const float e = ex2_approx(p);  // = Phi(-|x|)
  return x >= 0.f ? 1.0f - e : e;
}
NGU_DEVINL float gelu_erf(float x) { return x * norm_cdf(x); }
// d/dx gelu(x) = Phi(x) + x * phi(x) with ONE MUFU: Phi(-|x|) = phi(x) * m(|x|), m = Mills ratio fitted by a
// degree-8 polynomial on [0, 6.5] weighted by phi (|gelu' err| <= 5e-5, below bf16 resolution).
NGU_DEVINL float gelu_erf_grad(float x) {
  const float a = fminf(fabsf(x), 6.5f);
  const float pdf = 0.3989422804014327f * ex2_approx(-0.7213475204444817f * a * a);
  float m = fmaf(1.828833229e-05f, a, -4.542384704e-04f);
  m = fmaf(m, a, 4.796606954e-03f);
  m = fmaf(m, a, -2.867788263e-02f);
  m = fmaf(m, a, 1.102813110e-01f);
  m = fmaf(m, a, -2.983200252e-01f);
  m = fmaf(m, a, 6.123299003e-01f);
  m = fmaf(m, a, -9.974651933e-01f);
  m = fmaf(m, a, 1.253203034e+00f);
  const float tail = pdf * m;                    // Phi(-|x|)
  const float cdf = x >= 0.f ? 1.0f - tail : tail;
  return fmaf(x, pdf, cdf);
}
// gelu(x) and gelu'(x) together from ONE MUFU: Phi(x) = 0.5 + 0.5 tanh(x (a0 + a1 x^2 + a2 x^4)) with (a0,a1,a2)
// fitted to the exact erf CDF (|Phi err| <= 2e-5 + tanh.approx error, |gelu err| <= 5e-5, |gelu' err| <= 1.3e-4:
// below bf16 resolution).  The GEMM epilogue emits the activation and its derivative so the backward epilogue
// is a plain multiply.
NGU_DEVINL float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2): one issue slot per two lanes' worth of work; the FMA pipe rate is unchanged
// (128 lanes/clk/SM), so these pay where a loop is issue-bound, not FMA-bound.
NGU_DEVINL float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)), "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
NGU_DEVINL float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
NGU_DEVINL float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
constexpr float kGa0 = 7.97704294e-01f, kGa1 = 3.68194288e-02f, kGa2 = -3.20606757e-04f;
// The odd polynomial u(x) = x * P(x^2) turns over at x^2 = 133.7; tanh(u) is already exactly +-1 in fp32 for |x| >= 7.3, so x^2
// is clamped to 100 inside P (and P') and u stays monotone for every finite x.
constexpr float kGeluS2Max = 100.0f;
NGU_DEVINL float gelu_fast(float x) {
  const float x2 = fminf(x * x, kGeluS2Max);
  const float t = tanh_approx(x * fmaf(fmaf(kGa2, x2, kGa1), x2, kGa0));
  return x * fmaf(0.5f, t, 0.5f);
}
NGU_DEVINL void gelu_and_grad(float x, float& y, float& dy) {
  const float x2 = fminf(x * x, kGeluS2Max);
  const float t = tanh_approx(x * fmaf(fmaf(kGa2, x2, kGa1), x2, kGa0));
  const float cdf = fmaf(0.5f, t, 0.5f);
  const float hdu = fmaf(fmaf(2.5f * kGa2, x2, 1.5f * kGa1), x2, 0.5f * kGa0);   // u'(x) / 2
  y = x * cdf;
  dy = fmaf(x * fmaf(-t, t, 1.0f), hdu, cdf);
}
// The same on a register pair: 12 packed FMA-pipe instructions (+ 2 MUFU, 2 FMNMX) for two elements.
template <bool GRAD>
NGU_DEVINL void gelu_pair(float2 x, float2& y, float2& dy) {
  float2 s = fmul2(x, x);
  s.x = fminf(s.x, kGeluS2Max);
  s.y = fminf(s.y, kGeluS2Max);
  const float2 pl = ffma2(ffma2(make_float2(kGa2, kGa2), s, make_float2(kGa1, kGa1)), s, make_float2(kGa0, kGa0));
  const float2 u = fmul2(x, pl);
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 cdf = ffma2(make_float2(0.5f, 0.5f), t, make_float2(0.5f, 0.5f));
  y = fmul2(x, cdf);
  if (GRAD) {
    // dy = cdf + (1 - t^2) * x * u'/2, written as (t^2 - 1) * (x * -u'/2) + cdf: no negated operands for the packed forms
    const float2 nh = ffma2(ffma2(make_float2(-2.5f * kGa2, -2.5f * kGa2), s, make_float2(-1.5f * kGa1, -1.5f * kGa1)), s,
                            make_float2(-0.5f * kGa0, -0.5f * kGa0));
    const float2 q = ffma2(t, t, make_float2(-1.0f, -1.0f));
    dy = ffma2(q, fmul2(x, nh), cdf);
  }
}
NGU_DEVINL float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
NGU_DEVINL float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
NGU_DEVINL float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
NGU_DEVINL float quick_gelu(float x) { return x * sigmoid_fast(1.702f * x); }
NGU_DEVINL void quick_gelu_and_grad(float x, float& y, float& dy) {
  const float s = sigmoid_fast(1.702f * x);
  y = x * s;
  dy = s * (1.0f + 1.702f * x * (1.0f - s));
}
NGU_DEVINL float quick_gelu_grad(float x) {
  const float s = sigmoid_fast(1.702f * x);
  return s * (1.0f + 1.702f * x * (1.0f - s));
}

NGU_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
NGU_DEVINL float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 128-bit shared-memory load the compiler will neither hoist nor keep live across loop iterations
NGU_DEVINL float4 lds128_volatile(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

// dtype-generic scalar load/store used by the SIMT (fp32 check mode) kernels
template <typename T> NGU_DEVINL float to_f32(T v);
template <> NGU_DEVINL float to_f32<float>(float v) { return v; }
template <> NGU_DEVINL float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> NGU_DEVINL T from_f32(float v);
template <> NGU_DEVINL float from_f32<float>(float v) { return v; }
template <> NGU_DEVINL bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// ----------------------------------------------------------------------------------------
// host side: tensor map encode through the driver entry point (no link against libcuda)
// ----------------------------------------------------------------------------------------
// 2-D row-major tensor [rows, cols] of bf16 with row pitch `ld` elements; box = [box_rows, box_cols].
// swizzle: 0 none, 1 = SWIZZLE_128B (box_cols * 2 <= 128 bytes), 2 = SWIZZLE_64B (box_cols * 2 <= 64 bytes).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols, int swizzle);
// 3-D row-major tensor [batch, rows, cols] of bf16 (row pitch ld, batch pitch ldb elements), box = [1, box_rows, box_cols];
// rows past `rows` inside a batch element are clipped on store / zero-filled on load.
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t ldb,
                      uint32_t box_rows, uint32_t box_cols, int swizzle);
int sm_count();
bool pdl_enabled();               // NGU_PDL=1 (default off: measured slower on the benchmark step)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
const uint64_t* seed_counter();   // device pointer registered with ngu_set_seed_counter, or nullptr

}  // namespace ngu

// ----------------------------------------------------------------------------------------
// 16-byte vector access, generic over the activation dtype (bf16 product path / fp32 check mode)
// ----------------------------------------------------------------------------------------
namespace ngu {
template <typename T> struct Vec;
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  static NGU_DEVINL void load(const bf16* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  static NGU_DEVINL void unpack(const uint4& u, float (&f)[8]) {
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  static NGU_DEVINL void store(bf16* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec<float> {
  static constexpr int N = 4;
  static NGU_DEVINL void load(const float* p, float (&f)[4]) {
    const float4 u = *reinterpret_cast<const float4*>(p);
    f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w;
  }
  static NGU_DEVINL void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  static NGU_DEVINL void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
// exact (libm) variants for the fp32 check mode, fast variants for bf16
template <typename T> NGU_DEVINL float gelu_t(float x);
template <> NGU_DEVINL float gelu_t<bf16>(float x) { return gelu_erf(x); }
template <> NGU_DEVINL float gelu_t<float>(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }
template <typename T> NGU_DEVINL float gelu_grad_t(float x);
template <> NGU_DEVINL float gelu_grad_t<bf16>(float x) { return gelu_erf_grad(x); }
template <> NGU_DEVINL float gelu_grad_t<float>(float x) {
  return 0.5f * (1.f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
}  // namespace ngu

// ----------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG for dropout masks (regenerated in backward from seed + element index)
// ----------------------------------------------------------------------------------------
namespace ngu {
NGU_DEVINL uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c[4] = {c0, c1, 0x9E3779B9u, 0xBB67AE85u};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c[0], c[1], c[2], c[3]);
}
// keep-mask for element `idx` of a tensor under dropout probability p: returns 1/(1-p) or 0.
// Counter-based (stateless, regenerated in backward from seed + element index): one splitmix64 finaliser of
// (seed, idx / 4) yields the 16-bit uniforms of four consecutive elements, so vectorised kernels hash once per
// four elements; every user (LoRA dropout kernel, Mona conv stage forward / backward) goes through these two functions.
NGU_DEVINL uint64_t dropout_bits(uint64_t seed, uint64_t group) {
  uint64_t z = group * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// Optional per-process device counter mixed into every dropout seed (ngu_set_seed_counter): a captured CUDA graph replays
// with the seeds baked in, the counter (advanced on device once per micro-step) makes each replay draw fresh masks while
// forward and backward of the same step still agree.
NGU_DEVINL uint64_t mix_seed(uint64_t seed, const uint64_t* ctr) { return ctr ? seed + (*ctr) * 0xD1B54A32D192ED03ull : seed; }
NGU_DEVINL uint32_t dropout_threshold(float p) { return uint32_t(p * 65536.0f); }
NGU_DEVINL float dropout_pick(uint64_t bits, int lane4, uint32_t thr, float keep_scale) {
  return (uint32_t(bits >> (16 * lane4)) & 0xFFFFu) >= thr ? keep_scale : 0.0f;
}
NGU_DEVINL float dropout_scale(uint64_t seed, uint64_t idx, float p) {
  return dropout_pick(dropout_bits(seed, idx >> 2), int(idx & 3), dropout_threshold(p), 1.0f / (1.0f - p));
}
}  // namespace ngu
