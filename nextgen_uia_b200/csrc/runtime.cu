// Host-side runtime glue: error strings, launch accounting, tensor-map encoding, device queries.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include "common.cuh"
#include "kernels.h"

namespace ngu {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NGU_OK;
  set_last_error("%s: %s", what, cudaGetErrorString(e));
  return NGU_ERR_CUDA;
}
int check_launch(const char* what) {
  count_launch(1);
  return cuda_status(cudaGetLastError(), what);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(sym);
    else cudaGetLastError();
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols, int swizzle) {
  // swizzle: 0 none, 1 (true) 128-byte, 2 64-byte
  const bool swizzle128 = swizzle == 1;
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver / GPU?)"); return NGU_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((ld * 2) & 15u)) {
    set_last_error("tensor map: base %p / pitch %llu B not 16-byte aligned", base, (unsigned long long)(ld * 2));
    return NGU_ERR_ALIGN;
  }
  if (swizzle == 2 && box_cols * 2 > 64) { set_last_error("tensor map: 64B-swizzled box wider than 64 B"); return NGU_ERR_ARG; }
  if (swizzle128 && box_cols * 2 > 128) { set_last_error("tensor map: swizzled box wider than 128 B"); return NGU_ERR_ARG; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", int(r),
                   (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return NGU_ERR_CUDA;
  }
  return NGU_OK;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t ldb,
                      uint32_t box_rows, uint32_t box_cols, int swizzle) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver / GPU?)"); return NGU_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((ld * 2) & 15u) || ((ldb * 2) & 15u)) {
    set_last_error("tensor map: base %p / pitches %llu, %llu B not 16-byte aligned", base, (unsigned long long)(ld * 2),
                   (unsigned long long)(ldb * 2));
    return NGU_ERR_ALIGN;
  }
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {ld * 2, ldb * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(3d) failed (%d) batch=%llu rows=%llu cols=%llu", int(r), (unsigned long long)batch,
                   (unsigned long long)rows, (unsigned long long)cols);
    return NGU_ERR_CUDA;
  }
  return NGU_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NGU_PDL"); v = (e != nullptr && atoi(e) != 0) ? 1 : 0; }   // measured: PDL costs ~3 % on this step (r2 profiles), default off
  return v == 1;
}

static const uint64_t* g_seed_ctr = nullptr;
const uint64_t* seed_counter() { return g_seed_ctr; }
void set_seed_counter(const uint64_t* p) { g_seed_ctr = p; }

const char* last_error();
int64_t launch_count();

}  // namespace ngu
