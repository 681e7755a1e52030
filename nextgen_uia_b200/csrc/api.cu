// extern "C" entry points of libngu_b200.so (declared in include/ngu_b200.h).
#include "common.cuh"
#include "kernels.h"

namespace ngu {
const char* last_error();
int64_t launch_count();
}  // namespace ngu

using namespace ngu;

extern "C" {

int ngu_version(void) { return NGU_VERSION; }
const char* ngu_last_error(void) { return last_error(); }
int64_t ngu_launch_count(void) { return launch_count(); }

int ngu_selftest_device(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); set_last_error("no CUDA device visible"); return NGU_ERR_CUDA; }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) { set_last_error("device %d is sm_%d0, this library is sm_100a only", dev, major); return NGU_ERR_CUDA; }
  return NGU_OK;
}

int ngu_gemm(const ngu_gemm_desc* d, void* stream) {
  if (!d) { set_last_error("ngu_gemm: null descriptor"); return NGU_ERR_ARG; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (d->dtype == NGU_BF16) return gemm_tc(*d, s);
  if (d->dtype == NGU_F32) return gemm_simt(*d, s);
  if (d->dtype == 100 + NGU_BF16) { ngu_gemm_desc t = *d; t.dtype = NGU_BF16; return gemm_simt(t, s); }  // test-only: bf16 on CUDA cores
  set_last_error("ngu_gemm: unknown dtype %d", d->dtype);
  return NGU_ERR_DTYPE;
}

}  // extern "C"
