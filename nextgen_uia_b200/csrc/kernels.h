// Internal C++ declarations shared by the kernel translation units and api.cu.
#pragma once
#include <cuda_runtime.h>
#include "../../include/ngu_b200.h"

namespace ngu {

typedef ngu_gemm_desc GemmArgs;

// tcgen05 / TMA GEMM (bf16) — gemm_tc.cu
int gemm_tc(const GemmArgs& a, cudaStream_t stream);
// CUDA-core GEMM for the fp32 check mode (and a bf16 instantiation used only by tests) — gemm_simt.cu
int gemm_simt(const GemmArgs& a, cudaStream_t stream);

void count_launch(int n = 1);

}  // namespace ngu
