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

int ln_fwd(const ngu_ln_desc& d, cudaStream_t s);
int ln_bwd(const ngu_ln_bwd_desc& d, cudaStream_t s);
int mona_pre_bwd(const ngu_mona_pre_bwd_desc& d, cudaStream_t s);
int mona_conv_fwd(const ngu_mona_conv_desc& d, cudaStream_t s);
int mona_conv_bwd(const ngu_mona_conv_desc& d, cudaStream_t s);
int64_t mona_ws_floats(int D);
int mona_prep(const ngu_mona_prep_item* items, int n, int D, cudaStream_t s);
int mona_fwd_stage(const ngu_mona_stage_desc& d, cudaStream_t s);
int mona_bwd_stage(const ngu_mona_stage_desc& d, cudaStream_t s);
int mona_finish(const ngu_mona_params& p, const ngu_mona_grads& g, const float* ws, int D, cudaStream_t s);
int attn_validate(const ngu_attn_desc& d, const char* what, bool bwd);
int attn_fwd_simt(const ngu_attn_desc& d, cudaStream_t s);
int attn_bwd_simt(const ngu_attn_desc& d, cudaStream_t s);
bool attn_tc_supported(const ngu_attn_desc& d, bool bwd);
int attn_fwd_tc(const ngu_attn_desc& d, cudaStream_t s);
int attn_bwd_tc(const ngu_attn_desc& d, cudaStream_t s);
bool attn_long_supported(const ngu_attn_desc& d, bool bwd);
int attn_fwd_long(const ngu_attn_desc& d, cudaStream_t s);
int attn_bwd_long(const ngu_attn_desc& d, cudaStream_t s);
int infonce_normalize(const void* x, float* xhat, float* norm, int B, int E, int dtype, cudaStream_t s);
int infonce_core(const ngu_infonce_desc& d, cudaStream_t s);
int infonce_normalize_bwd(const float* dxhat, const float* xhat, const float* norm, const float* gscale, void* dx, int B, int E, int dtype, cudaStream_t s);
int patchify(const float* img, void* out, int B, int R, int P, int dtype, cudaStream_t s);
int assemble_tokens(const void* patch, const float* cls, const float* pos, void* out, int B, int np, int D, int dtype, cudaStream_t s);
int embed_tokens(const int64_t* ids, const float* word, const float* pos, const float* type0, void* out, int B, int S, int D, int vocab, int dtype, cudaStream_t s);
int cast_f32(const float* in, void* out, int rows, int cols, int transpose, float scale, int dtype, cudaStream_t s);
int cast_f32_batch(const ngu_cast_item* items, int n, int dtype, cudaStream_t s);
int wgrad_simt(const void* X, int ldx, const void* Y, int ldy, float* D, int ldd, int Tn, int Mo, int No, int dtype, cudaStream_t s);
int dropout(const void* x, void* out, size_t n, float p, uint64_t seed, int accumulate, int dtype, cudaStream_t s);
bool wgrad_tc_supported(int ldx, int ldy, int ldd, int Mo, int No, int dtype, const void* X, const void* Y, const float* D);
int wgrad_tc(const void* X, int ldx, const void* Y, int ldy, float* D, int ldd, int T, int Mo, int No, cudaStream_t s);
int colsum(const void* X, int ldx, float* out, int Tn, int Cn, int dtype, cudaStream_t s);

int sqnorm(const float* x, size_t n, float* out, cudaStream_t s);
int guard_tick(int64_t* state, const float* loss, const float* gsq, int mode, cudaStream_t s);
int kv_len(const int64_t* ids, int64_t pad, int* out, int* flag, int B, int S, cudaStream_t s);
void set_seed_counter(const uint64_t* p);
int zero_shot_prototypes(const void* tf, const int* cls, float* proto, int P, int E, int C, int dtype, cudaStream_t s);
int zero_shot_score(const void* f, const float* proto, float* logits, int* pred, int B, int E, int C, float scale, int dtype, cudaStream_t s);
int adamw_step(const ngu_adamw_desc& d, cudaStream_t s);

void count_launch(int n = 1);

}  // namespace ngu
