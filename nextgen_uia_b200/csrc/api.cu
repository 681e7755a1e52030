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
  if (d->aux_mode == NGU_AUX_MONA_DX && d->dtype != NGU_BF16) { set_last_error("ngu_gemm: NGU_AUX_MONA_DX is a bf16 (tcgen05) epilogue"); return NGU_ERR_DTYPE; }
  if (d->dtype == NGU_BF16) return gemm_tc(*d, s);
  if (d->dtype == NGU_F32) return gemm_simt(*d, s);
  if (d->dtype == 100 + NGU_BF16) { ngu_gemm_desc t = *d; t.dtype = NGU_BF16; return gemm_simt(t, s); }  // test-only: bf16 on CUDA cores
  set_last_error("ngu_gemm: unknown dtype %d", d->dtype);
  return NGU_ERR_DTYPE;
}


#define NGU_STREAM reinterpret_cast<cudaStream_t>(stream)
#define NGU_NONNULL(d, name) if (!(d)) { set_last_error(name ": null descriptor"); return NGU_ERR_ARG; }

int ngu_ln_fwd(const ngu_ln_desc* d, void* stream) { NGU_NONNULL(d, "ngu_ln_fwd"); return ln_fwd(*d, NGU_STREAM); }
int ngu_ln_bwd(const ngu_ln_bwd_desc* d, void* stream) { NGU_NONNULL(d, "ngu_ln_bwd"); return ln_bwd(*d, NGU_STREAM); }
int ngu_mona_pre_bwd(const ngu_mona_pre_bwd_desc* d, void* stream) { NGU_NONNULL(d, "ngu_mona_pre_bwd"); return mona_pre_bwd(*d, NGU_STREAM); }
int ngu_mona_conv_fwd(const ngu_mona_conv_desc* d, void* stream) { NGU_NONNULL(d, "ngu_mona_conv_fwd"); return mona_conv_fwd(*d, NGU_STREAM); }
int ngu_mona_conv_bwd(const ngu_mona_conv_desc* d, void* stream) { NGU_NONNULL(d, "ngu_mona_conv_bwd"); return mona_conv_bwd(*d, NGU_STREAM); }

int64_t ngu_mona_ws_floats(int D) { return mona_ws_floats(D); }
int ngu_mona_prep(const ngu_mona_prep_item* items, int n, int D, void* stream) { return mona_prep(items, n, D, NGU_STREAM); }
int ngu_mona_fwd_stage(const ngu_mona_stage_desc* d, void* stream) { NGU_NONNULL(d, "ngu_mona_fwd_stage"); return mona_fwd_stage(*d, NGU_STREAM); }
int ngu_mona_bwd_stage(const ngu_mona_stage_desc* d, void* stream) { NGU_NONNULL(d, "ngu_mona_bwd_stage"); return mona_bwd_stage(*d, NGU_STREAM); }
int ngu_mona_finish(const ngu_mona_params* p, const ngu_mona_grads* g, const float* ws, int D, void* stream) {
  if (!p || !g) { set_last_error("ngu_mona_finish: null descriptor"); return NGU_ERR_ARG; }
  return mona_finish(*p, *g, ws, D, NGU_STREAM);
}

int ngu_attn_fwd(const ngu_attn_desc* d, void* stream) {
  NGU_NONNULL(d, "ngu_attn_fwd");
  if (int rc = attn_validate(*d, "ngu_attn_fwd", false)) return rc;
  if ((d->impl == 0 || d->impl == 2) && attn_tc_supported(*d, false)) return attn_fwd_tc(*d, NGU_STREAM);
  if (d->impl == 0 && attn_long_supported(*d, false)) return attn_fwd_long(*d, NGU_STREAM);
  return attn_fwd_simt(*d, NGU_STREAM);
}
int ngu_attn_bwd(const ngu_attn_desc* d, void* stream) {
  NGU_NONNULL(d, "ngu_attn_bwd");
  if (int rc = attn_validate(*d, "ngu_attn_bwd", true)) return rc;
  if (d->impl == 0 && attn_tc_supported(*d, true)) return attn_bwd_tc(*d, NGU_STREAM);
  if (d->impl == 0 && attn_long_supported(*d, true)) return attn_bwd_long(*d, NGU_STREAM);
  return attn_bwd_simt(*d, NGU_STREAM);
}

int ngu_infonce_normalize(const void* x, float* xhat, float* norm, int B, int E, int dtype, void* stream) {
  return infonce_normalize(x, xhat, norm, B, E, dtype, NGU_STREAM);
}
int ngu_infonce_core(const ngu_infonce_desc* d, void* stream) { NGU_NONNULL(d, "ngu_infonce_core"); return infonce_core(*d, NGU_STREAM); }
int ngu_infonce_normalize_bwd(const float* dxhat, const float* xhat, const float* norm, const float* gscale, void* dx, int B,
                              int E, int dtype, void* stream) {
  return infonce_normalize_bwd(dxhat, xhat, norm, gscale, dx, B, E, dtype, NGU_STREAM);
}

int ngu_wgrad(const void* X, int ldx, const void* Y, int ldy, float* D, int ldd, int T, int Mo, int No, int dtype, int impl,
              void* stream) {
  if (impl == 0 && T > 0 && wgrad_tc_supported(ldx, ldy, ldd, Mo, No, dtype, X, Y, D)) return wgrad_tc(X, ldx, Y, ldy, D, ldd, T, Mo, No, NGU_STREAM);
  return wgrad_simt(X, ldx, Y, ldy, D, ldd, T, Mo, No, dtype, NGU_STREAM);
}
int ngu_colsum(const void* X, int ldx, float* out, int T, int C, int dtype, void* stream) {
  return colsum(X, ldx, out, T, C, dtype, NGU_STREAM);
}
int ngu_dropout(const void* x, void* out, int64_t n, float p, uint64_t seed, int accumulate, int dtype, void* stream) {
  return dropout(x, out, size_t(n), p, seed, accumulate, dtype, NGU_STREAM);
}
int ngu_sqnorm(const float* x, int64_t n, float* out, void* stream) { return sqnorm(x, size_t(n), out, NGU_STREAM); }
int ngu_guard_tick(int64_t* state, const float* loss, const float* gsq, int mode, void* stream) { return guard_tick(state, loss, gsq, mode, NGU_STREAM); }
int ngu_kv_len(const int64_t* ids, int64_t pad_id, int* kv_len_out, int* flag, int B, int S, void* stream) { return kv_len(ids, pad_id, kv_len_out, flag, B, S, NGU_STREAM); }
int ngu_zero_shot_prototypes(const void* text_feat, const int* class_of_prompt, float* proto, int P, int E, int C, int dtype, void* stream) {
  return zero_shot_prototypes(text_feat, class_of_prompt, proto, P, E, C, dtype, NGU_STREAM);
}
int ngu_zero_shot_score(const void* image_feat, const float* proto, float* logits, int* pred, int B, int E, int C, float scale, int dtype, void* stream) {
  return zero_shot_score(image_feat, proto, logits, pred, B, E, C, scale, dtype, NGU_STREAM);
}
int ngu_set_seed_counter(const void* counter) { set_seed_counter(reinterpret_cast<const uint64_t*>(counter)); return NGU_OK; }
int ngu_adamw_step(const ngu_adamw_desc* d, void* stream) { NGU_NONNULL(d, "ngu_adamw_step"); return adamw_step(*d, NGU_STREAM); }
int ngu_patchify(const float* img, void* out, int B, int R, int P, int dtype, void* stream) {
  return patchify(img, out, B, R, P, dtype, NGU_STREAM);
}
int ngu_assemble_tokens(const void* patch, const float* cls, const float* pos, void* out, int B, int np, int D, int dtype,
                        void* stream) {
  return assemble_tokens(patch, cls, pos, out, B, np, D, dtype, NGU_STREAM);
}
int ngu_embed_tokens(const int64_t* ids, const float* word, const float* pos, const float* type0, void* out, int B, int S,
                     int D, int vocab, int dtype, void* stream) {
  return embed_tokens(ids, word, pos, type0, out, B, S, D, vocab, dtype, NGU_STREAM);
}
int ngu_cast_f32(const float* in, void* out, int rows, int cols, int transpose, float scale, int dtype, void* stream) {
  return cast_f32(in, out, rows, cols, transpose, scale, dtype, NGU_STREAM);
}
int ngu_cast_f32_batch(const ngu_cast_item* items, int n, int dtype, void* stream) {
  return cast_f32_batch(items, n, dtype, NGU_STREAM);
}

}  // extern "C"
