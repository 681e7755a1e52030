/*
 * ngu_b200.h — C ABI of libngu_b200.so: the sm_100a (B200) kernels behind the NextGen-UIA adapter
 * fine-tuning hot path (ViT-B/16 block + Mona + LoRA, forward and backward, + InfoNCE).
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; row-major; 16-byte aligned.
 *   - the library never allocates or frees: outputs and workspaces are caller-provided
 *     (PyTorch owns all memory on the Python side).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises,
 *     so every entry point is CUDA-graph capturable and re-entrant.
 *   - return value: 0 = OK, negative = error (NGU_ERR_*); ngu_last_error() returns a thread-local
 *     message.  There is no CPU fallback: without a usable sm_100 device the calls fail.
 *   - dtype: NGU_BF16 is the product path (tcgen05 tensor-core GEMMs, bf16 activations, fp32
 *     accumulation/statistics); NGU_F32 is the "fp32 check mode" of the parity contract (same
 *     kernels' math in fp32 on CUDA cores), used by tests at small sizes.
 *
 * Each entry point cites the reference code (jinggqu/NextGen-UIA) whose device work it replaces.
 */
#ifndef NGU_B200_H_
#define NGU_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGU_VERSION 100

/* error codes */
#define NGU_OK 0
#define NGU_ERR_SHAPE (-1)
#define NGU_ERR_ALIGN (-2)
#define NGU_ERR_DTYPE (-3)
#define NGU_ERR_CUDA (-4)
#define NGU_ERR_ARG (-5)

/* dtypes */
#define NGU_BF16 0
#define NGU_F32 1

/* activation in a GEMM epilogue */
#define NGU_ACT_NONE 0
#define NGU_ACT_GELU 1      /* exact erf GELU: timm Mlp (pinned dep timm 1.0.20), F.gelu in src/adapters/mona.py:107 */
#define NGU_ACT_QUICKGELU 2 /* x*sigmoid(1.702x): src/third_party/openai_clip/model.py:172-174 */

/* role of the auxiliary [M,N] operand in a GEMM epilogue */
#define NGU_AUX_NONE 0
#define NGU_AUX_RESIDUAL 1 /* C = act(acc + bias) + aux            (x + attn(..), x + mlp(..)) */
#define NGU_AUX_DACT 2     /* C = (acc + bias) * act'(aux)          (backward through fc1's activation) */

int ngu_version(void);
const char* ngu_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
int64_t ngu_launch_count(void);
/* 0 if device 0..n is an sm_100 part and the driver can encode tensor maps, else NGU_ERR_CUDA */
int ngu_selftest_device(void);

/*
 * Dense projection  C[M,N] = epi( alpha * (A[M,K] · B[N,K]^T  +  A2[M,K2] · B2[N,K2]^T) ).
 * Replaces nn.Linear / addmm on the hot path: timm Block attn.qkv, attn.proj, mlp.fc1, mlp.fc2
 * (touched at src/adapters/lora.py:284-313, src/adapters/mona.py:620-630), the CLIP
 * ResidualAttentionBlock (src/third_party/openai_clip/model.py:177-202) and LinearLoRA.forward
 * (src/adapters/lora.py:78-90; the low-rank pair A2 = s·drop(x)·A^T, B2 = lora_B rides as extra K
 * blocks instead of the reference's dense B@A re-materialisation).  dgrad for frozen weights is the
 * same call with the host-kept transposed weight copy.
 */
typedef struct ngu_gemm_desc {
  const void* A;  int lda;   /* [M,K]  */
  const void* B;  int ldb;   /* [N,K]  */
  void* C;        int ldc;   /* [M,N]  */
  const void* A2; int lda2;  /* [M,K2] or NULL */
  const void* B2; int ldb2;  /* [N,K2] or NULL */
  const float* bias;         /* [N] fp32 or NULL */
  const void* aux; int ldaux;/* [M,N] or NULL (see NGU_AUX_*) */
  void* Pre;      int ldpre; /* [M,N] pre-activation output when save_pre != 0 */
  int M, N, K, K2;
  int act, aux_mode, save_pre;
  float alpha;
  int dtype;                 /* NGU_BF16 (tcgen05 path) or NGU_F32 (check mode) */
  int block_n;               /* 0 = auto; 64/128/256 force the N tile (tuning/tests) */
} ngu_gemm_desc;
int ngu_gemm(const ngu_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NGU_B200_H_ */
