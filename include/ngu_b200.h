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
#define NGU_AUX_DACT 2     /* C = (acc + bias) * aux,  aux = act'(pre) saved by the forward (backward through fc1's activation) */
#define NGU_AUX_DACT_U8 4  /* NGU_AUX_DACT with aux = the ONE-BYTE derivative [M, ldaux] a forward with save_pre == 2 wrote (bf16 path) */
#define NGU_AUX_MONA_DX 3  /* C = acc + aux + rowab[r].beta * aux2 + rowab[r].alpha: backward of the Mona input mix + residual
                              (src/adapters/mona.py:124-125,150) with A = [dh | dh*rstd], B = [W1*gammax ; W1*w*gamma]^T, aux = dy,
                              aux2 = x and the LayerNorm-backward row terms folded into two per-row scalars (ngu_mona_bwd_stage) */

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
  void* Pre;      int ldpre; /* [M,N] when save_pre != 0: act'(acc + bias) (the derivative backward needs); acc + bias if act == NONE.
                                save_pre == 2 (bf16 path, act != NONE): one byte per element, q = round(d * 170 + 43) */
  int M, N, K, K2;
  int act, aux_mode, save_pre;
  float alpha;
  int dtype;                 /* NGU_BF16 (tcgen05 path) or NGU_F32 (check mode) */
  int block_n;               /* 0 = auto; 64/128/256 force the N tile (tuning/tests) */
  const void* aux2; int ldaux2; /* [M,N] second elementwise operand (NGU_AUX_MONA_DX) or NULL */
  const float* rowab;        /* [M,2] fp32 per-row (alpha, beta) of NGU_AUX_MONA_DX or NULL */
  int c_dtype;               /* 0 = C has the operand dtype; NGU_F32 with dtype NGU_BF16 = fp32 C from bf16 operands (plain alpha * acc
                                epilogue): InfoNCE logits and feature gradients, src/losses/losses.py:34 */
} ngu_gemm_desc;
int ngu_gemm(const ngu_gemm_desc* d, void* stream);

/*
 * LayerNorm forward (+ optional Mona pre-scale).  One call replaces nn.LayerNorm (timm Block norm1 /
 * norm2 / trunk.norm, eps 1e-6, pinned dep; CLIP LayerNorm src/third_party/openai_clip/model.py:163-169)
 * and, with gamma/gammax set, the Mona input mix  src/adapters/mona.py:125
 *     y = (xhat * w + b) * gamma + x * gammax.
 * Rows may be strided (ldx / ldy elements) so the CLS rows of [B,N,D] can be normalised in place.
 * mean / rstd (fp32 [M]) are saved for backward when non-NULL.
 */
typedef struct ngu_ln_desc {
  const void* x; int64_t ldx;
  void* y;       int64_t ldy;
  const float* w; const float* b;          /* [D] */
  const float* gamma; const float* gammax; /* [D] or both NULL */
  float* mean; float* rstd;                /* [M] or NULL */
  int M, D; float eps; int dtype;
} ngu_ln_desc;
int ngu_ln_fwd(const ngu_ln_desc* d, void* stream);

/* LayerNorm backward with frozen affine: dx = LNbwd(g; x, mean, rstd, w) (+ dres).  The base model's
 * norms are frozen in Mona/LoRA fine-tuning (src/models/biomedclip/finetune.py:166-175) so no dw/db. */
typedef struct ngu_ln_bwd_desc {
  const void* g; int64_t ldg;
  const void* x; int64_t ldx;
  const void* dres; int64_t ldr;  /* optional residual-stream gradient added to dx */
  void* dx; int64_t lddx;
  const float* mean; const float* rstd; const float* w;
  int M, D; int dtype;
} ngu_ln_bwd_desc;
int ngu_ln_bwd(const ngu_ln_bwd_desc* d, void* stream);

/* Backward of the Mona input mix + residual (src/adapters/mona.py:124-125,150):
 *   dx = dy + du*gammax + LNbwd(du*gamma);  dw/db/dgamma/dgammax/dycol are fp32 [D], ACCUMULATED (+=). */
typedef struct ngu_mona_pre_bwd_desc {
  const void* du; const void* dy; const void* x;
  const float* mean; const float* rstd;
  const float* w; const float* b; const float* gamma; const float* gammax;
  void* dx;
  float* dw; float* db; float* dgamma; float* dgammax; float* dycol;
  int M, D; int dtype;
} ngu_mona_pre_bwd_desc;
int ngu_mona_pre_bwd(const ngu_mona_pre_bwd_desc* d, void* stream);

/*
 * Mona bottleneck stage between project1 and project2: merged depthwise 3x3+5x5+7x7 stencil,
 * 1x1 projector, GELU, dropout — src/adapters/mona.py:85-93 (BaselineMonaOp) and :129-147.
 * h, g, dg, dh: [B, N, C] with N = has_cls + H*W.  Weights/grads fp32 in the reference's own
 * parameter shapes (conv{1,2,3}.weight [C,1,k,k], projector.weight [C,C,1,1]); grads ACCUMULATE.
 * db1 receives the column sum of dh (= d project1.bias).
 */
typedef struct ngu_mona_conv_weights {
  const float* k3; const float* b3; const float* k5; const float* b5; const float* k7; const float* b7;
  const float* P; const float* bp;
  /* variants (NULL = baseline behaviour):
   *   freq   [C]        FreqEnhancedMonaOp.freq_filter (mona.py:279): rfft2 -> x f_c -> irfft2 == per-channel scale of the conv input
   *   ne_*              NoiseAwareMonaOp.noise_estimator (mona.py:170-176): GAP -> 1x1 (C -> C/4) -> ReLU -> 1x1 (C/4 -> 3) -> softmax
   *                     = per-image weights of the three depthwise branches (replacing the fixed 1/3) */
  const float* freq;
  const float* ne_w1; const float* ne_b1; const float* ne_w2; const float* ne_b2;
} ngu_mona_conv_weights;
typedef struct ngu_mona_conv_grads {
  float* dk3; float* db3; float* dk5; float* db5; float* dk7; float* db7; float* dP; float* dbp; float* db1;
  float* dfreq; float* dne_w1; float* dne_b1; float* dne_w2; float* dne_b2;
} ngu_mona_conv_grads;
typedef struct ngu_mona_conv_desc {
  const void* h; void* g;          /* forward: h -> g */
  const void* dg; void* dh;        /* backward: (h, dg) -> dh + grads */
  ngu_mona_conv_weights w;
  ngu_mona_conv_grads gr;
  int B, N, H, W, C, has_cls;
  float drop_p; uint64_t seed;     /* dropout p (0 = eval) and counter-RNG seed; backward regenerates the mask */
  int dtype;
  int force_simt;                  /* tests: run the bf16 stage on CUDA cores instead of mma.sync */
} ngu_mona_conv_desc;
int ngu_mona_conv_fwd(const ngu_mona_conv_desc* d, void* stream);
int ngu_mona_conv_bwd(const ngu_mona_conv_desc* d, void* stream);

/*
 * Fused Mona adapter, bf16 product path — src/adapters/mona.py:115-151 (BaselineMona.forward) with :85-93
 * (BaselineMonaOp) / :261-296 (FreqEnhancedMonaOp) inside, batch-first [B,N,D] (BatchFirstMonaWrapper :54-67).
 * The LayerNorm mix is folded INTO the 768->64 projection so x is consumed straight from HBM by TMA:
 *     h = rstd_r * (x Wa^T) + (x Wb^T) - mean_r rstd_r ca + cb,   Wa = W1 * (ln_w*gamma),  Wb = W1 * gammax,
 *     ca = Wa 1,  cb = b1 + W1 (ln_b*gamma)
 * (row statistics are accumulated from the same shared-memory tiles the tensor core reads), and its backward is
 *     dx = dy + [dh | dh*rstd] [Wb ; Wa] + beta_r x + alpha_r        (one GEMM, NGU_AUX_MONA_DX epilogue)
 * with every parameter gradient of the input mix derived from G = x^T [dh | dh*rstd] (one token reduction).
 *
 *   ngu_mona_prep        : derived operands of n adapters from their fp32 parameters (after each optimiser update)
 *   ngu_mona_fwd_stage   : x -> h, hA (= LN part of h, saved for backward), g = dropout(gelu(stage(h))), mean, rstd
 *                          (project2 + residual is an ngu_gemm with NGU_AUX_RESIDUAL on g)
 *   ngu_mona_bwd_stage   : (h, hA, dg, mean, rstd) -> dhcat = [dh | dh*rstd], rowab = (alpha_r, beta_r); accumulates the
 *                          stage reductions into `ws` and dP / dbp directly
 *   ngu_mona_finish      : ws (+ G written there by ngu_wgrad with No = 128) -> every remaining parameter gradient (+=)
 * Grids up to 16x16, bottleneck 64, D % 64 == 0, baseline and freq_enhanced variants; other cases use the unfused entry
 * points above.
 */
typedef struct ngu_mona_params {          /* fp32 parameters of one adapter, the reference's own tensors */
  const float* w1; const float* b1;       /* project1 [64, D], [64]  (mona.py:106) */
  const float* w2; const float* b2;       /* project2 [D, 64], [D]   (mona.py:108) */
  const float* ln_w; const float* ln_b;   /* norm [D]                (mona.py:111) */
  const float* gamma; const float* gammax;/* [D]                     (mona.py:112-113) */
  ngu_mona_conv_weights conv;             /* adapter_conv.*          (mona.py:78-83) */
} ngu_mona_params;
typedef struct ngu_mona_derived {         /* written by ngu_mona_prep; caller-allocated */
  void* wab;      /* bf16 [128, D]: rows 0..63 Wa, rows 64..127 Wb */
  void* wcat_t;   /* bf16 [D, 128]: [k][c] = Wb[c][k], [k][64+c] = Wa[c][k] */
  void* w2;       /* bf16 [D, 64] */
  void* w2_t;     /* bf16 [64, D] */
  float* ca; float* cb;   /* [64] */
  float* kc;      /* [49*64] merged 7x7 stencil, tap-major: f_c (k3 + k5 + k7)/3 + delta */
  float* bc;      /* [64] merged stencil bias */
  void* pb;       /* bf16 [64*64] projector weight as the 128-byte-swizzled shared-memory image */
  float* bp;      /* [64] */
} ngu_mona_derived;
typedef struct ngu_mona_prep_item { ngu_mona_params p; ngu_mona_derived d; } ngu_mona_prep_item;
/* `items` is a DEVICE array of n entries (all with the same D) */
int ngu_mona_prep(const ngu_mona_prep_item* items, int n, int D, void* stream);

typedef struct ngu_mona_stage_desc {
  ngu_mona_derived d;
  const void* x;                      /* fwd: [B, N, D] bf16 */
  void* h; void* hA; void* g;         /* [B, N, 64] bf16: fwd outputs; bwd reads h, hA */
  float* mean; float* rstd;           /* [B*N] fp32: fwd outputs, bwd inputs */
  const void* dg;                     /* bwd: [B, N, 64] bf16 */
  void* dhcat;                        /* bwd: [B*N, 128] bf16 */
  float* rowab;                       /* bwd: [B*N, 2] fp32 */
  float* ws;                          /* bwd: fp32 workspace of ngu_mona_ws_floats(D) elements, zeroed by the call */
  float* dP; float* dbp;              /* bwd: projector gradients, accumulated (+=) */
  int B, N, H, W, D, has_cls;
  float eps; float drop_p; uint64_t seed;
} ngu_mona_stage_desc;
int64_t ngu_mona_ws_floats(int D);
int ngu_mona_fwd_stage(const ngu_mona_stage_desc* d, void* stream);
int ngu_mona_bwd_stage(const ngu_mona_stage_desc* d, void* stream);
typedef struct ngu_mona_grads {           /* fp32, parameter shapes, accumulated (+=) */
  float* dw1; float* db1; float* dln_w; float* dln_b; float* dgamma; float* dgammax;
  float* dk3; float* db3; float* dk5; float* db5; float* dk7; float* db7; float* dfreq;
} ngu_mona_grads;
int ngu_mona_finish(const ngu_mona_params* p, const ngu_mona_grads* g, const float* ws, int D, void* stream);

/*
 * Attention core softmax(q k^T * scale) v, forward and backward (recompute from saved LSE).
 * Replaces F.scaled_dot_product_attention in timm Attention (pinned dep) and
 * src/adapters/lora.py:188-190, and nn.MultiheadAttention's core in
 * src/third_party/openai_clip/model.py:195-197.  q/k/v/o are addressed as
 * ptr + b*bs + n*ts + head*dh (+ d), so the fused timm qkv buffer [B,N,3,H,dh], separate q/k/v
 * tensors and the sequence-first [N,B,D] layout are all expressible.  lse: fp32 [B,H,N].
 * bf16, head dim 64, packed qkv buffer: tcgen05 kernels — N <= 256 whole-sequence tiles (attention_tc.cu), 256 < N <= 1024
 * key-tiled with online softmax (attention_long.cu); other layouts / fp32 check mode: CUDA cores (attention_simt.cu).
 */
typedef struct ngu_attn_desc {
  const void* q; int64_t q_bs, q_ts;
  const void* k; int64_t k_bs, k_ts;
  const void* v; int64_t v_bs, v_ts;
  void* o;       int64_t o_bs, o_ts;
  float* lse;
  const void* d_o;            /* backward: grad of o (same strides as o) */
  void* dq; void* dk; void* dv; /* backward outputs (same strides as q / k / v) */
  int B, H, N, S, dh;
  float scale; int causal;
  int dtype;
  int impl;                   /* 0 = default for dtype (bf16: tcgen05, fp32: CUDA cores); 1 = force CUDA cores;
                                 2 = tcgen05 with the persistent two-group forward kernel (forward only; tuning) */
  const int* kv_len;          /* NULL = every key valid: [B] device ints, keys j >= kv_len[b] are masked
                                 (right-padded token batches: the HF attention_mask of BiomedCLIP's text tower,
                                 open_clip HFTextEncoder.forward `attn_mask = (x != pad_token_id)`); 1 <= kv_len[b] <= S */
  float* ws;                  /* backward, bf16 packed layout with 256 < N <= 1024 (ViT-L/14@336: 577, ViT-B/16@352: 485): fp32 workspace
                                 of B*H*N elements (delta = rowsum(dO * O)) for the tiled tcgen05 kernels; NULL selects the CUDA-core path */
} ngu_attn_desc;
int ngu_attn_fwd(const ngu_attn_desc* d, void* stream);
int ngu_attn_bwd(const ngu_attn_desc* d, void* stream);

/*
 * InfoNCE — src/losses/losses.py:23-47.  Three calls so the all-gather of the normalised features can
 * sit between them (single GPU: Bg == Bl, r0 == 0):
 *   ngu_infonce_normalize : xhat (fp32, written at the caller's row offset of the gather buffer), norms
 *   ngu_infonce_core      : loss (global mean, fp32 scalar) + d loss / d xhat for the local rows
 *   ngu_infonce_normalize_bwd : d loss / d x for the local rows (times *gscale, the upstream scalar grad)
 * ws: fp32 workspace of 2*Bg*Bg + 2*Bg elements.  Each rank evaluates the global [Bg,Bg] logits once; the loss is exact
 * for the global batch and the feature gradients of the local rows need no second collective.
 */
int ngu_infonce_normalize(const void* x, float* xhat, float* norm, int B, int E, int dtype, void* stream);
typedef struct ngu_infonce_desc {
  const float* ihat; const float* that;   /* [Bg, E] gathered, normalised */
  float* dihat; float* dthat;             /* [Bl, E] or NULL (forward only) */
  float* loss; float* ws;
  int Bg, Bl, r0, E;
  float temperature;
  /* bf16 product path (all NULL = fp32 check mode on CUDA cores): bf16 copies of the gathered normalised features and their
   * transposes, so the logit matrix and both feature-gradient contractions run on the tcgen05 GEMM (fp32 accumulate/output) */
  const void* ihat16; const void* that16;       /* [Bg, E] bf16 */
  const void* ihat16_t; const void* that16_t;   /* [E, Bg] bf16 (gradients only) */
  void* g_ws;                                   /* bf16 workspace of 2*Bl*Bg elements (gradients only) */
} ngu_infonce_desc;
int ngu_infonce_core(const ngu_infonce_desc* d, void* stream);
int ngu_infonce_normalize_bwd(const float* dxhat, const float* xhat, const float* norm, const float* gscale,
                              void* dx, int B, int E, int dtype, void* stream);

/* Weight gradient of a trainable projection: D[Mo,No] += X[T,Mo]^T · Y[T,No]  (fp32 accumulate).
 * Mona project1/project2 and LoRA A/B grads (autograd of src/adapters/mona.py:127,148, lora.py:86). */
int ngu_wgrad(const void* X, int ldx, const void* Y, int ldy, float* D, int ldd, int T, int Mo, int No,
              int dtype, int impl, void* stream);
/* out[C] += column sums of X[T,C] (bias gradients) */
int ngu_colsum(const void* X, int ldx, float* out, int T, int C, int dtype, void* stream);

/* out = (accumulate ? out : 0) + x * keep_mask(seed) / (1-p): LoRA input dropout (src/adapters/lora.py:82-83) and its
 * backward (same mask from the same seed). */
int ngu_dropout(const void* x, void* out, int64_t n, float p, uint64_t seed, int accumulate, int dtype, void* stream);

/*
 * Fused optimiser step on flat fp32 buffers: clip_grad_norm_(max_norm) + AdamW + optional skip on a non-finite loss
 * (src/models/biomedclip/finetune.py:244-255, :281-285, :296-303; torch.optim.AdamW semantics).
 *   ngu_sqnorm : *out += sum x^2   (caller zeroes *out; feed the result to ngu_adamw_step.gsq)
 */
int ngu_sqnorm(const float* x, int64_t n, float* out, void* stream);
typedef struct ngu_adamw_desc {
  float* param; float* grad; float* m; float* v;
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay;
  int step;                /* 1-based update count (bias correction) */
  float max_norm;          /* <= 0 disables clipping */
  const float* gsq;        /* device scalar: squared global grad norm (after the all-reduce), or NULL */
  const float* loss;       /* device scalar: if non-finite the update is skipped, or NULL */
  int zero_grad;           /* zero the gradient buffer after use */
  /* device-resident schedule (CUDA-graph replays cannot take a new lr / step from the host): when state != NULL the
   * update count t = state[0] drives the bias correction (step = t + 1) and, with t_max > 0, the cosine schedule
   * lr(t) = lr_min + (lr - lr_min)(1 + cos(pi t / t_max))/2 (torch CosineAnnealingLR, finetune.py:255); the update is
   * skipped while state[1] != 0 (a micro-step of this accumulation window had a non-finite loss) or *gsq is non-finite. */
  const int64_t* state;
  float lr_min; int t_max;
} ngu_adamw_desc;
int ngu_adamw_step(const ngu_adamw_desc* d, void* stream);
/* Guard / counters of the training loop on device (finetune.py:281-288): state = int64[4] {updates applied, poison flag,
 * updates skipped, micro-steps run}.  mode 0: after a micro-step's loss (poison |= !isfinite(loss)); mode 1: after the
 * optimiser launch of an update step (applied or skipped count, clear poison, micro-steps + 1); mode 2: end of a micro-step
 * that does not update (micro-steps + 1). */
int ngu_guard_tick(int64_t* state, const float* loss, const float* gsq, int mode, void* stream);
/* Optional device counter mixed into every dropout seed of this process (Mona / LoRA dropout, mona.py:147, lora.py:82):
 * a captured CUDA graph bakes the seeds in, the counter (e.g. &state[3] above) makes every replay draw fresh masks.
 * NULL disables. */
int ngu_set_seed_counter(const void* counter);
/* Key-padding lengths of a right-padded token batch (open_clip HFTextEncoder.forward: attn_mask = (x != pad_token_id)):
 * kv_len_out[b] = number of non-pad ids of row b; *flag |= 1 if some row's valid ids are not a non-empty prefix. */
int ngu_kv_len(const int64_t* ids, int64_t pad_id, int* kv_len_out, int* flag, int B, int S, void* stream);

/*
 * Zero-shot prompt-ensemble scorer — src/models/biomedclip/zero_shot.py:176-228: per class, the mean over its prompts of
 * 100 * Ihat . That_p, stacked to [B, n_classes] logits (argmax = prediction, src/utils/tools.py:211).  The mean commutes with the
 * dot product, so one prototype per class is built once (mean of the L2-normalised prompt features) and an image batch is scored
 * in one launch (normalise + dot + argmax).  text_feat [P,E] / image_feat [B,E] in `dtype`; class_of_prompt int32 [P] in [0,C).
 */
int ngu_zero_shot_prototypes(const void* text_feat, const int* class_of_prompt, float* proto, int P, int E, int C, int dtype, void* stream);
int ngu_zero_shot_score(const void* image_feat, const float* proto, float* logits, int* pred, int B, int E, int C, float scale,
                        int dtype, void* stream);

/* Patch-embed im2col (stride == kernel, timm PatchEmbed / CLIP conv1): images fp32 NCHW [B,3,R,R]
 * -> [B*(R/P)^2, Kp] with row pitch Kp = 3*P*P rounded up to a multiple of 8 (P = 14: 588 -> 592; the caller zero-fills
 * the buffer once, the kernel writes the 3*P*P live columns; pixels past the last whole patch are dropped like the conv);
 * and token assembly x0[b,0]=cls+pos[0], x0[b,1+p]=patch[b,p]+pos[1+p]. */
int ngu_patchify(const float* img, void* out, int B, int R, int P, int dtype, void* stream);
int ngu_assemble_tokens(const void* patch, const float* cls, const float* pos, void* out, int B, int np, int D,
                        int dtype, void* stream);
/* BERT input embeddings (HF BertEmbeddings before its LayerNorm; text tower of open_clip HFTextEncoder, pinned dep):
 * out[b,s,:] = word[ids[b,s]] + pos[s] + type0.  ids int64 [B,S]; tables fp32. */
int ngu_embed_tokens(const int64_t* ids, const float* word, const float* pos, const float* type0, void* out, int B, int S,
                     int D, int vocab, int dtype, void* stream);
/* out (dtype) = scale * in (fp32 [rows, cols]), optionally transposed to [cols, rows] */
int ngu_cast_f32(const float* in, void* out, int rows, int cols, int transpose, float scale, int dtype, void* stream);

/* The same conversion for a table of tensors in ONE launch: the per-step refresh of the bf16 shadows (and transposes)
 * of the trainable adapter projections after the optimizer update (mona.py:127,148 weights; lora.py:48-51 factors).
 * `items` is a DEVICE array of n entries; pointers in it are device pointers. */
typedef struct ngu_cast_item {
  const float* in;            /* fp32 [rows, cols] */
  void* out;                  /* dtype [rows, cols], or [cols, rows] when transpose */
  int rows, cols;
  int transpose;
  float scale;
} ngu_cast_item;
int ngu_cast_f32_batch(const ngu_cast_item* items, int n, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NGU_B200_H_ */
