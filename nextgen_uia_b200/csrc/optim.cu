// Fused optimiser step for the flat adapter parameter buffer: global grad-norm clipping + AdamW + (optional)
// on-device non-finite-loss guard, replacing clip_grad_norm_ / optimizer.step() / zero_grad() of the reference loop
// (src/models/biomedclip/finetune.py:244-255, :281-285, :296-303).  Two launches over ~1.3 M fp32 elements:
//   sqnorm:  *out += sum g^2            (block reduce + one atomic per block)
//   adamw :  c = min(1, max_norm / (sqrt(*gsq) + 1e-6));  g *= c;  decoupled weight decay;  Adam moments; update
// torch.optim.AdamW / torch.nn.utils.clip_grad_norm_ semantics (error_if_nonfinite = False).
#include "common.cuh"
#include "kernels.h"

namespace ngu {
namespace {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ x, size_t n, float* __restrict__ out) {
  pdl_prologue();
  float s = 0.f;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) s = fmaf(x[i], x[i], s);
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = part[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// state (int64[4], device): [0] updates applied so far, [1] poison flag of the current accumulation window (a micro-step saw
// a non-finite loss), [2] updates skipped, [3] micro-steps run (the dropout seed counter)
__global__ void guard_tick_kernel(int64_t* state, const float* loss, const float* gsq, int mode) {
  pdl_prologue();
  if (mode == 0) {          // after the loss of a micro-step is known
    if (loss != nullptr && !isfinite(*loss)) state[1] = 1;
  } else if (mode == 1) {   // after the optimiser launch of an update step
    const bool bad = state[1] != 0 || (gsq != nullptr && !isfinite(*gsq));
    if (bad) state[2] += 1; else state[0] += 1;
    state[1] = 0;
    state[3] += 1;
  } else {                  // end of a micro-step that does not update
    state[3] += 1;
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(ngu_adamw_desc d) {
  pdl_prologue();
  // skip the whole update when a loss of this accumulation window or the gradient norm is not finite (reference:
  // `if not torch.isfinite(loss): continue` before backward, finetune.py:281-285; here the poisoned window is dropped
  // and neither the step count nor the LR schedule advances, like the reference which also skips scheduler.step())
  bool skip = d.loss != nullptr && !isfinite(*d.loss);
  if (d.gsq != nullptr && !isfinite(*d.gsq)) skip = true;
  int step = d.step;
  float lr = d.lr;
  if (d.state != nullptr) {
    if (d.state[1] != 0) skip = true;
    const int t = int(d.state[0]);
    step = t + 1;
    if (d.t_max > 0) lr = d.lr_min + (d.lr - d.lr_min) * (1.0f + cospif(float(t) / float(d.t_max))) * 0.5f;   // CosineAnnealingLR closed form
  }
  float clip = 1.f;
  if (d.max_norm > 0.f && d.gsq != nullptr) {
    const float c = d.max_norm / (sqrtf(*d.gsq) + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  const float bc1 = 1.f - powf(d.beta1, float(step));
  const float bc2 = 1.f - powf(d.beta2, float(step));
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < d.n; i += size_t(gridDim.x) * blockDim.x) {
    if (!skip) {
      const float g = d.grad[i] * clip;
      float p = d.param[i] * (1.f - lr * d.weight_decay);
      const float m = d.beta1 * d.m[i] + (1.f - d.beta1) * g;
      const float v = d.beta2 * d.v[i] + (1.f - d.beta2) * g * g;
      d.m[i] = m;
      d.v[i] = v;
      p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + d.eps);
      d.param[i] = p;
    }
    if (d.zero_grad) d.grad[i] = 0.f;
  }
}

}  // namespace

int sqnorm(const float* x, size_t n, float* out, cudaStream_t st) {
  if (n == 0) { set_last_error("sqnorm: empty"); return NGU_ERR_SHAPE; }
  size_t g = (n + 256 * 8 - 1) / (256 * 8);
  const size_t cap = size_t(sm_count()) * 4;
  if (g > cap) g = cap;
  launch_pdl(sqnorm_kernel, dim3(int(g)), dim3(256), size_t(0), st, x, n, out);
  return check_launch("sqnorm");
}

int guard_tick(int64_t* state, const float* loss, const float* gsq, int mode, cudaStream_t st) {
  if (!state || mode < 0 || mode > 2) { set_last_error("guard_tick: bad arguments"); return NGU_ERR_ARG; }
  launch_pdl(guard_tick_kernel, dim3(1), dim3(1), size_t(0), st, state, loss, gsq, mode);
  return check_launch("guard_tick");
}

int adamw_step(const ngu_adamw_desc& d, cudaStream_t st) {
  if (d.n <= 0 || !d.param || !d.grad || !d.m || !d.v || (d.step < 1 && d.state == nullptr)) { set_last_error("adamw: bad arguments"); return NGU_ERR_ARG; }
  size_t g = (size_t(d.n) + 256 * 4 - 1) / (256 * 4);
  const size_t cap = size_t(sm_count()) * 4;
  if (g > cap) g = cap;
  launch_pdl(adamw_kernel, dim3(int(g)), dim3(256), size_t(0), st, d);
  return check_launch("adamw");
}

}  // namespace ngu
