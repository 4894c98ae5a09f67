// Fused optimiser tail on the flat fp32 LoRA bucket: global-norm clip + AdamW (decoupled weight
// decay), matching torch.optim.AdamW + clip_grad_norm_(1.0) as used by the reference trainer
// (train_joint.py:198-226, gradient_clip_val=1.0 at :353-355). A non-finite gradient norm skips
// the update and raises found_inf (the GradScaler behaviour of the reference's '16-mixed' run).
#include "kernels.h"
#include "common.cuh"

namespace cvflow {

#define LAUNCH_RET() do { cudaError_t e_ = cudaGetLastError(); return e_ == cudaSuccess ? 0 : -(int)e_; } while (0)

static constexpr int kSumsqBlocks = 296;

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long n, float* __restrict__ partials) {
  __shared__ float red[8];
  float s = 0.f;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) { const float v = g[i]; s += v * v; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    partials[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)partials[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[0] = (float)t;
  }
}
int launch_sumsq(const float* g, long n, float* partials, float* out_sumsq, cudaStream_t st) {
  sumsq_partial_kernel<<<kSumsqBlocks, 256, 0, st>>>(g, n, partials);
  sumsq_final_kernel<<<1, 256, 0, st>>>(partials, kSumsqBlocks, out_sumsq);
  LAUNCH_RET();
}

// Per-step optimiser scalars advanced ON THE DEVICE, inside the captured step (no host write can race a replay):
//   state[0] = optimiser steps applied so far (k), state[1] = steps skipped because the gradient norm was not finite
//   hyper    = {lr of this step, 1 - beta1^t, sqrt(1 - beta2^t), 1 = apply / 0 = skip}, t = k + 1
// The learning rate follows the reference's LambdaLR with interval 'step' (train_joint.py:210-226): the (k+1)-th
// optimiser step runs with base_lr * lambda(k) -- linear warm-up, then cosine decay to min_lr (the reference's
// literal 3.14159). A skipped step (GradScaler semantics of the '16-mixed' run) advances neither k nor the schedule.
__global__ void optim_advance_kernel(int* __restrict__ state, float* __restrict__ hyper, const float* __restrict__ sumsq,
                                     const float* __restrict__ sumsq2, float grad_unscale, float base_lr, int warmup_steps,
                                     int total_steps, float min_lr, float beta1, float beta2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float ss = sumsq[0] + (sumsq2 ? sumsq2[0] : 0.f);
  const float total = sqrtf(ss) * grad_unscale;
  if (!isfinite(total)) {
    state[1] += 1;
    hyper[3] = 0.f;
    return;
  }
  const int k = state[0];
  float lam = 1.f;
  if (total_steps > 0) {
    if (k < warmup_steps) lam = (float)k / (float)max(1, warmup_steps);
    else {
      const float progress = (float)(k - warmup_steps) / (float)max(1, total_steps - warmup_steps);
      lam = fmaxf(min_lr / base_lr, 0.5f * (1.f + cosf(progress * 3.14159f)));
    }
  }
  const float t = (float)(k + 1);
  hyper[0] = base_lr * lam;
  hyper[1] = 1.f - powf(beta1, t);
  hyper[2] = sqrtf(1.f - powf(beta2, t));
  hyper[3] = 1.f;
  state[0] = k + 1;
}
int launch_optim_advance(int* state, float* hyper, const float* sumsq, const float* sumsq2, float grad_unscale, float base_lr,
                         int warmup_steps, int total_steps, float min_lr, float beta1, float beta2, cudaStream_t st) {
  optim_advance_kernel<<<1, 32, 0, st>>>(state, hyper, sumsq, sumsq2, grad_unscale, base_lr, warmup_steps, total_steps, min_lr,
                                         beta1, beta2);
  LAUNCH_RET();
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long n, const float* __restrict__ sumsq,
                                                    float grad_unscale, float max_norm, float lr, float beta1, float beta2,
                                                    float eps, float wd, float bc1, float bc2_sqrt, int* __restrict__ found_inf,
                                                    const float* __restrict__ hyper) {
  if (hyper) {   // {lr, 1-beta1^t, sqrt(1-beta2^t), apply flag} written by optim_advance_kernel earlier in the same step
    lr = hyper[0]; bc1 = hyper[1]; bc2_sqrt = hyper[2];
    if (hyper[3] == 0.f) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && found_inf) *found_inf = 1;
      return;
    }
  }
  const float total = sqrtf(sumsq[0]) * grad_unscale;
  if (!isfinite(total)) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && found_inf) *found_inf = 1;
    return;
  }
  float coef = grad_unscale;
  if (max_norm > 0.f) coef *= fminf(1.f, max_norm / (total + 1e-6f));
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}
int launch_adamw(float* p, const float* g, float* m, float* v, long n, const float* sumsq, float grad_unscale,
                 float max_norm, float lr, float beta1, float beta2, float eps, float wd, int step, int* found_inf,
                 const float* hyper_dev, cudaStream_t st) {
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 4 ? (n + 255) / 256 : 148 * 4);
  adamw_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, sumsq, grad_unscale, max_norm, lr, beta1, beta2, eps, wd, bc1,
                                     bc2_sqrt, found_inf, hyper_dev);
  LAUNCH_RET();
}

}  // namespace cvflow
