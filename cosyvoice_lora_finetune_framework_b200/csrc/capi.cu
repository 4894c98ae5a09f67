// extern "C" entry points of libcvflow.so (declared in include/cvflow.h).
#include "../../include/cvflow.h"
#include "gemm.h"
#include "estimator.h"
#include "kernels.h"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <vector>

namespace cvflow {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
char* error_buf() { return g_err; }
int error_buf_len() { return (int)sizeof(g_err); }
}  // namespace cvflow

using namespace cvflow;

extern "C" CVFLOW_API const char* cvflow_last_error(void) { return g_err; }
extern "C" CVFLOW_API int cvflow_abi_version(void) { return 1; }

extern "C" CVFLOW_API int cvflow_gemm(const cvflow_gemm_desc* d, void* stream) {
  if (!d) { set_error("cvflow_gemm: null desc"); return CVFLOW_ERR_ARG; }
  GemmArgs a;
  for (int s = 0; s < 2; ++s) {
    a.A[s] = d->A[s]; a.a_rows[s] = d->a_rows[s]; a.a_cols[s] = d->a_cols[s];
    a.a_ld[s] = d->a_ld[s]; a.a_bstride[s] = d->a_bstride[s];
  }
  a.nbatch = d->nbatch; a.bf16 = d->dtype == CVFLOW_DTYPE_BF16;
  a.W = d->W; a.N = d->N; a.Ktot = d->Ktot; a.nseg = d->nseg;
  for (int s = 0; s < 8; ++s) {
    a.seg[s].a_map = d->seg[s].a_map; a.seg[s].row_shift = d->seg[s].row_shift;
    a.seg[s].a_col0 = d->seg[s].a_col0; a.seg[s].nkb = d->seg[s].nkb;
  }
  a.R = d->R; a.rmul = d->rmul; a.roff = d->roff; a.out_rows = d->out_rows;
  a.out = d->out; a.out_f32 = d->out_f32; a.transposed_out = d->transposed_out; a.ldc = d->ldc;
  a.col_off = d->col_off; a.n_valid = d->n_valid; a.alpha = d->alpha; a.act = d->act;
  a.bias = d->bias; a.aux_out = d->aux_out; a.mul_src = d->mul_src; a.ld_aux = d->ld_aux;
  a.rowmask = d->rowmask; a.resid = d->resid; a.ldr = d->ldr; a.dbg = (long long*)d->dbg;
  a.gn_part = d->gn_part; a.ln_gamma = d->ln_gamma; a.ln_beta = d->ln_beta;
  GemmParams p;
  if (gemm_prepare(a, &p, error_buf(), error_buf_len())) return CVFLOW_ERR_ARG;
  int r = gemm_launch(p, (cudaStream_t)stream);
  if (r) { set_error("cvflow_gemm: launch failed: %s", cudaGetErrorString((cudaError_t)(-r))); return CVFLOW_ERR_CUDA; }
  return CVFLOW_OK;
}

// ---------------------------------------------------------------------------------------------
struct cvflow_estimator { Estimator* e; };

extern "C" CVFLOW_API int cvflow_create(const cvflow_config* c, cvflow_estimator** out) {
  if (!c || !out) { set_error("cvflow_create: null argument"); return CVFLOW_ERR_ARG; }
  if (c->n_blocks < 1 || c->n_mid < 0 || c->lora_r < 0 || c->lora_r > 16 ||
      (c->lora_r != 0 && c->lora_r != 4 && c->lora_r != 8 && c->lora_r != 16)) {
    set_error("cvflow_create: unsupported config (n_blocks %d, n_mid %d, lora_r %d; rank must be 0/4/8/16)",
              c->n_blocks, c->n_mid, c->lora_r);
    return CVFLOW_ERR_UNSUPPORTED;
  }
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_error("cvflow_create: no CUDA device");
    return CVFLOW_ERR_CUDA;
  }
  if (prop.major != 10) {
    set_error("cvflow_create: kernels are built for sm_100a only, device is sm_%d%d (no fallback)", prop.major,
              prop.minor);
    return CVFLOW_ERR_UNSUPPORTED;
  }
  EstimatorConfig ec;
  ec.n_blocks = c->n_blocks; ec.n_mid = c->n_mid; ec.bf16 = c->dtype == CVFLOW_DTYPE_BF16; ec.gelu_erf = c->gelu_erf;
  ec.lora_r = c->lora_r; ec.lora_scaling = c->lora_scaling;
  cvflow_estimator* h = new cvflow_estimator;
  h->e = new Estimator(ec);
  *out = h;
  return CVFLOW_OK;
}
extern "C" CVFLOW_API void cvflow_destroy(cvflow_estimator* h) {
  if (h) { delete h->e; delete h; }
}
extern "C" CVFLOW_API int cvflow_bind(cvflow_estimator* h, const char* name, void* ptr, int64_t numel, int32_t dtype) {
  if (!h || !name || !ptr) { set_error("cvflow_bind: null argument"); return CVFLOW_ERR_ARG; }
  return h->e->bind(name, ptr, (long)numel, dtype);
}
extern "C" CVFLOW_API int64_t cvflow_workspace_bytes(cvflow_estimator* h, int32_t B, int32_t T, int32_t training) {
  if (!h || B < 1 || T < 1) { set_error("cvflow_workspace_bytes: bad argument"); return CVFLOW_ERR_ARG; }
  return h->e->workspace_bytes(B, T, training);
}
extern "C" CVFLOW_API int cvflow_set_workspace(cvflow_estimator* h, void* ptr, int64_t bytes) {
  if (!h || !ptr) { set_error("cvflow_set_workspace: null argument"); return CVFLOW_ERR_ARG; }
  h->e->set_workspace(ptr, (long)bytes);
  return CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_lora_refresh(cvflow_estimator* h, void* stream) {
  if (!h) { set_error("cvflow_lora_refresh: null handle"); return CVFLOW_ERR_ARG; }
  return h->e->lora_refresh((cudaStream_t)stream) ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_lora_refresh_factors(cvflow_estimator* h, void* stream) {
  if (!h) { set_error("cvflow_lora_refresh_factors: null handle"); return CVFLOW_ERR_ARG; }
  return h->e->lora_refresh((cudaStream_t)stream, false) ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_lora_prepare(cvflow_estimator* h, void* stream) {
  if (!h) { set_error("cvflow_lora_prepare: null handle"); return CVFLOW_ERR_ARG; }
  return h->e->lora_refresh((cudaStream_t)stream, false) ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_estimator_forward(cvflow_estimator* h, const cvflow_estimator_io* io, void* stream) {
  if (!h || !io || !io->x || !io->mask || !io->mu || !io->t || !io->out || io->B < 1 || io->T < 1) {
    set_error("cvflow_estimator_forward: null/invalid argument");
    return CVFLOW_ERR_ARG;
  }
  EstimatorIO e;
  e.x = io->x; e.x_nb = io->x_nb; e.mask = io->mask; e.mask_nb = io->mask_nb; e.mu = io->mu; e.mu_nb = io->mu_nb;
  e.t = io->t; e.t_nb = io->t_nb; e.spks = io->spks; e.spks_nb = io->spks_nb > 0 ? io->spks_nb : 1;
  e.cond = io->cond; e.cond_nb = io->cond_nb > 0 ? io->cond_nb : 1; e.keep = io->keep; e.out = io->out;
  e.B = io->B; e.T = io->T; e.iso_len = io->iso_len; e.training = io->training;
  if (e.x_nb < 1 || e.mask_nb < 1 || e.mu_nb < 1 || e.t_nb < 1) { set_error("cvflow_estimator_forward: *_nb must be >= 1"); return CVFLOW_ERR_ARG; }
  return h->e->forward(e, (cudaStream_t)stream) ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_estimator_backward(cvflow_estimator* h, const void* dpred16, float grad_scale,
                                                    const float* grad_scale_dev, void* stream) {
  if (!h || !dpred16) { set_error("cvflow_estimator_backward: null argument"); return CVFLOW_ERR_ARG; }
  return h->e->backward(dpred16, grad_scale, grad_scale_dev, (cudaStream_t)stream) ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_estimator_backward_inputs(cvflow_estimator* h, const void* dpred16, float grad_scale,
                                                           const float* grad_scale_dev, const cvflow_input_grads* ig,
                                                           void* stream) {
  if (!h || !dpred16 || !ig) { set_error("cvflow_estimator_backward_inputs: null argument"); return CVFLOW_ERR_ARG; }
  InputGrads g{ig->dx, ig->dmu, ig->dspks, ig->dcond};
  return h->e->backward(dpred16, grad_scale, grad_scale_dev, (cudaStream_t)stream, &g) ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_set_lora_dropout(cvflow_estimator* h, float p, uint64_t seed, const uint8_t* debug_mask,
                                                  int64_t debug_rows) {
  if (!h) { set_error("cvflow_set_lora_dropout: null handle"); return CVFLOW_ERR_ARG; }
  return h->e->set_lora_dropout(p, (unsigned long long)seed, debug_mask, (long)debug_rows) ? CVFLOW_ERR_ARG : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_set_grad_chunks(cvflow_estimator* h, int32_t n, const int32_t* first_block, void** events) {
  if (!h) { set_error("cvflow_set_grad_chunks: null handle"); return CVFLOW_ERR_ARG; }
  return h->e->set_grad_chunks(n, first_block, reinterpret_cast<cudaEvent_t*>(events)) ? CVFLOW_ERR_ARG : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_solve_capture(cvflow_estimator* h, int32_t T, int32_t n_steps, float cfg_rate, float* x,
                                               const float* mask, const float* mu, const float* spks, const float* cond,
                                               const float* t, const float* dt, float* d_scratch, void* stream) {
  if (!h) { set_error("cvflow_solve_capture: null handle"); return CVFLOW_ERR_ARG; }
  if (!stream) { set_error("cvflow_solve_capture: stream capture needs a non-default stream"); return CVFLOW_ERR_ARG; }
  return h->e->solve_capture(T, n_steps, cfg_rate, x, mask, mu, spks, cond, t, dt, d_scratch, (cudaStream_t)stream)
             ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_solve_replay(cvflow_estimator* h, int32_t T, int32_t n_steps, void* stream) {
  if (!h) { set_error("cvflow_solve_replay: null handle"); return CVFLOW_ERR_ARG; }
  return h->e->solve_replay(T, n_steps, (cudaStream_t)stream) ? CVFLOW_ERR_ARG : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_solve_release(cvflow_estimator* h) {
  if (!h) { set_error("cvflow_solve_release: null handle"); return CVFLOW_ERR_ARG; }
  h->e->solve_release();
  return CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_time_embed(cvflow_estimator* h, const float* t, int32_t t_nb, float* out, float* scratch,
                                            int32_t B, void* stream) {
  if (!h || !t || !out || !scratch || B < 1 || t_nb < 1) { set_error("cvflow_time_embed: null/invalid argument"); return CVFLOW_ERR_ARG; }
  return h->e->time_embed(t, t_nb, out, scratch, B, (cudaStream_t)stream) ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_lora_dropout_seed(cvflow_estimator* h, uint64_t* out, const uint64_t* in) {
  if (!h) { set_error("cvflow_lora_dropout_seed: null handle"); return CVFLOW_ERR_ARG; }
  unsigned long long o = 0ull, i = in ? (unsigned long long)*in : 0ull;
  if (h->e->lora_dropout_seed(out ? &o : nullptr, in ? &i : nullptr)) return CVFLOW_ERR_CUDA;
  if (out) *out = (uint64_t)o;
  return CVFLOW_OK;
}
extern "C" CVFLOW_API int64_t cvflow_launch_count(cvflow_estimator* h) { return h ? h->e->launches() : 0; }

#define RET_LAUNCH(call, what)                                                                       \
  do {                                                                                               \
    int r_ = (call);                                                                                 \
    if (r_) { set_error(what ": %s", cudaGetErrorString((cudaError_t)(-r_))); return CVFLOW_ERR_CUDA; } \
    return CVFLOW_OK;                                                                                \
  } while (0)

extern "C" CVFLOW_API int cvflow_cfm_prep(const float* x1, const float* z, const float* t, float* y, int32_t B,
                                          int32_t T, float sigma_min, void* stream) {
  if (!x1 || !z || !t || !y) { set_error("cvflow_cfm_prep: null argument"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_cfm_prep(x1, z, t, y, B, T, sigma_min, (cudaStream_t)stream), "cvflow_cfm_prep");
}
extern "C" CVFLOW_API int cvflow_cfm_loss(const float* pred, const float* x1, const float* z, const float* w,
                                          const float* mask, float* scal, float* partials, void* dpred16, int32_t B,
                                          int32_t T, float sigma_min, float loss_scale, int32_t dtype,
                                          const float* wsum_dev, void* stream) {
  if (!pred || !x1 || !z || !w || !mask || !scal || !partials) { set_error("cvflow_cfm_loss: null argument"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_cfm_loss(pred, x1, z, w, mask, scal, partials, dpred16, B, T, sigma_min, loss_scale,
                             dtype == CVFLOW_DTYPE_BF16, wsum_dev, (cudaStream_t)stream), "cvflow_cfm_loss");
}
extern "C" CVFLOW_API int cvflow_euler_update(float* x, const float* d, const float* dt, int32_t step, float cfg_rate,
                                              int64_t n, void* stream) {
  if (!x || !d || !dt) { set_error("cvflow_euler_update: null argument"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_euler_update(x, d, dt, step, cfg_rate, (long)n, (cudaStream_t)stream), "cvflow_euler_update");
}
extern "C" CVFLOW_API int cvflow_sumsq(const float* g, int64_t n, float* partials, float* out, void* stream) {
  if (!g || !partials || !out) { set_error("cvflow_sumsq: null argument"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_sumsq(g, (long)n, partials, out, (cudaStream_t)stream), "cvflow_sumsq");
}
extern "C" CVFLOW_API int cvflow_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq,
                                            float grad_unscale, float max_norm, float lr, float beta1, float beta2,
                                            float eps, float weight_decay, int32_t step, int32_t* found_inf,
                                            const float* hyper_dev, void* stream) {
  if (!p || !g || !m || !v || !sumsq) { set_error("cvflow_adamw_step: null argument"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_adamw(p, g, m, v, (long)n, sumsq, grad_unscale, max_norm, lr, beta1, beta2, eps, weight_decay, step,
                          found_inf, hyper_dev, (cudaStream_t)stream), "cvflow_adamw_step");
}

extern "C" CVFLOW_API int cvflow_optim_advance(int32_t* state, float* hyper, const float* sumsq, const float* sumsq2,
                                               float grad_unscale, float base_lr, int32_t warmup_steps, int32_t total_steps,
                                               float min_lr, float beta1, float beta2, void* stream) {
  if (!state || !hyper || !sumsq) { set_error("cvflow_optim_advance: null argument"); return CVFLOW_ERR_ARG; }
  if (!(base_lr > 0.f)) { set_error("cvflow_optim_advance: base_lr must be positive"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_optim_advance(state, hyper, sumsq, sumsq2, grad_unscale, base_lr, warmup_steps, total_steps, min_lr, beta1,
                                  beta2, (cudaStream_t)stream), "cvflow_optim_advance");
}

extern "C" CVFLOW_API int cvflow_set_profile(cvflow_estimator* h, int32_t on) {
  if (!h) return CVFLOW_ERR_ARG;
  h->e->set_profile(on);
  return CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_profile_read(cvflow_estimator* h, double* ms, int64_t* counts, double* flops, int32_t n) {
  if (!h || !ms || !counts || !flops) return CVFLOW_ERR_ARG;
  long c[16]; if (n > 16) n = 16;
  int r = h->e->profile_read(ms, c, flops, n);
  for (int i = 0; i < n; ++i) counts[i] = c[i];
  return r ? CVFLOW_ERR_CUDA : CVFLOW_OK;
}

extern "C" CVFLOW_API int cvflow_debug_attention_stamps(void* buf) { attn_set_debug_buffer(buf); return CVFLOW_OK; }
extern "C" CVFLOW_API int cvflow_debug_mlp_stamps(void* buf) { mlp_set_debug_buffer(buf); return CVFLOW_OK; }

// ---------------------------------------------------------------------------------------------
extern "C" CVFLOW_API int64_t cvflow_attention_scratch_ints(int32_t B, int32_t L) { return attn_kinfo_ints(B, L); }
extern "C" CVFLOW_API int cvflow_attention_forward(const void* qkv, int64_t ldq, int32_t B, int32_t L, int32_t dtype,
                                                   const float* keymask, int32_t* kmax_scratch, int32_t iso_p, void* o,
                                                   float* lse, void* stream) {
  if (!qkv || !keymask || !kmax_scratch || !o || B < 1 || L < 1 || ldq < 1536) {
    set_error("cvflow_attention_forward: null/invalid argument");
    return CVFLOW_ERR_ARG;
  }
  std::vector<uint8_t> plan(attn_plan_bytes());
  if (attn_fwd_prepare(plan.data(), qkv, (long)ldq, B, L, dtype == CVFLOW_DTYPE_BF16, error_buf(), error_buf_len()))
    return CVFLOW_ERR_ARG;
  int r = launch_attn_kinfo(keymask, B, L, kmax_scratch, (cudaStream_t)stream);
  if (!r) r = attn_fwd_launch(plan.data(), kmax_scratch, iso_p, o, lse, (cudaStream_t)stream);
  if (r) { set_error("cvflow_attention_forward: %s", cudaGetErrorString((cudaError_t)(-r))); return CVFLOW_ERR_CUDA; }
  return CVFLOW_OK;
}
extern "C" CVFLOW_API int cvflow_attention_backward(const void* qkv, int64_t ldq, int32_t B, int32_t L, int32_t dtype,
                                                    const float* keymask, int32_t* kmax_scratch, int32_t iso_p,
                                                    const void* o, const float* lse, const void* dout,
                                                    float* delta_scratch, void* dqkv, void* stream) {
  if (!qkv || !keymask || !kmax_scratch || !o || !lse || !dout || !delta_scratch || !dqkv || B < 1 || L < 1 || ldq < 1536) {
    set_error("cvflow_attention_backward: null/invalid argument");
    return CVFLOW_ERR_ARG;
  }
  std::vector<uint8_t> plan(attn_plan_bytes());
  if (attn_bwd_prepare(plan.data(), qkv, (long)ldq, dout, B, L, dtype == CVFLOW_DTYPE_BF16, error_buf(), error_buf_len()))
    return CVFLOW_ERR_ARG;
  int r = launch_attn_kinfo(keymask, B, L, kmax_scratch, (cudaStream_t)stream);
  if (!r) r = attn_bwd_launch(plan.data(), dout, kmax_scratch, iso_p, o, lse, delta_scratch, dqkv, (cudaStream_t)stream);
  if (r) { set_error("cvflow_attention_backward: %s", cudaGetErrorString((cudaError_t)(-r))); return CVFLOW_ERR_CUDA; }
  return CVFLOW_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" CVFLOW_API int cvflow_mlp_forward(const void* x16, const void* w1, const float* b1, const void* w2,
                                             const float* b2, const float* resid32, float* out32, void* pre16, int64_t M,
                                             int32_t dtype, int32_t gelu_erf, void* stream) {
  if (!x16 || !w1 || !w2 || !out32 || !pre16 || M < 1) { set_error("cvflow_mlp_forward: null/invalid argument"); return CVFLOW_ERR_ARG; }
  std::vector<uint8_t> plan(mlp_plan_bytes());
  if (mlp_prepare(plan.data(), 0, x16, w1, b1, w2, b2, resid32, out32, pre16, (long)M, dtype == CVFLOW_DTYPE_BF16, gelu_erf,
                  error_buf(), error_buf_len()))
    return CVFLOW_ERR_ARG;
  RET_LAUNCH(mlp_launch(plan.data(), (cudaStream_t)stream), "cvflow_mlp_forward");
}
extern "C" CVFLOW_API int cvflow_mlp_backward(const void* dy16, const void* w2_t, const void* pre16, const void* w1_t,
                                              void* dx16, int64_t M, int32_t dtype, int32_t gelu_erf, void* stream) {
  if (!dy16 || !w2_t || !pre16 || !w1_t || !dx16 || M < 1) { set_error("cvflow_mlp_backward: null/invalid argument"); return CVFLOW_ERR_ARG; }
  std::vector<uint8_t> plan(mlp_plan_bytes());
  if (mlp_prepare(plan.data(), 1, dy16, w2_t, nullptr, w1_t, nullptr, nullptr, dx16, const_cast<void*>(pre16), (long)M,
                  dtype == CVFLOW_DTYPE_BF16, gelu_erf, error_buf(), error_buf_len()))
    return CVFLOW_ERR_ARG;
  RET_LAUNCH(mlp_launch(plan.data(), (cudaStream_t)stream), "cvflow_mlp_backward");
}

// ---- the inputs of the path (regulator.cu) ---------------------------------------------------------------------
static int regulator_args(const cvflow_regulator_weights* w, const cvflow_regulator_io* io, RegulatorWeights* W, RegulatorIO* I,
                          const char* who) {
  if (!w || !io || !io->src || !io->saved) { set_error("%s: null argument", who); return CVFLOW_ERR_ARG; }
  if (io->B < 1 || io->T < 1 || io->n_src < 1 || io->n_seg < 1 || io->n_seg > 4) {
    set_error("%s: B %d, T %d, n_src %d, n_seg %d out of range", who, io->B, io->T, io->n_src, io->n_seg);
    return CVFLOW_ERR_ARG;
  }
  for (int l = 0; l < 5; ++l) {
    if (!w->wf[l] || !w->wb[l] || !w->bias[l] || (l < 4 && (!w->gamma[l] || !w->beta[l]))) {
      set_error("%s: weight image %d missing", who, l);
      return CVFLOW_ERR_ARG;
    }
    W->wf[l] = w->wf[l]; W->wb[l] = w->wb[l]; W->bias[l] = w->bias[l];
    if (l < 4) { W->gamma[l] = w->gamma[l]; W->beta[l] = w->beta[l]; }
  }
  int covered = 0;
  for (int s = 0; s < io->n_seg; ++s) {
    const int32_t* g = io->seg[s];
    if (g[0] < 0 || g[1] < 1 || g[0] + g[1] > io->n_src || g[2] != covered || g[3] < 1) {
      set_error("%s: segment %d {%d,%d,%d,%d} invalid (segments must tile the frames in order)", who, s, g[0], g[1], g[2], g[3]);
      return CVFLOW_ERR_ARG;
    }
    I->seg[s] = RegSeg{g[0], g[1], g[2], g[3]};
    covered += g[3];
  }
  if (covered != io->T) { set_error("%s: segments cover %d frames, T = %d", who, covered, io->T); return CVFLOW_ERR_ARG; }
  for (int s = io->n_seg; s < 4; ++s) I->seg[s] = RegSeg{0, 0, 0, 0};
  I->src = io->src; I->B = io->B; I->n_src = io->n_src; I->T = io->T; I->n_seg = io->n_seg;
  I->lens = io->lens; I->blind = io->blind; I->out = io->out; I->channel_major = io->channel_major; I->saved = io->saved;
  return CVFLOW_OK;
}
extern "C" CVFLOW_API int64_t cvflow_regulator_saved_floats(int32_t B, int32_t T) { return regulator_saved_floats(B, T); }
extern "C" CVFLOW_API int64_t cvflow_regulator_scratch_floats(int32_t B, int32_t T) { return regulator_scratch_floats(B, T); }
extern "C" CVFLOW_API int cvflow_regulator_forward(const cvflow_regulator_weights* w, const cvflow_regulator_io* io, void* stream) {
  RegulatorWeights W; RegulatorIO I;
  if (int r = regulator_args(w, io, &W, &I, "cvflow_regulator_forward")) return r;
  if (!io->out) { set_error("cvflow_regulator_forward: null output"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_regulator_forward(W, I, (cudaStream_t)stream), "cvflow_regulator_forward");
}
extern "C" CVFLOW_API int cvflow_regulator_backward(const cvflow_regulator_weights* w, const cvflow_regulator_io* io,
                                                    const float* dout, float* dsrc, float* scratch, void* stream) {
  RegulatorWeights W; RegulatorIO I;
  if (int r = regulator_args(w, io, &W, &I, "cvflow_regulator_backward")) return r;
  if (!dout || !dsrc || !scratch) { set_error("cvflow_regulator_backward: null argument"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_regulator_backward(W, I, dout, dsrc, scratch, (cudaStream_t)stream), "cvflow_regulator_backward");
}
extern "C" CVFLOW_API int cvflow_path_inputs_pack(const float* feat, const float* cross, int32_t cross_T, const int32_t* desc,
                                                  float mel_mean, float mel_std, float silence, float* x1, float* cond, float* mask,
                                                  int32_t B, int32_t T, void* stream) {
  if (!feat || !desc || !x1 || !cond || !mask || B < 1 || T < 1 || (cross && cross_T < 1) || mel_std == 0.f) {
    set_error("cvflow_path_inputs_pack: null/invalid argument");
    return CVFLOW_ERR_ARG;
  }
  RET_LAUNCH(launch_path_inputs_pack(feat, cross, cross_T, desc, mel_mean, mel_std, silence, x1, cond, mask, B, T,
                                     (cudaStream_t)stream), "cvflow_path_inputs_pack");
}
extern "C" CVFLOW_API int cvflow_spk_affine(const float* e, const float* W, const float* bias, float* out, int32_t B, int32_t K,
                                            int32_t N, void* stream) {
  if (!e || !W || !out || B < 1 || K < 1 || K > 8192 || N < 1) { set_error("cvflow_spk_affine: null/invalid argument"); return CVFLOW_ERR_ARG; }
  RET_LAUNCH(launch_spk_affine(e, W, bias, out, B, K, N, (cudaStream_t)stream), "cvflow_spk_affine");
}
