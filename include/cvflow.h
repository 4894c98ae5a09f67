/* cvflow C ABI — the drop-in boundary of the B200-native flow-LoRA hot path.
 *
 * Plain C: raw device pointers, sizes, an explicit cudaStream_t (passed as void*), int status
 * codes (0 = ok, <0 = error; text via cvflow_last_error()). No torch types cross this boundary.
 *
 * The reference (leeoisaboy/cosyvoice-lora-finetune-framework) has no FFI of its own for this
 * path; its only precedent for a non-PyTorch estimator is the TensorRT hook of the vendored
 * upstream, which binds raw data_ptr()s by tensor name and runs on the caller's stream
 * (cosyvoice/flow/flow_matching.py:125-152, cosyvoice/utils/common.py:171-186). The entry points
 * below follow that calling convention; each one cites the reference code it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 */
#ifndef CVFLOW_H_
#define CVFLOW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CVFLOW_API __attribute__((visibility("default")))
#else
#define CVFLOW_API
#endif

#define CVFLOW_OK 0
#define CVFLOW_ERR_ARG (-1)
#define CVFLOW_ERR_CUDA (-2)
#define CVFLOW_ERR_UNSUPPORTED (-3)

#define CVFLOW_DTYPE_F16 0
#define CVFLOW_DTYPE_BF16 1

/* Thread-local text of the last error returned by any cvflow_* call on this thread. */
CVFLOW_API const char* cvflow_last_error(void);
/* ABI version of this library (bumped on any signature change). */
CVFLOW_API int cvflow_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction engine (tcgen05 / TMEM / TMA). Exposed so each fused linear / conv of the
 * estimator can be parity-tested on its own. Replaces F.linear / nn.Conv1d / nn.ConvTranspose1d
 * as called from modules.py:65,87,101,112,138,218,266-268,291,943,981 and lora.py:66-74.
 * ------------------------------------------------------------------------------------------- */
typedef struct cvflow_gemm_seg {
  int32_t a_map;     /* A source 0/1 */
  int32_t row_shift; /* row offset of this tap (rows outside the source read as zero) */
  int32_t a_col0;    /* first source column */
  int32_t nkb;       /* number of 64-column blocks */
} cvflow_gemm_seg;

typedef struct cvflow_gemm_desc {
  const void* A[2];     /* 16-bit [nbatch][a_rows][a_cols] */
  int32_t a_rows[2];
  int32_t a_cols[2];
  int64_t a_ld[2];      /* row stride, elements */
  int64_t a_bstride[2]; /* batch stride, elements */
  int32_t nbatch;
  int32_t dtype;        /* CVFLOW_DTYPE_* of A, W and 16-bit outputs */
  const void* W;        /* [N][Ktot] row-major 16-bit */
  int32_t N;
  int32_t Ktot;
  cvflow_gemm_seg seg[8];
  int32_t nseg;
  int32_t R;            /* tile rows per batch */
  int32_t rmul, roff;   /* output row = i*rmul + roff */
  int32_t out_rows;     /* rows per batch of the output */
  void* out;
  int32_t out_f32;
  int32_t transposed_out; /* fp32 out[(b*n_valid+n)*out_rows + row] */
  int64_t ldc;
  int32_t col_off;
  int32_t n_valid;
  float alpha;
  int32_t act;          /* 0 none, 1 gelu-tanh, 2 gelu-erf, 3 *gelu-tanh'(mul_src), 4 *gelu-erf'(mul_src) */
  const float* bias;
  void* aux_out;
  const void* mul_src;
  int64_t ld_aux;
  const float* rowmask;
  const float* resid;
  int64_t ldr;
  int64_t* dbg;         /* optional: 16 x int64 globaltimer stamps per CTA (profiling aid), else NULL */
  const float* ln_gamma; /* optional: LayerNorm fused into the epilogue (modules.py:349-375: h += to_out(..); x~ = LN(h)): with */
  const float* ln_beta;  /* N = n_valid = 256 and an fp32 output the tile owns whole rows; besides out (+ bias + resid) the
                          * epilogue writes x~ = (row - mean) rstd ln_gamma + ln_beta (eps 1e-5, two-pass statistics) as
                          * 16-bit to aux_out / ld_aux. Needs act = 0, rmul = 1, roff = 0, col_off = 0. Else NULL. */
  float* gn_part;       /* optional: GroupNorm partial statistics of the output taken in the epilogue (before the 16-bit
                         * rounding): {n, mean, M2} per (batch, 32-row slice, 32-channel group),
                         * [nbatch][4 * ceil(R / 128)][n_valid / 32][3] floats, Chan-mergeable (replaces the separate
                         * statistics pass of nn.GroupNorm after nn.Conv1d, modules.py:60-73); else NULL */
} cvflow_gemm_desc;

CVFLOW_API int cvflow_gemm(const cvflow_gemm_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Estimator handle: ConditionalDecoder.forward and its backward (reference modules.py:886-1106),
 * i.e. what `self.estimator(x, mask, mu, t, spks, cond)` runs in flow_model.py:116,174.
 * One handle = one device + one in-flight call (the caller serialises, as the reference's
 * TrtContextWrapper queue does). All kernels go to the given stream; no host synchronisation.
 * ------------------------------------------------------------------------------------------- */
typedef struct cvflow_estimator cvflow_estimator;

typedef struct cvflow_config {
  int32_t n_blocks;      /* transformer blocks per stage (CosyVoice-300M: 4) */
  int32_t n_mid;         /* mid stages (CosyVoice-300M: 12) */
  int32_t dtype;         /* CVFLOW_DTYPE_F16 / BF16 operand type (fp32 accumulate, fp32 residual stream) */
  int32_t gelu_erf;      /* 0 = tanh approximation (reference default), 1 = erf */
  int32_t lora_r;        /* LoRA rank on attn1.to_q/k/v, 0 = none */
  float lora_scaling;    /* alpha / r */
} cvflow_config;

CVFLOW_API int cvflow_create(const cvflow_config* cfg, cvflow_estimator** out);
CVFLOW_API void cvflow_destroy(cvflow_estimator* h);
/* Bind a device tensor by name (names are listed in DESIGN.md "weight binding").
 * dtype: 0 f16, 1 bf16, 2 f32. The memory stays owned by the caller and must outlive the handle. */
CVFLOW_API int cvflow_bind(cvflow_estimator* h, const char* name, void* ptr, int64_t numel, int32_t dtype);
CVFLOW_API int64_t cvflow_workspace_bytes(cvflow_estimator* h, int32_t B, int32_t T, int32_t training);
CVFLOW_API int cvflow_set_workspace(cvflow_estimator* h, void* ptr, int64_t bytes);
/* Rebuild the merged 16-bit q/k/v operands W + (alpha/r) B A from the bound fp32 masters
 * (lora.py:64-76 folded into the GEMM operand). Call after binding and after every optimiser step. */
CVFLOW_API int cvflow_lora_refresh(cvflow_estimator* h, void* stream);
/* Same pointer-table setup without the merge launch: for a second handle that shares the weight
 * images of another one (batch shards running concurrently on several streams). */
CVFLOW_API int cvflow_lora_prepare(cvflow_estimator* h, void* stream);
/* The light refresh a training step with lora_dropout > 0 needs after its optimiser update: only the 16-bit FACTOR images
 * (A_cat, B_blk and the LoRA parts of "<block>.w0d" / "<block>.w0t_ext"), not the 100 MB of folded W_eff images that such
 * a step never reads. Call cvflow_lora_refresh before the next eval() / lora_dropout = 0 forward. */
CVFLOW_API int cvflow_lora_refresh_factors(cvflow_estimator* h, void* stream);

typedef struct cvflow_estimator_io {
  const float* x;    int32_t x_nb;     /* [x_nb][80][T]; batch row b reads row b % x_nb */
  const float* mask; int32_t mask_nb;  /* [mask_nb][T] in {0,1} */
  const float* mu;   int32_t mu_nb;
  const float* t;    int32_t t_nb;     /* [t_nb] */
  const float* spks; int32_t spks_nb;  /* [spks_nb][80], may be NULL */
  const float* cond; int32_t cond_nb;  /* may be NULL */
  const float* keep;                   /* [B] multiplier of mu/spks/cond (CFG drop / uncond row), may be NULL */
  float* out;                          /* [B][80][T] fp32, zero where mask = 0 */
  int32_t B, T;
  int32_t iso_len;                     /* prompt_isolation_len (0 = off), modules.py:1034-1042 */
  int32_t training;                    /* 1: keep activations for cvflow_estimator_backward */
} cvflow_estimator_io;

CVFLOW_API int cvflow_estimator_forward(cvflow_estimator* h, const cvflow_estimator_io* io, void* stream);
/* dpred16: dL/dout as 16-bit token-major [B][T][128] (columns 80..127 zero), as cvflow_cfm_loss
 * writes it. Accumulates grad_scale * (*grad_scale_dev) * dL/d(lora_A|lora_B) into the bound
 * "<...>.grad" tensors; grad_scale_dev is an optional device scalar (NULL = 1), so an upstream
 * autograd factor never forces a host synchronisation. */
CVFLOW_API int cvflow_estimator_backward(cvflow_estimator* h, const void* dpred16, float grad_scale,
                                         const float* grad_scale_dev, void* stream);
/* The same backward, continued through the first transformer block's input, the first ResnetBlock1D and the
 * input pack (autograd of modules.py:1008-1019 and 89-94), so that the modules that PRODUCE the estimator inputs --
 * the Conformer encoder / length regulator behind mu, the speaker affine layer behind spks (flow_model.py:248-400;
 * LoRA targets of config.py:207-216) -- can be trained upstream. Outputs are fp32 in the input layouts
 * (dx, dmu, dcond [B][80][T]; dspks [B][80]), each nullable, overwritten (not accumulated), zero on masked frames,
 * including the CFG keep factor of the forward call and grad_scale * (*grad_scale_dev). */
typedef struct cvflow_input_grads {
  float* dx;
  float* dmu;
  float* dspks;
  float* dcond;
} cvflow_input_grads;
CVFLOW_API int cvflow_estimator_backward_inputs(cvflow_estimator* h, const void* dpred16, float grad_scale,
                                                const float* grad_scale_dev, const cvflow_input_grads* grads,
                                                void* stream);
/* lora_dropout of the q/k/v LoRA branches (reference lora.py:66-74: y = W x + s B (A drop(x)), one independent
 * nn.Dropout per LoRALinear). p = 0 (default): B A is folded into the GEMM operand. p > 0: training forwards apply an
 * independent keep mask per (attention block, projection, token, feature), drawn from a counter-based hash of `seed`
 * (advanced on the device at every training forward; the decisions are kept bit-packed, 96 B per token and block, for
 * the backward), scaled by 1/(1-p). p is quantised to 16 bits (a feature is dropped iff its hash field < round(p 2^16)).
 * Needs the extra operand images "<block>.w0d" ([1536][320] = [W0 | s B_cat]) and "<block>.w0t_ext" ([320][1536] =
 * [W0^T ; B_blk]) to be bound and kept current by the caller. debug_mask (optional, else NULL): explicit keep masks
 * [n_blocks][3][debug_rows][256] bytes (1 = keep) with debug_rows = B*T, for parity tests against the reference. */
CVFLOW_API int cvflow_set_lora_dropout(cvflow_estimator* h, float p, uint64_t seed, const uint8_t* debug_mask,
                                       int64_t debug_rows);
/* Data-parallel overlap: split the flat LoRA-gradient bucket into n chunks of whole attention blocks. Chunk k covers
 * blocks [first_block[k], first_block[k+1]) in execution order, which is also their order in the bucket
 * (first_block[0] = 0; a block holds 3 (r 256 + 512 r) floats). The backward visits blocks last to first; as soon as a
 * chunk's lowest block is done it reduces the chunk's split partials into the bucket and records events[k] (a
 * cudaEvent_t, timing disabled) on the backward's stream, so the caller can allreduce that slice on a side stream while
 * the rest of the backward runs (replaces DDP's bucketed gradient hooks; the reference trainer is single-device,
 * train_joint.py:349-352). n = 0 restores the single final reduction. */
CVFLOW_API int cvflow_set_grad_chunks(cvflow_estimator* h, int32_t n, const int32_t* first_block, void** events);
/* ---------------------------------------------------------------------------------------------
 * Euler-ODE solve with classifier-free guidance as ONE CUDA graph owned by the handle
 * (replaces ConditionalCFM.solve_euler's Python loop, flow_model.py:94-125; stands where the reference's TensorRT
 * estimator hook stands, cosyvoice/flow/flow_matching.py:125-152: raw device pointers, caller's stream).
 *   per step k:  d = estimator([x; x], mask, [mu; 0], t[k], [spks; 0], [cond; 0])      (batch 2 = cond / uncond)
 *                x += dt[k] * ((1 + cfg_rate) d[0] - cfg_rate d[1])
 * cvflow_solve_capture records n_steps of that on `stream` (must be a non-default stream; one un-captured warm-up
 * forward runs first and writes only d_scratch) and instantiates the graph, kept per (T, n_steps);
 * cvflow_solve_replay launches the one for (T, n_steps); cvflow_solve_release drops them all. The
 * buffers are the caller's and must stay alive and in place: x [1][80][T] (in: noise z, out: mel), mask [1][T],
 * mu / cond [1][80][T], spks [1][80] (spks / cond nullable), t / dt [n_steps] (the reference's accumulated time grid),
 * d_scratch [2][80][T]. Refill x / mu / ... between replays for a new utterance of the same T. The workspace for
 * (B = 2, T, training = 0) must be set; binding new weights or moving the workspace invalidates the captures
 * (release and capture again).
 * ------------------------------------------------------------------------------------------- */
CVFLOW_API int cvflow_solve_capture(cvflow_estimator* h, int32_t T, int32_t n_steps, float cfg_rate, float* x,
                                    const float* mask, const float* mu, const float* spks, const float* cond,
                                    const float* t, const float* dt, float* d_scratch, void* stream);
CVFLOW_API int cvflow_solve_replay(cvflow_estimator* h, int32_t T, int32_t n_steps, void* stream);
CVFLOW_API int cvflow_solve_release(cvflow_estimator* h);
/* time_mlp(SinusoidalPosEmb(320, scale 1000)(t)) -> out [B][1024]   (modules.py:27-57, the estimator's timestep path);
 * t has t_nb entries (row b uses t[b % t_nb]); scratch: B * 1344 floats. */
CVFLOW_API int cvflow_time_embed(cvflow_estimator* h, const float* t, int32_t t_nb, float* out, float* scratch, int32_t B,
                                 void* stream);
/* Read (out != NULL) and / or overwrite (in != NULL) the device-resident seed of the mask hash; host-synchronous.
 * Lets a caller rewind the dropout stream (e.g. after the warm-up executions that precede a CUDA-graph capture). */
CVFLOW_API int cvflow_lora_dropout_seed(cvflow_estimator* h, uint64_t* out, const uint64_t* in);
CVFLOW_API int64_t cvflow_launch_count(cvflow_estimator* h);
/* Measurement aid (bench.py roofline): when on, CUDA events bracket every tensor-core launch on
 * the launching stream. cvflow_profile_read synchronises the stream and sums per class
 * (0 gemm, 1 attention fwd, 2 attention bwd, 4 lora wgrad): milliseconds, launches, algorithmic FLOPs. */
CVFLOW_API int cvflow_set_profile(cvflow_estimator* h, int32_t on);
/* Profiling aid: attention-forward plans prepared after this call write 16 x int64 globaltimer phase
 * stamps per CTA into buf (NULL switches it off). */
CVFLOW_API int cvflow_debug_attention_stamps(void* buf);
/* Same for the fused feed-forward kernel: 64 x int64 clock stamps per CTA. */
CVFLOW_API int cvflow_debug_mlp_stamps(void* buf);
CVFLOW_API int cvflow_profile_read(cvflow_estimator* h, double* ms, int64_t* counts, double* flops, int32_t n);

/* ---------------------------------------------------------------------------------------------
 * attn1 self-attention on its own (reference Attention.forward, modules.py:253-293): the kernels the
 * estimator runs per transformer block, exposed for parity tests and reuse.
 *   qkv   16-bit [B][L][ldq], columns [0,512) q | [512,1024) k | [1024,1536) v, 8 heads x 64
 *   keymask fp32 [B][L] (0 = padded key, utils.py:103-109); iso_p = prompt-isolation boundary (0 = off)
 *   kmax_scratch int32 [cvflow_attention_scratch_ints(B, L)] (written: per sample 1 + last valid index, then key
 *   validity bit words, then the (tile pair, sample) work items ranked by cost: the persistent CTAs take them longest
 *   first, in zig-zag rounds, so ragged batches are load-balanced)
 *   o 16-bit [B][L][512]; lse fp32 [B][8][L] (base-2 log-sum-exp of the scaled scores, +inf on empty rows)
 * backward: dout 16-bit [B][L][512] (rows at or beyond a sample's last valid index are taken as zero, as they are
 * in the estimator where every consumer of padded rows is masked); delta_scratch fp32 [B][8][L];
 * dqkv 16-bit [B][L][1536].
 * ------------------------------------------------------------------------------------------- */
CVFLOW_API int64_t cvflow_attention_scratch_ints(int32_t B, int32_t L);
CVFLOW_API int cvflow_attention_forward(const void* qkv, int64_t ldq, int32_t B, int32_t L, int32_t dtype,
                                        const float* keymask, int32_t* kmax_scratch, int32_t iso_p, void* o, float* lse,
                                        void* stream);
CVFLOW_API int cvflow_attention_backward(const void* qkv, int64_t ldq, int32_t B, int32_t L, int32_t dtype,
                                         const float* keymask, int32_t* kmax_scratch, int32_t iso_p, const void* o,
                                         const float* lse, const void* dout, float* delta_scratch, void* dqkv,
                                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused FeedForward of a transformer block (reference modules.py:192-224,372-374), model 256 / hidden 1024:
 *   forward : out32[M][256] = resid32 + b2 + gelu(x16 W1^T + b1) W2^T, pre16[M][1024] = x16 W1^T + b1 (stash)
 *             W1 16-bit [1024][256], W2 16-bit [256][1024] (nn.Linear layouts)
 *   backward: dx16[M][256] = ((dy16 W2) o gelu'(pre16)) W1, given the transposed images W2^T [1024][256] and
 *             W1^T [256][1024]
 * ------------------------------------------------------------------------------------------- */
CVFLOW_API int cvflow_mlp_forward(const void* x16, const void* w1, const float* b1, const void* w2, const float* b2,
                                  const float* resid32, float* out32, void* pre16, int64_t M, int32_t dtype,
                                  int32_t gelu_erf, void* stream);
CVFLOW_API int cvflow_mlp_backward(const void* dy16, const void* w2_t, const void* pre16, const void* w1_t, void* dx16,
                                   int64_t M, int32_t dtype, int32_t gelu_erf, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CFM passes (reference flow_model.py:94-204)
 * ------------------------------------------------------------------------------------------- */
/* y = (1-(1-sigma_min) t) z + t x1            flow_model.py:154 */
CVFLOW_API int cvflow_cfm_prep(const float* x1, const float* z, const float* t, float* y, int32_t B, int32_t T,
                               float sigma_min, void* stream);
/* loss = sum(((pred-u) w)^2) / (sum(w) 80), u = x1-(1-sigma_min) z   flow_model.py:155,197-200
 * scal[0] = sum w, scal[1] = numerator, scal[2] = numerator / (scal[0] 80); partials: >= B*ceil(T/32) floats scratch;
 * dpred16 (nullable): loss_scale * dL/dpred * mask, 16-bit token-major [B][T][128]. */
CVFLOW_API int cvflow_cfm_loss(const float* pred, const float* x1, const float* z, const float* w, const float* mask,
                               float* scal, float* partials, void* dpred16, int32_t B, int32_t T, float sigma_min,
                               float loss_scale, int32_t dtype,
                               const float* wsum_dev /* optional: sum(w) of the WHOLE batch when this call covers one
                                                        shard of it (scal[0] is then copied from it), else NULL */,
                               void* stream);
/* x += dt[step] * ((1+cfg) d[0] - cfg d[1]) over n = 80*T elements   flow_model.py:117-119 */
CVFLOW_API int cvflow_euler_update(float* x, const float* d, const float* dt, int32_t step, float cfg_rate, int64_t n,
                                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimiser tail on the flat fp32 LoRA bucket (train_joint.py:198-226, gradient_clip_val 1.0)
 * ------------------------------------------------------------------------------------------- */
CVFLOW_API int cvflow_sumsq(const float* g, int64_t n, float* partials /* >= 296 */, float* out, void* stream);
CVFLOW_API int cvflow_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq,
                                 float grad_unscale, float max_norm, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, int32_t step, int32_t* found_inf,
                                 const float* hyper_dev /* optional {lr, 1-b1^t, sqrt(1-b2^t)} on device, NULL = use args */,
                                 void* stream);
/* Step counter, LR schedule and Adam bias corrections kept ON THE DEVICE so that a captured whole-step CUDA graph
 * needs no per-replay host write (replaces Lightning's LambdaLR.step + optimizer bookkeeping, train_joint.py:198-226):
 *   state[0] = optimiser steps applied (k), state[1] = steps skipped on a non-finite gradient norm;
 *   hyper[4] = {base_lr * lambda(k), 1-b1^(k+1), sqrt(1-b2^(k+1)), apply flag}; total_steps <= 0: constant lr.
 * sumsq2 (nullable): squared norm of a second bucket sharing the same global norm. Call between cvflow_sumsq and
 * cvflow_adamw_step(hyper_dev = hyper). */
CVFLOW_API int cvflow_optim_advance(int32_t* state, float* hyper, const float* sumsq, const float* sumsq2,
                                    float grad_unscale, float base_lr, int32_t warmup_steps, int32_t total_steps,
                                    float min_lr, float beta1, float beta2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The inputs of the path: what MaskedDiffWithXvec.forward prepares for compute_loss.
 * Stateless entry points (no estimator handle); fp32; all pointers device pointers unless stated.
 *
 * Length regulator = InterpolateRegulator.forward / .inference (modules.py:800-837):
 *   F.interpolate(mode='linear') of the projected encoder output, 4 x [Conv1d(80,80,3,pad 1) -> GroupNorm(1,80) -> Mish],
 *   Conv1d(80,80,1), * pad mask; the interpolation indices are bit-identical to at::upsample_linear1d's.
 * The weights are the module's frozen parameters in a kernel-friendly image:
 *   wf[l] forward image  [ci][tap][16][6]: wf[((ci*taps + k)*16 + co/5)*6 + co%5] = weight[co][ci][k], pad entries 0 (taps 3,3,3,3,1)
 *   wb[l] dgrad image, same layout with the roles of ci / co swapped and the taps reversed: weight[ci][co][taps-1-k]
 * ------------------------------------------------------------------------------------------- */
typedef struct cvflow_regulator_weights {
  const float* wf[5];
  const float* wb[5];
  const float* bias[5];
  const float* gamma[4];
  const float* beta[4];
} cvflow_regulator_weights;
typedef struct cvflow_regulator_io {
  const float* src;      /* encoder_proj output [B][n_src][80] */
  int32_t B, n_src, T, n_seg;
  int32_t seg[4][4];     /* {src0, src_n, dst0, dst_n}: F.interpolate(src[:, src0:src0+src_n], size=dst_n) -> frames [dst0, dst0+dst_n);
                          * training: one segment {0, n_src, 0, T} (modules.py:822); inference: prompt | head | mid | tail (:827-836) */
  const int32_t* lens;   /* optional [B]: frames >= lens[b] are zero (modules.py:821,824) */
  const int32_t* blind;  /* optional [B]: frames < blind[b] are zero (text-side blinding, flow_model.py:372-373) */
  float* out;            /* [B][T][80] like the module, or with channel_major = 1 [B][80][T] as compute_loss takes mu */
  int32_t channel_major;
  float* saved;          /* cvflow_regulator_saved_floats(B, T) floats; the backward reads them */
} cvflow_regulator_io;
CVFLOW_API int64_t cvflow_regulator_saved_floats(int32_t B, int32_t T);
CVFLOW_API int64_t cvflow_regulator_scratch_floats(int32_t B, int32_t T);
CVFLOW_API int cvflow_regulator_forward(const cvflow_regulator_weights* w, const cvflow_regulator_io* io, void* stream);
/* autograd of the above with respect to src (the weights are frozen under LoRA fine-tuning): dout in the layout of io->out,
 * dsrc [B][n_src][80]; io as passed to the forward (io->out unused); scratch: cvflow_regulator_scratch_floats(B, T) */
CVFLOW_API int cvflow_regulator_backward(const cvflow_regulator_weights* w, const cvflow_regulator_io* io, const float* dout,
                                         float* dsrc, float* scratch, void* stream);
/* x1 = ((feat - mel_mean) / mel_std)^T, cond (prompt frames from feat or cross, silence gap, zeros), mask, each [B][80|1][T],
 * from per-utterance descriptors desc[b] = {len, prompt frames, silence-gap frames, flags (bit 0: prompt from cross)}
 * (flow_model.py:266-269, 319-387: mel normalisation, the conds loop, make_pad_mask, the transposes). feat / cross are
 * the raw log-mel [B][T][80] / [B][cross_T][80] (cross nullable). */
CVFLOW_API int cvflow_path_inputs_pack(const float* feat, const float* cross, int32_t cross_T, const int32_t* desc,
                                       float mel_mean, float mel_std, float silence_normalised, float* x1, float* cond,
                                       float* mask, int32_t B, int32_t T, void* stream);
/* spks = Linear(F.normalize(embedding, dim=1))   (flow_model.py:297-298): e [B][K], W [N][K], out [B][N] */
CVFLOW_API int cvflow_spk_affine(const float* e, const float* W, const float* bias, float* out, int32_t B, int32_t K,
                                 int32_t N, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CVFLOW_H_ */
