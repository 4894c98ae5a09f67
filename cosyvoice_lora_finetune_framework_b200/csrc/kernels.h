// Host launchers of the non-GEMM kernels (HBM-bound passes, attention, LoRA, optimiser).
// All pointers are device pointers; every launcher is asynchronous on `st` and returns the
// cudaError_t of the launch as a negative int (0 = ok).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvflow {

// ---- elementwise.cu ------------------------------------------------------------------------
// [x | mu*keep | spks*keep | cond*keep] * mask -> token-major 16-bit [B][T][320]
// (replaces einops.pack / repeat at modules.py:1008-1014 and the x*mask of the first Block1D).
// Source batch rows are taken modulo their own batch count (CFG packing of solve_euler,
// flow_model.py:108-113, without materialising the batch-2 copies).
int launch_pack_inputs(const float* x, int x_nb, const float* mu, int mu_nb, const float* spks, int spks_nb,
                       const float* cond, int cond_nb, const float* mask, int mask_nb, const float* keep,
                       void* out, int B, int T, int bf16, cudaStream_t st);
// Backward of the pack: 16-bit dL/d(packed) [B][T][320] (already row-masked) -> fp32 channel-major dL/dx, dL/dmu,
// dL/dcond [B][80][T] and dL/dspks [B][80] (sum over frames), each nullable; mu / spks / cond are multiplied by keep[b];
// everything by scale * (*gs_dev). spk_part: scratch of B * ceil(T/32) * 80 floats (autograd of modules.py:1008-1014).
int launch_unpack_input_grads(const void* g16, const float* keep, float scale, const float* gs_dev, float* dx,
                              float* dmu, float* dspks, float* dcond, float* spk_part, int B, int T, int bf16,
                              cudaStream_t st);
// mask2[b][j] = mask[b % mask_nb][2j]   (modules.py:1049, masks.append(mask[:, :, ::2]))
int launch_mask_down(const float* mask, int mask_nb, float* mask1, float* mask2, int B, int T, int T2,
                     cudaStream_t st);
// sinusoidal embedding, scale 1000, dim 320 (modules.py:27-42)
int launch_sinus_embed(const float* t, int t_nb, float* out, int B, cudaStream_t st);
// y[b][n] = out_act(sum_k in_act(x[b][k]) W[n][k] + bias[n]); act: 0 none, 1 silu, 2 mish
int launch_small_linear(const float* x, const float* W, const float* bias, float* y, int B, int K, int N,
                        int in_act, int out_act, cudaStream_t st);
// 16-bit masked copy of the fp32 residual stream into (dst, ldc, col_off)
int launch_stage_out(const float* h, const float* rowmask, void* dst, long ldc, int col_off, long M, int bf16,
                     cudaStream_t st);
// fp32 [M][256] (+)= 16-bit src[M][ld_src, col_off..+256] * rowmask   (skip / concat gradient routing)
int launch_grad_route(const void* src, long ld_src, int col_off, const float* rowmask, float* dst, int accumulate,
                      void* dst16, long M, int bf16, cudaStream_t st);
// y = (1-(1-sigma)t) z + t x1   (flow_model.py:154)
int launch_cfm_prep(const float* x1, const float* z, const float* t, float* y, int B, int T, float sigma_min,
                    cudaStream_t st);
// w-sum, masked loss and dL/dpred (flow_model.py:197-200); dpred is 16-bit token-major [B][T][128]
// (cols 80..127 zero) scaled by loss_scale. scal[0]=sum w, scal[1]=loss numerator, scal[2]=loss.
int launch_cfm_loss(const float* pred, const float* x1, const float* z, const float* w, const float* mask,
                    float* scal, float* partials, void* dpred, int B, int T, float sigma_min, float loss_scale,
                    int bf16, const float* wsum_dev, cudaStream_t st);
// x += dt * ((1+g) d[0] - g d[1])   (flow_model.py:117-119)
int launch_euler_update(float* x, const float* d, const float* dt_arr, int step, float cfg_rate, long n,
                        cudaStream_t st);

// ---- norm.cu -------------------------------------------------------------------------------
int launch_layernorm_fwd(const float* h, const float* gamma, const float* beta, void* out, long M, int bf16,
                         cudaStream_t st);
// dh = dres + LN'(dxn16; h_in); also writes the 16-bit copy dh16 (nullable)
int launch_layernorm_bwd(const void* dxn16, long ld_dxn, const float* h_in, const float* gamma, const float* dres, float* dh,
                         void* dh16, long M, int bf16, cudaStream_t st);
// GroupNorm(8 groups of 32 channels, C=256) over the padded extent L of [B][L][256]. Forward statistics arrive as
// partials {n, mean, M2} [B][gn_fwd_splits(L)][8][3] written by the epilogue of the conv GEMM (GemmArgs::gn_part).
int gn_num_splits(int B, int L);     // backward split count: partials [B][gn_num_splits][8][2]
int gn_fwd_splits(int L);
// mode 0: out16 = (mish(gn(c)) + tb[b][c]) * mask        (tb nullable)
// mode 1: out32 = mish(gn(c)) * mask + add16             (resnet output, fp32 residual stream)
int launch_gn_apply(const void* c16, const float* partials, float* stats, const float* gamma, const float* beta,
                    const float* mask, const float* tb, long tb_stride, const void* add16, void* out, int mode, int B, int L,
                    int bf16, cudaStream_t st);
// dy: fp32 (dy_f32=1) or 16-bit; masked by rowmask. Produces dc16.
int launch_gn_bwd(const void* dy, int dy_f32, const void* c16, const float* stats, const float* gamma,
                  const float* beta, const float* mask, float* partials, void* dc16, int B, int L, int bf16,
                  cudaStream_t st);

// ---- attention.cu --------------------------------------------------------------------------
// qkv 16-bit [B][L][1536] (q | k | v, 8 heads x 64) -> o 16-bit [B][L][512], lse fp32 [B][8][L]
// keymask fp32 [B][L]; iso_p: prompt-isolation boundary at this resolution (0 = off).
struct AttnPlan;  // holds the encoded tensor maps
int attn_plan_bytes();
void attn_plan_set_early_kinfo(void* plan, int on);   // kinfo is not produced by the launch right before the attention kernels
void attn_set_debug_buffer(void* p);   // profiling aid: 16 x int64 globaltimer stamps per CTA of the next forward plans
int attn_fwd_prepare(void* plan, const void* qkv, long ldq, int B, int L, int bf16, char* err, int errlen);
// kinfo (attn_kinfo_ints(B, L) ints): [b] = kmax[b] = 1 + last index with keymask[b][.] != 0 -- rows and keys beyond
// it are skipped (rows written as zeros) -- followed by the per-sample key validity bit words.
long attn_kinfo_ints(int B, int L);
int launch_attn_kinfo(const float* keymask, int B, int L, int* kinfo, cudaStream_t st);
int attn_fwd_launch(void* plan, const int* kinfo, int iso_p, void* o, float* lse, cudaStream_t st);
int attn_bwd_prepare(void* plan, const void* qkv, long ldq, const void* dout, int B, int L, int bf16, char* err,
                     int errlen);
// dout rows at or beyond kinfo[b] are taken as zero (padding rows; every consumer of them is masked in the estimator)
int attn_bwd_launch(void* plan, const void* dout, const int* kinfo, int iso_p, const void* o, const float* lse,
                    float* delta, void* dqkv, cudaStream_t st);

// ---- mlp.cu --------------------------------------------------------------------------------
// Fused FeedForward (modules.py:192-224) + residual, and its backward, for hidden 1024 / model 256:
//   forward : out32[M][256] = resid + b2 + gelu(x16 W1^T + b1) W2^T; pre16[M][1024] = x16 W1^T + b1 (stash)
//   backward: out16[M][256] = ((x16 W1'^T) o gelu'(pre16)) W2'^T with W1' = W2^T image [1024][256], W2' = W1^T image [256][1024]
int mlp_plan_bytes();
int mlp_prepare(void* plan, int backward, const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                const float* resid, void* out, void* pre, long M, int bf16, int gelu_erf, char* err, int errlen);
int mlp_launch(const void* plan, cudaStream_t st);
void mlp_set_debug_buffer(void* p);   // profiling aid: 64 clock64 stamps per CTA for plans prepared afterwards

// ---- lora.cu -------------------------------------------------------------------------------
struct LoraLayerPtrs {       // one q/k/v projection of one attention block
  const float* W;            // frozen [512][256] fp32
  const float* A;            // [r][256] fp32 (nullable -> plain cast)
  const float* Bm;           // [512][r] fp32
  float* dA;                 // grads (nullable)
  float* dB;
  float scaling;
};
struct LoraBlockPtrs {       // one attention block: q, k, v + the merged 16-bit images
  LoraLayerPtrs p[3];
  void* weff;                // [1536][256]
  void* weff_t;              // [256][1536]
  void* acat16;              // [64][256]  rows p*r+j = A_p[j]      (nullable)
  void* bblk16;              // [64][1536] rows p*r+j, cols p*512+n = B_p[n][j]
  void* w0d;                 // lora_dropout > 0: [1536][320] = [W0 | s B_cat] (columns 256 + p*r+j of row p*512+n = s B_p[n][j]), nullable
  void* w0t_ext;             // lora_dropout > 0: [320][1536] = [W0^T ; B_blk] (rows 256 + p*r+j as bblk16), nullable
};
// All 16-bit LoRA operand images of every block in one launch. factors_only = 0: W_eff = W + s B A in both layouts plus
// the factor images (A_cat, B_blk, LoRA parts of w0d / w0t_ext); factors_only = 1: the factor images alone (what a
// training step with lora_dropout > 0 consumes: it never reads the folded W_eff).
int launch_lora_merge(const LoraBlockPtrs* blocks_dev, int nblocks, int r, int bf16, int factors_only, cudaStream_t st);
// dA, dB of the three projections of one block: tcgen05 reductions over the token axis of
// dY^T u and x^T v (u = x A_cat^T, v = dY B_blk^T are produced by two engine GEMMs, [M][64] each).
int lora_wgrad_plan_bytes();
long lora_wgrad_scratch_floats(long M, int r);
int lora_wgrad_prepare(void* plan, const void* dqkv, const void* xn, const void* u16, long ld_u, const void* v16,
                       long ld_v, long M, int r, float* scratch, int bf16, char* err, int errlen);
// per block: split partial reductions into its scratch region (any stream) ...
int lora_wgrad_splits(long M);
int lora_wgrad_launch_partial(const void* plan, cudaStream_t st);
// ... and one final reduction for all nb blocks (block i: scratch + i*stride; the first / last nfull blocks are full-rate)
int launch_lora_wgrad_final(const LoraBlockPtrs* blocks_dev, int nb, int nfull, const float* scratch, long stride, int S_full,
                            int S_half, int r, float grad_scale, const float* gs_dev, int bi0, int count, cudaStream_t st);

// ---- lora_dropout.cu -------------------------------------------------------------------------
// lora_dropout > 0: independent keep masks per (attention block, projection, token, feature) from a counter-based hash of
// a device-resident seed, drawn once by the forward and kept bit-packed (24 words per token) for the backward; or an
// explicit byte mask for the parity tests.
struct LoraDropSpec {
  const unsigned long long* seed;  // device scalar
  const uint8_t* dbg;              // optional explicit keep mask [nblocks][3][mcap][256] (1 = keep), else NULL
  long mcap;                       // rows of the counter space per (block, projection): B * T
  int blk;                         // attention block index
  unsigned thr16;                  // drop iff the feature's 16-bit hash field < thr16 (= round(p * 2^16))
  float inv_keep;                  // 1 / (1 - p)
};
int launch_lora_seed_bump(unsigned long long* seed, cudaStream_t st);
// x~ = LayerNorm(h) (16-bit), keep bits [M][3][8 words], u_d[M][64] (16-bit, columns >= 3r zero) =
// 1/(1-p) (keep_p o x~) A_p^T for the three projections -- one pass over the residual stream
int launch_ln_lora_drop_fwd(const float* h, const float* gamma, const float* beta, const void* acat16, void* x16, void* ud16,
                            uint32_t* bits, long M, int r, int bf16, const LoraDropSpec& d, cudaStream_t st);
// dxe[M][320] (columns [0,256) dx, [256,256+3r) v): dx~ = dx + s/(1-p) sum_p keep_p o (v_p A_p), then the LayerNorm backward
// of dx~ (h_in, gamma) + residual gradient dres -> dh (fp32) / dh16
int launch_ln_lora_drop_bwd(const void* dxe16, const void* acat16, const uint32_t* bits, const float* h_in, const float* gamma,
                            const float* dres, float* dh, void* dh16, long M, int r, float scaling, float inv_keep, int bf16,
                            cudaStream_t st);
// replaces the x~^T v partials of lora_wgrad_launch_partial (run after it) by the masked ones;
// scratch: lora_wgrad_a_dropout_scratch_floats(M) floats, free to reuse once the call's kernels have run
long lora_wgrad_a_dropout_scratch_floats(long M);
int launch_lora_wgrad_a_drop(const void* x16, const void* v16, long ld_v, const uint32_t* bits, float* scratch, float* part_a,
                             int S, long M, int r, float inv_keep, int bf16, cudaStream_t st);
float* lora_wgrad_plan_part_a(const void* plan, int* S);

// ---- optim.cu ------------------------------------------------------------------------------
// Fused global-norm clip + AdamW over a flat fp32 bucket (train_joint.py:198-226,353-355).
int launch_sumsq(const float* g, long n, float* partials, float* out_sumsq, cudaStream_t st);
int launch_adamw(float* p, const float* g, float* m, float* v, long n, const float* sumsq, float grad_unscale,
                 float max_norm, float lr, float beta1, float beta2, float eps, float wd, int step,
                 int* found_inf, const float* hyper_dev, cudaStream_t st);
// device-resident step counter / LR schedule / Adam bias corrections (see optim.cu)
int launch_optim_advance(int* state, float* hyper, const float* sumsq, const float* sumsq2, float grad_unscale, float base_lr,
                         int warmup_steps, int total_steps, float min_lr, float beta1, float beta2, cudaStream_t st);

// ---- regulator.cu: the inputs of the path (length regulator, speaker affine, conditioning pack) -----------------
struct RegSeg { int src0, srcn, dst0, dstn; };   // F.interpolate(src[:, src0:src0+srcn], size=dstn, mode='linear') -> frames [dst0, dst0+dstn)
struct RegulatorWeights {          // fp32 device pointers; conv images [ci][tap][16][6] (see regulator.cu), wb = dgrad images
  const float* wf[5];
  const float* wb[5];
  const float* bias[5];
  const float* gamma[4];
  const float* beta[4];
};
struct RegulatorIO {
  const float* src;                // [B][n_src][80]
  int B, n_src, T, n_seg;
  RegSeg seg[4];
  const int* lens;                 // nullable: frames >= lens[b] are zero
  const int* blind;                // nullable: frames < blind[b] are zero
  float* out;                      // [B][T][80] or, channel_major, [B][80][T]
  int channel_major;
  float* saved;                    // regulator_saved_floats(B, T): raw layer outputs + GroupNorm partials (kept for the backward)
};
long regulator_saved_floats(int B, int T);
long regulator_scratch_floats(int B, int T);
int launch_regulator_forward(const RegulatorWeights& w, const RegulatorIO& io, cudaStream_t st);
// dout in the layout of io.out (masked / blinded here); dsrc [B][n_src][80]; scratch: regulator_scratch_floats(B, T)
int launch_regulator_backward(const RegulatorWeights& w, const RegulatorIO& io, const float* dout, float* dsrc, float* scratch,
                              cudaStream_t st);
// desc[b] = {len, prompt frames, silence-gap frames, flags (bit 0: prompt from cross)}; feat / cross raw log-mel [B][T][80]
int launch_path_inputs_pack(const float* feat, const float* cross, int cross_T, const int* desc, float mel_mean, float mel_std,
                            float silence, float* x1, float* cond, float* mask, int B, int T, cudaStream_t st);
int launch_spk_affine(const float* e, const float* W, const float* bias, float* out, int B, int K, int N, cudaStream_t st);

}  // namespace cvflow
