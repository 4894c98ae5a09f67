// Estimator handle: bound weights, workspace, memoised launch plans (see estimator.cu).
#pragma once
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include "gemm.h"
#include "kernels.h"

namespace cvflow {

char* error_buf();
int error_buf_len();

struct EstimatorConfig {
  int n_blocks = 4;     // transformer blocks per stage
  int n_mid = 12;       // mid stages
  int bf16 = 0;         // 16-bit operand type: 0 fp16, 1 bf16
  int gelu_erf = 0;     // 0: tanh approximation (reference modules.py:132), 1: erf
  int lora_r = 0;       // 0: no LoRA on q/k/v
  float lora_scaling = 1.f;
};

struct EstimatorIO {
  const float* x; int x_nb;        // [x_nb][80][T]
  const float* mask; int mask_nb;  // [mask_nb][T]
  const float* mu; int mu_nb;
  const float* t; int t_nb;        // [t_nb]
  const float* spks; int spks_nb;  // [spks_nb][80] (nullable)
  const float* cond; int cond_nb;  // nullable
  const float* keep;               // [B] CFG keep factors (nullable)
  float* out;                      // [B][80][T]
  int B, T, iso_len, training;
};

// optional outputs of the backward: dL/d(estimator inputs), fp32, same layouts as the inputs (each nullable)
struct InputGrads { float* dx; float* dmu; float* dspks; float* dcond; };

struct BoundTensor { void* ptr; long numel; int dtype; };  // dtype: 0 f16, 1 bf16, 2 f32

struct ResnetRec { std::string prefix; int cin, B, L; void *c1, *c2; float *st1, *st2; const float* mask; };
struct TBRec {
  std::string prefix; int lora_idx, B, L; long ldq; float* h0; void *x1, *qkv, *o; float* lse; float* h1; void* pre;
  const float* mask; const int* kmax; int iso_p;
  void* ud; int drop;   // lora_dropout > 0: u_d = (drop(x) A^T) stash [M][64], and the flag
  uint32_t* bits;       // lora_dropout > 0: keep decisions of the forward, [M][3][8 words]
};
struct StageRec { ResnetRec resnet; std::vector<TBRec> tbs; float* h_out; };
struct FinalRec { void* cf; float* st; };
struct BwdTemps {
  void *dpre, *dx, *dO; void* dqkv[2]; float* delta; void *dc, *da; float* wg_scratch; long wg_stride; void* dxe[2];
  float* wga_scratch;   // lora_dropout > 0: chunk partials of the masked dA reduction (shared by all blocks)
};

struct PlanKey {
  int B, T, training; uintptr_t ws;
  bool operator<(const PlanKey& o) const {
    if (B != o.B) return B < o.B;
    if (T != o.T) return T < o.T;
    if (training != o.training) return training < o.training;
    return ws < o.ws;
  }
};
struct Plan {
  std::vector<GemmParams> gemms;
  std::vector<std::vector<uint8_t>> attn, wgrads, mlps;
};

class Estimator {
 public:
  explicit Estimator(const EstimatorConfig& c);
  ~Estimator();
  int bind(const char* name, void* ptr, long numel, int dtype);
  void set_workspace(void* p, long bytes) { ws_ = p; ws_bytes_ = bytes; plans_.clear(); }
  long workspace_bytes(int B, int T, int training);
  // merge = true: every 16-bit LoRA operand image (folded W_eff in both layouts + factor images); false: factor images only
  int lora_refresh(cudaStream_t st, bool merge = true);
  int forward(const EstimatorIO& io, cudaStream_t st);
  int backward(const void* dpred16, float grad_scale, const float* grad_scale_dev, cudaStream_t st,
               const InputGrads* in_grads = nullptr);
  // lora_dropout p of the q/k/v LoRA branches in training forwards (0 = off: B A is folded into the GEMM operand);
  // dbg_mask: optional explicit keep masks [n_tbs][3][dbg_rows][256] bytes for parity tests (dbg_rows must be B*T)
  int set_lora_dropout(float p, unsigned long long seed, const uint8_t* dbg_mask, long dbg_rows);
  // Gradient chunks for overlapping the data-parallel allreduce with the rest of the backward: chunk k covers attention
  // blocks [lo[k], lo[k+1]) (execution order = position in the flat LoRA bucket; lo[0] = 0, the last chunk ends at n_tbs).
  // The backward finalises a chunk's gradients as soon as its lowest block is done (blocks are visited last to first)
  // and records events[k] on the stream; the caller makes a side stream wait on it and reduces that slice of the bucket.
  int set_grad_chunks(int n, const int* lo, cudaEvent_t* events);
  // N-step CFG Euler solve (flow_model.py:94-125) captured ONCE into a CUDA graph owned by the handle and replayed:
  // per step one batch-2 (cond / uncond) estimator forward + the guidance / Euler update. All pointers are the caller's
  // static device buffers; `stream` must be a capturing-capable (non-legacy) stream. The workspace for (B=2, T,
  // training=0) must have been set.
  int solve_capture(int T, int n_steps, float cfg_rate, float* x, const float* mask, const float* mu, const float* spks,
                    const float* cond, const float* t_arr, const float* dt_arr, float* d_scratch, cudaStream_t st);
  int solve_replay(int T, int n_steps, cudaStream_t st);
  void solve_release();     // drops every captured solve
  // t [t_nb] -> time_mlp(SinusoidalPosEmb(t)) [B][1024] (modules.py:27-57); scratch: B * (320 + 1024) floats
  int time_embed(const float* t, int t_nb, float* out, float* scratch, int B, cudaStream_t st);
  // device-resident seed of the mask hash (advanced by every training forward): read / overwrite (host-synchronous)
  int lora_dropout_seed(unsigned long long* out, const unsigned long long* in);
  long launches() const { return launches_; }
  void set_profile(int on);
  // class ids: 0 gemm, 1 attn_fwd, 2 attn_bwd, 3 layernorm (fwd + bwd, incl. the fused LoRA-dropout forms), 4 lora_wgrad,
  // 5 groupnorm + mish (apply, bwd). Work is FLOPs for the tensor-core classes, algorithmic bytes for 3 and 5.
  int profile_read(double* ms, long* counts, double* flops, int n);
  EstimatorConfig cfg;

 private:
  int n_resnets() const { return 4 + cfg.n_mid; }
  int n_tbs() const { return (4 + cfg.n_mid) * cfg.n_blocks; }
  void* get(const std::string& name, int dtype, long numel);
  bool has(const std::string& name) const;
  void* alloc(long bytes);
  int run_gemm(GemmArgs& a);
  int run_mlp(int backward, const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
              const float* resid, void* out, void* pre, long M);
  int iso_at(int L, int T, int iso_len) const;
  void for_each_tb(const std::function<void(const std::string&)>& f);
  int forward_impl(const EstimatorIO& io);
  int backward_impl(const void* dpred16, float grad_scale, const InputGrads* in_grads);
  int resnet_fwd(const std::string& P, const void* xin, long ld_in, int col0, int cin, int B, int L, const float* mask,
                 const float* tb, long tb_stride, float** h_out, ResnetRec* rec);
  int tb_fwd(const std::string& Q, int lora_idx, float* h0, int B, int L, const float* mask, const int* kmax, int iso_p,
             float** h_out, TBRec* rec, const std::string& Qnext);
  int stage_fwd(const std::string& S, int res_idx, const void* xin, long ld_in, int col0, int cin, int B, int L, int T,
                const float* mask, int iso_len, float** h_out);
  int tb_bwd(const TBRec& t, float* dh32, void* dh16, bool need_input_grad, float grad_scale, BwdTemps& tmp);
  int resnet_bwd(const ResnetRec& r, const float* dout32, const void* dout16, void* dxin16, BwdTemps& tmp);
  int stage_bwd(const StageRec& s, float* dh32, void* dh16, void* dxin16, bool first_stage, float grad_scale,
                BwdTemps& tmp);

  std::unordered_map<std::string, BoundTensor> bound_;
  std::map<PlanKey, Plan> plans_;
  Plan* plan_ = nullptr;
  void* ws_ = nullptr;
  long ws_bytes_ = 0, ws_off_ = 0, fwd_ws_end_ = 0;
  bool training_ = false;
  bool dry_ = false, missing_ = false, oom_ = false, have_fwd_ = false, lora_table_ready_ = false;
  int gemm_idx_ = 0, attn_idx_ = 0, tb_counter_ = 0, wg_idx_ = 0, mlp_idx_ = 0;
  bool ln_fuse_ = false;    // CVFLOW_LN_FUSE=1: LayerNorm in the row-owning GEMM epilogue (measured slower, see estimator.cu)
  void* next_x1_ = nullptr; // x1 of the next transformer block when the previous FF2 epilogue already wrote it
  bool fused_mlp_ = false;  // CVFLOW_FUSED_MLP=1 runs the feed-forward as one fused launch (mlp.cu) instead of two engine GEMMs
  long launches_ = 0;
  cudaStream_t stream_ = nullptr;
  LoraBlockPtrs* lora_table_dev_ = nullptr;
  // per-forward state
  std::vector<StageRec> stages_;
  FinalRec final_{};
  EstimatorIO last_io_{};
  const float* grad_scale_dev_ = nullptr;
  struct ProfRec { int cls; double flops; cudaEvent_t a, b; };
  bool profile_ = false;
  std::vector<ProfRec> prof_;
  std::vector<cudaEvent_t> ev_pool_;
  size_t ev_used_ = 0;
  cudaEvent_t ev_get();
  void prof_begin(int cls, double flops, cudaStream_t st = nullptr);
  void prof_end(cudaStream_t st = nullptr);
  // LoRA weight-gradient reductions are leaves of the backward graph: they run on a side stream (fork after the q/k/v
  // dgrad GEMM, join before the batched final reduction) over double-buffered dqkv / v operands
  cudaStream_t side_ = nullptr;
  cudaEvent_t ev_fork_ = nullptr, ev_done_[2] = {nullptr, nullptr};
  bool ev_done_valid_[2] = {false, false};
  bool wgrad_side_ = true;   // CVFLOW_WGRAD_SIDE=0 disables (see the constructor for the measurement)
  struct SolveGraph { cudaGraph_t graph; cudaGraphExec_t exec; };
  std::map<std::pair<int, int>, SolveGraph> solves_;   // (T, n_steps) -> captured solve
  float* solve_keep_dev_ = nullptr;     // {1, 0}: CFG keep factors of the (cond, uncond) rows
  std::vector<int> chunk_lo_;
  std::vector<cudaEvent_t> chunk_ev_;
  int finalize_blocks_from(int lora_idx, const BwdTemps& tmp, float grad_scale, long MT, long MH);
  float drop_p_ = 0.f;
  unsigned long long* drop_seed_dev_ = nullptr;
  const uint8_t* drop_dbg_ = nullptr;
  long drop_dbg_rows_ = 0, drop_mcap_ = 0;
  LoraDropSpec drop_spec(int blk) const;
  unsigned skip_ = 0;        // profiling build only (CVFLOW_PROFILING_BUILD): CVFLOW_SKIP bit mask, 1 attention, 2 layernorm, 4 groupnorm, 8 wgrad, 16 gemm
#ifdef CVFLOW_PROFILING_BUILD
  static unsigned kSkip(unsigned m) { return m; }
#else
  static constexpr unsigned kSkip(unsigned) { return 0u; }   // product build: every launch always runs
#endif
  // Buffers that do not outlive their block are allocated ONCE per forward and reused by every block, so that they stay
  // dirty-in-L2 and are overwritten in place instead of streaming to HBM: in training the FF hidden activation g16, the
  // LayerNorm-3 output and the resnets' a1 / r (nothing in the backward reads them); in eval() every block tensor
  // (two ping-pong sets, block j+1 reads block j's h2 while writing its own).
  struct TbSet { void *x1, *qkv, *o; float *lse, *h1, *h2; };
  void *scr_g16_ = nullptr, *scr_x3_ = nullptr, *scr_a1_ = nullptr, *scr_r_ = nullptr, *scr_c1_ = nullptr, *scr_c2_ = nullptr;
  float* scr_rh_ = nullptr;
  TbSet scr_tb_[2] = {};
  float *tb_all_ = nullptr, *gn_partials_ = nullptr, *mask1_ = nullptr, *mask2_ = nullptr;
  int *kmax1_ = nullptr, *kmax2_ = nullptr;
  void *cat0_ = nullptr, *cat1_ = nullptr;
};

}  // namespace cvflow
