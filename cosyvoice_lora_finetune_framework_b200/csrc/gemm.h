// Host-side description of one implicit-GEMM launch of the tcgen05 engine (gemm.cu).
//
//   D[b, i, n] = epilogue( sum_seg sum_k A_seg[b, i + shift_seg, col0_seg + k] * W[n, kofs_seg + k] )
//
// Every dense contraction of the estimator maps onto it (SURVEY.md §2.3 K1,K3,K4,K6,K7 and their
// dgrads): nn.Linear (1 segment, 1 batch), Conv1d k=3 (3 row-shifted segments), the two phases
// of ConvTranspose1d / strided-conv dgrad (2 segments + interleaved output rows), stride-2 conv
// (segments on even/odd row views) and channel-concatenated inputs (segments on two sources).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvflow {

enum GemmAct { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2, ACT_MUL_GELU_TANH_GRAD = 3,
               ACT_MUL_GELU_ERF_GRAD = 4 };

struct GemmSeg {
  int a_map;      // which A source (0/1)
  int row_shift;  // row offset added to the tile's first row (may be negative; OOB rows read 0)
  int a_col0;     // first column of A consumed by this segment
  int nkb;        // number of 64-column k-blocks
};

struct GemmArgs {
  // A sources: 16-bit tensors viewed as [nbatch][a_rows][a_cols], row stride a_ld and batch stride
  // a_bstride in elements. Rows outside [0, a_rows) read as zero (TMA out-of-bounds fill).
  const void* A[2] = {nullptr, nullptr};
  int a_rows[2] = {0, 0};
  int a_cols[2] = {0, 0};
  long a_ld[2] = {0, 0};
  long a_bstride[2] = {0, 0};
  int nbatch = 1;
  // B operand: weights [N][Ktot] row-major, 16-bit, K contiguous (nn.Linear layout).
  const void* W = nullptr;
  bool w_static = false;   // W is not written by the kernel launched just before this one: its first tiles may be fetched before the grid-dependency wait
  int N = 0;       // rows of W (padded to a multiple of the N tile by the caller or zero-filled by TMA)
  int Ktot = 0;
  GemmSeg seg[8];
  int nseg = 0;
  int bf16 = 0;
  // iteration space: R output-tile rows per batch; tile row i maps to output row i*rmul+roff,
  // written iff that is < out_rows.
  int R = 0;
  int rmul = 1, roff = 0;
  int out_rows = 0;  // rows per batch of the output tensor
  // epilogue: x = acc*alpha + bias[n]; x = act(x); x *= rowmask[row]; x += resid[row][n]
  void* out = nullptr;
  int out_f32 = 0;
  long ldc = 0;
  int col_off = 0;
  int n_valid = 0;            // columns actually stored (<= N)
  int transposed_out = 0;     // fp32 out[(b*n_valid + n)*out_rows + row] (channel-major API edge)
  float alpha = 1.f;
  const float* bias = nullptr;
  int act = ACT_NONE;
  void* aux_out = nullptr;    // ACT_GELU_*: 16-bit pre-activation stash, ld = ld_aux
  const void* mul_src = nullptr;  // ACT_MUL_*: 16-bit pre-activation, ld = ld_aux
  long ld_aux = 0;
  const float* rowmask = nullptr;  // [nbatch*out_rows]
  const float* resid = nullptr;    // fp32 [nbatch*out_rows][ldr] (may alias out)
  long ldr = 0;
  // LayerNorm of the finished output rows fused into the epilogue (modules.py:349-375: h += to_out(...); x~ = LN3(h)):
  // with N = n_valid = 256 one 128x256 tile owns whole rows, so the epilogue writes the fp32 row (out, with bias and
  // residual) AND x~ = (row - mean) * rstd * ln_gamma + ln_beta as 16-bit to aux_out (ld_aux), statistics exact two-pass
  // over the row kept in TMEM. Needs out_f32, act == ACT_NONE, rmul == 1, roff == 0.
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  // GroupNorm statistics of the output (groups of 32 channels) taken in the epilogue, before the 16-bit rounding:
  // gn_part[((b * gn_nsplit + tile_in_batch * 4 + lane_quadrant) * (n_valid / 32) + group) * 3] = {n, mean, M2} over the
  // valid rows of that 32-row slice (Chan-mergeable partials; the consumer merges them). gn_nsplit = 4 * ceil(R / 128).
  float* gn_part = nullptr;
  long long* dbg = nullptr;        // optional per-CTA phase timestamps (8 x int64 per CTA), profiling aid
};

struct alignas(64) GemmParams {
  CUtensorMap tmA[2];
  CUtensorMap tmW;
  CUtensorMap tmOut;    // row-major output, box {32 cols, 32 rows, 1} (TMA store from the epilogue)
  CUtensorMap tmAux;    // pre-activation stash, same box
  int tma_out;          // 1: epilogue stores through TMA (CVFLOW_GEMM_EPI=0: the round-1 path, kept for A/B timing)
  int w_rows;           // rows of the W image (GemmArgs::N)
  const void* pf_ptr;   // weights of the NEXT GEMM of the plan (nullable): every CTA asks L2 for one slice of them
  unsigned pf_bytes;    // (cp.async.bulk.prefetch.L2), so the next launch streams its W operand from L2, not from HBM
  int w_prefetch;       // producer issues the first tile's W loads before griddepcontrol.wait (GemmArgs::w_static)
  int epi_direct;       // 1: epilogue stores as coalesced st.global.v4 after a transposition through the staging tile
  GemmSeg seg[8];
  int nseg, nkb_total;
  int bf16;
  int R, rmul, roff, out_rows, nbatch, tiles_per_batch;
  void* out; int out_f32; long ldc; int col_off; int n_valid; int transposed_out;
  float alpha; const float* bias; int act; void* aux_out; const void* mul_src; long ld_aux;
  const float* rowmask; const float* resid; long ldr;
  float* gn_part;
  const float* ln_gamma; const float* ln_beta;
  long long* dbg;
  const void* src_A[2]; const void* src_W;   // operand pointers the tensor maps were encoded for
  int block_n;   // 64, 128 or 256
  int deep;      // 1: BN = 128 with one CTA per SM and six pipeline stages (narrow-N, long-K shapes)
  int cluster;   // CTAs per cluster along N sharing (multicasting) the A tile: 1, 2 or 4
  int grid_x, grid_y;
};

// Encode tensor maps and pick the tile shape. Returns 0 or a negative error (message via err).
int gemm_prepare(const GemmArgs& a, GemmParams* p, char* err, int errlen);
int gemm_launch(const GemmParams& p, cudaStream_t stream);
// One-off: fetch cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency).
int tma_encode_3d(CUtensorMap* tm, const void* base, int bf16, uint64_t d0, uint64_t d1, uint64_t d2,
                  uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1,
                  uint32_t b2);
// dtype: 0 f16, 1 bf16, 2 f32; swizzle_bytes: 0, 32, 64 or 128
int tma_encode_3d_ex(CUtensorMap* tm, const void* base, int dtype, int swizzle_bytes, uint64_t d0, uint64_t d1,
                     uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1,
                     uint32_t b2);

int tma_encode_2d_ex(CUtensorMap* tm, const void* base, int dtype, int swizzle_bytes, uint64_t d0, uint64_t d1,
                     uint64_t stride1_bytes, uint32_t b0, uint32_t b1);

}  // namespace cvflow
