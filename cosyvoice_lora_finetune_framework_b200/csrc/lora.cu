// LoRA kernels for the attn1 q/k/v projections (reference lora.py:18-76, 259-281).
//
// Forward / dgrad use the merged operand W_eff = W + (alpha/r) B A, rebuilt from the fp32 master
// copies by one batched launch per step (both the [N][K] image for y = x W_eff^T and the [K][N]
// image for dx = dy W_eff), so the tcgen05 GEMM sees a single 16-bit weight tile stream and the
// low-rank update costs no extra activation traffic. (Valid for lora_dropout == 0 and in eval();
// with lora_dropout > 0 in training the un-folded branch of lora_dropout.cu runs.)
// The wgrad kernel produces dA = s (dY B)^T x and dB = s dY^T (x A^T) per projection with a
// deterministic two-stage reduction (no atomics).
#include "kernels.h"
#include "common.cuh"
#include "gemm.h"
#include <string.h>

namespace cvflow {

#define LAUNCH_RET() do { cudaError_t e_ = cudaGetLastError(); return e_ == cudaSuccess ? 0 : -(int)e_; } while (0)

// grid (8 row-tiles of 64 output features, 3 projections, nblocks); 256 threads = one input feature k each.
// The K-major image is written row by row (512 B per warp store); the transposed image goes through a 16-bit shared-memory
// tile so that every row of it receives 64 contiguous values (128 B) instead of 2-element fragments.
static constexpr int kMergeRows = 64;
__global__ void __launch_bounds__(256) lora_merge_kernel(const LoraBlockPtrs* __restrict__ blocks, int r, int bf, int factors_only) {
  __shared__ __align__(16) uint16_t tile[kMergeRows][256 + 8];
  const LoraBlockPtrs& blk = blocks[blockIdx.z];
  const int p = blockIdx.y;
  const LoraLayerPtrs lp = blk.p[p];
  const int n0 = blockIdx.x * kMergeRows;
  const int k = threadIdx.x;
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = (lp.A && j < r) ? lp.A[j * 256 + k] : 0.f;
  uint16_t* weff = reinterpret_cast<uint16_t*>(blk.weff);
  uint16_t* weff_t = reinterpret_cast<uint16_t*>(blk.weff_t);
  // 16-bit factor images for the tensor-core wgrad: Acat16[p*r+j][k] = A_p[j][k],
  // Bblk16[p*r+j][p*512+n] = B_p[n][j] (block diagonal; the zero pattern is written once at bind time)
  if (lp.A && blk.acat16) {
    uint16_t* acat = reinterpret_cast<uint16_t*>(blk.acat16);
    uint16_t* bblk = reinterpret_cast<uint16_t*>(blk.bblk16);
    if (blockIdx.x == 0)
      for (int j = 0; j < r; ++j) acat[(p * r + j) * 256 + k] = f32_to_h16(a[j], bf);
    uint16_t* w0d = reinterpret_cast<uint16_t*>(blk.w0d);
    uint16_t* w0t = reinterpret_cast<uint16_t*>(blk.w0t_ext);
    for (int t = k; t < kMergeRows * r; t += 256) {
      const int i = t / r, j = t - i * r;
      const float bv = lp.Bm[(n0 + i) * r + j];
      const uint16_t b16 = f32_to_h16(bv, bf);
      bblk[(long)(p * r + j) * 1536 + p * 512 + n0 + i] = b16;
      if (w0t) w0t[(long)(256 + p * r + j) * 1536 + p * 512 + n0 + i] = b16;
      if (w0d) w0d[(long)(p * 512 + n0 + i) * 320 + 256 + p * r + j] = f32_to_h16(bv * lp.scaling, bf);
    }
  }
  if (factors_only) return;
  for (int i = 0; i < kMergeRows; ++i) {
    const int n = n0 + i;
    float w = lp.W[n * 256 + k];
    if (lp.A) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < r) acc += lp.Bm[n * r + j] * a[j];
      w += lp.scaling * acc;
    }
    const uint16_t h = f32_to_h16(w, bf);
    tile[i][k] = h;
    weff[(long)(p * 512 + n) * 256 + k] = h;
  }
  __syncthreads();
  // transposed image: row kk, columns p*512 + n0 .. +64 as 8 x 16-byte stores
  for (int idx = threadIdx.x; idx < 256 * 8; idx += 256) {
    const int kk = idx >> 3, u = idx & 7;
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (uint32_t)tile[8 * u + 2 * e][kk] | ((uint32_t)tile[8 * u + 2 * e + 1][kk] << 16);
    *reinterpret_cast<uint4*>(weff_t + (long)kk * 1536 + p * 512 + n0 + 8 * u) = make_uint4(v[0], v[1], v[2], v[3]);
  }
}
int launch_lora_merge(const LoraBlockPtrs* blocks_dev, int nblocks, int r, int bf16, int factors_only, cudaStream_t st) {
  if (r > 16) return -1;
  lora_merge_kernel<<<dim3(512 / kMergeRows, 3, nblocks), 256, 0, st>>>(blocks_dev, r, bf16, factors_only);
  LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// wgrad on tcgen05. With u = x A_cat^T and v = dY B_blk^T (two small GEMMs of the implicit-GEMM
// engine, 64 output columns each), the LoRA gradients are reductions over the token axis:
//     dB_all[1536][*] = dY^T u        dA_all^T[256][*] = x^T v
// Both operands of these contractions are token-major tensors whose contraction index is the row,
// i.e. MN-major UMMA operands: a TMA box {64 channels, 64 tokens} (128B swizzle) is consumed as
// is, with a_major = b_major = MN. CTA (s, t) reduces token split s for output row tile t
// (t < 12: 128 channels of dY, t >= 12: 128 channels of x) into a fp32 partial; a small kernel
// sums the splits in fixed order (deterministic) and accumulates into the fp32 gradient bucket.
// ------------------------------------------------------------------------------------------
struct alignas(64) WgradParams {
  CUtensorMap tmA[2];   // problem 0: dY [M][1536], problem 1: x [M][256]   (box {64, 64})
  CUtensorMap tmW[2];   // problem 0: u [M][64],    problem 1: v [M][64]
  long M;
  int rows_per_split, S, r, bf16;
  float* part_b;        // [S][1536][r]
  float* part_a;        // [S][256][64]
};
static constexpr int kWgStages = 4;
static constexpr int kWgStageBytes = 16384 + 8192;
static constexpr int kWgSmem = kWgStages * kWgStageBytes + 1024 + 256;

__global__ void __launch_bounds__(192, 2) lora_wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + kWgStages * kWgStageBytes;
  auto full_bar = [&](int s) { return bar + 8u * s; };
  auto empty_bar = [&](int s) { return bar + 8u * (kWgStages + s); };
  const uint32_t done_bar = bar + 8u * (2 * kWgStages);
  const uint32_t tmem_slot = bar + 8u * (2 * kWgStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, tile = blockIdx.y;
  const int prob = tile < 12 ? 0 : 1;
  const int mt = prob == 0 ? tile : tile - 12;
  const long tok_begin = (long)split * p.rows_per_split;
  long tok_end = tok_begin + p.rows_per_split;
  if (tok_end > p.M) tok_end = p.M;
  const int nkb = tok_end > tok_begin ? (int)((tok_end - tok_begin + 63) / 64) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), kWgStageBytes);
        const uint32_t sa = base + stage * kWgStageBytes;
        const int tok = (int)(tok_begin + (long)kb * 64);
        tma_load_3d(sa, &p.tmA[prob], full_bar(stage), mt * 128, tok, 0);
        tma_load_3d(sa + 8192, &p.tmA[prob], full_bar(stage), mt * 128 + 64, tok, 0);
        tma_load_3d(sa + 16384, &p.tmW[prob], full_bar(stage), 0, tok, 0);
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(p.bf16, 128, 64, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = base + stage * kWgStageBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 16 tokens per MMA = 16 rows of 128 bytes
          const uint64_t da = umma_desc_mnmajor_sw128(sa + k * 2048, 8192);
          const uint64_t db = umma_desc_mnmajor_sw128(sa + 16384 + k * 2048, 8192);
          umma_f16_ss(tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(done_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;   // channel within the tile
    if (nkb > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      if (nkb > 0) {
        __syncwarp();
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (prob == 0) {
        const int pj = mt >> 2;                      // projection of this 128-row tile
        float* dst = p.part_b + ((long)split * 1536 + mt * 128 + row) * p.r;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c * 32 + j - pj * p.r;
          if (col >= 0 && col < p.r) dst[col] = __uint_as_float(v[j]);
        }
      } else {
        float4* dst = reinterpret_cast<float4*>(p.part_a + ((long)split * 256 + mt * 128 + row) * 64 + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

// One launch for every attention block of the step: block i keeps its split partials in its own scratch region
// (scratch + i * stride), so the tensor-core reductions of the blocks can run on a side stream while the main
// backward chain continues, and the 64 small final reductions collapse into this one.
__global__ void __launch_bounds__(256) lora_wgrad_final_kernel(const LoraBlockPtrs* __restrict__ blocks, int nb, int nfull,
                                                               const float* __restrict__ scratch, long stride, int S_full,
                                                               int S_half, int r, float grad_scale,
                                                               const float* __restrict__ gs_dev, int bi0) {
  pdl_wait();
  pdl_launch();
  const int nbw = 1536 * r, na = 3 * r * 256;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= nbw + na) return;
  const int bi = bi0 + blockIdx.y;      // blocks [bi0, bi0 + gridDim.y) of the step
  const int S = (bi < nfull || bi >= nb - nfull) ? S_full : S_half;   // full-rate stages come first and last
  const float* part_b = scratch + (long)bi * stride;
  const float* part_a = part_b + (long)S * 1536 * r;
  if (gs_dev) grad_scale *= gs_dev[0];
  const LoraBlockPtrs& blk = blocks[bi];
  float s = 0.f;
  if (i < nbw) {
    for (int sp = 0; sp < S; ++sp) s += part_b[(long)sp * nbw + i];
    const int n_all = i / r, j = i - n_all * r;
    const int pj = n_all >> 9, n = n_all & 511;
    if (blk.p[pj].dB) blk.p[pj].dB[n * r + j] += s * grad_scale * blk.p[pj].scaling;
  } else {
    const int k2 = i - nbw;                 // (p, j, k)
    const int pj = k2 / (r * 256), rem = k2 - pj * r * 256;
    const int j = rem >> 8, k = rem & 255;
    for (int sp = 0; sp < S; ++sp) s += part_a[((long)sp * 256 + k) * 64 + pj * r + j];
    if (blk.p[pj].dA) blk.p[pj].dA[j * 256 + k] += s * grad_scale * blk.p[pj].scaling;
  }
}

static int wgrad_splits(long M) {
  long s = (M + 767) / 768;
  if (s < 1) s = 1;
  if (s > 32) s = 32;
  return (int)s;
}
int lora_wgrad_splits(long M) { return wgrad_splits(M); }
long lora_wgrad_scratch_floats(long M, int r) {
  const long S = wgrad_splits(M);
  return S * (1536L * r + 256L * 64) + 64;
}
int lora_wgrad_plan_bytes() { return (int)sizeof(WgradParams); }

int lora_wgrad_prepare(void* plan, const void* dqkv, const void* xn, const void* u16, long ld_u, const void* v16,
                       long ld_v, long M, int r, float* scratch, int bf16, char* err, int errlen) {
  WgradParams* p = reinterpret_cast<WgradParams*>(plan);
  memset(p, 0, sizeof(*p));
  p->M = M; p->r = r; p->bf16 = bf16;
  p->S = wgrad_splits(M);
  long rows = (M + p->S - 1) / p->S;
  p->rows_per_split = (int)((rows + 63) / 64 * 64);
  p->part_b = scratch;
  p->part_a = scratch + (long)p->S * 1536 * r;
  int rc = 0;
  rc |= tma_encode_3d(&p->tmA[0], dqkv, bf16, 1536, (uint64_t)M, 1, 1536 * 2, (uint64_t)M * 1536 * 2, 64, 64, 1);
  rc |= tma_encode_3d(&p->tmA[1], xn, bf16, 256, (uint64_t)M, 1, 256 * 2, (uint64_t)M * 256 * 2, 64, 64, 1);
  rc |= tma_encode_3d(&p->tmW[0], u16, bf16, 64, (uint64_t)M, 1, (uint64_t)ld_u * 2, (uint64_t)M * ld_u * 2, 64, 64, 1);
  rc |= tma_encode_3d(&p->tmW[1], v16, bf16, 64, (uint64_t)M, 1, (uint64_t)ld_v * 2, (uint64_t)M * ld_v * 2, 64, 64, 1);
  if (rc) { if (err) snprintf(err, errlen, "lora wgrad: cuTensorMapEncodeTiled failed"); return -1; }
  return 0;
}

float* lora_wgrad_plan_part_a(const void* plan, int* S) {
  const WgradParams* p = reinterpret_cast<const WgradParams*>(plan);
  if (S) *S = p->S;
  return p->part_a;
}

int lora_wgrad_launch_partial(const void* plan, cudaStream_t st) {
  const WgradParams* p = reinterpret_cast<const WgradParams*>(plan);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(lora_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    attr_done = true;
  }
  launch_pdl(lora_wgrad_tc_kernel, dim3(p->S, 14), 192, kWgSmem, st, *p);
  LAUNCH_RET();
}

int launch_lora_wgrad_final(const LoraBlockPtrs* blocks_dev, int nb, int nfull, const float* scratch, long stride, int S_full,
                            int S_half, int r, float grad_scale, const float* gs_dev, int bi0, int count, cudaStream_t st) {
  const int total = 1536 * r + 3 * r * 256;
  launch_pdl(lora_wgrad_final_kernel, dim3((total + 255) / 256, count), 256, 0, st, blocks_dev, nb, nfull, scratch, stride,
             S_full, S_half, r, grad_scale, gs_dev, bi0);
  LAUNCH_RET();
}


}  // namespace cvflow
