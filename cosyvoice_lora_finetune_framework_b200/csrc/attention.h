// Shared pieces of the attn1 kernels (attention.cu forward, attention_bwd.cu backward).
//
// Work decomposition (all three tensor-core kernels): an *item* is one (batch b, head h, pair of
// adjacent 128-row tiles). A persistent CTA walks items; its two row warpgroups (WG0 / WG1, 128
// threads each, thread = row = TMEM lane) own the two row tiles of the item and share the column
// blocks that a producer warp streams through a shared-memory ring with TMA. One MMA-issuing warp
// per warpgroup drives that warpgroup's chain (score MMA -> row math -> output MMA), so while one
// warpgroup is in its exp-heavy phase the tensor core runs the other one's MMAs.
//
// Padding is skipped, not computed: kmax[b] = 1 + (last index with mask != 0) bounds both the rows
// and the columns that are touched (the `kinfo` array: per-sample extents followed by key-validity bit words). Rows >= kmax[b] are written as zeros (the reference computes
// garbage there and masks it downstream, modules.py:1046-1049,1104-1106), columns >= kmax[b] carry
// the -1e10 bias in the reference (utils.py:103-109), i.e. probability exactly 0 in fp32.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "common.cuh"

namespace cvflow {

struct AttnPlan {
  CUtensorMap tm_qkv;   // 16-bit [B][L][ldq] (q | k | v column ranges), box {64 cols, 64 rows, 1}
  CUtensorMap tm_do;    // 16-bit [B][L][512], same box (backward only)
  CUtensorMap tm_o;     // forward output [B][L][512], same box (TMA store)
  CUtensorMap tm_qkv32; // as tm_qkv with box {64 cols, 32 rows, 1} (96-key blocks of the dQ kernel)
  CUtensorMap tm_dqkv;  // backward output [B][L][1536], box {64, 64, 1} (TMA store)
  const void* o_ptr;    // destination tm_o / tm_dqkv was encoded for
  int B, L, bf16;
  int tpi;              // 128-row tiles per item: 2 (one per warpgroup) or, when that leaves SMs idle, 1 (warpgroup 0 only)
  long long* dbg;       // optional per-CTA globaltimer stamps (profiling aid)
  int early_kinfo;      // kinfo is NOT written by the kernel launched just before: its first read may precede the grid-dependency wait
};

static constexpr float kAttnScale = 0.125f;                               // d^-1/2, d = 64
static constexpr float kAttnScaleLog2 = 0.125f * 1.4426950408889634f;     // d^-1/2 * log2(e)

// thread layout shared by the kernels: warps 0-3 = WG0, 4-7 = WG1, 8 = TMA producer, 9 / 10 = MMA issuers
static constexpr int kAttnThreads = 352;

struct AttnItem {
  int b, h, tile0;      // the item's 128-row tiles: tile0 (warpgroup 0) and, with two tiles per item, tile0 + 1 (warpgroup 1)
  int kmax;             // valid extent of sample b (0 = nothing valid)
  int ext;              // kmax rounded up to 16 (MMA granularity), <= round16(L)
  bool act[2];          // does tile (2*pair + w) contain a valid row?
};
// Items are enumerated in DESCENDING COST: `ord` (written by attn_order_kernel after attn_kinfo_kernel) lists the
// (tile pair, sample) entries by decreasing cost = active 128-row tiles of the pair x valid keys of the sample
// (ord[2 r] = pair << 16 | b, ord[2 r + 1] = kmax[b]); item index = (entry rank, head). Together with the zig-zag
// assignment of attn_sched() a CTA that took a long item in one round gets a short one in the next
// (longest-processing-time pairing): with ragged batches the kernel time is the MEAN of a long and a short item, not the
// sum of two long ones. Entries whose tiles are all padding sort last and are skipped by every role.
__device__ __forceinline__ AttnItem attn_item(int it, int npairs, int tpi, int B, const int* __restrict__ ord) {
  AttnItem a;
  const int rank = it >> 3;
  const int2 e = *reinterpret_cast<const int2*>(ord + 2 * rank);
  a.h = it & 7;
  a.b = e.x & 0xffff;
  a.tile0 = (e.x >> 16) * tpi;
  a.kmax = e.y;
  a.ext = (a.kmax + 15) & ~15;
  a.act[0] = a.tile0 * 128 < a.kmax;
  a.act[1] = tpi == 2 && (a.tile0 + 1) * 128 < a.kmax;
  return a;
}
// k-th item of this CTA (-1: none): rounds alternate direction over the CTAs
__device__ __forceinline__ int attn_sched(int k, int n_items) {
  const int G = (int)gridDim.x, base = k * G;
  const int idx = base + ((k & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x);
  return (base < n_items && idx < n_items) ? idx : -1;
}
__device__ __forceinline__ const int* attn_order_ptr(const int* kinfo, int B, int L) {
  return kinfo + ((B + 3) & ~3) + B * (8 * ((L + 255) / 256));
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (16-bit, two K elements per 32-bit column, lane = row)
// was written with tcgen05.st by the row threads, so P / dS never travel through shared memory.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// named barrier of one 128-thread row warpgroup (ids 1 and 2; 0 is __syncthreads)
__device__ __forceinline__ void wg_bar_sync(int w) { asm volatile("bar.sync %0, 128;" ::"r"(w + 1) : "memory"); }

// prompt-isolation (modules.py:844-879): a row on one side of the boundary p only sees columns on the same side
__device__ __forceinline__ uint32_t attn_iso_word(uint32_t vw, int c0, int iso_p, bool row_below) {
  if (iso_p <= 0) return vw;
  const int nb = iso_p - c0;
  const uint32_t below = nb <= 0 ? 0u : (nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u));
  return vw & (row_below ? below : ~below);
}

// 3-input max (FMNMX3) and packed fp32x2 arithmetic (FFMA2 / FADD2): halve the issue slots of the row math
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// (x0, x1) * s + t for both lanes of a packed pair
__device__ __forceinline__ void ffma2(float& y0, float& y1, float x0, float x1, float s, float t) {
  asm("{\n\t.reg .b64 rx, rs, rt, ry;\n\t"
      "mov.b64 rx, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\tmov.b64 rt, {%5, %5};\n\t"
      "fma.rn.f32x2 ry, rx, rs, rt;\n\tmov.b64 {%0, %1}, ry;\n\t}"
      : "=f"(y0), "=f"(y1)
      : "f"(x0), "f"(x1), "f"(s), "f"(t));
}
__device__ __forceinline__ void fadd2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\t"
      "add.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(b0), "f"(b1));
}

// 2^x for x <= 0 on the FMA pipe (no MUFU): Cody-Waite split x = n + f, |f| <= 0.5, degree-4 polynomial for 2^f
// (relative error 5e-5, below the 16-bit rounding of the probabilities it feeds), exponent patched in with integer
// adds. The exp phases of the attention kernels are MUFU-bound (16 ex2 per clock per SM); sending a share of the
// elements down this path balances the two pipes. Both lanes of a pair are processed with packed fp32x2 arithmetic.
__device__ __forceinline__ void exp2_poly2(float& y0, float& y1, float x0, float x1) {
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  const f32x2 x = f2_pack(x0, x1);
  const f32x2 magic = f2_pack(12582912.f, 12582912.f), nmagic = f2_pack(-12582912.f, -12582912.f);
  const f32x2 t = f2_add(x, magic);                       // round-to-nearest integer n lands in the low mantissa bits
  const f32x2 f = f2_add(x, f2_mul(f2_add(t, nmagic), f2_pack(-1.f, -1.f)));
  f32x2 p = f2_fma(f2_pack(0.009618129f, 0.009618129f), f, f2_pack(0.05550411f, 0.05550411f));
  p = f2_fma(p, f, f2_pack(0.2402265f, 0.2402265f));
  p = f2_fma(p, f, f2_pack(0.6931472f, 0.6931472f));
  p = f2_fma(p, f, f2_pack(1.f, 1.f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
// pairs (of the 16 pairs of a 32-column chunk) whose exponentials take the polynomial path
#ifndef CVFLOW_POLY_PAIR_MASK
#define CVFLOW_POLY_PAIR_MASK 0x0000u   // measured: 19 % and 44 % polynomial shares are both slower (issue-bound), so off
#endif

int attn_num_sms();
// items of a launch and tiles per item: two tiles per item unless that leaves SMs without work (small batches, inference)
inline int attn_tiles_per_item(int B, int L) { return B * 8 * (((L + 127) / 128 + 1) / 2) >= attn_num_sms() ? 2 : 1; }
inline int attn_num_items(int B, int L, int tpi) { return B * 8 * (((L + 127) / 128 + tpi - 1) / tpi); }

}  // namespace cvflow
