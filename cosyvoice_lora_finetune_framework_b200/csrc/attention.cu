// Flash-style attn1 self-attention on tcgen05 (sm_100a), forward and backward.
//
// Reference: Attention.forward, modules.py:253-293 -- softmax(q k^T * d^-1/2 + bias) v with
// 8 heads x 64, bias = -1e10 on padded keys (utils.py:103-109) plus the optional prompt-isolation
// block mask (modules.py:844-879,1034-1042). The [B,L,L] bias and the [B,8,L,L] scores are never
// materialised: the key mask comes from the [B,L] float mask and the isolation boundary from one
// integer.
//
// Layout of one CTA (forward and dQ kernels): 128 query rows of one (batch, head); sixteen softmax
// warps (warp w: TMEM lane quadrant w & 3 = 32 query rows, column group w >> 2), one MMA/TMA warp.
//   S  = Q K^T   : A = Q [128 x 64] K-major, B = K block [128 keys x 64] K-major  -> TMEM cols [0,128)
//   P  = softmax : tcgen05.ld row -> registers -> 16-bit, written to smem in the K-major
//                  128B-swizzled operand layout (row = query, K = key)
//   O' = P V     : A = P, B = V block [128 keys x 64] as MN-major operand            -> TMEM cols [128,192)
// The running max / sum / output row live in registers of the row's thread (no cross-lane
// shuffles: one thread sees its whole score row).
#include "kernels.h"
#include "gemm.h"
#include "common.cuh"
#include <string.h>

namespace cvflow {

struct AttnPlan {
  CUtensorMap tm_qkv;   // 16-bit [B][L][1536], box {64, 128, 1}
  CUtensorMap tm_do;    // 16-bit [B][L][512],  box {64, 128, 1} (backward only)
  int B, L, bf16;
  long long* dbg;       // optional per-CTA globaltimer stamps (profiling aid)
};
static long long* g_attn_dbg = nullptr;
void attn_set_debug_buffer(void* p) { g_attn_dbg = reinterpret_cast<long long*>(p); }
int attn_plan_bytes() { return (int)sizeof(AttnPlan); }

static constexpr int kAQ = 128;   // query rows per CTA
static constexpr int kAK = 128;   // keys per block
static constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;  // d^-1/2 * log2(e)

// named barrier among the 512 softmax threads only
__device__ __forceinline__ void softmax_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// forward. 16 softmax warps: warp w owns TMEM lane quadrant w & 3 (query rows) and column group
// w >> 2 (32 of the 128 keys of a block / 16 of the 64 output columns), so every scheduler has
// four warps to interleave; warp 16 drives TMA and issues the MMAs.
// ------------------------------------------------------------------------------------------
struct AttnFwdSmem {
  static constexpr int kQ = 0;
  static constexpr int kK0 = 16384;
  static constexpr int kK1 = 32768;
  static constexpr int kV = 49152;
  static constexpr int kP = 65536;          // 2 x 16 KB
  static constexpr int kBar = 98304;        // barriers, then row-max / row-sum exchange
  static constexpr int kMx = kBar + 128;    // float [2][128][4]
  static constexpr int kBytes = kMx + 4096 + 1024;
};
static constexpr int kFwdThreads = 544;

__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ AttnPlan plan, const float* __restrict__ keymask, int iso_p,
                uint16_t* __restrict__ o_out, float* __restrict__ lse_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + AttnFwdSmem::kBar;
  const uint32_t bar_q = bar, bar_k0 = bar + 8, bar_k1 = bar + 16, bar_v = bar + 24, bar_s = bar + 32,
                 bar_p = bar + 40, bar_o = bar + 48, tmem_slot = bar + 56;
  float* mx = reinterpret_cast<float*>(gbase + AttnFwdSmem::kMx);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAQ, h = blockIdx.y, b = blockIdx.z;
  const int L = plan.L, bf = plan.bf16;
  const int nkb = (L + kAK - 1) / kAK;
  long long* dbg = plan.dbg ? plan.dbg + ((long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 : nullptr;
  auto stamp = [&](int k) {
    if (dbg) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[k] = t; }
  };
  if (threadIdx.x == 0) stamp(0);

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_k0, 1); mbar_init(bar_k1, 1); mbar_init(bar_v, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 512); mbar_init(bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&plan.tm_qkv);
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tmem_S = tmem, tmem_O = tmem + 128;
  if (threadIdx.x == 0) stamp(1);
  pdl_wait();
  pdl_launch();
  if (threadIdx.x == 0) stamp(2);

  if (warp == 16) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(bf, 128, 128, 0, 0);
      const uint32_t idesc_o = umma_idesc_f16(bf, 128, 64, 0, 1);
      mbar_expect_tx(bar_q, 16384);
      tma_load_3d(base + AttnFwdSmem::kQ, &plan.tm_qkv, bar_q, h * 64, q0, b);
      mbar_expect_tx(bar_k0, 16384);
      tma_load_3d(base + AttnFwdSmem::kK0, &plan.tm_qkv, bar_k0, 512 + h * 64, 0, b);
      mbar_expect_tx(bar_v, 16384);
      tma_load_3d(base + AttnFwdSmem::kV, &plan.tm_qkv, bar_v, 1024 + h * 64, 0, b);
      mbar_wait(bar_q, 0);
      for (int i = 0; i < nkb; ++i) {
        const uint32_t sK = base + ((i & 1) ? AttnFwdSmem::kK1 : AttnFwdSmem::kK0);
        if (i + 1 < nkb) {  // prefetch next K block (its buffer was last read by S(i-1), long complete)
          const uint32_t bk = ((i + 1) & 1) ? bar_k1 : bar_k0;
          mbar_expect_tx(bk, 16384);
          tma_load_3d(base + (((i + 1) & 1) ? AttnFwdSmem::kK1 : AttnFwdSmem::kK0), &plan.tm_qkv, bk,
                      512 + h * 64, (i + 1) * kAK, b);
        }
        mbar_wait((i & 1) ? bar_k1 : bar_k0, (uint32_t)((i >> 1) & 1));
        tc_fence_after();
        {
          const uint64_t dq = umma_desc_kmajor_sw128(base + AttnFwdSmem::kQ);
          const uint64_t dk = umma_desc_kmajor_sw128(sK);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
        }
        umma_commit(bar_s);
        mbar_wait(bar_p, (uint32_t)(i & 1));
        mbar_wait(bar_v, (uint32_t)(i & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t dp = umma_desc_kmajor_sw128(base + AttnFwdSmem::kP + (k >> 2) * 16384) + 2 * (k & 3);
          const uint64_t dv = umma_desc_mnmajor_sw128(base + AttnFwdSmem::kV + k * 2048, 1024);
          umma_f16_ss(tmem_O, dp, dv, idesc_o, k > 0);
        }
        umma_commit(bar_o);
        if (i + 1 < nkb) {
          mbar_wait(bar_o, (uint32_t)(i & 1));  // V buffer is free once P V has completed
          mbar_expect_tx(bar_v, 16384);
          tma_load_3d(base + AttnFwdSmem::kV, &plan.tm_qkv, bar_v, 1024 + h * 64, (i + 1) * kAK, b);
        }
      }
    }
  } else {
    const int qd = warp & 3, g = warp >> 2;
    const int r = qd * 32 + lane;  // query row within the tile == TMEM lane
    const int qi = q0 + r;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = 0.f;
    const bool q_side = qi < iso_p;
    uint8_t* sP = gbase + AttnFwdSmem::kP;

    for (int i = 0; i < nkb; ++i) {
      const int k0 = i * kAK;
      uint32_t vw;
      {  // validity bits of this warp's 32 keys
        const int key = k0 + g * 32 + lane;
        const bool ok = key < L && keymask[(long)b * L + key] != 0.f;
        vw = __ballot_sync(0xffffffffu, ok);
        if (iso_p > 0) {
          const int nb = iso_p - (k0 + g * 32);
          const uint32_t below = nb <= 0 ? 0u : (nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u));
          vw &= q_side ? below : ~below;
        }
      }
      mbar_wait(bar_s, (uint32_t)(i & 1));
      if (threadIdx.x == 0 && i < 2) stamp(3 + 4 * i);
      tc_fence_after();
      uint32_t v[32];
      __syncwarp();
      tmem_ld_32x32b_x32(tmem_S + lane_addr + g * 32, v);
      tmem_ld_wait();
      float m_loc = -INFINITY;
      const bool all_valid = vw == 0xffffffffu;   // warp-uniform: no per-element masking on the common path
      if (all_valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) m_loc = fmaxf(m_loc, __uint_as_float(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((vw >> j) & 1u) m_loc = fmaxf(m_loc, __uint_as_float(v[j]));
      }
      float* mrow = mx + ((i & 1) * 128 + r) * 4;
      mrow[g] = m_loc;
      softmax_bar_sync();
      if (threadIdx.x == 0 && i < 2) stamp(4 + 4 * i);
      const float4 m4 = *reinterpret_cast<const float4*>(mrow);
      const float m_blk = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w)) * kScaleLog2;
      const float m_new = fmaxf(m_run, m_blk);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = exp2_fast(m_run - m_use);
      float p[32];
      float rowsum = 0.f;
      if (all_valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[j] = exp2_fast(fmaf(__uint_as_float(v[j]), kScaleLog2, -m_use));
          rowsum += p[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pv = ((vw >> j) & 1u) ? exp2_fast(fmaf(__uint_as_float(v[j]), kScaleLog2, -m_use)) : 0.f;
          p[j] = pv;
          rowsum += pv;
        }
      }
      {
        uint8_t* chunk = sP + (g >> 1) * 16384 + r * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int unit = (g & 1) * 4 + u;
          *reinterpret_cast<uint4*>(chunk + ((unit ^ (r & 7)) << 4)) = pack8_h16(p + 8 * u, bf);
        }
      }
      l_run = l_run * alpha + rowsum;   // partial sum over this thread's key group
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
      if (threadIdx.x == 0 && i < 2) stamp(5 + 4 * i);
      // O' = P V of this block: this thread accumulates output columns [16g, 16g+16)
      mbar_wait(bar_o, (uint32_t)(i & 1));
      if (threadIdx.x == 0 && i < 2) stamp(6 + 4 * i);
      tc_fence_after();
      uint32_t ov[16];
      __syncwarp();
      tmem_ld_32x32b_x16(tmem_O + lane_addr + g * 16, ov);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = o[j] * alpha + __uint_as_float(ov[j]);
    }
    // total row sum = sum of the four groups' partial sums (same running max in all four)
    float* lrow = mx + ((nkb & 1) * 128 + r) * 4;
    softmax_bar_sync();
    lrow[g] = l_run;
    softmax_bar_sync();
    const float4 l4 = *reinterpret_cast<const float4*>(lrow);
    const float l_tot = (l4.x + l4.y) + (l4.z + l4.w);
    if (qi < L) {
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] *= inv;
      uint16_t* dst = o_out + ((long)b * L + qi) * 512 + h * 64 + g * 16;
      reinterpret_cast<uint4*>(dst)[0] = pack8_h16(o, bf);
      reinterpret_cast<uint4*>(dst)[1] = pack8_h16(o + 8, bf);
      if (lse_out && g == 0) lse_out[((long)b * 8 + h) * L + qi] = l_tot > 0.f ? m_run + log2f(l_tot) : INFINITY;
    }
  }
  if (threadIdx.x == 0) stamp(11);
  tc_fence_before();
  __syncthreads();
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem, 256); }
  if (threadIdx.x == 0) stamp(12);
}

int attn_fwd_prepare(void* plan_, const void* qkv, long ldq, int B, int L, int bf16, char* err, int errlen) {
  AttnPlan* p = reinterpret_cast<AttnPlan*>(plan_);
  memset(p, 0, sizeof(*p));
  p->B = B; p->L = L; p->bf16 = bf16; p->dbg = g_attn_dbg;
  int r = tma_encode_3d(&p->tm_qkv, qkv, bf16, 1536, (uint64_t)L, (uint64_t)B, (uint64_t)ldq * 2,
                        (uint64_t)L * ldq * 2, 64, 128, 1);
  if (r) { if (err) snprintf(err, errlen, "attn: cuTensorMapEncodeTiled(qkv) failed (%d)", r); return -1; }
  return 0;
}

int attn_fwd_launch(const void* plan_, const float* keymask, int iso_p, void* o, float* lse, cudaStream_t st) {
  const AttnPlan* p = reinterpret_cast<const AttnPlan*>(plan_);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnFwdSmem::kBytes);
    attr_done = true;
  }
  dim3 grid((p->L + kAQ - 1) / kAQ, 8, p->B);
  launch_pdl(attn_fwd_kernel, grid, kFwdThreads, AttnFwdSmem::kBytes, st, *p, keymask, iso_p, reinterpret_cast<uint16_t*>(o), lse);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// backward: see attention_bwd.cu
}  // namespace cvflow
