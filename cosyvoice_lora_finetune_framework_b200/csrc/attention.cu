// Flash-style attn1 self-attention on tcgen05 (sm_100a), forward and backward.
//
// Reference: Attention.forward, modules.py:253-293 -- softmax(q k^T * d^-1/2 + bias) v with
// 8 heads x 64, bias = -1e10 on padded keys (utils.py:103-109) plus the optional prompt-isolation
// block mask (modules.py:844-879,1034-1042). The [B,L,L] bias and the [B,8,L,L] scores are never
// materialised: the key mask comes from the [B,L] float mask and the isolation boundary from one
// integer.
//
// Layout of one CTA (forward and dQ kernels): 128 query rows of one (batch, head); thread r of
// warps 0-3 owns query row r (= TMEM lane r), warp 4 drives TMA and issues every tcgen05.mma.
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
};
int attn_plan_bytes() { return (int)sizeof(AttnPlan); }

static constexpr int kAQ = 128;   // query rows per CTA
static constexpr int kAK = 128;   // keys per block
static constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;  // d^-1/2 * log2(e)

// named barrier among the 128 softmax threads only
__device__ __forceinline__ void softmax_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct AttnFwdSmem {
  static constexpr int kQ = 0;
  static constexpr int kK0 = 16384;
  static constexpr int kK1 = 32768;
  static constexpr int kV = 49152;
  static constexpr int kP = 65536;          // 2 x 16 KB
  static constexpr int kBar = 98304;        // barriers + scratch
  static constexpr int kBytes = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(160, 2)
attn_fwd_kernel(const __grid_constant__ AttnPlan plan, const float* __restrict__ keymask, int iso_p,
                uint16_t* __restrict__ o_out, float* __restrict__ lse_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + AttnFwdSmem::kBar;
  const uint32_t bar_q = bar, bar_k0 = bar + 8, bar_k1 = bar + 16, bar_v = bar + 24, bar_s = bar + 32,
                 bar_p = bar + 40, bar_o = bar + 48, tmem_slot = bar + 56;
  uint32_t* kvalid = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + AttnFwdSmem::kBar + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAQ, h = blockIdx.y, b = blockIdx.z;
  const int L = plan.L, bf = plan.bf16;
  const int nkb = (L + kAK - 1) / kAK;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_k0, 1); mbar_init(bar_k1, 1); mbar_init(bar_v, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&plan.tm_qkv);
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tmem_S = tmem, tmem_O = tmem + 128;

  if (warp == 4) {
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
    const int r = threadIdx.x;  // query row within the tile == TMEM lane
    const int qi = q0 + r;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) o[j] = 0.f;
    const bool q_side = qi < iso_p;
    uint8_t* sP = smem_raw + (base - smem_u32(smem_raw)) + AttnFwdSmem::kP;

    for (int i = 0; i < nkb; ++i) {
      const int k0 = i * kAK;
      {
        const int key = k0 + r;
        const bool ok = key < L && keymask[(long)b * L + key] != 0.f;
        const uint32_t word = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) kvalid[(i & 1) * 4 + warp] = word;
      }
      softmax_bar_sync();
      uint32_t vw[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) vw[c] = kvalid[(i & 1) * 4 + c];
      if (iso_p > 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // keys [k0+32c, k0+32c+32): keep those on the query's side of the boundary
          const int lo = k0 + 32 * c;
          uint32_t below;  // bit j set iff key lo+j < iso_p
          const int nb = iso_p - lo;
          below = nb <= 0 ? 0u : (nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u));
          vw[c] &= q_side ? below : ~below;
        }
      }
      mbar_wait(bar_s, (uint32_t)(i & 1));
      tc_fence_after();
      // pass 1: block row max
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((vw[c] >> j) & 1u) m_blk = fmaxf(m_blk, __uint_as_float(v[j]));
      }
      m_blk *= kScaleLog2;  // scale > 0, max commutes
      const float m_new = fmaxf(m_run, m_blk);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = exp2f(m_run - m_use);
      float rowsum = 0.f;
      // pass 2: probabilities -> smem (16-bit, swizzled K-major operand)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pv = ((vw[c] >> j) & 1u) ? exp2f(__uint_as_float(v[j]) * kScaleLog2 - m_use) : 0.f;
          p[j] = pv;
          rowsum += pv;
        }
        uint8_t* chunk = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack2_h16(p[8 * u + 0], p[8 * u + 1], bf);
          w.y = pack2_h16(p[8 * u + 2], p[8 * u + 3], bf);
          w.z = pack2_h16(p[8 * u + 4], p[8 * u + 5], bf);
          w.w = pack2_h16(p[8 * u + 6], p[8 * u + 7], bf);
          const int unit = (c & 1) * 4 + u;
          *reinterpret_cast<uint4*>(chunk + ((unit ^ (r & 7)) << 4)) = w;
        }
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
      // O' = P V of this block
      mbar_wait(bar_o, (uint32_t)(i & 1));
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_O + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o[c * 32 + j] = o[c * 32 + j] * alpha + __uint_as_float(v[j]);
      }
    }
    if (qi < L) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      uint16_t* dst = o_out + ((long)b * L + qi) * 512 + h * 64;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint4 w;
        w.x = pack2_h16(o[8 * u + 0] * inv, o[8 * u + 1] * inv, bf);
        w.y = pack2_h16(o[8 * u + 2] * inv, o[8 * u + 3] * inv, bf);
        w.z = pack2_h16(o[8 * u + 4] * inv, o[8 * u + 5] * inv, bf);
        w.w = pack2_h16(o[8 * u + 6] * inv, o[8 * u + 7] * inv, bf);
        reinterpret_cast<uint4*>(dst)[u] = w;
      }
      if (lse_out) lse_out[((long)b * 8 + h) * L + qi] = l_run > 0.f ? m_run + log2f(l_run) : INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int attn_fwd_prepare(void* plan_, const void* qkv, int B, int L, int bf16, char* err, int errlen) {
  AttnPlan* p = reinterpret_cast<AttnPlan*>(plan_);
  memset(p, 0, sizeof(*p));
  p->B = B; p->L = L; p->bf16 = bf16;
  int r = tma_encode_3d(&p->tm_qkv, qkv, bf16, 1536, (uint64_t)L, (uint64_t)B, 1536 * 2, (uint64_t)L * 1536 * 2, 64,
                        128, 1);
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
  attn_fwd_kernel<<<grid, 160, AttnFwdSmem::kBytes, st>>>(*p, keymask, iso_p, reinterpret_cast<uint16_t*>(o), lse);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// backward: see attention_bwd.cu
}  // namespace cvflow
