// attn1 backward on tcgen05 (sm_100a): three launches per attention layer.
//
//   delta[b,h,q] = sum_d dO*O                                   (HBM-bound prep pass)
//   dQ kernel   : CTA = 128 queries, loops over key blocks      (rows = queries)
//        S = Q K^T, dP = dO V^T  -> TMEM;  dS = P o (dP - delta) * d^-1/2  -> smem (16-bit)
//        dQ += dS K                (A = dS K-major, B = K block as MN-major operand) in TMEM
//   dK/dV kernel: CTA = 128 keys, loops over query blocks       (rows = keys)
//        S^T = K Q^T, dP^T = V dO^T -> TMEM;  P^T, dS^T -> smem (16-bit)
//        dV += P^T dO, dK += dS^T Q (B = dO / Q blocks as MN-major operands) in TMEM
// P is recomputed from the stored base-2 log-sum-exp of the forward pass. Both kernels are
// deterministic (no atomics). Masking matches the forward kernel: padded keys and the
// prompt-isolation boundary give P = 0 (reference modules.py:275-288 under autograd).
#include "kernels.h"
#include "gemm.h"
#include "attention.h"
#include <string.h>

namespace cvflow {

static constexpr float kScale = 0.125f;
static constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;

__device__ __forceinline__ void rows_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
static constexpr int kBwdThreads = 544;   // 16 row warps (quadrant = w & 3, column group = w >> 2) + 1 MMA/TMA warp

// write 32 consecutive K-elements (columns c*32 .. c*32+31) of row r into a [128 x 128] 16-bit
// K-major SW128 operand made of two 16 KB column chunks
__device__ __forceinline__ void store_row_chunk(uint8_t* tile, int r, int c, const float (&v)[32], int bf) {
  uint8_t* chunk = tile + (c >> 1) * 16384 + r * 128;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int unit = (c & 1) * 4 + u;
    *reinterpret_cast<uint4*>(chunk + ((unit ^ (r & 7)) << 4)) = pack8_h16(v + 8 * u, bf);
  }
}

// ------------------------------------------------------------------------------------------
// delta = rowsum(dO * O) per head
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const uint16_t* __restrict__ dO, const uint16_t* __restrict__ O,
                                                         float* __restrict__ delta, int L, long M, int bf) {
  pdl_wait();
  pdl_launch();
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float s = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint4 a = *reinterpret_cast<const uint4*>(dO + row * 512 + lane * 16 + half * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(O + row * 512 + lane * 16 + half * 8);
    float x[8], y[8];
    unpack2_h16(a.x, bf, x[0], x[1]); unpack2_h16(a.y, bf, x[2], x[3]);
    unpack2_h16(a.z, bf, x[4], x[5]); unpack2_h16(a.w, bf, x[6], x[7]);
    unpack2_h16(b.x, bf, y[0], y[1]); unpack2_h16(b.y, bf, y[2], y[3]);
    unpack2_h16(b.z, bf, y[4], y[5]); unpack2_h16(b.w, bf, y[6], y[7]);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += x[e] * y[e];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if ((lane & 3) == 0) {
    const int h = lane >> 2;
    const long b = row / L;
    const int q = (int)(row - b * L);
    delta[(b * 8 + h) * L + q] = s;
  }
}

// ------------------------------------------------------------------------------------------
// dQ kernel
// ------------------------------------------------------------------------------------------
struct DqSmem {
  static constexpr int kQ = 0, kdO = 16384;
  static constexpr int kK0 = 32768, kK1 = 49152, kV0 = 65536, kV1 = 81920;
  static constexpr int kdS = 98304;   // 32 KB
  static constexpr int kBar = 131072;
  static constexpr int kBytes = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ AttnPlan plan, const float* __restrict__ keymask, int iso_p,
                   const float* __restrict__ lse, const float* __restrict__ delta, uint16_t* __restrict__ dqkv) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + DqSmem::kBar;
  const uint32_t bar_q = bar, bar_kv0 = bar + 8, bar_kv1 = bar + 16, bar_sp = bar + 24, bar_ds = bar + 32,
                 bar_dq = bar + 40, tmem_slot = bar + 48;
  uint32_t* kvalid = reinterpret_cast<uint32_t*>(gbase + DqSmem::kBar + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int L = plan.L, bf = plan.bf16;
  const int nkb = (L + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_kv0, 1); mbar_init(bar_kv1, 1); mbar_init(bar_sp, 1);
    mbar_init(bar_ds, 512); mbar_init(bar_dq, 1);
    fence_barrier_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tmem_S = tmem, tmem_dP = tmem + 128, tmem_dQ = tmem + 256;
  pdl_wait();
  pdl_launch();

  if (warp == 16) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(bf, 128, 128, 0, 0);
      const uint32_t idesc_q = umma_idesc_f16(bf, 128, 64, 0, 1);
      mbar_expect_tx(bar_q, 32768);
      tma_load_3d(base + DqSmem::kQ, &plan.tm_qkv, bar_q, h * 64, q0, b);
      tma_load_3d(base + DqSmem::kdO, &plan.tm_do, bar_q, h * 64, q0, b);
      mbar_expect_tx(bar_kv0, 32768);
      tma_load_3d(base + DqSmem::kK0, &plan.tm_qkv, bar_kv0, 512 + h * 64, 0, b);
      tma_load_3d(base + DqSmem::kV0, &plan.tm_qkv, bar_kv0, 1024 + h * 64, 0, b);
      mbar_wait(bar_q, 0);
      for (int i = 0; i < nkb; ++i) {
        const int s = i & 1;
        const uint32_t sK = base + (s ? DqSmem::kK1 : DqSmem::kK0);
        const uint32_t sV = base + (s ? DqSmem::kV1 : DqSmem::kV0);
        mbar_wait(s ? bar_kv1 : bar_kv0, (uint32_t)((i >> 1) & 1));
        tc_fence_after();
        {
          const uint64_t dq = umma_desc_kmajor_sw128(base + DqSmem::kQ);
          const uint64_t dk = umma_desc_kmajor_sw128(sK);
          const uint64_t ddo = umma_desc_kmajor_sw128(base + DqSmem::kdO);
          const uint64_t dv = umma_desc_kmajor_sw128(sV);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_dP, ddo + 2 * k, dv + 2 * k, idesc_s, k > 0);
        }
        umma_commit(bar_sp);
        if (i + 1 < nkb) {
          if (i >= 1) mbar_wait(bar_dq, (uint32_t)((i - 1) & 1));  // buffers of block i-1 are free
          const int s2 = (i + 1) & 1;
          const uint32_t bk = s2 ? bar_kv1 : bar_kv0;
          mbar_expect_tx(bk, 32768);
          tma_load_3d(base + (s2 ? DqSmem::kK1 : DqSmem::kK0), &plan.tm_qkv, bk, 512 + h * 64, (i + 1) * 128, b);
          tma_load_3d(base + (s2 ? DqSmem::kV1 : DqSmem::kV0), &plan.tm_qkv, bk, 1024 + h * 64, (i + 1) * 128, b);
        }
        mbar_wait(bar_ds, (uint32_t)(i & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = umma_desc_kmajor_sw128(base + DqSmem::kdS + (k >> 2) * 16384) + 2 * (k & 3);
          const uint64_t db = umma_desc_mnmajor_sw128(sK + k * 2048, 1024);
          umma_f16_ss(tmem_dQ, da, db, idesc_q, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_dq);
      }
    }
  } else {
    const int qd = warp & 3, g = warp >> 2;
    const int r = qd * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const bool q_side = qi < iso_p;
    const long stat_idx = ((long)b * 8 + h) * L + qi;
    const float my_lse = qi < L ? lse[stat_idx] : INFINITY;
    const float my_delta = qi < L ? delta[stat_idx] : 0.f;
    uint8_t* sdS = gbase + DqSmem::kdS;
    for (int i = 0; i < nkb; ++i) {
      const int k0 = i * 128;
      uint32_t vw;
      {
        const int key = k0 + g * 32 + lane;
        const bool ok = key < L && keymask[(long)b * L + key] != 0.f;
        vw = __ballot_sync(0xffffffffu, ok);
        if (iso_p > 0) {
          const int nb = iso_p - (k0 + 32 * g);
          const uint32_t below = nb <= 0 ? 0u : (nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u));
          vw &= q_side ? below : ~below;
        }
      }
      mbar_wait(bar_sp, (uint32_t)(i & 1));
      if (i >= 1) mbar_wait(bar_dq, (uint32_t)((i - 1) & 1));  // dS tile no longer read by dQ MMA(i-1)
      tc_fence_after();
      uint32_t sv[32], pv[32];
      __syncwarp();
      tmem_ld_32x32b_x32(tmem_S + lane_addr + g * 32, sv);
      tmem_ld_32x32b_x32(tmem_dP + lane_addr + g * 32, pv);
      tmem_ld_wait();
      float ds[32];
      // ds = P (dP - delta) d^-1/2 with the scale folded into the exponent: 2^(s c - (lse - log2 d^-1/2))
      const float lse_s = my_lse + 3.0f;   // -log2(0.125) = 3
      if (vw == 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          ds[j] = exp2_fast(fmaf(__uint_as_float(sv[j]), kScaleLog2, -lse_s)) * (__uint_as_float(pv[j]) - my_delta);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          ds[j] = ((vw >> j) & 1u)
                      ? exp2_fast(fmaf(__uint_as_float(sv[j]), kScaleLog2, -lse_s)) * (__uint_as_float(pv[j]) - my_delta)
                      : 0.f;
      }
      store_row_chunk(sdS, r, g, ds, bf);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_ds);
    }
    mbar_wait(bar_dq, (uint32_t)((nkb - 1) & 1));
    tc_fence_after();
    {
      uint32_t v[16];
      __syncwarp();
      tmem_ld_32x32b_x16(tmem_dQ + lane_addr + g * 16, v);
      tmem_ld_wait();
      if (qi < L) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        uint16_t* dst = dqkv + ((long)b * L + qi) * 1536 + h * 64 + g * 16;
        reinterpret_cast<uint4*>(dst)[0] = pack8_h16(f, bf);
        reinterpret_cast<uint4*>(dst)[1] = pack8_h16(f + 8, bf);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------
// dK / dV kernel
// ------------------------------------------------------------------------------------------
struct DkvSmem {
  static constexpr int kK = 0, kV = 16384;
  static constexpr int kQ0 = 32768, kQ1 = 49152, kdO0 = 65536, kdO1 = 81920;
  static constexpr int kPT = 98304;    // 32 KB
  static constexpr int kdST = 131072;  // 32 KB
  static constexpr int kBar = 163840;
  static constexpr int kBytes = kBar + 2304 + 1024;   // barriers + 2 x (lse, delta)[128]
};

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dkv_kernel(const __grid_constant__ AttnPlan plan, const float* __restrict__ keymask, int iso_p,
                    const float* __restrict__ lse, const float* __restrict__ delta, uint16_t* __restrict__ dqkv) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + DkvSmem::kBar;
  const uint32_t bar_kv = bar, bar_q0 = bar + 8, bar_q1 = bar + 16, bar_sp = bar + 24, bar_pd = bar + 32,
                 bar_acc = bar + 40, tmem_slot = bar + 48;
  float* s_lse = reinterpret_cast<float*>(gbase + DkvSmem::kBar + 256);    // [2][128]
  float* s_delta = s_lse + 256;                                            // [2][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int L = plan.L, bf = plan.bf16;
  const int nqb = (L + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(bar_kv, 1); mbar_init(bar_q0, 1); mbar_init(bar_q1, 1); mbar_init(bar_sp, 1);
    mbar_init(bar_pd, 512); mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tmem_ST = tmem, tmem_dPT = tmem + 128, tmem_dV = tmem + 256, tmem_dK = tmem + 320;
  pdl_wait();
  pdl_launch();

  if (warp == 16) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(bf, 128, 128, 0, 0);
      const uint32_t idesc_a = umma_idesc_f16(bf, 128, 64, 0, 1);
      mbar_expect_tx(bar_kv, 32768);
      tma_load_3d(base + DkvSmem::kK, &plan.tm_qkv, bar_kv, 512 + h * 64, k0, b);
      tma_load_3d(base + DkvSmem::kV, &plan.tm_qkv, bar_kv, 1024 + h * 64, k0, b);
      mbar_expect_tx(bar_q0, 32768);
      tma_load_3d(base + DkvSmem::kQ0, &plan.tm_qkv, bar_q0, h * 64, 0, b);
      tma_load_3d(base + DkvSmem::kdO0, &plan.tm_do, bar_q0, h * 64, 0, b);
      mbar_wait(bar_kv, 0);
      for (int i = 0; i < nqb; ++i) {
        const int s = i & 1;
        const uint32_t sQ = base + (s ? DkvSmem::kQ1 : DkvSmem::kQ0);
        const uint32_t sdO = base + (s ? DkvSmem::kdO1 : DkvSmem::kdO0);
        mbar_wait(s ? bar_q1 : bar_q0, (uint32_t)((i >> 1) & 1));
        tc_fence_after();
        {
          const uint64_t dk = umma_desc_kmajor_sw128(base + DkvSmem::kK);
          const uint64_t dq = umma_desc_kmajor_sw128(sQ);
          const uint64_t dv = umma_desc_kmajor_sw128(base + DkvSmem::kV);
          const uint64_t ddo = umma_desc_kmajor_sw128(sdO);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_ST, dk + 2 * k, dq + 2 * k, idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_dPT, dv + 2 * k, ddo + 2 * k, idesc_s, k > 0);
        }
        umma_commit(bar_sp);
        if (i + 1 < nqb) {
          if (i >= 1) mbar_wait(bar_acc, (uint32_t)((i - 1) & 1));
          const int s2 = (i + 1) & 1;
          const uint32_t bq = s2 ? bar_q1 : bar_q0;
          mbar_expect_tx(bq, 32768);
          tma_load_3d(base + (s2 ? DkvSmem::kQ1 : DkvSmem::kQ0), &plan.tm_qkv, bq, h * 64, (i + 1) * 128, b);
          tma_load_3d(base + (s2 ? DkvSmem::kdO1 : DkvSmem::kdO0), &plan.tm_do, bq, h * 64, (i + 1) * 128, b);
        }
        mbar_wait(bar_pd, (uint32_t)(i & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t dp = umma_desc_kmajor_sw128(base + DkvSmem::kPT + (k >> 2) * 16384) + 2 * (k & 3);
          const uint64_t db = umma_desc_mnmajor_sw128(sdO + k * 2048, 1024);
          umma_f16_ss(tmem_dV, dp, db, idesc_a, (i > 0 || k > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t ds = umma_desc_kmajor_sw128(base + DkvSmem::kdST + (k >> 2) * 16384) + 2 * (k & 3);
          const uint64_t db = umma_desc_mnmajor_sw128(sQ + k * 2048, 1024);
          umma_f16_ss(tmem_dK, ds, db, idesc_a, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_acc);
      }
    }
  } else {
    const int qd = warp & 3, g = warp >> 2;
    const int r = qd * 32 + lane;   // key row
    const int kj = k0 + r;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const bool key_ok = kj < L && keymask[(long)b * L + kj] != 0.f;
    const bool k_side = kj < iso_p;
    uint8_t* sPT = gbase + DkvSmem::kPT;
    uint8_t* sdST = gbase + DkvSmem::kdST;
    for (int i = 0; i < nqb; ++i) {
      const int q0 = i * 128;
      if (threadIdx.x < 128) {
        const int q = q0 + threadIdx.x;
        const long idx = ((long)b * 8 + h) * L + q;
        s_lse[(i & 1) * 128 + threadIdx.x] = q < L ? lse[idx] : INFINITY;
        s_delta[(i & 1) * 128 + threadIdx.x] = q < L ? delta[idx] : 0.f;
      }
      rows_bar_sync();
      const float* lse_i = s_lse + (i & 1) * 128 + g * 32;
      const float* del_i = s_delta + (i & 1) * 128 + g * 32;
      mbar_wait(bar_sp, (uint32_t)(i & 1));
      if (i >= 1) mbar_wait(bar_acc, (uint32_t)((i - 1) & 1));
      tc_fence_after();
      uint32_t col_ok = key_ok ? 0xffffffffu : 0u;
      if (iso_p > 0) {
        const int nb = iso_p - (q0 + 32 * g);
        const uint32_t below = nb <= 0 ? 0u : (nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u));
        col_ok &= k_side ? below : ~below;
      }
      uint32_t sv[32], pv[32];
      __syncwarp();
      tmem_ld_32x32b_x32(tmem_ST + lane_addr + g * 32, sv);
      tmem_ld_32x32b_x32(tmem_dPT + lane_addr + g * 32, pv);
      tmem_ld_wait();
      float pp[32], ds[32];
      if (col_ok == 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pj = exp2_fast(fmaf(__uint_as_float(sv[j]), kScaleLog2, -lse_i[j]));
          pp[j] = pj;
          ds[j] = pj * ((__uint_as_float(pv[j]) - del_i[j]) * kScale);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pj = ((col_ok >> j) & 1u) ? exp2_fast(fmaf(__uint_as_float(sv[j]), kScaleLog2, -lse_i[j])) : 0.f;
          pp[j] = pj;
          ds[j] = pj * ((__uint_as_float(pv[j]) - del_i[j]) * kScale);
        }
      }
      store_row_chunk(sPT, r, g, pp, bf);
      store_row_chunk(sdST, r, g, ds, bf);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_pd);
    }
    mbar_wait(bar_acc, (uint32_t)((nqb - 1) & 1));
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {   // 0: dK -> cols 512.., 1: dV -> cols 1024..
      uint32_t v[16];
      __syncwarp();
      tmem_ld_32x32b_x16((which == 0 ? tmem_dK : tmem_dV) + lane_addr + g * 16, v);
      tmem_ld_wait();
      if (kj < L) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        uint16_t* dst = dqkv + ((long)b * L + kj) * 1536 + (which == 0 ? 512 : 1024) + h * 64 + g * 16;
        reinterpret_cast<uint4*>(dst)[0] = pack8_h16(f, bf);
        reinterpret_cast<uint4*>(dst)[1] = pack8_h16(f + 8, bf);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------
int attn_bwd_prepare(void* plan_, const void* qkv, long ldq, const void* dout, int B, int L, int bf16, char* err,
                     int errlen) {
  AttnPlan* p = reinterpret_cast<AttnPlan*>(plan_);
  memset(p, 0, sizeof(*p));
  p->B = B; p->L = L; p->bf16 = bf16;
  int r = tma_encode_3d(&p->tm_qkv, qkv, bf16, 1536, (uint64_t)L, (uint64_t)B, (uint64_t)ldq * 2,
                        (uint64_t)L * ldq * 2, 64, 128, 1);
  if (!r) r = tma_encode_3d(&p->tm_do, dout, bf16, 512, (uint64_t)L, (uint64_t)B, 512 * 2, (uint64_t)L * 512 * 2, 64, 128, 1);
  if (r) { if (err) snprintf(err, errlen, "attn bwd: cuTensorMapEncodeTiled failed (%d)", r); return -1; }
  return 0;
}

int attn_bwd_launch(const void* plan_, const void* dout, const float* keymask, int iso_p, const void* o, const float* lse,
                    float* delta, void* dqkv, cudaStream_t st) {
  const AttnPlan* p = reinterpret_cast<const AttnPlan*>(plan_);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::kBytes);
    cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::kBytes);
    attr_done = true;
  }
  const long M = (long)p->B * p->L;
  launch_pdl(attn_delta_kernel, (unsigned)((M + 7) / 8), 256, 0, st, reinterpret_cast<const uint16_t*>(dout),
                                                             reinterpret_cast<const uint16_t*>(o), delta, p->L, M, p->bf16);
  dim3 grid((p->L + 127) / 128, 8, p->B);
  launch_pdl(attn_bwd_dq_kernel, grid, kBwdThreads, DqSmem::kBytes, st, *p, keymask, iso_p, lse, delta, reinterpret_cast<uint16_t*>(dqkv));
  launch_pdl(attn_bwd_dkv_kernel, grid, kBwdThreads, DkvSmem::kBytes, st, *p, keymask, iso_p, lse, delta,
                                                          reinterpret_cast<uint16_t*>(dqkv));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace cvflow
