// attn1 backward on tcgen05 (sm_100a): two launches per attention layer, both with the persistent
// two-warpgroup structure of the forward kernel (attention.h).
//
//   dQ kernel   : item = (b, h, pair of 128-query tiles); thread = query row; key blocks of <= 96
//        delta = rowsum(dO o O) (fused here, also written out for the dK/dV kernel)
//        S = Q K^T, dP = dO V^T                        -> TMEM [0,96) / [96,192)
//        dS = P o (dP - delta) d^-1/2, P = 2^(s c - lse) -> 16-bit, stored over S with tcgen05.st
//        dQ += dS K   (A = dS from TMEM, B = K block as MN-major smem operand) -> TMEM [192,256),
//        accumulated in TMEM over the key blocks
//   dK/dV kernel: item = (b, h, pair of 128-key tiles); thread = key row; query blocks of <= 32, score tiles
//        double-buffered in TMEM (buffer b: S^T = K Q^T at [64b,+32), dP^T = V dO^T at [64b+32,+32))
//        P^T, dS^T -> 16-bit over S^T / dP^T
//        dV += P^T dO, dK += dS^T Q (A from TMEM, B = dO / Q blocks as MN-major operands) -> [128,192) / [192,256)
// P is recomputed from the stored base-2 log-sum-exp of the forward pass. Both kernels are
// deterministic (no atomics). Masking matches the forward kernel: padded keys and the
// prompt-isolation boundary give P = 0 (reference modules.py:275-288 under autograd). Rows at or
// beyond the sample's valid extent are skipped: their dQ/dK/dV are written as zeros and their dO is
// taken as zero (in the estimator every consumer of those rows is masked, so it is).
#include "kernels.h"
#include "gemm.h"
#include "attention.h"
#include <string.h>

namespace cvflow {

// ------------------------------------------------------------------------------------------
// dQ kernel
// ------------------------------------------------------------------------------------------
static constexpr int kDqKB = 96;    // keys per block
struct DqSmem {
  static constexpr int kQ = 0;                 // [2 stages][2 tiles][16 KB]
  static constexpr int kdO = 65536;            // [2 stages][2 tiles][16 KB]
  static constexpr int kKV = 131072;           // [2 slots][K 12 KB | V 12 KB]
  static constexpr int kSlot = 24576;
  static constexpr int kStg = 180224;          // [2 warpgroups][16 KB] output staging
  static constexpr int kBar = 212992;
  static constexpr int kBytes = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ AttnPlan plan, const int* __restrict__ kinfo, int iso_p,
                   const uint16_t* __restrict__ o_in, const uint16_t* __restrict__ do_in, const float* __restrict__ lse,
                   float* __restrict__ delta_out, uint16_t* __restrict__ dqkv) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + DqSmem::kBar;
  auto q_full = [&](int s) { return bar + 8u * s; };
  auto q_empty = [&](int s) { return bar + 16u + 8u * s; };
  auto kv_full = [&](int s) { return bar + 32u + 8u * s; };
  auto kv_empty = [&](int s) { return bar + 48u + 8u * s; };
  auto s_full = [&](int w) { return bar + 64u + 8u * w; };
  auto p_full = [&](int w) { return bar + 80u + 8u * w; };
  auto acc_full = [&](int w) { return bar + 96u + 8u * w; };
  auto acc_free = [&](int w) { return bar + 112u + 8u * w; };
  const uint32_t tmem_slot = bar + 128u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = plan.L, bf = plan.bf16;
  const int tpi = plan.tpi;
  const int npairs = ((L + 127) / 128 + tpi - 1) / tpi;   // items per (batch, head)
  const int n_items = plan.B * 8 * npairs;
  const int* ord = attn_order_ptr(kinfo, plan.B, L);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1); mbar_init(q_empty(s), 2);
      mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2);
      mbar_init(s_full(s), 1); mbar_init(p_full(s), 128);
      mbar_init(acc_full(s), 1); mbar_init(acc_free(s), 128);
    }
    fence_barrier_init();
    tma_prefetch_desc(&plan.tm_qkv);
    tma_prefetch_desc(&plan.tm_qkv32);
    tma_prefetch_desc(&plan.tm_do);
    tma_prefetch_desc(&plan.tm_dqkv);
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
// the per-sample extents were written at the start of the step (estimator) -- not by the previous kernel --
  // so the first item's extent is fetched before the grid-dependency wait and the producer's first TMA is not behind it
  AttnItem first;
  if (plan.early_kinfo) first = attn_item(min((int)blockIdx.x, n_items - 1), npairs, tpi, plan.B, ord);
  pdl_wait();
  if (!plan.early_kinfo) first = attn_item(min((int)blockIdx.x, n_items - 1), npairs, tpi, plan.B, ord);

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int qs = 0, ring = 0;
      uint32_t qph = 0, rph = 0;
      AttnItem nxt = first;
      for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
        const AttnItem a = nxt;
        nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);
        if (!a.act[0]) continue;
        mbar_wait(q_empty(qs), qph ^ 1u);
        const int ntile = a.act[1] ? 2 : 1;
        mbar_expect_tx(q_full(qs), (uint32_t)ntile * 32768u);
        for (int t = 0; t < ntile; ++t)
          for (int hf = 0; hf < 2; ++hf) {
            const int row = (a.tile0 + t) * 128 + hf * 64;
            tma_load_3d(base + DqSmem::kQ + qs * 32768 + t * 16384 + hf * 8192, &plan.tm_qkv, q_full(qs), a.h * 64, row, a.b);
            tma_load_3d(base + DqSmem::kdO + qs * 32768 + t * 16384 + hf * 8192, &plan.tm_do, q_full(qs), a.h * 64, row, a.b);
          }
        const int nkb = (a.ext + kDqKB - 1) / kDqKB;
        for (int blk = 0; blk < nkb; ++blk) {
          const int k0 = blk * kDqKB;
          const int keb = min(kDqKB, a.ext - k0);
          const int nbox = (keb + 31) >> 5;
          mbar_wait(kv_empty(ring), rph ^ 1u);
          mbar_expect_tx(kv_full(ring), (uint32_t)nbox * 8192u);
          const uint32_t slot = base + DqSmem::kKV + ring * DqSmem::kSlot;
          for (int x = 0; x < nbox; ++x) {
            tma_load_3d(slot + x * 4096, &plan.tm_qkv32, kv_full(ring), 512 + a.h * 64, k0 + x * 32, a.b);
            tma_load_3d(slot + 12288 + x * 4096, &plan.tm_qkv32, kv_full(ring), 1024 + a.h * 64, k0 + x * 32, a.b);
          }
          if (++ring == 2) { ring = 0; rph ^= 1u; }
        }
        if (++qs == 2) { qs = 0; qph ^= 1u; }
      }
    }
  } else if (warp >= 9) {
    // ---------------- MMA issuer of warpgroup w ----------------
    if (lane == 0) {
      const int w = warp - 9;
      const uint32_t treg = tmem + (uint32_t)(w * 256);
      const uint32_t idesc_q = umma_idesc_f16(bf, 128, 64, 0, 1);
      int qs = 0, ring = 0;
      uint32_t qph = 0, rph = 0, n = 0, m = 0;
      AttnItem nxt = first;
      for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
        const AttnItem a = nxt;
        nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);
        if (!a.act[0]) continue;
        mbar_wait(q_full(qs), qph);
        const int nkb = (a.ext + kDqKB - 1) / kDqKB;
        const uint32_t sQ = base + DqSmem::kQ + qs * 32768 + w * 16384;
        const uint32_t sdO = base + DqSmem::kdO + qs * 32768 + w * 16384;
        for (int blk = 0; blk < nkb; ++blk) {
          const int keb = min(kDqKB, a.ext - blk * kDqKB);
          const uint32_t sK = base + DqSmem::kKV + ring * DqSmem::kSlot;
          const uint32_t sV = sK + 12288;
          mbar_wait(kv_full(ring), rph);
          if (a.act[w]) {
            tc_fence_after();
            const uint32_t idesc_s = umma_idesc_f16(bf, 128, keb, 0, 0);
            const uint64_t dq = umma_desc_kmajor_sw128(sQ), dk = umma_desc_kmajor_sw128(sK);
            const uint64_t ddo = umma_desc_kmajor_sw128(sdO), dv = umma_desc_kmajor_sw128(sV);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(treg, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(treg + 96, ddo + 2 * k, dv + 2 * k, idesc_s, k > 0);
            umma_commit(s_full(w));
            if (blk == nkb - 1) umma_commit(q_empty(qs));
            mbar_wait(p_full(w), n & 1u);
            if (blk == 0) mbar_wait(acc_free(w), (m & 1u) ^ 1u);   // the previous item's dQ has been read out
            tc_fence_after();
            const int nk = keb >> 4;
            for (int k = 0; k < nk; ++k) {
              const uint64_t db = umma_desc_mnmajor_sw128(sK + k * 2048, 1024);
              umma_f16_ts(treg + 192, treg + (uint32_t)(k * 8), db, idesc_q, (blk > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(kv_empty(ring));
            if (blk == nkb - 1) { umma_commit(acc_full(w)); ++m; }
            ++n;
          } else {
            if (blk == nkb - 1) mbar_arrive(q_empty(qs));
            mbar_arrive(kv_empty(ring));
          }
          if (++ring == 2) { ring = 0; rph ^= 1u; }
        }
        if (++qs == 2) { qs = 0; qph ^= 1u; }
      }
    }
  } else {
    // ---------------- row warpgroups: thread = query row ----------------
    const int w = warp >> 2, qd = warp & 3;
    const int r = qd * 32 + lane;
    const int wtid = threadIdx.x & 127;
    const uint32_t treg = tmem + (uint32_t)(w * 256) + ((uint32_t)(qd * 32) << 16);
    const int nwords = 8 * ((L + 255) / 256);
    const int* bits_base = kinfo + ((plan.B + 3) & ~3);
    uint32_t n = 0, m = 0;
    AttnItem nxt = first;
    for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
      const AttnItem a = nxt;
      nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);
      const int q0 = (a.tile0 + w) * 128;
      const int qi = q0 + r;
      const long stat_idx = ((long)a.b * 8 + a.h) * L + qi;
      if (!a.act[w]) {   // tile of padding rows only
        if (w < tpi && qi < L) {
          uint16_t* dst = dqkv + ((long)a.b * L + qi) * 1536 + a.h * 64;
#pragma unroll
          for (int u = 0; u < 8; ++u) reinterpret_cast<uint4*>(dst)[u] = make_uint4(0u, 0u, 0u, 0u);
          delta_out[stat_idx] = 0.f;
        }
        continue;
      }
      // delta = rowsum(dO o O) and the row's log-sum-exp
      float my_delta = 0.f, my_lse = INFINITY;
      if (qi < L) {
        if (qi < a.kmax) {
          my_lse = lse[stat_idx];
          const uint4* po = reinterpret_cast<const uint4*>(o_in + ((long)a.b * L + qi) * 512 + a.h * 64);
          const uint4* pd = reinterpret_cast<const uint4*>(do_in + ((long)a.b * L + qi) * 512 + a.h * 64);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint4 x = __ldg(po + u), y = __ldg(pd + u);
            float xa, xb, ya, yb;
            unpack2_h16(x.x, bf, xa, xb); unpack2_h16(y.x, bf, ya, yb); my_delta = fmaf(xa, ya, fmaf(xb, yb, my_delta));
            unpack2_h16(x.y, bf, xa, xb); unpack2_h16(y.y, bf, ya, yb); my_delta = fmaf(xa, ya, fmaf(xb, yb, my_delta));
            unpack2_h16(x.z, bf, xa, xb); unpack2_h16(y.z, bf, ya, yb); my_delta = fmaf(xa, ya, fmaf(xb, yb, my_delta));
            unpack2_h16(x.w, bf, xa, xb); unpack2_h16(y.w, bf, ya, yb); my_delta = fmaf(xa, ya, fmaf(xb, yb, my_delta));
          }
        }   // else: padding row inside an active tile: P = 0 (lse = +inf), so dS = 0
        delta_out[stat_idx] = my_delta;
      }
      const bool q_below = qi < iso_p;
      const float lse_s = my_lse + 3.0f;   // d^-1/2 folded into the exponent: -log2(0.125) = 3
      const int nkb = (a.ext + kDqKB - 1) / kDqKB;
      const int* bits = bits_base + (long)a.b * nwords;
      for (int blk = 0; blk < nkb; ++blk) {
        const int k0 = blk * kDqKB;
        const int keb = min(kDqKB, a.ext - k0);
        uint32_t vw[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          vw[c] = (3 * blk + c < nwords) ? attn_iso_word((uint32_t)__ldg(bits + 3 * blk + c), k0 + 32 * c, iso_p, q_below) : 0u;
        mbar_wait(s_full(w), n & 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c * 32 < keb) {   // the last chunk of a block may hold only 16 live columns; the rest is masked by vw
            uint32_t sv[32], pv[32];
            tmem_ld_32x32b_x32(treg + (uint32_t)(c * 32), sv);
            tmem_ld_32x32b_x32(treg + 96u + (uint32_t)(c * 32), pv);
            tmem_ld_wait();
            float ds[32];
            if (vw[c] == 0xffffffffu) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {   // packed fp32x2: one FFMA2, one FADD2, one FMUL2 and two MUFU.EX2 per pair
                float a0, a1;
                ffma2(a0, a1, __uint_as_float(sv[j]), __uint_as_float(sv[j + 1]), kAttnScaleLog2, -lse_s);
                const f32x2 e = f2_pack(exp2_fast(a0), exp2_fast(a1));
                const f32x2 d = f2_add(f2_pack(__uint_as_float(pv[j]), __uint_as_float(pv[j + 1])), f2_pack(-my_delta, -my_delta));
                f2_unpack(f2_mul(e, d), ds[j], ds[j + 1]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                ds[j] = ((vw[c] >> j) & 1u)
                            ? exp2_fast(fmaf(__uint_as_float(sv[j]), kAttnScaleLog2, -lse_s)) * (__uint_as_float(pv[j]) - my_delta)
                            : 0.f;
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack2_h16(ds[2 * j], ds[2 * j + 1], bf);
            tmem_st_32x32b_x16(treg + (uint32_t)(c * 16), pk);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(w));
        ++n;
      }
      // dQ of the tile: TMEM -> 16-bit -> swizzled staging tile -> TMA store
      mbar_wait(acc_full(w), m & 1u);
      tc_fence_after();
      float dqv[64];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(treg + 192u + (uint32_t)(hf * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) dqv[hf * 32 + j] = __uint_as_float(v[j]);
      }
      tc_fence_before();
      mbar_arrive(acc_free(w));
      ++m;
      const uint32_t stg = base + DqSmem::kStg + w * 16384;
      if (wtid == 0) tma_store_wait_read();
      wg_bar_sync(w);
      const uint32_t rowaddr = stg + (uint32_t)r * 128u;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint4 pk = pack8_h16(dqv + 8 * u, bf);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((u ^ (r & 7)) << 4)), "r"(pk.x),
                     "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
      }
      fence_proxy_async_smem();
      wg_bar_sync(w);
      if (wtid == 0) {
        tma_store_3d(&plan.tm_dqkv, stg, a.h * 64, q0, a.b);
        if (q0 + 64 < L) tma_store_3d(&plan.tm_dqkv, stg + 8192, a.h * 64, q0 + 64, a.b);
        tma_store_commit();
      }
    }
    if (wtid == 0) tma_store_wait_all();
  }
  pdl_launch();   // dependents are released late: CTAs of the next kernel that spin at their grid-dependency wait next to the working ones cost more than their prologue overlap gains (same-box A/B)
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------
// dK / dV kernel
//
// The score tiles are 32 queries wide and DOUBLE-BUFFERED in TMEM (buffer b: S^T at [64b, +32), dP^T at [64b+32, +32),
// accumulators dV [128,192), dK [192,256)): the MMA warp keeps the score MMAs of block n+1 (and n+2 once block n's
// operands are consumed) in flight while the row threads work on block n, so the row threads never wait a full
// MMA round trip per block (measured: 1,800 of 4,400 clocks per 64-query block were that wait).
// ------------------------------------------------------------------------------------------
static constexpr int kDkvQB = 32;   // queries per block
static constexpr int kDkvSlotQ = 64;   // queries per shared-memory ring slot (two blocks)
struct DkvSmem {
  static constexpr int kK = 0;                 // [2 stages][2 tiles][16 KB]
  static constexpr int kV = 65536;             // [2 stages][2 tiles][16 KB]
  static constexpr int kQdO = 131072;          // [3 slots][Q 8 KB | dO 8 KB]
  static constexpr int kSlot = 16384;
  static constexpr int kNumSlots = 3;
  static constexpr int kStg = 180224;          // [2 warpgroups][16 KB] output staging
  static constexpr int kStat = 212992;         // float [2 warpgroups][2 buffers][-lse 64 | -delta d^-1/2 64]
  static constexpr int kBar = 215040;
  static constexpr int kBytes = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_dkv_kernel(const __grid_constant__ AttnPlan plan, const int* __restrict__ kinfo, int iso_p,
                    const float* __restrict__ lse, const float* __restrict__ delta, uint16_t* __restrict__ dqkv) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + DkvSmem::kBar;
  auto kv_full = [&](int s) { return bar + 8u * s; };
  auto kv_empty = [&](int s) { return bar + 16u + 8u * s; };
  auto c_full = [&](int s) { return bar + 32u + 8u * s; };     // 3 slots
  auto c_empty = [&](int s) { return bar + 56u + 8u * s; };    // 3 slots
  auto s_full = [&](int w, int b) { return bar + 80u + 8u * (2 * w + b); };    // per warpgroup and score buffer
  auto p_full = [&](int w, int b) { return bar + 112u + 8u * (2 * w + b); };
  auto acc_full = [&](int w) { return bar + 144u + 8u * w; };
  auto acc_free = [&](int w) { return bar + 160u + 8u * w; };
  const uint32_t tmem_slot = bar + 176u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = plan.L, bf = plan.bf16;
  const int tpi = plan.tpi;
  const int npairs = ((L + 127) / 128 + tpi - 1) / tpi;   // items per (batch, head)
  const int n_items = plan.B * 8 * npairs;
  const int* ord = attn_order_ptr(kinfo, plan.B, L);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2);
      mbar_init(acc_full(s), 1); mbar_init(acc_free(s), 128);
      for (int b = 0; b < 2; ++b) { mbar_init(s_full(s, b), 1); mbar_init(p_full(s, b), 128); }
    }
    for (int s = 0; s < DkvSmem::kNumSlots; ++s) { mbar_init(c_full(s), 1); mbar_init(c_empty(s), 2); }
    fence_barrier_init();
    tma_prefetch_desc(&plan.tm_qkv);
    tma_prefetch_desc(&plan.tm_do);
    tma_prefetch_desc(&plan.tm_dqkv);
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
// the per-sample extents were written at the start of the step (estimator) -- not by the previous kernel --
  // so the first item's extent is fetched before the grid-dependency wait and the producer's first TMA is not behind it
  AttnItem first;
  if (plan.early_kinfo) first = attn_item(min((int)blockIdx.x, n_items - 1), npairs, tpi, plan.B, ord);
  pdl_wait();
  if (!plan.early_kinfo) first = attn_item(min((int)blockIdx.x, n_items - 1), npairs, tpi, plan.B, ord);

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int ks = 0, ring = 0;
      uint32_t kph = 0, rph = 0;
      AttnItem nxt = first;
      for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
        const AttnItem a = nxt;
        nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);
        if (!a.act[0]) continue;
        mbar_wait(kv_empty(ks), kph ^ 1u);
        const int ntile = a.act[1] ? 2 : 1;
        mbar_expect_tx(kv_full(ks), (uint32_t)ntile * 32768u);
        for (int t = 0; t < ntile; ++t)
          for (int hf = 0; hf < 2; ++hf) {
            const int row = (a.tile0 + t) * 128 + hf * 64;
            tma_load_3d(base + DkvSmem::kK + ks * 32768 + t * 16384 + hf * 8192, &plan.tm_qkv, kv_full(ks), 512 + a.h * 64, row, a.b);
            tma_load_3d(base + DkvSmem::kV + ks * 32768 + t * 16384 + hf * 8192, &plan.tm_qkv, kv_full(ks), 1024 + a.h * 64, row, a.b);
          }
        const int nslots = (a.ext + kDkvSlotQ - 1) / kDkvSlotQ;
        for (int sl = 0; sl < nslots; ++sl) {
          mbar_wait(c_empty(ring), rph ^ 1u);
          mbar_expect_tx(c_full(ring), 16384u);
          const uint32_t slot = base + DkvSmem::kQdO + ring * DkvSmem::kSlot;
          tma_load_3d(slot, &plan.tm_qkv, c_full(ring), a.h * 64, sl * kDkvSlotQ, a.b);
          tma_load_3d(slot + 8192, &plan.tm_do, c_full(ring), a.h * 64, sl * kDkvSlotQ, a.b);
          if (++ring == DkvSmem::kNumSlots) { ring = 0; rph ^= 1u; }
        }
        if (++ks == 2) { ks = 0; kph ^= 1u; }
      }
    }
  } else if (warp >= 9) {
    // ---------------- MMA issuer of warpgroup w ----------------
    if (lane == 0) {
      const int w = warp - 9;
      const uint32_t treg = tmem + (uint32_t)(w * 256);
      const uint32_t idesc_a = umma_idesc_f16(bf, 128, 64, 0, 1);
      int ks = 0, ring = 0;         // ring = slot of the item's first query block
      uint32_t kph = 0, rph = 0, m = 0;
      uint32_t pcnt[2] = {0u, 0u};  // completed uses of the two P^T/dS^T buffers (parity of p_full)
      AttnItem nxt = first;
      for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
        const AttnItem a = nxt;
        nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);
        if (!a.act[0]) continue;
        mbar_wait(kv_full(ks), kph);
        const int nqb = (a.ext + kDkvQB - 1) / kDkvQB;
        const int nslots = (a.ext + kDkvSlotQ - 1) / kDkvSlotQ;
        const uint32_t sK = base + DkvSmem::kK + ks * 32768 + w * 16384;
        const uint32_t sV = base + DkvSmem::kV + ks * 32768 + w * 16384;
        // ring position / phase of slot i of this item
        auto slot_of = [&](int i, int& pos, uint32_t& ph) {
          pos = ring + i; ph = rph;
          while (pos >= DkvSmem::kNumSlots) { pos -= DkvSmem::kNumSlots; ph ^= 1u; }
        };
        if (a.act[w]) {
          int waited = -1;   // slots [0, waited] have been seen full
          auto issue_score = [&](int n) {
            int pos; uint32_t ph;
            slot_of(n >> 1, pos, ph);
            if ((n >> 1) > waited) { mbar_wait(c_full(pos), ph); waited = n >> 1; }
            tc_fence_after();
            const int qeb = min(kDkvQB, a.ext - n * kDkvQB);
            const uint32_t sQ = base + DkvSmem::kQdO + pos * DkvSmem::kSlot + (n & 1) * 4096;
            const uint32_t sdO = sQ + 8192;
            const uint32_t idesc_s = umma_idesc_f16(bf, 128, qeb, 0, 0);
            const uint64_t dk = umma_desc_kmajor_sw128(sK), dq = umma_desc_kmajor_sw128(sQ);
            const uint64_t dv = umma_desc_kmajor_sw128(sV), ddo = umma_desc_kmajor_sw128(sdO);
            const uint32_t tb = treg + (uint32_t)((n & 1) * 64);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tb, dk + 2 * k, dq + 2 * k, idesc_s, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tb + 32, dv + 2 * k, ddo + 2 * k, idesc_s, k > 0);
            umma_commit(s_full(w, n & 1));
            if (n == nqb - 1) umma_commit(kv_empty(ks));   // K / V tiles are only operands of the score MMAs
          };
          issue_score(0);
          if (nqb > 1) issue_score(1);
          for (int n = 0; n < nqb; ++n) {
            const int b = n & 1;
            mbar_wait(p_full(w, b), pcnt[b] & 1u);
            ++pcnt[b];
            if (n == 0) mbar_wait(acc_free(w), (m & 1u) ^ 1u);
            tc_fence_after();
            int pos; uint32_t ph;
            slot_of(n >> 1, pos, ph);
            const int qeb = min(kDkvQB, a.ext - n * kDkvQB);
            const uint32_t sQ = base + DkvSmem::kQdO + pos * DkvSmem::kSlot + (n & 1) * 4096;
            const uint32_t sdO = sQ + 8192;
            const uint32_t tb = treg + (uint32_t)(b * 64);
            const int nk = qeb >> 4;
            for (int k = 0; k < nk; ++k) {
              const uint64_t db = umma_desc_mnmajor_sw128(sdO + k * 2048, 1024);
              umma_f16_ts(treg + 128, tb + (uint32_t)(k * 8), db, idesc_a, (n > 0 || k > 0) ? 1u : 0u);
            }
            for (int k = 0; k < nk; ++k) {
              const uint64_t db = umma_desc_mnmajor_sw128(sQ + k * 2048, 1024);
              umma_f16_ts(treg + 192, tb + 32u + (uint32_t)(k * 8), db, idesc_a, (n > 0 || k > 0) ? 1u : 0u);
            }
            if ((n & 1) == 1 || n == nqb - 1) umma_commit(c_empty(pos));   // both halves of the slot are consumed
            if (n + 2 < nqb) issue_score(n + 2);   // into the buffer the accumulate MMAs above have just read (in-order pipe)
          }
          umma_commit(acc_full(w));
          ++m;
        } else {
          for (int i = 0; i < nslots; ++i) {
            int pos; uint32_t ph;
            slot_of(i, pos, ph);
            mbar_wait(c_full(pos), ph);
            mbar_arrive(c_empty(pos));
          }
          mbar_arrive(kv_empty(ks));
        }
        ring += nslots;
        while (ring >= DkvSmem::kNumSlots) { ring -= DkvSmem::kNumSlots; rph ^= 1u; }
        if (++ks == 2) { ks = 0; kph ^= 1u; }
      }
    }
  } else {
    // ---------------- row warpgroups: thread = key row ----------------
    const int w = warp >> 2, qd = warp & 3;
    const int r = qd * 32 + lane;
    const int wtid = threadIdx.x & 127;
    const uint32_t treg = tmem + (uint32_t)(w * 256) + ((uint32_t)(qd * 32) << 16);
    const int nwords = 8 * ((L + 255) / 256);
    const int* bits_base = kinfo + ((plan.B + 3) & ~3);
    float* stat = reinterpret_cast<float*>(gbase + DkvSmem::kStat) + w * 256;   // [2 buffers][-lse 64 | -delta d^-1/2 64]
    uint32_t m = 0, sgrp = 0;
    uint32_t scnt[2] = {0u, 0u};   // completed uses of the two score buffers (parity of s_full)
    AttnItem nxt = first;
    for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
      const AttnItem a = nxt;
      nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);
      const int k0 = (a.tile0 + w) * 128;
      const int kj = k0 + r;
      if (!a.act[w]) {   // tile of padding keys only
        if (w < tpi && kj < L) {
          uint16_t* dst = dqkv + ((long)a.b * L + kj) * 1536 + 512 + a.h * 64;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            reinterpret_cast<uint4*>(dst)[u] = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4*>(dst + 512)[u] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        continue;
      }
      const bool key_ok = kj < a.kmax && ((__ldg(bits_base + (long)a.b * nwords + (kj >> 5)) >> (kj & 31)) & 1);
      const bool k_below = kj < iso_p;
      const int nqb = (a.ext + kDkvQB - 1) / kDkvQB;
      const float* lse_row = lse + ((long)a.b * 8 + a.h) * L;
      const float* del_row = delta + ((long)a.b * 8 + a.h) * L;
      for (int n = 0; n < nqb; ++n) {
        const int b = n & 1;
        const int q0 = n * kDkvQB;
        if (b == 0) {   // per-column statistics of the next 64 queries (two blocks): -lse and -delta d^-1/2 (dead queries: P = 0)
          float* sbw = stat + (sgrp & 1u) * 128;
          const int q = q0 + (wtid & 63);
          const bool live = q < a.kmax;
          sbw[wtid] = wtid < 64 ? (live ? -lse_row[q] : -INFINITY) : (live ? -del_row[q] * kAttnScale : 0.f);
          wg_bar_sync(w);
          ++sgrp;
        }
        const float* sb = stat + ((sgrp - 1u) & 1u) * 128 + b * 32;
        mbar_wait(s_full(w, b), scnt[b] & 1u);
        ++scnt[b];
        tc_fence_after();
        const uint32_t tb = treg + (uint32_t)(b * 64);
        {
          const int nlive = a.kmax - q0;   // live query columns of this block
          uint32_t col_ok = !key_ok || nlive <= 0 ? 0u : (nlive >= 32 ? 0xffffffffu : ((1u << nlive) - 1u));
          col_ok = attn_iso_word(col_ok, q0, iso_p, k_below);
          uint32_t sv[32], pv[32];
          tmem_ld_32x32b_x32(tb, sv);
          tmem_ld_32x32b_x32(tb + 32u, pv);
          tmem_ld_wait();
          uint32_t pk[16], dk[16];
          const float4* l4 = reinterpret_cast<const float4*>(sb);
          const float4* d4 = reinterpret_cast<const float4*>(sb + 64);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 lv = l4[g], dv = d4[g];
            const float ls[4] = {lv.x, lv.y, lv.z, lv.w}, de[4] = {dv.x, dv.y, dv.z, dv.w};
            float pp[4], dd[4];
            if (col_ok == 0xffffffffu) {   // packed fp32x2: per pair two FFMA2, one FMUL2 and two MUFU.EX2
              const f32x2 c2 = f2_pack(kAttnScaleLog2, kAttnScaleLog2), k2 = f2_pack(kAttnScale, kAttnScale);
#pragma unroll
              for (int e = 0; e < 4; e += 2) {
                const int j = 4 * g + e;
                float a0, a1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1])), c2, f2_pack(ls[e], ls[e + 1])), a0, a1);
                const f32x2 pr = f2_pack(exp2_fast(a0), exp2_fast(a1));
                const f32x2 t = f2_fma(f2_pack(__uint_as_float(pv[j]), __uint_as_float(pv[j + 1])), k2, f2_pack(de[e], de[e + 1]));
                f2_unpack(pr, pp[e], pp[e + 1]);
                f2_unpack(f2_mul(pr, t), dd[e], dd[e + 1]);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = 4 * g + e;
                const bool on = (col_ok >> j) & 1u;
                pp[e] = on ? exp2_fast(fmaf(__uint_as_float(sv[j]), kAttnScaleLog2, ls[e])) : 0.f;
                dd[e] = on ? pp[e] * fmaf(__uint_as_float(pv[j]), kAttnScale, de[e]) : 0.f;
              }
            }
            pk[2 * g] = pack2_h16(pp[0], pp[1], bf); pk[2 * g + 1] = pack2_h16(pp[2], pp[3], bf);
            dk[2 * g] = pack2_h16(dd[0], dd[1], bf); dk[2 * g + 1] = pack2_h16(dd[2], dd[3], bf);
          }
          __syncwarp();
          tmem_st_32x32b_x16(tb, pk);
          tmem_st_32x32b_x16(tb + 32u, dk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(w, b));
      }
      // dV, dK of the tile: TMEM -> 16-bit -> swizzled staging tile -> TMA stores (dV first, then dK through the same tile)
      mbar_wait(acc_full(w), m & 1u);
      tc_fence_after();
      const uint32_t stg = base + DkvSmem::kStg + w * 16384;
      const uint32_t rowaddr = stg + (uint32_t)r * 128u;
#pragma unroll
      for (int which = 0; which < 2; ++which) {   // 0: dV -> columns 1024.., 1: dK -> columns 512..
        float f[64];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(treg + (which == 0 ? 128u : 192u) + (uint32_t)(hf * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[hf * 32 + j] = __uint_as_float(v[j]);
        }
        if (which == 1) {
          tc_fence_before();
          mbar_arrive(acc_free(w));
          ++m;
        }
        if (wtid == 0) tma_store_wait_read();
        wg_bar_sync(w);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 pk = pack8_h16(f + 8 * u, bf);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((u ^ (r & 7)) << 4)), "r"(pk.x),
                       "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
        }
        fence_proxy_async_smem();
        wg_bar_sync(w);
        if (wtid == 0) {
          const int col = (which == 0 ? 1024 : 512) + a.h * 64;
          tma_store_3d(&plan.tm_dqkv, stg, col, k0, a.b);
          if (k0 + 64 < L) tma_store_3d(&plan.tm_dqkv, stg + 8192, col, k0 + 64, a.b);
          tma_store_commit();
        }
      }
    }
    if (wtid == 0) tma_store_wait_all();
  }
  pdl_launch();   // dependents are released late: CTAs of the next kernel that spin at their grid-dependency wait next to the working ones cost more than their prologue overlap gains (same-box A/B)
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------
int attn_bwd_prepare(void* plan_, const void* qkv, long ldq, const void* dout, int B, int L, int bf16, char* err,
                     int errlen) {
  AttnPlan* p = reinterpret_cast<AttnPlan*>(plan_);
  memset(p, 0, sizeof(*p));
  p->B = B; p->L = L; p->bf16 = bf16;
  int r = tma_encode_3d(&p->tm_qkv, qkv, bf16, 1536, (uint64_t)L, (uint64_t)B, (uint64_t)ldq * 2,
                        (uint64_t)L * ldq * 2, 64, 64, 1);
  if (!r) r = tma_encode_3d(&p->tm_qkv32, qkv, bf16, 1536, (uint64_t)L, (uint64_t)B, (uint64_t)ldq * 2,
                            (uint64_t)L * ldq * 2, 64, 32, 1);
  if (!r) r = tma_encode_3d(&p->tm_do, dout, bf16, 512, (uint64_t)L, (uint64_t)B, 512 * 2, (uint64_t)L * 512 * 2, 64, 64, 1);
  if (r) { if (err) snprintf(err, errlen, "attn bwd: cuTensorMapEncodeTiled failed (%d)", r); return -1; }
  p->o_ptr = nullptr;
  return 0;
}

int attn_bwd_launch(void* plan_, const void* dout, const int* kinfo, int iso_p, const void* o, const float* lse,
                    float* delta, void* dqkv, cudaStream_t st) {
  AttnPlan* p = reinterpret_cast<AttnPlan*>(plan_);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::kBytes);
    cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::kBytes);
    attr_done = true;
  }
  if (p->o_ptr != dqkv) {
    int r = tma_encode_3d(&p->tm_dqkv, dqkv, p->bf16, 1536, (uint64_t)p->L, (uint64_t)p->B, 1536 * 2,
                          (uint64_t)p->L * 1536 * 2, 64, 64, 1);
    if (r) return -(int)cudaErrorInvalidValue;
    p->o_ptr = dqkv;
  }
  p->tpi = attn_tiles_per_item(p->B, p->L);
  const int n_items = attn_num_items(p->B, p->L, p->tpi);
  const int grid = n_items < attn_num_sms() ? n_items : attn_num_sms();
  launch_pdl(attn_bwd_dq_kernel, dim3((unsigned)grid), kAttnThreads, DqSmem::kBytes, st, *p, kinfo, iso_p,
             reinterpret_cast<const uint16_t*>(o), reinterpret_cast<const uint16_t*>(dout), lse, delta,
             reinterpret_cast<uint16_t*>(dqkv));
  launch_pdl(attn_bwd_dkv_kernel, dim3((unsigned)grid), kAttnThreads, DkvSmem::kBytes, st, *p, kinfo, iso_p, lse, delta,
             reinterpret_cast<uint16_t*>(dqkv));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace cvflow
