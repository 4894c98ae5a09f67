// Flash-style attn1 self-attention forward on tcgen05 (sm_100a).
//
// Reference: Attention.forward, modules.py:253-293 -- softmax(q k^T * d^-1/2 + bias) v with
// 8 heads x 64, bias = -1e10 on padded keys (utils.py:103-109) plus the optional prompt-isolation
// block mask (modules.py:844-879,1034-1042). The [B,L,L] bias and the [B,8,L,L] scores are never
// materialised: the key mask comes from the [B,L] float mask and the isolation boundary from one
// integer.
//
// Structure (see attention.h): persistent CTAs over (batch, head, query-tile pair) items; per
// warpgroup
//   S  = Q K^T   : A = Q tile [128 x 64] (smem, K-major), B = key block [<=256 x 64] (smem, K-major)
//                  -> TMEM columns [0, keys) of the warpgroup's 256-column region
//   P  = softmax : thread = query row; two passes over the TMEM row (max, then exp2 / sum); P is
//                  rounded to 16 bits and written back over S with tcgen05.st (A operand of P V)
//   O' = P V     : A = P (TMEM), B = V block as MN-major smem operand -> TMEM columns [128, 192)
// With L <= 256 there is one key block and the softmax is single-pass; longer sequences use the
// online rescaling with the running max / sum / output row in registers of the row's thread.
#include "kernels.h"
#include "gemm.h"
#include "attention.h"
#include <string.h>

namespace cvflow {

static long long* g_attn_dbg = nullptr;
void attn_set_debug_buffer(void* p) { g_attn_dbg = reinterpret_cast<long long*>(p); }
int attn_plan_bytes() { return (int)sizeof(AttnPlan); }
int attn_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// ------------------------------------------------------------------------------------------
// kinfo[b] = kmax[b] = 1 + last index with mask != 0 (0 when the row is empty); behind the B extents
// (rounded up to 4), kinfo[Bp + b*nw + i] = validity bits of keys [32 i, 32 i + 32) of sample b (nw = attn_kinfo_words(L))
// ------------------------------------------------------------------------------------------
int attn_kinfo_words(int L) { return 8 * ((L + 255) / 256); }
static int kinfo_bits_off(int B) { return (B + 3) & ~3; }   // bit words start 16-byte aligned
long attn_kinfo_ints(int B, int L) { return kinfo_bits_off(B) + (long)B * attn_kinfo_words(L) + 2L * B * ((L + 127) / 128); }   // + order table (<= one entry per tile and sample)

__global__ void __launch_bounds__(256) attn_kinfo_kernel(const float* __restrict__ mask, int Bp, int L, int nw,
                                                         int* __restrict__ kinfo) {
  __shared__ int s_max[8];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31;
  int m = 0;
  for (int i0 = 0; i0 < nw * 32; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const bool ok = i < L && mask[(long)b * L + i] != 0.f;
    if (ok) m = i + 1;   // i increases per thread: the last hit is the largest
    const uint32_t word = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) kinfo[Bp + b * nw + (i >> 5)] = (int)word;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = max(m, s_max[i]);
    kinfo[b] = m;
  }
}
// (tile pair, sample) entries by decreasing cost (ties by index: a fixed, deterministic order); see attn_item()
__global__ void __launch_bounds__(256) attn_order_kernel(const int* __restrict__ kmax, int B, int L, int tpi, int* __restrict__ ord) {
  const int tiles = (L + 127) / 128;
  const int npairs = (tiles + tpi - 1) / tpi;
  const int n = npairs * B;
  auto cost = [&](int e) {
    const int pair = e / B, b = e - pair * B, km = kmax[b];
    int act = 0;
    for (int t = 0; t < tpi; ++t) act += ((pair * tpi + t) * 128 < km) ? 1 : 0;
    return act * ((km + 15) & ~15);
  };
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int mine = cost(e);
    int rank = 0;
    for (int o = 0; o < n; ++o) {
      const int v = cost(o);
      rank += (v > mine || (v == mine && o < e)) ? 1 : 0;
    }
    const int pair = e / B, b = e - pair * B;
    ord[2 * rank] = (pair << 16) | b;
    ord[2 * rank + 1] = kmax[b];
  }
}
int launch_attn_kinfo(const float* mask, int B, int L, int* kinfo, cudaStream_t st) {
  attn_kinfo_kernel<<<B, 256, 0, st>>>(mask, kinfo_bits_off(B), L, attn_kinfo_words(L), kinfo);
  attn_order_kernel<<<1, 256, 0, st>>>(kinfo, B, L, attn_tiles_per_item(B, L), kinfo + kinfo_bits_off(B) + B * attn_kinfo_words(L));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct FwdSmem {
  static constexpr int kQ = 0;               // [2 stages][2 tiles][16 KB]
  static constexpr int kKV = 65536;          // [2 slots][K 32 KB | V 32 KB]
  static constexpr int kSlot = 65536;
  static constexpr int kO = 196608;          // [2 warpgroups][16 KB] output staging for the TMA store
  static constexpr int kBar = 229376;
  static constexpr int kBytes = kBar + 256 + 1024;
};
static constexpr int kFwdKB = 256;   // keys per block

// one 32- (or 16-) column chunk of the softmax's second pass: P = 2^(s c - m) on the valid keys, row sum,
// 16-bit P stored back into TMEM as the A operand of P V
template <int NC>
__device__ __forceinline__ void fwd_softmax_chunk(uint32_t t_s, uint32_t t_p, uint32_t vw, float m_use, float& rs0, float& rs1,
                                                  int bf) {
  uint32_t v[32];
  if (NC == 32) {
    tmem_ld_32x32b_x32(t_s, v);
  } else {
    uint32_t v16[16];
    tmem_ld_32x32b_x16(t_s, v16);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = v16[j];
  }
  tmem_ld_wait();
  float p[NC];
  const uint32_t full = NC == 32 ? 0xffffffffu : 0xffffu;
  if ((vw & full) == full) {
#pragma unroll
    for (int j = 0; j < NC; j += 2) {
      float a0, a1;
      ffma2(a0, a1, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), kAttnScaleLog2, -m_use);
      if ((CVFLOW_POLY_PAIR_MASK >> (j >> 1)) & 1u) {   // compile-time choice per pair: FMA-pipe polynomial ...
        exp2_poly2(p[j], p[j + 1], a0, a1);
      } else {                                          // ... or MUFU.EX2
        p[j] = exp2_fast(a0);
        p[j + 1] = exp2_fast(a1);
      }
      fadd2(rs0, rs1, p[j], p[j + 1]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      p[j] = ((vw >> j) & 1u) ? exp2_fast(fmaf(__uint_as_float(v[j]), kAttnScaleLog2, -m_use)) : 0.f;
      rs0 += p[j];
    }
  }
  if (NC == 32) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack2_h16(p[2 * j], p[2 * j + 1], bf);
    tmem_st_32x32b_x16(t_p, pk);
  } else {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack2_h16(p[2 * j], p[2 * j + 1], bf);
    tmem_st_32x32b_x8(t_p, pk);
  }
}
template <int NC>
__device__ __forceinline__ float fwd_max_chunk(uint32_t t_s, uint32_t vw, float m_loc) {
  uint32_t v[32];
  if (NC == 32) {
    tmem_ld_32x32b_x32(t_s, v);
  } else {
    uint32_t v16[16];
    tmem_ld_32x32b_x16(t_s, v16);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = v16[j];
  }
  tmem_ld_wait();
  const uint32_t full = NC == 32 ? 0xffffffffu : 0xffffu;
  if ((vw & full) == full) {   // four independent FMNMX3 chains
    float m1 = m_loc, m2 = m_loc, m3 = m_loc;
#pragma unroll
    for (int j = 0; j < NC; j += 8) {
      m_loc = fmax3(m_loc, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
      m1 = fmax3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      m2 = fmax3(m2, __uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
      m3 = fmax3(m3, __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
    }
    m_loc = fmaxf(fmaxf(m_loc, m1), fmaxf(m2, m3));
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if ((vw >> j) & 1u) m_loc = fmaxf(m_loc, __uint_as_float(v[j]));
  }
  return m_loc;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ AttnPlan plan, const int* __restrict__ kmax_arr, int iso_p,
                uint16_t* __restrict__ o_out, float* __restrict__ lse_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + FwdSmem::kBar;
  auto q_full = [&](int s) { return bar + 8u * s; };
  auto q_empty = [&](int s) { return bar + 16u + 8u * s; };
  auto kv_full = [&](int s) { return bar + 32u + 8u * s; };
  auto kv_empty = [&](int s) { return bar + 48u + 8u * s; };
  auto s_full = [&](int w) { return bar + 64u + 8u * w; };
  auto p_full = [&](int w) { return bar + 80u + 8u * w; };
  auto o_full = [&](int w) { return bar + 96u + 8u * w; };
  auto o_free = [&](int w) { return bar + 112u + 8u * w; };
  const uint32_t tmem_slot = bar + 128u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = plan.L, bf = plan.bf16;
  const int tpi = plan.tpi;
  const int npairs = ((L + 127) / 128 + tpi - 1) / tpi;   // items per (batch, head)
  const int n_items = plan.B * 8 * npairs;
  const int* ord = attn_order_ptr(kmax_arr, plan.B, L);
  long long* dbg = plan.dbg ? plan.dbg + (long)blockIdx.x * 32 : nullptr;
  auto stamp = [&](int k) {
    if (dbg) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[k] = t; }
  };
  if (threadIdx.x == 0) { stamp(0); if (dbg) dbg[30] = clock64(); }

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1); mbar_init(q_empty(s), 2);
      mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2);
      mbar_init(s_full(s), 1); mbar_init(p_full(s), 128);
      mbar_init(o_full(s), 1); mbar_init(o_free(s), 128);
    }
    fence_barrier_init();
    tma_prefetch_desc(&plan.tm_qkv);
    tma_prefetch_desc(&plan.tm_o);
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  if (threadIdx.x == 0) stamp(1);
// the per-sample extents were written at the start of the step (estimator) -- not by the previous kernel --
  // so the first item's extent is fetched before the grid-dependency wait and the producer's first TMA is not behind it
  AttnItem first;
  if (plan.early_kinfo) first = attn_item(min((int)blockIdx.x, n_items - 1), npairs, tpi, plan.B, ord);
  pdl_wait();
  if (!plan.early_kinfo) first = attn_item(min((int)blockIdx.x, n_items - 1), npairs, tpi, plan.B, ord);
  if (threadIdx.x == 0) stamp(2);

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int qs = 0, ring = 0;
      uint32_t qph = 0, rph = 0;
      AttnItem nxt = first;
      for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
        const AttnItem a = nxt;
        nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);   // extent of the next item: load issued early
        if (!a.act[0]) continue;
        mbar_wait(q_empty(qs), qph ^ 1u);
        const int ntile = a.act[1] ? 2 : 1;
        mbar_expect_tx(q_full(qs), (uint32_t)ntile * 16384u);
        for (int t = 0; t < ntile; ++t)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(base + FwdSmem::kQ + qs * 32768 + t * 16384 + hf * 8192, &plan.tm_qkv, q_full(qs), a.h * 64,
                        (a.tile0 + t) * 128 + hf * 64, a.b);
        const int nkb = (a.ext + kFwdKB - 1) / kFwdKB;
        for (int blk = 0; blk < nkb; ++blk) {
          const int k0 = blk * kFwdKB;
          const int keb = min(kFwdKB, a.ext - k0);
          const int nbox = (keb + 63) >> 6;
          mbar_wait(kv_empty(ring), rph ^ 1u);
          mbar_expect_tx(kv_full(ring), (uint32_t)nbox * 16384u);
          const uint32_t slot = base + FwdSmem::kKV + ring * FwdSmem::kSlot;
          for (int x = 0; x < nbox; ++x) {
            tma_load_3d(slot + x * 8192, &plan.tm_qkv, kv_full(ring), 512 + a.h * 64, k0 + x * 64, a.b);
            tma_load_3d(slot + 32768 + x * 8192, &plan.tm_qkv, kv_full(ring), 1024 + a.h * 64, k0 + x * 64, a.b);
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
      const uint32_t idesc_o = umma_idesc_f16(bf, 128, 64, 0, 1);
      int qs = 0, ring = 0;
      uint32_t qph = 0, rph = 0, n = 0;
      AttnItem nxt = first;
      for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
        const AttnItem a = nxt;
        nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);
        if (!a.act[0]) continue;
        mbar_wait(q_full(qs), qph);
        if (w == 0 && n == 0) stamp(17);
        const int nkb = (a.ext + kFwdKB - 1) / kFwdKB;
        const uint32_t sQ = base + FwdSmem::kQ + qs * 32768 + w * 16384;
        for (int blk = 0; blk < nkb; ++blk) {
          const int keb = min(kFwdKB, a.ext - blk * kFwdKB);
          const uint32_t slot = base + FwdSmem::kKV + ring * FwdSmem::kSlot;
          mbar_wait(kv_full(ring), rph);
          if (w == 0 && n == 0) stamp(18);
          if (a.act[w]) {
            mbar_wait(o_free(w), (n & 1u) ^ 1u);   // the warpgroup has drained O' (and S / P) of its previous block
            tc_fence_after();
            const uint32_t idesc_s = umma_idesc_f16(bf, 128, keb, 0, 0);
            const uint64_t dq = umma_desc_kmajor_sw128(sQ);
            const uint64_t dk = umma_desc_kmajor_sw128(slot);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(treg, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
            umma_commit(s_full(w));
            if (blk == nkb - 1) umma_commit(q_empty(qs));
            mbar_wait(p_full(w), n & 1u);
            tc_fence_after();
            const int nk = keb >> 4;
            for (int k = 0; k < nk; ++k) {
              const uint64_t dv = umma_desc_mnmajor_sw128(slot + 32768 + k * 2048, 1024);
              umma_f16_ts(treg + 128, treg + (uint32_t)(k * 8), dv, idesc_o, k > 0);
            }
            umma_commit(o_full(w));
            umma_commit(kv_empty(ring));
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
    uint32_t n = 0;
    const int nwords = 8 * ((L + 255) / 256);
    const int* bits_base = kmax_arr + ((plan.B + 3) & ~3);
    AttnItem nxt = first;
    uint4 nb0 = __ldg(reinterpret_cast<const uint4*>(bits_base + (long)nxt.b * nwords));
    uint4 nb1 = __ldg(reinterpret_cast<const uint4*>(bits_base + (long)nxt.b * nwords) + 1);
    for (int k_it = 0, it = attn_sched(0, n_items), itn; it >= 0; ++k_it, it = itn) {
        itn = attn_sched(k_it + 1, n_items);
      const AttnItem a = nxt;
      const uint4 fb0 = nb0, fb1 = nb1;   // validity words of the item's first key block
      nxt = attn_item(itn >= 0 ? itn : it, npairs, tpi, plan.B, ord);
      nb0 = __ldg(reinterpret_cast<const uint4*>(bits_base + (long)nxt.b * nwords));
      nb1 = __ldg(reinterpret_cast<const uint4*>(bits_base + (long)nxt.b * nwords) + 1);
      const int qi = (a.tile0 + w) * 128 + r;
      uint16_t* dst = o_out + ((long)a.b * L + qi) * 512 + a.h * 64;
      if (!a.act[w]) {   // tile of padding rows only: defined zeros instead of the reference's masked garbage
        if (w < tpi && qi < L) {   // (with one tile per item the second warpgroup owns no tile at all)
#pragma unroll
          for (int u = 0; u < 8; ++u) reinterpret_cast<uint4*>(dst)[u] = make_uint4(0u, 0u, 0u, 0u);
          if (lse_out) lse_out[((long)a.b * 8 + a.h) * L + qi] = INFINITY;
        }
        continue;
      }
      const uint4* bits = reinterpret_cast<const uint4*>(bits_base + (long)a.b * nwords);
      const bool q_below = qi < iso_p;
      float m_run = -INFINITY, l_run = 0.f;
      float o[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) o[j] = 0.f;
      const int nkb = (a.ext + kFwdKB - 1) / kFwdKB;
      for (int blk = 0; blk < nkb; ++blk) {
        const int k0 = blk * kFwdKB;
        const int keb = min(kFwdKB, a.ext - k0);
        const int nfull = keb >> 5;
        const bool tail = (keb & 16) != 0;
        uint32_t vw[8];
        {
          const uint4 b0 = blk == 0 ? fb0 : __ldg(bits + 2 * blk), b1 = blk == 0 ? fb1 : __ldg(bits + 2 * blk + 1);
          vw[0] = b0.x; vw[1] = b0.y; vw[2] = b0.z; vw[3] = b0.w; vw[4] = b1.x; vw[5] = b1.y; vw[6] = b1.z; vw[7] = b1.w;
        }
        uint32_t vt = 0u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          vw[c] = attn_iso_word(vw[c], k0 + 32 * c, iso_p, q_below);
          if (c == nfull) vt = vw[c];
        }
        mbar_wait(s_full(w), n & 1u);
        const bool st_on = threadIdx.x == 0 && n < 2;
        if (st_on) stamp(3 + 6 * n);
        tc_fence_after();
        // pass 1: row maximum over the valid keys of the block
        float m_loc = -INFINITY;
#pragma unroll 1   // rolled on purpose: eight inlined copies of the chunk body blow the instruction footprint of the hot loop
        for (int c = 0; c < nfull; ++c) m_loc = fwd_max_chunk<32>(treg + (uint32_t)(c * 32), vw[c], m_loc);
        if (tail) m_loc = fwd_max_chunk<16>(treg + (uint32_t)(nfull * 32), vt, m_loc);
        if (st_on) stamp(4 + 6 * n);
        const float m_new = fmaxf(m_run, m_loc * kAttnScaleLog2);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = exp2_fast(m_run - m_use);
        // pass 2: P = 2^(s c - m), row sum, 16-bit P written over S
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < nfull; ++c)
          fwd_softmax_chunk<32>(treg + (uint32_t)(c * 32), treg + (uint32_t)(c * 16), vw[c], m_use, rs0, rs1, bf);
        if (tail) fwd_softmax_chunk<16>(treg + (uint32_t)(nfull * 32), treg + (uint32_t)(nfull * 16), vt, m_use, rs0, rs1, bf);
        const float rowsum = rs0 + rs1;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(w));
        if (st_on) stamp(5 + 6 * n);
        l_run = l_run * alpha + rowsum;
        m_run = m_new;
        // O' = P V of this block
        mbar_wait(o_full(w), n & 1u);
        if (st_on) stamp(6 + 6 * n);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(treg + 128u + (uint32_t)(hf * 32), ov);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) o[hf * 32 + j] = o[hf * 32 + j] * alpha + __uint_as_float(ov[j]);
        }
        tc_fence_before();
        mbar_arrive(o_free(w));
        if (st_on) stamp(7 + 6 * n);
        ++n;
      }
      {
        // normalised output row -> 128B-swizzled staging tile -> TMA store (rows >= L are clipped by the tensor map)
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) o[j] *= inv;
        const uint32_t stg = base + FwdSmem::kO + w * 16384;
        if (wtid == 0) tma_store_wait_read();   // the previous item's store has finished reading the tile
        wg_bar_sync(w);
        const uint32_t rowaddr = stg + (uint32_t)r * 128u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 pk = pack8_h16(o + 8 * u, bf);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((u ^ (r & 7)) << 4)), "r"(pk.x),
                       "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
        }
        fence_proxy_async_smem();
        wg_bar_sync(w);
        if (wtid == 0) {
          const int q0 = (a.tile0 + w) * 128;
          tma_store_3d(&plan.tm_o, stg, a.h * 64, q0, a.b);
          if (q0 + 64 < L) tma_store_3d(&plan.tm_o, stg + 8192, a.h * 64, q0 + 64, a.b);
          tma_store_commit();
        }
        if (lse_out && qi < L) lse_out[((long)a.b * 8 + a.h) * L + qi] = l_run > 0.f ? m_run + log2f(l_run) : INFINITY;
      }
      if (threadIdx.x == 0 && n <= 2) stamp(8 + 6 * (n - 1));
    }
    if (wtid == 0) tma_store_wait_all();   // the staging tile must outlive the bulk stores
  }
  if (threadIdx.x == 0) { stamp(15); if (dbg) dbg[31] = clock64(); }
  pdl_launch();   // dependents are released late: CTAs of the next kernel that spin at their grid-dependency wait next to the working ones cost more than their prologue overlap gains (same-box A/B)
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

void attn_plan_set_early_kinfo(void* plan, int on) { reinterpret_cast<AttnPlan*>(plan)->early_kinfo = on; }

int attn_fwd_prepare(void* plan_, const void* qkv, long ldq, int B, int L, int bf16, char* err, int errlen) {
  AttnPlan* p = reinterpret_cast<AttnPlan*>(plan_);
  memset(p, 0, sizeof(*p));
  p->B = B; p->L = L; p->bf16 = bf16; p->dbg = g_attn_dbg;
  int r = tma_encode_3d(&p->tm_qkv, qkv, bf16, 1536, (uint64_t)L, (uint64_t)B, (uint64_t)ldq * 2,
                        (uint64_t)L * ldq * 2, 64, 64, 1);
  if (r) { if (err) snprintf(err, errlen, "attn: cuTensorMapEncodeTiled(qkv) failed (%d)", r); return -1; }
  p->o_ptr = nullptr;
  return 0;
}

int attn_fwd_launch(void* plan_, const int* kinfo, int iso_p, void* o, float* lse, cudaStream_t st) {
  AttnPlan* p = reinterpret_cast<AttnPlan*>(plan_);
  if (p->o_ptr != o) {   // (re-)encode the output map when the destination moves
    int r = tma_encode_3d(&p->tm_o, o, p->bf16, 512, (uint64_t)p->L, (uint64_t)p->B, 512 * 2, (uint64_t)p->L * 512 * 2, 64, 64, 1);
    if (r) return -(int)cudaErrorInvalidValue;
    p->o_ptr = o;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kBytes);
    attr_done = true;
  }
  p->tpi = attn_tiles_per_item(p->B, p->L);
  const int n_items = attn_num_items(p->B, p->L, p->tpi);
  const int grid = n_items < attn_num_sms() ? n_items : attn_num_sms();
  launch_pdl(attn_fwd_kernel, dim3((unsigned)grid), kAttnThreads, FwdSmem::kBytes, st, *p, kinfo, iso_p,
             reinterpret_cast<uint16_t*>(o), lse);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// backward: see attention_bwd.cu
}  // namespace cvflow
