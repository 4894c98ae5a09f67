// tcgen05 / TMEM / TMA implicit-GEMM engine for sm_100a (see gemm.h for the contraction it runs).
//
// One CTA computes a 128 x BN output tile. Warp 0 (one lane) streams A/B k-blocks through a
// ring of shared-memory stages with TMA (128B-swizzled, K-major, 64 columns = one swizzle atom
// per row); warp 1 allocates BN TMEM columns and one lane issues tcgen05.mma (M=128, N=BN, K=16,
// fp32 accumulate in TMEM); tcgen05.commit hands stages back to the producer and finally the
// accumulator to warps 2-5, which read it with tcgen05.ld (thread = output row) and apply the
// fused epilogue (bias / GELU / GELU' / row mask / fp32 residual / channel-major store).
//
// Replaces the library calls the reference makes through torch: F.linear (lora.py:66-74,
// modules.py:138,266-268,291,218), nn.Conv1d / ConvTranspose1d (modules.py:65,87,101,112,943,981).
#include "gemm.h"
#include "common.cuh"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace cvflow {

static constexpr int BM = 128;
static constexpr int BK = 64;

// DEEP (BN = 128 only): one CTA per SM with six stages (192 KB in flight instead of 96 KB). A CTA's main loop runs at
// (bytes in flight) / (TMA round trip of ~1.1 us); the narrow-N GEMMs (N = 256: 200 tiles of 128x64, most SMs holding a
// single CTA) spent 4-6 us of their 8-10 us there. 100 tiles of 128x128 with twice the bytes in flight halve that.
template <int BN, int DEEP = 0>
struct GemmCfg {
  static constexpr int kStages = DEEP ? 6 : ((BN == 256) ? 4 : (BN == 128 ? 3 : 4));
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = 8 * 2048;   // per epilogue warp: 32 rows x 64 B (32 x 16-bit, or 16 x fp32 per pass)
  // two CTAs per SM leave (228 KB - 2 x 1 KB reserved) / 2 = 115,712 B each: stages + staging + 128 B of
  // barriers + 896 B of alignment slack (the dynamic window is at least 128-byte aligned; checked at run time)
  static constexpr int kExchBytes = (BN == 256) ? 2048 : 0;   // fused LayerNorm: row sums exchanged between the two column halves
  static constexpr int kPayloadBytes = kStages * kStageBytes + kStagingBytes + 128 + kExchBytes;
  static constexpr int kSmemBytes = kPayloadBytes + 896;
  static constexpr int kCtasPerSm = (BN == 256 || DEEP) ? 1 : 2;
  static constexpr int kTmemCols = BN;          // 64 / 128 / 256: powers of two >= 32
  static constexpr int kChunksPerWarp = BN / 64; // 8 epilogue warps: 2 per TMEM lane quadrant
};
static constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue

// epilogue operands of one 32-column chunk that do not depend on the accumulator: fetched before
// the accumulator is ready so their latency overlaps the main loop / the previous chunk
struct EpiPrefetch {
  uint4 v[8];   // fp32 residual (8 x float4) or 16-bit pre-activation (first 4 x uint4)
};

// Persistent CTAs: each loops over output tiles (tile id = m-tile * n_tiles + n-tile, so CTAs that
// run together share the A tile in L2). Accumulators are double-buffered in TMEM (2 x BN columns),
// so the epilogue of tile i (tcgen05.ld, activation, global stores) overlaps the TMA + MMA main
// loop of tile i+1.
// CL > 1: thread-block clusters of CL CTAs along N work on the same m-tile; every CTA fetches 1/CL of the A rows of
// a k-block and TMA-multicasts it to the whole cluster, so the L2 -> SM traffic of the (bandwidth-bound, short-K)
// GEMMs of this model drops from A + B to A/CL + B per k-block.
template <int BN, int CL, int DEEP = 0>
__global__ void __launch_bounds__(kGemmThreads, (BN == 256 || DEEP) ? 1 : 2)
gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, DEEP>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (smem_base - smem_u32(smem_raw) + Cfg::kPayloadBytes > Cfg::kSmemBytes) __trap();   // alignment slack exceeded
  const uint32_t staging_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = staging_base + Cfg::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntn_g = (p.grid_y + CL - 1) / CL;  // groups of CL n-tiles (the last one may hold phantom tiles beyond N)
  const int total_super = p.grid_x * ntn_g;    // (m-tile, n-group) work items of a cluster
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  constexpr int kSliceRows = BM / CL;
  long long* dbg = p.dbg ? p.dbg + (long)blockIdx.x * 16 : nullptr;
  // profiling aid: slot 0 = globaltimer at CTA start (ns, aligns the CTAs of a launch), slots 1..13 = SM cycle counter
  // (clock64: a few cycles per read; %globaltimer costs hundreds and perturbed the epilogue it was meant to time),
  // slot 14 = clock64 at CTA start, slot 7 = tiles run by the CTA
  auto stamp = [&](int k) {
    if (dbg) dbg[k] = clock64();
  };
  if (threadIdx.x == 0 && dbg) {
    long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[0] = t; dbg[14] = clock64();
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kGemmThreads - 64);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmW);
    if (p.tma_out) tma_prefetch_desc(&p.tmOut);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 2 * Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // every CTA's barriers are initialised before a peer multicasts / arrives into them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  // Weights do not depend on the predecessor: the W half of the first tile's stages is requested before the
  // grid-dependency wait, so only the A loads (half the bytes of the first ring fill) start after it.
  int w_pre = 0;
  if (CL == 1 && p.w_prefetch && warp == 0 && lane == 0 && cluster_id < total_super) {
    const int nt0 = cluster_id % ntn_g;
    w_pre = p.nkb_total < Cfg::kStages ? p.nkb_total : Cfg::kStages;
    for (int kb = 0; kb < w_pre; ++kb) {
      mbar_expect_tx(full_bar(kb), Cfg::kStageBytes);
      tma_load_2d(smem_base + kb * Cfg::kStageBytes + Cfg::kABytes, &p.tmW, full_bar(kb), kb * BK, nt0 * BN);
    }
  }
  if (p.pf_ptr && warp == 0 && lane == 1) {     // weights are static: no need to wait for the predecessor
    const unsigned per = ((p.pf_bytes + gridDim.x - 1) / gridDim.x + 15u) & ~15u;
    const unsigned off = blockIdx.x * per;
    if (off < p.pf_bytes) {
      const unsigned n = min(per, p.pf_bytes - off) & ~15u;
      if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t*>(p.pf_ptr) + off), "r"(n) : "memory");
    }
  }
  pdl_wait();     // everything above overlapped the previous kernel's tail; global memory from here on
  if (threadIdx.x == 0) stamp(1);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int st = cluster_id; st < total_super; st += n_clusters) {
        const int mt = st / ntn_g, nt = (st - mt * ntn_g) * CL + rank;
        const int b = mt / p.tiles_per_batch;
        const int i0 = (mt - b * p.tiles_per_batch) * BM;
        const int n0 = nt * BN;
        int kb_global = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg sg = p.seg[s];
          const CUtensorMap* tm = &p.tmA[sg.a_map];
          for (int kb = 0; kb < sg.nkb; ++kb, ++kb_global) {
            const bool w_done = CL == 1 && st == cluster_id && kb_global < w_pre;   // requested before the wait
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            if (!w_done) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            }
            if (CL > 1)
              tma_load_3d_mc(sa + rank * kSliceRows * 128, tm, full_bar(stage), sg.a_col0 + kb * BK,
                             i0 + sg.row_shift + rank * kSliceRows, b, kMask);
            else
              tma_load_3d(sa, tm, full_bar(stage), sg.a_col0 + kb * BK, i0 + sg.row_shift, b);
            if (!w_done) tma_load_2d(sa + Cfg::kABytes, &p.tmW, full_bar(stage), kb_global * BK, n0);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(p.bf16, BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int st = cluster_id; st < total_super; st += n_clusters, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.nkb_total; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (it == 0 && kb == 0) stamp(2);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = umma_desc_kmajor_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom
            umma_f16_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (CL > 1) umma_commit_mc(empty_bar(stage), kMask);   // the stage is refilled by every CTA of the cluster
          else umma_commit(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        stamp(3);   // overwritten per tile: the last one stays
      }
    }
  } else {
    // ------- epilogue: warps 2..9; TMEM lane quadrant = warp % 4, column half = (warp-2)/4 -------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int bf = p.bf16;
    const bool is_gelu = p.act == ACT_GELU_TANH || p.act == ACT_GELU_ERF;
    const bool is_mul = p.act == ACT_MUL_GELU_TANH_GRAD || p.act == ACT_MUL_GELU_ERF_GRAD;
    // per-warp staging tile for TMA stores: [32 rows][64 B] (16-bit, 64B swizzle) or [32][128 B] (fp32, 128B swizzle)
    const uint32_t stg = staging_base + (uint32_t)(warp - 2) * 2048u;
    auto stage_h16 = [&](const float* x) {   // this lane's row: 32 x 16-bit = 4 x 16 B
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 w = pack8_h16(x + 8 * u, bf);
        const uint32_t addr = stg + (uint32_t)lane * 64u + (uint32_t)((u ^ ((lane >> 1) & 3)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
      }
    };
    auto stage_f32_half = [&](const float* x) {   // 16 fp32 of this lane's row = 4 x 16 B (64-byte rows, 64B swizzle)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t addr = stg + (uint32_t)lane * 64u + (uint32_t)((u ^ ((lane >> 1) & 3)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x[4 * u]), "f"(x[4 * u + 1]),
                     "f"(x[4 * u + 2]), "f"(x[4 * u + 3]) : "memory");
      }
    };
    auto stage_release = [&]() {   // the previous TMA store has finished reading the staging tile
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    };
    auto stage_store = [&](const CUtensorMap* tm, int col, int row0, int b) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_3d(tm, stg, col, row0, b); tma_store_commit(); }
    };
    // epi_direct: the warp's 32 x 64-byte slice is transposed through the staging tile (row-wise st.shared.v4, then
    // ld.shared.v4 with lane -> (row i*8 + lane/4, 16-byte unit lane%4)) and leaves as st.global.v4 that cover 8 rows x 64
    // contiguous bytes per instruction: 8 L1 wavefronts instead of 32, and no TMA round trip between chunks (the
    // 2 KB TMA stores made the epilogue a chain of ~1 us store latencies: 2.2 us per 128x128 tile against 0.9 us of MMA).
    auto lds_unit = [&](int r, int u) {
      uint4 v;
      const uint32_t addr = stg + (uint32_t)r * 64u + (uint32_t)((u ^ ((r >> 1) & 3)) << 4);
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
      return v;
    };
    // rows of this warp's slice: slice row r <-> tile row q*32 + r; writes 64 bytes per valid row at byte offset col_bytes
    auto store_slice = [&](void* base, long ld_bytes, long col_bytes, int b, int i0) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = k * 8 + (lane >> 2), u = lane & 3;
        const uint4 v = lds_unit(r, u);
        const int ii = i0 + q * 32 + r;
        const int orow = ii * p.rmul + p.roff;
        if (ii < p.R && orow < p.out_rows) {
          uint8_t* dst = reinterpret_cast<uint8_t*>(base) + ((long)b * p.out_rows + orow) * ld_bytes + col_bytes + u * 16;
          *reinterpret_cast<uint4*>(dst) = v;
        }
      }
      __syncwarp();   // the staging tile may be overwritten
    };
    int it = 0;
    for (int st = cluster_id; st < total_super; st += n_clusters, ++it) {
      const int acc = it & 1;
      const int mt = st / ntn_g, nt = (st - mt * ntn_g) * CL + rank;
      const int b = mt / p.tiles_per_batch;
      const int i0 = (mt - b * p.tiles_per_batch) * BM;
      const int n0 = nt * BN;
      const int i = i0 + q * 32 + lane;
      const int orow = i * p.rmul + p.roff;
      const bool valid = (i < p.R) && (orow < p.out_rows);
      const long frow = (long)b * p.out_rows + orow;
      const bool full_cols = n0 + BN <= p.n_valid;   // whole tile inside the valid columns
      const float rm = (valid && p.rowmask) ? p.rowmask[frow] : 1.f;

      auto prefetch = [&](int c, EpiPrefetch& pf) {
        const int nn = n0 + c * 32;
        if (nn >= p.n_valid || !valid) return;
        if (p.resid) {
          const uint4* rs = reinterpret_cast<const uint4*>(p.resid + frow * p.ldr + p.col_off + nn);
#pragma unroll
          for (int j = 0; j < 8; ++j) pf.v[j] = rs[j];
        } else if (is_mul) {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.mul_src) +
                                                            frow * p.ld_aux + nn);
#pragma unroll
          for (int j = 0; j < 4; ++j) pf.v[j] = __ldg(src + j);
        }
      };

      // DEEP: both chunks' residual / pre-activation operands are requested before the accumulator wait (registers allow
      // it at one CTA per SM); otherwise the next chunk's are requested after the current chunk's math
      constexpr int kNPf = DEEP ? 2 : 1;
      constexpr int kUnroll = DEEP ? 2 : 1;
      EpiPrefetch pfs[kNPf];
      float ln_s1 = 0.f;
      prefetch(half, pfs[0]);
      if constexpr (DEEP) prefetch(half + 2, pfs[1]);
      mbar_wait(tfull_bar(acc), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (it == 0 && threadIdx.x == 64) stamp(4);
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll kUnroll
      for (int cc = 0; cc < Cfg::kChunksPerWarp; ++cc) {
        const int c = cc * 2 + half;
        const int nn = n0 + c * 32;
        if (nn >= p.n_valid) break;  // warp-uniform
        EpiPrefetch& pf = pfs[DEEP ? cc : 0];
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32b_x32(t_acc + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (it == 0 && threadIdx.x == 64) stamp(cc == 0 ? 8 : 11);
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]);
        if (p.alpha != 1.f) {      // kernel-uniform; the estimator never scales (the epilogue is issue-bound: 32 FMULs matter)
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] *= p.alpha;
        }
        if (p.bias) {
          if (full_cols) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + nn);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(bp + j);
              x[4 * j + 0] += bv.x; x[4 * j + 1] += bv.y; x[4 * j + 2] += bv.z; x[4 * j + 3] += bv.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nn + j < p.n_valid) x[j] += __ldg(p.bias + nn + j);
          }
        }
        if (p.gn_part) {   // GroupNorm partial statistics of this 32-row x 32-channel (= one group) slice
          float s1 = 0.f, s2 = 0.f;
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { s1 += x[j]; s2 += x[j] * x[j]; }
          }
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          const float cnt = 32.f * (float)__popc(__ballot_sync(0xffffffffu, valid));
          if (lane == 0) {
            const int tib = mt - b * p.tiles_per_batch;
            float* o = p.gn_part + ((((long)b * p.tiles_per_batch + tib) * 4 + q) * (p.n_valid >> 5) + (nn >> 5)) * 3;
            const float mean = cnt > 0.f ? s1 / cnt : 0.f;
            o[0] = cnt; o[1] = mean; o[2] = cnt > 0.f ? fmaxf(s2 - s1 * mean, 0.f) : 0.f;
          }
        }
        if (p.epi_direct && is_gelu && p.aux_out) {   // pre-activation stash through the staging tile
          stage_h16(x);
          store_slice(p.aux_out, p.ld_aux * 2, (long)nn * 2, b, i0);
        } else if (p.tma_out && is_gelu && p.aux_out) {
          stage_release();
          stage_h16(x);
          stage_store(&p.tmAux, nn, i0 + q * 32, b);
        }
        if (valid || p.tma_out || p.epi_direct) {
          if (is_gelu) {
            if (p.aux_out && !p.tma_out && !p.epi_direct) {
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.aux_out) + frow * p.ld_aux + nn);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = pack8_h16(x + 8 * j, bf);
            }
            if (p.act == ACT_GELU_TANH) {   // packed fp32x2 arithmetic: half the issue slots
#pragma unroll
              for (int j = 0; j < 32; j += 2) f2_unpack(gelu_tanh_f2(f2_pack(x[j], x[j + 1])), x[j], x[j + 1]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = gelu_erf_f(x[j]);
            }
          } else if (is_mul) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 v = valid ? pf.v[j] : make_uint4(0u, 0u, 0u, 0u);
              float pre[8];
              unpack2_h16(v.x, bf, pre[0], pre[1]);
              unpack2_h16(v.y, bf, pre[2], pre[3]);
              unpack2_h16(v.z, bf, pre[4], pre[5]);
              unpack2_h16(v.w, bf, pre[6], pre[7]);
              if (p.act == ACT_MUL_GELU_TANH_GRAD) {   // kernel-uniform branch hoisted out of the element loop
#pragma unroll
                for (int e = 0; e < 8; e += 2)
                  f2_unpack(gelu_tanh_grad_mul_f2(f2_pack(x[8 * j + e], x[8 * j + e + 1]), f2_pack(pre[e], pre[e + 1])),
                            x[8 * j + e], x[8 * j + e + 1]);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[8 * j + e] *= gelu_erf_grad_f(pre[e]);
              }
            }
          }
          if (p.rowmask) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] *= rm;
          }
          if (p.resid && valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              x[4 * j + 0] += __uint_as_float(pf.v[j].x); x[4 * j + 1] += __uint_as_float(pf.v[j].y);
              x[4 * j + 2] += __uint_as_float(pf.v[j].z); x[4 * j + 3] += __uint_as_float(pf.v[j].w);
            }
          }
          if constexpr (BN == 256) {
            if (p.ln_gamma) {   // fused LayerNorm, pass 1: the finished fp32 row goes back into TMEM, its sum is taken
              uint32_t w0[16], w1[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) { w0[j] = __float_as_uint(x[j]); w1[j] = __float_as_uint(x[16 + j]); ln_s1 += x[j] + x[16 + j]; }
              tmem_st_32x32b_x16(t_acc + (uint32_t)(c * 32), w0);
              tmem_st_32x32b_x16(t_acc + (uint32_t)(c * 32 + 16), w1);
            }
          }
          if (!DEEP && cc + 1 < Cfg::kChunksPerWarp) prefetch(c + 2, pf);   // next chunk's operands fly during the stores
          if (it == 0 && threadIdx.x == 64 && cc == 0) stamp(9);
          if (p.epi_direct == 2) {
            // diagnostic build of the launch (CVFLOW_GEMM_EPI=2): no stores at all
          } else if (p.epi_direct) {
            if (p.out_f32) {   // two passes of 16 columns (64 bytes per row) through the 2 KB staging tile
              stage_f32_half(x);
              store_slice(p.out, p.ldc * 4, (long)(p.col_off + nn) * 4, b, i0);
              stage_f32_half(x + 16);
              store_slice(p.out, p.ldc * 4, (long)(p.col_off + nn + 16) * 4, b, i0);
            } else {
              stage_h16(x);
              store_slice(p.out, p.ldc * 2, (long)(p.col_off + nn) * 2, b, i0);
            }
          } else if (p.tma_out) {
            if (p.out_f32) {   // two passes of 16 columns through the 2 KB staging tile
              stage_release();
              stage_f32_half(x);
              stage_store(&p.tmOut, p.col_off + nn, i0 + q * 32, b);
              stage_release();
              stage_f32_half(x + 16);
              stage_store(&p.tmOut, p.col_off + nn + 16, i0 + q * 32, b);
            } else {
              stage_release();
              stage_h16(x);
              stage_store(&p.tmOut, p.col_off + nn, i0 + q * 32, b);
            }
          } else if (p.transposed_out) {
            if (valid) {
              float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nn + j < p.n_valid) o[((long)b * p.n_valid + nn + j) * p.out_rows + orow] = x[j];
            }
          } else if (p.out_f32) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + frow * p.ldc + p.col_off + nn);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(x[4 * j + 0], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + frow * p.ldc + p.col_off + nn);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = pack8_h16(x + 8 * j, bf);
          }
          if (it == 0 && threadIdx.x == 64) stamp(cc == 0 ? 10 : 12);
        }
      }
      if constexpr (BN == 256) {
        if (p.ln_gamma) {
          // fused LayerNorm, passes 2 and 3 over the row kept in TMEM (this warp: 32 rows x its 4 chunks = 128 columns;
          // the other column half of the same rows belongs to warp +-4: sums cross through shared memory)
          const uint32_t exch = bar_base + 128u;
          const int row = q * 32 + lane;
          tmem_st_wait();
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(exch + (uint32_t)(half * 128 + row) * 4u), "f"(ln_s1) : "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");
          float sa, sb;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sa) : "r"(exch + (uint32_t)row * 4u) : "memory");
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sb) : "r"(exch + (uint32_t)(128 + row) * 4u) : "memory");
          const float mean = (sa + sb) * (1.f / 256.f);
          float s2 = 0.f;
#pragma unroll 1
          for (int cc = 0; cc < Cfg::kChunksPerWarp; ++cc) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_acc + (uint32_t)((cc * 2 + half) * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float d = __uint_as_float(r[j]) - mean; s2 = fmaf(d, d, s2); }
          }
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(exch + (uint32_t)(256 + half * 128 + row) * 4u), "f"(s2) : "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sa) : "r"(exch + (uint32_t)(256 + row) * 4u) : "memory");
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sb) : "r"(exch + (uint32_t)(384 + row) * 4u) : "memory");
          const float rstd = rsqrtf((sa + sb) * (1.f / 256.f) + 1e-5f);
#pragma unroll 1
          for (int cc = 0; cc < Cfg::kChunksPerWarp; ++cc) {
            const int nn = (cc * 2 + half) * 32;
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_acc + (uint32_t)nn, r);
            tmem_ld_wait();
            float y[32];
            const float4* gp = reinterpret_cast<const float4*>(p.ln_gamma + nn);
            const float4* bp = reinterpret_cast<const float4*>(p.ln_beta + nn);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g4 = __ldg(gp + j), b4 = __ldg(bp + j);
              y[4 * j + 0] = (__uint_as_float(r[4 * j + 0]) - mean) * rstd * g4.x + b4.x;
              y[4 * j + 1] = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * g4.y + b4.y;
              y[4 * j + 2] = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * g4.z + b4.z;
              y[4 * j + 3] = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * g4.w + b4.w;
            }
            if (p.epi_direct) {
              stage_h16(y);
              store_slice(p.aux_out, p.ld_aux * 2, (long)nn * 2, b, i0);
            } else {
              stage_release();
              stage_h16(y);
              stage_store(&p.tmAux, nn, i0 + q * 32, b);
            }
          }
          // the exchange slots are reused by the next tile: every warp has read them before anyone writes again
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      // all of this thread's TMEM reads of the buffer are complete: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (it == 0 && threadIdx.x == 64) stamp(13);
      if (threadIdx.x == 64) { stamp(5); if (dbg) dbg[7] = it + 1; }
    }
  }
  pdl_launch();   // dependents are released late: CTAs of the next kernel that spin at their grid-dependency wait next to the working ones cost more than their prologue overlap gains (same-box A/B)
  if (warp >= 2 && lane == 0) tma_store_wait_all();   // staging smem must outlive the bulk stores
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into / arrive on its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * Cfg::kTmemCols);
  }
  if (threadIdx.x == 32) stamp(6);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int tma_encode_3d_ex(CUtensorMap* tm, const void* base, int dtype, int swizzle_bytes, uint64_t d0, uint64_t d1,
                     uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1,
                     uint32_t b2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -100;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dtc = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                             : (dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(tm, dtc, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}
int tma_encode_3d(CUtensorMap* tm, const void* base, int bf16, uint64_t d0, uint64_t d1, uint64_t d2,
                  uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1,
                  uint32_t b2) {
  return tma_encode_3d_ex(tm, base, bf16 ? 1 : 0, 128, d0, d1, d2, stride1_bytes, stride2_bytes, b0, b1, b2);
}

int tma_encode_2d_ex(CUtensorMap* tm, const void* base, int dtype, int swizzle_bytes, uint64_t d0, uint64_t d1,
                     uint64_t stride1_bytes, uint32_t b0, uint32_t b1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -100;
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dtc = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                             : (dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(tm, dtc, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

static int encode_2d(CUtensorMap* tm, const void* base, int bf16, uint64_t d0, uint64_t d1,
                     uint64_t stride1_bytes, uint32_t b0, uint32_t b1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -100;
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

#define GEMM_FAIL(...)                          \
  do {                                          \
    if (err) snprintf(err, errlen, __VA_ARGS__); \
    return -1;                                  \
  } while (0)

int gemm_prepare(const GemmArgs& a, GemmParams* p, char* err, int errlen) {
  memset(p, 0, sizeof(*p));
  if (a.nseg < 1 || a.nseg > 8) GEMM_FAIL("gemm: nseg %d out of range", a.nseg);
  if (a.Ktot % BK) GEMM_FAIL("gemm: Ktot %d not a multiple of %d", a.Ktot, BK);
  if (!a.out || !a.W || !a.A[0]) GEMM_FAIL("gemm: null operand");
  int nkb = 0;
  bool use1 = false;
  for (int s = 0; s < a.nseg; ++s) {
    const GemmSeg& sg = a.seg[s];
    if (sg.a_map < 0 || sg.a_map > 1 || !a.A[sg.a_map]) GEMM_FAIL("gemm: segment %d bad source", s);
    if (sg.a_col0 % 8 || sg.a_col0 + sg.nkb * BK > a.a_cols[sg.a_map])
      GEMM_FAIL("gemm: segment %d columns [%d,+%d) outside source (%d cols)", s, sg.a_col0,
                sg.nkb * BK, a.a_cols[sg.a_map]);
    use1 |= sg.a_map == 1;
    nkb += sg.nkb;
    p->seg[s] = sg;
  }
  if (nkb * BK != a.Ktot) GEMM_FAIL("gemm: segments cover %d columns, Ktot %d", nkb * BK, a.Ktot);
  if (!a.transposed_out) {
    if (a.n_valid % 32) GEMM_FAIL("gemm: n_valid %d must be a multiple of 32", a.n_valid);
    if ((a.ldc % 8) || (a.col_off % 8)) GEMM_FAIL("gemm: ldc/col_off must be multiples of 8");
  }
  if (a.n_valid > a.N && !a.transposed_out) GEMM_FAIL("gemm: n_valid > N");
  p->nseg = a.nseg;
  p->nkb_total = nkb;
  p->src_A[0] = a.A[0]; p->src_A[1] = a.A[1]; p->src_W = a.W;
  p->dbg = a.dbg;
  p->bf16 = a.bf16;
  p->R = a.R; p->rmul = a.rmul; p->roff = a.roff; p->out_rows = a.out_rows; p->nbatch = a.nbatch;
  p->tiles_per_batch = (a.R + BM - 1) / BM;
  p->out = a.out; p->out_f32 = a.out_f32 || a.transposed_out; p->ldc = a.ldc; p->col_off = a.col_off;
  p->n_valid = a.n_valid; p->transposed_out = a.transposed_out;
  p->alpha = a.alpha; p->bias = a.bias; p->act = a.act; p->aux_out = a.aux_out;
  p->mul_src = a.mul_src; p->ld_aux = a.ld_aux; p->rowmask = a.rowmask; p->resid = a.resid;
  p->ldr = a.ldr;
  p->gn_part = a.gn_part;
  p->w_rows = a.N;
  static int wpre_env = -1;
  if (wpre_env < 0) { const char* e = getenv("CVFLOW_GEMM_WPRE"); wpre_env = e ? atoi(e) : 1; }
  p->w_prefetch = (a.w_static && wpre_env) ? 1 : 0;
  p->ln_gamma = a.ln_gamma; p->ln_beta = a.ln_beta;
  if (a.ln_gamma) {
    if (!a.ln_beta || !a.aux_out || a.N != 256 || a.n_valid != 256 || a.transposed_out || !a.out_f32 || a.act != ACT_NONE ||
        a.rmul != 1 || a.roff != 0 || a.col_off != 0)
      GEMM_FAIL("gemm: fused LayerNorm needs N = n_valid = 256, fp32 row-major output, no activation, aux_out = 16-bit x~");
  }
  if (a.gn_part && (a.transposed_out || a.n_valid % 32)) GEMM_FAIL("gemm: gn_part needs a row-major output with n_valid %% 32 == 0");
  // tile shape: 256-wide tiles only when N divides evenly and there is more than a wave of them
  const long mtiles = (long)p->tiles_per_batch * a.nbatch;
  // tile width: these GEMMs are short (K = 256..1536) and latency-bound, so prefer enough CTAs to
  // keep >= 2 resident per SM over wide tiles
  static int bn256_min = -1;   // tiles needed before 256-wide tiles are used (tuning knob)
  if (bn256_min < 0) { const char* e = getenv("CVFLOW_GEMM_BN256_MIN"); bn256_min = e ? atoi(e) : 4 * 148; }
  int bn = 64;
  if (a.n_valid % 256 == 0 && mtiles * (a.n_valid / 256) >= bn256_min) bn = 256;
  static int bn128_min = -1;   // tuning knob (CVFLOW_GEMM_BN128_MIN)
  if (bn128_min < 0) { const char* e = getenv("CVFLOW_GEMM_BN128_MIN"); bn128_min = e ? atoi(e) : 2 * 148; }
  if (bn == 64 && (mtiles * ((a.n_valid + 127) / 128) >= bn128_min || a.n_valid > 512)) bn = 128;
  if (a.transposed_out) bn = 128;
  if (a.ln_gamma) bn = 256;      // the tile must own whole rows
  // deep pipeline for the narrow-N, long-K shapes (N <= 256): see GemmCfg
  static int deep_env = -1;
  if (deep_env < 0) { const char* e = getenv("CVFLOW_GEMM_DEEP"); deep_env = e ? atoi(e) : 0; }   // measured: faster stand-alone (FF2 10.1 -> 9.7, q/k/v dgrad 9.7 -> 8.7 us), slower inside the PDL-chained step (16.17 -> 16.75 ms: 100 single-CTA SMs with 208 KB of shared memory keep the next kernel's CTAs out)
  p->deep = 0;
  if (deep_env && bn == 64 && !a.transposed_out && !a.ln_gamma && a.n_valid % 128 == 0 && a.n_valid <= 256 && a.Ktot >= 512) {
    bn = 128;
    p->deep = 1;
  }
  p->block_n = bn;
  p->grid_x = (int)mtiles;
  p->grid_y = (a.n_valid + bn - 1) / bn;
  // cluster width (A-tile multicast along N): opt-in with CVFLOW_GEMM_CLUSTER=2|4. Measured on the estimator's shapes it
  // does not pay (these launches are latency- not L2-bandwidth-bound: 9.7 -> 10.3 us for the q/k/v GEMM), so the default is 1.
  static int cl_env = -1;
  if (cl_env < 0) { const char* e = getenv("CVFLOW_GEMM_CLUSTER"); cl_env = e ? atoi(e) : 1; }
  int cl = 1;
  if (bn != 256 && !a.transposed_out && !p->deep) cl = p->grid_y >= 4 ? 4 : (p->grid_y >= 2 ? 2 : 1);
  if (bn == 128 && cl > 2) cl = 2;
  if (cl_env != 2 && cl_env != 4) cl = 1;
  else if (cl_env < cl) cl = cl_env;
  p->cluster = cl;
  for (int s = 0; s < 2; ++s) {
    if (!a.A[s] || (s == 1 && !use1)) continue;
    if ((reinterpret_cast<uintptr_t>(a.A[s]) & 15) || (a.a_ld[s] % 8) || (a.a_bstride[s] % 8))
      GEMM_FAIL("gemm: A[%d] must be 16-byte aligned with strides multiple of 8 elements", s);
    int r = tma_encode_3d(&p->tmA[s], a.A[s], a.bf16, (uint64_t)a.a_cols[s], (uint64_t)a.a_rows[s],
                          (uint64_t)a.nbatch, (uint64_t)a.a_ld[s] * 2,
                          (uint64_t)(a.nbatch > 1 ? a.a_bstride[s] : a.a_ld[s] * (long)a.a_rows[s]) * 2,
                          BK, BM / p->cluster, 1);
    if (r) GEMM_FAIL("gemm: cuTensorMapEncodeTiled(A[%d]) failed (%d)", s, r);
  }
  if (!use1) p->tmA[1] = p->tmA[0];
  // output tensor maps for the TMA-store epilogue: rows i -> i*rmul + roff as a strided row view
  p->tma_out = 0;
  p->epi_direct = 0;
  static int epi_env = -1;
  if (epi_env < 0) { const char* e = getenv("CVFLOW_GEMM_EPI"); epi_env = e ? atoi(e) : 0; }   // measured on B200: the TMA-store epilogue is faster (q/k/v 10.2 vs 12.5 us)
  if (!a.transposed_out && epi_env >= 1) {
    const int esz = p->out_f32 ? 4 : 2;
    const bool ok = ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) && ((a.ldc * esz) % 16 == 0) &&
                    (!a.aux_out || (((reinterpret_cast<uintptr_t>(a.aux_out) & 15) == 0) && (a.ld_aux % 8 == 0)));
    if (a.aux_out && (a.rmul != 1 || a.roff != 0)) GEMM_FAIL("gemm: aux_out needs contiguous output rows");
    p->epi_direct = ok ? epi_env : 0;
  }
  if (!a.transposed_out && !p->epi_direct) {
    const int esz = p->out_f32 ? 4 : 2;
    long rows_view = (a.out_rows - a.roff + a.rmul - 1) / a.rmul;
    if (rows_view > a.R) rows_view = a.R;
    if (rows_view < 1) rows_view = 1;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a.out) + (long)a.roff * a.ldc * esz;
    const bool ok = ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((a.ldc * esz) % 16 == 0);
    if (ok) {
      int r2 = tma_encode_3d_ex(&p->tmOut, base, p->out_f32 ? 2 : (a.bf16 ? 1 : 0), 64,
                                (uint64_t)(a.col_off + a.n_valid), (uint64_t)rows_view, (uint64_t)a.nbatch,
                                (uint64_t)a.rmul * a.ldc * esz,
                                (uint64_t)(a.nbatch > 1 ? (long)a.out_rows * a.ldc * esz : (long)a.rmul * a.ldc * esz * rows_view),
                                p->out_f32 ? 16 : 32, 32, 1);
      if (r2) GEMM_FAIL("gemm: cuTensorMapEncodeTiled(out) failed (%d)", r2);
      p->tma_out = 1;
      if (a.aux_out) {
        if (a.rmul != 1 || a.roff != 0) GEMM_FAIL("gemm: aux_out needs contiguous output rows");
        r2 = tma_encode_3d_ex(&p->tmAux, a.aux_out, a.bf16 ? 1 : 0, 64, (uint64_t)a.n_valid, (uint64_t)rows_view,
                              (uint64_t)a.nbatch, (uint64_t)a.ld_aux * 2,
                              (uint64_t)(a.nbatch > 1 ? (long)a.out_rows * a.ld_aux * 2 : (long)a.ld_aux * 2 * rows_view),
                              32, 32, 1);
        if (r2) GEMM_FAIL("gemm: cuTensorMapEncodeTiled(aux) failed (%d)", r2);
      }
    }
  }
  if (a.ln_gamma && !p->tma_out && !p->epi_direct) GEMM_FAIL("gemm: fused LayerNorm needs 16-byte aligned outputs");
  if (reinterpret_cast<uintptr_t>(a.W) & 15) GEMM_FAIL("gemm: W must be 16-byte aligned");
  int r = encode_2d(&p->tmW, a.W, a.bf16, (uint64_t)a.Ktot, (uint64_t)a.N, (uint64_t)a.Ktot * 2, BK,
                    bn);
  if (r) GEMM_FAIL("gemm: cuTensorMapEncodeTiled(W) failed (%d)", r);
  return 0;
}

int gemm_launch(const GemmParams& p, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_tc_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::kSmemBytes);
    cudaFuncSetAttribute(gemm_tc_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::kSmemBytes);
    cudaFuncSetAttribute(gemm_tc_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::kSmemBytes);
    cudaFuncSetAttribute(gemm_tc_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::kSmemBytes);
    cudaFuncSetAttribute(gemm_tc_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::kSmemBytes);
    cudaFuncSetAttribute(gemm_tc_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256>::kSmemBytes);
    cudaFuncSetAttribute(gemm_tc_kernel<128, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128, 1>::kSmemBytes);
    attr_done = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int cl = p.cluster;
  const long total = (long)p.grid_x * ((p.grid_y + cl - 1) / cl) * cl;   // CTAs if every (m-tile, n-group) had its own cluster
  long slots = (long)num_sms * ((p.block_n == 256 || p.deep) ? 1 : 2);
  slots -= slots % cl;
  dim3 grid((unsigned)(total < slots ? total : slots));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(kGemmThreads); cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = (unsigned)cl; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = cl > 1 ? 2 : 1;
#define CVFLOW_GEMM_LAUNCH(BN_, CL_)                               \
  do {                                                             \
    cfg.dynamicSmemBytes = GemmCfg<BN_>::kSmemBytes;               \
    cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN_, CL_>, p);         \
  } while (0)
  if (p.deep) {
    cfg.dynamicSmemBytes = GemmCfg<128, 1>::kSmemBytes;
    cudaLaunchKernelEx(&cfg, gemm_tc_kernel<128, 1, 1>, p);
  } else if (p.block_n == 256) CVFLOW_GEMM_LAUNCH(256, 1);
  else if (p.block_n == 128) { if (cl == 2) CVFLOW_GEMM_LAUNCH(128, 2); else CVFLOW_GEMM_LAUNCH(128, 1); }
  else { if (cl == 4) CVFLOW_GEMM_LAUNCH(64, 4); else if (cl == 2) CVFLOW_GEMM_LAUNCH(64, 2); else CVFLOW_GEMM_LAUNCH(64, 1); }
#undef CVFLOW_GEMM_LAUNCH
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace cvflow
