// Fused feed-forward of the transformer blocks on tcgen05 (sm_100a): both GEMMs of
// FeedForward (reference modules.py:192-224: Linear(256,1024) -> GELU(tanh) -> Linear(1024,256))
// and the residual add of modules.py:372-374 in ONE kernel, and likewise its backward
// (dpre = (dh W2) o gelu'(pre); dx = dpre W1) -- the [M,1024] hidden activation never leaves the SM.
//
// One CTA owns a 128-row tile. The hidden dimension is walked in 8 chunks of 128:
//   G1(c): acc1[c&1] = X[128x256] W1_c^T         A = X tile resident in smem, B = W1 rows 128c.. streamed by TMA
//   E1(c): y = act(acc1 + b1)                     8 epilogue warps, thread = row; the 16-bit result is written back
//                                                 over acc1 with tcgen05.st and is the A operand of G2 (TS MMA)
//   G2(c): acc2 += Y_c[128x128] W2[:,128c..]^T    accumulated in TMEM over the chunks
// acc1 is double-buffered so E1(c) overlaps G1(c+1) and G2(c-1); TMEM = 2x128 + 256 = 512 columns.
// Forward E1 also stashes the pre-activation (16-bit) for the backward pass; backward E1 multiplies
// by gelu'(pre). The final epilogue adds bias + fp32 residual (forward) and stores through TMA.
#include "kernels.h"
#include "gemm.h"
#include "common.cuh"
#include "attention.h"   // umma_f16_ts, tmem_st_*
#include <string.h>
#include <stdlib.h>

namespace cvflow {

struct MlpParams {
  CUtensorMap tmX;     // 16-bit [M][256], box {64, 128}
  CUtensorMap tmW1;    // 16-bit [1024][256], box {64, 128}
  CUtensorMap tmW2;    // 16-bit [256][1024], box {64, 128}
  CUtensorMap tmAux;   // forward: pre-activation stash 16-bit [M][1024], box {32, 32} (64B swizzle)
  CUtensorMap tmOut;   // forward: fp32 [M][256], box {16, 32}; backward: 16-bit [M][256], box {32, 32} (64B swizzle)
  const float* b1;     // [1024] (forward) or null
  const float* b2;     // [256] (forward) or null
  const float* resid;  // fp32 [M][256] (forward) or null
  const uint16_t* pre; // backward: stashed pre-activation [M][1024]
  uint16_t* pre_out;   // forward: where the pre-activation is stashed
  void* out;
  int M, backward, bf16, gelu_erf;
  const void* x_ptr; const void* out_ptr; const void* aux_ptr;
  long long* dbg;      // optional: 64 clock64 stamps per CTA (profiling aid)
};
int mlp_plan_bytes() { return (int)sizeof(MlpParams); }
static long long* g_mlp_dbg = nullptr;
void mlp_set_debug_buffer(void* p) { g_mlp_dbg = reinterpret_cast<long long*>(p); }

struct MlpSmem {
  static constexpr int kX = 0;               // 4 k-blocks x [128 rows x 128 B] = 64 KB
  static constexpr int kRing = 65536;        // 8 units x 16 KB
  static constexpr int kUnits = 8;
  static constexpr int kStg = 196608;        // 16 epilogue warps x 2 KB staging tiles for TMA stores
  static constexpr int kBar = 229376;
  static constexpr int kBytes = kBar + 256 + 1024;
};
static constexpr int kMlpThreads = 576;      // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue (4 per TMEM lane quadrant)
static constexpr int kMlpEpiThreads = 512;

// SPLIT = 2: a cluster of two CTAs shares one 128-row tile. CTA r takes hidden chunks [4r, 4r+4) -- half the weight
// stream per SM (the stream is L2->SM bandwidth-bound) and twice as many busy SMs -- and the two partial 128x256
// results are reduced through distributed shared memory: each CTA ships the half it does not finalise to its peer.
template <int SPLIT>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_tc_kernel(const __grid_constant__ MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + MlpSmem::kBar;
  auto full = [&](int u) { return bar + 8u * u; };
  auto empty = [&](int u) { return bar + 64u + 8u * u; };
  const uint32_t x_full = bar + 128u;
  auto a1_full = [&](int b) { return bar + 144u + 8u * b; };
  auto p_full = [&](int b) { return bar + 160u + 8u * b; };
  const uint32_t acc2_full = bar + 176u;
  const uint32_t tmem_slot = bar + 192u;
  constexpr int NCH = 8 / SPLIT;   // hidden chunks of this CTA

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bf = p.bf16;
  const int rank = SPLIT > 1 ? (int)cluster_ctarank() : 0;
  const int tile = blockIdx.x / SPLIT;
  const int ch0 = rank * NCH;      // first global chunk
  long long* dbg = p.dbg ? p.dbg + (long)blockIdx.x * 64 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();

  if (threadIdx.x == 0) {
    for (int u = 0; u < MlpSmem::kUnits; ++u) { mbar_init(full(u), 1); mbar_init(empty(u), 1); }
    mbar_init(x_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(a1_full(b), 1); mbar_init(p_full(b), kMlpEpiThreads); }
    mbar_init(acc2_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW1); tma_prefetch_desc(&p.tmW2); tma_prefetch_desc(&p.tmOut);
    if (!p.backward) tma_prefetch_desc(&p.tmAux);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t t_acc2 = tmem + 256u;
  pdl_wait();
  pdl_launch();

  // epilogue-warp identity (warps 2..17): TMEM lane quadrant q, 32-column group sub of a chunk / 64-column group of the result
  const int q = warp & 3;
  const int sub = (warp - 2) >> 2;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t stg = base + MlpSmem::kStg + (uint32_t)(warp >= 2 ? warp - 2 : 0) * 2048u;
  const int row0 = tile * 128 + q * 32;
  const long row = (long)row0 + lane;
  const bool valid = row < p.M;
  auto stage_h16 = [&](const float* x) {   // this lane's row: 32 x 16-bit = 4 x 16 B (64-byte rows, 64B swizzle)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 wv = pack8_h16(x + 8 * u, bf);
      const uint32_t addr = stg + (uint32_t)lane * 64u + (uint32_t)((u ^ ((lane >> 1) & 3)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv.x), "r"(wv.y), "r"(wv.z), "r"(wv.w) : "memory");
    }
  };
  auto stage_f32_half = [&](const float* x) {   // 16 fp32 of this lane's row
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t addr = stg + (uint32_t)lane * 64u + (uint32_t)((u ^ ((lane >> 1) & 3)) << 4);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x[4 * u]), "f"(x[4 * u + 1]),
                   "f"(x[4 * u + 2]), "f"(x[4 * u + 3]) : "memory");
    }
  };
  auto stage_release = [&]() {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  };
  auto stage_store = [&](const CUtensorMap* tm, int col, int r0) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) { tma_store_2d(tm, stg, col, r0); tma_store_commit(); }
  };

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int u = 0;
      uint32_t uph = 0;
      auto load_unit = [&](const CUtensorMap* tm, int col, int r) {
        mbar_wait(empty(u), uph ^ 1u);
        mbar_expect_tx(full(u), 16384u);
        tma_load_2d(base + MlpSmem::kRing + u * 16384, tm, full(u), col, r);
        if (++u == MlpSmem::kUnits) { u = 0; uph ^= 1u; }
      };
      mbar_expect_tx(x_full, 65536u);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(base + MlpSmem::kX + kb * 16384, &p.tmX, x_full, kb * 64, tile * 128);
      for (int s = 0; s <= NCH; ++s) {
        if (s < NCH)
          for (int kb = 0; kb < 4; ++kb) load_unit(&p.tmW1, kb * 64, (ch0 + s) * 128);
        if (s >= 1) {
          const int c = ch0 + s - 1;
          for (int kb = 0; kb < 2; ++kb) {
            load_unit(&p.tmW2, c * 128 + kb * 64, 0);
            load_unit(&p.tmW2, c * 128 + kb * 64, 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_f16(bf, 128, 128, 0, 0);
      const uint32_t idesc2 = umma_idesc_f16(bf, 128, 256, 0, 0);
      int u = 0;
      uint32_t uph = 0;
      mbar_wait(x_full, 0);
      tc_fence_after();
      if (dbg) dbg[1] = clock64();
      for (int s = 0; s <= NCH; ++s) {
        if (s < NCH) {
          const uint32_t d = tmem + (uint32_t)((s & 1) * 128);
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(full(u), uph);
            tc_fence_after();
            const uint64_t da = umma_desc_kmajor_sw128(base + MlpSmem::kX + kb * 16384);
            const uint64_t db = umma_desc_kmajor_sw128(base + MlpSmem::kRing + u * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(d, da + 2 * k, db + 2 * k, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(empty(u));
            if (++u == MlpSmem::kUnits) { u = 0; uph ^= 1u; }
          }
          umma_commit(a1_full(s & 1));
          if (dbg) dbg[2 + 2 * s] = clock64();
        }
        if (s >= 1) {
          const int g = s - 1;
          mbar_wait(p_full(g & 1), (uint32_t)((g >> 1) & 1));
          if (dbg) dbg[3 + 2 * g] = clock64();
          tc_fence_after();
          const uint32_t a2 = tmem + (uint32_t)((g & 1) * 128);
          for (int kb = 0; kb < 2; ++kb) {
            mbar_wait(full(u), uph);
            const int u2 = u + 1;   // units come in even/odd pairs: 256 contiguous rows of W2
            mbar_wait(full(u2), uph);
            tc_fence_after();
            const uint64_t db = umma_desc_kmajor_sw128(base + MlpSmem::kRing + u * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(t_acc2, a2 + (uint32_t)(32 * (kb * 2 + (k >> 1)) + 8 * (k & 1)), db + 2 * k, idesc2,
                          (s > 1 || kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(empty(u));
            umma_commit(empty(u2));
            u += 2;
            if (u == MlpSmem::kUnits) { u = 0; uph ^= 1u; }
          }
        }
      }
      umma_commit(acc2_full);
    }
  } else {
    // ---------------- epilogue warps, part 1: activation of every hidden chunk ----------------
    for (int c = 0; c < NCH; ++c) {
      const uint32_t t_a1 = tmem + (uint32_t)((c & 1) * 128) + lane_addr + (uint32_t)(sub * 32);
      const int col0 = (ch0 + c) * 128 + sub * 32;   // hidden columns of this warp's group
      uint4 pre_v[4];
      if (p.backward) {   // stashed pre-activation of this thread's 32 columns: issued before the accumulator wait
        const uint4* src = reinterpret_cast<const uint4*>(p.pre + row * 1024 + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) pre_v[j] = valid ? __ldg(src + j) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(a1_full(c & 1), (uint32_t)((c >> 1) & 1));
      if (dbg && threadIdx.x == 64) dbg[20 + 2 * c] = clock64();
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_a1, r);
      tmem_ld_wait();
      uint32_t pk[16];
      if (!p.backward) {
        float x[32];
        const float4* bp = reinterpret_cast<const float4*>(p.b1 + col0);   // L1-resident broadcast loads
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = __ldg(bp + j);
          x[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bv.x; x[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bv.y;
          x[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bv.z; x[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bv.w;
        }
        if (p.gelu_erf != 3) {   // pre-activation stash for the backward pass (3 = profiling aid: no stash, identity)
          stage_release();
          stage_h16(x);
          stage_store(&p.tmAux, col0, row0);
        }
        if (p.gelu_erf >= 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack2_h16(x[2 * j], x[2 * j + 1], bf);
        } else if (p.gelu_erf) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack2_h16(gelu_erf_f(x[2 * j]), gelu_erf_f(x[2 * j + 1]), bf);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float y0, y1;
            f2_unpack(gelu_tanh_f2(f2_pack(x[2 * j], x[2 * j + 1])), y0, y1);
            pk[j] = pack2_h16(y0, y1, bf);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 v = pre_v[j];
          float pre[8];
          unpack2_h16(v.x, bf, pre[0], pre[1]); unpack2_h16(v.y, bf, pre[2], pre[3]);
          unpack2_h16(v.z, bf, pre[4], pre[5]); unpack2_h16(v.w, bf, pre[6], pre[7]);
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const float d0 = __uint_as_float(r[8 * j + e]), d1 = __uint_as_float(r[8 * j + e + 1]);
            float y0, y1;
            if (p.gelu_erf) {
              y0 = d0 * gelu_erf_grad_f(pre[e]);
              y1 = d1 * gelu_erf_grad_f(pre[e + 1]);
            } else {
              f2_unpack(gelu_tanh_grad_mul_f2(f2_pack(d0, d1), f2_pack(pre[e], pre[e + 1])), y0, y1);
            }
            pk[4 * j + (e >> 1)] = pack2_h16(y0, y1, bf);
          }
        }
      }
      // A operand of G2: K elements [32 sub, 32 sub + 32) of the chunk go to the first 16 columns of this warp's own
      // (already consumed) accumulator group
      tmem_st_32x32b_x16(t_a1, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(c & 1));
      if (dbg && threadIdx.x == 64) dbg[21 + 2 * c] = clock64();
    }
    mbar_wait(acc2_full, 0);
    if (dbg && threadIdx.x == 64) dbg[40] = clock64();
    tc_fence_after();
  }

  // ---------------- final epilogue ----------------
  // acc2 (+ peer partial + b2 + residual) -> global through the staging tile. Without the split every epilogue warp owns
  // columns [64 sub, +64). With it, CTA r finalises columns [128 r, +128): its warps with sub in {2r, 2r+1} keep their
  // columns, the other eight ship theirs to the peer (column-major fp32 [128 cols][128 rows] in the idle weight ring).
  const bool own = SPLIT == 1 || (sub >> 1) == rank;
  if (SPLIT > 1) {
    cluster_sync_all();   // both CTAs are past their last MMA: the weight rings are free to receive
    if (warp >= 2 && !own) {
      const uint32_t remote = dsmem_map(base + MlpSmem::kRing, (uint32_t)(rank ^ 1));
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_acc2 + lane_addr + (uint32_t)((sub * 2 + i) * 32), r);
        tmem_ld_wait();
        const uint32_t colbase = (uint32_t)(((sub & 1) * 2 + i) * 32);   // column within the peer's 128-column half
#pragma unroll
        for (int j = 0; j < 32; ++j)
          dsmem_st_u32(remote + ((colbase + (uint32_t)j) * 128u + (uint32_t)(q * 32 + lane)) * 4u, r[j]);
      }
    }
    cluster_sync_all();   // partials have landed
  }
  if (warp >= 2 && own) {
    const float* recv = reinterpret_cast<const float*>(gbase + MlpSmem::kRing);
    float4 rv[8];   // fp32 residual of the next 32-column group: in flight while the current one is processed
    auto load_resid = [&](int g) {
      const float4* rs = reinterpret_cast<const float4*>(p.resid + row * 256 + g * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) rv[j] = (p.resid && valid) ? __ldg(rs + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (!p.backward) load_resid(sub * 2);
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const int g = sub * 2 + i;
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_acc2 + lane_addr + (uint32_t)(g * 32), r);
      tmem_ld_wait();
      float x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]);
      if (SPLIT > 1) {
        const int colbase = ((sub & 1) * 2 + i) * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] += recv[(colbase + j) * 128 + q * 32 + lane];
      }
      if (!p.backward) {
        const float4* bp = reinterpret_cast<const float4*>(p.b2 + g * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = __ldg(bp + j);
          x[4 * j + 0] += bv.x + rv[j].x; x[4 * j + 1] += bv.y + rv[j].y;
          x[4 * j + 2] += bv.z + rv[j].z; x[4 * j + 3] += bv.w + rv[j].w;
        }
        if (i == 0) load_resid(g + 1);
        stage_release();
        stage_f32_half(x);
        stage_store(&p.tmOut, g * 32, row0);
        stage_release();
        stage_f32_half(x + 16);
        stage_store(&p.tmOut, g * 32 + 16, row0);
      } else {
        stage_release();
        stage_h16(x);
        stage_store(&p.tmOut, g * 32, row0);
      }
    }
    if (dbg && threadIdx.x == 64) dbg[41] = clock64();
  }
  if (warp >= 2 && lane == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (SPLIT > 1) cluster_sync_all();   // the peer may still be reading what this CTA shipped / vice versa
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int tma_encode_2d_ex(CUtensorMap* tm, const void* base, int dtype, int swizzle_bytes, uint64_t d0, uint64_t d1,
                     uint64_t stride1_bytes, uint32_t b0, uint32_t b1);

int mlp_prepare(void* plan_, int backward, const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                const float* resid, void* out, void* pre, long M, int bf16, int gelu_erf, char* err, int errlen) {
  MlpParams* p = reinterpret_cast<MlpParams*>(plan_);
  memset(p, 0, sizeof(*p));
  p->M = (int)M; p->backward = backward; p->bf16 = bf16; p->gelu_erf = gelu_erf;
  p->b1 = b1; p->b2 = b2; p->resid = resid; p->pre = reinterpret_cast<const uint16_t*>(pre);
  p->x_ptr = x; p->out_ptr = out; p->aux_ptr = pre;
  p->pre_out = reinterpret_cast<uint16_t*>(pre); p->out = out;
  p->dbg = g_mlp_dbg;
  const int dt = bf16 ? 1 : 0;
  int r = tma_encode_2d_ex(&p->tmX, x, dt, 128, 256, (uint64_t)M, 512, 64, 128);
  if (!r) r = tma_encode_2d_ex(&p->tmW1, w1, dt, 128, 256, 1024, 512, 64, 128);
  if (!r) r = tma_encode_2d_ex(&p->tmW2, w2, dt, 128, 1024, 256, 2048, 64, 128);
  if (!r && !backward) r = tma_encode_2d_ex(&p->tmAux, pre, dt, 64, 1024, (uint64_t)M, 2048, 32, 32);
  if (!r) r = backward ? tma_encode_2d_ex(&p->tmOut, out, dt, 64, 256, (uint64_t)M, 512, 32, 32)
                       : tma_encode_2d_ex(&p->tmOut, out, 2, 64, 256, (uint64_t)M, 1024, 16, 32);
  if (r) { if (err) snprintf(err, errlen, "mlp: cuTensorMapEncodeTiled failed (%d)", r); return -1; }
  return 0;
}

int mlp_launch(const void* plan_, cudaStream_t st) {
  const MlpParams* p = reinterpret_cast<const MlpParams*>(plan_);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(mlp_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpSmem::kBytes);
    cudaFuncSetAttribute(mlp_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpSmem::kBytes);
    attr_done = true;
  }
  const int ntiles = (p->M + 127) / 128;
  // split the hidden dimension over a CTA pair while that still fits one wave (it halves the per-SM weight stream)
  static int split_env = -1;
  if (split_env < 0) { const char* e = getenv("CVFLOW_MLP_SPLIT"); split_env = e ? atoi(e) : 0; }
  const int split = split_env == 1 ? 1 : (split_env == 2 ? 2 : (2 * ntiles <= attn_num_sms() ? 2 : 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(ntiles * split)); cfg.blockDim = dim3(kMlpThreads); cfg.stream = st;
  cfg.dynamicSmemBytes = MlpSmem::kBytes;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = split > 1 ? 2 : 1;
  if (split > 1) cudaLaunchKernelEx(&cfg, mlp_tc_kernel<2>, *p);
  else cudaLaunchKernelEx(&cfg, mlp_tc_kernel<1>, *p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace cvflow
