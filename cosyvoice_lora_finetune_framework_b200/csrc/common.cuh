// Shared device helpers for the cvflow sm_100a kernels: 16-bit operand conversion, warp
// reductions, mbarrier / TMA / tcgen05 inline-PTX wrappers and the UMMA descriptor encoders.
//
// Nothing here comes from the reference (it has no native code); the PTX forms
// follow the PTX ISA for sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

namespace cvflow {

// ------------------------------------------------------------------------------------------
// 16-bit operand type is a run-time flag (0 = fp16, 1 = bf16): tcgen05 kind::f16 runs both at
// the same rate, so one binary serves both and the parity tests can pick per run.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float h16_to_f32(uint16_t v, int bf) {
  if (bf) return __uint_as_float(((uint32_t)v) << 16);
  return __half2float(__ushort_as_half(v));
}
__device__ __forceinline__ uint16_t f32_to_h16(float f, int bf) {
  if (bf) return __bfloat16_as_ushort(__float2bfloat16_rn(f));
  return __half_as_ushort(__float2half_rn(f));
}
__device__ __forceinline__ uint32_t pack2_h16(float a, float b, int bf) {
  if (bf) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);   // one cvt.rn.bf16x2.f32
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 8 floats -> 8 packed 16-bit values with a single (kernel-uniform) dtype branch
__device__ __forceinline__ uint4 pack8_h16(const float* v, int bf) {
  uint4 w;
  if (bf) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    w.x = *reinterpret_cast<uint32_t*>(&a); w.y = *reinterpret_cast<uint32_t*>(&b);
    w.z = *reinterpret_cast<uint32_t*>(&c); w.w = *reinterpret_cast<uint32_t*>(&d);
  } else {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    w.x = *reinterpret_cast<uint32_t*>(&a); w.y = *reinterpret_cast<uint32_t*>(&b);
    w.z = *reinterpret_cast<uint32_t*>(&c); w.w = *reinterpret_cast<uint32_t*>(&d);
  }
  return w;
}
__device__ __forceinline__ void unpack2_h16(uint32_t v, int bf, float& a, float& b) {
  a = h16_to_f32((uint16_t)(v & 0xffffu), bf);
  b = h16_to_f32((uint16_t)(v >> 16), bf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Mish(x) = x * tanh(softplus(x)); with e = exp(x): tanh(log(1+e)) = ((1+e)^2-1)/((1+e)^2+1).
__device__ __forceinline__ float mish_f(float x) {
  if (x > 20.f) return x;
  float e = __expf(x);
  float n = e * (e + 2.f);
  return x * n / (n + 2.f);
}
// d/dx Mish(x) = tsp + x * sigmoid(x) * (1 - tsp^2), tsp = tanh(softplus(x)).
__device__ __forceinline__ float mish_grad_f(float x) {
  if (x > 20.f) return 1.f;
  float e = __expf(x);
  float n = e * (e + 2.f);
  float tsp = n / (n + 2.f);
  float sig = e / (1.f + e);
  return tsp + x * sig * (1.f - tsp * tsp);
}
// 2^x on the SFU (MUFU.EX2, flush-to-zero): exact enough for probabilities that are rounded to 16 bits
__device__ __forceinline__ float exp2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// hardware tanh (MUFU.TANH, rel. error ~2^-11: below the 16-bit rounding of the tensors it feeds)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU tanh approximation (reference modules.py:132,139 uses approximate="tanh").
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.f + tanh_fast(u));
}
__device__ __forceinline__ float gelu_tanh_grad_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float u = k0 * (x + k1 * x * x2);
  float th = tanh_fast(u);
  float du = k0 * (1.f + 3.f * k1 * x2);
  return 0.5f * (1.f + th) + 0.5f * x * (1.f - th * th) * du;
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.f + erff(x * 0.7071067811865475f));
}
__device__ __forceinline__ float gelu_erf_grad_f(float x) {
  return 0.5f * (1.f + erff(x * 0.7071067811865475f)) +
         x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// ------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 issue one instruction for two lanes): the
// epilogues that run an activation over a whole hidden tile on one SM are issue-bound otherwise.
// ------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// GELU(tanh) of a pair; 5 packed ops + 2 MUFU.TANH
__device__ __forceinline__ f32x2 gelu_tanh_f2(f32x2 x) {
  const f32x2 c0 = f2_pack(0.7978845608028654f, 0.7978845608028654f);
  const f32x2 c1 = f2_pack(0.7978845608028654f * 0.044715f, 0.7978845608028654f * 0.044715f);
  const f32x2 half = f2_pack(0.5f, 0.5f);
  const f32x2 u = f2_mul(x, f2_fma(f2_mul(x, x), c1, c0));
  float u0, u1;
  f2_unpack(u, u0, u1);
  const f32x2 th = f2_pack(tanh_fast(u0), tanh_fast(u1));
  const f32x2 hx = f2_mul(x, half);
  return f2_fma(hx, th, hx);
}
// d * GELU'(x) of a pair (tanh form); 10 packed ops + 2 MUFU.TANH
__device__ __forceinline__ f32x2 gelu_tanh_grad_mul_f2(f32x2 d, f32x2 x) {
  const f32x2 c0 = f2_pack(0.7978845608028654f, 0.7978845608028654f);
  const f32x2 c1 = f2_pack(0.7978845608028654f * 0.044715f, 0.7978845608028654f * 0.044715f);
  const f32x2 c3 = f2_pack(3.f * 0.7978845608028654f * 0.044715f, 3.f * 0.7978845608028654f * 0.044715f);
  const f32x2 half = f2_pack(0.5f, 0.5f), one = f2_pack(1.f, 1.f), mone = f2_pack(-1.f, -1.f);
  const f32x2 x2 = f2_mul(x, x);
  const f32x2 u = f2_mul(x, f2_fma(x2, c1, c0));
  float u0, u1;
  f2_unpack(u, u0, u1);
  const f32x2 th = f2_pack(tanh_fast(u0), tanh_fast(u1));
  const f32x2 du = f2_fma(x2, c3, c0);
  const f32x2 a = f2_fma(th, half, half);                  // 0.5 (1 + th)
  const f32x2 s = f2_fma(f2_mul(th, mone), th, one);       // 1 - th^2
  const f32x2 m = f2_mul(f2_mul(x, half), s);
  return f2_mul(d, f2_fma(m, du, a));
}

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with launch_pdl() may begin (barrier init,
// TMEM allocation, descriptor prefetch) while its predecessor drains; it must execute pdl_wait()
// before its first global-memory access. pdl_launch() lets the successor start early.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint
// expires) instead of returning at once, so a waiting warp does not compete for issue slots with the warps
// doing the math on its scheduler (measured equal to a plain poll loop in the attention kernel; kept).
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok;
}
// Bounded wait: a wrong descriptor / byte count must fault loudly rather than hang the box (a hung GPU box is a
// strike). The bound is WALL TIME on %globaltimer (5 s), not a try count: how long one try_wait parks is up to the
// hardware (the hint is only an upper limit), and under a profiler's instrumented replay passes a kernel runs two orders
// of magnitude slower -- a count-based bound fired there (ncu --set full on the dK/dV kernel at L = 400).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  unsigned long long t0 = 0ull;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0ull) t0 = now;
      else if (now - t0 > 5000000000ull) {
        printf("cvflow: mbarrier timeout block (%d,%d) thread %d bar %u parity %u\n", blockIdx.x, blockIdx.y, threadIdx.x, bar,
               parity);
        __trap();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, completion on an mbarrier.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Multicast TMA load: the box lands at the same shared-memory offset in every CTA of the cluster named by
// cta_mask and completes tx bytes on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
      "{%3, %4, %5}], [%2], %6;" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// distributed shared memory: address of `local` in the CTA `rank` of this cluster, and a 32-bit store to it
__device__ __forceinline__ uint32_t dsmem_map(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsmem_st_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// TMA store (shared -> global), bulk-group completion
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// Same, arriving on the mbarrier at this offset in every CTA of cta_mask (stage release of a multicast pipeline).
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 | [49,52) base offset | [61,64) swizzle (2 = 128B).
// K-major operand, 128B swizzle, one 64-element (128-byte) K atom per row, rows packed densely:
// 8-row core groups are 1024 bytes apart (SBO); LBO is ignored for swizzled K-major (set to 1).
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand, 128B swizzle: the tile is stored as [k rows][64 contiguous MN elements]
// (128 bytes per k row), i.e. exactly what a TMA box {64, rows} with SWIZZLE_128B deposits for a
// row-major [k][mn] global tensor. 8 k-rows form a 1024-byte atom (SBO = 1024); successive
// 64-element MN chunks are `lbo_bytes` apart.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr,
                                                            uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, A/B format f16 (0) or bf16 (1).
__device__ __forceinline__ uint32_t umma_idesc_f16(int bf16, int M, int N, int a_mn_major,
                                                   int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                           // D format = F32
  d |= (uint32_t)(bf16 ? 1 : 0) << 7;     // A format
  d |= (uint32_t)(bf16 ? 1 : 0) << 10;    // B format
  d |= (uint32_t)(a_mn_major ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn_major ? 1 : 0) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

}  // namespace cvflow

#include <cstdlib>
#include <utility>
namespace cvflow {
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CVFLOW_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// Launch with programmatic stream serialization (PDL). The kernel MUST call pdl_wait() before
// touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
}  // namespace cvflow
