// LayerNorm(256) and GroupNorm(8 x 32 channels) + Mish passes, forward and backward.
// Token-major [rows][256] tensors; one lane owns 8 contiguous channels (16-byte accesses), so a
// warp covers a full row and a GroupNorm group is 4 adjacent lanes (warp-shuffle reductions).
//
// Reference semantics: nn.LayerNorm eps 1e-5 (modules.py:318,346); Block1D = Conv1d -> GroupNorm(8)
// -> Mish with the statistics taken over ALL padded positions (modules.py:60-73, SURVEY trap 1);
// resnet time bias added after the first block's mask (modules.py:90-92).
#include "kernels.h"
#include "common.cuh"
#include "rowops.cuh"

namespace cvflow {

#define LAUNCH_RET() do { cudaError_t e_ = cudaGetLastError(); return e_ == cudaSuccess ? 0 : -(int)e_; } while (0)

// ------------------------------------------------------------------------------------------
// LayerNorm
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) layernorm_fwd_kernel(const float* __restrict__ h, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, uint16_t* __restrict__ out,
                                                            long M, int bf) {
  pdl_wait();
  pdl_launch();
  const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float x[8], g[8], b[8];
  load8_f32(h + row * 256 + lane * 8, x);
  load8_f32(gamma + lane * 8, g);
  load8_f32(beta + lane * 8, b);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] -= mean; v += x[i] * x[i]; }
  const float rstd = rsqrtf(warp_sum(v) * (1.f / 256.f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = x[i] * rstd * g[i] + b[i];
  store8_h16(out + row * 256 + lane * 8, bf, x);
}
int launch_layernorm_fwd(const float* h, const float* gamma, const float* beta, void* out, long M, int bf16,
                         cudaStream_t st) {
  launch_pdl(layernorm_fwd_kernel, (unsigned)((M + 3) / 4), 128, 0, st, h, gamma, beta, reinterpret_cast<uint16_t*>(out), M,
                                                               bf16);
  LAUNCH_RET();
}

__global__ void __launch_bounds__(128) layernorm_bwd_kernel(const uint16_t* __restrict__ dxn, long ld_dxn, const float* __restrict__ h_in,
                                                            const float* __restrict__ gamma, const float* __restrict__ dres,
                                                            float* __restrict__ dh, uint16_t* __restrict__ dh16, long M,
                                                            int bf) {
  pdl_wait();
  pdl_launch();
  const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float x[8], g[8], d[8];
  load8_f32(h_in + row * 256 + lane * 8, x);
  load8_f32(gamma + lane * 8, g);
  load8_h16(dxn + row * ld_dxn + lane * 8, bf, d);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] -= mean; v += x[i] * x[i]; }
  const float rstd = rsqrtf(warp_sum(v) * (1.f / 256.f) + 1e-5f);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] *= rstd;       // xhat
    d[i] *= g[i];       // dxhat
    s1 += d[i];
    s2 += d[i] * x[i];
  }
  s1 = warp_sum(s1) * (1.f / 256.f);
  s2 = warp_sum(s2) * (1.f / 256.f);
  float r[8];
  if (dres) load8_f32(dres + row * 256 + lane * 8, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = rstd * (d[i] - s1 - x[i] * s2) + (dres ? r[i] : 0.f);
  if (dh) store8_f32(dh + row * 256 + lane * 8, d);
  if (dh16) store8_h16(dh16 + row * 256 + lane * 8, bf, d);
}
int launch_layernorm_bwd(const void* dxn16, long ld_dxn, const float* h_in, const float* gamma, const float* dres, float* dh,
                         void* dh16, long M, int bf16, cudaStream_t st) {
  launch_pdl(layernorm_bwd_kernel, (unsigned)((M + 3) / 4), 128, 0, st, reinterpret_cast<const uint16_t*>(dxn16), ld_dxn, h_in, gamma,
                                                               dres, dh, reinterpret_cast<uint16_t*>(dh16), M, bf16);
  LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// GroupNorm(8 groups of 32 channels) + Mish. The statistics come as Chan-mergeable partials {n, mean, M2} per
// (sample, 32-row slice, group) written by the epilogue of the conv GEMM that produced the tensor (gemm.cu, gn_part):
// no separate statistics pass re-reads the conv output. Every CTA of the apply pass works on rows of ONE sample and
// merges that sample's partials once, in shared memory.
// ------------------------------------------------------------------------------------------
int gn_num_splits(int B, int L) {   // backward partial sums: (sample, row split) CTAs
  int s = (2 * 148 + B - 1) / B;
  const int max_s = (L + 15) / 16;
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}
int gn_fwd_splits(int L) { return 4 * ((L + 127) / 128); }   // partials per sample written by the GEMM epilogue

// Mish(x) = x tanh(softplus(x)) = x n / (n + 2), n = e (e + 2), e = exp(x): one MUFU.EX2 + one MUFU.RCP
__device__ __forceinline__ float mish_fast(float x) {
  const float e = __expf(fminf(x, 20.f));
  const float n = e * (e + 2.f);
  return x * __fdividef(n, n + 2.f);
}
// d/dx Mish(x) = tsp + x sigmoid(x) (1 - tsp^2), tsp = n / (n + 2), sigmoid = e / (1 + e): one reciprocal serves both
__device__ __forceinline__ float mish_grad_fast(float x) {
  const float e = __expf(fminf(x, 20.f));
  const float n = e * (e + 2.f);
  const float a = n + 2.f, b = 1.f + e;
  const float r = __fdividef(1.f, a * b);
  const float tsp = n * b * r;
  const float sig = e * a * r;
  return tsp + x * sig * (1.f - tsp * tsp);
}

// merge the partials of sample b into sm[g] = {mean, rstd}: the {n, mean, M2} triples are fetched by all threads in
// parallel (one L2 round trip instead of a dependent chain of nsplit of them), then 8 threads run the Chan merge from
// shared memory. Call with all 256 threads; ends with a __syncthreads.
static constexpr int kGnMaxSplits = 128;   // 4 * ceil(L / 128): L <= 4096
__device__ __forceinline__ void gn_merge_to_smem(const float* __restrict__ partials, int b, int nsplit, float (*part)[3],
                                                 float (*sm)[2]) {
  for (int i = threadIdx.x; i < nsplit * 8; i += blockDim.x) {
    const float* p = partials + ((long)b * nsplit * 8 + i) * 3;
    part[i][0] = p[0]; part[i][1] = p[1]; part[i][2] = p[2];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int g = threadIdx.x;
    float n = 0.f, mu = 0.f, m2 = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float nb = part[s * 8 + g][0];
      if (nb <= 0.f) continue;
      const float delta = part[s * 8 + g][1] - mu;
      const float nn = n + nb;
      mu += delta * nb / nn;
      m2 += part[s * 8 + g][2] + delta * delta * n * nb / nn;
      n = nn;
    }
    sm[g][0] = mu;
    sm[g][1] = rsqrtf(m2 / n + 1e-5f);
  }
  __syncthreads();
}

static constexpr int kGnRows = 16;   // rows of one sample per CTA (8 warps x 2 rows, both rows' loads in flight together)

__global__ void __launch_bounds__(256) gn_apply_kernel(const uint16_t* __restrict__ c, const float* __restrict__ partials,
                                                       int nsplit, float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ mask, const float* __restrict__ tb, long tb_stride,
                                                       const uint16_t* __restrict__ add16, void* __restrict__ out, int mode,
                                                       int L, int bf) {
  __shared__ float sm[8][2];
  __shared__ float part[kGnMaxSplits * 8][3];
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.y;
  gn_merge_to_smem(partials, b, nsplit, part, sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c8 = lane * 8, grp = lane >> 2;
  const float mean = sm[grp][0], rstd = sm[grp][1];
  if (blockIdx.x == 0 && threadIdx.x < 8) {   // keep (mean, rstd) for the backward pass
    stats[(b * 8 + threadIdx.x) * 2] = sm[threadIdx.x][0];
    stats[(b * 8 + threadIdx.x) * 2 + 1] = sm[threadIdx.x][1];
  }
  float sc[8], sh[8], t8[8];
  load8_f32(gamma + c8, sc);
  load8_f32(beta + c8, sh);
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] *= rstd; sh[k] -= mean * sc[k]; }
  if (mode == 0 && tb) load8_f32(tb + (long)b * tb_stride + c8, t8);
  const int l0 = blockIdx.x * kGnRows;
  constexpr int NR = kGnRows / 8;
  uint4 cv[NR], av[NR];
  float m[NR];
  long row[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {        // every load of the CTA's rows is issued before any of the math
    const int l = l0 + r * 8 + warp;
    row[r] = l < L ? (long)b * L + l : -1;
    if (row[r] >= 0) {
      cv[r] = *reinterpret_cast<const uint4*>(c + row[r] * 256 + c8);
      m[r] = mask[row[r]];
      if (mode != 0) av[r] = *reinterpret_cast<const uint4*>(add16 + row[r] * 256 + c8);
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (row[r] < 0) continue;
    float x[8];
    unpack8_h16(cv[r], bf, x);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = mish_fast(x[k] * sc[k] + sh[k]);
    if (mode == 0) {
      if (tb) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] += t8[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] *= m[r];
      store8_h16(reinterpret_cast<uint16_t*>(out) + row[r] * 256 + c8, bf, x);
    } else {
      float a[8];
      unpack8_h16(av[r], bf, a);
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = x[k] * m[r] + a[k];
      store8_f32(reinterpret_cast<float*>(out) + row[r] * 256 + c8, x);
    }
  }
}
int launch_gn_apply(const void* c16, const float* partials, float* stats, const float* gamma, const float* beta,
                    const float* mask, const float* tb, long tb_stride, const void* add16, void* out, int mode, int B, int L,
                    int bf16, cudaStream_t st) {
  launch_pdl(gn_apply_kernel, dim3((L + kGnRows - 1) / kGnRows, B), 256, 0, st, reinterpret_cast<const uint16_t*>(c16),
             partials, gn_fwd_splits(L), stats, gamma, beta, mask, tb, tb_stride, reinterpret_cast<const uint16_t*>(add16), out,
             mode, L, bf16);
  LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// GroupNorm + Mish backward: dz = dy*mask*mish'(z); dxhat = dz*gamma;
// dc = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat*xhat))
// ------------------------------------------------------------------------------------------
// per-lane constants of a sample: xhat = c*a1 + a0, z = c*sc + sh
struct GnLane { float a1[8], a0[8], sc[8], sh[8], g[8]; };
__device__ __forceinline__ void gn_lane_consts(const float* __restrict__ stats, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, int b, int lane, GnLane& k) {
  const int grp = lane >> 2, c8 = lane * 8;
  const float mean = stats[(b * 8 + grp) * 2], rstd = stats[(b * 8 + grp) * 2 + 1];
  float be[8];
  load8_f32(gamma + c8, k.g);
  load8_f32(beta + c8, be);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    k.a1[i] = rstd; k.a0[i] = -mean * rstd;
    k.sc[i] = rstd * k.g[i]; k.sh[i] = be[i] - mean * k.sc[i];
  }
}
__device__ __forceinline__ void gn_bwd_load(const void* dy, int dy_f32, const uint16_t* c, const GnLane& k, float m, long off,
                                            int bf, float (&xh)[8], float (&dxh)[8]) {
  float x[8], d[8];
  load8_h16(c + off, bf, x);
  if (dy_f32) load8_f32(reinterpret_cast<const float*>(dy) + off, d);
  else load8_h16(reinterpret_cast<const uint16_t*>(dy) + off, bf, d);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xh[i] = x[i] * k.a1[i] + k.a0[i];
    dxh[i] = d[i] * m * mish_grad_fast(x[i] * k.sc[i] + k.sh[i]) * k.g[i];
  }
}

__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const void* __restrict__ dy, int dy_f32,
                                                            const uint16_t* __restrict__ c, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mask, float* __restrict__ partials,
                                                            int L, int nsplit, int bf) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[8][8][2];
  const int b = blockIdx.y, sp = blockIdx.x;
  const int rows_per = (L + nsplit - 1) / nsplit;
  const int l0 = sp * rows_per, l1 = min(L, l0 + rows_per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 2, c8 = lane * 8;
  GnLane k;
  gn_lane_consts(stats, gamma, beta, b, lane, k);
  float s1 = 0.f, s2 = 0.f;
  for (int l = l0 + warp; l < l1; l += 8) {
    const long row = (long)b * L + l;
    float xh[8], dxh[8];
    gn_bwd_load(dy, dy_f32, c, k, mask[row], row * 256 + c8, bf, xh, dxh);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1 += dxh[i]; s2 += dxh[i] * xh[i]; }
  }
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  if ((lane & 3) == 0) { red[warp][grp][0] = s1; red[warp][grp][1] = s2; }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int g = threadIdx.x;
    float a = 0.f, q = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[w][g][0]; q += red[w][g][1]; }
    float* o = partials + (((long)b * nsplit + sp) * 8 + g) * 2;
    o[0] = a; o[1] = q;
  }
}
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const void* __restrict__ dy, int dy_f32,
                                                           const uint16_t* __restrict__ c, const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mask, const float* __restrict__ partials,
                                                           int nsplit, float inv_n, uint16_t* __restrict__ dc, int L, int bf) {
  __shared__ float sm[8][2];
  pdl_wait();
  pdl_launch();
  __shared__ float part[64 * 8][2];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < nsplit * 8; i += blockDim.x) {   // all split partials in one L2 round trip
    const float* pp = partials + ((long)b * nsplit * 8 + i) * 2;
    part[i][0] = pp[0]; part[i][1] = pp[1];
  }
  __syncthreads();
  if (threadIdx.x < 8) {      // group means of dxhat and dxhat * xhat, fixed summation order
    float m1 = 0.f, m2 = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) { m1 += part[sp * 8 + threadIdx.x][0]; m2 += part[sp * 8 + threadIdx.x][1]; }
    sm[threadIdx.x][0] = m1 * inv_n; sm[threadIdx.x][1] = m2 * inv_n;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 2, c8 = lane * 8;
  const float m1 = sm[grp][0], m2 = sm[grp][1];
  GnLane k;
  gn_lane_consts(stats, gamma, beta, b, lane, k);
  const int l0 = blockIdx.x * kGnRows;
#pragma unroll
  for (int r = 0; r < kGnRows / 8; ++r) {
    const int l = l0 + r * 8 + warp;
    if (l < L) {
      const long row = (long)b * L + l;
      float xh[8], dxh[8];
      gn_bwd_load(dy, dy_f32, c, k, mask[row], row * 256 + c8, bf, xh, dxh);
#pragma unroll
      for (int i = 0; i < 8; ++i) dxh[i] = k.a1[i] * (dxh[i] - m1 - xh[i] * m2);
      store8_h16(dc + row * 256 + c8, bf, dxh);
    }
  }
}
int launch_gn_bwd(const void* dy, int dy_f32, const void* c16, const float* stats, const float* gamma,
                  const float* beta, const float* mask, float* partials, void* dc16, int B, int L, int bf16,
                  cudaStream_t st) {
  const int ns = gn_num_splits(B, L);
  const uint16_t* c = reinterpret_cast<const uint16_t*>(c16);
  launch_pdl(gn_bwd_reduce_kernel, dim3(ns, B), 256, 0, st, dy, dy_f32, c, stats, gamma, beta, mask, partials, L, ns, bf16);
  launch_pdl(gn_bwd_apply_kernel, dim3((L + kGnRows - 1) / kGnRows, B), 256, 0, st, dy, dy_f32, c, stats, gamma, beta, mask,
             (const float*)partials, ns, 1.f / (32.f * (float)L), reinterpret_cast<uint16_t*>(dc16), L, bf16);
  LAUNCH_RET();
}

}  // namespace cvflow
