// LayerNorm(256) and GroupNorm(8 x 32 channels) + Mish passes, forward and backward.
// Token-major [rows][256] tensors; one lane owns 8 contiguous channels (16-byte accesses), so a
// warp covers a full row and a GroupNorm group is 4 adjacent lanes (warp-shuffle reductions).
//
// Reference semantics: nn.LayerNorm eps 1e-5 (modules.py:318,346); Block1D = Conv1d -> GroupNorm(8)
// -> Mish with the statistics taken over ALL padded positions (modules.py:60-73, SURVEY trap 1);
// resnet time bias added after the first block's mask (modules.py:90-92).
#include "kernels.h"
#include "common.cuh"

namespace cvflow {

#define LAUNCH_RET() do { cudaError_t e_ = cudaGetLastError(); return e_ == cudaSuccess ? 0 : -(int)e_; } while (0)

__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8_f32(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load8_h16(const uint16_t* p, int bf, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  unpack2_h16(u.x, bf, v[0], v[1]);
  unpack2_h16(u.y, bf, v[2], v[3]);
  unpack2_h16(u.z, bf, v[4], v[5]);
  unpack2_h16(u.w, bf, v[6], v[7]);
}
__device__ __forceinline__ void store8_h16(uint16_t* p, int bf, const float (&v)[8]) {
  uint4 u;
  u.x = pack2_h16(v[0], v[1], bf); u.y = pack2_h16(v[2], v[3], bf);
  u.z = pack2_h16(v[4], v[5], bf); u.w = pack2_h16(v[6], v[7], bf);
  *reinterpret_cast<uint4*>(p) = u;
}

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
// GroupNorm statistics: per (b, split) shifted sums -> (n, mean, M2); Chan merge over splits.
// ------------------------------------------------------------------------------------------
int gn_num_splits(int B, int L) {
  int s = (2 * 148 + B - 1) / B;
  const int max_s = (L + 15) / 16;
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

__global__ void __launch_bounds__(256) gn_stats_kernel(const uint16_t* __restrict__ c, float* __restrict__ partials,
                                                       int L, int nsplit, int bf) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[8][8][2];
  const int b = blockIdx.y, sp = blockIdx.x;
  const int rows_per = (L + nsplit - 1) / nsplit;
  const int l0 = sp * rows_per, l1 = min(L, l0 + rows_per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 2;
  float s1 = 0.f, s2 = 0.f, K = 0.f;
  if (l0 < l1) K = h16_to_f32(c[((long)b * L + l0) * 256 + grp * 32], bf);
  for (int l = l0 + warp; l < l1; l += 8) {
    float x[8];
    load8_h16(c + ((long)b * L + l) * 256 + lane * 8, bf, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = x[i] - K; s1 += d; s2 += d * d; }
  }
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  if ((lane & 3) == 0) { red[warp][grp][0] = s1; red[warp][grp][1] = s2; }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int g = threadIdx.x;
    float a = 0.f, q = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[w][g][0]; q += red[w][g][1]; }
    const float n = 32.f * (float)max(0, l1 - l0);
    const float Kg = (l0 < l1) ? h16_to_f32(c[((long)b * L + l0) * 256 + g * 32], bf) : 0.f;
    float* o = partials + (((long)b * nsplit + sp) * 8 + g) * 3;
    o[0] = n;
    o[1] = n > 0.f ? Kg + a / n : 0.f;
    o[2] = n > 0.f ? q - a * a / n : 0.f;
  }
}
// Chan merge of the split partials of one (sample, group): every consumer thread does it for its own group
// (a few dozen cached loads) instead of a separate single-CTA launch between the reduction and the apply pass.
__device__ __forceinline__ void gn_merge_partials(const float* __restrict__ partials, int b, int g, int nsplit, float& mean,
                                                  float& rstd) {
  float n = 0.f, mu = 0.f, m2 = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float* p = partials + (((long)b * nsplit + s) * 8 + g) * 3;
    const float nb = p[0];
    if (nb <= 0.f) continue;
    const float delta = p[1] - mu;
    const float nn = n + nb;
    mu += delta * nb / nn;
    m2 += p[2] + delta * delta * n * nb / nn;
    n = nn;
  }
  mean = mu;
  rstd = rsqrtf(m2 / n + 1e-5f);
}
int launch_gn_stats(const void* c16, float* partials, float* stats, int B, int L, int bf16, cudaStream_t st) {
  const int ns = gn_num_splits(B, L);
  (void)stats;   // written by the apply pass (first row of every sample)
  launch_pdl(gn_stats_kernel, dim3(ns, B), 256, 0, st, reinterpret_cast<const uint16_t*>(c16), partials, L, ns, bf16);
  LAUNCH_RET();
}

__global__ void __launch_bounds__(128) gn_apply_kernel(const uint16_t* __restrict__ c, const float* __restrict__ partials,
                                                       int nsplit, float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ mask, const float* __restrict__ tb, long tb_stride,
                                                       const uint16_t* __restrict__ add16, void* __restrict__ out, int mode,
                                                       int L, long M, int bf) {
  pdl_wait();
  pdl_launch();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * 32) return;
  const long row = i >> 5;
  const int lane = (int)(i & 31);
  const int b = (int)(row / L);
  const int c8 = lane * 8, grp = lane >> 2;
  float mean, rstd;
  gn_merge_partials(partials, b, grp, nsplit, mean, rstd);
  if (row == (long)b * L && (lane & 3) == 0) {   // keep (mean, rstd) for the backward pass
    stats[(b * 8 + grp) * 2] = mean;
    stats[(b * 8 + grp) * 2 + 1] = rstd;
  }
  const float m = mask[row];
  float x[8], g[8], be[8];
  load8_h16(c + row * 256 + c8, bf, x);
  load8_f32(gamma + c8, g);
  load8_f32(beta + c8, be);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = mish_f((x[k] - mean) * rstd * g[k] + be[k]);
  if (mode == 0) {
    if (tb) {
      float t8[8];
      load8_f32(tb + (long)b * tb_stride + c8, t8);
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] += t8[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] *= m;
    store8_h16(reinterpret_cast<uint16_t*>(out) + row * 256 + c8, bf, x);
  } else {
    float a[8];
    load8_h16(add16 + row * 256 + c8, bf, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = x[k] * m + a[k];
    store8_f32(reinterpret_cast<float*>(out) + row * 256 + c8, x);
  }
}
int launch_gn_apply(const void* c16, const float* partials, float* stats, const float* gamma, const float* beta,
                    const float* mask, const float* tb, long tb_stride, const void* add16, void* out, int mode, int B, int L,
                    int bf16, cudaStream_t st) {
  const long M = (long)B * L, n = M * 32;
  launch_pdl(gn_apply_kernel, (unsigned)((n + 127) / 128), 128, 0, st, reinterpret_cast<const uint16_t*>(c16), partials,
                                                               gn_num_splits(B, L), stats, gamma,
                                                               beta, mask, tb, tb_stride,
                                                               reinterpret_cast<const uint16_t*>(add16), out, mode, L,
                                                               M, bf16);
  LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// GroupNorm + Mish backward: dz = dy*mask*mish'(z); dxhat = dz*gamma;
// dc = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat*xhat))
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void gn_bwd_load(const void* dy, int dy_f32, const uint16_t* c, const float* gamma,
                                            const float* beta, float mean, float rstd, float m, long off, int c8,
                                            int bf, float (&xh)[8], float (&dxh)[8]) {
  float x[8], g[8], be[8], d[8];
  load8_h16(c + off, bf, x);
  load8_f32(gamma + c8, g);
  load8_f32(beta + c8, be);
  if (dy_f32) load8_f32(reinterpret_cast<const float*>(dy) + off, d);
  else load8_h16(reinterpret_cast<const uint16_t*>(dy) + off, bf, d);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    xh[k] = (x[k] - mean) * rstd;
    const float zz = xh[k] * g[k] + be[k];
    dxh[k] = d[k] * m * mish_grad_f(zz) * g[k];
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
  const float mean = stats[(b * 8 + grp) * 2], rstd = stats[(b * 8 + grp) * 2 + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int l = l0 + warp; l < l1; l += 8) {
    const long row = (long)b * L + l;
    float xh[8], dxh[8];
    gn_bwd_load(dy, dy_f32, c, gamma, beta, mean, rstd, mask[row], row * 256 + c8, c8, bf, xh, dxh);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1 += dxh[k]; s2 += dxh[k] * xh[k]; }
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
__global__ void __launch_bounds__(128) gn_bwd_apply_kernel(const void* __restrict__ dy, int dy_f32,
                                                           const uint16_t* __restrict__ c, const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mask, const float* __restrict__ partials,
                                                           int nsplit, float inv_n, uint16_t* __restrict__ dc, int L, long M,
                                                           int bf) {
  pdl_wait();
  pdl_launch();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * 32) return;
  const long row = i >> 5;
  const int lane = (int)(i & 31);
  const int b = (int)(row / L);
  const int c8 = lane * 8, grp = lane >> 2;
  const float mean = stats[(b * 8 + grp) * 2], rstd = stats[(b * 8 + grp) * 2 + 1];
  float m1 = 0.f, m2 = 0.f;   // group means of dxhat and dxhat * xhat from the split partials
  for (int sp = 0; sp < nsplit; ++sp) {
    const float* pp = partials + (((long)b * nsplit + sp) * 8 + grp) * 2;
    m1 += pp[0]; m2 += pp[1];
  }
  m1 *= inv_n; m2 *= inv_n;
  float xh[8], dxh[8];
  gn_bwd_load(dy, dy_f32, c, gamma, beta, mean, rstd, mask[row], row * 256 + c8, c8, bf, xh, dxh);
#pragma unroll
  for (int k = 0; k < 8; ++k) dxh[k] = rstd * (dxh[k] - m1 - xh[k] * m2);
  store8_h16(dc + row * 256 + c8, bf, dxh);
}
int launch_gn_bwd(const void* dy, int dy_f32, const void* c16, const float* stats, const float* gamma,
                  const float* beta, const float* mask, float* partials, void* dc16, int B, int L, int bf16,
                  cudaStream_t st) {
  const int ns = gn_num_splits(B, L);
  const uint16_t* c = reinterpret_cast<const uint16_t*>(c16);
  launch_pdl(gn_bwd_reduce_kernel, dim3(ns, B), 256, 0, st, dy, dy_f32, c, stats, gamma, beta, mask, partials, L, ns, bf16);
  const long M = (long)B * L, n = M * 32;
  launch_pdl(gn_bwd_apply_kernel, (unsigned)((n + 127) / 128), 128, 0, st, dy, dy_f32, c, stats, gamma, beta, mask,
                                                                   (const float*)partials, ns, 1.f / (32.f * (float)L),
                                                                   reinterpret_cast<uint16_t*>(dc16), L, M, bf16);
  LAUNCH_RET();
}

}  // namespace cvflow
