// HBM-bound passes of the flow path: input packing, timestep embedding, residual-stream casts,
// CFM interpolation / masked loss, CFG + Euler update. Vectorised 16-byte accesses, coalesced
// along the contiguous dimension, warp-shuffle reductions.
#include "kernels.h"
#include "common.cuh"

namespace cvflow {

#define LAUNCH_RET() do { cudaError_t e_ = cudaGetLastError(); return e_ == cudaSuccess ? 0 : -(int)e_; } while (0)

// ------------------------------------------------------------------------------------------
// pack: channel-major fp32 sources -> token-major 16-bit [B][T][320], masked.
// One block = one (b, 32-frame tile); smem transpose so both sides are coalesced.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_inputs_kernel(
    const float* __restrict__ x, int x_nb, const float* __restrict__ mu, int mu_nb,
    const float* __restrict__ spks, int spks_nb, const float* __restrict__ cond, int cond_nb,
    const float* __restrict__ mask, int mask_nb, const float* __restrict__ keep, uint32_t* __restrict__ out,
    int T, int bf) {
  __shared__ float tile[320][33];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int t = t0 + tx;
  const float kp = keep ? keep[b] : 1.f;
  const float* xs = x + (long)(b % x_nb) * 80 * T;
  const float* ms = mu + (long)(b % mu_nb) * 80 * T;
  const float* cs = cond ? cond + (long)(b % cond_nb) * 80 * T : nullptr;
  const float* ss = spks ? spks + (long)(b % spks_nb) * 80 : nullptr;
  const float m = (t < T) ? mask[(long)(b % mask_nb) * T + t] : 0.f;
  for (int c = wy; c < 80; c += 8) {
    const bool ok = t < T;
    tile[c][tx] = ok ? xs[(long)c * T + t] * m : 0.f;
    tile[80 + c][tx] = ok ? ms[(long)c * T + t] * kp * m : 0.f;
    tile[160 + c][tx] = (ok && ss) ? ss[c] * kp * m : 0.f;
    tile[240 + c][tx] = (ok && cs) ? cs[(long)c * T + t] * kp * m : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * 160; idx += 256) {
    const int r = idx / 160, cp = idx - r * 160;
    if (t0 + r < T)
      out[((long)b * T + t0 + r) * 160 + cp] = pack2_h16(tile[2 * cp][r], tile[2 * cp + 1][r], bf);
  }
}

int launch_pack_inputs(const float* x, int x_nb, const float* mu, int mu_nb, const float* spks, int spks_nb,
                       const float* cond, int cond_nb, const float* mask, int mask_nb, const float* keep,
                       void* out, int B, int T, int bf16, cudaStream_t st) {
  dim3 grid((T + 31) / 32, B);
  pack_inputs_kernel<<<grid, 256, 0, st>>>(x, x_nb, mu, mu_nb, spks, spks_nb, cond, cond_nb, mask, mask_nb, keep,
                                           reinterpret_cast<uint32_t*>(out), T, bf16);
  LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// unpack (backward of pack): 16-bit token-major dL/d(packed input) [B][T][320] -> fp32 channel-major
// dL/dx, dL/dmu, dL/dcond [B][80][T] and per-tile partial sums of dL/dspks (summed in fixed order by
// spk_grad_reduce_kernel: deterministic, no atomics). The row mask is already applied by the
// producing dgrad GEMM; mu / spks / cond carry the CFG keep factor of the forward pack.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpack_input_grads_kernel(
    const uint32_t* __restrict__ g, const float* __restrict__ keep, float scale, const float* __restrict__ gs_dev,
    float* __restrict__ dx, float* __restrict__ dmu, float* __restrict__ dcond, float* __restrict__ spk_part,
    int T, int bf) {
  __shared__ float tile[320][33];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const float s = scale * (gs_dev ? gs_dev[0] : 1.f);
  const float sk = s * (keep ? keep[b] : 1.f);
  for (int idx = threadIdx.x; idx < 32 * 160; idx += 256) {
    const int r = idx / 160, cp = idx - r * 160;
    float a = 0.f, c = 0.f;
    if (t0 + r < T) unpack2_h16(g[((long)b * T + t0 + r) * 160 + cp], bf, a, c);
    tile[2 * cp][r] = a;
    tile[2 * cp + 1][r] = c;
  }
  __syncthreads();
  const int t = t0 + tx;
  const bool ok = t < T;
  for (int c = wy; c < 80; c += 8) {
    const long o = ((long)b * 80 + c) * T + t;
    if (dx && ok) dx[o] = tile[c][tx] * s;
    if (dmu && ok) dmu[o] = tile[80 + c][tx] * sk;
    if (dcond && ok) dcond[o] = tile[240 + c][tx] * sk;
    if (spk_part) {
      const float v = warp_sum(tile[160 + c][tx]);   // rows beyond T hold zeros
      if (tx == 0) spk_part[((long)b * gridDim.x + blockIdx.x) * 80 + c] = v * sk;
    }
  }
}
__global__ void spk_grad_reduce_kernel(const float* __restrict__ spk_part, int ntiles, float* __restrict__ dspks) {
  const int b = blockIdx.x, c = threadIdx.x;
  if (c >= 80) return;
  float acc = 0.f;
  for (int i = 0; i < ntiles; ++i) acc += spk_part[((long)b * ntiles + i) * 80 + c];
  dspks[(long)b * 80 + c] = acc;
}
int launch_unpack_input_grads(const void* g16, const float* keep, float scale, const float* gs_dev, float* dx,
                              float* dmu, float* dspks, float* dcond, float* spk_part, int B, int T, int bf16,
                              cudaStream_t st) {
  dim3 grid((T + 31) / 32, B);
  unpack_input_grads_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(g16), keep, scale, gs_dev, dx, dmu,
                                                  dcond, dspks ? spk_part : nullptr, T, bf16);
  if (dspks) spk_grad_reduce_kernel<<<B, 96, 0, st>>>(spk_part, (int)grid.x, dspks);
  LAUNCH_RET();
}

__global__ void mask_down_kernel(const float* __restrict__ mask, int mask_nb, float* __restrict__ mask1,
                                 float* __restrict__ mask2, int B, int T, int T2) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long)B * T) {
    const int b = (int)(i / T), t = (int)(i - (long)b * T);
    mask1[i] = mask[(long)(b % mask_nb) * T + t];
  }
  if (i < (long)B * T2) {
    const int b = (int)(i / T2), j = (int)(i - (long)b * T2);
    mask2[i] = mask[(long)(b % mask_nb) * T + 2 * j];
  }
}
int launch_mask_down(const float* mask, int mask_nb, float* mask1, float* mask2, int B, int T, int T2,
                     cudaStream_t st) {
  const long n = (long)B * T;
  mask_down_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mask, mask_nb, mask1, mask2, B, T, T2);
  LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// timestep embedding + small-batch linears (time MLP, per-resnet time projections)
// ------------------------------------------------------------------------------------------
__global__ void sinus_embed_kernel(const float* __restrict__ t, int t_nb, float* __restrict__ out, int B) {
  const int b = blockIdx.x;
  const int i = threadIdx.x;  // 0..159
  if (i >= 160) return;
  // exp(i * -(ln(1e4)/159)) in fp32 like torch.exp(arange.float() * -emb)
  const float step = -(9.210340371976184f / 159.f);
  const float f = expf((float)i * step);
  const float arg = 1000.f * t[b % t_nb] * f;
  out[(long)b * 320 + i] = sinf(arg);
  out[(long)b * 320 + 160 + i] = cosf(arg);
}
int launch_sinus_embed(const float* t, int t_nb, float* out, int B, cudaStream_t st) {
  sinus_embed_kernel<<<B, 160, 0, st>>>(t, t_nb, out, B);
  LAUNCH_RET();
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return v / (1.f + __expf(-v));  // silu
  if (act == 2) return mish_f(v);
  return v;
}

// y[b][n] = out_act(sum_k in_act(x[b][k]) W[n][k] + bias[n]) for small batches (time MLP: B <= 128, K <= 1024, N up to 4096).
// One warp per output feature; the weight row is read exactly once: 256-column k-tiles of up to 32 activated input rows sit
// in shared memory, every lane keeps 8 weights of the tile in registers and accumulates all 32 batch rows (32 accumulators),
// and the cross-lane sums are taken once at the end. HBM-bound on the fp32 weights (22 MB for the three layers).
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           int B, int K, int N, int in_act, int out_act) {
  __shared__ float4 xs[32][64];   // [row][256 columns of the k-tile]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  const int nc = n < N ? n : N - 1;   // clamp: every warp takes part in the barriers
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int nb = min(32, B - b0);
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 256) {
      const int kw = min(256, K - k0);   // multiple of 4 (K % 4 == 0 checked by the launcher)
      __syncthreads();
      for (int i = threadIdx.x; i < nb * 64; i += 256) {
        const int r = i >> 6, c4 = (i & 63) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < kw) {
          v = *reinterpret_cast<const float4*>(x + (long)(b0 + r) * K + k0 + c4);
          v.x = apply_act(v.x, in_act); v.y = apply_act(v.y, in_act); v.z = apply_act(v.z, in_act); v.w = apply_act(v.w, in_act);
        }
        xs[r][i & 63] = v;
      }
      __syncthreads();
      float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
      if (lane * 4 < kw) w0 = __ldg(reinterpret_cast<const float4*>(W + (long)nc * K + k0 + lane * 4));
      if (128 + lane * 4 < kw) w1 = __ldg(reinterpret_cast<const float4*>(W + (long)nc * K + k0 + 128 + lane * 4));
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nb) {
          const float4 a = xs[j][lane], b = xs[j][32 + lane];
          acc[j] += w0.x * a.x + w0.y * a.y + w0.z * a.z + w0.w * a.w + w1.x * b.x + w1.y * b.y + w1.z * b.z + w1.w * b.w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float sum = warp_sum(acc[j]);
      if (lane == 0 && j < nb && n < N) y[(long)(b0 + j) * N + n] = apply_act(sum + (bias ? bias[n] : 0.f), out_act);
    }
  }
}
int launch_small_linear(const float* x, const float* W, const float* bias, float* y, int B, int K, int N,
                        int in_act, int out_act, cudaStream_t st) {
  if (K % 4) return -(int)cudaErrorInvalidValue;
  small_linear_kernel<<<(N + 7) / 8, 256, 0, st>>>(x, W, bias, y, B, K, N, in_act, out_act);
  LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// residual stream (fp32 [M][256]) -> masked 16-bit into a strided destination
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stage_out_kernel(const float* __restrict__ h, const float* __restrict__ rowmask,
                                                        uint16_t* __restrict__ dst, long ldc, int col_off, long M,
                                                        int bf) {
  pdl_wait();
  pdl_launch();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread = 8 channels
  if (i >= M * 32) return;
  const long row = i >> 5;
  const int c8 = (int)(i & 31) * 8;
  const float m = rowmask ? rowmask[row] : 1.f;
  const float4 a = *reinterpret_cast<const float4*>(h + row * 256 + c8);
  const float4 b = *reinterpret_cast<const float4*>(h + row * 256 + c8 + 4);
  uint4 v;
  v.x = pack2_h16(a.x * m, a.y * m, bf);
  v.y = pack2_h16(a.z * m, a.w * m, bf);
  v.z = pack2_h16(b.x * m, b.y * m, bf);
  v.w = pack2_h16(b.z * m, b.w * m, bf);
  *reinterpret_cast<uint4*>(dst + row * ldc + col_off + c8) = v;
}
int launch_stage_out(const float* h, const float* rowmask, void* dst, long ldc, int col_off, long M, int bf16,
                     cudaStream_t st) {
  const long n = M * 32;
  launch_pdl(stage_out_kernel, (unsigned)((n + 127) / 128), 128, 0, st, h, rowmask, reinterpret_cast<uint16_t*>(dst), ldc,
                                                                col_off, M, bf16);
  LAUNCH_RET();
}

__global__ void __launch_bounds__(128) grad_route_kernel(const uint16_t* __restrict__ src, long ld_src, int col_off,
                                                         const float* __restrict__ rowmask, float* __restrict__ dst,
                                                         int accumulate, uint16_t* __restrict__ dst16, long M, int bf) {
  pdl_wait();
  pdl_launch();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * 32) return;
  const long row = i >> 5;
  const int c8 = (int)(i & 31) * 8;
  const float m = rowmask ? rowmask[row] : 1.f;
  const uint4 v = *reinterpret_cast<const uint4*>(src + row * ld_src + col_off + c8);
  float f[8];
  unpack2_h16(v.x, bf, f[0], f[1]);
  unpack2_h16(v.y, bf, f[2], f[3]);
  unpack2_h16(v.z, bf, f[4], f[5]);
  unpack2_h16(v.w, bf, f[6], f[7]);
  float* d = dst + row * 256 + c8;
  float4 a = make_float4(f[0] * m, f[1] * m, f[2] * m, f[3] * m);
  float4 b = make_float4(f[4] * m, f[5] * m, f[6] * m, f[7] * m);
  if (accumulate) {
    const float4 pa = *reinterpret_cast<const float4*>(d);
    const float4 pb = *reinterpret_cast<const float4*>(d + 4);
    a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
    b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
  }
  *reinterpret_cast<float4*>(d) = a;
  *reinterpret_cast<float4*>(d + 4) = b;
  if (dst16) {
    uint4 o;
    o.x = pack2_h16(a.x, a.y, bf); o.y = pack2_h16(a.z, a.w, bf);
    o.z = pack2_h16(b.x, b.y, bf); o.w = pack2_h16(b.z, b.w, bf);
    *reinterpret_cast<uint4*>(dst16 + row * 256 + c8) = o;
  }
}
int launch_grad_route(const void* src, long ld_src, int col_off, const float* rowmask, float* dst, int accumulate,
                      void* dst16, long M, int bf16, cudaStream_t st) {
  const long n = M * 32;
  launch_pdl(grad_route_kernel, (unsigned)((n + 127) / 128), 128, 0, st, reinterpret_cast<const uint16_t*>(src), ld_src,
                                                                 col_off, rowmask, dst, accumulate,
                                                                 reinterpret_cast<uint16_t*>(dst16), M, bf16);
  LAUNCH_RET();
}

// ------------------------------------------------------------------------------------------
// CFM: interpolation, masked loss + dL/dpred, Euler/CFG update
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cfm_prep_kernel(const float4* __restrict__ x1, const float4* __restrict__ z,
                                                       const float* __restrict__ t, float4* __restrict__ y,
                                                       long per_b4, long total4, float one_minus_sigma) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const float tt = t[i / per_b4];
    const float cz = 1.f - one_minus_sigma * tt;
    const float4 a = x1[i], n = z[i];
    y[i] = make_float4(cz * n.x + tt * a.x, cz * n.y + tt * a.y, cz * n.z + tt * a.z, cz * n.w + tt * a.w);
  }
}
__global__ void __launch_bounds__(256) cfm_prep_tail_kernel(const float* __restrict__ x1, const float* __restrict__ z,
                                                            const float* __restrict__ t, float* __restrict__ y,
                                                            long per_b, long total, float one_minus_sigma) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float tt = t[i / per_b];
    y[i] = (1.f - one_minus_sigma * tt) * z[i] + tt * x1[i];
  }
}
int launch_cfm_prep(const float* x1, const float* z, const float* t, float* y, int B, int T, float sigma_min,
                    cudaStream_t st) {
  const long per_b = 80L * T, total = per_b * B;
  const float oms = 1.f - sigma_min;
  if (per_b % 4 == 0) {
    const long total4 = total / 4;
    const unsigned grid = (unsigned)min((total4 + 255) / 256, 148L * 8);
    cfm_prep_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(x1), reinterpret_cast<const float4*>(z), t,
                                          reinterpret_cast<float4*>(y), per_b / 4, total4, oms);
  } else {
    const unsigned grid = (unsigned)min((total + 255) / 256, 148L * 8);
    cfm_prep_tail_kernel<<<grid, 256, 0, st>>>(x1, z, t, y, per_b, total, oms);
  }
  LAUNCH_RET();
}

__global__ void __launch_bounds__(256) wsum_kernel(const float* __restrict__ w, long n, float* __restrict__ scal,
                                                   const float* __restrict__ wsum_dev) {
  if (wsum_dev) {   // this call covers one shard: the normaliser is the whole batch's sum(w)
    if (threadIdx.x == 0) scal[0] = wsum_dev[0];
    return;
  }
  __shared__ float red[8];
  float s = 0.f;
  for (long i = threadIdx.x; i < n; i += 256) s += w[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    scal[0] = tot;
  }
}

// block = (b, 32-frame tile); reads pred/x1/z coalesced along t, writes dpred token-major
__global__ void __launch_bounds__(256) cfm_loss_kernel(const float* __restrict__ pred, const float* __restrict__ x1,
                                                       const float* __restrict__ z, const float* __restrict__ w,
                                                       const float* __restrict__ mask, const float* __restrict__ scal,
                                                       float* __restrict__ partials, uint32_t* __restrict__ dpred,
                                                       int T, float one_minus_sigma, float loss_scale, int bf) {
  __shared__ float tile[80][33];
  __shared__ float red[8];
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int t = t0 + tx;
  const bool ok = t < T;
  const float wt = ok ? w[(long)b * T + t] : 0.f;
  const float mk = ok ? mask[(long)b * T + t] : 0.f;
  const float denom = scal[0] * 80.f;
  const float gcoef = denom > 0.f ? 2.f * wt * wt * mk * loss_scale / denom : 0.f;
  float acc = 0.f;
  for (int c = wy; c < 80; c += 8) {
    float g = 0.f;
    if (ok) {
      const long i = ((long)b * 80 + c) * T + t;
      const float u = x1[i] - one_minus_sigma * z[i];
      const float diff = pred[i] - u;
      const float dw = diff * wt;
      acc += dw * dw;
      g = gcoef * diff;
    }
    tile[c][tx] = g;
  }
  acc = warp_sum(acc);
  if (tx == 0) red[wy] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    partials[(long)blockIdx.y * gridDim.x + blockIdx.x] = tot;
  }
  if (dpred) {
    for (int idx = threadIdx.x; idx < 32 * 64; idx += 256) {
      const int r = idx >> 6, cp = idx & 63;  // 64 pairs = 128 columns
      if (t0 + r < T) {
        const float a = (2 * cp < 80) ? tile[2 * cp][r] : 0.f;
        const float c2 = (2 * cp + 1 < 80) ? tile[2 * cp + 1][r] : 0.f;
        dpred[((long)b * T + t0 + r) * 64 + cp] = pack2_h16(a, c2, bf);
      }
    }
  }
}
__global__ void __launch_bounds__(256) loss_final_kernel(const float* __restrict__ partials, int n,
                                                         float* __restrict__ scal) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)partials[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    scal[1] = (float)tot;
    const float denom = scal[0] * 80.f;
    scal[2] = denom > 0.f ? (float)(tot / (double)denom) : 0.f;
  }
}
int launch_cfm_loss(const float* pred, const float* x1, const float* z, const float* w, const float* mask,
                    float* scal, float* partials, void* dpred, int B, int T, float sigma_min, float loss_scale,
                    int bf16, const float* wsum_dev, cudaStream_t st) {
  wsum_kernel<<<1, 256, 0, st>>>(w, (long)B * T, scal, wsum_dev);
  dim3 grid((T + 31) / 32, B);
  cfm_loss_kernel<<<grid, 256, 0, st>>>(pred, x1, z, w, mask, scal, partials, reinterpret_cast<uint32_t*>(dpred), T,
                                        1.f - sigma_min, loss_scale, bf16);
  loss_final_kernel<<<1, 256, 0, st>>>(partials, (int)(grid.x * grid.y), scal);
  LAUNCH_RET();
}

__global__ void __launch_bounds__(256) euler_update_kernel(float* __restrict__ x, const float* __restrict__ d,
                                                           const float* __restrict__ dt_arr, int step, float g,
                                                           long n) {
  const float dt = dt_arr[step];
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = (1.f + g) * d[i] - g * d[n + i];
    x[i] = x[i] + dt * v;
  }
}
int launch_euler_update(float* x, const float* d, const float* dt_arr, int step, float cfg_rate, long n,
                        cudaStream_t st) {
  const unsigned grid = (unsigned)min((n + 255) / 256, 148L * 4);
  euler_update_kernel<<<grid, 256, 0, st>>>(x, d, dt_arr, step, cfg_rate, n);
  LAUNCH_RET();
}

}  // namespace cvflow
