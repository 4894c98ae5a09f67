// The inputs of the path: the length regulator and the tensors the flow model prepares for
// ConditionalCFM.compute_loss, as fp32 CUDA-core kernels (the layers are 80 channels wide: 0.5 GFLOP per
// convolution at 32 x 400 frames, FMA-pipe work, no tensor-core shape).
//
//   InterpolateRegulator (reference modules.py:800-837)
//     F.interpolate(mode='linear') -> 4 x [Conv1d(80,80,3,pad 1) -> GroupNorm(1,80) -> Mish] -> Conv1d(80,80,1) -> * mask
//   is five launches of ONE kernel template: a convolution over a 128-frame tile of one utterance whose PROLOGUE
//   produces the operand on the fly (linear interpolation of the encoder output with at::upsample_linear1d's index
//   arithmetic; GroupNorm + Mish of the previous layer's raw output from merged Chan partials) and whose EPILOGUE
//   writes the raw output plus this tile's GroupNorm partial {n, mean, M2} (so the statistics cost no extra pass), or
//   the masked / text-blinded result in the layout compute_loss wants. The backward (frozen weights, LoRA fine-tuning:
//   input gradients only) is the same template with transposed-flipped weights: the GroupNorm backward is applied in
//   the prologue, the Mish / GroupNorm-affine backward and the two reduction sums in the epilogue; the interpolation
//   adjoint is a gather. Everything is deterministic (no atomics).
//
//   path_inputs_pack: mel normalisation, [B][T][80] -> [B][80][T], the prompt / silence-gap conditioning and the
//   pad mask in one launch from per-utterance descriptors (reference flow_model.py:266-387: ~100 small torch
//   launches and one host synchronisation per utterance).
//   spk_affine: F.normalize + Linear(192, 80) (flow_model.py:297-298).
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace cvflow {
namespace {

constexpr int kC = 80;        // channels (flow_model.py:715-720: InterpolateRegulator(channels = 80))
constexpr int kTT = 128;      // frames per CTA tile
constexpr int kXP = 132;      // pitch of the transposed operand tile xs[ci][kXP]: frames t0-1 .. t0+128, 16-byte aligned rows
constexpr int kOP = 81;       // pitch of the output tile os[frame][kOP]
constexpr int kCG = 16;       // channel groups (one warp each) of kCPW output channels ...
constexpr int kCPW = 5;
constexpr int kCW = 6;        // ... padded to 6 floats so a group's weights are three aligned 8-byte loads
constexpr int kThreads = 512; // 16 warps: four per scheduler hide the shared-memory / FMA latency of the 20-accumulator inner loop
constexpr float kGnEps = 1e-5f;

enum { PRO_INTERP = 0, PRO_GN_MISH = 1, PRO_MASK = 2, PRO_GN_BWD = 3 };
enum { EPI_STATS = 0, EPI_MASK = 1, EPI_MISH_BWD = 2, EPI_PLAIN = 3 };

struct LerpTap { int i0, i1; float w0, w1; };
// at::native::area_pixel_compute_source_index (align_corners = false, not cubic) and upsample_linear1d's weights, in
// fp32 with ATen's single fused multiply-add: indices and weights are bit-identical to F.interpolate's
// (tests/test_oracle.py pins the oracle's taps to torch, tests/test_path_inputs_gpu.py this kernel to the oracle).
__device__ __forceinline__ LerpTap lerp_tap(int j, int n_in, int n_out) {
  const float scale = __fdiv_rn((float)n_in, (float)n_out);
  float s = __fmaf_rn(scale, __fadd_rn((float)j, 0.5f), -0.5f);
  if (s < 0.f) s = 0.f;
  LerpTap t;
  t.i0 = min((int)s, n_in - 1);
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
  t.w1 = fminf(fmaxf(__fsub_rn(s, (float)t.i0), 0.f), 1.f);
  t.w0 = __fsub_rn(1.f, t.w1);
  return t;
}

// Mish and its derivative in full fp32 (these layers are not MUFU-bound; PyTorch: x * tanh(softplus(x)), threshold 20)
__device__ __forceinline__ float mish_acc(float x) {
  if (x > 20.f) return x;
  const float e = expf(x), n = e * (e + 2.f);
  return x * (n / (n + 2.f));
}
__device__ __forceinline__ float mish_grad_acc(float x) {
  if (x > 20.f) return 1.f;
  const float e = expf(x), n = e * (e + 2.f);
  const float tsp = n / (n + 2.f), sig = e / (1.f + e);
  return tsp + x * sig * (1.f - tsp * tsp);
}

struct RegConv {
  // operand
  const float* in;          // INTERP: src [B][n_src][80] | GN_MISH: raw y [B][T][80] | MASK: dout | GN_BWD: G [B][T][80]
  const float* in_y;        // GN_BWD: the raw output of the layer whose GroupNorm is differentiated (x^)
  int in_channel_major;     // MASK: dout is [B][80][T]
  int B, T, n_src, n_seg;
  RegSeg seg[4];
  const int* lens;          // MASK / EPI_MASK: frames >= lens[b] are zero (nullable)
  const int* blind;         // MASK / EPI_MASK: frames < blind[b] are zero (nullable)
  const float* part_in;     // GN_MISH / GN_BWD: Chan partials [B][NT][3] of the operand's layer
  const float* spart_in;    // GN_BWD: [B][NT][2] partial sums of G and G x^
  const float* gamma_in;    // GN_MISH
  const float* beta_in;
  const float* w;           // weight image [ci][tap][16][6]
  const float* bias;        // nullable
  // result
  float* out;
  int out_channel_major;    // EPI_MASK
  float* part_out;          // EPI_STATS: [B][NT][3]
  const float* y_prev;      // EPI_MISH_BWD: raw output, partials and affine of the layer below
  const float* part_prev;
  const float* gamma_prev;
  const float* beta_prev;
  float* spart_out;         // EPI_MISH_BWD: [B][NT][2]
};

// {n, mean, M2} partials of one utterance -> mean, rstd (biased variance, nn.GroupNorm). Called by one warp.
__device__ __forceinline__ void merge_partials(const float* part, int nt, float* mean_out, float* rstd_out) {
  const int lane = threadIdx.x & 31;
  float n = 0.f, m = 0.f, m2 = 0.f;
  for (int i = lane; i < nt; i += 32) {     // nt <= 32 for T <= 4096; the loop keeps longer inputs correct
    const float nb = part[i * 3], mb = part[i * 3 + 1], m2b = part[i * 3 + 2];
    const float nn = n + nb;
    if (nn > 0.f) {
      const float d = mb - m;
      m += d * (nb / nn);
      m2 += m2b + d * d * (n * nb / nn);
      n = nn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, m, o),
                m2b = __shfl_xor_sync(0xffffffffu, m2, o);
    const float nn = n + nb;
    if (nn > 0.f) {
      // symmetric form: both lanes of a pair compute the same merged triple
      const float mean = (n * m + nb * mb) / nn;
      const float d = mb - m;
      m2 = m2 + m2b + d * d * (n * nb / nn);
      m = mean;
      n = nn;
    }
  }
  if (lane == 0) {
    *mean_out = m;
    *rstd_out = rsqrtf(m2 / fmaxf(n, 1.f) + kGnEps);
  }
}

template <int TAPS, int PRO, int EPI>
__global__ void __launch_bounds__(kThreads, 1) reg_conv_kernel(const RegConv p) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                          // [80][kXP]; reused as the output tile os[128][kOP]
  float* ws = smem + kC * kXP;               // [80][TAPS][8][12]
  __shared__ float sc[8];                    // 0 mean_in, 1 rstd_in, 2 m1, 3 m2, 4 mean_prev, 5 rstd_prev
  __shared__ float red[2][kThreads / 32];
  __shared__ __align__(8) unsigned long long wbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, tile = blockIdx.x, t0 = tile * kTT;
  const int T = p.T, nt = (T + kTT - 1) / kTT;

  // the weight image is a constant of the model: one thread requests it as bulk copies (all 92 KB in flight at once)
  // while the previous kernel drains; it lands under the prologue and is waited for right before the inner loop
  constexpr uint32_t kWBytes = (uint32_t)(kC * TAPS * kCG * kCW * sizeof(float));
  if (tid == 0) {
    mbar_init(smem_u32(&wbar), 1);
    fence_barrier_init();
    mbar_expect_tx(smem_u32(&wbar), kWBytes);
    constexpr uint32_t kChunk = kWBytes / TAPS / 2;     // 15,360-byte pieces
#pragma unroll
    for (uint32_t o = 0; o < kWBytes; o += kChunk)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ws) + o),
                   "l"(reinterpret_cast<const uint8_t*>(p.w) + o), "r"(kChunk), "r"(smem_u32(&wbar))
                   : "memory");
  }
  pdl_wait();

  if (PRO == PRO_GN_MISH || PRO == PRO_GN_BWD) {
    if (warp == 0) merge_partials(p.part_in + (long)b * nt * 3, nt, &sc[0], &sc[1]);
  }
  if (PRO == PRO_GN_BWD) {
    if (warp == 1) {
      float s1 = 0.f, s2 = 0.f;
      for (int i = lane; i < nt; i += 32) { s1 += p.spart_in[((long)b * nt + i) * 2]; s2 += p.spart_in[((long)b * nt + i) * 2 + 1]; }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { const float inv = 1.f / ((float)T * kC); sc[2] = s1 * inv; sc[3] = s2 * inv; }
    }
  }
  if (EPI == EPI_MISH_BWD) {
    if (warp == 2) merge_partials(p.part_prev + (long)b * nt * 3, nt, &sc[4], &sc[5]);
  }
  if (PRO == PRO_GN_MISH || PRO == PRO_GN_BWD || EPI == EPI_MISH_BWD) __syncthreads();

  // ---- prologue: operand tile, transposed to xs[ci][frame], frames outside [0, T) are the convolution's zero padding ----
  const int len_b = p.lens ? p.lens[b] : T;
  const int blind_b = p.blind ? p.blind[b] : 0;
  constexpr int kFr = kTT + 2;
  // every global load of the tile is issued before the first value is used (one L2 round trip, not 21 dependent ones)
  constexpr int kPer = (kC * kFr + kThreads - 1) / kThreads;
  const bool cm_in = PRO == PRO_MASK && p.in_channel_major;
  float ra[kPer], rb[kPer], rw[kPer];
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    const int idx = tid + u * kThreads;
    ra[u] = 0.f; rb[u] = 0.f; rw[u] = 0.f;
    if (idx < kC * kFr) {
      const int tt = cm_in ? idx % kFr : idx / kC, c = cm_in ? idx / kFr : idx % kC, t = t0 - 1 + tt;
      if (t >= 0 && t < T) {
        if (PRO == PRO_INTERP) {
#pragma unroll
          for (int sg = 0; sg < 4; ++sg) {
            if (sg < p.n_seg && t >= p.seg[sg].dst0 && t < p.seg[sg].dst0 + p.seg[sg].dstn) {
              const LerpTap k = lerp_tap(t - p.seg[sg].dst0, p.seg[sg].srcn, p.seg[sg].dstn);
              const float* row = p.in + ((long)b * p.n_src + p.seg[sg].src0) * kC + c;
              ra[u] = row[(long)k.i0 * kC];
              rb[u] = row[(long)k.i1 * kC];
              rw[u] = k.w1;
            }
          }
        } else if (PRO == PRO_MASK) {
          if (t >= blind_b && t < len_b) ra[u] = cm_in ? p.in[((long)b * kC + c) * T + t] : p.in[((long)b * T + t) * kC + c];
        } else {
          ra[u] = p.in[((long)b * T + t) * kC + c];
          if (PRO == PRO_GN_BWD) rb[u] = p.in_y[((long)b * T + t) * kC + c];
        }
      }
    }
  }
  {
    const float mean = sc[0], rstd = sc[1], m1 = sc[2], m2 = sc[3];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int idx = tid + u * kThreads;
      if (idx < kC * kFr) {
        const int tt = cm_in ? idx % kFr : idx / kC, c = cm_in ? idx / kFr : idx % kC, t = t0 - 1 + tt;
        float v = 0.f;
        if (t >= 0 && t < T) {
          if (PRO == PRO_INTERP) v = __fsub_rn(1.f, rw[u]) * ra[u] + rw[u] * rb[u];
          else if (PRO == PRO_GN_MISH) v = mish_acc((ra[u] - mean) * rstd * p.gamma_in[c] + p.beta_in[c]);
          else if (PRO == PRO_MASK) v = ra[u];
          else v = rstd * (ra[u] - m1 - (rb[u] - mean) * rstd * m2);   // GroupNorm backward: rstd (G - mean(G) - x^ mean(G x^))
        }
        xs[c * kXP + tt] = v;
      }
    }
  }
  __syncthreads();             // operand tile complete (and the barrier initialisation visible to every thread)
  mbar_wait(smem_u32(&wbar), 0u);

  // ---- 4 frames x 5 output channels per thread: warp = channel group (weights broadcast), lane = frame quad ----
  float acc[4][kCPW];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int j = 0; j < kCPW; ++j) acc[f][j] = 0.f;
  {
    const float* xr = xs + 4 * lane + (TAPS == 1 ? 1 : 0);
    const float* wr = ws + warp * kCW;
#pragma unroll 2
    for (int ci = 0; ci < kC; ++ci) {
      float x[6];
      if (TAPS == 1) {
        x[0] = xr[0]; x[1] = xr[1]; x[2] = xr[2]; x[3] = xr[3]; x[4] = 0.f; x[5] = 0.f;
      } else {
        const float4 xa = *reinterpret_cast<const float4*>(xr);
        const float2 xb = *reinterpret_cast<const float2*>(xr + 4);
        x[0] = xa.x; x[1] = xa.y; x[2] = xa.z; x[3] = xa.w; x[4] = xb.x; x[5] = xb.y;
      }
#pragma unroll
      for (int k = 0; k < TAPS; ++k) {
        const float2 wa = *reinterpret_cast<const float2*>(wr + k * kCG * kCW);
        const float2 wb = *reinterpret_cast<const float2*>(wr + k * kCG * kCW + 2);
        const float2 wc = *reinterpret_cast<const float2*>(wr + k * kCG * kCW + 4);
        float w[kCPW];
        w[0] = wa.x; w[1] = wa.y; w[2] = wb.x; w[3] = wb.y; w[4] = wc.x;
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int j = 0; j < kCPW; ++j) acc[f][j] = fmaf(w[j], x[f + k], acc[f][j]);
      }
      xr += kXP;
      wr += TAPS * kCG * kCW;
    }
  }
  __syncthreads();          // every warp is done with xs: it becomes the output tile
  float* os = xs;
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int j = 0; j < kCPW; ++j) {
      const int c = warp * kCPW + j;
      os[(4 * lane + f) * kOP + c] = acc[f][j] + (p.bias ? p.bias[c] : 0.f);
    }
  __syncthreads();
  pdl_launch();

  const int nv = min(kTT, T - t0);          // valid frames of this tile
  if (EPI == EPI_STATS || EPI == EPI_PLAIN) {
    float s = 0.f;
    for (int idx = tid; idx < nv * kC; idx += kThreads) {
      const int tt = idx / kC, c = idx - tt * kC;
      const float v = os[tt * kOP + c];
      p.out[((long)b * T + t0 + tt) * kC + c] = v;
      s += v;
    }
    if (EPI == EPI_STATS) {                  // two-pass partial of the tile (Chan form: merged by the consumer)
      s = warp_sum(s);
      if (lane == 0) red[0][warp] = s;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < kThreads / 32; ++i) tot += red[0][i];
      const float mean = tot / (float)(nv * kC);
      float q = 0.f;
      for (int idx = tid; idx < nv * kC; idx += kThreads) {
        const int tt = idx / kC, c = idx - tt * kC;
        const float d = os[tt * kOP + c] - mean;
        q += d * d;
      }
      q = warp_sum(q);
      if (lane == 0) red[1][warp] = q;
      __syncthreads();
      if (tid == 0) {
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < kThreads / 32; ++i) m2 += red[1][i];
        float* o = p.part_out + ((long)b * nt + tile) * 3;
        o[0] = (float)(nv * kC); o[1] = mean; o[2] = m2;
      }
    }
  } else if (EPI == EPI_MASK) {
    if (p.out_channel_major) {
      for (int idx = tid; idx < kC * kTT; idx += kThreads) {
        const int c = idx / kTT, tt = idx - c * kTT, t = t0 + tt;
        if (tt < nv) p.out[((long)b * kC + c) * T + t] = (t >= blind_b && t < len_b) ? os[tt * kOP + c] : 0.f;
      }
    } else {
      for (int idx = tid; idx < nv * kC; idx += kThreads) {
        const int tt = idx / kC, c = idx - tt * kC, t = t0 + tt;
        p.out[((long)b * T + t) * kC + c] = (t >= blind_b && t < len_b) ? os[tt * kOP + c] : 0.f;
      }
    }
  } else {   // EPI_MISH_BWD: acc = dL/d(activation of the layer below); G = acc mish'(u) gamma, sums of G and G x^
    const float mean = sc[4], rstd = sc[5];
    float s1 = 0.f, s2 = 0.f;
    constexpr int kPerE = kTT * kC / kThreads;
    float yp[kPerE];
#pragma unroll
    for (int u = 0; u < kPerE; ++u) {      // the tile's rows are contiguous in y_prev / out: element idx of the tile
      const int idx = tid + u * kThreads;
      yp[u] = idx < nv * kC ? p.y_prev[((long)b * T + t0) * kC + idx] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kPerE; ++u) {
      const int idx = tid + u * kThreads;
      if (idx < nv * kC) {
        const int tt = idx / kC, c = idx - tt * kC;
        const float xh = (yp[u] - mean) * rstd;
        const float ga = p.gamma_prev[c];
        const float g = os[tt * kOP + c] * mish_grad_acc(ga * xh + p.beta_prev[c]) * ga;
        p.out[((long)b * T + t0) * kC + idx] = g;
        s1 += g;
        s2 += g * xh;
      }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, c2 = 0.f;
#pragma unroll
      for (int i = 0; i < kThreads / 32; ++i) { a += red[0][i]; c2 += red[1][i]; }
      float* o = p.spart_out + ((long)b * nt + tile) * 2;
      o[0] = a; o[1] = c2;
    }
  }
}

// Adjoint of the interpolation: dsrc[b][s][c] = sum over the frames j whose taps touch s. Gather form: the source index
// is monotone in j, so the candidates are a short window around (s + 0.5) / scale.
__global__ void __launch_bounds__(kC * 4) reg_interp_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dsrc, int T,
                                                                int n_src, int n_seg, RegSeg s0, RegSeg s1, RegSeg s2, RegSeg s3) {
  pdl_wait();
  const int c = threadIdx.x % kC, s = blockIdx.x * 4 + threadIdx.x / kC, b = blockIdx.y;
  if (s >= n_src) return;
  const RegSeg segs[4] = {s0, s1, s2, s3};
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q >= n_seg) break;
    const RegSeg g = segs[q];
    if (s < g.src0 || s >= g.src0 + g.srcn) continue;
    const int sl = s - g.src0;
    const float inv = (float)g.dstn / (float)g.srcn;
    int lo = (int)floorf(((float)sl - 0.5f) * inv - 0.5f) - 2, hi = (int)ceilf(((float)sl + 1.5f) * inv - 0.5f) + 2;
    lo = max(lo, 0); hi = min(hi, g.dstn - 1);
    for (int j = lo; j <= hi; ++j) {
      const LerpTap k = lerp_tap(j, g.srcn, g.dstn);
      const float v = dx[((long)b * T + g.dst0 + j) * kC + c];
      if (k.i0 == sl) acc += k.w0 * v;
      if (k.i1 == sl) acc += k.w1 * v;
    }
  }
  dsrc[((long)b * n_src + s) * kC + c] = acc;
}

template <int TAPS, int PRO, int EPI>
int launch_reg(const RegConv& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)(kC * kXP + kC * TAPS * kCG * kCW) * sizeof(float);
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(reg_conv_kernel<TAPS, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    done = true;
  }
  const int nt = (p.T + kTT - 1) / kTT;
  cudaError_t e = launch_pdl(reg_conv_kernel<TAPS, PRO, EPI>, dim3(nt, p.B), dim3(kThreads), smem, st, p);
  return e == cudaSuccess ? 0 : -(int)e;
}

// ------------------------------------------------------------------------------------------
// path_inputs_pack / spk_affine
// ------------------------------------------------------------------------------------------
// desc[b] = {len, prompt frames copied, silence-gap frames, flags (bit 0: prompt taken from cross_mel)}
__global__ void __launch_bounds__(256) path_inputs_pack_kernel(const float* __restrict__ feat, const float* __restrict__ cross,
                                                               int cross_T, const int* __restrict__ desc, float mel_mean,
                                                               float mel_std, float silence, float* __restrict__ x1,
                                                               float* __restrict__ cond, float* __restrict__ mask, int T) {
  __shared__ float tf[32][kC + 1], tc[32][kC + 1];
  pdl_wait();
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int len = desc[b * 4], pl = desc[b * 4 + 1], gap = desc[b * 4 + 2], fl = desc[b * 4 + 3];
  for (int idx = threadIdx.x; idx < 32 * kC; idx += 256) {
    const int tt = idx / kC, c = idx - tt * kC, t = t0 + tt;
    float f = 0.f, cv = 0.f;
    if (t < T) {
      f = (feat[((long)b * T + t) * kC + c] - mel_mean) / mel_std;
      if (t < pl) cv = (fl & 1) ? (t < cross_T ? (cross[((long)b * cross_T + t) * kC + c] - mel_mean) / mel_std : 0.f) : f;
      else if (t < pl + gap) cv = silence;
    }
    tf[tt][c] = f;
    tc[tt][c] = cv;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * kC; idx += 256) {
    const int c = idx >> 5, tt = idx & 31, t = t0 + tt;
    if (t < T) {
      x1[((long)b * kC + c) * T + t] = tf[tt][c];
      cond[((long)b * kC + c) * T + t] = tc[tt][c];
    }
  }
  if (threadIdx.x < 32 && t0 + threadIdx.x < T) mask[(long)b * T + t0 + threadIdx.x] = (t0 + (int)threadIdx.x < len) ? 1.f : 0.f;
}

// out[b][n] = sum_k W[n][k] e[b][k] / max(||e[b]||, 1e-12) + bias[n]   (F.normalize(dim=1) then nn.Linear)
__global__ void __launch_bounds__(128) spk_affine_kernel(const float* __restrict__ e, const float* __restrict__ W,
                                                         const float* __restrict__ bias, float* __restrict__ out, int K, int N) {
  extern __shared__ float se[];
  __shared__ float part[4];
  pdl_wait();
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float q = 0.f;
  for (int k = threadIdx.x; k < K; k += 128) { const float v = e[(long)b * K + k]; se[k] = v; q += v * v; }
  q = warp_sum(q);
  if (lane == 0) part[warp] = q;
  __syncthreads();
  const float inv = 1.f / fmaxf(sqrtf(part[0] + part[1] + part[2] + part[3]), 1e-12f);
  for (int n = warp; n < N; n += 4) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(W[(long)n * K + k], se[k], s);
    s = warp_sum(s);
    if (lane == 0) out[(long)b * N + n] = s * inv + (bias ? bias[n] : 0.f);
  }
}

}  // namespace

long regulator_saved_floats(int B, int T) {
  const long nt = (T + kTT - 1) / kTT;
  return 4L * B * T * kC + 4L * B * nt * 3;
}
long regulator_scratch_floats(int B, int T) {
  const long nt = (T + kTT - 1) / kTT;
  return 2L * B * T * kC + 2L * B * nt * 2;
}

int launch_regulator_forward(const RegulatorWeights& w, const RegulatorIO& io, cudaStream_t st) {
  const int B = io.B, T = io.T;
  const long nt = (T + kTT - 1) / kTT, plane = (long)B * T * kC;
  float* y[4] = {io.saved, io.saved + plane, io.saved + 2 * plane, io.saved + 3 * plane};
  float* part = io.saved + 4 * plane;
  RegConv p{};
  p.B = B; p.T = T; p.n_src = io.n_src; p.n_seg = io.n_seg;
  for (int s = 0; s < 4; ++s) p.seg[s] = io.seg[s];
  int r;
  // layer 0: interpolation in the prologue
  p.in = io.src; p.w = w.wf[0]; p.bias = w.bias[0]; p.out = y[0]; p.part_out = part;
  if ((r = launch_reg<3, PRO_INTERP, EPI_STATS>(p, st))) return r;
  for (int l = 1; l < 4; ++l) {
    p.in = y[l - 1]; p.part_in = part + (l - 1) * B * nt * 3; p.gamma_in = w.gamma[l - 1]; p.beta_in = w.beta[l - 1];
    p.w = w.wf[l]; p.bias = w.bias[l]; p.out = y[l]; p.part_out = part + l * B * nt * 3;
    if ((r = launch_reg<3, PRO_GN_MISH, EPI_STATS>(p, st))) return r;
  }
  p.in = y[3]; p.part_in = part + 3 * B * nt * 3; p.gamma_in = w.gamma[3]; p.beta_in = w.beta[3];
  p.w = w.wf[4]; p.bias = w.bias[4]; p.out = io.out; p.part_out = nullptr;
  p.lens = io.lens; p.blind = io.blind; p.out_channel_major = io.channel_major;
  return launch_reg<1, PRO_GN_MISH, EPI_MASK>(p, st);
}

int launch_regulator_backward(const RegulatorWeights& w, const RegulatorIO& io, const float* dout, float* dsrc, float* scratch,
                              cudaStream_t st) {
  const int B = io.B, T = io.T;
  const long nt = (T + kTT - 1) / kTT, plane = (long)B * T * kC;
  const float* y[4] = {io.saved, io.saved + plane, io.saved + 2 * plane, io.saved + 3 * plane};
  const float* part = io.saved + 4 * plane;
  float* G[2] = {scratch, scratch + plane};
  float* sp[2] = {scratch + 2 * plane, scratch + 2 * plane + B * nt * 2};
  RegConv p{};
  p.B = B; p.T = T; p.n_src = io.n_src; p.n_seg = io.n_seg;
  int r;
  // 1x1 conv: d(activation 3) from the masked output gradient, then through Mish / GroupNorm-affine of layer 3
  p.in = dout; p.in_channel_major = io.channel_major; p.lens = io.lens; p.blind = io.blind;
  p.w = w.wb[4]; p.out = G[1]; p.y_prev = y[3]; p.part_prev = part + 3 * B * nt * 3; p.gamma_prev = w.gamma[3];
  p.beta_prev = w.beta[3]; p.spart_out = sp[1];
  if ((r = launch_reg<1, PRO_MASK, EPI_MISH_BWD>(p, st))) return r;
  p.lens = nullptr; p.blind = nullptr; p.in_channel_major = 0;
  for (int l = 3; l >= 1; --l) {            // conv l: input = activation l-1
    p.in = G[l & 1]; p.in_y = y[l]; p.part_in = part + l * B * nt * 3; p.spart_in = sp[l & 1];
    p.w = w.wb[l]; p.out = G[(l - 1) & 1]; p.y_prev = y[l - 1]; p.part_prev = part + (l - 1) * B * nt * 3;
    p.gamma_prev = w.gamma[l - 1]; p.beta_prev = w.beta[l - 1]; p.spart_out = sp[(l - 1) & 1];
    if ((r = launch_reg<3, PRO_GN_BWD, EPI_MISH_BWD>(p, st))) return r;
  }
  // conv 0: input = the interpolated encoder output
  p.in = G[0]; p.in_y = y[0]; p.part_in = part; p.spart_in = sp[0]; p.w = w.wb[0]; p.out = G[1];
  p.y_prev = nullptr; p.part_prev = nullptr; p.spart_out = nullptr;
  if ((r = launch_reg<3, PRO_GN_BWD, EPI_PLAIN>(p, st))) return r;
  cudaError_t e = launch_pdl(reg_interp_bwd_kernel, dim3((io.n_src + 3) / 4, B), dim3(kC * 4), 0, st, (const float*)G[1], dsrc, T,
                             io.n_src, io.n_seg, io.seg[0], io.seg[1], io.seg[2], io.seg[3]);
  return e == cudaSuccess ? 0 : -(int)e;
}

int launch_path_inputs_pack(const float* feat, const float* cross, int cross_T, const int* desc, float mel_mean, float mel_std,
                            float silence, float* x1, float* cond, float* mask, int B, int T, cudaStream_t st) {
  cudaError_t e = launch_pdl(path_inputs_pack_kernel, dim3((T + 31) / 32, B), dim3(256), 0, st, feat, cross, cross_T, desc, mel_mean,
                             mel_std, silence, x1, cond, mask, T);
  return e == cudaSuccess ? 0 : -(int)e;
}

int launch_spk_affine(const float* e, const float* W, const float* bias, float* out, int B, int K, int N, cudaStream_t st) {
  cudaError_t err = launch_pdl(spk_affine_kernel, dim3(B), dim3(128), (size_t)K * sizeof(float), st, e, W, bias, out, K, N);
  return err == cudaSuccess ? 0 : -(int)err;
}

}  // namespace cvflow
