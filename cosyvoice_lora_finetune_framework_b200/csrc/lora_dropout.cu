// lora_dropout > 0 on the attn1 q/k/v LoRA branches (reference lora.py:66-74: y = W x + s B (A drop(x)), one
// independent nn.Dropout per LoRALinear, i.e. three keep masks per attention block; the reference's default
// configuration, config.py:207-216 uses 0.05). The low-rank branch cannot be folded into the frozen operand then; it
// stays rank-r work on the CUDA cores, FUSED INTO THE LAYERNORM PASSES that already hold the token row in registers:
//
//   forward   ln_lora_drop_fwd    x~ = LN1(h);  keep bits drawn once and stored bit-packed (96 B per token);
//                                 u_d[m][p r + j] = 1/(1-p) sum_k keep_p[m][k] x~[m][k] A_p[j][k]
//                                 (u_d is the second K segment of the q/k/v GEMM against [W0 | s B_cat])
//   backward  ln_lora_drop_bwd    dx~ = dx + s/(1-p) sum_p keep_p o (v_p A_p)   (v = dY B_blk^T from the dgrad GEMM),
//                                 then the LayerNorm backward on dx~ in the same pass
//   wgrad     lora_wgrad_a_drop   dA_p[j][k] = s/(1-p) sum_m v[m][p r + j] keep_p[m][k] x~[m][k]  (partials in the layout
//                                 of the tensor-core wgrad kernel, whose un-masked x~^T v they replace);
//                                 dB = s dY^T u_d comes from that kernel unchanged.
//
// Masks: one splitmix64 hash of (seed, block, projection, token, feature quad) yields four 16-bit fields, one per
// feature; a feature is dropped iff its field < round(p 2^16) (p = 0.05 -> 3277 / 65536, relative error 6e-5).
// The forward stores the decisions as bits (8 words per token and projection), the two backward kernels read them back
// instead of re-hashing. Parity tests pass explicit byte masks (LoraDropSpec::dbg) to the forward kernel instead.
#include "kernels.h"
#include "common.cuh"
#include "rowops.cuh"

namespace cvflow {

#define LAUNCH_RET() do { cudaError_t e_ = cudaGetLastError(); return e_ == cudaSuccess ? 0 : -(int)e_; } while (0)

// 4 keep decisions (bit j = feature 4 quad + j) of one (block, projection, token, quad) counter
__device__ __forceinline__ uint32_t lora_keep4(unsigned long long seed, unsigned long long ctr, unsigned thr16) {
  unsigned long long z = seed + (ctr + 1ull) * 0x9E3779B97F4A7C15ull;   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const unsigned lo = (unsigned)z, hi = (unsigned)(z >> 32);
  return (uint32_t)((lo & 0xffffu) >= thr16) | ((uint32_t)((lo >> 16) >= thr16) << 1) |
         ((uint32_t)((hi & 0xffffu) >= thr16) << 2) | ((uint32_t)((hi >> 16) >= thr16) << 3);
}

// sum over the warp of v[i] for every i in [0, 32): lane l returns the total of v[l] (31 shuffles instead of 160)
__device__ __forceinline__ float reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}
// 16 values: lanes l and l ^ 16 both return the total of v[l & 15]
__device__ __forceinline__ float reduce_scatter16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// stage the 3r rows of A_cat (16-bit [64][256], rows >= 3r unused) in shared memory
template <int R>
__device__ __forceinline__ void load_acat(uint16_t* As, const uint16_t* __restrict__ acat) {
  for (int i = threadIdx.x; i < 3 * R * 32; i += blockDim.x)
    reinterpret_cast<uint4*>(As)[i] = reinterpret_cast<const uint4*>(acat)[i];
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// forward: LayerNorm + mask draw + masked down-projection. Warp per token, TOK tokens per pass so that every row of
// A_cat fetched (and widened to fp32) from shared memory serves TOK dot products.
// ------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256, 3) ln_lora_drop_fwd_kernel(const float* __restrict__ h, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const uint16_t* __restrict__ acat,
                                                               uint16_t* __restrict__ x1, uint32_t* __restrict__ ud,
                                                               uint32_t* __restrict__ bits, long M, int bf, const LoraDropSpec d) {
  constexpr int NV = 3 * R;
  constexpr int TOK = 1;     // measured: two tokens per pass (shared A-row fetches) is no faster, registers cost occupancy
  __shared__ __align__(16) uint16_t As[NV * 256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long stride = (long)gridDim.x * 8;
  // constants of the step (LayerNorm affine, the seed bumped at the start of the forward, the LoRA factor image written
  // by lora_merge: none is produced by the kernel launched just before this one) are fetched and staged while the
  // predecessor drains; only the token rows wait for it
  float g[8], b[8], xfirst[8];
  load8_f32(gamma + lane * 8, g);
  load8_f32(beta + lane * 8, b);
  const unsigned long long seed = d.dbg ? 0ull : d.seed[0];
  load_acat<R>(As, acat);
  pdl_wait();
  const long mfirst = (long)blockIdx.x * 8 + warp;
  if (mfirst < M) load8_f32(h + mfirst * 256 + lane * 8, xfirst);
  pdl_launch();
  for (long mb = mfirst; mb < M; mb += stride * TOK) {
    float xv[TOK][8];
    uint32_t kb[TOK][3];
    float x[TOK][8];
#pragma unroll
    for (int t = 0; t < TOK; ++t) {      // all the row loads first
      const long m = mb + t * stride;
      if (m == mfirst) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[t][e] = xfirst[e];
      } else if (m < M) load8_f32(h + m * 256 + lane * 8, x[t]);
    }
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
      const long m = mb + t * stride;
      if (m >= M) {      // warp-uniform
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[t][e] = 0.f;
        kb[t][0] = kb[t][1] = kb[t][2] = 0u;
        continue;
      }
      const float rstd = row_center_rstd(x[t]);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[t][i] = x[t][i] * rstd * g[i] + b[i];
      const uint4 packed = pack8_h16(x[t], bf);
      *reinterpret_cast<uint4*>(x1 + m * 256 + lane * 8) = packed;
      unpack8_h16(packed, bf, xv[t]);          // the branch sees the 16-bit x~ the q/k/v GEMM sees
      // keep decisions of this lane's 8 features for the three projections
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const unsigned long long row = ((unsigned long long)(d.blk * 3 + p)) * (unsigned long long)d.mcap + (unsigned long long)m;
        uint32_t k = 0;
        if (d.dbg) {
          const uint2 by = *reinterpret_cast<const uint2*>(d.dbg + row * 256ull + (unsigned)(lane * 8));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            k |= (uint32_t)(((by.x >> (8 * e)) & 0xffu) != 0u) << e;
            k |= (uint32_t)(((by.y >> (8 * e)) & 0xffu) != 0u) << (4 + e);
          }
        } else {
          k = lora_keep4(seed, row * 64ull + (unsigned)(lane * 2), d.thr16) |
              (lora_keep4(seed, row * 64ull + (unsigned)(lane * 2 + 1), d.thr16) << 4);
        }
        kb[t][p] = k;
        uint32_t w = k << (8 * (lane & 3));   // word lane/4 of the projection's 8 words: bytes of lanes 4w .. 4w+3
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        if ((lane & 3) == 0) bits[m * 24 + p * 8 + (lane >> 2)] = w;
      }
    }
    // masked down-projection: this lane's share of the 3r dot products of each token, then one reduce-scatter per token
    float acc[TOK][32];
    float acc2[TOK][16];
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[t][i] = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc2[t][i] = 0.f;
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float xm[TOK][8];
#pragma unroll
      for (int t = 0; t < TOK; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) xm[t][e] = ((kb[t][p] >> e) & 1u) ? xv[t][e] : 0.f;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int c = p * R + j;
        float a[8];
        load8_h16(As + c * 256 + lane * 8, bf, a);
#pragma unroll
        for (int t = 0; t < TOK; ++t) {
          float s0 = xm[t][0] * a[0], s1 = xm[t][1] * a[1];
#pragma unroll
          for (int e = 2; e < 8; e += 2) { s0 = fmaf(xm[t][e], a[e], s0); s1 = fmaf(xm[t][e + 1], a[e + 1], s1); }
          if (c < 32) acc[t][c] = s0 + s1; else acc2[t][c - 32] = s0 + s1;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
      const long m = mb + t * stride;
      if (m >= M) continue;
      const float t32 = reduce_scatter32(acc[t], lane);
      float o0 = __shfl_sync(0xffffffffu, t32, (2 * lane) & 31), o1 = __shfl_sync(0xffffffffu, t32, (2 * lane + 1) & 31);
      if (2 * lane >= 32 || 2 * lane >= NV) { o0 = 0.f; o1 = 0.f; }
      if (NV > 32) {
        const float t16 = reduce_scatter16(acc2[t], lane);
        const float q0 = __shfl_sync(0xffffffffu, t16, (2 * lane) & 15), q1 = __shfl_sync(0xffffffffu, t16, (2 * lane + 1) & 15);
        if (2 * lane >= 32 && 2 * lane < NV) { o0 = q0; o1 = q1; }
      }
      ud[m * 32 + lane] = pack2_h16(o0 * d.inv_keep, o1 * d.inv_keep, bf);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: masked up-projection of v added to dx, then the LayerNorm backward; warp per token, TOK tokens per pass
// dxe: [M][320] 16-bit, columns [0,256) dx (from the dgrad GEMM on W0^T), columns [256, 256+3r) v = dY B_blk^T
// ------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256, 3) ln_lora_drop_bwd_kernel(const uint16_t* __restrict__ dxe, const uint16_t* __restrict__ acat,
                                                               const uint32_t* __restrict__ bits, const float* __restrict__ h_in,
                                                               const float* __restrict__ gamma, const float* __restrict__ dres,
                                                               float* __restrict__ dh, uint16_t* __restrict__ dh16, long M,
                                                               float sc, int bf) {
  constexpr int NV = 3 * R;
  constexpr int TOK = 1;     // measured: two tokens per pass is slower here (19.0 vs 14.8 us)
  __shared__ __align__(16) uint16_t As[NV * 256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float g[8];
  load8_f32(gamma + lane * 8, g);      // step constants: staged before the grid-dependency wait (see the forward kernel)
  load_acat<R>(As, acat);
  pdl_wait();
  pdl_launch();
  const long stride = (long)gridDim.x * 8;
  for (long mb = (long)blockIdx.x * 8 + warp; mb < M; mb += stride * TOK) {
    float v0[TOK], v1[TOK], d[TOK][8], x[TOK][8], r[TOK][8];
    uint32_t kw[TOK][3];
#pragma unroll
    for (int t = 0; t < TOK; ++t) {      // every load of the pass up front
      const long m = mb + t * stride;
      v0[t] = v1[t] = 0.f;
      kw[t][0] = kw[t][1] = kw[t][2] = 0u;
      if (m < M) {
        const uint16_t* row = dxe + m * 320;
        unpack2_h16(reinterpret_cast<const uint32_t*>(row + 256)[lane], bf, v0[t], v1[t]);
        load8_h16(row + lane * 8, bf, d[t]);
        load8_f32(h_in + m * 256 + lane * 8, x[t]);
        if (dres) load8_f32(dres + m * 256 + lane * 8, r[t]);
#pragma unroll
        for (int p = 0; p < 3; ++p) kw[t][p] = bits[m * 24 + p * 8 + (lane >> 2)];
      }
    }
    float acc[TOK][8];
#pragma unroll
    for (int t = 0; t < TOK; ++t)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float tt[TOK][8];
#pragma unroll
      for (int t = 0; t < TOK; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) tt[t][e] = 0.f;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int c = p * R + j;
        float a[8];
        load8_h16(As + c * 256 + lane * 8, bf, a);
#pragma unroll
        for (int t = 0; t < TOK; ++t) {
          const float vj = __shfl_sync(0xffffffffu, (c & 1) ? v1[t] : v0[t], c >> 1);
#pragma unroll
          for (int e = 0; e < 8; ++e) tt[t][e] = fmaf(vj, a[e], tt[t][e]);
        }
      }
#pragma unroll
      for (int t = 0; t < TOK; ++t) {
        const uint32_t kb = (kw[t][p] >> (8 * (lane & 3))) & 0xffu;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] += ((kb >> e) & 1u) ? tt[t][e] : 0.f;
      }
    }
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
      const long m = mb + t * stride;
      if (m >= M) continue;
#pragma unroll
      for (int e = 0; e < 8; ++e) d[t][e] += sc * acc[t][e];
      // LayerNorm backward (same arithmetic as layernorm_bwd_kernel, norm.cu)
      const float rstd = row_center_rstd(x[t]);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[t][i] *= rstd;
        d[t][i] *= g[i];
        s1 += d[t][i];
        s2 += d[t][i] * x[t][i];
      }
      s1 = warp_sum(s1) * (1.f / 256.f);
      s2 = warp_sum(s2) * (1.f / 256.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[t][i] = rstd * (d[t][i] - s1 - x[t][i] * s2) + (dres ? r[t][i] : 0.f);
      if (dh) store8_f32(dh + m * 256 + lane * 8, d[t]);
      if (dh16) store8_h16(dh16 + m * 256 + lane * 8, bf, d[t]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// masked dA partials: dA_p[j][k] = sum_m v[m][p r + j] keep_p[m][k] x~[m][k]. (64-token chunk, projection) CTAs; warp w
// owns the feature quarter w % 4 (a lane = 2 adjacent features, one 32-bit load per token) of the chunk's token half
// w / 4, 2 r accumulators per lane; the two halves are added through shared memory. Scratch [chunk][3][16][256]; a
// second kernel sums the chunks in fixed order (deterministic) into split 0 of the tensor-core wgrad kernel's part_a
// layout ([split][256][64]) and zeroes the other splits, replacing that kernel's un-masked x~^T v.
// ------------------------------------------------------------------------------------------
static constexpr int kWgaChunk = 64;
template <int R>
__global__ void __launch_bounds__(256) lora_wgrad_a_drop_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ v,
                                                                long ld_v, const uint32_t* __restrict__ bits,
                                                                float* __restrict__ scratch, long M, int bf) {
  __shared__ __align__(16) float vs[kWgaChunk][R];
  __shared__ float red[4][32][2 * R];
  pdl_wait();
  pdl_launch();
  const int c = blockIdx.x, p = blockIdx.y;
  const long m0 = (long)c * kWgaChunk;
  const int n = (int)(M - m0 < kWgaChunk ? M - m0 : kWgaChunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fq = warp & 3, th = warp >> 2;
  const int k0 = fq * 64 + lane * 2;
  float acc0[R], acc1[R];
#pragma unroll
  for (int j = 0; j < R; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
  const int t_begin = th * (kWgaChunk / 2);
  constexpr int NT = kWgaChunk / 2;
  uint32_t xws[NT], kws[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {      // all 2 x 32 loads of the warp in flight together (one L2 round trip)
    const int t = t_begin + i;
    xws[i] = 0u; kws[i] = 0u;
    if (t < n) {
      const long row = m0 + t;
      xws[i] = *reinterpret_cast<const uint32_t*>(x + row * 256 + k0);
      kws[i] = bits[row * 24 + p * 8 + 2 * fq + (lane >> 4)];
    }
  }
  for (int i = threadIdx.x; i < kWgaChunk * R; i += 256) {
    const int t = i / R, j = i - t * R;
    vs[t][j] = t < n ? h16_to_f32(v[(m0 + t) * ld_v + p * R + j], bf) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int t = t_begin + i;
    const uint32_t kb = (kws[i] >> ((2 * lane) & 31)) & 3u;
    float x0, x1;
    unpack2_h16(xws[i], bf, x0, x1);
    x0 = (kb & 1u) ? x0 : 0.f;
    x1 = (kb & 2u) ? x1 : 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const float vj = vs[t][j];
      acc0[j] = fmaf(vj, x0, acc0[j]);
      acc1[j] = fmaf(vj, x1, acc1[j]);
    }
  }
  if (th == 1) {
#pragma unroll
    for (int j = 0; j < R; ++j) { red[fq][lane][2 * j] = acc0[j]; red[fq][lane][2 * j + 1] = acc1[j]; }
  }
  __syncthreads();
  if (th == 0) {
    float* out = scratch + (((long)c * 3 + p) * 16) * 256;
#pragma unroll
    for (int j = 0; j < R; ++j)
      *reinterpret_cast<float2*>(out + j * 256 + k0) =
          make_float2(acc0[j] + red[fq][lane][2 * j], acc1[j] + red[fq][lane][2 * j + 1]);
  }
}
// grid (3r, 8): block (pj, kb) reduces feature slice [32 kb, 32 kb + 32) of rank row pj; its 256 threads are 32 features x
// 8 chunk slices, so every thread sums nchunks / 8 partials with independent loads (one or two L2 round trips instead of
// a dependent chain of nchunks), then the 8 slices are added in fixed order (deterministic).
__global__ void __launch_bounds__(256) lora_wgrad_a_drop_reduce_kernel(const float* __restrict__ scratch, int nchunks,
                                                                       float* __restrict__ part_a, int S, int r,
                                                                       float inv_keep) {
  __shared__ float red[8][32];
  pdl_wait();
  pdl_launch();
  const int pj = blockIdx.x, k = blockIdx.y * 32 + (threadIdx.x & 31), sl = threadIdx.x >> 5;
  const int p = pj / r, j = pj - p * r;
  float s = 0.f;
#pragma unroll 4
  for (int c = sl; c < nchunks; c += 8) s += scratch[(((long)c * 3 + p) * 16 + j) * 256 + k];
  red[sl][threadIdx.x & 31] = s;
  __syncthreads();
  if (sl == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x & 31];
    part_a[(long)k * 64 + pj] = t * inv_keep;
    for (int sp = 1; sp < S; ++sp) part_a[((long)sp * 256 + k) * 64 + pj] = 0.f;
  }
}

__global__ void lora_seed_bump_kernel(unsigned long long* seed) { seed[0] += 0x632BE59BD9B4E019ull; }

int launch_lora_seed_bump(unsigned long long* seed, cudaStream_t st) {
  lora_seed_bump_kernel<<<1, 1, 0, st>>>(seed);
  LAUNCH_RET();
}
// warps take two tokens per pass (one for r = 16 in the forward): size the grid so that a pass covers M when it can
static unsigned drop_grid(long M, int tok_unused) {
  const int tok = 1; (void)tok_unused; long g = (M + 8 * tok - 1) / (8 * tok); return (unsigned)(g < 148 * 3 ? g : 148 * 3); }   // three 256-thread CTAs per SM are resident (80 registers): one wave. Four (64 registers, ~90 B of spills) measured +0.5 ms per step

int launch_ln_lora_drop_fwd(const float* h, const float* gamma, const float* beta, const void* acat16, void* x16, void* ud16,
                            uint32_t* bits, long M, int r, int bf16, const LoraDropSpec& d, cudaStream_t st) {
  if (M <= 0) return -(int)cudaErrorInvalidValue;
  const uint16_t* a = reinterpret_cast<const uint16_t*>(acat16);
  uint16_t* x = reinterpret_cast<uint16_t*>(x16);
  uint32_t* u = reinterpret_cast<uint32_t*>(ud16);
  if (r == 4) launch_pdl(ln_lora_drop_fwd_kernel<4>, drop_grid(M, 2), 256, 0, st, h, gamma, beta, a, x, u, bits, M, bf16, d);
  else if (r == 8) launch_pdl(ln_lora_drop_fwd_kernel<8>, drop_grid(M, 2), 256, 0, st, h, gamma, beta, a, x, u, bits, M, bf16, d);
  else if (r == 16) launch_pdl(ln_lora_drop_fwd_kernel<16>, drop_grid(M, 1), 256, 0, st, h, gamma, beta, a, x, u, bits, M, bf16, d);
  else return -(int)cudaErrorInvalidValue;
  LAUNCH_RET();
}
int launch_ln_lora_drop_bwd(const void* dxe16, const void* acat16, const uint32_t* bits, const float* h_in, const float* gamma,
                            const float* dres, float* dh, void* dh16, long M, int r, float scaling, float inv_keep, int bf16,
                            cudaStream_t st) {
  if (M <= 0) return -(int)cudaErrorInvalidValue;
  const uint16_t* e = reinterpret_cast<const uint16_t*>(dxe16);
  const uint16_t* a = reinterpret_cast<const uint16_t*>(acat16);
  uint16_t* o = reinterpret_cast<uint16_t*>(dh16);
  const float sc = scaling * inv_keep;
  if (r == 4) launch_pdl(ln_lora_drop_bwd_kernel<4>, drop_grid(M, 2), 256, 0, st, e, a, bits, h_in, gamma, dres, dh, o, M, sc, bf16);
  else if (r == 8) launch_pdl(ln_lora_drop_bwd_kernel<8>, drop_grid(M, 2), 256, 0, st, e, a, bits, h_in, gamma, dres, dh, o, M, sc, bf16);
  else if (r == 16) launch_pdl(ln_lora_drop_bwd_kernel<16>, drop_grid(M, 2), 256, 0, st, e, a, bits, h_in, gamma, dres, dh, o, M, sc, bf16);
  else return -(int)cudaErrorInvalidValue;
  LAUNCH_RET();
}
long lora_wgrad_a_dropout_scratch_floats(long M) { return ((M + kWgaChunk - 1) / kWgaChunk) * 3L * 16 * 256; }
int launch_lora_wgrad_a_drop(const void* x16, const void* v16, long ld_v, const uint32_t* bits, float* scratch, float* part_a,
                             int S, long M, int r, float inv_keep, int bf16, cudaStream_t st) {
  const int nchunks = (int)((M + kWgaChunk - 1) / kWgaChunk);
  const uint16_t* xx = reinterpret_cast<const uint16_t*>(x16);
  const uint16_t* vv = reinterpret_cast<const uint16_t*>(v16);
  if (r == 4) launch_pdl(lora_wgrad_a_drop_kernel<4>, dim3(nchunks, 3), 256, 0, st, xx, vv, ld_v, bits, scratch, M, bf16);
  else if (r == 8) launch_pdl(lora_wgrad_a_drop_kernel<8>, dim3(nchunks, 3), 256, 0, st, xx, vv, ld_v, bits, scratch, M, bf16);
  else if (r == 16) launch_pdl(lora_wgrad_a_drop_kernel<16>, dim3(nchunks, 3), 256, 0, st, xx, vv, ld_v, bits, scratch, M, bf16);
  else return -(int)cudaErrorInvalidValue;
  launch_pdl(lora_wgrad_a_drop_reduce_kernel, dim3(3 * r, 8), 256, 0, st, (const float*)scratch, nchunks, part_a, S, r, inv_keep);
  LAUNCH_RET();
}

}  // namespace cvflow
