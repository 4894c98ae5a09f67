// Row helpers shared by the warp-per-token kernels (norm.cu, lora_dropout.cu): one lane owns 8 contiguous
// channels of a 256-wide row (16-byte accesses), a warp covers the row.
#pragma once
#include "common.cuh"

namespace cvflow {

__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8_f32(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void unpack8_h16(const uint4& u, int bf, float (&v)[8]) {
  unpack2_h16(u.x, bf, v[0], v[1]);
  unpack2_h16(u.y, bf, v[2], v[3]);
  unpack2_h16(u.z, bf, v[4], v[5]);
  unpack2_h16(u.w, bf, v[6], v[7]);
}
__device__ __forceinline__ void load8_h16(const uint16_t* p, int bf, float (&v)[8]) {
  unpack8_h16(*reinterpret_cast<const uint4*>(p), bf, v);
}
__device__ __forceinline__ void store8_h16(uint16_t* p, int bf, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = pack8_h16(v, bf);
}

// LayerNorm statistics of a 256-wide row held as 8 values per lane: x is centred in place, rstd returned.
__device__ __forceinline__ float row_center_rstd(float (&x)[8]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] -= mean; v += x[i] * x[i]; }
  return rsqrtf(warp_sum(v) * (1.f / 256.f) + 1e-5f);
}

}  // namespace cvflow
