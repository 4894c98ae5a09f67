// Micro-benchmarks behind the attention kernel design (run on the B200 box):
//  (1) MUFU.EX2 throughput per SM vs. resident warps per scheduler, (2) tcgen05.ld 32x32b.x32 + wait::ld round trip.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_tmem mufu_tmem.cu && ./mufu_tmem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void mufu_kernel(float* out, long long* cycles, int iters) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// out-of-place ex2 on normal inputs (x stays in [-8, 0]): no special-value path
__global__ void mufu2_kernel(float* out, long long* cycles, int iters) {
  float x[32], y[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { x[j] = -0.01f * ((threadIdx.x & 31) + j); y[j] = 0.f; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float r;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[j]));
      y[j] += r;
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += y[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// MUFU throughput with loop-varying (non-CSE-able) inputs: op 0 ex2, 1 tanh, 2 rcp
template <int OP>
__global__ void mufu3_kernel(float* out, long long* cycles, int iters) {
  float x[32], y[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { x[j] = -0.01f * ((threadIdx.x & 31) + j) - 0.5f; y[j] = 0.f; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float r;
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[j]));
      if (OP == 1) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x[j]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[j]));
      y[j] += r;
      x[j] += 0.001f;
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += y[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void cvt_kernel(float* out, long long* cycles, int iters) {
  float x[32];
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      uint32_t r;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[j]), "f"(x[j + 1]));
      acc ^= r;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
// the softmax chunk body without TMEM: 16 FFMA2 + 32 EX2 + 16 FADD2 + 16 cvt
__global__ void chunk_kernel(float* out, long long* cycles, int iters) {
  float x[32];
  uint32_t acc = 0;
  float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float p[32];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a0 = fmaf(x[j], 0.18f, -1.f), a1 = fmaf(x[j + 1], 0.18f, -1.f);
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[j]) : "f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[j + 1]) : "f"(a1));
      rs0 += p[j]; rs1 += p[j + 1];
    }
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      uint32_t r;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(p[j]), "f"(p[j + 1]));
      acc ^= r;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc) + rs0 + rs1;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void tmem_ld_kernel(float* out, long long* cycles, int iters) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + (((threadIdx.x >> 5) & 3) * 32 << 16);
  uint32_t r[32];
  float acc = 0.f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + (uint32_t)((i & 7) * 32))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
  }
  long long t1 = clock64();
  // tcgen05.st 32x32b.x16 back to back, one wait::st at the end of each group of 8
  uint32_t v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = r[j];
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(tmem + (uint32_t)((i & 7) * 16)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
                 "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    if ((i & 7) == 7) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  long long t3 = clock64();
  // interleaved ld / st as in the softmax pass (ld chunk, wait, st packed chunk)
  long long t4 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + (uint32_t)(((i & 3) + 4) * 32))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(tmem + (uint32_t)((i & 7) * 16)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
                 "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  long long t5 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t3 - t2; cycles[2] = t5 - t4; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(slot) : "memory");
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024 * 8);
  const int iters = 200;
  for (int warps_per_sched = 1; warps_per_sched <= 4; ++warps_per_sched) {
    const int threads = 128 * warps_per_sched;
    mufu_kernel<<<1, threads>>>(out, cyc, iters);
    mufu_kernel<<<1, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = (double)threads * iters * 32;
    printf("MUFU.EX2: %d warp(s)/scheduler: %.2f ex2 per clk per SM (%.1f clk per warp instr per scheduler)\n", warps_per_sched,
           ops / c, (double)c / (iters * 32.0 * warps_per_sched));
  }
  for (int w = 1; w <= 4; w *= 2) {
    mufu2_kernel<<<1, 128 * w>>>(out, cyc, iters); mufu2_kernel<<<1, 128 * w>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.EX2 (normal inputs, + FADD each): %d warp(s)/scheduler: %.2f ex2 per clk per SM\n", w, (double)128 * w * iters * 32 / c);
  }
  for (int w = 1; w <= 4; w *= 2) {
    long long c[3];
    mufu3_kernel<0><<<1, 128 * w>>>(out, cyc, iters); mufu3_kernel<0><<<1, 128 * w>>>(out, cyc, iters);
    cudaDeviceSynchronize(); cudaMemcpy(&c[0], cyc, 8, cudaMemcpyDeviceToHost);
    mufu3_kernel<1><<<1, 128 * w>>>(out, cyc, iters); mufu3_kernel<1><<<1, 128 * w>>>(out, cyc, iters);
    cudaDeviceSynchronize(); cudaMemcpy(&c[1], cyc, 8, cudaMemcpyDeviceToHost);
    mufu3_kernel<2><<<1, 128 * w>>>(out, cyc, iters); mufu3_kernel<2><<<1, 128 * w>>>(out, cyc, iters);
    cudaDeviceSynchronize(); cudaMemcpy(&c[2], cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = 128.0 * w * iters * 32;
    printf("MUFU varying inputs, %d warp(s)/scheduler: ex2 %.1f, tanh %.1f, rcp %.1f per clk per SM\n", w, ops / c[0], ops / c[1], ops / c[2]);
  }
  for (int w = 1; w <= 2; ++w) {
    cvt_kernel<<<1, 128 * w>>>(out, cyc, iters); cvt_kernel<<<1, 128 * w>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("cvt.rn.bf16x2.f32: %d warp(s)/scheduler: %.1f clk per warp instr per scheduler\n", w, (double)c / (iters * 16.0 * w));
    chunk_kernel<<<1, 128 * w>>>(out, cyc, iters); chunk_kernel<<<1, 128 * w>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("softmax chunk body (32 ex2 + 16 cvt + fma/add): %d warp(s)/scheduler: %.1f clk per chunk per warp\n", w, (double)c / (iters * 1.0 * w));
  }
  for (int warps = 4; warps <= 8; warps += 4) {
    tmem_ld_kernel<<<1, 32 * warps>>>(out, cyc, iters);
    tmem_ld_kernel<<<1, 32 * warps>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c[3]; cudaMemcpy(c, cyc, 24, cudaMemcpyDeviceToHost);
    printf("tcgen05.ld 32x32b.x32 + wait::ld, %d warps: %.1f clk per round trip; st.x16: %.1f clk each; ld+wait+st: %.1f clk\n", warps,
           (double)c[0] / iters, (double)c[1] / iters, (double)c[2] / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
