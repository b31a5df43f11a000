// Micro-benchmark: issue cost of FFMA2 (fma.rn.f32x2) vs FFMA on sm_100a, 4 or 8 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  unsigned long long acc[18];
  float f[36];
#pragma unroll
  for (int i = 0; i < 18; ++i) acc[i] = (unsigned long long)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 36; ++i) f[i] = threadIdx.x + i;
  unsigned long long a = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s + 1.f);
  unsigned long long bb = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 18; ++i) acc[i] = ffma2(a, bb, acc[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 36; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(s), "f"(s + 1.f));
    }
  }
  long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int i = 0; i < 18; ++i) r += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
#pragma unroll
  for (int i = 0; i < 36; ++i) r += f[i];
  if (r == 12345.f) out[0] = r;
  __shared__ long long tmin, tmax;
  if (threadIdx.x == 0) { tmin = t0; tmax = t1; }
  __syncthreads();
  atomicMin(&tmin, t0);
  atomicMax(&tmax, t1);
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (float)(tmax - tmin) / iters;
}
int main() {
  float* d;
  cudaMalloc(&d, 64);
  for (int threads : {128, 512, 1024}) {
    float h[2];
    k<0><<<148, threads>>>(d, 2000, 1.0001f);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    float c2 = h[1];
    k<1><<<148, threads>>>(d, 2000, 1.0001f);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("threads/SM %4d (%d warps/SMSP): 18 FFMA2 = %.1f clk/iter (%.2f clk per FFMA2 per SMSP-warp-set), 36 FFMA = %.1f clk/iter\n", threads,
           threads / 128, c2, c2 / 18 / (threads / 128), h[1]);
  }
  return 0;
}
