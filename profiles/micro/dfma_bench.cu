// Micro-benchmark: DFMA issue rate per SM on sm_100a (16 independent fp64 accumulators per thread).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters, double s) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], s, 1.0);
  }
  long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += acc[i];
  if (r == 12345.0) out[0] = r;
  __shared__ long long tmin, tmax;
  if (threadIdx.x == 0) { tmin = t0; tmax = t1; }
  __syncthreads();
  atomicMin(&tmin, t0);
  atomicMax(&tmax, t1);
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(tmax - tmin) / iters;
}
int main() {
  double* d;
  cudaMalloc(&d, 64);
  for (int threads : {32, 128, 256, 512}) {
    double h[2];
    k<<<148, threads>>>(d, 2000, 1.0000001);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("threads/SM %4d: 16 DFMA/thread/iter = %.1f clk/iter -> %.1f DFMA lanes per clk per SM\n", threads, h[1], 16.0 * threads / h[1]);
  }
  return 0;
}
