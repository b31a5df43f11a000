// Micro-benchmark for the round-2 kernel idea in DESIGN.md section 5: the FeaSt aggregation itself on tcgen05.
// Per group of 14 nodes: D[128 (node, head) rows x 64 channels] = A[128 x K] . B[K x 64] with K = 192 (the group's edges),
// A = block-diagonal soft assignments, B = gathered rows, both bf16 split hi / lo -> 3 passes (hi.hi, hi.lo, lo.hi):
// 36 tcgen05.mma (M=128, N=64, K=16) per group.  This measures ONLY the tensor-pipe side - how many clocks one SM needs per
// group with the operands already resident in shared memory (contents are irrelevant for timing) - i.e. the floor that the
// gather / assignment / drain stages of a real kernel would have to hide under.  Modes: 0 = groups issued back to back, one
// commit at the end (pure issue rate); 1 = commit + wait after every group (latency of a non-overlapped group);
// 2 = two accumulators, wait for group g-1 while group g runs (what a pipelined kernel would do).
// Also runs N = 128 and N = 256 for comparison (is a 64-wide tile penalised?).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../geobi_gnn_b200/csrc -I../../include -o tc_agg_bench tc_agg_bench.cu && ./tc_agg_bench
// NOT YET RUN (written when the round's GPU budget was spent); compiles for sm_100a, SASS holds the UTCMMA instructions.
#include <cstdio>
#include <cuda_runtime.h>

#include "tc.cuh"

using namespace geobi::tc;

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) agg_mma_kernel(int groups, long long* clocks) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  constexpr int KB = 3;                                   // K = 192 = 3 swizzle rows of 64 bf16
  constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128;   // per K block, per hi / lo plane
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + KB * A_BYTES;
  unsigned char* b_hi = a_lo + KB * A_BYTES;
  unsigned char* b_lo = b_hi + KB * B_BYTES;
  for (int i = threadIdx.x; i < (2 * KB * (A_BYTES + B_BYTES)) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = make_idesc(128, N);
  if (threadIdx.x < 32 && elect_one()) {
    const uint64_t ah = make_desc(smem_u32(a_hi)), al = make_desc(smem_u32(a_lo));
    const uint64_t bh = make_desc(smem_u32(b_hi)), bl = make_desc(smem_u32(b_lo));
    uint32_t phase[2] = {0, 0};
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const int buf = (MODE == 2) ? (g & 1) : 0;
      const uint32_t acc = tmem + buf * N;                // accumulators side by side in TMEM columns
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint64_t oa = (uint64_t)(kb * A_BYTES) >> 4, ob = (uint64_t)(kb * B_BYTES) >> 4;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) mma_f16(acc, ah + oa + 2 * k16, bh + ob + 2 * k16, idesc, (kb | k16) ? 1u : 0u);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) mma_f16(acc, ah + oa + 2 * k16, bl + ob + 2 * k16, idesc, 1u);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) mma_f16(acc, al + oa + 2 * k16, bh + ob + 2 * k16, idesc, 1u);
      }
      if (MODE == 1) {
        mma_commit(&bar[0]);
        mbar_wait(&bar[0], phase[0]);
        phase[0] ^= 1;
      } else if (MODE == 2) {
        mma_commit(&bar[buf]);
        if (g > 0) {                                      // the previous group's accumulator is complete: a drain could start
          mbar_wait(&bar[buf ^ 1], phase[buf ^ 1]);
          phase[buf ^ 1] ^= 1;
        }
      }
    }
    if (MODE == 0) {
      mma_commit(&bar[0]);
      mbar_wait(&bar[0], 0);
    } else if (MODE == 2) {
      const int last = (groups - 1) & 1;
      mbar_wait(&bar[last], phase[last]);
    }
    const long long t1 = clock64();
    clocks[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int MODE>
void run(const char* label, int n_sm, long long* d_clocks) {
  const int groups = 4096;
  const size_t smem = 2 * 3 * (128 * 128 + N * 128) + 1024;
  cudaFuncSetAttribute(agg_mma_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  agg_mma_kernel<N, MODE><<<n_sm, 128, smem>>>(64, d_clocks);          // warm-up
  cudaEventRecord(e0);
  agg_mma_kernel<N, MODE><<<n_sm, 128, smem>>>(groups, d_clocks);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("%-28s FAILED: %s\n", label, cudaGetErrorString(err));
    return;
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[256];
  cudaMemcpy(h, d_clocks, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < n_sm; ++i) mx = h[i] > mx ? h[i] : mx;
  const double macs = (double)n_sm * groups * 36.0 * 128 * N * 16;
  printf("%-28s %8.1f clk/group (max over SMs)  %7.3f us/group  %7.1f TMAC/s  (%d SMs, %.3f ms)\n", label, (double)mx / groups,
         ms * 1e3 / groups, macs / (ms * 1e-3) * 1e-12, n_sm, ms);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int n_sm = prop.multiProcessorCount;
  long long* d_clocks;
  cudaMalloc(&d_clocks, sizeof(long long) * 256);
  printf("%s, %d SMs; one group = 36 x tcgen05.mma M=128 N=<N> K=16 (K=192, bf16 hi/lo x3 passes)\n", prop.name, n_sm);
  run<64, 0>("N=64  back to back", n_sm, d_clocks);
  run<64, 1>("N=64  commit+wait per group", n_sm, d_clocks);
  run<64, 2>("N=64  two accumulators", n_sm, d_clocks);
  run<128, 0>("N=128 back to back", n_sm, d_clocks);
  run<128, 2>("N=128 two accumulators", n_sm, d_clocks);
  printf("per launch of the 512 000-node facet layer: 512000 / 14 nodes per group / %d SMs = %.0f groups per SM\n", n_sm, 512000.0 / 14 / n_sm);
  return 0;
}
