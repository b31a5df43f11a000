// Probe for the round-2 fused FeaSt kernel (aggregation on tcgen05, one small MMA chain per node):
//   1. LAYOUT: are MN-major operands read the way the kernel will write them?
//        A = X^T  [M = channels x K = 16 neighbour slots], MN-major SWIZZLE_128B: slot k = one gathered 128-byte bf16 row
//        B = Q    [N = 32 (q_hi | q_lo, 16 heads each) x K = 16 slots], MN-major SWIZZLE_64B: slot k = one 64-byte row
//      D (M = 64) at TMEM lane offset 0 and 16 (two nodes interleaved in one column group), N = 16 sub-tile of the same B,
//      and M = 128 (C_in = 128: two 64-channel atoms LBO apart).  The kernel dumps TMEM; the host compares with a CPU product
//      and prints the error per descriptor variant.
//   2. THROUGHPUT: clocks per tcgen05.mma for the shapes the design needs (resident operands, back-to-back issue), and the
//      per-node cost of the planned agg (2 MMAs per node) + proj (72 MMAs per 32 nodes) pattern.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../geobi_gnn_b200/csrc -I../../include -o tc_layout_probe tc_layout_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "tc.cuh"

using namespace geobi::tc;

__host__ __device__ inline uint64_t mk_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t mk_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

struct LayoutParams {
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  int M;        // 64 or 128
};

// A rows: Ag[node][slot][M] bf16 (a gathered row per slot); B rows: Bg[node][slot][32] bf16.  Two "nodes" (0 -> lane offset 0,
// 1 -> lane offset 16 for M = 64).  dump[128 lanes][128 cols].
__global__ void __launch_bounds__(128, 1) layout_kernel(const uint16_t* __restrict__ Ag, const uint16_t* __restrict__ Bg, float* __restrict__ dump,
                                                        LayoutParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = p.M;
  const int A_NODE = (M / 64) * 2048;   // per node: (M/64) blocks of [16 slots x 128 B]
  unsigned char* a_t = sm;                       // 2 nodes
  unsigned char* b_t = sm + 2 * 4096;            // 2 nodes x 1024 B (1024-aligned)
  for (int i = tid; i < (2 * 4096 + 2 * 1024) / 4; i += 128) ((uint32_t*)sm)[i] = 0;
  __syncthreads();
  // A: slot k, channel chunk c (8 channels = 16 B)
  for (int idx = tid; idx < 2 * 16 * (M / 8); idx += 128) {
    const int node = idx / (16 * (M / 8)), rem = idx % (16 * (M / 8)), k = rem / (M / 8), c = rem % (M / 8);
    const int half = c / 8, cc = c % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(Ag + ((size_t)(node * 16 + k) * M + c * 8));
    *reinterpret_cast<uint4*>(a_t + node * A_NODE + half * 2048 + (k >> 3) * 1024 + (k & 7) * 128 + ((cc ^ (k & 7)) << 4)) = v;
  }
  // B: slot k, chunk c (0..3) of the 64-byte row
  for (int idx = tid; idx < 2 * 16 * 4; idx += 128) {
    const int node = idx / 64, rem = idx % 64, k = rem / 4, c = rem % 4;
    const uint4 v = *reinterpret_cast<const uint4*>(Bg + ((size_t)(node * 16 + k) * 32 + c * 8));
    *reinterpret_cast<uint4*>(b_t + node * 1024 + (k >> 3) * 512 + (k & 7) * 64 + ((c ^ ((k >> 1) & 3)) << 4)) = v;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && elect_one()) {
    const uint32_t i32 = mk_idesc(M, 32, 1, 1), i16 = mk_idesc(M, 16, 1, 1);
    for (int node = 0; node < 2; ++node) {
      const uint64_t ad = mk_desc(smem_u32(a_t + node * A_NODE), p.a_lbo, p.a_sbo, 2);
      const uint64_t bd = mk_desc(smem_u32(b_t + node * 1024), p.b_lbo, p.b_sbo, 4);
      // cols 0..31: N = 32 product; cols 32..47: N = 16 product of the same operands (q_hi part only)
      const uint32_t d = (M == 64) ? tmem + ((uint32_t)(node * 16) << 16) : tmem + node * 64;
      mma_f16(d, ad, bd, i32, 0u);
      mma_f16(d + 32, ad, bd, i16, 0u);
    }
    mma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  for (int c = 0; c < 128; c += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 32; ++i) dump[(size_t)tid * 128 + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static void run_layout(int M, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo) {
  std::vector<uint16_t> A(2 * 16 * M), B(2 * 16 * 32);
  srand(1234);
  for (auto& v : A) v = f2bf((float)(rand() % 17 - 8) / 4.0f);
  for (auto& v : B) v = f2bf((float)(rand() % 13 - 6) / 8.0f);
  uint16_t *dA, *dB;
  float* dD;
  cudaMalloc(&dA, A.size() * 2);
  cudaMalloc(&dB, B.size() * 2);
  cudaMalloc(&dD, 128 * 128 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * 128 * 4);
  LayoutParams p{a_lbo, a_sbo, b_lbo, b_sbo, M};
  cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  layout_kernel<<<1, 128, 16384>>>(dA, dB, dD, p);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("layout M=%d a(lbo %u sbo %u) b(lbo %u sbo %u): FAILED %s\n", M, a_lbo, a_sbo, b_lbo, b_sbo, cudaGetErrorString(err));
    exit(1);
  }
  std::vector<float> D(128 * 128);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double e32 = 0, e16 = 0;
  int bad_m = -1, bad_n = -1, bad_node = -1;
  for (int node = 0; node < 2; ++node)
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 16; ++k) ref += (double)bf2f(A[(node * 16 + k) * M + m]) * bf2f(B[(node * 16 + k) * 32 + n]);
        int lane, col0;
        if (M == 64) {
          lane = (m % 16) + 32 * (m / 16) + 16 * node;
          col0 = 0;
        } else {
          lane = m;
          col0 = node * 64;
        }
        const double d32 = fabs(D[lane * 128 + col0 + n] - ref);
        if (d32 > e32) {
          e32 = d32;
          bad_m = m; bad_n = n; bad_node = node;
        }
        if (n < 16) e16 = fmax(e16, fabs(D[lane * 128 + col0 + 32 + n] - ref));
      }
  printf("layout M=%3d a(lbo %3u sbo %3u) b(lbo %3u sbo %3u): max|err| N=32 %.3g  N=16 %.3g   (worst node %d m %d n %d)\n", M, a_lbo, a_sbo,
         b_lbo, b_sbo, e32, e16, bad_node, bad_m, bad_n);
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
}

// ---------------------------------------------------------------------------------------------------------------------
// throughput: NBUF distinct operand tiles, chain of n MMAs into rotating accumulators
template <int M, int N, int AMN>
__global__ void __launch_bounds__(128, 1) chain_kernel(int n_mma, long long* clocks) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int NBUF = 8;
  // MN-major: one tile per K = 16 step; K-major: tiles of [rows x 128 B] hold four K = 16 steps (+32 B each), two tiles each
  constexpr int A_BYTES = AMN ? (M * 32 < 1024 ? 1024 : M * 32) : M * 128;
  constexpr int B_BYTES = AMN ? (N * 32 < 1024 ? 1024 : N * 32) : N * 128;
  constexpr int A_TILES = AMN ? NBUF : 2, B_TILES = AMN ? NBUF : 2;
  for (int i = threadIdx.x; i < (A_TILES * A_BYTES + B_TILES * B_BYTES) / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32 && elect_one()) {
    constexpr uint32_t idesc = mk_idesc(M, N, AMN, AMN);
    uint64_t ad[NBUF], bd[NBUF];
    for (int b = 0; b < NBUF; ++b) {
      // MN-major: SW128 for A (rows of 128 B per slot), K-major: SW128 rows of 128 B per M row (only 32 B of each used per step)
      ad[b] = AMN ? mk_desc(smem_u32(sm + b * A_BYTES), M > 64 ? 128 : 0, 64, 2) : mk_desc(smem_u32(sm + (b >> 2) * A_BYTES + (b & 3) * 32), 1, 64, 2);
      bd[b] = AMN ? mk_desc(smem_u32(sm + A_TILES * A_BYTES + b * B_BYTES), 0, N >= 64 ? 64 : (N >= 32 ? 32 : 16), N >= 64 ? 2 : (N >= 32 ? 4 : 6))
                  : mk_desc(smem_u32(sm + A_TILES * A_BYTES + (b >> 2) * B_BYTES + (b & 3) * 32), 1, 64, 2);
    }
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += NBUF) {
#pragma unroll
      for (int b = 0; b < NBUF; ++b) mma_f16(tmem + ((i / NBUF) & 1) * 256, ad[b], bd[b], idesc, 1u);
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    clocks[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int M, int N, int AMN>
static void run_chain(int n_sm, long long* d_clocks) {
  const int n = 8192;
  const size_t smem = (AMN ? 8 * ((M * 32 < 1024 ? 1024 : M * 32) + (N * 32 < 1024 ? 1024 : N * 32)) : 2 * (M + N) * 128) + 1024;
  cudaFuncSetAttribute(chain_kernel<M, N, AMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  chain_kernel<M, N, AMN><<<n_sm, 128, smem>>>(64, d_clocks);
  chain_kernel<M, N, AMN><<<n_sm, 128, smem>>>(n, d_clocks);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("chain M=%d N=%d %s FAILED: %s\n", M, N, AMN ? "MN" : "K ", cudaGetErrorString(err));
    exit(1);
  }
  long long h[256];
  cudaMemcpy(h, d_clocks, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < n_sm; ++i) mx = h[i] > mx ? h[i] : mx;
  const double clk = (double)mx / n;
  printf("chain M=%3d N=%3d K=16 %s-major: %6.1f clk/MMA   operand bytes %5d -> %5.1f B/clk   MAC/clk %6.0f\n", M, N, AMN ? "MN" : "K ", clk,
         (M + N) * 32, (M + N) * 32 / clk, (double)M * N * 16 / clk);
}

// the planned per-tile pattern: 32 nodes x {MMA M=64 N=32 MN-major (x_hi . [q_hi|q_lo]), MMA M=64 N=NLO (x_lo . q_hi)} into rotating TMEM
// slots, then the projection chain of `proj` MMAs M=64 N=32 K-major; mode 1: agg only, 2: proj only, 3: both
template <int NLO>
__global__ void __launch_bounds__(128, 1) pattern_kernel(int tiles, int mode, int proj, long long* clocks) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int XN = 16;                       // node operand sets resident (ring)
  constexpr int X_BYTES = 4096, Q_BYTES = 1024;
  constexpr int W_BYTES = 9 * 8192, Z_BYTES = 2 * 9 * 4096;
  unsigned char* xs = sm;
  unsigned char* qs = xs + XN * X_BYTES;
  unsigned char* ws = qs + XN * Q_BYTES;
  unsigned char* zs = ws + W_BYTES;
  for (int i = threadIdx.x; i < (XN * (X_BYTES + Q_BYTES) + W_BYTES + Z_BYTES) / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32 && elect_one()) {
    constexpr uint32_t ia = mk_idesc(64, 32, 1, 1), ib = mk_idesc(64, NLO, 1, 1), ip = mk_idesc(64, 32, 0, 0);
    const uint64_t x0 = mk_desc(smem_u32(xs), 0, 64, 2), q0 = mk_desc(smem_u32(qs), 0, 32, 4);
    const uint64_t w0 = mk_desc(smem_u32(ws), 1, 64, 2), z0 = mk_desc(smem_u32(zs), 1, 64, 2);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      if (mode & 1) {
#pragma unroll 4
        for (int n = 0; n < 32; ++n) {
          const int s = n % XN;
          const uint32_t d = tmem + (uint32_t)((n >> 1) % 8) * 32 + ((uint32_t)((n & 1) * 16) << 16);
          mma_f16(d, x0 + (uint64_t)(s * (X_BYTES >> 4)), q0 + (uint64_t)(s * (Q_BYTES >> 4)), ia, 0u);
          mma_f16(d, x0 + (uint64_t)(s * (X_BYTES >> 4) + (2048 >> 4)), q0 + (uint64_t)(s * (Q_BYTES >> 4)), ib, 1u);
        }
      }
      if (mode & 2) {
        // W: 9 K blocks of [64 rows x 128 B]; Z: two planes of 9 K blocks of [32 rows x 128 B]
        for (int i = 0; i < proj; ++i) {
          const int kb = (i >> 2) % 9, k16 = i & 3, plane = (i / 36) & 1;
          mma_f16(tmem + 256 + (t & 1) * 32, w0 + (uint64_t)(kb * (8192 >> 4) + 2 * k16), z0 + (uint64_t)((plane * 9 + kb) * (4096 >> 4) + 2 * k16), ip,
                  i ? 1u : 0u);
        }
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    clocks[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int NLO>
static void run_pattern(int n_sm, int mode, int proj, long long* d_clocks) {
  const int tiles = 256;
  const size_t smem = 16 * 5120 + 9 * 8192 + 2 * 9 * 4096 + 1024;
  cudaFuncSetAttribute(pattern_kernel<NLO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pattern_kernel<NLO><<<n_sm, 128, smem>>>(4, mode, proj, d_clocks);
  pattern_kernel<NLO><<<n_sm, 128, smem>>>(tiles, mode, proj, d_clocks);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("pattern mode %d FAILED: %s\n", mode, cudaGetErrorString(err));
    exit(1);
  }
  long long h[256];
  cudaMemcpy(h, d_clocks, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < n_sm; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("pattern NLO=%2d mode %d (1 agg, 2 proj, 3 both) proj=%3d: %7.1f clk per 32-node tile = %5.1f clk/node\n", NLO, mode, proj, (double)mx / tiles,
         (double)mx / tiles / 32);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int n_sm = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, n_sm);
  // expected: A SBO = 1024 B (64), LBO unused for M = 64 / 2048 B (128) for M = 128; B (SW64) SBO = 512 B (32)
  run_layout(64, 0, 64, 0, 32);
  run_layout(64, 1, 64, 1, 32);
  run_layout(64, 64, 64, 32, 32);
  run_layout(64, 128, 64, 64, 32);
  run_layout(64, 64, 128, 32, 64);    // a wrong one on purpose: must show an error
  run_layout(128, 128, 64, 0, 32);
  run_layout(128, 64, 128, 0, 32);    // roles swapped: which of the two is right for M = 128?
  long long* d_clocks;
  cudaMalloc(&d_clocks, sizeof(long long) * 256);
  run_chain<64, 16, 1>(n_sm, d_clocks);
  run_chain<64, 32, 1>(n_sm, d_clocks);
  run_chain<64, 64, 1>(n_sm, d_clocks);
  run_chain<128, 16, 1>(n_sm, d_clocks);
  run_chain<128, 32, 1>(n_sm, d_clocks);
  run_chain<128, 64, 1>(n_sm, d_clocks);
  run_chain<64, 16, 0>(n_sm, d_clocks);
  run_chain<64, 32, 0>(n_sm, d_clocks);
  run_chain<64, 64, 0>(n_sm, d_clocks);
  run_chain<128, 32, 0>(n_sm, d_clocks);
  run_chain<128, 64, 0>(n_sm, d_clocks);
  run_chain<128, 128, 0>(n_sm, d_clocks);
  run_chain<128, 256, 0>(n_sm, d_clocks);
  run_pattern<16>(n_sm, 1, 72, d_clocks);
  run_pattern<32>(n_sm, 1, 72, d_clocks);
  run_pattern<16>(n_sm, 2, 72, d_clocks);
  run_pattern<16>(n_sm, 2, 108, d_clocks);
  run_pattern<16>(n_sm, 3, 72, d_clocks);
  run_pattern<32>(n_sm, 3, 72, d_clocks);
  run_pattern<16>(1, 3, 72, d_clocks);
  return 0;
}
