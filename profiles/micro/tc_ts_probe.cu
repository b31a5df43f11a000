// Probe 2 for the round-2 fused FeaSt kernel: can the projection's weight operand live in TENSOR MEMORY (tcgen05.mma with
// A = [tmem]), and what do the A-collector hints / concurrent shared-memory stores do to the MMA rate?
//   (a) TS correctness: W[64 x 64] bf16 written to TMEM with tcgen05.st (two candidate lane mappings, two packings),
//       B[32 x 64] K-major SWIZZLE_128B in shared memory, D in the same / the other lane half; the host says which combination
//       reproduces W.B^T.
//   (b) throughput: TS chain; SS chain with collector::a::fill / lastuse pairs (same A, two B tiles); plain SS pairs.
//   (c) interference: SS chain while 8 other warps stream st.shared.v4.
//   (d) issue cost: the agg + proj pattern of probe 1 with precomputed 32-bit descriptor words.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../geobi_gnn_b200/csrc -I../../include -o tc_ts_probe tc_ts_probe.cu
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
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_fill(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_lastuse(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n\t"
      "tcgen05.wait::st.sync.aligned;"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// ------------------------------------------------------------------------------------------------- (a) TS correctness
// lane_mode 0: row m -> lane (m % 16) + 32 (m / 16) [half-subpartition, as D for M = 64]; 1: row m -> lane m.
// pack 0: column c holds (k = 2c in the low half, 2c+1 in the high half); 1: swapped.   d_half: lane offset of D (0 / 16).
__global__ void __launch_bounds__(128, 1) ts_kernel(const uint16_t* __restrict__ Wg, const uint16_t* __restrict__ Bg, float* __restrict__ dump,
                                                    int lane_mode, int pack, int d_half, int a_half) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < 32 * 8; idx += 128) {     // B: row n, chunk c
    const int n = idx / 8, c = idx % 8;
    *reinterpret_cast<uint4*>(sm + sw128_off(n, c)) = *reinterpret_cast<const uint4*>(Bg + n * 64 + c * 8);
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
  // zero the D region first (so that unwritten lanes read as 0), then A -> TMEM columns 0..31
  {
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = 0;
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 64, z);
    uint32_t r[32];
    int m = -1;
    if (lane_mode == 0) m = (a_half ? lane >= 16 : lane < 16) ? 16 * warp + (lane & 15) : -1;
    else m = tid < 64 ? tid : -1;
    for (int c = 0; c < 32; ++c) {
      uint32_t lo = 0, hi = 0;
      if (m >= 0) {
        lo = Wg[m * 64 + 2 * c];
        hi = Wg[m * 64 + 2 * c + 1];
      }
      r[c] = pack == 0 ? (lo | (hi << 16)) : (hi | (lo << 16));
    }
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16), r);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = mk_idesc(64, 32, 0, 0);
    const uint64_t bd = mk_desc(smem_u32(sm), 1, 64, 2);
    for (int k16 = 0; k16 < 4; ++k16) mma_ts(tmem + 64 + ((uint32_t)d_half << 16), tmem + ((uint32_t)a_half << 16) + 8 * k16, bd + 2 * k16, idesc, k16 ? 1u : 0u);
    mma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 64, v);
  for (int i = 0; i < 32; ++i) dump[(size_t)tid * 32 + i] = v[i];
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

static void run_ts(int lane_mode, int pack, int d_half, int a_half = 0) {
  std::vector<uint16_t> W(64 * 64), B(32 * 64);
  srand(4321);
  for (auto& v : W) v = f2bf((float)(rand() % 17 - 8) / 4.0f);
  for (auto& v : B) v = f2bf((float)(rand() % 13 - 6) / 8.0f);
  uint16_t *dW, *dB;
  float* dD;
  cudaMalloc(&dW, W.size() * 2);
  cudaMalloc(&dB, B.size() * 2);
  cudaMalloc(&dD, 128 * 32 * 4);
  cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  ts_kernel<<<1, 128, 8192>>>(dW, dB, dD, lane_mode, pack, d_half, a_half);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("ts lane_mode %d pack %d d_half %d a_half %d: FAILED %s\n", lane_mode, pack, d_half, a_half, cudaGetErrorString(err));
    exit(1);
  }
  std::vector<float> D(128 * 32);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double e = 0;
  int nz = 0;
  for (auto v : D) nz += v != 0.f;
  for (int m = 0; m < 64; ++m)
    for (int n = 0; n < 32; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)bf2f(W[m * 64 + k]) * bf2f(B[n * 64 + k]);
      const int lane = (m % 16) + 32 * (m / 16) + d_half;
      e = fmax(e, fabs(D[lane * 32 + n] - ref));
    }
  printf("ts  A-lanes %s (offset %2d)  pack %s  D lane offset %2d: max|err| %.3g   (%d non-zero cells of 4096)\n",
         lane_mode == 0 ? "16/quadrant" : "0..63      ", a_half, pack == 0 ? "low=even k" : "low=odd k ", d_half, e, nz);
  cudaFree(dW);
  cudaFree(dB);
  cudaFree(dD);
}

// ------------------------------------------------------------------------------------------------- (b), (c) throughput
// mode 0: TS chain (A in TMEM, 8 columns per K step, cycling over 32 K steps; B cycling over 8 tiles)
// mode 1: SS pairs with collector::a::fill / lastuse (same A, two different B tiles)
// mode 2: SS pairs without hints (reference)
// mode 3: mode 2 + 8 warps streaming st.shared.v4 into a 32 KB scratch region
__global__ void __launch_bounds__(128 + 256, 1) chain2_kernel(int n_mma, int mode, long long* clocks, long long* stores) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  constexpr int A_BYTES = 64 * 128, B_BYTES = 32 * 128;   // K-major tiles (four K = 16 steps each)
  unsigned char* at = sm;                  // 4 A tiles
  unsigned char* bt = sm + 4 * A_BYTES;    // 8 B tiles
  unsigned char* scratch = bt + 8 * B_BYTES;
  for (int i = threadIdx.x; i < (4 * A_BYTES + 8 * B_BYTES) / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop = 0;
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 128) {
    uint32_t r[32];
    for (int c = 0; c < 32; ++c) r[c] = 0x3c003c00u + c;
    for (int c0 = 0; c0 < 256; c0 += 32) tmem_st32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + c0, r);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x >= 128) {
    if (mode == 3) {
      long long n = 0;
      const uint4 v = make_uint4(1, 2, 3, 4);
      uint4* p = reinterpret_cast<uint4*>(scratch) + (threadIdx.x - 128);
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i * 256] = v;      // 8 x 256 threads x 16 B = 32 KB per round, conflict-free
        n += 8;
      }
      if (threadIdx.x == 128) stores[blockIdx.x] = n * 256 * 16;
    }
  } else if (threadIdx.x < 32 && elect_one()) {
    constexpr uint32_t idesc = mk_idesc(64, 32, 0, 0);
    uint64_t ad[8], bd[8];
    for (int b = 0; b < 8; ++b) {
      ad[b] = mk_desc(smem_u32(at + (b >> 1) * A_BYTES + (b & 1) * 64), 1, 64, 2);
      bd[b] = mk_desc(smem_u32(bt + b * B_BYTES), 1, 64, 2);
    }
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
      const uint32_t d = tmem + 256 + ((i >> 3) & 1) * 64;
      if (mode == 0) {
#pragma unroll
        for (int b = 0; b < 8; ++b) mma_ts(d, tmem + ((i + b) & 31) * 8, bd[b], idesc, 1u);
      } else if (mode == 1) {
#pragma unroll
        for (int b = 0; b < 8; b += 2) {
          mma_fill(d, ad[b], bd[b], idesc, 1u);
          mma_lastuse(d, ad[b], bd[b + 1], idesc, 1u);
        }
      } else {
#pragma unroll
        for (int b = 0; b < 8; b += 2) {
          mma_f16(d, ad[b], bd[b], idesc, 1u);
          mma_f16(d, ad[b], bd[b + 1], idesc, 1u);
        }
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    clocks[blockIdx.x] = clock64() - t0;
    stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

static void run_chain2(int n_sm, int mode, long long* d_clocks, long long* d_stores) {
  const int n = 8192;
  const size_t smem = 4 * 64 * 128 + 8 * 32 * 128 + 32768 + 1024;
  cudaFuncSetAttribute(chain2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  chain2_kernel<<<n_sm, 384, smem>>>(64, mode, d_clocks, d_stores);
  chain2_kernel<<<n_sm, 384, smem>>>(n, mode, d_clocks, d_stores);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("chain2 mode %d FAILED: %s\n", mode, cudaGetErrorString(err));
    exit(1);
  }
  long long h[256], s[256];
  cudaMemcpy(h, d_clocks, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
  cudaMemcpy(s, d_stores, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < n_sm; ++i) mx = h[i] > mx ? h[i] : mx;
  const char* names[] = {"TS (A in TMEM), B 1 KB", "SS pairs, collector fill/lastuse", "SS pairs, no hint", "SS pairs + 8 warps of st.shared.v4"};
  printf("chain2 M=64 N=32 %-36s: %6.1f clk/MMA", names[mode], (double)mx / n);
  if (mode == 3) printf("   concurrent stores %.1f B/clk", (double)s[0] / (double)h[0]);
  printf("\n");
}

// ------------------------------------------------------------------------------------------------- (d) issue cost
// the per-tile pattern with 32-bit descriptor words computed by adds only (as the product kernel will do)
__global__ void __launch_bounds__(128, 1) pattern2_kernel(int tiles, int mode, long long* clocks) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[32];
  __shared__ uint32_t tmem_slot;
  constexpr int XN = 14;
  constexpr int X_BYTES = 4096, Q_BYTES = 1024;
  constexpr int W_BYTES = 9 * 8192, Z_BYTES = 2 * 9 * 4096;
  unsigned char* xs = sm;
  unsigned char* qs = xs + 16 * X_BYTES;
  unsigned char* ws = qs + 16 * Q_BYTES;
  unsigned char* zs = ws + W_BYTES;
  for (int i = threadIdx.x; i < (16 * (X_BYTES + Q_BYTES) + W_BYTES + Z_BYTES) / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 32; ++i) mbar_init(&dummy[i], 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  {
    uint32_t r[32];
    for (int c = 0; c < 32; ++c) r[c] = 0x3c003c00u + c;
    for (int c0 = 0; c0 < 288; c0 += 32) tmem_st32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + c0, r);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32 && elect_one()) {
    constexpr uint32_t ia = mk_idesc(64, 32, 1, 1), ib = mk_idesc(64, 16, 1, 1), ip = mk_idesc(64, 32, 0, 0);
    const uint64_t x0 = mk_desc(smem_u32(xs), 0, 64, 2), q0 = mk_desc(smem_u32(qs), 0, 32, 4);
    const uint64_t w0 = mk_desc(smem_u32(ws), 1, 64, 2), z0 = mk_desc(smem_u32(zs), 1, 64, 2);
    const uint32_t xhi = (uint32_t)(x0 >> 32), qhi = (uint32_t)(q0 >> 32), whi = (uint32_t)(w0 >> 32);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      if (mode & 1) {
        uint32_t xl = (uint32_t)x0, ql = (uint32_t)q0;
        int s = 0;
#pragma unroll 2
        for (int n = 0; n < 32; n += 2) {
          const uint32_t d = tmem + 352 + (uint32_t)((n >> 1) % 5) * 32;
          // node n (lane half 0) and node n+1 (lane half 16): x_hi.[q_hi|q_lo] then x_lo.q_hi
          asm volatile(
              "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db, dc;\n\t.reg .b32 t0, t1;\n\t"
              "setp.ne.b32 p, 0, 0;\n\tsetp.eq.b32 q, 0, 0;\n\t"
              "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
              "add.u32 t0, %1, 128;\n\tmov.b64 dc, {t0, %3};\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], dc, db, %6, q;\n\t}"
              ::"r"(d), "r"(xl), "r"(ql), "r"(xhi), "r"(qhi), "r"(ia), "r"(ib)
              : "memory");
          asm volatile(
              "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db, dc;\n\t.reg .b32 t0, t1;\n\t"
              "setp.ne.b32 p, 0, 0;\n\tsetp.eq.b32 q, 0, 0;\n\t"
              "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
              "add.u32 t0, %1, 128;\n\tmov.b64 dc, {t0, %3};\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], dc, db, %6, q;\n\t}"
              ::"r"(d + (16u << 16)), "r"(xl + 256), "r"(ql + 64), "r"(xhi), "r"(qhi), "r"(ia), "r"(ib)
              : "memory");
          if (mode & 8) {
            mma_commit(&dummy[n & 15]);
            mma_commit(&dummy[16 + (n & 15)]);
            mma_commit(&dummy[(n + 1) & 15]);
            mma_commit(&dummy[16 + ((n + 1) & 15)]);
          } else if (mode & 16) {
            mma_commit(&dummy[n & 15]);
            mma_commit(&dummy[(n + 1) & 15]);
          }
          s += 2;
          xl += 512; ql += 128;
          if (s >= XN) {
            s = 0;
            xl = (uint32_t)x0; ql = (uint32_t)q0;
          }
        }
      }
      if (mode & 2) {
        uint32_t wl = (uint32_t)w0, zh = (uint32_t)z0, zl = (uint32_t)z0 + (9 * 4096 >> 4);
        const uint32_t acc = tmem + 288 + (t & 1) * 32;
#pragma unroll 1
        for (int kb = 0; kb < 9; ++kb) {
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            if (mode & 4) {
              const uint64_t bh = ((uint64_t)whi << 32) | (zh + 2 * k16), bl = ((uint64_t)whi << 32) | (zl + 2 * k16);
              mma_ts(acc, tmem + (kb * 4 + k16) * 8, bh, ip, (kb | k16) ? 1u : 0u);
              mma_ts(acc, tmem + (kb * 4 + k16) * 8, bl, ip, 1u);
            } else {
              if (kb == 0 && k16 == 0) mma_f16_first(acc, wl, zh, whi, ip);
              else mma_f16_acc(acc, wl + 2 * k16, zh + 2 * k16, whi, ip);
              mma_f16_acc(acc, wl + 2 * k16, zl + 2 * k16, whi, ip);
            }
          }
          wl += 8192 >> 4; zh += 4096 >> 4; zl += 4096 >> 4;
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

static void run_pattern2(int n_sm, int mode, long long* d_clocks) {
  const int tiles = 256;
  const size_t smem = 16 * 5120 + 9 * 8192 + 2 * 9 * 4096 + 1024;
  cudaFuncSetAttribute(pattern2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pattern2_kernel<<<n_sm, 128, smem>>>(4, mode, d_clocks);
  pattern2_kernel<<<n_sm, 128, smem>>>(tiles, mode, d_clocks);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("pattern2 mode %d FAILED: %s\n", mode, cudaGetErrorString(err));
    exit(1);
  }
  long long h[256];
  cudaMemcpy(h, d_clocks, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < n_sm; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("pattern2 mode %d (1 agg, 2 proj, 3 both; 32-bit descriptor adds): %7.1f clk per 32-node tile = %5.1f clk/node\n", mode,
         (double)mx / tiles, (double)mx / tiles / 32);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int n_sm = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, n_sm);
  long long *d_clocks, *d_stores;
  cudaMalloc(&d_clocks, sizeof(long long) * 256);
  cudaMalloc(&d_stores, sizeof(long long) * 256);
  cudaMemset(d_stores, 0, sizeof(long long) * 256);
  run_pattern2(n_sm, 1, d_clocks);
  run_pattern2(n_sm, 2, d_clocks);
  run_pattern2(n_sm, 3, d_clocks);
  run_pattern2(n_sm, 6, d_clocks);
  run_pattern2(n_sm, 7, d_clocks);
  run_pattern2(n_sm, 1 | 16, d_clocks);
  run_pattern2(n_sm, 1 | 8, d_clocks);
  return 0;
  run_chain2(n_sm, 2, d_clocks, d_stores);
  run_chain2(n_sm, 1, d_clocks, d_stores);
  run_chain2(n_sm, 3, d_clocks, d_stores);
  run_chain2(n_sm, 0, d_clocks, d_stores);
  for (int lane_mode = 0; lane_mode < 2; ++lane_mode)
    for (int pack = 0; pack < 2; ++pack) run_ts(lane_mode, pack, 0);
  run_ts(0, 0, 16, 16);
  return 0;
}
