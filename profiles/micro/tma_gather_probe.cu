// Probe 3 (round 2): can TMA's tile::gather4 feed the aggregation operand of feast_tcagg?
//   (a) semantics: which tensor-map box ({64,1} or {64,4}) the instruction wants, how many bytes it counts on the mbarrier, where the
//       four rows land under SWIZZLE_128B (the MN-major SW128 operand wants slot k's 128-B row at (k>>3)*1024 + (k&7)*128 with 16-B
//       chunk c at c ^ (k&7)), what an out-of-range row index gives;
//   (b) throughput per SM: one warp issuing 16 gather4 (64 rows x 128 B = one node pair's x_hi|x_lo rows) per item into a ring of D slots,
//       all 148 SMs at once, rows drawn from a window (mesh locality) of a 512k-row table.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../geobi_gnn_b200/csrc -I../../include -o tma_gather_probe tma_gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc.cuh"

using namespace geobi::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<EncodeTiledFn>(p);
}
static bool make_map(CUtensorMap* m, const void* g, int64_t rows, int box_rows, CUtensorMapSwizzle sw) {
  const cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(g), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("  cuTensorMapEncodeTiled(box rows %d) -> %d\n", box_rows, (int)r);
  return r == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_gather4(uint32_t dst, const void* tmap, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// ---------------------------------------------------------------- (a) semantics
__global__ void sem_kernel(const __grid_constant__ CUtensorMap tm, int4 rows_a, int4 rows_b, uint32_t expect, uint32_t* dump, int* flags) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tile = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) ((uint32_t*)tile)[i] = 0xFFFFFFFFu;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar[0], expect);
    tma_gather4(smem_u32(tile), &tm, 0, rows_a.x, rows_a.y, rows_a.z, rows_a.w, smem_u32(&bar[0]));
    int ok = 0;
    for (int it = 0; it < 2000000 && !ok; ++it) ok = try_wait(smem_u32(&bar[0]), 0);
    flags[0] = ok;
    mbar_expect_tx(&bar[1], expect);
    tma_gather4(smem_u32(tile) + 512, &tm, 0, rows_b.x, rows_b.y, rows_b.z, rows_b.w, smem_u32(&bar[1]));
    ok = 0;
    for (int it = 0; it < 2000000 && !ok; ++it) ok = try_wait(smem_u32(&bar[1]), 0);
    flags[1] = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) dump[i] = ((uint32_t*)tile)[i];
}

// ---------------------------------------------------------------- (b) throughput
template <int D>
__global__ void __launch_bounds__(64) thr_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, int items, int n_rows,
                                                 int window, long long* clk_out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* ring = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[D];
  if (threadIdx.x == 0) { for (int i = 0; i < D; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const int base0 = (int)(((long long)blockIdx.x * n_rows) / gridDim.x);
  long long t0 = clock64();
  for (int it = 0; it < items; ++it) {
    const int s = it % D, k = it / D;
    if (k > 0) mbar_wait(&bar[s], (k - 1) & 1);      // the slot's previous fill has landed (nothing consumes it here)
    if (lane == 0) mbar_expect_tx(&bar[s], 8192);
    __syncwarp();
    if (lane < 16) {
      // 16 instructions: lanes 0-7 the hi rows of slots 4*lane.., lanes 8-15 the lo rows
      const int q = lane & 7;
      const int base = base0 + 2 * it;
      int r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t h = (uint32_t)(base * 31 + (q * 4 + i) * 2654435761u);
        r[i] = (base + (int)(h % (uint32_t)window) - window / 2 + n_rows) % n_rows;
      }
      tma_gather4(smem_u32(ring) + s * 8192 + (lane >> 3) * 4096 + q * 512, lane < 8 ? &tm_hi : &tm_lo, 0, r[0], r[1], r[2], r[3], smem_u32(&bar[s]));
    }
    __syncwarp();
  }
  for (int s = 0; s < D && s < items; ++s) { const int last = (items - 1 - s) / D * D + s; mbar_wait(&bar[s], (last / D) & 1); }
  long long t1 = clock64();
  if (lane == 0) clk_out[blockIdx.x] = t1 - t0;
}

template <int D>
static void run_thr(const CUtensorMap& hi, const CUtensorMap& lo, int n_rows, int window, long long* d_clk) {
  const int items = 4000, grid = 148;
  const size_t smem = (size_t)D * 8192 + 1024;
  CK(cudaFuncSetAttribute(thr_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  thr_kernel<D><<<grid, 64, smem>>>(hi, lo, items, n_rows, window, d_clk);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  thr_kernel<D><<<grid, 64, smem>>>(hi, lo, items, n_rows, window, d_clk);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> clk(grid);
  CK(cudaMemcpy(clk.data(), d_clk, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0; long long mx = 0;
  for (long long c : clk) { mean += (double)c; if (c > mx) mx = c; }
  mean /= grid;
  printf("  depth %2d window %6d: %.0f clk / item (max CTA %.0f), %.1f B/clk/SM, kernel %.3f ms -> %.0f GB/s over %d SMs\n", D, window, mean / items,
         (double)mx / items, 8192.0 * items / mean, ms, (double)grid * items * 8192 / (ms * 1e6), grid);
}

int main() {
  if (!encode_fn()) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int N = 512000;
  std::vector<uint16_t> h((size_t)N * 64);
  for (int r = 0; r < N; ++r) for (int c = 0; c < 64; ++c) h[(size_t)r * 64 + c] = (uint16_t)((r * 131 + c * 7 + 1) & 0x7FFF);
  uint16_t *d_hi, *d_lo;
  CK(cudaMalloc(&d_hi, h.size() * 2)); CK(cudaMalloc(&d_lo, h.size() * 2));
  CK(cudaMemcpy(d_hi, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_lo, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  uint32_t* d_dump; int* d_flags; long long* d_clk;
  CK(cudaMalloc(&d_dump, 2048)); CK(cudaMalloc(&d_flags, 8)); CK(cudaMalloc(&d_clk, 148 * 8));

  printf("(a) gather4 semantics\n");
  const int4 ra = {5, 1000, 7, 123456}, rb = {9, 10, N + 5, 11};
  auto semantics = [&](int box_rows) {
  bool all_ok = false;
  {
    for (uint32_t expect : {512u, 2048u, 128u}) {
      if (box_rows == 1 && expect == 2048u) continue;
      if (box_rows == 4 && expect == 128u) continue;
      CUtensorMap tm;
      if (!make_map(&tm, d_hi, N, box_rows, CU_TENSOR_MAP_SWIZZLE_128B)) continue;
      CK(cudaMemset(d_flags, 0, 8));
      CK(cudaFuncSetAttribute(sem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096));
      sem_kernel<<<1, 128, 4096>>>(tm, ra, rb, expect, d_dump, d_flags);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  box rows %d expect %u: kernel error %s\n", box_rows, expect, cudaGetErrorString(e)); exit(1); }
      int flags[2]; std::vector<uint32_t> dump(512);
      CK(cudaMemcpy(flags, d_flags, 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(dump.data(), d_dump, 2048, cudaMemcpyDeviceToHost));
      printf("  box {64,%d}, expect_tx %u: barrier completed %d %d\n", box_rows, expect, flags[0], flags[1]);
      if (expect == 512u && flags[0] && flags[1]) all_ok = true;
      // where did each row's chunks go?
      const uint16_t* s = (const uint16_t*)dump.data();
      const int rows[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
      for (int k = 0; k < 8; ++k) {
        printf("    slot %d (row %6d):", k, rows[k]);
        for (int c = 0; c < 8; ++c) {     // chunk c of the source row: find its 16-byte position in the dump
          int found = -1;
          if (rows[k] < N)
            for (int pos = 0; pos < 128 && found < 0; ++pos) {
              bool eq = true;
              for (int e2 = 0; e2 < 8 && eq; ++e2) eq = s[pos * 8 + e2] == h[(size_t)rows[k] * 64 + c * 8 + e2];
              if (eq) found = pos;
            }
          const int want = k * 8 + (c ^ (k & 7));
          printf(" %3d%s", found, found == want ? "" : "!");
        }
        // out-of-range row: what is at the expected place?
        if (rows[k] >= N) printf("   [first words at slot: %08x %08x]", dump[k * 32], dump[k * 32 + 1]);
        printf("\n");
      }
    }
  }
  return all_ok;
  };

  const bool ok1 = semantics(1);
  printf("(b) gather4 throughput, 64 rows x 128 B per item, one issuing warp per SM\n");
  CUtensorMap hi, lo;
  const int box_rows = 1;
  if (!ok1) printf("  skipped: a {64,1} box does not count 512 bytes per instruction\n");
  else if (!make_map(&hi, d_hi, N, box_rows, CU_TENSOR_MAP_SWIZZLE_128B) || !make_map(&lo, d_lo, N, box_rows, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  else for (int window : {2048, 65536}) {
    run_thr<2>(hi, lo, N, window, d_clk);
    run_thr<4>(hi, lo, N, window, d_clk);
    run_thr<8>(hi, lo, N, window, d_clk);
    run_thr<16>(hi, lo, N, window, d_clk);
  }
  printf("(a') the other box\n");
  semantics(4);
  return 0;
}
