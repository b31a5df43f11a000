// Dense half of the FeaStConv backward on the tcgen05 tensor cores (training step, /root/reference/code/train_dual.py:199-218;
// the layer is torch_geometric's FeaStConv, constructed at network.py:258-268 and called at :271-299).
//
// With g = dL/d(pre-activation) [N, C_out], Z = the forward's aggregate [N, 9 C_in] and W_flat [C_out, 9 C_in]:
//     dbias = sum_n g                       (bwd_prep_g_kernel, which also writes g split into bf16 hi | lo planes)
//     dZ    = g . W_flat                    (tc_gemm_tma_kernel of feast_tc.cu: A = the g planes, B = W_flat^T in row chunks)
//     dW    = g^T . Z                       (dw_splitk_kernel, below: the reduction runs over the NODES)
//     dx, dP, dc from dZ                    (feast_bwd_edges kernels, feast_bwd.cu)
//     dx   += dP . U,   dU = dP^T . x       (bwd_dpu_kernel: K = 9, CUDA cores, one pass over x)
// Every product with a long reduction or a wide output runs as three bf16 passes on split operands (hi.hi + hi.lo + lo.hi,
// fp32 accumulation in TMEM: fp32-grade results), like the forward projections.
//
// dw_splitk_kernel.  D[o, k] = sum_n g[n, o] Z[n, k]: both operands are stored node-major, i.e. the reduction index is the SLOW
// one, so both are MN-major UMMA operands: a TMA box of 32 nodes x 64 columns (bf16, SWIZZLE_128B) of the row-major planes IS the
// canonical MN-major SWIZZLE_128B tile (128-byte rows = 64 M/N elements of one node, 8 nodes per 1 KB atom, SBO = 1 KB between
// atoms along K).  A CTA owns a node range (split K) and a slice of <= 6 column chunks of 64; per 16 nodes it issues, for every
// (64-row half of C_out, chunk), three M64 N64 K16 MMAs into that pair's accumulator - rows 0-63 in TMEM lanes 0-15 of every
// quadrant, rows 64-127 in lanes 16-31 of the same columns.  Partial sums go to [split][C_out][kpad] and a second kernel adds them
// in a fixed order (deterministic) while permuting to lin.weight's [9 C_out, C_in] layout.
#include <cuda.h>
#include <stdlib.h>

#include "tc.cuh"

namespace geobi {
namespace tc {
int gemm_dispatch(const __nv_bfloat16* A, int64_t a_plane, int64_t M, int kpad, const __nv_bfloat16* Bq, int N, const float* bias, float slope,
                  float* out, int64_t ldo, int passes, cudaStream_t st);                                       // feast_tc.cu
bool make_tmap(CUtensorMap* m, const __nv_bfloat16* g, int64_t rows, int kpad, int box_rows);                  // feast_tc.cu
}  // namespace tc
int feast_project_and_aggregate(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                const int32_t* row_map, int64_t n_src, const float* U, const float* c, double* P, void* Z, int64_t ldz,
                                int out_mode, cudaStream_t st);                                                  // feast.cu
bool feast_aggregate_fills_padding(int c_in, int64_t ldx, int64_t ldz, const float* x, bool row_map);           // feast.cu
int feast_bwd_edges_launch(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr, const double* P,
                           const float* c, const float* dZ, int64_t lddz, float* dx, int64_t lddx, float* dP, float* dc,
                           cudaStream_t st);                                                                     // feast_bwd.cu

namespace bwd {
using namespace tc;

// ------------------------------------------------------------------------------ g = go * act'(out) -> bf16 planes, dbias
// One thread = 4 adjacent columns of a row; a thread keeps its column group over all its rows, so the bias gradient is summed in
// registers and leaves the CTA as one atomic per column.
__global__ void __launch_bounds__(256) bwd_prep_g_kernel(const float* __restrict__ go, int64_t ldg, const float* __restrict__ out, int64_t ldo,
                                                         int64_t N, int c_out, int gpad, float slope, __nv_bfloat16* __restrict__ G,
                                                         float* __restrict__ dbias) {
  __shared__ float4 part[256];
  const int cg = gpad >> 2, rpb = 256 / cg;
  const int tid = threadIdx.x, q = tid % cg, r0 = tid / cg, c0 = 4 * q;
  const bool live = c0 < c_out;
  const int64_t plane = N * (int64_t)gpad;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t n = (int64_t)blockIdx.x * rpb + r0; n < N; n += (int64_t)gridDim.x * rpb) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      v = *reinterpret_cast<const float4*>(go + n * ldg + c0);
      if (slope != 1.0f) {
        const float4 y = *reinterpret_cast<const float4*>(out + n * ldo + c0);      // leaky_relu keeps the sign of its input
        v.x = y.x > 0.f ? v.x : v.x * slope;
        v.y = y.y > 0.f ? v.y : v.y * slope;
        v.z = y.z > 0.f ? v.z : v.z * slope;
        v.w = y.w > 0.f ? v.w : v.w * slope;
      }
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    uint2 hi, lo;
    split_bf16x4(v, hi, lo);
    *reinterpret_cast<uint2*>(G + n * gpad + c0) = hi;
    *reinterpret_cast<uint2*>(G + plane + n * gpad + c0) = lo;
  }
  part[tid] = acc;
  __syncthreads();
  if (tid < cg && live) {
    float4 s = part[tid];
    for (int r = 1; r < rpb; ++r) {
      const float4 t = part[r * cg + tid];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    atomicAdd(dbias + c0, s.x);
    atomicAdd(dbias + c0 + 1, s.y);
    atomicAdd(dbias + c0 + 2, s.z);
    atomicAdd(dbias + c0 + 3, s.w);
  }
}

// ------------------------------------------------------------------------------ W_flat^T as the B operand of dZ = g . W_flat
// Rows k = h*C_in + c of W_flat^T (zero rows up to kpad), columns o (zero up to gpad), cut into row chunks of 256 / 128 / 64
// (the widths tc_gemm_tma_kernel is instantiated for); chunk j occupies [2 * r0_j * gpad, ...) as its hi plane followed by its lo plane.
struct Chunks {
  int n;
  int r0[8];
  int nt[8];
};
static Chunks chunks_of(int kpad) {
  Chunks c{};
  int r = 0;
  while (r < kpad && c.n < 8) {
    const int left = kpad - r;
    const int nt = left >= 256 ? 256 : (left >= 128 ? 128 : 64);
    c.r0[c.n] = r;
    c.nt[c.n] = nt;
    ++c.n;
    r += nt;
  }
  return c;
}
__global__ void bwd_prep_wt_kernel(const float* __restrict__ W, int c_in, int c_out, int kpad, int gpad, Chunks ch, __nv_bfloat16* __restrict__ Wt) {
  const int total = kpad * gpad, K = H * c_in;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int k = t / gpad, o = t - k * gpad;
    float v = 0.f;
    if (k < K && o < c_out) {
      const int h = k / c_in, c = k - h * c_in;
      v = W[(int64_t)(h * c_out + o) * c_in + c];
    }
    int j = 0;
    while (j + 1 < ch.n && k >= ch.r0[j + 1]) ++j;
    __nv_bfloat16* base = Wt + (int64_t)2 * ch.r0[j] * gpad;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    base[(k - ch.r0[j]) * gpad + o] = hi;
    base[(ch.nt[j] + k - ch.r0[j]) * gpad + o] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// ------------------------------------------------------------------------------ dW = g^T . Z, split over the nodes
constexpr int KB = 32;                 // nodes per pipeline stage (four 8-node swizzle atoms)
constexpr int TILE_BYTES = KB * 128;   // one operand tile: KB nodes x 64 bf16
constexpr uint32_t HALF = 16u << 16;   // TMEM lane offset of the second 64-row half

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {   // kind::f16, D = f32, A = B = bf16, both MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// high word of the shared-memory matrix descriptor: SBO = 1024 B (next 8-node atom) | version 1 | SWIZZLE_128B
constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);

__device__ __forceinline__ void mma_mn(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(acc), "r"(idesc_mn(64, 64))
      : "memory");
}

// grid = (splits, slices).  mh = 64-row halves of C_out (1 or 2), nc = column chunks per slice, ct = kpad / 64.
// Stage layout: [G_hi half 0..mh-1 | G_lo half 0..mh-1 | Z_hi chunk 0..nc-1 | Z_lo chunk 0..nc-1], TILE_BYTES each.
__global__ void __launch_bounds__(128) dw_splitk_kernel(const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
                                                        const __grid_constant__ CUtensorMap tm_z_hi, const __grid_constant__ CUtensorMap tm_z_lo,
                                                        int64_t N, int64_t nodes_per_split, int mh, int nc, int ct, int stages, int c_out,
                                                        int kpad, uint32_t tmem_cols, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[8];
  __shared__ __align__(8) uint64_t empty[8];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c_first = blockIdx.y * nc;
  const int ncs = min(nc, ct - c_first);                 // chunks this slice really has
  const int64_t n0 = (int64_t)blockIdx.x * nodes_per_split;
  const int64_t n1 = min(N, n0 + nodes_per_split);
  const int nkb = (int)((n1 - n0 + KB - 1) / KB);        // >= 1 by construction of the grid
  const uint32_t stage_bytes = (uint32_t)(2 * mh + 2 * nc) * TILE_BYTES;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {                                   // ---- producer: TMA boxes of KB nodes x 64 columns
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        if (kb >= stages) mbar_wait(&empty[s], (uint32_t)(((kb / stages) - 1) & 1));
        const uint32_t st = base + (uint32_t)s * stage_bytes;
        const int node = (int)(n0 + (int64_t)kb * KB);
        mbar_expect_tx(&full[s], (uint32_t)(2 * mh + 2 * ncs) * TILE_BYTES);
        for (int m = 0; m < mh; ++m) {
          tma_load_2d(st + (uint32_t)m * TILE_BYTES, &tm_g_hi, m * 64, node, &full[s]);
          tma_load_2d(st + (uint32_t)(mh + m) * TILE_BYTES, &tm_g_lo, m * 64, node, &full[s]);
        }
        for (int c = 0; c < ncs; ++c) {
          tma_load_2d(st + (uint32_t)(2 * mh + c) * TILE_BYTES, &tm_z_hi, (c_first + c) * 64, node, &full[s]);
          tma_load_2d(st + (uint32_t)(2 * mh + nc + c) * TILE_BYTES, &tm_z_lo, (c_first + c) * 64, node, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {                                   // ---- MMA issue
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        mbar_wait(&full[s], (uint32_t)((kb / stages) & 1));
        tc_fence_after();
        const uint32_t st16 = ((base + (uint32_t)s * stage_bytes) & 0x3FFFFu) >> 4;
#pragma unroll
        for (int k16 = 0; k16 < KB / 16; ++k16) {
          const uint32_t koff = (uint32_t)k16 * (2048u >> 4);           // two 8-node atoms per K = 16 step
          const uint32_t acc = (kb | k16) ? 1u : 0u;
          for (int m = 0; m < mh; ++m) {
            const uint32_t a_hi = st16 + (uint32_t)m * (TILE_BYTES >> 4) + koff;
            const uint32_t a_lo = st16 + (uint32_t)(mh + m) * (TILE_BYTES >> 4) + koff;
            for (int c = 0; c < ncs; ++c) {
              const uint32_t b_hi = st16 + (uint32_t)(2 * mh + c) * (TILE_BYTES >> 4) + koff;
              const uint32_t b_lo = st16 + (uint32_t)(2 * mh + nc + c) * (TILE_BYTES >> 4) + koff;
              const uint32_t d = tmem_d + (uint32_t)c * 64u + (m ? HALF : 0u);
              mma_mn(d, a_hi, b_hi, acc);
              mma_mn(d, a_hi, b_lo, 1u);
              mma_mn(d, a_lo, b_hi, 1u);
            }
          }
        }
        mma_commit(&empty[s]);
      }
      mma_commit(&done);
    }
    __syncwarp();
  }
  mbar_wait(&done, 0u);
  tc_fence_after();

  // epilogue: quadrant warp q, lane l -> row 16 q + (l & 15) of half l >> 4
  const int half = lane >> 4;
  const int row = half * 64 + warp * 16 + (lane & 15);
  const bool live = half < mh && row < c_out;
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
  float* prow = partial + ((int64_t)blockIdx.x * c_out + row) * kpad + (int64_t)c_first * 64;
  for (int c = 0; c < ncs; ++c) {
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 32) {
      float v[32];
      tmem_ld32(lane_addr + (uint32_t)(c * 64 + cc), v);
      if (live) {
        float* o = prow + c * 64 + cc;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

// dW[(h*C_out + o)*C_in + c] = sum over the splits, in split order, of partial[s][o][h*C_in + c]
__global__ void __launch_bounds__(256) dw_reduce_kernel(const float* __restrict__ partial, int splits, int c_out, int c_in, int kpad,
                                                        float* __restrict__ dW) {
  const int K = H * c_in;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c_out * K) return;
  const int o = t / K, k = t - o * K;
  const int64_t stride = (int64_t)c_out * kpad;
  const float* p = partial + (int64_t)o * kpad + k;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
    s0 += p[(int64_t)s * stride];
    s1 += p[(int64_t)(s + 1) * stride];
    s2 += p[(int64_t)(s + 2) * stride];
    s3 += p[(int64_t)(s + 3) * stride];
  }
  for (; s < splits; ++s) s0 += p[(int64_t)s * stride];
  const int h = k / c_in, c = k - h * c_in;
  dW[(int64_t)(h * c_out + o) * c_in + c] = (s0 + s1) + (s2 + s3);
}

// ------------------------------------------------------------------------------ dx += dP . U,  dU = dP^T . x   (K = 9)
// CP = power of two >= C threads per row; a thread keeps its channel over all its rows: dU partial sums stay in registers.
__global__ void __launch_bounds__(256) bwd_dpu_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int C, int CP,
                                                      const float* __restrict__ dP, const float* __restrict__ U, float* __restrict__ dx,
                                                      int64_t lddx, float* __restrict__ dU) {
  __shared__ float part[H][256];
  const int tid = threadIdx.x, rpb = 256 / CP, c = tid % CP, r0 = tid / CP;
  const bool live = c < C;
  float u[H], acc[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    u[h] = live ? U[h * C + c] : 0.f;
    acc[h] = 0.f;
  }
  for (int64_t n = (int64_t)blockIdx.x * rpb + r0; n < N; n += (int64_t)gridDim.x * rpb) {
    if (!live) continue;
    const float xv = x[n * ldx + c];
    float d = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float p = dP[n * H + h];
      acc[h] = fmaf(p, xv, acc[h]);
      d = fmaf(p, u[h], d);
    }
    if (dx != nullptr) dx[n * lddx + c] += d;
  }
#pragma unroll
  for (int h = 0; h < H; ++h) part[h][tid] = acc[h];
  __syncthreads();
  for (int t = tid; t < H * CP; t += 256) {
    const int h = t / CP, cc = t - h * CP;
    if (cc < C) {
      float s = 0.f;
      for (int r = 0; r < rpb; ++r) s += part[h][r * CP + cc];
      atomicAdd(dU + h * C + cc, s);
    }
  }
}

// ------------------------------------------------------------------------------ host side
struct Plan {
  int kpad, gpad, mh, ct, nsl, nc, stages;
  int64_t splits, nodes_per_split;
  uint32_t tmem_cols;
  size_t smem;
};
static Plan plan_of(int64_t N, int c_in, int c_out) {
  Plan p{};
  p.kpad = (int)(cdiv(H * c_in, BK) * BK);
  p.gpad = c_out <= 64 ? 64 : 128;
  p.mh = p.gpad / 64;
  p.ct = p.kpad / 64;
  p.nsl = (int)cdiv(p.ct, 6);
  p.nc = (int)cdiv(p.ct, p.nsl);
  const size_t stage = (size_t)(2 * p.mh + 2 * p.nc) * TILE_BYTES;
  p.stages = (int)((200 * 1024) / stage);
  if (p.stages > 6) p.stages = 6;
  p.smem = (size_t)p.stages * stage + 1024;
  const uint32_t cols = (uint32_t)p.nc * 64u;
  p.tmem_cols = cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512));
  // splits: about one CTA per SM over all slices; at least 8 stages of work each; partial sums stay under a quarter of the input
  const int64_t blocks = cdiv(N > 0 ? N : 1, KB);
  int64_t s = 148 / p.nsl;
  if (s > blocks / 8) s = blocks / 8;
  if (s > N / (4 * (int64_t)c_out)) s = N / (4 * (int64_t)c_out);
  if (s < 1) s = 1;
  p.nodes_per_split = cdiv(blocks, s) * KB;
  p.splits = cdiv(N > 0 ? N : 1, p.nodes_per_split);
  return p;
}

struct Ws {
  double* P;
  __nv_bfloat16 *Z, *G, *Wt;
  float *dZ, *zero, *dP, *partial;
};
template <class C>
static void carve(C& c, int64_t N, int c_in, int c_out, Ws* out) {
  const Plan p = plan_of(N, c_in, c_out);
  Ws w;
  w.P = c.template take<double>((size_t)N * H);
  w.Z = c.template take<__nv_bfloat16>((size_t)2 * N * p.kpad);
  w.G = c.template take<__nv_bfloat16>((size_t)2 * N * p.gpad);
  w.Wt = c.template take<__nv_bfloat16>((size_t)2 * p.kpad * p.gpad);
  w.dZ = c.template take<float>((size_t)N * p.kpad);
  w.zero = c.template take<float>(256);
  w.dP = c.template take<float>((size_t)N * H);
  w.partial = c.template take<float>((size_t)p.splits * c_out * p.kpad);
  if (out) *out = w;
}
struct NullCarver {
  Sizer s;
  template <typename T>
  T* take(size_t n) { s.take<T>(n); return nullptr; }
};

}  // namespace bwd
}  // namespace geobi

using namespace geobi;

extern "C" size_t geobi_feast_bwd_ws_bytes(int64_t n_nodes, int c_in, int c_out) {
  if (n_nodes < 0 || c_in < 1 || c_in > 128 || c_out < 4 || c_out > 128) return 0;
  bwd::NullCarver c;
  bwd::carve(c, n_nodes, c_in, c_out, nullptr);
  return c.s.total();
}

extern "C" int geobi_feast_bwd(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr, const float* W,
                               const float* U, const float* c, int c_out, float act_slope, const float* out, int64_t ldo,
                               const float* g_out, int64_t ldg, float* dx, int64_t lddx, float* dW, float* dU, float* dc, float* dbias,
                               void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(x && rowptr && W && U && c && g_out && dW && dU && dc && dbias && N >= 0, "feast_bwd: bad arguments");
  GEOBI_REQUIRE(c_in >= 1 && c_in <= 128 && c_out >= 4 && c_out <= 128 && c_out % 4 == 0, "feast_bwd: C_in must be 1..128, C_out a multiple of 4 up to 128");
  GEOBI_REQUIRE(act_slope == 1.0f || out != nullptr, "feast_bwd: the layer's output is needed to differentiate its activation");
  GEOBI_REQUIRE(ldg % 4 == 0 && (reinterpret_cast<uintptr_t>(g_out) & 15) == 0 && (out == nullptr || (ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)),
                "feast_bwd: rows of g_out / out must be 16-byte aligned");
  if (!ws || ws_bytes < geobi_feast_bwd_ws_bytes(N, c_in, c_out) || (reinterpret_cast<uintptr_t>(ws) & 127) != 0) {
    set_error("feast_bwd: workspace missing, too small or not 128-byte aligned");
    return GEOBI_ERR_WORKSPACE;
  }
  const int K = bwd::H * c_in;
  GEOBI_CUDA_OK(cudaMemsetAsync(dU, 0, sizeof(float) * (size_t)K, st));
  GEOBI_CUDA_OK(cudaMemsetAsync(dc, 0, sizeof(float) * bwd::H, st));
  GEOBI_CUDA_OK(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)c_out, st));
  if (N == 0) return GEOBI_OK;
  const bwd::Plan p = bwd::plan_of(N, c_in, c_out);
  Carver cv(ws, ws_bytes);
  bwd::Ws w;
  bwd::carve(cv, N, c_in, c_out, &w);

  // g planes + dbias, W_flat^T chunks
  {
    const int rpb = 256 / (p.gpad >> 2);
    const int64_t want = cdiv(N, rpb);
    bwd::bwd_prep_g_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(g_out, ldg, out, ldo, N, c_out, p.gpad, act_slope, w.G, dbias);
    GEOBI_LAUNCH_OK("bwd_prep_g");
  }
  const bwd::Chunks ch = bwd::chunks_of(p.kpad);
  bwd::bwd_prep_wt_kernel<<<64, 256, 0, st>>>(W, c_in, c_out, p.kpad, p.gpad, ch, w.Wt);
  GEOBI_LAUNCH_OK("bwd_prep_wt");
  GEOBI_CUDA_OK(cudaMemsetAsync(w.zero, 0, sizeof(float) * 256, st));

  // forward intermediates: P (packed double-float) and Z as bf16 hi | lo planes
  if (p.kpad != K && !feast_aggregate_fills_padding(c_in, ldx, p.kpad, x, false))
    GEOBI_CUDA_OK(cudaMemsetAsync(w.Z, 0, sizeof(__nv_bfloat16) * (size_t)2 * N * p.kpad, st));
  int rc = feast_project_and_aggregate(x, ldx, N, c_in, rowptr, nbr, nullptr, N, U, c, w.P, w.Z, p.kpad, 2, st);
  if (rc) return rc;

  // dZ = g . W_flat (fp32 [N, kpad]; columns >= 9 C_in are zero)
  for (int j = 0; j < ch.n; ++j) {
    rc = tc::gemm_dispatch(w.G, N * (int64_t)p.gpad, N, p.gpad, w.Wt + (int64_t)2 * ch.r0[j] * p.gpad, ch.nt[j], w.zero, 1.0f, w.dZ + ch.r0[j],
                           p.kpad, 3, st);
    if (rc) return rc;
  }

  // edge part: dx (through the gathered rows), dP, dc
  GEOBI_CUDA_OK(cudaMemsetAsync(w.dP, 0, sizeof(float) * (size_t)N * bwd::H, st));
  if (dx != nullptr) GEOBI_CUDA_OK(cudaMemset2DAsync(dx, sizeof(float) * (size_t)lddx, 0, sizeof(float) * (size_t)c_in, (size_t)N, st));
  rc = feast_bwd_edges_launch(x, ldx, N, c_in, rowptr, nbr, w.P, c, w.dZ, p.kpad, dx, lddx, w.dP, dc, st);
  if (rc) return rc;

  // dW = g^T . Z
  {
    CUtensorMap tg_hi, tg_lo, tz_hi, tz_lo;
    const bool ok = tc::make_tmap(&tg_hi, w.G, N, p.gpad, bwd::KB) && tc::make_tmap(&tg_lo, w.G + N * (int64_t)p.gpad, N, p.gpad, bwd::KB) &&
                    tc::make_tmap(&tz_hi, w.Z, N, p.kpad, bwd::KB) && tc::make_tmap(&tz_lo, w.Z + N * (int64_t)p.kpad, N, p.kpad, bwd::KB);
    if (!ok) {
      set_error("feast_bwd: cuTensorMapEncodeTiled unavailable");
      return GEOBI_ERR_CUDA;
    }
    GEOBI_CUDA_OK(cudaFuncSetAttribute(bwd::dw_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    bwd::dw_splitk_kernel<<<dim3((unsigned)p.splits, (unsigned)p.nsl), 128, p.smem, st>>>(tg_hi, tg_lo, tz_hi, tz_lo, N, p.nodes_per_split, p.mh,
                                                                                          p.nc, p.ct, p.stages, c_out, p.kpad, p.tmem_cols,
                                                                                          w.partial);
    GEOBI_LAUNCH_OK("dw_splitk");
    bwd::dw_reduce_kernel<<<(unsigned)cdiv((int64_t)c_out * K, 256), 256, 0, st>>>(w.partial, (int)p.splits, c_out, c_in, p.kpad, dW);
    GEOBI_LAUNCH_OK("dw_reduce");
  }

  // dx += dP . U,  dU = dP^T . x
  {
    int cp = 8;
    while (cp < c_in) cp <<= 1;
    const int rpb = 256 / cp;
    const int64_t want = cdiv(N, rpb);
    bwd::bwd_dpu_kernel<<<(unsigned)(want < 148 * 4 ? want : 148 * 4), 256, 0, st>>>(x, ldx, N, c_in, cp, w.dP, U, dx, lddx, dU);
    GEOBI_LAUNCH_OK("bwd_dpu");
  }
  return GEOBI_OK;
}
