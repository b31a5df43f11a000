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
// fp32 accumulation in TMEM: fp32-grade results), like the forward projections.  The second half of the file is the backward of the
// two FC heads (geobi_mlp_head_bwd), built from the same two product kernels.
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
// split-K product D[rows <= gpad, kpad] = G^T . Z over N nodes (G planes [N, gpad], Z planes [N, kpad])
static Plan plan_dims(int64_t N, int kpad, int gpad, int rows) {
  Plan p{};
  p.kpad = kpad;
  p.gpad = gpad;
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
  if (s > N / (4 * (int64_t)rows)) s = N / (4 * (int64_t)rows);
  if (s < 1) s = 1;
  p.nodes_per_split = cdiv(blocks, s) * KB;
  p.splits = cdiv(N > 0 ? N : 1, p.nodes_per_split);
  return p;
}
static Plan plan_of(int64_t N, int c_in, int c_out) {
  return plan_dims(N, (int)(cdiv(H * c_in, BK) * BK), c_out <= 64 ? 64 : 128, c_out);
}
// partial[split][rows][kpad] = sum over the split's nodes of G[n, row] Z[n, col]
static int launch_dw_splitk(const Plan& p, const __nv_bfloat16* G, const __nv_bfloat16* Z, int64_t N, int rows, float* partial, cudaStream_t st) {
  CUtensorMap tg_hi, tg_lo, tz_hi, tz_lo;
  const bool ok = make_tmap(&tg_hi, G, N, p.gpad, KB) && make_tmap(&tg_lo, G + N * (int64_t)p.gpad, N, p.gpad, KB) &&
                  make_tmap(&tz_hi, Z, N, p.kpad, KB) && make_tmap(&tz_lo, Z + N * (int64_t)p.kpad, N, p.kpad, KB);
  if (!ok) {
    set_error("split-K product: cuTensorMapEncodeTiled unavailable");
    return GEOBI_ERR_CUDA;
  }
  // the limit is raised to the most any plan asks for (200 KB of stages + alignment slack), not to this plan's need: concurrent
  // callers with different plans must not lower it under each other
  GEOBI_CUDA_OK(cudaFuncSetAttribute(dw_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  dw_splitk_kernel<<<dim3((unsigned)p.splits, (unsigned)p.nsl), 128, p.smem, st>>>(tg_hi, tg_lo, tz_hi, tz_lo, N, p.nodes_per_split, p.mh, p.nc, p.ct,
                                                                                    p.stages, rows, p.kpad, p.tmem_cols, partial);
  GEOBI_LAUNCH_OK("dw_splitk");
  return GEOBI_OK;
}

// ============================================================================== FC head backward (network.py:324-325,340-341)
// y = W2 . a + b2,  a = leaky_relu(h),  h = W1 . f + b1.  Given dy [N, c_out]:
//   da = dy . W2,  dh = da * act'(h),  dW2 = dy^T . a,  db2 = sum dy,  dW1 = dh^T . f,  db1 = sum dh,  df = dh . W1.
// head_bwd_hidden_kernel recomputes h on tcgen05 (f and W1 split hi + lo; b1 rides along as column C_IN of W1 against a column of
// ones in f) and writes a and dh as bf16 hi | lo planes [N, hidden]; the three products that follow are the kernels above
// (split-K over the nodes for dW2 and [dW1; db1], tc_gemm_tma_kernel for df).

constexpr int HEAD_CIN = 32;      // fc_v1 / fc_f1 input width
constexpr int HEAD_NT = 256;      // hidden units per CTA

// fp32 [N, 32] -> planes [N, 64]: columns 0-31 = f, column 32 = 1 (bias / db1 column), rest 0
__global__ void head_split_f_kernel(const float* __restrict__ f, int64_t ldf, int64_t N, __nv_bfloat16* __restrict__ F) {
  const int64_t total = N * 16;                      // one thread = 4 columns
  const int64_t plane = N * 64;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = t >> 4;
    const int c0 = (int)(t & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 < HEAD_CIN) v = *reinterpret_cast<const float4*>(f + n * ldf + c0);
    else if (c0 == HEAD_CIN) v.x = 1.0f;
    uint2 hi, lo;
    split_bf16x4(v, hi, lo);
    *reinterpret_cast<uint2*>(F + n * 64 + c0) = hi;
    *reinterpret_cast<uint2*>(F + plane + n * 64 + c0) = lo;
  }
}
// dy fp32 [N, c_out <= 4] -> planes [N, 64] (columns >= c_out zero); db2 = sum_n dy.  One thread = 8 columns of a row.
__global__ void __launch_bounds__(256) head_split_dy_kernel(const float* __restrict__ dy, int64_t lddy, int64_t N, int c_out,
                                                            __nv_bfloat16* __restrict__ DY, float* __restrict__ db2) {
  __shared__ float part[4][32];
  const int64_t plane = N * 64;
  const int tid = threadIdx.x, grp = tid & 7;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t n = (int64_t)blockIdx.x * 32 + (tid >> 3); n < N; n += (int64_t)gridDim.x * 32) {
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
    if (grp == 0) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      v.x = dy[n * lddy];
      if (c_out > 1) v.y = dy[n * lddy + 1];
      if (c_out > 2) v.z = dy[n * lddy + 2];
      if (c_out > 3) v.w = dy[n * lddy + 3];
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      uint2 h2, l2;
      split_bf16x4(v, h2, l2);
      hi.x = h2.x; hi.y = h2.y;
      lo.x = l2.x; lo.y = l2.y;
    }
    *reinterpret_cast<uint4*>(DY + n * 64 + grp * 8) = hi;
    *reinterpret_cast<uint4*>(DY + plane + n * 64 + grp * 8) = lo;
  }
  if (grp == 0)
    for (int k = 0; k < 4; ++k) part[k][tid >> 3] = acc[k];
  __syncthreads();
  if (tid < c_out) {
    float s = 0.f;
    for (int r = 0; r < 32; ++r) s += part[tid][r];
    atomicAdd(db2 + tid, s);
  }
}
// W1p planes [hidden, 64]: columns 0-31 = W1[j, :], column 32 = b1[j];  W1t planes [32, hidden]: W1t[c, j] = W1[j, c]
__global__ void head_prep_w1_kernel(const float* __restrict__ W1, const float* __restrict__ b1, int hidden, __nv_bfloat16* __restrict__ W1p,
                                    __nv_bfloat16* __restrict__ W1t) {
  const int total = hidden * 64;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int j = t >> 6, c = t & 63;
    const float v = c < HEAD_CIN ? W1[j * HEAD_CIN + c] : (c == HEAD_CIN ? b1[j] : 0.f);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    W1p[t] = hi;
    W1p[total + t] = lo;
    if (c < HEAD_CIN) {
      W1t[c * hidden + j] = hi;
      W1t[HEAD_CIN * hidden + c * hidden + j] = lo;
    }
  }
}

// grid = (row tiles of 128, hidden / 256).  One K block (64 columns of the f planes): TMA loads the four operand tiles, one thread
// issues the 12 MMAs (M128 N256 K16 x 4 K steps x 3 passes), then thread = row reads h from TMEM 32 hidden units at a time.
__global__ void __launch_bounds__(128) head_bwd_hidden_kernel(const __grid_constant__ CUtensorMap tm_f_hi, const __grid_constant__ CUtensorMap tm_f_lo,
                                                              const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                                                              int64_t N, int hidden, const float* __restrict__ W2, int c_out,
                                                              const float* __restrict__ dy, int64_t lddy, float slope,
                                                              __nv_bfloat16* __restrict__ Aq, __nv_bfloat16* __restrict__ Dq) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full;
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float w2s[4][HEAD_NT];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int A_BYTES = BM * 128, B_BYTES = HEAD_NT * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int j0 = blockIdx.y * HEAD_NT;

  if (tid == 0) {
    mbar_init(&full, 1);
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, HEAD_NT);
  for (int t = tid; t < 4 * HEAD_NT; t += 128) {
    const int k = t / HEAD_NT, j = t - k * HEAD_NT;
    w2s[k][j] = k < c_out ? W2[(int64_t)k * hidden + j0 + j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&full, (uint32_t)(2 * A_BYTES + 2 * B_BYTES));
      tma_load_2d(base, &tm_f_hi, 0, (int)m0, &full);
      tma_load_2d(base + A_BYTES, &tm_f_lo, 0, (int)m0, &full);
      tma_load_2d(base + 2 * A_BYTES, &tm_w_hi, 0, j0, &full);
      tma_load_2d(base + 2 * A_BYTES + B_BYTES, &tm_w_lo, 0, j0, &full);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(&full, 0u);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc(BM, HEAD_NT);
      const uint64_t ah = make_desc(base), al = make_desc(base + A_BYTES);
      const uint64_t bh = make_desc(base + 2 * A_BYTES), bl = make_desc(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
      for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, ah + 2 * k16, bh + 2 * k16, idesc, k16 ? 1u : 0u);
#pragma unroll
      for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, ah + 2 * k16, bl + 2 * k16, idesc, 1u);
#pragma unroll
      for (int k16 = 0; k16 < BK / 16; ++k16) mma_f16(tmem_d, al + 2 * k16, bh + 2 * k16, idesc, 1u);
      mma_commit(&done);
    }
    __syncwarp();
  }
  mbar_wait(&done, 0u);
  tc_fence_after();

  const int64_t m = m0 + tid;
  const bool live = m < N;
  float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
  if (live) {
    d0 = dy[m * lddy];
    if (c_out > 1) d1 = dy[m * lddy + 1];
    if (c_out > 2) d2 = dy[m * lddy + 2];
    if (c_out > 3) d3 = dy[m * lddy + 3];
  }
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
  const int64_t plane = N * (int64_t)hidden;
  // The planes are written through a per-warp staging area that aliases the operand tiles (dead once `done` has fired): a thread
  // owns a ROW in tensor memory, so direct stores are 16-byte pieces 2 KB apart (measured 1.0 TB/s); staged, 8 lanes write one full
  // 128-byte line.  Staging rows are 128 B = 64 hidden units, 16-byte chunks XOR-ed with row % 8 (conflict-free both ways).
  uint8_t* stg = smem_raw + (base - smem_u32(smem_raw)) + warp * 16384;     // 4 planes x 32 rows x 128 B
  const int lane = tid & 31;
#pragma unroll 1
  for (int c0 = 0; c0 < HEAD_NT; c0 += 64) {
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      float v[32];
      tmem_ld32(lane_addr + (uint32_t)(c0 + 32 * hf), v);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint2 ah[2], al[2], dh[2], dl[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int jj = c0 + 32 * hf + j + 4 * q;
          const float4 w0 = *reinterpret_cast<const float4*>(&w2s[0][jj]);
          const float4 w1 = *reinterpret_cast<const float4*>(&w2s[1][jj]);
          const float4 w2 = *reinterpret_cast<const float4*>(&w2s[2][jj]);
          const float4 w3 = *reinterpret_cast<const float4*>(&w2s[3][jj]);
          const float h0 = v[j + 4 * q], h1 = v[j + 4 * q + 1], h2 = v[j + 4 * q + 2], h3 = v[j + 4 * q + 3];
          float4 a, d;
          d.x = fmaf(d3, w3.x, fmaf(d2, w2.x, fmaf(d1, w1.x, d0 * w0.x)));
          d.y = fmaf(d3, w3.y, fmaf(d2, w2.y, fmaf(d1, w1.y, d0 * w0.y)));
          d.z = fmaf(d3, w3.z, fmaf(d2, w2.z, fmaf(d1, w1.z, d0 * w0.z)));
          d.w = fmaf(d3, w3.w, fmaf(d2, w2.w, fmaf(d1, w1.w, d0 * w0.w)));
          a.x = h0 > 0.f ? h0 : h0 * slope;  d.x = h0 > 0.f ? d.x : d.x * slope;
          a.y = h1 > 0.f ? h1 : h1 * slope;  d.y = h1 > 0.f ? d.y : d.y * slope;
          a.z = h2 > 0.f ? h2 : h2 * slope;  d.z = h2 > 0.f ? d.z : d.z * slope;
          a.w = h3 > 0.f ? h3 : h3 * slope;  d.w = h3 > 0.f ? d.w : d.w * slope;
          split_bf16x4(a, ah[q], al[q]);
          split_bf16x4(d, dh[q], dl[q]);
        }
        const int off = lane * 128 + (((4 * hf + (j >> 3)) ^ (lane & 7)) << 4);
        *reinterpret_cast<uint4*>(stg + off) = make_uint4(ah[0].x, ah[0].y, ah[1].x, ah[1].y);
        *reinterpret_cast<uint4*>(stg + 4096 + off) = make_uint4(al[0].x, al[0].y, al[1].x, al[1].y);
        *reinterpret_cast<uint4*>(stg + 8192 + off) = make_uint4(dh[0].x, dh[0].y, dh[1].x, dh[1].y);
        *reinterpret_cast<uint4*>(stg + 12288 + off) = make_uint4(dl[0].x, dl[0].y, dl[1].x, dl[1].y);
      }
    }
    __syncwarp();
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3);
      const int64_t mr = m0 + warp * 32 + r;
      if (mr < N) {
        const int off = r * 128 + ((ch ^ (r & 7)) << 4);
        const int64_t g = mr * hidden + j0 + c0 + ch * 8;
        *reinterpret_cast<uint4*>(Aq + g) = *reinterpret_cast<const uint4*>(stg + off);
        *reinterpret_cast<uint4*>(Aq + plane + g) = *reinterpret_cast<const uint4*>(stg + 4096 + off);
        *reinterpret_cast<uint4*>(Dq + g) = *reinterpret_cast<const uint4*>(stg + 8192 + off);
        *reinterpret_cast<uint4*>(Dq + plane + g) = *reinterpret_cast<const uint4*>(stg + 12288 + off);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, HEAD_NT);
}

// out[r][col] = sum over the splits (fixed order) of partial[s][r][col];  transpose: rows 0..c_in-1 go to dW1[col][r] and row c_in
// to db1[col] (the [dW1; db1] product), otherwise rows go to dW2[r][col]
__global__ void __launch_bounds__(256) head_reduce_kernel(const float* __restrict__ partial, int splits, int rows, int kpad, int transpose,
                                                          float* __restrict__ out_a, float* __restrict__ out_b) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * kpad) return;
  const int r = t / kpad, col = t - r * kpad;
  const int64_t stride = (int64_t)rows * kpad;
  const float* p = partial + (int64_t)r * kpad + col;
  float s0 = 0.f, s1 = 0.f;
  int s = 0;
  for (; s + 2 <= splits; s += 2) {
    s0 += p[(int64_t)s * stride];
    s1 += p[(int64_t)(s + 1) * stride];
  }
  if (s < splits) s0 += p[(int64_t)s * stride];
  const float v = s0 + s1;
  if (!transpose) out_a[(int64_t)r * kpad + col] = v;
  else if (r < rows - 1) out_a[(int64_t)col * (rows - 1) + r] = v;
  else out_b[col] = v;
}

struct HeadWs {
  __nv_bfloat16 *F, *W1p, *W1t, *DY, *A, *D;
  float *zero, *partial;
};
template <class C>
static void carve_head(C& c, int64_t N, int hidden, HeadWs* out) {
  const Plan pa = plan_dims(N, hidden, 64, 4), pb = plan_dims(N, hidden, 64, HEAD_CIN + 1);
  HeadWs w;
  w.F = c.template take<__nv_bfloat16>((size_t)2 * N * 64);
  w.W1p = c.template take<__nv_bfloat16>((size_t)2 * hidden * 64);
  w.W1t = c.template take<__nv_bfloat16>((size_t)2 * HEAD_CIN * hidden);
  w.DY = c.template take<__nv_bfloat16>((size_t)2 * N * 64);
  w.A = c.template take<__nv_bfloat16>((size_t)2 * N * hidden);
  w.D = c.template take<__nv_bfloat16>((size_t)2 * N * hidden);
  w.zero = c.template take<float>(256);
  const size_t pa_n = (size_t)pa.splits * 4 * hidden, pb_n = (size_t)pb.splits * (HEAD_CIN + 1) * hidden;
  w.partial = c.template take<float>(pa_n > pb_n ? pa_n : pb_n);
  if (out) *out = w;
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
  rc = bwd::launch_dw_splitk(p, w.G, w.Z, N, c_out, w.partial, st);
  if (rc) return rc;
  bwd::dw_reduce_kernel<<<(unsigned)cdiv((int64_t)c_out * K, 256), 256, 0, st>>>(w.partial, (int)p.splits, c_out, c_in, p.kpad, dW);
  GEOBI_LAUNCH_OK("dw_reduce");

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

extern "C" size_t geobi_mlp_head_bwd_ws_bytes(int64_t n, int c_in, int hidden) {
  if (n < 0 || c_in != bwd::HEAD_CIN || hidden < bwd::HEAD_NT || hidden % bwd::HEAD_NT != 0 || hidden > 4096) return 0;
  bwd::NullCarver c;
  bwd::carve_head(c, n, hidden, nullptr);
  return c.s.total();
}

extern "C" int geobi_mlp_head_bwd(const float* f, int64_t ldf, int64_t N, int c_in, const float* W1, const float* b1, int hidden,
                                  const float* W2, int c_out, float act_slope, const float* dy, int64_t lddy, float* df, int64_t lddf,
                                  float* dW1, float* db1, float* dW2, float* db2, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(f && W1 && b1 && W2 && dy && dW1 && db1 && dW2 && db2 && N >= 0, "mlp_head_bwd: bad arguments");
  GEOBI_REQUIRE(c_in == bwd::HEAD_CIN, "mlp_head_bwd: c_in must be 32 (got %d)", c_in);
  GEOBI_REQUIRE(hidden >= bwd::HEAD_NT && hidden % bwd::HEAD_NT == 0 && hidden <= 4096, "mlp_head_bwd: hidden must be a multiple of 256, <= 4096");
  GEOBI_REQUIRE(c_out >= 1 && c_out <= 4, "mlp_head_bwd: c_out must be 1..4 (got %d)", c_out);
  GEOBI_REQUIRE(ldf % 4 == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0, "mlp_head_bwd: feature rows must be 16-byte aligned");
  GEOBI_REQUIRE(df == nullptr || (lddf % 4 == 0 && (reinterpret_cast<uintptr_t>(df) & 15) == 0), "mlp_head_bwd: df rows must be 16-byte aligned");
  if (!ws || ws_bytes < geobi_mlp_head_bwd_ws_bytes(N, c_in, hidden) || (reinterpret_cast<uintptr_t>(ws) & 127) != 0) {
    set_error("mlp_head_bwd: workspace missing, too small or not 128-byte aligned");
    return GEOBI_ERR_WORKSPACE;
  }
  GEOBI_CUDA_OK(cudaMemsetAsync(dW1, 0, sizeof(float) * (size_t)hidden * c_in, st));
  GEOBI_CUDA_OK(cudaMemsetAsync(db1, 0, sizeof(float) * (size_t)hidden, st));
  GEOBI_CUDA_OK(cudaMemsetAsync(dW2, 0, sizeof(float) * (size_t)c_out * hidden, st));
  GEOBI_CUDA_OK(cudaMemsetAsync(db2, 0, sizeof(float) * (size_t)c_out, st));
  if (N == 0) return GEOBI_OK;
  Carver cv(ws, ws_bytes);
  bwd::HeadWs w;
  bwd::carve_head(cv, N, hidden, &w);
  GEOBI_CUDA_OK(cudaMemsetAsync(w.zero, 0, sizeof(float) * 256, st));

  const int64_t nb = cdiv(N * 16, 256);
  bwd::head_split_f_kernel<<<(unsigned)(nb < 148 * 16 ? nb : 148 * 16), 256, 0, st>>>(f, ldf, N, w.F);
  bwd::head_prep_w1_kernel<<<64, 256, 0, st>>>(W1, b1, hidden, w.W1p, w.W1t);
  {
    const int64_t want = cdiv(N, 32);
    bwd::head_split_dy_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(dy, lddy, N, c_out, w.DY, db2);
  }
  GEOBI_LAUNCH_OK("mlp_head_bwd prep");

  // a = leaky_relu(h) and dh = (dy . W2) * act'(h) as bf16 planes
  {
    CUtensorMap tf_hi, tf_lo, tw_hi, tw_lo;
    const bool ok = tc::make_tmap(&tf_hi, w.F, N, 64, tc::BM) && tc::make_tmap(&tf_lo, w.F + N * 64, N, 64, tc::BM) &&
                    tc::make_tmap(&tw_hi, w.W1p, hidden, 64, bwd::HEAD_NT) && tc::make_tmap(&tw_lo, w.W1p + (int64_t)hidden * 64, hidden, 64, bwd::HEAD_NT);
    if (!ok) {
      set_error("mlp_head_bwd: cuTensorMapEncodeTiled unavailable");
      return GEOBI_ERR_CUDA;
    }
    const size_t smem = (size_t)2 * tc::BM * 128 + (size_t)2 * bwd::HEAD_NT * 128 + 1024;
    GEOBI_CUDA_OK(cudaFuncSetAttribute(bwd::head_bwd_hidden_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd::head_bwd_hidden_kernel<<<dim3((unsigned)cdiv(N, tc::BM), (unsigned)(hidden / bwd::HEAD_NT)), 128, smem, st>>>(
        tf_hi, tf_lo, tw_hi, tw_lo, N, hidden, W2, c_out, dy, lddy, act_slope, w.A, w.D);
    GEOBI_LAUNCH_OK("head_bwd_hidden");
  }

  // dW2 = dy^T . a
  {
    const bwd::Plan p = bwd::plan_dims(N, hidden, 64, 4);
    int rc = bwd::launch_dw_splitk(p, w.DY, w.A, N, c_out, w.partial, st);
    if (rc) return rc;
    bwd::head_reduce_kernel<<<(unsigned)cdiv((int64_t)c_out * hidden, 256), 256, 0, st>>>(w.partial, (int)p.splits, c_out, hidden, 0, dW2, nullptr);
    GEOBI_LAUNCH_OK("head_reduce (dW2)");
  }
  // [dW1; db1] = [f | 1]^T . dh
  {
    const int rows = bwd::HEAD_CIN + 1;
    const bwd::Plan p = bwd::plan_dims(N, hidden, 64, rows);
    int rc = bwd::launch_dw_splitk(p, w.F, w.D, N, rows, w.partial, st);
    if (rc) return rc;
    bwd::head_reduce_kernel<<<(unsigned)cdiv((int64_t)rows * hidden, 256), 256, 0, st>>>(w.partial, (int)p.splits, rows, hidden, 1, dW1, db1);
    GEOBI_LAUNCH_OK("head_reduce (dW1, db1)");
  }
  // df = dh . W1
  if (df != nullptr) {
    int rc = tc::gemm_dispatch(w.D, N * (int64_t)hidden, N, hidden, w.W1t, bwd::HEAD_CIN, w.zero, 1.0f, df, lddf, 3, st);
    if (rc) return rc;
  }
  return GEOBI_OK;
}
