// FeaSt convolution (aggregate-first) and the fused FC head — fp32 CUDA-core path.
//
//   P[n,h]   = U_h . x_n                      (fp64, so P_j - P_i is as accurate as U(x_j - x_i))
//   q_ijh    = softmax_h((float)(P_j - P_i) + c)
//   Z[i,h,:] = 1/(deg_i+1) * sum_{j in N(i)+{i}} q_ijh x_j         <- the only per-edge work
//   out_i    = act(W_flat . Z_i + b),  W_flat[o, h*C_in+c] = lin.weight[h*C_out+o, c]
//
// No per-edge tensor is written to HBM (the reference materialises [E, 9*C_out]).
// The bf16 tensor-core projection lives in feast_tc.cu.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace geobi {

constexpr int H = GEOBI_HEADS;

// ------------------------------------------------------------------------------ P = X U^T (fp64)
// One thread owns two nodes x all 9 heads (18 fp64 accumulators), so every staged x value is converted once and feeds
// 9 DFMAs, and every broadcast U load feeds 4.  x streams through shared memory in 32-channel chunks, double buffered
// with cp.async (the copy of chunk k+1 runs under the DFMAs of chunk k; conflict-free float4 row reads at a 36-float
// pitch), which leaves the fp64 pipe (64 DFMA/clk/SM measured, profiles/micro/dfma_bench.cu) as the limiter.
// Channels are accumulated in ascending order with one fma each - the same order as the oracle's float64 reference.
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
constexpr int PROJ_NODES = 256;
constexpr int PROJ_THREADS = 128;
constexpr int PROJ_CC = 32;
constexpr int PROJ_LD = PROJ_CC + 4;

static inline size_t proj_smem_bytes(int c_in) {
  const int cp = (c_in + 3) & ~3;
  return (size_t)H * cp * sizeof(double) + (size_t)(cp > PROJ_CC ? 2 : 1) * PROJ_NODES * PROJ_LD * sizeof(float);
}

// CT: compile-time channel count (32 / 64 / 128: the network's layers; folds the divisions of the staging loop and the bounds tests -
// the generic instance spent 43 M of its 62 M warp instructions per 1 M x 64 launch outside the DFMAs), 0 = runtime C.
template <int CT>
__global__ void __launch_bounds__(PROJ_THREADS) feast_project_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int C_rt,
                                                                     const float* __restrict__ U, double* __restrict__ P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = CT ? CT : C_rt;
  const int Cp = (C + 3) & ~3;
  double* Us = reinterpret_cast<double*>(smem_raw);          // [H][Cp], zero padded
  float* xs0 = reinterpret_cast<float*>(Us + H * Cp);        // [2][PROJ_NODES][PROJ_LD] (one buffer when a single chunk covers C)
  const int tid = threadIdx.x;
  const int64_t node0 = (int64_t)blockIdx.x * PROJ_NODES;
  const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int n_chunks = (Cp + PROJ_CC - 1) / PROJ_CC;

  auto stage = [&](int chunk) {
    float* xs = xs0 + (size_t)(chunk & 1) * PROJ_NODES * PROJ_LD;
    const int c0 = chunk * PROJ_CC;
    const int cw = min(PROJ_CC, Cp - c0);                    // multiple of 4
    if (vec) {
      const int units = cw >> 2;
      for (int i = tid; i < PROJ_NODES * units; i += PROJ_THREADS) {
        const int r = i / units, q = i - r * units;
        const int64_t n = node0 + r;
        const int c = c0 + q * 4;
        float* dst = xs + r * PROJ_LD + q * 4;
        if (n < N && c + 3 < C) {
          cp_async_16(dst, x + n * ldx + c);
        } else {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n < N) {
            const float* src = x + n * ldx + c;
            v.x = src[0];
            if (c + 1 < C) v.y = src[1];
            if (c + 2 < C) v.z = src[2];
          }
          *reinterpret_cast<float4*>(dst) = v;
        }
      }
    } else {
      for (int i = tid; i < PROJ_NODES * cw; i += PROJ_THREADS) {
        const int r = i / cw, c = i - r * cw;
        const int64_t n = node0 + r;
        xs[r * PROJ_LD + c] = (n < N && c0 + c < C) ? x[n * ldx + c0 + c] : 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  stage(0);
  for (int i = tid; i < H * Cp; i += PROJ_THREADS) {
    const int h = i / Cp, c = i - h * Cp;
    Us[i] = c < C ? (double)U[h * C + c] : 0.0;
  }
  double acc[2][H];
#pragma unroll
  for (int h = 0; h < H; ++h) acc[0][h] = acc[1][h] = 0.0;

  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    if (chunk + 1 < n_chunks) {
      stage(chunk + 1);                                      // its buffer was released by the barrier that ended chunk - 1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                                         // chunk's rows (and, first time, Us) visible to every thread
    const int c0 = chunk * PROJ_CC;
    const int cw = min(PROJ_CC, Cp - c0);
    const float* xs = xs0 + (size_t)(chunk & 1) * PROJ_NODES * PROJ_LD;
    const float* xa = xs + tid * PROJ_LD;
    const float* xb = xs + (tid + PROJ_THREADS) * PROJ_LD;
    const double* ub = Us + c0;
#pragma unroll 2
    for (int q = 0; q < cw; q += 4) {
      const float4 fa = *reinterpret_cast<const float4*>(xa + q);
      const float4 fb = *reinterpret_cast<const float4*>(xb + q);
      const double a0 = fa.x, a1 = fa.y, a2 = fa.z, a3 = fa.w;
      const double b0 = fb.x, b1 = fb.y, b2 = fb.z, b3 = fb.w;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const double2 u01 = *reinterpret_cast<const double2*>(ub + h * Cp + q);
        const double2 u23 = *reinterpret_cast<const double2*>(ub + h * Cp + q + 2);
        acc[0][h] = fma(a0, u01.x, acc[0][h]);
        acc[1][h] = fma(b0, u01.x, acc[1][h]);
        acc[0][h] = fma(a1, u01.y, acc[0][h]);
        acc[1][h] = fma(b1, u01.y, acc[1][h]);
        acc[0][h] = fma(a2, u23.x, acc[0][h]);
        acc[1][h] = fma(b2, u23.x, acc[1][h]);
        acc[0][h] = fma(a3, u23.y, acc[0][h]);
        acc[1][h] = fma(b3, u23.y, acc[1][h]);
      }
    }
    __syncthreads();                                         // every thread is done with this buffer
  }
  // transpose through shared memory so P is written in whole rows of the block (coalesced)
  double* ps = reinterpret_cast<double*>(xs0);               // 256 x 9 doubles = 18 KB <= one 36 KB x buffer
#pragma unroll
  for (int h = 0; h < H; ++h) {
    ps[tid * H + h] = p_pack(acc[0][h]);
    ps[(tid + PROJ_THREADS) * H + h] = p_pack(acc[1][h]);
  }
  __syncthreads();
  const int64_t left = N - node0;
  const int cnt = (int)(left < PROJ_NODES ? left : PROJ_NODES) * H;
  double* dst = P + node0 * H;
  for (int i = tid; i < cnt; i += PROJ_THREADS) dst[i] = ps[i];
}

// ------------------------------------------------------------------------------ Z = softmax-weighted neighbour sums
// One warp per target node; lane l owns the CPL adjacent channels [l*CPL, (l+1)*CPL) (one coalesced vector load per
// gathered row).  Per 32-edge chunk: lanes compute one edge's 9 soft assignments each (-> smem), then the warp walks the
// chunk accumulating 9 x CPL FMAs per edge per lane.
// OUT = 0: Z fp32 [N, ldz];  1: bf16 plane (hi);  2: two bf16 planes hi | lo (lo = bf16(z - hi)), plane stride = N*ldz.
template <int CPL>
struct VecLoad;
template <>
struct VecLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = *p; }
};
template <>
struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
};
template <>
struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};

// packed fp32x2 FMA (Blackwell FFMA2): d = a * b + c on both halves of 64-bit register pairs
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi32(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }

template <int CPL, int OUT, bool VEC>
__global__ void __launch_bounds__(256) feast_aggregate_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int C,
                                                              const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                              const double* __restrict__ P, const float* __restrict__ cvec,
                                                              void* __restrict__ Zout, int64_t ldz) {
  // soft assignments of up to 32 edges per warp: heads (0,1)(2,3)(4,5)(6,7) as 64-bit pairs + head 8
  __shared__ __align__(16) float qs[8][32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= N) return;
  const int b = rowptr[i];
  const int total = rowptr[i + 1] - b + 1;  // neighbours + implicit self loop (slot 0)
  const int c0 = lane * CPL;
  const int cl = c0 < C ? c0 : 0;           // lanes past the last channel gather channel 0 and never store
  const unsigned ldx32 = (unsigned)ldx;     // host guarantees N * ldx < 2^32 elements
  double Pi[H];
  float ch[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    Pi[h] = P[i * H + h];
    ch[h] = cvec[h];
  }
  // acc2[p][k]: heads (2p, 2p+1) of channel c0+k;  acc8[k]: head 8
  unsigned long long acc2[4][CPL];
  float acc8[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    acc8[k] = 0.f;
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) acc2[p2][k] = 0ull;
  }

  for (int s0 = 0; s0 < total; s0 += 32) {
    const int s = s0 + lane;
    int j = (int)i;
    float l[H];
    if (s < total) {
      if (s > 0) j = nbr[b + s - 1];
      float m = -INFINITY;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = p_diff(P[(int64_t)j * H + h], Pi[h]) + ch[h];
        m = fmaxf(m, l[h]);
      }
      float sum = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = __expf(l[h] - m);
        sum += l[h];
      }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int h = 0; h < H; ++h) l[h] *= inv;
    } else {
#pragma unroll
      for (int h = 0; h < H; ++h) l[h] = 0.f;   // padding slot: weight 0 on a valid row, keeps the pair loop branch-free
    }
    float4* q4 = reinterpret_cast<float4*>(&qs[warp][lane][0]);
    q4[0] = make_float4(l[0], l[1], l[2], l[3]);
    q4[1] = make_float4(l[4], l[5], l[6], l[7]);
    qs[warp][lane][8] = l[8];
    __syncwarp();
    const int cnt = min(32, total - s0);
    // two edges per iteration, both gathers in flight before either is consumed
    for (int t = 0; t < cnt; t += 2) {
      const unsigned ja = (unsigned)__shfl_sync(0xffffffffu, j, t);
      const unsigned jb = (unsigned)__shfl_sync(0xffffffffu, j, t + 1);   // t+1 <= 31 since t is even
      float xa[CPL], xb[CPL];
      if (VEC) {
        VecLoad<CPL>::ld(x + (ja * ldx32 + (unsigned)cl), xa);
        VecLoad<CPL>::ld(x + (jb * ldx32 + (unsigned)cl), xb);
      } else {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const unsigned ck = cl + k < C ? cl + k : cl;
          xa[k] = x[ja * ldx32 + ck];
          xb[k] = x[jb * ldx32 + ck];
        }
      }
      const float* qrow = &qs[warp][t][0];
      const ulonglong2 qa0 = *reinterpret_cast<const ulonglong2*>(qrow);
      const ulonglong2 qa1 = *reinterpret_cast<const ulonglong2*>(qrow + 4);
      const float qa8 = qrow[8];
      const ulonglong2 qb0 = *reinterpret_cast<const ulonglong2*>(qrow + 12);
      const ulonglong2 qb1 = *reinterpret_cast<const ulonglong2*>(qrow + 16);
      const float qb8 = qrow[20];
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const unsigned long long xxa = pack2(xa[k], xa[k]);
        acc2[0][k] = ffma2(qa0.x, xxa, acc2[0][k]);
        acc2[1][k] = ffma2(qa0.y, xxa, acc2[1][k]);
        acc2[2][k] = ffma2(qa1.x, xxa, acc2[2][k]);
        acc2[3][k] = ffma2(qa1.y, xxa, acc2[3][k]);
        acc8[k] = fmaf(qa8, xa[k], acc8[k]);
      }
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const unsigned long long xxb = pack2(xb[k], xb[k]);
        acc2[0][k] = ffma2(qb0.x, xxb, acc2[0][k]);
        acc2[1][k] = ffma2(qb0.y, xxb, acc2[1][k]);
        acc2[2][k] = ffma2(qb1.x, xxb, acc2[2][k]);
        acc2[3][k] = ffma2(qb1.y, xxb, acc2[3][k]);
        acc8[k] = fmaf(qb8, xb[k], acc8[k]);
      }
    }
    __syncwarp();
  }
  if (c0 >= C) return;
  const float rcnt = 1.0f / (float)total;   // mean over the neighbourhood (scatter-mean upstream)
  float z[H][CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) {
      z[2 * p2][k] = lo32(acc2[p2][k]) * rcnt;
      z[2 * p2 + 1][k] = hi32(acc2[p2][k]) * rcnt;
    }
    z[8][k] = acc8[k] * rcnt;
  }
  if (OUT == 0) {
    float* zrow = static_cast<float*>(Zout) + i * ldz;
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
      for (int k = 0; k < CPL; ++k)
        if (c0 + k < C) zrow[h * C + c0 + k] = z[h][k];
  } else {
    __nv_bfloat16* zhi = static_cast<__nv_bfloat16*>(Zout) + i * ldz;
    __nv_bfloat16* zlo = zhi + N * ldz;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      __nv_bfloat16 hi[CPL], lo[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        hi[k] = __float2bfloat16_rn(z[h][k]);
        lo[k] = __float2bfloat16_rn(z[h][k] - __bfloat162float(hi[k]));
      }
      if (CPL > 1 && VEC) {    // C % CPL == 0 here: one 4- or 8-byte store per plane
        if (CPL == 2) {
          *reinterpret_cast<uint32_t*>(zhi + h * C + c0) = *reinterpret_cast<const uint32_t*>(hi);
          if (OUT == 2) *reinterpret_cast<uint32_t*>(zlo + h * C + c0) = *reinterpret_cast<const uint32_t*>(lo);
        } else {
          *reinterpret_cast<uint2*>(zhi + h * C + c0) = *reinterpret_cast<const uint2*>(hi);
          if (OUT == 2) *reinterpret_cast<uint2*>(zlo + h * C + c0) = *reinterpret_cast<const uint2*>(lo);
        }
      } else {
#pragma unroll
        for (int k = 0; k < CPL; ++k)
          if (c0 + k < C) {
            zhi[h * C + c0 + k] = hi[k];
            if (OUT == 2) zlo[h * C + c0 + k] = lo[k];
          }
      }
    }
  }
}

// ------------------------------------------------------------------------------ packed variant: 128/C nodes per warp
// For C in {32, 64, 128} every lane owns 4 adjacent channels (one 128-bit gather per row) and a warp carries
// NPW = 128 / C nodes side by side: lane group g (LPN = C/4 lanes) owns node g.  The softmax pass then fills all 32 lanes
// (group g's lanes take that node's edges) and every instruction of the accumulation loop advances NPW nodes, which
// halves (C=64) or quarters (C=32) the instruction count per node of the one-node-per-warp version.
template <int NPW, int OUT>
__global__ void __launch_bounds__(256) feast_aggregate_packed_kernel(const float* __restrict__ x, int64_t ldx, int64_t N,
                                                                     const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                                     const double* __restrict__ P, const float* __restrict__ cvec,
                                                                     void* __restrict__ Zout, int64_t ldz) {
  constexpr int LPN = 32 / NPW;      // lanes per node = slots per node per chunk
  constexpr int C = 4 * LPN;
  __shared__ __align__(16) float qs[8][32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPN, sl = lane % LPN;
  const int64_t i_raw = ((int64_t)blockIdx.x * 8 + warp) * NPW + g;
  const bool live = i_raw < N;
  const int64_t i = live ? i_raw : N - 1;          // dead groups shadow the last node and never store
  if (((int64_t)blockIdx.x * 8 + warp) * NPW >= N) return;   // whole warp past the end
  const int b = rowptr[i];
  const int total = rowptr[i + 1] - b + 1;         // neighbours + implicit self loop (slot 0)
  int maxtotal = total;
#pragma unroll
  for (int o = 16; o >= LPN; o >>= 1) maxtotal = max(maxtotal, __shfl_xor_sync(0xffffffffu, maxtotal, o));
  const unsigned ldx32 = (unsigned)ldx;
  const int c0 = sl * 4;
  double Pi[H];
  float ch[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    Pi[h] = P[i * H + h];
    ch[h] = cvec[h];
  }
  unsigned long long acc2[4][4];
  float acc8[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc8[k] = 0.f;
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) acc2[p2][k] = 0ull;
  }

  for (int s0 = 0; s0 < maxtotal; s0 += LPN) {
    // lane (g, sl) handles edge slot s0 + sl of node g
    const int s = s0 + sl;
    int j = (int)i;
    float l[H];
    if (s < total) {
      if (s > 0) j = nbr[b + s - 1];
      float m = -INFINITY;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = p_diff(P[(int64_t)j * H + h], Pi[h]) + ch[h];
        m = fmaxf(m, l[h]);
      }
      float sum = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = __expf(l[h] - m);
        sum += l[h];
      }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int h = 0; h < H; ++h) l[h] *= inv;
    } else {
#pragma unroll
      for (int h = 0; h < H; ++h) l[h] = 0.f;       // padding slot: weight 0 on a valid row
    }
    float4* q4 = reinterpret_cast<float4*>(&qs[warp][lane][0]);
    q4[0] = make_float4(l[0], l[1], l[2], l[3]);
    q4[1] = make_float4(l[4], l[5], l[6], l[7]);
    qs[warp][lane][8] = l[8];
    __syncwarp();
    const int cnt = min(LPN, maxtotal - s0);        // warp-uniform
    const float* qbase = &qs[warp][g * LPN][0];
#pragma unroll 1
    for (int t = 0; t < cnt; t += 2) {
      const unsigned ja = (unsigned)__shfl_sync(0xffffffffu, j, t, LPN);
      const unsigned jb = (unsigned)__shfl_sync(0xffffffffu, j, (t + 1) & (LPN - 1), LPN);
      const float4 xa = *reinterpret_cast<const float4*>(x + (ja * ldx32 + (unsigned)c0));
      const float4 xb = *reinterpret_cast<const float4*>(x + (jb * ldx32 + (unsigned)c0));
      const bool has_b = t + 1 < cnt;
      const float* qa = qbase + t * 12;
      const float* qb = qbase + ((t + 1) & (LPN - 1)) * 12;
      const ulonglong2 qa0 = *reinterpret_cast<const ulonglong2*>(qa);
      const ulonglong2 qa1 = *reinterpret_cast<const ulonglong2*>(qa + 4);
      const float qa8 = qa[8];
      ulonglong2 qb0 = *reinterpret_cast<const ulonglong2*>(qb);
      ulonglong2 qb1 = *reinterpret_cast<const ulonglong2*>(qb + 4);
      float qb8 = qb[8];
      if (!has_b) { qb0.x = qb0.y = qb1.x = qb1.y = 0ull; qb8 = 0.f; }
      const float xav[4] = {xa.x, xa.y, xa.z, xa.w};
      const float xbv[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned long long xx = pack2(xav[k], xav[k]);
        acc2[0][k] = ffma2(qa0.x, xx, acc2[0][k]);
        acc2[1][k] = ffma2(qa0.y, xx, acc2[1][k]);
        acc2[2][k] = ffma2(qa1.x, xx, acc2[2][k]);
        acc2[3][k] = ffma2(qa1.y, xx, acc2[3][k]);
        acc8[k] = fmaf(qa8, xav[k], acc8[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned long long xx = pack2(xbv[k], xbv[k]);
        acc2[0][k] = ffma2(qb0.x, xx, acc2[0][k]);
        acc2[1][k] = ffma2(qb0.y, xx, acc2[1][k]);
        acc2[2][k] = ffma2(qb1.x, xx, acc2[2][k]);
        acc2[3][k] = ffma2(qb1.y, xx, acc2[3][k]);
        acc8[k] = fmaf(qb8, xbv[k], acc8[k]);
      }
    }
    __syncwarp();
  }
  if (!live) return;
  const float rcnt = 1.0f / (float)total;
  float z[H][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) {
      z[2 * p2][k] = lo32(acc2[p2][k]) * rcnt;
      z[2 * p2 + 1][k] = hi32(acc2[p2][k]) * rcnt;
    }
    z[8][k] = acc8[k] * rcnt;
  }
  if (OUT == 0) {
    float* zrow = static_cast<float*>(Zout) + i * ldz;
#pragma unroll
    for (int h = 0; h < H; ++h) *reinterpret_cast<float4*>(zrow + h * C + c0) = make_float4(z[h][0], z[h][1], z[h][2], z[h][3]);
  } else {
    __nv_bfloat16* zhi = static_cast<__nv_bfloat16*>(Zout) + i * ldz;
    __nv_bfloat16* zlo = zhi + N * ldz;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        hi[k] = __float2bfloat16_rn(z[h][k]);
        lo[k] = __float2bfloat16_rn(z[h][k] - __bfloat162float(hi[k]));
      }
      *reinterpret_cast<uint2*>(zhi + h * C + c0) = *reinterpret_cast<const uint2*>(hi);
      if (OUT == 2) *reinterpret_cast<uint2*>(zlo + h * C + c0) = *reinterpret_cast<const uint2*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------ small-C variant (first layer: C_in = 6 / 12)
// Four lanes per node (8 nodes per warp), lane sl owns channels [sl*V, sl*V+V).  Per chunk of four edge slots each lane
// computes the soft assignments of one slot; the slot's nine weights and its row index then travel by quad shuffles, so
// there is no shared memory and no warp barrier.  The generic kernel above spends a whole warp on such a node with
// 6 or 12 of its 32 lanes active.
template <int V, int OUT>
__global__ void __launch_bounds__(256, 2) feast_aggregate_small_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int C,
                                                                    const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                                    const double* __restrict__ P, const float* __restrict__ cvec,
                                                                    void* __restrict__ Zout, int64_t ldz) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, sl = lane & 3;
  const int c0 = sl * V;
  const bool active = c0 < C;                        // C % V == 0 (host)
  const unsigned ldx32 = (unsigned)ldx;
  const float* xl = x + (active ? c0 : 0);
  float ch[H];
#pragma unroll
  for (int h = 0; h < H; ++h) ch[h] = cvec[h];

  // Persistent warps over contiguous ranges of 8-node tiles with a two-deep index pipeline (as feast_aggregate_ps_kernel):
  // tile t runs on (b, total, its 16 first neighbour ids) fetched during tiles t-2 / t-1, so per chunk only the P rows and the
  // feature rows - issued together - are a fresh round trip.
  const int64_t n_tiles = (N + 7) / 8;
  const int64_t wid = (int64_t)blockIdx.x * 8 + warp, n_warps = (int64_t)gridDim.x * 8;
  const int64_t t_begin = (n_tiles * wid) / n_warps, t_end = (n_tiles * (wid + 1)) / n_warps;
  if (t_begin >= t_end) return;
  auto node_of = [&](int64_t tile) -> int64_t {
    const int64_t r = tile * 8 + g;
    return r < N ? r : N - 1;                        // dead quads shadow the last node and never store
  };
  auto load_rowptr = [&](int64_t tile, int& b_, int& total_) {
    if (tile < t_end) {
      const int64_t i_ = node_of(tile);
      b_ = rowptr[i_];
      total_ = rowptr[i_ + 1] - b_ + 1;              // neighbours + implicit self loop (slot 0)
    } else {
      b_ = 0;
      total_ = 1;
    }
  };
  constexpr int PRE = 4;                             // chunks whose neighbour ids are fetched ahead (16 slots: every mesh row)
  auto load_js = [&](int64_t tile, int b_, int total_, int* jv) {
    const int self = (int)node_of(tile < t_end ? tile : t_begin);
#pragma unroll
    for (int k = 0; k < PRE; ++k) {
      const int s_ = 4 * k + sl;
      jv[k] = (tile < t_end && s_ > 0 && s_ < total_) ? nbr[b_ + s_ - 1] : self;
    }
  };
  int b_cur, total_cur, b_nxt, total_nxt, b_nx2 = 0, total_nx2 = 1;
  int j_cur[PRE], j_nxt[PRE];
  load_rowptr(t_begin, b_cur, total_cur);
  load_rowptr(t_begin + 1, b_nxt, total_nxt);
  load_js(t_begin, b_cur, total_cur, j_cur);

  for (int64_t tile = t_begin; tile < t_end; ++tile) {
  const int64_t i_raw = tile * 8 + g;
  const bool live = i_raw < N;
  const int64_t i = live ? i_raw : N - 1;
  const int b = b_cur;
  const int total = total_cur;
  load_js(tile + 1, b_nxt, total_nxt, j_nxt);
  load_rowptr(tile + 2, b_nx2, total_nx2);
  int maxtotal = total;
#pragma unroll
  for (int o = 16; o >= 4; o >>= 1) maxtotal = max(maxtotal, __shfl_xor_sync(0xffffffffu, maxtotal, o));
  double Pi[H];
#pragma unroll
  for (int h = 0; h < H; ++h) Pi[h] = P[i * H + h];
  float acc[H][V];
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int k = 0; k < V; ++k) acc[h][k] = 0.f;

#pragma unroll 1
  for (int s0 = 0; s0 < maxtotal; s0 += 4) {
    const int s = s0 + sl;
    int j = (int)i;                                  // padding slots: weight 0 on the node's own (valid) row
    if (s0 < 4 * PRE) {
#pragma unroll
      for (int k = 0; k < PRE; ++k)
        if (s0 == 4 * k) j = j_cur[k];
    } else if (s < total) {
      j = nbr[b + s - 1];
    }
    // the chunk's four feature rows go out together with the P row (both depend only on j)
    float xv[4][V];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const unsigned jt = (unsigned)__shfl_sync(0xffffffffu, j, t, 4);
      VecLoad<V>::ld(xl + jt * ldx32, xv[t]);
    }
    float l[H];
    if (s < total) {
      float m = -INFINITY;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = p_diff(P[(int64_t)j * H + h], Pi[h]) + ch[h];
        m = fmaxf(m, l[h]);
      }
      float sum = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = __expf(l[h] - m);
        sum += l[h];
      }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int h = 0; h < H; ++h) l[h] *= inv;
    } else {
#pragma unroll
      for (int h = 0; h < H; ++h) l[h] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float q = __shfl_sync(0xffffffffu, l[h], t, 4);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[h][k] = fmaf(q, xv[t][k], acc[h][k]);
      }
    }
  }
  b_cur = b_nxt; total_cur = total_nxt;
  b_nxt = b_nx2; total_nxt = total_nx2;
#pragma unroll
  for (int k = 0; k < PRE; ++k) j_cur[k] = j_nxt[k];
  if (!live) continue;
  if (!active) {
    // bf16 planes feed a GEMM whose K is padded to a multiple of 64: the first idle lane zero-fills columns [9C, ldz)
    if (OUT != 0 && c0 == C) {
      __nv_bfloat16* zhi = static_cast<__nv_bfloat16*>(Zout) + i * ldz;
      __nv_bfloat16* zlo = zhi + N * ldz;
      for (int k = H * C; k < ldz; k += 2) {          // 9C and ldz are even
        *reinterpret_cast<uint32_t*>(zhi + k) = 0u;
        if (OUT == 2) *reinterpret_cast<uint32_t*>(zlo + k) = 0u;
      }
    }
    continue;
  }
  const float rcnt = 1.0f / (float)total;            // mean over the neighbourhood (scatter-mean upstream)
  // one 4V-byte (fp32) / 2V-byte (bf16) store per head: C, c0 and ldz are multiples of V (host)
  if (OUT == 0) {
    float* zrow = static_cast<float*>(Zout) + i * ldz;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      if (V == 4) *reinterpret_cast<float4*>(zrow + h * C + c0) = make_float4(acc[h][0] * rcnt, acc[h][1] * rcnt, acc[h][2] * rcnt, acc[h][3] * rcnt);
      else *reinterpret_cast<float2*>(zrow + h * C + c0) = make_float2(acc[h][0] * rcnt, acc[h][1] * rcnt);
    }
  } else {
    __nv_bfloat16* zhi = static_cast<__nv_bfloat16*>(Zout) + i * ldz;
    __nv_bfloat16* zlo = zhi + N * ldz;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      __nv_bfloat16 hi[V], lo[V];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float z = acc[h][k] * rcnt;
        hi[k] = __float2bfloat16_rn(z);
        lo[k] = __float2bfloat16_rn(z - __bfloat162float(hi[k]));
      }
      if (V == 4) {
        *reinterpret_cast<uint2*>(zhi + h * C + c0) = *reinterpret_cast<const uint2*>(hi);
        if (OUT == 2) *reinterpret_cast<uint2*>(zlo + h * C + c0) = *reinterpret_cast<const uint2*>(lo);
      } else {
        *reinterpret_cast<uint32_t*>(zhi + h * C + c0) = *reinterpret_cast<const uint32_t*>(hi);
        if (OUT == 2) *reinterpret_cast<uint32_t*>(zlo + h * C + c0) = *reinterpret_cast<const uint32_t*>(lo);
      }
    }
  }
  }   // tile loop
}

template <int V>
static void launch_small(int out_mode, cudaStream_t st, const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr,
                         const int32_t* nbr, const double* P, const float* c, void* Z, int64_t ldz) {
  static int resident = 0;
  if (resident == 0) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    resident = sms * 2;                              // __launch_bounds__(256, 2)
  }
  const int64_t want = cdiv(N, 64);
  const unsigned blocks = (unsigned)(want < resident ? want : resident);
  if (out_mode == 0) feast_aggregate_small_kernel<V, 0><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
  else if (out_mode == 1) feast_aggregate_small_kernel<V, 1><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
  else feast_aggregate_small_kernel<V, 2><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
}

template <int NPW>
static void launch_packed(int out_mode, cudaStream_t st, const float* x, int64_t ldx, int64_t N, const int32_t* rowptr, const int32_t* nbr,
                          const double* P, const float* c, void* Z, int64_t ldz) {
  const unsigned blocks = (unsigned)cdiv(N, 8 * NPW);
  if (out_mode == 0) feast_aggregate_packed_kernel<NPW, 0><<<blocks, 256, 0, st>>>(x, ldx, N, rowptr, nbr, P, c, Z, ldz);
  else if (out_mode == 1) feast_aggregate_packed_kernel<NPW, 1><<<blocks, 256, 0, st>>>(x, ldx, N, rowptr, nbr, P, c, Z, ldz);
  else feast_aggregate_packed_kernel<NPW, 2><<<blocks, 256, 0, st>>>(x, ldx, N, rowptr, nbr, P, c, Z, ldz);
}

// ------------------------------------------------------------------------------ packed + staged variant (cp.async)
// Packed layout as above, and the feature rows of the chunk's edges are first copied into shared memory with 16-byte
// cp.async — every gather of the chunk in flight at once, no registers held — while the soft assignments are computed;
// the accumulation loop then runs load-free out of shared memory.  This is what hides the gather latency: the direct
// version keeps only two rows per warp in flight (ncu: long-scoreboard stalls dominate, issue slots < 50 % busy).
// SLOTS = edge slots per node per chunk (<= LPN).

template <int NPW, int OUT, int SLOTS, bool HAS_MAP>
__global__ void __launch_bounds__(256) feast_aggregate_ps_kernel(const float* __restrict__ x, int64_t ldx, int64_t N,
                                                                 const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                                 const int* __restrict__ row_map,
                                                                 const double* __restrict__ P, const float* __restrict__ cvec,
                                                                 void* __restrict__ Zout, int64_t ldz) {
  constexpr int LPN = 32 / NPW;      // lanes per node
  constexpr int C = 4 * LPN;
  static_assert(SLOTS <= LPN, "one lane per edge slot in the softmax pass");
  extern __shared__ __align__(16) float smem_f[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* xs = smem_f + (size_t)warp * (NPW * SLOTS) * C;                              // [NPW*SLOTS][C]
  float* qs = smem_f + (size_t)8 * (NPW * SLOTS) * C + (size_t)warp * (NPW * SLOTS) * 12;  // [NPW*SLOTS][12]
  const int g = lane / LPN, sl = lane % LPN;
  const unsigned ldx32 = (unsigned)ldx;
  const int c0 = sl * 4;
  float ch[H];
#pragma unroll
  for (int h = 0; h < H; ++h) ch[h] = cvec[h];
  float* xrow = xs + (size_t)(g * SLOTS) * C + c0;   // this lane's 16-byte piece of slot 0 of its node
  float* qrow = qs + (size_t)(g * SLOTS) * 12;

  // Persistent warps: each warp walks a contiguous range of NPW-node tiles (neighbouring nodes share gathered rows in L1)
  // with a two-deep index pipeline - iteration t holds (b, total, first-chunk row index) of tile t, loads nbr of tile t+1
  // with the rowptr fetched one iteration earlier and fetches rowptr of tile t+2 - so the rowptr -> nbr -> row chain
  // (three dependent L2 round trips, 37 % of the stall samples of the one-node-per-warp version) is off the critical path.
  const int64_t n_tiles = (N + NPW - 1) / NPW;
  const int64_t wid = (int64_t)blockIdx.x * 8 + warp, n_warps = (int64_t)gridDim.x * 8;
  const int64_t t_begin = (n_tiles * wid) / n_warps, t_end = (n_tiles * (wid + 1)) / n_warps;
  auto node_of = [&](int64_t tile) -> int64_t {
    const int64_t r = tile * NPW + g;
    return r < N ? r : N - 1;                       // dead groups shadow the last node and never store
  };
  auto load_rowptr = [&](int64_t tile, int& b_, int& total_) {
    if (tile < t_end) {
      const int64_t i_ = node_of(tile);
      b_ = rowptr[i_];
      total_ = rowptr[i_ + 1] - b_ + 1;             // neighbours + implicit self loop (slot 0)
    } else {
      b_ = 0;
      total_ = 1;
    }
  };
  // row_map (PoolingLayer.unpooling fused into the conv): x and P rows of node v live at row_map[v]
  auto load_first_j = [&](int64_t tile, int b_, int total_) -> int {
    int j_ = (int)node_of(tile < t_end ? tile : t_begin);
    if (tile < t_end && sl < SLOTS && sl > 0 && sl < total_) j_ = nbr[b_ + sl - 1];
    return HAS_MAP ? row_map[j_] : j_;
  };
  if (t_begin >= t_end) return;
  int b_cur, total_cur, b_nxt, total_nxt, b_nx2 = 0, total_nx2 = 1;
  load_rowptr(t_begin, b_cur, total_cur);
  load_rowptr(t_begin + 1, b_nxt, total_nxt);
  int j_cur = load_first_j(t_begin, b_cur, total_cur);

  for (int64_t tile = t_begin; tile < t_end; ++tile) {
  const int64_t i_raw = tile * NPW + g;
  const bool live = i_raw < N;
  const int64_t i = live ? i_raw : N - 1;
  const int b = b_cur;
  const int total = total_cur;
  const int j_nxt = load_first_j(tile + 1, b_nxt, total_nxt);
  load_rowptr(tile + 2, b_nx2, total_nx2);
  int maxtotal = total;
#pragma unroll
  for (int o = 16; o >= LPN; o >>= 1) maxtotal = max(maxtotal, __shfl_xor_sync(0xffffffffu, maxtotal, o));
  const int i_src = HAS_MAP ? row_map[i] : (int)i;
  double Pi[H];
#pragma unroll
  for (int h = 0; h < H; ++h) Pi[h] = P[(int64_t)i_src * H + h];
  unsigned long long acc2[4][4];
  float acc8[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc8[k] = 0.f;
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) acc2[p2][k] = 0ull;
  }
  for (int s0 = 0; s0 < maxtotal; s0 += SLOTS) {
    const int cnt = min(SLOTS, maxtotal - s0);      // warp-uniform
    const int s = s0 + sl;
    int j = i_src;                                  // padding slots re-read the node's own row with weight 0
    if (s0 == 0) j = j_cur;                         // first chunk: fetched one tile ago
    else if (sl < SLOTS && s < total) {
      j = nbr[b + s - 1];
      if (HAS_MAP) j = row_map[j];
    }
    // 1. launch every gather of the chunk: group g's lanes copy the LPN pieces of each of their node's rows
    for (int t = 0; t < cnt; ++t) {
      const unsigned jt = (unsigned)__shfl_sync(0xffffffffu, j, t, LPN);
      cp_async_16(xrow + t * C, x + (jt * ldx32 + (unsigned)c0));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // 2. soft assignments (lane (g, sl) -> edge slot s0 + sl of node g) while the rows are in flight
    if (sl < SLOTS) {
      float l[H];
      if (s < total) {
        float m = -INFINITY;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          l[h] = p_diff(P[(int64_t)j * H + h], Pi[h]) + ch[h];
          m = fmaxf(m, l[h]);
        }
        float sum = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          l[h] = __expf(l[h] - m);
          sum += l[h];
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int h = 0; h < H; ++h) l[h] *= inv;
      } else {
#pragma unroll
        for (int h = 0; h < H; ++h) l[h] = 0.f;
      }
      float4* q4 = reinterpret_cast<float4*>(qrow + sl * 12);
      q4[0] = make_float4(l[0], l[1], l[2], l[3]);
      q4[1] = make_float4(l[4], l[5], l[6], l[7]);
      qrow[sl * 12 + 8] = l[8];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // 3. load-free accumulation out of shared memory
#pragma unroll 2
    for (int t = 0; t < cnt; ++t) {
      const float4 xv = *reinterpret_cast<const float4*>(xrow + t * C);
      const float* qt = qrow + t * 12;
      const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(qt);
      const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(qt + 4);
      const float q8 = qt[8];
      const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned long long xx = pack2(xe[k], xe[k]);
        acc2[0][k] = ffma2(q0.x, xx, acc2[0][k]);
        acc2[1][k] = ffma2(q0.y, xx, acc2[1][k]);
        acc2[2][k] = ffma2(q1.x, xx, acc2[2][k]);
        acc2[3][k] = ffma2(q1.y, xx, acc2[3][k]);
        acc8[k] = fmaf(q8, xe[k], acc8[k]);
      }
    }
    __syncwarp();
  }
  b_cur = b_nxt; total_cur = total_nxt; j_cur = j_nxt;
  b_nxt = b_nx2; total_nxt = total_nx2;
  if (!live) continue;
  const float rcnt = 1.0f / (float)total;
  float z[H][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) {
      z[2 * p2][k] = lo32(acc2[p2][k]) * rcnt;
      z[2 * p2 + 1][k] = hi32(acc2[p2][k]) * rcnt;
    }
    z[8][k] = acc8[k] * rcnt;
  }
  if (OUT == 0) {
    float* zrow = static_cast<float*>(Zout) + i * ldz;
#pragma unroll
    for (int h = 0; h < H; ++h) *reinterpret_cast<float4*>(zrow + h * C + c0) = make_float4(z[h][0], z[h][1], z[h][2], z[h][3]);
  } else {
    __nv_bfloat16* zhi = static_cast<__nv_bfloat16*>(Zout) + i * ldz;
    __nv_bfloat16* zlo = zhi + N * ldz;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        hi[k] = __float2bfloat16_rn(z[h][k]);
        lo[k] = __float2bfloat16_rn(z[h][k] - __bfloat162float(hi[k]));
      }
      *reinterpret_cast<uint2*>(zhi + h * C + c0) = *reinterpret_cast<const uint2*>(hi);
      if (OUT == 2) *reinterpret_cast<uint2*>(zlo + h * C + c0) = *reinterpret_cast<const uint2*>(lo);
    }
  }
  }   // tile loop
}

template <int NPW, int OUT, int SLOTS, bool HAS_MAP>
static int launch_ps3(cudaStream_t st, const float* x, int64_t ldx, int64_t N, const int32_t* rowptr, const int32_t* nbr,
                      const int32_t* row_map, const double* P, const float* c, void* Z, int64_t ldz) {
  constexpr int C = 128 / NPW;
  const size_t smem = (size_t)8 * NPW * SLOTS * (C + 12) * sizeof(float);
  static int resident = 0;   // CTAs that fit on the device at once (per template instance)
  if (resident == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    GEOBI_CUDA_OK(cudaGetDevice(&dev));
    GEOBI_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GEOBI_CUDA_OK(cudaFuncSetAttribute(feast_aggregate_ps_kernel<NPW, OUT, SLOTS, HAS_MAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GEOBI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, feast_aggregate_ps_kernel<NPW, OUT, SLOTS, HAS_MAP>, 256, smem));
    resident = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int64_t want = cdiv(N, 8 * NPW);
  const unsigned grid = (unsigned)(want < resident ? want : resident);
  feast_aggregate_ps_kernel<NPW, OUT, SLOTS, HAS_MAP><<<grid, 256, smem, st>>>(x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz);
  return GEOBI_OK;
}
// the row_map switch is a template parameter: a predicated-off row_map[j] still waits for the index load it would consume
template <int NPW, int OUT, int SLOTS>
static int launch_ps2(cudaStream_t st, const float* x, int64_t ldx, int64_t N, const int32_t* rowptr, const int32_t* nbr,
                      const int32_t* row_map, const double* P, const float* c, void* Z, int64_t ldz) {
  return row_map ? launch_ps3<NPW, OUT, SLOTS, true>(st, x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz)
                 : launch_ps3<NPW, OUT, SLOTS, false>(st, x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz);
}
template <int NPW, int SLOTS>
static int launch_ps(int out_mode, cudaStream_t st, const float* x, int64_t ldx, int64_t N, const int32_t* rowptr, const int32_t* nbr,
                     const int32_t* row_map, const double* P, const float* c, void* Z, int64_t ldz) {
  if (out_mode == 0) return launch_ps2<NPW, 0, SLOTS>(st, x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz);
  if (out_mode == 1) return launch_ps2<NPW, 1, SLOTS>(st, x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz);
  return launch_ps2<NPW, 2, SLOTS>(st, x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz);
}

// Wt[(h*C_in + c), o] = W[(h*C_out + o), c]
__global__ void feast_transpose_w_kernel(const float* __restrict__ W, int C_in, int C_out, float* __restrict__ Wt) {
  const int total = H * C_in * C_out;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int o = t % C_out;
    const int k = t / C_out;
    const int h = k / C_in, c = k - h * C_in;
    Wt[t] = W[(int64_t)(h * C_out + o) * C_in + c];
  }
}

// ------------------------------------------------------------------------------ out = act(A . Bt + bias)
// Plain smem-tiled fp32 GEMM, 64 x BN tile, 4x4 per thread.  A [M,K] row-major, Bt [K,Ntot].
template <int BN>
__global__ void __launch_bounds__(16 * (BN / 4)) gemm_bias_act_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bt,
                                                                       int Ntot, const float* __restrict__ bias, int64_t M, int K,
                                                                       float slope, float* __restrict__ out, int64_t ldo) {
  constexpr int BM = 64, BK = 16, TM = 4, TN = 4, NT = 16 * (BN / 4);
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int idx = tid; idx < BM * BK; idx += NT) {
      const int r = idx / BK, kk = idx - r * BK;
      const int64_t m = m0 + r;
      As[kk][r] = (m < M && k0 + kk < K) ? A[m * lda + k0 + kk] : 0.f;
    }
    for (int idx = tid; idx < BK * BN; idx += NT) {
      const int kk = idx / BN, n = idx - kk * BN;
      Bs[kk][n] = (k0 + kk < K) ? Bt[(int64_t)(k0 + kk) * Ntot + n0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      const float a[TM] = {a4.x, a4.y, a4.z, a4.w};
      const float b[TN] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int p = 0; p < TM; ++p)
#pragma unroll
        for (int q = 0; q < TN; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < TM; ++p) {
    const int64_t m = m0 + ty * TM + p;
    if (m >= M) continue;
#pragma unroll
    for (int q = 0; q < TN; ++q) {
      const int n = n0 + tx * TN + q;
      float v = acc[p][q] + bias[n];
      v = v > 0.f ? v : v * slope;
      out[m * ldo + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------ fused FC head
// y = W2 . lrelu(W1 . f + b1) + b2 (+ epilogue).  64 nodes per CTA, 4 hidden sub-ranges per node;
// W1 streams through smem in 128-row chunks, the [N,hidden] activation never leaves registers.
template <int CIN>
__global__ void __launch_bounds__(256) fc_head_kernel(const float* __restrict__ f, int64_t ldf, int64_t N, const float* __restrict__ W1,
                                                      const float* __restrict__ b1, int hidden, const float* __restrict__ W2,
                                                      const float* __restrict__ b2, int CO, int epilogue, const float* __restrict__ res,
                                                      int64_t ldres, const float* __restrict__ res2, int64_t ldres2, float* __restrict__ out,
                                                      int64_t ldo) {
  constexpr int NODES = 64, HC = 128;
  __shared__ __align__(16) float W1s[HC][CIN];
  __shared__ float b1s[HC];
  __shared__ float W2s[4][HC];
  __shared__ float part[4][NODES][4];
  const int tid = threadIdx.x;
  const int node = tid & (NODES - 1), sub = tid / NODES;
  const int64_t n = (int64_t)blockIdx.x * NODES + node;
  float fr[CIN];
#pragma unroll
  for (int k = 0; k < CIN; ++k) fr[k] = n < N ? f[n * ldf + k] : 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int h0 = 0; h0 < hidden; h0 += HC) {
    __syncthreads();
    for (int idx = tid; idx < HC * CIN; idx += 256) (&W1s[0][0])[idx] = W1[(int64_t)h0 * CIN + idx];
    for (int idx = tid; idx < HC; idx += 256) b1s[idx] = b1[h0 + idx];
    for (int idx = tid; idx < 4 * HC; idx += 256) {
      const int c = idx / HC, j = idx - c * HC;
      W2s[c][j] = c < CO ? W2[(int64_t)c * hidden + h0 + j] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < HC / 4; ++jj) {
      const int j = sub * (HC / 4) + jj;
      float hsum = b1s[j];
#pragma unroll
      for (int k = 0; k < CIN; k += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(&W1s[j][k]);
        hsum = fmaf(w4.x, fr[k], hsum);
        hsum = fmaf(w4.y, fr[k + 1], hsum);
        hsum = fmaf(w4.z, fr[k + 2], hsum);
        hsum = fmaf(w4.w, fr[k + 3], hsum);
      }
      hsum = hsum > 0.f ? hsum : 0.2f * hsum;
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = fmaf(W2s[c][j], hsum, acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) part[sub][node][c] = acc[c];
  __syncthreads();
  if (sub != 0 || n >= N) return;
  float y[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) y[c] = c < CO ? ((part[0][node][c] + part[1][node][c]) + (part[2][node][c] + part[3][node][c])) + b2[c] : 0.f;
  int co = CO;
  if (epilogue == 2) {
    if (CO == 1) {
      const float s = y[0];
      for (int c = 0; c < 3; ++c) y[c] = s * res2[n * ldres2 + c];
      co = 3;
    } else {
      for (int c = 0; c < co; ++c) y[c] *= res2[n * ldres2 + c];
    }
  }
  if (epilogue == 1 || epilogue == 2)
    for (int c = 0; c < co; ++c) y[c] += res[n * ldres + c];
  if (epilogue == 3) {
    float s = 0.f;
    for (int c = 0; c < co; ++c) s += y[c] * y[c];
    const float d = fmaxf(sqrtf(s), 1e-12f);
    for (int c = 0; c < co; ++c) y[c] /= d;
  }
  for (int c = 0; c < co; ++c) out[n * ldo + c] = y[c];
}

struct FeastWs {
  double* P;
  float* Z;
  float* Wt;
};
template <class C>
static void carve_feast(C& c, int64_t N, int c_in, int c_out, FeastWs* out) {
  double* P = c.template take<double>((size_t)N * H);
  float* Z = c.template take<float>((size_t)N * H * c_in);
  float* Wt = c.template take<float>((size_t)H * c_in * c_out);
  if (out) *out = FeastWs{P, Z, Wt};
}
struct NullCarverF {
  Sizer s;
  template <typename T>
  T* take(size_t n) { s.take<T>(n); return nullptr; }
};

// P = X U^T (fp64) then Z[i, h*C_in + c] (row stride ldz) = mean over N(i)+{i} of q_ijh x_j[c].
// out_mode 0: fp32 Z;  1: bf16 hi plane;  2: bf16 hi | lo planes.
template <int CPL, bool VEC>
static void launch_aggregate(int out_mode, unsigned blocks, cudaStream_t st, const float* x, int64_t ldx, int64_t N, int c_in,
                             const int32_t* rowptr, const int32_t* nbr, const double* P, const float* c, void* Z, int64_t ldz) {
  if (out_mode == 0) feast_aggregate_kernel<CPL, 0, VEC><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
  else if (out_mode == 1) feast_aggregate_kernel<CPL, 1, VEC><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
  else feast_aggregate_kernel<CPL, 2, VEC><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
}

static int launch_project(const float* x, int64_t ldx, int64_t N, int c_in, const float* U, double* P, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    GEOBI_CUDA_OK(cudaFuncSetAttribute(feast_project_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)proj_smem_bytes(128)));
    GEOBI_CUDA_OK(cudaFuncSetAttribute(feast_project_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)proj_smem_bytes(32)));
    GEOBI_CUDA_OK(cudaFuncSetAttribute(feast_project_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)proj_smem_bytes(64)));
    GEOBI_CUDA_OK(cudaFuncSetAttribute(feast_project_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)proj_smem_bytes(128)));
    attr_set = true;
  }
  if (N == 0) return GEOBI_OK;
  const unsigned grid = (unsigned)cdiv(N, PROJ_NODES);
  const size_t smem = proj_smem_bytes(c_in);
  if (c_in == 32) feast_project_kernel<32><<<grid, PROJ_THREADS, smem, st>>>(x, ldx, N, c_in, U, P);
  else if (c_in == 64) feast_project_kernel<64><<<grid, PROJ_THREADS, smem, st>>>(x, ldx, N, c_in, U, P);
  else if (c_in == 128) feast_project_kernel<128><<<grid, PROJ_THREADS, smem, st>>>(x, ldx, N, c_in, U, P);
  else feast_project_kernel<0><<<grid, PROJ_THREADS, smem, st>>>(x, ldx, N, c_in, U, P);
  GEOBI_LAUNCH_OK("feast_project");
  return GEOBI_OK;
}

// 4 / 2: channels per lane of the small-C kernel that will run for this layer; 0: another kernel
static int small_c_variant(int c_in, int64_t ldx, int64_t ldz, const float* x, bool row_map) {
  if (row_map || getenv("GEOBI_NO_SMALLC") != nullptr) return 0;
  if (c_in <= 12 && c_in % 4 == 0 && ldx % 4 == 0 && ldz % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) return 4;
  if (c_in <= 6 && c_in % 2 == 0 && ldx % 2 == 0 && ldz % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0) return 2;
  return 0;
}
// true when the aggregation kernel itself zero-fills the K padding of the bf16 planes (no memset needed)
bool feast_aggregate_fills_padding(int c_in, int64_t ldx, int64_t ldz, const float* x, bool row_map) {
  return small_c_variant(c_in, ldx, ldz, x, row_map) != 0;
}

int feast_project_only(const float* x, int64_t ldx, int64_t N, int c_in, const float* U, double* P, cudaStream_t st) {
  return launch_project(x, ldx, N, c_in, U, P, st);
}

int feast_project_and_aggregate(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                const int32_t* row_map, int64_t n_src, const float* U, const float* c, double* P, void* Z, int64_t ldz,
                                int out_mode, cudaStream_t st) {
  {
    const int prc = launch_project(x, ldx, n_src, c_in, U, P, st);
    if (prc) return prc;
  }
  const unsigned ab = (unsigned)cdiv(N, 8);
  const int cpl = c_in <= 32 ? 1 : (c_in <= 64 ? 2 : 4);
  // vector path: every lane's CPL-channel group is whole and 4*CPL-byte aligned in x (and in Z for the packed stores)
  const bool vec_ok = (c_in % cpl == 0) && (ldx % cpl == 0) && ((reinterpret_cast<uintptr_t>(x) % (4 * cpl)) == 0) && (ldz % cpl == 0);
  GEOBI_REQUIRE(n_src * ldx < ((int64_t)1 << 32), "feast_fwd: rows * ldx must stay below 2^32 elements (32-bit gather offsets)");
  // staged (cp.async) variant: whole 16-byte pieces, 32 % (C/4) == 0, lanes own whole CPL groups
  const bool staged_ok = (c_in == 32 || c_in == 64 || c_in == 128) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                         (ldz % cpl == 0);
  GEOBI_REQUIRE(!row_map || (staged_ok && ldz % 4 == 0 && getenv("GEOBI_AGG_DIRECT") == nullptr),
                "feast_fwd: row_map needs C_in in {32,64,128} and 16-byte aligned rows");
  if (staged_ok && (ldz % 4 == 0)) {
    // C=32: 4 nodes/warp, 8 slots each; C=64: 2 nodes/warp, 16 slots; C=128: 1 node/warp, 16 slots (64 KB of rows per CTA)
    const bool direct = getenv("GEOBI_AGG_DIRECT") != nullptr;   // A/B switch for profiling
    int rc = GEOBI_OK;
    if (direct) {
      if (c_in == 32) launch_packed<4>(out_mode, st, x, ldx, N, rowptr, nbr, P, c, Z, ldz);
      else if (c_in == 64) launch_packed<2>(out_mode, st, x, ldx, N, rowptr, nbr, P, c, Z, ldz);
      else launch_packed<1>(out_mode, st, x, ldx, N, rowptr, nbr, P, c, Z, ldz);
    } else if (c_in == 32) rc = launch_ps<4, 8>(out_mode, st, x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz);
    else if (c_in == 64) rc = launch_ps<2, 16>(out_mode, st, x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz);
    else rc = launch_ps<1, 16>(out_mode, st, x, ldx, N, rowptr, nbr, row_map, P, c, Z, ldz);
    if (rc) return rc;
  } else if (small_c_variant(c_in, ldx, ldz, x, row_map != nullptr) == 4)
    launch_small<4>(out_mode, st, x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);     // first facet layer (C_in = 12)
  else if (small_c_variant(c_in, ldx, ldz, x, row_map != nullptr) == 2)
    launch_small<2>(out_mode, st, x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);     // first vertex layer (C_in = 6)
  else if (cpl == 1) launch_aggregate<1, true>(out_mode, ab, st, x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);   // scalar loads are always aligned
  else if (cpl == 2 && vec_ok) launch_aggregate<2, true>(out_mode, ab, st, x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
  else if (cpl == 2) launch_aggregate<2, false>(out_mode, ab, st, x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
  else if (vec_ok) launch_aggregate<4, true>(out_mode, ab, st, x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
  else launch_aggregate<4, false>(out_mode, ab, st, x, ldx, N, c_in, rowptr, nbr, P, c, Z, ldz);
  GEOBI_LAUNCH_OK("feast_aggregate");
  return GEOBI_OK;
}

int feast_fwd_tc(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr, const int32_t* row_map,
                 int64_t n_src, const float* W, const float* U, const float* c, const float* bias, int c_out, float act_slope, float* out,
                 int64_t ldo, int passes, void* ws, size_t ws_bytes, cudaStream_t st);  // feast_tc.cu
size_t feast_fwd_tc_ws_bytes(int64_t N, int c_in, int c_out);
bool feast_fused_supported(int c_in, int c_out, int64_t ldx, int64_t ldo, const float* x, int64_t N);   // feast_fused.cu
int feast_fwd_fused(const float* x, int64_t ldx, int64_t N, const int32_t* rowptr, const int32_t* nbr, const int32_t* row_map, int64_t n_src,
                    const float* W, const float* U, const float* c, const float* bias, float act_slope, float* out, int64_t ldo, bool reuse_ws,
                    void* ws, size_t ws_bytes, cudaStream_t st);
// feast_tcagg.cu: the same layer with the aggregation itself on tcgen05 (round 2).  Opt-in (GEOBI_TCAGG=1) until it beats the
// FP32-pipe kernel above: 0.544 ms vs 0.527 ms per launch at N = 512 000 (profiles/r02_NOTES.md)
bool feast_tcagg_supported(int c_in, int c_out, int64_t ldx, int64_t ldo, const float* x, int64_t n_src);
size_t feast_fwd_tcagg_ws_bytes(int64_t n_src);
int feast_fwd_tcagg(const float* x, int64_t ldx, int64_t N, const int32_t* rowptr, const int32_t* nbr, const int32_t* row_map, int64_t n_src,
                    const float* W, const float* U, const float* c, const float* bias, float act_slope, float* out, int64_t ldo, bool reuse_ws,
                    void* ws, size_t ws_bytes, cudaStream_t st);
int fc_head_fwd_tc(const float* f, int64_t ldf, int64_t n, int c_in, const float* W1, const float* b1, int hidden, const float* W2,
                   const float* b2, int c_out, int epilogue, const float* res, int64_t ldres, const float* res2, int64_t ldres2, float* out,
                   int64_t ldo, void* ws, size_t ws_bytes, cudaStream_t st);
size_t fc_head_tc_ws_bytes(int hidden);

}  // namespace geobi

using namespace geobi;

extern "C" size_t geobi_feast_fwd_ws_bytes(int64_t n_nodes, int c_in, int c_out, int precision) {
  precision &= ~GEOBI_FEAST_REUSE_WS;
  if (precision != GEOBI_PREC_FP32) {
    const size_t a = feast_fwd_tc_ws_bytes(n_nodes, c_in, c_out);
    const size_t b = (c_in == 64 && c_out == 32) ? feast_fwd_tcagg_ws_bytes(n_nodes) : 0;
    return a > b ? a : b;
  }
  NullCarverF c;
  carve_feast(c, n_nodes, c_in, c_out, nullptr);
  return c.s.total();
}

extern "C" int geobi_feast_fwd(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                               const int32_t* row_map, int64_t n_src, const float* W, const float* U, const float* c, const float* bias,
                               int c_out, float act_slope, float* out, int64_t ldo, int precision, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!row_map) n_src = N;
  GEOBI_REQUIRE(n_src >= 0 && (N == 0 || n_src > 0), "feast_fwd: n_src must be the row count of x when row_map is given");
  GEOBI_REQUIRE(x && rowptr && W && U && c && bias && out && N >= 0, "feast_fwd: null argument");
  GEOBI_REQUIRE(c_in >= 1 && c_in <= 128, "feast_fwd: C_in must be in 1..128 (got %d)", c_in);
  GEOBI_REQUIRE(c_out == 32 || c_out == 64 || c_out == 128, "feast_fwd: C_out must be 32, 64 or 128 (got %d)", c_out);
  GEOBI_REQUIRE(ldx >= c_in && ldo >= c_out, "feast_fwd: leading dimension smaller than channel count");
  const bool reuse_ws = (precision & GEOBI_FEAST_REUSE_WS) != 0;
  precision &= ~GEOBI_FEAST_REUSE_WS;
  GEOBI_REQUIRE(precision >= GEOBI_PREC_FP32 && precision <= GEOBI_PREC_BF16X3, "feast_fwd: unknown precision %d", precision);
  if (N == 0) return GEOBI_OK;
  const bool fused = precision == GEOBI_PREC_BF16X3 && feast_fused_supported(c_in, c_out, ldx, ldo, x, n_src) && getenv("GEOBI_NO_FUSED") == nullptr;
  GEOBI_REQUIRE(!reuse_ws || fused, "feast_fwd: GEOBI_FEAST_REUSE_WS is only defined for the fused 64->32 bf16x3 kernel");
  const char* tcagg_env = getenv("GEOBI_TCAGG");
  if (fused && tcagg_env != nullptr && tcagg_env[0] == '1' && feast_tcagg_supported(c_in, c_out, ldx, ldo, x, n_src))
    return feast_fwd_tcagg(x, ldx, N, rowptr, nbr, row_map, n_src, W, U, c, bias, act_slope, out, ldo, reuse_ws, ws, ws_bytes, st);
  if (fused) return feast_fwd_fused(x, ldx, N, rowptr, nbr, row_map, n_src, W, U, c, bias, act_slope, out, ldo, reuse_ws, ws, ws_bytes, st);
  if (precision != GEOBI_PREC_FP32)
    return feast_fwd_tc(x, ldx, N, c_in, rowptr, nbr, row_map, n_src, W, U, c, bias, c_out, act_slope, out, ldo,
                        precision == GEOBI_PREC_BF16X3 ? 3 : 1, ws, ws_bytes, st);
  if (!ws || ws_bytes < geobi_feast_fwd_ws_bytes(N, c_in, c_out, precision)) {
    set_error("feast_fwd: workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  Carver cv(ws, ws_bytes);
  FeastWs Wk;
  carve_feast(cv, N, c_in, c_out, &Wk);
  feast_transpose_w_kernel<<<64, 256, 0, st>>>(W, c_in, c_out, Wk.Wt);
  int rc = feast_project_and_aggregate(x, ldx, N, c_in, rowptr, nbr, row_map, n_src, U, c, Wk.P, Wk.Z, (int64_t)H * c_in, 0, st);
  if (rc) return rc;
  const int K = H * c_in;
  if (c_out == 32) {
    dim3 g((unsigned)cdiv(N, 64), 1);
    gemm_bias_act_kernel<32><<<g, 128, 0, st>>>(Wk.Z, K, Wk.Wt, c_out, bias, N, K, act_slope, out, ldo);
  } else {
    dim3 g((unsigned)cdiv(N, 64), c_out / 64);
    gemm_bias_act_kernel<64><<<g, 256, 0, st>>>(Wk.Z, K, Wk.Wt, c_out, bias, N, K, act_slope, out, ldo);
  }
  GEOBI_LAUNCH_OK("feast_gemm");
  return GEOBI_OK;
}

extern "C" size_t geobi_fc_head_ws_bytes(int hidden) { return fc_head_tc_ws_bytes(hidden); }

extern "C" int geobi_fc_head_fwd(const float* f, int64_t ldf, int64_t n, int c_in, const float* W1, const float* b1, int hidden, const float* W2,
                                 const float* b2, int c_out, int epilogue, const float* res, int64_t ldres, const float* res2, int64_t ldres2,
                                 float* out, int64_t ldo, int precision, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(f && W1 && b1 && W2 && b2 && out && n >= 0, "fc_head_fwd: null argument");
  GEOBI_REQUIRE(c_in == 32 || c_in == 64, "fc_head_fwd: c_in must be 32 or 64 (got %d)", c_in);
  GEOBI_REQUIRE(hidden > 0 && hidden % 128 == 0, "fc_head_fwd: hidden must be a multiple of 128 (got %d)", hidden);
  GEOBI_REQUIRE(c_out >= 1 && c_out <= 4, "fc_head_fwd: c_out must be in 1..4 (got %d)", c_out);
  GEOBI_REQUIRE(epilogue >= 0 && epilogue <= 3, "fc_head_fwd: unknown epilogue %d", epilogue);
  GEOBI_REQUIRE(!(epilogue == 1 || epilogue == 2) || res, "fc_head_fwd: epilogue %d needs res", epilogue);
  GEOBI_REQUIRE(epilogue != 2 || res2, "fc_head_fwd: epilogue 2 needs res2");
  GEOBI_REQUIRE(precision >= GEOBI_PREC_FP32 && precision <= GEOBI_PREC_BF16X3, "fc_head_fwd: unknown precision %d", precision);
  if (n == 0) return GEOBI_OK;
  if (precision != GEOBI_PREC_FP32 && c_in == 32 && c_out <= 3 && hidden % 256 == 0)
    return fc_head_fwd_tc(f, ldf, n, c_in, W1, b1, hidden, W2, b2, c_out, epilogue, res, ldres, res2, ldres2, out, ldo, ws, ws_bytes, st);
  const unsigned blocks = (unsigned)cdiv(n, 64);
  if (c_in == 32)
    fc_head_kernel<32><<<blocks, 256, 0, st>>>(f, ldf, n, W1, b1, hidden, W2, b2, c_out, epilogue, res, ldres, res2, ldres2, out, ldo);
  else
    fc_head_kernel<64><<<blocks, 256, 0, st>>>(f, ldf, n, W1, b1, hidden, W2, b2, c_out, epilogue, res, ldres, res2, ldres2, out, ldo);
  GEOBI_LAUNCH_OK("fc_head");
  return GEOBI_OK;
}
