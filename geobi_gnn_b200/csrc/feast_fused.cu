// Fused FeaSt convolution, C_in = 64 -> C_out = 32 (r_conv3 / r_conv4 of both U-Nets, network.py:267-268: the two
// largest layers of each graph): aggregation + projection + bias + leaky_relu in ONE persistent kernel.
//
//   per tile of 32 target nodes (16 aggregation warps x 2 nodes, lanes own 4 adjacent channels):
//     1. aggregate Z_i[h, :] = 1/d_i sum_j q_ijh x_j in registers (1/d_i folded into the soft assignments),
//     2. write the rows, split x = hi + lo (bf16), straight into the K-major SWIZZLE_128B *B-operand* tiles in shared memory,
//     3. warp 18 issues tcgen05.mma with the weights as the A operand IN TENSOR MEMORY (loaded once per CTA: output o = row o of an
//        M = 64 tile whose rows 32-63 are unused; the bf16 hi plane of W sits in lanes 0-15 and the lo plane in lanes 16-31 of TMEM
//        quadrants 0 and 1), N = 32 nodes, 36 K steps x {W_hi Z_hi, W_hi Z_lo, W_lo Z_hi} = 108 MMAs -> D^T[channel, node] in one of
//        two TMEM accumulators (the hi-plane and lo-plane partial sums in the two lane halves),
//     4. warps 16 and 17 drain that accumulator (partials added by a lane shuffle, +bias, leaky_relu) and write out[node, channel]
//        while the next tile's MMAs fill the other one; the aggregation warps never touch TMEM, they only wait for the MMAs of
//        tile t-1 before overwriting the operand tiles with tile t.
//
// Round 2: the weights moved from shared memory (108 SS MMAs per tile, each re-reading a 2 KB weight block: 81 of the kernel's
// 165 shared-memory wavefronts per node, ncu: data pipe 87 % busy) to tensor memory (108 TS MMAs that read only the 1 KB Z block:
// 27 wavefronts per node).  Stacking [W_hi; W_lo] on M (72 MMAs) needs four drain warps = 672 threads, which drops the register
// cap from 96 to 80 and spills the aggregation warps' accumulators.
//
// Z (2.3 KB per node) never reaches HBM: per node the kernel reads 256 B of x per gathered row (L2) + 72 B of P, and writes
// 128 B.  The unfused path writes and re-reads 2 x 2.3 KB per node (ncu: 1.12 GB written by the aggregation of one layer).
#include "tc.cuh"

namespace geobi {
namespace fused {

using namespace tc;

#ifndef PREFETCH_MODE
#define PREFETCH_MODE 3   // 3: P rows of the next tile into L1 at the end of a tile (A/B: -2 %); x-row prefetches (1, 2) lost
#endif
constexpr int C_IN = 64, C_OUT = 32;
#ifndef PAIR_SETS
#define PAIR_SETS 2   // 3 = gathers issued two pair-consumptions ahead: no gain (0.526 vs 0.527 ms; 96 registers, spills)
#endif
#ifndef FUSED_WARPS
#define FUSED_WARPS 16
#endif
constexpr int WARPS = FUSED_WARPS;   // aggregation warps, 2 nodes each (a multiple of 4: the drain warps must be warps 0, 1 mod 4)
constexpr int NT = 2 * WARPS;        // nodes per tile = MMA N
static_assert(WARPS % 4 == 0 && NT % 8 == 0 && NT <= 64, "tile shape");
constexpr int EPI_WARPS = 2;      // warps 16, 17 drain TMEM quadrants 0, 1 (a warp reaches lanes 32*(warp%4)..+31)
constexpr int MMA_WARP = WARPS + EPI_WARPS;   // warp 18 issues the MMAs
constexpr int THREADS = (WARPS + EPI_WARPS + 1) * 32;
constexpr int KB = H;             // one 64-wide K block per head
constexpr int KSTEPS = KB * 4;    // K = 16 steps
constexpr int ZTILE_BYTES = NT * 128;  // Z: [NT node rows x 128 B] of one K block
constexpr int ZPLANE_BYTES = KB * ZTILE_BYTES;
// tensor memory: weights in columns [0, 288) (hi plane in lanes 0-15, lo plane in lanes 16-31), then the two accumulators (NT columns each)
constexpr uint32_t COL_W = 0, COL_D = KSTEPS * 8, TMEM_COLS = 512, HALF = 16u << 16;
static_assert(COL_D + 2 * NT <= TMEM_COLS, "tensor memory budget");
constexpr int QROW = 12;          // floats per soft-assignment row (9 used; 48-byte rows keep the float4 reads aligned)
constexpr int QPAD = 4;           // floats between the two nodes' halves: without it their rows sit 768 B apart = on the same banks, and every
                                  // broadcast read of the pair loop (one address per half warp) took two shared-memory wavefronts
constexpr int SCRATCH = (32 * QROW + QPAD) * 4;   // per-warp scratch: the soft-assignment rows of the chunk's 32 slots
constexpr int SMEM_BYTES = 2 * ZPLANE_BYTES + WARPS * SCRATCH + WARPS * 32 * 4 + 64 + 1024;

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
// named barrier 1: the 16 aggregation warps arrive (non-blocking) when their rows are in shared memory; the MMA warp sleeps on it
__device__ __forceinline__ void z_ready_arrive() { asm volatile("bar.arrive 1, %0;" ::"n"((WARPS + 1) * 32) : "memory"); }
__device__ __forceinline__ void z_ready_wait() { asm volatile("bar.sync 1, %0;" ::"n"((WARPS + 1) * 32) : "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

#ifdef EXP_TIMELINE
__device__ unsigned long long g_tl[64 * 8];
__device__ unsigned long long g_tw[64 * 16 * 2];
__device__ unsigned long long g_ts[64 * 2 * 8];
#define TS(slot) do { if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15) && tile - t_begin >= 20 && tile - t_begin < 84) g_ts[((tile - t_begin - 20) * 2 + (warp == 15)) * 8 + (slot)] = now_ns(); } while (0)
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL(slot) do { if (blockIdx.x == 0 && lane == 0 && tile - t_begin >= 20 && tile - t_begin < 84) g_tl[(tile - t_begin - 20) * 8 + (slot)] = now_ns(); } while (0)
#else
#define TL(slot) do { } while (0)
#define TS(slot) do { } while (0)
#endif

// HAS_MAP: compile-time row_map switch - a predicated-off `row_map[j]` address computation still waits on the scoreboard of
// the index load it would consume (5 % of the stall samples of the map-less instance before the split)
template <bool HAS_MAP>
__global__ void __launch_bounds__(THREADS, 1) feast_fused_64_32_kernel(const float* __restrict__ x, int64_t ldx, int64_t N,
                                                                       const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                                       const int* __restrict__ row_map,
                                                                       const double* __restrict__ P, const float* __restrict__ cvec,
                                                                       const uint32_t* __restrict__ Wp /* [36 K steps][64 stacked rows][8] bf16 pairs */,
                                                                       const float* __restrict__ bias, float slope, float* __restrict__ out,
                                                                       int64_t ldo) {
  extern __shared__ uint8_t smem_raw[];
  // mma_done[b]: the MMAs of a tile whose accumulator is TMEM buffer b have completed (operand tiles reusable, accumulator
  // readable); acc_free[b]: both drain warps have read buffer b.  Two accumulators, so the MMAs of tile t+1 run while tile t
  // is drained; each barrier completes once every second tile: the parity of local tile k is (k >> 1) & 1.
  __shared__ __align__(8) uint64_t mma_done[2];
  __shared__ __align__(8) uint64_t acc_free[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* z_hi = sm;
  uint8_t* z_lo = sm + ZPLANE_BYTES;
  uint8_t* scratch_all = sm + 2 * ZPLANE_BYTES;
  unsigned* joff_all = reinterpret_cast<unsigned*>(scratch_all + WARPS * SCRATCH);
  float* chs = reinterpret_cast<float*>(joff_all + WARPS * 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- one-time setup: barriers, TMEM, resident weights
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init(&acc_free[0], EPI_WARPS);
    mbar_init(&acc_free[1], EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid < H) chs[tid] = cvec[tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  if (warp >= WARPS && warp < MMA_WARP) {
    // weights -> tensor memory: lane l of quadrant warp q holds output row 16 q + (l & 15) of plane l >> 4 (stacked row 32 plane + output in Wp)
    const int q = warp - WARPS, m = 32 * (lane >> 4) + 16 * q + (lane & 15);
    const uint4* src = reinterpret_cast<const uint4*>(Wp);
#pragma unroll 2
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const uint4 a = __ldg(src + (ks * 64 + m) * 2), b = __ldg(src + (ks * 64 + m) * 2 + 1);
      tmem_st8(tmem_d + COL_W + 8 * ks + ((uint32_t)(q * 32) << 16), a, b);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int64_t n_tiles = (N + NT - 1) / NT;
  const int64_t t_begin = (n_tiles * blockIdx.x) / gridDim.x, t_end = (n_tiles * (blockIdx.x + 1)) / gridDim.x;

  if (warp == MMA_WARP) {
    // ===== MMA warp: per tile, wait for its rows, issue 36 K steps x {Z_hi, Z_lo} against the weights in tensor memory, commit =====
    constexpr uint32_t idesc = make_idesc(64, NT);
    const uint64_t d0 = make_desc(smem_u32(z_hi));
    const uint32_t desc_hi = (uint32_t)(d0 >> 32);
    const uint32_t zh_lo = (uint32_t)d0, zl_lo = (uint32_t)make_desc(smem_u32(z_lo));
    for (int64_t tile = t_begin; tile < t_end; ++tile) {
      const uint32_t k = (uint32_t)(tile - t_begin), buf = k & 1;
      z_ready_wait();
      TL(0);
      if (k >= 2) mbar_wait(&acc_free[buf], ((k >> 1) - 1) & 1);   // tile k-2 has been drained out of this accumulator
      if (elect_one()) {
        tc_fence_after();
        // the B descriptors differ only in the 14-bit start-address field: + (ZTILE_BYTES >> 4) per K block, + 2 per K = 16 step
        const uint32_t acc_h = tmem_d + COL_D + buf * NT, acc_l = acc_h + HALF, a_h = tmem_d + COL_W, a_l = a_h + HALF;
#pragma unroll 4
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint32_t zoff = (uint32_t)(ks >> 2) * (ZTILE_BYTES >> 4) + 2 * (ks & 3);
          mma_ts(acc_h, a_h + 8 * ks, zh_lo + zoff, desc_hi, idesc, ks > 0);
#ifndef EXP_ONEPASS
          mma_ts(acc_h, a_h + 8 * ks, zl_lo + zoff, desc_hi, idesc, 1u);
          mma_ts(acc_l, a_l + 8 * ks, zh_lo + zoff, desc_hi, idesc, ks > 0);
#endif
        }
        mma_commit(&mma_done[buf]);
      }
      __syncwarp();
      TL(1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem_dealloc(tmem_d, TMEM_COLS);
    return;
  }
  if (warp >= WARPS) {
    // ===== drain warps: accumulator -> hi-plane + lo-plane partial -> +bias, leaky_relu -> out[node, channel] =====
    // quadrant q: lane l < 16 holds W_hi . Z of output 16 q + l, lane l + 16 the W_lo . Z_hi part of the same output
    const int quad = warp - WARPS;
    const int o = 16 * quad + (lane & 15);
    const float my_bias = bias[o];
    for (int64_t tile = t_begin; tile < t_end; ++tile) {
      const uint32_t k = (uint32_t)(tile - t_begin), buf = k & 1;
      mbar_wait(&mma_done[buf], (k >> 1) & 1);
      tc_fence_after();
      if (quad == 0) TL(2);
      float v[NT];
      {
        const uint32_t ta = tmem_d + COL_D + buf * NT + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c = 0; c + 32 <= NT; c += 32) tmem_ld32(ta + c, v + c);
#pragma unroll
        for (int c = NT & ~31; c < NT; c += 8) tmem_ld8(ta + c, v + c);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free[buf]);  // values are in registers: the accumulator may be overwritten
#pragma unroll
      for (int col = 0; col < NT; ++col) v[col] += __shfl_xor_sync(0xffffffffu, v[col], 16);
      if (lane < 16) {
        const int64_t n0 = tile * NT;
#pragma unroll
        for (int col = 0; col < NT; ++col) {
#ifdef EXP_NOSTORE
          if (n0 + col < N && v[col] == 123.456f) {
#else
          if (n0 + col < N) {
#endif
            float r = v[col] + my_bias;
            r = r > 0.f ? r : r * slope;
            out[(n0 + col) * ldo + o] = r;
          }
        }
      }
      if (quad == 0) TL(3);
    }
    tc_fence_before();
    __syncthreads();
    return;
  }

  // ===== aggregation warps =====
  constexpr int LPN = 16;
  const int g = lane / LPN, sl = lane % LPN, c0 = sl * 4;
  const int row = warp * 2 + g;              // node slot inside the tile = row of the B operand
  const unsigned ldx32 = (unsigned)ldx;
  const float* xl = x + c0;                  // this lane's 4 channels
  asm volatile("" : "+l"(xl));               // keep the pointer in registers (ptxas otherwise re-derives it from %tid in the loop)
  float* qs = reinterpret_cast<float*>(scratch_all + warp * SCRATCH);
  unsigned* joffs = joff_all + warp * 32;
  // Index prefetch pipeline (breaks the rowptr -> nbr -> data dependency chain across tiles):
  //   iteration t holds (b, total, j) of tile t, loads nbr of tile t+1 with the rowptr loaded one iteration earlier,
  //   and loads rowptr of tile t+2.
  auto node_of = [&](int64_t tile) -> int64_t {
    const int64_t r = tile * NT + row;
    return r < N ? r : N - 1;          // slots past the end shadow the last node (their rows are zeroed)
  };
  auto load_rowptr = [&](int64_t tile, int& b_, int& total_) {
    if (tile < t_end) {
      const int64_t i_ = node_of(tile);
      b_ = rowptr[i_];
      total_ = rowptr[i_ + 1] - b_ + 1;
    } else {
      b_ = 0;
      total_ = 1;
    }
  };
  // row_map (PoolingLayer.unpooling fused into the conv): x and P rows of node v live at row_map[v]
  auto load_first_j = [&](int64_t tile, int b_, int total_) -> int {
    int j_ = (int)node_of(tile < t_end ? tile : t_begin);
    if (tile < t_end && sl > 0 && sl < total_) j_ = nbr[b_ + sl - 1];
    return HAS_MAP ? row_map[j_] : j_;
  };
  int b_cur, total_cur, b_nxt, total_nxt, b_nx2 = 0, total_nx2 = 1;
  load_rowptr(t_begin, b_cur, total_cur);
  load_rowptr(t_begin + 1, b_nxt, total_nxt);
  int j_cur = load_first_j(t_begin, b_cur, total_cur);
  int j_nxt = 0;

  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    // ---------------- 1. aggregation of this warp's two nodes (registers only)
    if (warp == 0) TL(4);
    TS(0);
#ifdef EXP_TIMELINE
    if (blockIdx.x == 0 && lane == 0 && tile - t_begin >= 20 && tile - t_begin < 84) g_tw[((tile - t_begin - 20) * 16 + warp) * 2] = now_ns();
#endif
    const int64_t i_raw = tile * NT + row;
    const bool live = i_raw < N;
    const int64_t i = live ? i_raw : N - 1;
    const int b = b_cur;
    const int total = total_cur;
    // prefetch: indices of the next tiles (consumed one / two iterations from now)
    j_nxt = load_first_j(tile + 1, b_nxt, total_nxt);
    load_rowptr(tile + 2, b_nx2, total_nx2);
    const int maxtotal = max(total, __shfl_xor_sync(0xffffffffu, total, 16));
    const int i_src = HAS_MAP ? row_map[i] : (int)i;
    const float rcnt = live ? 1.0f / (float)total : 0.f;   // mean over N(i)+{i}, folded into q; dead rows become zeros
    // accumulators: acc[h][0] = channels (c0, c0+1), acc[h][1] = (c0+2, c0+3) of head h, as packed fp32 pairs (FFMA2)
    unsigned long long acc[H][2];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h][0] = acc[h][1] = 0ull;
    for (int s0 = 0; s0 < maxtotal; s0 += LPN) {
      const int s = s0 + sl;
      int j = i_src;
      if (s0 == 0) j = j_cur;                      // first chunk: prefetched one tile ago
      else if (s < total) {
        j = nbr[b + s - 1];
        if (HAS_MAP) j = row_map[j];
      }
      const int cnt = min(LPN, maxtotal - s0);
#ifdef EXP_NOGATHER
      j = i_src;
#endif
      const unsigned joff = (unsigned)j * ldx32;
      // the first pair of rows does not depend on the soft assignments: get it in flight together with the P rows
      const unsigned o0 = __shfl_sync(0xffffffffu, joff, 0, LPN);
      const unsigned o1 = __shfl_sync(0xffffffffu, joff, 1, LPN);
      ulonglong2 xa = __ldg(reinterpret_cast<const ulonglong2*>(xl + o0));
      ulonglong2 xb = __ldg(reinterpret_cast<const ulonglong2*>(xl + o1));
#if PAIR_SETS == 3
      const unsigned o2 = __shfl_sync(0xffffffffu, joff, 2, LPN);   // slots past the row's end hold the node's own (valid) row
      const unsigned o3 = __shfl_sync(0xffffffffu, joff, 3, LPN);
      ulonglong2 xc = __ldg(reinterpret_cast<const ulonglong2*>(xl + o2));
      ulonglong2 xd = __ldg(reinterpret_cast<const ulonglong2*>(xl + o3));
#endif
#if PREFETCH_MODE == 1
      if (sl >= 2 && s < total) {
        prefetch_l1(x + joff);
        prefetch_l1(x + joff + 32);
      }
#endif
      joffs[lane] = joff;
      if (s0 == 0) TS(1);
      float l[H];
#ifdef EXP_NOSOFTMAX
      if (false) {
#else
      if (s < total) {
#endif
        const double* Pj = P + (int64_t)j * H;
        const double* Pi = P + (int64_t)i_src * H;
        float m = -INFINITY;
#pragma unroll
        for (int h = 0; h < H; ++h) {
#ifdef EXP_NOPLOAD
          l[h] = (float)(__longlong_as_double((long long)(j + h)) - __longlong_as_double((long long)(i_src + h))) + chs[h];
#elif defined(EXP_NOEXP)
          l[h] = p_diff(Pj[h], Pi[h]) * 1e-30f + 0.111f;
#else
          l[h] = p_diff(Pj[h], Pi[h]) + chs[h];
#endif
          m = fmaxf(m, l[h]);
        }
        float sum = 0.f;
#ifndef EXP_NOEXP
#pragma unroll
        for (int h = 0; h < H; ++h) {
          l[h] = __expf(l[h] - m);
          sum += l[h];
        }
        const float inv = rcnt / sum;
#else
        const float inv = rcnt;
#endif
#pragma unroll
        for (int h = 0; h < H; ++h) l[h] *= inv;
      } else {
#pragma unroll
        for (int h = 0; h < H; ++h) l[h] = 0.f;
      }
#ifdef EXP_NOSOFTMAX
      if (s < total) {
#pragma unroll
        for (int h = 0; h < H; ++h) l[h] = rcnt * (1.0f / 9.0f);
      }
#endif
      float4* q4 = reinterpret_cast<float4*>(qs + lane * QROW + g * QPAD);
      q4[0] = make_float4(l[0], l[1], l[2], l[3]);
      q4[1] = make_float4(l[4], l[5], l[6], l[7]);
      qs[lane * QROW + g * QPAD + 8] = l[8];
      if (s0 == 0) TS(2);
      __syncwarp();
      const float* qbase = qs + g * (LPN * QROW + QPAD);
      const unsigned* jbase = joffs + g * LPN;
      // slots >= total hold zero assignments and a valid row (the node's own), so pairs need no tail handling
      auto consume = [&](const ulonglong2& va, const ulonglong2& vb, int t) {
        const float* qa = qbase + t * QROW;
        const float4 qa0 = *reinterpret_cast<const float4*>(qa);
        const float4 qa1 = *reinterpret_cast<const float4*>(qa + 4);
        const float qa8 = qa[8];
        const float4 qb0 = *reinterpret_cast<const float4*>(qa + QROW);
        const float4 qb1 = *reinterpret_cast<const float4*>(qa + QROW + 4);
        const float qb8 = qa[QROW + 8];
        const float qav[H] = {qa0.x, qa0.y, qa0.z, qa0.w, qa1.x, qa1.y, qa1.z, qa1.w, qa8};
        const float qbv[H] = {qb0.x, qb0.y, qb0.z, qb0.w, qb1.x, qb1.y, qb1.z, qb1.w, qb8};
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const unsigned long long qq = pack2(qav[h], qav[h]);
          acc[h][0] = ffma2(va.x, qq, acc[h][0]);
          acc[h][1] = ffma2(va.y, qq, acc[h][1]);
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const unsigned long long qq = pack2(qbv[h], qbv[h]);
          acc[h][0] = ffma2(vb.x, qq, acc[h][0]);
          acc[h][1] = ffma2(vb.y, qq, acc[h][1]);
        }
      };
#if PAIR_SETS == 3
      // software pipeline over pairs of rows with THREE register sets: a pair's gathers are issued two pair-consumptions
      // (~160 issue slots) before it is needed - with two sets an L1 miss (11 % of the rows) stalled the warp on L2
      ulonglong2 xe = xa, xf = xb;
#pragma unroll 1
      for (int t = 0; t < cnt; t += 6) {
        if (t + 4 < cnt) {
          const uint2 o = *reinterpret_cast<const uint2*>(jbase + t + 4);
          xe = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.x));
          xf = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.y));
        }
        consume(xa, xb, t);
        if (t + 2 >= cnt) break;
        if (t + 6 < cnt) {
          const uint2 o = *reinterpret_cast<const uint2*>(jbase + t + 6);
          xa = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.x));
          xb = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.y));
        }
        consume(xc, xd, t + 2);
        if (t + 4 >= cnt) break;
        if (t + 8 < cnt) {
          const uint2 o = *reinterpret_cast<const uint2*>(jbase + t + 8);
          xc = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.x));
          xd = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.y));
        }
        consume(xe, xf, t + 4);
      }
#else
      // software pipeline over pairs of rows with two register sets: the next pair's gathers are issued before the
      // current pair is consumed
      ulonglong2 xc = xa, xd = xb;
#pragma unroll 1
      for (int t = 0; t < cnt; t += 4) {
        const bool more = t + 2 < cnt;
        if (more) {
          const uint2 o = *reinterpret_cast<const uint2*>(jbase + t + 2);
          xc = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.x));
          xd = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.y));
        }
        consume(xa, xb, t);
        if (more) {
          if (t + 4 < cnt) {
            const uint2 o = *reinterpret_cast<const uint2*>(jbase + t + 4);
            xa = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.x));
            xb = __ldg(reinterpret_cast<const ulonglong2*>(xl + o.y));
          }
          consume(xc, xd, t + 2);
        }
      }
#endif
      __syncwarp();
    }
    b_cur = b_nxt; total_cur = total_nxt; j_cur = j_nxt;
    b_nxt = b_nx2; total_nxt = total_nx2;
    if (warp == 0) TL(5);
    TS(3);
    // ---------------- 2. the previous tile's MMAs have finished reading the operand tiles
    if (tile > t_begin) {
      const uint32_t kp = (uint32_t)(tile - t_begin) - 1;
      mbar_wait(&mma_done[kp & 1], (kp >> 1) & 1);
      tc_fence_after();
    }
    if (warp == 0) TL(6);
    TS(4);
    // ---------------- 3. this tile's rows -> B-operand tiles (split bf16), then the MMAs
    {
      const uint32_t off = sw128_off(row, c0 >> 3) + (uint32_t)(c0 & 7) * 2;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float4 z = make_float4(lo32(acc[h][0]), hi32(acc[h][0]), lo32(acc[h][1]), hi32(acc[h][1]));
        uint2 hi, lo;
        split_bf16x4(z, hi, lo);
        *reinterpret_cast<uint2*>(z_hi + h * ZTILE_BYTES + off) = hi;
        *reinterpret_cast<uint2*>(z_lo + h * ZTILE_BYTES + off) = lo;
      }
    }
    // the next tile's first 16 rows of x and P: pull them into L1 now (their indices, loaded at the top of this
    // iteration, have arrived), so the next iteration starts on L1 hits
#if PREFETCH_MODE >= 2
    if (tile + 1 < t_end) {
#if PREFETCH_MODE == 2
      const float* xr = x + (unsigned)j_nxt * ldx32;
      prefetch_l1(xr);
      prefetch_l1(xr + 32);
#endif
      const double* pr = P + (int64_t)j_nxt * H;
      prefetch_l1(pr);
      prefetch_l1(pr + H - 1);
    }
#endif
    TS(5);
    fence_proxy_async();
    tc_fence_before();
    z_ready_arrive();      // non-blocking: go on with the next tile while warp 16 issues this one
    if (warp == 0) TL(7);
    TS(6);
#ifdef EXP_TIMELINE
    if (blockIdx.x == 0 && lane == 0 && tile - t_begin >= 20 && tile - t_begin < 84) g_tw[((tile - t_begin - 20) * 16 + warp) * 2 + 1] = now_ns();
#endif
  }
  __syncthreads();
}

}  // namespace fused

// feast.cu
int feast_project_only(const float* x, int64_t ldx, int64_t N, int c_in, const float* U, double* P, cudaStream_t st);

struct FusedWs {
  double* P;
  uint32_t* Wp;
};
template <class C>
static void carve_fused(C& c, int64_t N, FusedWs* out) {
  double* P = c.template take<double>((size_t)N * tc::H);
  uint32_t* Wp = c.template take<uint32_t>((size_t)fused::KSTEPS * 64 * 8);
  if (out) *out = FusedWs{P, Wp};
}
struct NullCarverFu {
  Sizer s;
  template <typename T>
  T* take(size_t n) { s.take<T>(n); return nullptr; }
};

size_t feast_fwd_fused_ws_bytes(int64_t N) {
  NullCarverFu c;
  carve_fused(c, N, nullptr);
  return c.s.total();
}

namespace fused {
// Wp[ks][m][8]: 32-bit words (bf16 pair, even k low) of stacked row m (m < 32: hi plane of output m, else lo plane of output m - 32) for
// K = 16 step ks; k = h * 64 + c (head-major, as the Z tiles are laid out).  W is FeaStConv's lin.weight [9 * 32, 64].
__global__ void prep_w_tmem_kernel(const float* __restrict__ W, uint32_t* __restrict__ Wp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= KSTEPS * 64 * 8) return;
  const int ks = idx / 512, m = (idx >> 3) & 63, i = idx & 7;
  const int o = m & 31, plane = m >> 5;
  uint32_t word = 0;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = 16 * ks + 2 * i + e, h = k / C_IN, c = k % C_IN;
    const float v = W[(h * C_OUT + o) * C_IN + c];
    const __nv_bfloat16 bh = __float2bfloat16_rn(v);
    const __nv_bfloat16 b = plane ? __float2bfloat16_rn(v - __bfloat162float(bh)) : bh;
    word |= (uint32_t)(*reinterpret_cast<const uint16_t*>(&b)) << (16 * e);
  }
  Wp[idx] = word;
}
}  // namespace fused

bool feast_fused_supported(int c_in, int c_out, int64_t ldx, int64_t ldo, const float* x, int64_t n_src) {
  return c_in == fused::C_IN && c_out == fused::C_OUT && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
         n_src * ldx < ((int64_t)1 << 32) && n_src > 0;
}

int feast_fwd_fused(const float* x, int64_t ldx, int64_t N, const int32_t* rowptr, const int32_t* nbr, const int32_t* row_map, int64_t n_src,
                    const float* W, const float* U, const float* c, const float* bias, float act_slope, float* out, int64_t ldo, bool reuse_ws,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < feast_fwd_fused_ws_bytes(n_src)) {
    set_error("feast_fwd (fused): workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  Carver cv(ws, ws_bytes);
  FusedWs Wk;
  carve_fused(cv, n_src, &Wk);
  if (!reuse_ws) {   // P = X.U^T (fp64) and the split-bf16 weight planes; skipped when the caller re-runs on the same inputs
    fused::prep_w_tmem_kernel<<<(fused::KSTEPS * 64 * 8 + 255) / 256, 256, 0, st>>>(W, Wk.Wp);
    GEOBI_LAUNCH_OK("feast_fused prep_w");
    int rc = feast_project_only(x, ldx, n_src, fused::C_IN, U, Wk.P, st);
    if (rc) return rc;
  }
  // one-time set-up as a thread-safe static initialiser: concurrent callers (bench.py runs two forwards per GPU from two host
  // threads) must not see the SM count before the kernels' shared-memory limit is raised - a plain `if (sms == 0)` flag let the
  // second thread launch with the 48 KB default ("invalid argument", seen once in ~10 bench runs)
  static const int sms = []() -> int {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(fused::feast_fused_64_32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(fused::feast_fused_64_32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::SMEM_BYTES) != cudaSuccess)
      return -1;
    return n;
  }();
  if (sms <= 0) {
    (void)cudaGetLastError();
    set_error("feast_fused: device query or shared-memory opt-in failed");
    return GEOBI_ERR_CUDA;
  }
  const int64_t n_tiles = (N + fused::NT - 1) / fused::NT;
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  if (row_map)
    fused::feast_fused_64_32_kernel<true><<<grid, fused::THREADS, fused::SMEM_BYTES, st>>>(x, ldx, N, rowptr, nbr, row_map, Wk.P, c, Wk.Wp, bias,
                                                                                           act_slope, out, ldo);
  else
    fused::feast_fused_64_32_kernel<false><<<grid, fused::THREADS, fused::SMEM_BYTES, st>>>(x, ldx, N, rowptr, nbr, row_map, Wk.P, c, Wk.Wp, bias,
                                                                                            act_slope, out, ldo);
  GEOBI_LAUNCH_OK("feast_fused");
  return GEOBI_OK;
}

#ifdef EXP_TIMELINE
extern "C" __attribute__((visibility("default"))) int geobi_debug_fused_timeline(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, fused::g_tl, sizeof(unsigned long long) * 64 * 8);
}
extern "C" __attribute__((visibility("default"))) int geobi_debug_fused_stages(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, fused::g_ts, sizeof(unsigned long long) * 64 * 2 * 8);
}
extern "C" __attribute__((visibility("default"))) int geobi_debug_fused_warps(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, fused::g_tw, sizeof(unsigned long long) * 64 * 16 * 2);
}
#endif

}  // namespace geobi
