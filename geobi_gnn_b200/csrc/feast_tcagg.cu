// Fused FeaSt convolution 64 -> 32 with the AGGREGATION on tcgen05 (round 2; replaces the FP32-pipe aggregation of
// feast_fused.cu on r_conv3 / r_conv4 of both U-Nets, /root/reference/code/network.py:267-268,295-299).
//
//   out_i = b + W_flat . Z_i,     Z_i[h, c] = 1/d_i  sum_{j in N(i)+{i}}  q_ijh x_j[c],     q_ij = softmax_h(u_h.(x_j - x_i) + c_h)
//
// Per target node the aggregation is a tiny GEMM  Z_i^T[c, h] = X_i^T[c, slot] . Q_i[slot, h]  over the node's <= 16
// neighbour slots.  Both operands are MN-major in shared memory, so
//   * the A operand is the gathered rows themselves: the layer's input is pre-split into bf16 hi | lo planes (prep_x_kernel:
//     256 B per row), and a neighbour row is copied with cp.async straight into its slot of a SWIZZLE_128B tile - no ALU work
//     per gathered element;
//   * the B operand is the soft-assignment tile [slot][q_hi(16) | q_lo(16)] (SWIZZLE_64B), written by the lane that owns the slot.
//   2 tcgen05.mma per node:  D[64 ch, 32] = x_hi^T.[q_hi | q_lo],  D[:, 0:16] += x_lo^T.q_hi   (bf16x3 split product).
// D lives in TMEM (M = 64 uses 16 lanes per quadrant, so two nodes share a 32-column group); drain warps add the two column
// halves, split Z into bf16 hi / lo and write the K-major B operand of the projection  out^T[o, node] = W[o, K=576] . Z[node, K],
// whose A operand - the weights, hi rows stacked on lo rows - is resident in TENSOR MEMORY (tcgen05.mma with A = [tmem]):
// the projection reads only Z from shared memory.  Measured on B200 (profiles/micro/tc_layout_probe.cu, tc_ts_probe.cu): an
// SS-mode MMA costs (A + B bytes) / 128 per clock, a TS-mode MMA 16 clocks at N = 32.
//
// Warp roles (768 threads, 1 CTA / SM, persistent over contiguous tiles of 32 nodes):
//   warps 0-3    epilogue: projection accumulator -> +bias, leaky_relu -> out; also load W into TMEM at start
//   warps 4-11   drain: aggregation accumulators -> Z operand tiles (two sets of 4 quadrant warps, alternating node pairs)
//   warp  12     issues the aggregation MMAs (four node pairs per batch of barrier waits)
//   warp  13     issues the projection MMAs
//   warps 14-23  producers: indices, cp.async gathers, soft assignments; two node pairs in flight per warp
#include "tc.cuh"

namespace geobi {
namespace tcagg {

using namespace tc;

constexpr int C_IN = 64, C_OUT = 32;
constexpr int TILE = 32;              // nodes per projection tile = MMA N
constexpr int PAIRS = TILE / 2;
// Build knobs (profiles/build_variant.sh builds A/B copies of the library; history and measurements in profiles/r02_NOTES.md A5).
// The defaults are the fastest combination measured; -DTCAGG_CLASSIC restores the first design (software-pipelined producers,
// whole-tile Z hand-over, 32-column drain loads, hinted try_wait everywhere, 10 producers).
#ifndef TCAGG_CLASSIC
#define TCAGG_SEQ 1          // producers take one item at a time (latency hidden by 12 warps, not by a 3-stage pipeline)
#define TCAGG_ZHALF 1        // Z handed to the projection in two 16-node halves
#define TCAGG_LD18 1         // drain warps read only the 18 useful accumulator columns
#define TCAGG_SLEEPWAIT 1    // long waits: test_wait + plain nanosleep instead of the hinted try_wait loop
#ifndef TCAGG_G
#define TCAGG_G 12
#endif
#endif
#ifndef TCAGG_G
#define TCAGG_G 10
#endif
constexpr int G = TCAGG_G;            // producer warps (768 threads: ptxas grants 80 registers up to that count anyway)
constexpr int R = 14;                 // ring slots, one node pair (2 x (x_hi, x_lo, q)) each; all the shared memory left
constexpr int DS = 8;                 // aggregation accumulator slots in TMEM (one node pair each)
#ifndef TCAGG_AB
#define TCAGG_AB 4
#endif
constexpr uint32_t AB = TCAGG_AB;     // node pairs the aggregation warp waits for and issues together (2 or 4; commits are per two slots)
// tcgen05.commit is not free (probe: +58 clk per pair when every pair commits, profiles/micro/tc_ts_probe.cu): ring slots and
// accumulator slots are released / published two at a time - xfree[slot >> 1], dfull[dslot >> 1] - by one commit each.
#ifndef TCAGG_NSETS
#define TCAGG_NSETS 2
#endif
constexpr int NSETS = TCAGG_NSETS;               // drain sets (4 quadrant warps each) taking node pairs in turn
constexpr int EPI_WARPS = 4, DRAIN_WARPS = 4 * NSETS;
// TCAGG_ZHALF: the Z tile is handed to the projection in two halves of 16 nodes (N = 16 MMAs): the projection of nodes 0-15 runs
// while nodes 16-31 are drained and is long finished when the drain warps come back to rows 0-15 with the next tile - without it
// the drain warps stop at every tile boundary until all 72 projection MMAs of the previous tile have read the (single) Z buffer
#ifdef TCAGG_ZHALF
constexpr int ZH = 2;
#else
constexpr int ZH = 1;
#endif
constexpr int ZPAIRS = PAIRS / ZH;               // node pairs per hand-over unit
#ifdef TCAGG_ZARRIVE_PAIR
static_assert(ZH == 1, "per-pair arrival is the old whole-tile protocol");
constexpr uint32_t ZFULL_COUNT = PAIRS * 4;      // every drain warp arrives after every pair
#else
constexpr uint32_t ZFULL_COUNT = DRAIN_WARPS;    // every drain warp arrives once per tile, after its last pair of the tile
static_assert(ZPAIRS % NSETS == 0, "each drain set takes the same pairs of every tile (half)");
#endif
constexpr int AGG_WARP = EPI_WARPS + DRAIN_WARPS;   // issues the aggregation MMAs
constexpr int PROJ_WARP = AGG_WARP + 1;              // issues the projection MMAs
constexpr int PROD_WARP0 = PROJ_WARP + 1;
constexpr int WARPS = PROD_WARP0 + G;
constexpr int THREADS = WARPS * 32;

// TMEM columns.  W: K steps 0-17 in lane half 0, 18-35 in lane half 16 (8 columns = 16 bf16 of K per step).
constexpr uint32_t COL_W = 0, COL_O = 144, COL_D = 208, TMEM_COLS = 512;
constexpr uint32_t HALF = 16u << 16;  // lane offset of the second node / second K half

constexpr int XT = 2048;                       // one plane of one node's slot tile: 16 slots x 128 B
constexpr int SLOT_BYTES = 4 * XT + 2 * 1024;  // A hi | A lo | B hi | B lo | q A | q B
constexpr int ZCHUNK = TILE * 128;             // one 64-wide K block of the Z operand
constexpr int ZPLANE = 9 * ZCHUNK;
constexpr int STAGE_BYTES = 2 * TILE * C_OUT * 4;
constexpr int SMEM_BYTES = 2 * ZPLANE + R * SLOT_BYTES + STAGE_BYTES + 1024;
constexpr int PROW = 12;                       // floats per row of the head projections u_h . x (computed in fp64): 9 + 3 pad

__host__ __device__ constexpr uint32_t idesc_of(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// high word of a shared-memory matrix descriptor: SBO (16-byte units) | version 1 | layout type
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo16, uint32_t layout) { return sbo16 | (1u << 14) | (layout << 29); }

__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the 18 useful columns of an aggregation accumulator (q_hi heads at columns 0-8, q_lo heads at 16-24): r[h] and r[9 + h]
__device__ __forceinline__ void tmem_ld18_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%18];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%8}, [%19];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%9, %10, %11, %12, %13, %14, %15, %16}, [%20];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%17}, [%21];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17])
      : "r"(taddr), "r"(taddr + 8), "r"(taddr + 16), "r"(taddr + 24)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~10 ms pass) instead of
// re-issuing the probe (ncu on the first version: a third of all issued warp instructions sat in wait loops)
__device__ __forceinline__ void mbar_wait_u(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_LOOP_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(addr), "r"(parity), "r"(0x989680u)
      : "memory");
}
// wait for long expected waits: non-blocking probe + a plain timed sleep.  The hinted try_wait above compiles to
// SYNCS.TRYWAIT / NANOSLEEP.SYNCS / SYNCS.PHASECHK loops in which every mbarrier event of the CTA wakes every sleeper: ncu counted
// ~450 SYNCS per node pair (20 % of the shared-memory data pipe, 44 % of the issued instructions with their branches) with ~100
// iterations per producer / drain / epilogue wait.  A plain nanosleep is not woken by barrier traffic.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t addr, uint32_t parity, uint32_t ns) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "SWAIT_LOOP_%=:\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SDONE_%=;\n\t"
      "nanosleep.u32 %2;\n\t"
      "bra SWAIT_LOOP_%=;\n\t"
      "SDONE_%=:\n\t}"
      ::"r"(addr), "r"(parity), "r"(ns)
      : "memory");
}
// non-blocking probe of an mbarrier phase (warp-uniform answer: lane 0 tests, the result is broadcast)
__device__ __forceinline__ bool phase_done(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return __shfl_sync(0xffffffffu, done, 0) != 0;
}

#ifdef TCAGG_TIMELINE
// per-role clock64() stamps of CTA 0 for pairs [128, 192) / tiles [8, 72): g_tl[role][index][event]
__device__ long long g_tl[5 * 64 * 8];
#define TLW(role, idx, ev) do { const int _i = (idx); if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && _i >= 0 && _i < 64) g_tl[((role) * 64 + _i) * 8 + (ev)] = clock64(); } while (0)
#else
#define TLW(role, idx, ev) do { } while (0)
#endif

#ifdef TCAGG_DEBUG
// bounded waits: a wait that does not complete records (warp, tag) and raises a flag that lets every other wait fall through,
// so a protocol bug ends the kernel instead of hanging the box
__device__ unsigned int g_dbg[64];
__device__ __forceinline__ void wait_dbg(uint64_t* bar, uint32_t parity, int tag) {
  const uint32_t addr = smem_u32(bar);
  for (long long it = 0;; ++it) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if ((it & 255) == 0) {
      unsigned int who = *(volatile unsigned int*)&g_dbg[63];
      if (who == 0 && it > (1ll << 17)) {
        atomicCAS(&g_dbg[63], 0u, blockIdx.x + 1);
        who = *(volatile unsigned int*)&g_dbg[63];
      }
      if (who) {       // the CTA whose wait timed out first records the FIRST wait each of its warps gave up on; everybody falls through
        if (who == blockIdx.x + 1 && *(volatile unsigned int*)&g_dbg[threadIdx.x >> 5] == 0)
          g_dbg[threadIdx.x >> 5] = (unsigned)tag | (parity << 31);
        return;
      }
    }
  }
}
#define PROGRESS(idx, val) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_dbg[idx] = (unsigned)(val); } while (0)
#define WAIT(bar, parity, tag) wait_dbg(bar, parity, (tag) | ((int)((bar) - bars) << 8))
#define WAIT_U(addr, parity, tag) wait_dbg(reinterpret_cast<uint64_t*>(__cvta_shared_to_generic(addr)), parity, (tag) | ((int)(((addr) - smem_u32(bars)) >> 3) << 8))
#else
#define PROGRESS(idx, val) do { } while (0)
#define WAIT(bar, parity, tag) mbar_wait_u(smem_u32(bar), parity)
#define WAIT_U(addr, parity, tag) mbar_wait_u(addr, parity)
#endif
#if defined(TCAGG_SLEEPWAIT) && !defined(TCAGG_DEBUG)
#define WAIT_LONG(bar, parity, tag, ns) mbar_wait_sleep(smem_u32(bar), parity, ns)
#else
#define WAIT_LONG(bar, parity, tag, ns) WAIT(bar, parity, tag)
#endif
#ifndef TCAGG_NS_X
#define TCAGG_NS_X 64
#endif
#ifndef TCAGG_NS_Z
#define TCAGG_NS_Z 64
#endif
#ifndef TCAGG_NS_O
#define TCAGG_NS_O 256
#endif
__device__ __forceinline__ void mma_commit_u(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// prep: x (fp32 rows) -> Xs[row] = bf16 hi[64] | lo[64]  (256 B)  and  P[row] = the head projections u_h . x_row, accumulated in fp64
// and rounded once to fp32 (|error| <= 2^-24 |P|: the soft assignments use P_j - P_i, so their absolute error is ~1e-7 |P|).
constexpr int PREP_ROWS = 64, PREP_THREADS = 256;
__global__ void __launch_bounds__(PREP_THREADS) prep_x_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, const float* __restrict__ U,
                                                             uint8_t* __restrict__ Xs, float* __restrict__ P) {
  __shared__ float xs[PREP_ROWS][C_IN + 1];
  __shared__ double us[C_IN][12];
  const int tid = threadIdx.x;
  for (int i = tid; i < C_IN * 12; i += PREP_THREADS) {
    const int c = i / 12, h = i % 12;
    us[c][h] = h < H ? (double)U[h * C_IN + c] : 0.0;
  }
  const int64_t row0 = (int64_t)blockIdx.x * PREP_ROWS;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + PREP_THREADS * i, r = idx >> 4, c4 = idx & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n) v = __ldg(reinterpret_cast<const float4*>(x + (row0 + r) * ldx + 4 * c4));
    xs[r][4 * c4 + 0] = v.x; xs[r][4 * c4 + 1] = v.y; xs[r][4 * c4 + 2] = v.z; xs[r][4 * c4 + 3] = v.w;
  }
  __syncthreads();
  const int r = tid >> 2, part = tid & 3;
  if (row0 + r >= n) return;
  // heads part, part + 4, part + 8 (only part 0 owns a third head)
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 8
  for (int c = 0; c < C_IN; ++c) {
    const double xv = (double)xs[r][c];
    a0 = fma(xv, us[c][part], a0);
    a1 = fma(xv, us[c][part + 4], a1);
    a2 = fma(xv, us[c][part + 8], a2);
  }
  float* pr = P + (row0 + r) * PROW;
  pr[part] = (float)a0;
  pr[part + 4] = (float)a1;
  if (part == 0) pr[8] = (float)a2;
  else pr[8 + part] = 0.f;
  // split this thread's 16 channels
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float v0 = xs[r][part * 16 + 2 * i], v1 = xs[r][part * 16 + 2 * i + 1];
    const uint32_t h2 = pack_bf16x2(v0, v1);
    const float f0 = __uint_as_float(h2 << 16), f1 = __uint_as_float(h2 & 0xffff0000u);
    hi[i] = h2;
    lo[i] = pack_bf16x2(v0 - f0, v1 - f1);
  }
  uint4* dst = reinterpret_cast<uint4*>(Xs + (row0 + r) * 256 + part * 32);
  dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  dst[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  dst[8] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  dst[9] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}

// K ordering of the projection ("K3"): k < 512: head pair hp = k / 128, channel c = (k % 128) / 2, head h = 2 hp + (k & 1);
// k >= 512: head 8, channel c = k - 512.  A drain thread (one channel, 9 heads) then writes its head pairs as packed 32-bit words.
__host__ __device__ inline void k3_decode(int k, int& h, int& c) {
  if (k < 512) {
    h = 2 * (k >> 7) + (k & 1);
    c = (k & 127) >> 1;
  } else {
    h = 8;
    c = k - 512;
  }
}
// Wp[ks][m][8]: 32-bit words (bf16 pair: even k low) of stacked row m (m < 32: hi plane of output m, else lo plane of m - 32)
__global__ void prep_w_kernel(const float* __restrict__ W /* [9*32][64] */, uint32_t* __restrict__ Wp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 36 * 64 * 8) return;
  const int ks = idx / 512, m = (idx >> 3) & 63, i = idx & 7;
  const int o = m & 31, plane = m >> 5;
  uint32_t word = 0;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    int h, c;
    k3_decode(16 * ks + 2 * i + e, h, c);
    const float v = W[(h * C_OUT + o) * C_IN + c];
    const __nv_bfloat16 bh = __float2bfloat16_rn(v);
    const __nv_bfloat16 b = plane ? __float2bfloat16_rn(v - __bfloat162float(bh)) : bh;
    word |= (uint32_t)(*reinterpret_cast<const uint16_t*>(&b)) << (16 * e);
  }
  Wp[idx] = word;
}

// ---------------------------------------------------------------------------------------------------------------------
template <bool HAS_MAP>
__global__ void __launch_bounds__(THREADS, 1) feast_tcagg_64_32_kernel(const uint8_t* __restrict__ Xs, const float* __restrict__ P, int64_t N,
                                                                       const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                                       const int* __restrict__ row_map, const float* __restrict__ cvec,
                                                                       const uint32_t* __restrict__ Wp, const float* __restrict__ bias,
                                                                       float slope, float* __restrict__ out, int64_t ldo) {
  extern __shared__ uint8_t smem_raw[];
  // one array so that the debug build can name a barrier by its index: full | xfree | dfull | dfree | zfull | zfree | ofull | ofree
  __shared__ __align__(8) uint64_t bars[R + R / 2 + DS / 2 + DS + 8];
  uint64_t* const full = bars;                          // [R]      per ring slot
  uint64_t* const xfree = bars + R;                     // [R / 2]  per pair of ring slots
  uint64_t* const dfull = bars + R + R / 2;             // [DS / 2] per pair of accumulator slots
  uint64_t* const dfree = bars + R + R / 2 + DS / 2;    // [DS]     per accumulator slot
  uint64_t* const zfull = bars + R + R / 2 + DS / 2 + DS;        // [ZH] per node half of the Z tile (ZH = 1: the whole tile)
  uint64_t* const zfree = bars + R + R / 2 + DS / 2 + DS + 2;    // [ZH]
  uint64_t* const ofull = bars + R + R / 2 + DS / 2 + DS + 4;
  uint64_t* const ofree = bars + R + R / 2 + DS / 2 + DS + 6;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t agg_pos;      // ring sequence numbers issued so far by the aggregation-issue warp (see `slot_reusable`)
  __shared__ __align__(16) float chs[12];
  __shared__ __align__(16) int jbuf[G][32];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* z_hi = sm;
  uint8_t* z_lo = sm + ZPLANE;
  uint8_t* ring = sm + 2 * ZPLANE;
  float* stage = reinterpret_cast<float*>(ring + R * SLOT_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < R; ++i) mbar_init(&full[i], 1);
    for (int i = 0; i < R / 2; ++i) mbar_init(&xfree[i], 1);
    for (int i = 0; i < DS / 2; ++i) mbar_init(&dfull[i], 1);
    for (int i = 0; i < DS; ++i) mbar_init(&dfree[i], 4);
    for (int i = 0; i < 2; ++i) mbar_init(&zfull[i], ZFULL_COUNT);
    for (int i = 0; i < 2; ++i) mbar_init(&zfree[i], 1);
    mbar_init(&ofull[0], 1);
    mbar_init(&ofull[1], 1);
    mbar_init(&ofree[0], EPI_WARPS);
    mbar_init(&ofree[1], EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == AGG_WARP) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid < H) chs[tid] = cvec[tid];
  if (tid == 32) agg_pos = 0;
  // slots that a node does not use keep whatever an earlier node left there (finite, and multiplied by q = 0): start finite
  for (int i = tid; i < R * SLOT_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp < EPI_WARPS) {
    // weights -> TMEM: lane l of quadrant warp w holds stacked row 16 w + (l & 15); lanes 0-15 K steps 0-17, lanes 16-31 K steps 18-35
    const int m = 16 * warp + (lane & 15), ks0 = 18 * (lane >> 4);
    const uint4* src = reinterpret_cast<const uint4*>(Wp);
#pragma unroll 2
    for (int i = 0; i < 18; ++i) {
      const uint4 a = __ldg(src + ((ks0 + i) * 64 + m) * 2), b = __ldg(src + ((ks0 + i) * 64 + m) * 2 + 1);
      tmem_st8(tmem + COL_W + 8 * i + ((uint32_t)(warp * 32) << 16), a, b);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int64_t n_tiles = (N + TILE - 1) / TILE;
  const int64_t t_begin = (n_tiles * blockIdx.x) / gridDim.x, t_end = (n_tiles * (blockIdx.x + 1)) / gridDim.x;
  const int T = (int)(t_end - t_begin);

  // Per tile, every warp that needs the ring sequence derives it from the tile's degrees: pair pp = nodes (rA, rA + 4) with
  // rA = (pp / 4) * 8 + pp % 4 (rows 4 apart land in different bank groups of the swizzled Z tile); a pair takes
  // rounds = max(1, ceil(max(d_A, d_B) / 16)) ring slots in sequence (d = neighbours + 1; nodes past N have d = 0).
  struct TileInfo {
    int rb, d;          // lane = tile row: first CSR entry and slot count of node base + lane
    float inv_d;        // lane = tile row: 1 / d (0 for rows past N)
    int rounds, excl;   // lane = pair (lanes 0-15): ring slots of the pair and their exclusive prefix inside the tile
    int total;
  };
  auto tile_info = [&](int64_t tile, int rb_in, int re_in) {
    TileInfo ti;
    const int64_t node = tile * TILE + lane;
    ti.rb = rb_in;
    ti.d = node < N ? re_in - rb_in + 1 : 0;
    ti.inv_d = ti.d > 0 ? __frcp_rn((float)ti.d) : 0.f;
    const int rA = ((lane >> 2) & 3) * 8 + (lane & 3);
    const int dA = __shfl_sync(0xffffffffu, ti.d, rA), dB = __shfl_sync(0xffffffffu, ti.d, rA + 4);
    int rounds = lane < PAIRS ? max(1, (max(dA, dB) + 15) >> 4) : 0;
    int incl = rounds;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    ti.rounds = rounds;
    ti.excl = incl - rounds;
    ti.total = __shfl_sync(0xffffffffu, incl, 15);
    return ti;
  };
  auto load_rowptr = [&](int64_t tile, int& rb, int& re) {
    const int64_t node = tile * TILE + lane;
    if (tile < t_end && node < N) {
      rb = __ldg(rowptr + node);
      re = __ldg(rowptr + node + 1);
    } else {
      rb = 0;
      re = 0;
    }
  };

  if (warp == AGG_WARP) {
    // ================================================== aggregation MMA issue ==================================================
    // A single warp issues everything serially (~5 clk per dependent instruction, ~90 clk per mbarrier wait), so the common case -
    // every pair of the tile takes one ring slot - is handled four pairs at a time: lanes 0-3 wait on the pairs' `full`
    // barriers and lanes 4-7 on their `dfree` barriers concurrently, then one elected lane issues the 16 MMAs and 8 commits with
    // incrementally maintained slot indices (no divisions).
    constexpr uint32_t IA = idesc_of(64, 32, 1, 1), IB = idesc_of(64, 16, 1, 1);
    constexpr uint32_t XHI = desc_hi(64, 2), QHI = desc_hi(32, 4);
    const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFFu) >> 4;
    const uint32_t full_u = smem_u32(&full[0]), xfree_u = smem_u32(&xfree[0]), dfull_u = smem_u32(&dfull[0]), dfree_u = smem_u32(&dfree[0]);
    int rb_n, re_n;
    load_rowptr(t_begin, rb_n, re_n);
    uint32_t seq_base = 0;
    TileInfo ti{};
    // `last`: this is the pair's final round.  Commits go out when the second slot of a group has been issued.
    auto issue_pair = [&](uint32_t slot, uint32_t dslot, uint32_t acc, bool last) {
      const uint32_t xa = ring_lo + slot * (SLOT_BYTES >> 4), qa = xa + (4 * XT >> 4);
      const uint32_t dA = tmem + COL_D + dslot * 32, dB = dA + HALF;
      mma_ss(dA, xa, XHI, qa, QHI, IA, acc);
      mma_ss(dA, xa + (XT >> 4), XHI, qa, QHI, IB, 1u);
      mma_ss(dB, xa + (2 * XT >> 4), XHI, qa + (1024 >> 4), QHI, IA, acc);
      mma_ss(dB, xa + (3 * XT >> 4), XHI, qa + (1024 >> 4), QHI, IB, 1u);
      if (slot & 1u) mma_commit_u(xfree_u + 8 * (slot >> 1));
      if (last && (dslot & 1u)) mma_commit_u(dfull_u + 8 * (dslot >> 1));
    };
    auto do_pair = [&](uint32_t p, int pp) {       // general path: pairs that take several rounds
      const uint32_t dslot = p % DS, dk = p / DS;
      const int rounds = __shfl_sync(0xffffffffu, ti.rounds, pp);
      const uint32_t seq0 = seq_base + (uint32_t)__shfl_sync(0xffffffffu, ti.excl, pp);
      for (int r = 0; r < rounds; ++r) {
        const uint32_t seq = seq0 + r, slot = seq % R, k = seq / R;
        TLW(1, (int)p - 128, 0);
        WAIT(&full[slot], k & 1, 1);
        TLW(1, (int)p - 128, 1);
        if (r == 0) WAIT(&dfree[dslot], (dk & 1) ^ 1, 2);
        TLW(1, (int)p - 128, 2);
        tc_fence_after();
        if (elect_one()) {
          issue_pair(slot, dslot, r > 0, r == rounds - 1);
          *(volatile uint32_t*)&agg_pos = seq + 1;
        }
        __syncwarp();
        PROGRESS(32, (seq + 1) | (p << 16));
        TLW(1, (int)p - 128, 3);
      }
    };
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      ti = tile_info(t_begin + t, rb_n, re_n);
      load_rowptr(t_begin + t + 1, rb_n, re_n);
      if (ti.total == PAIRS) {
        const uint32_t p0 = (uint32_t)(PAIRS * t);
        const uint32_t s0 = seq_base % R, k0 = (seq_base / R) & 1u, d0 = p0 % DS, dk0 = (p0 / DS) & 1u;
#pragma unroll 1
        for (uint32_t b4 = 0; b4 < PAIRS; b4 += AB) {
          {   // lanes 0..AB-1: full[pair b4 + lane]; lanes AB..2AB-1: dfree[pair b4 + lane - AB]
            const uint32_t q = b4 + (lane % AB);
            uint32_t s = s0 + q, ks = k0, d = d0 + q, kd = dk0 ^ 1u;
            if (s >= R) { s -= R; ks ^= 1u; }
            if (s >= R) { s -= R; ks ^= 1u; }
            if (d >= DS) { d -= DS; kd ^= 1u; }
            if (d >= DS) { d -= DS; kd ^= 1u; }
            TLW(1, (int)(p0 + b4) - 128, 0);
            if (lane < 2 * AB) WAIT_U(lane < AB ? full_u + 8 * s : dfree_u + 8 * d, lane < AB ? ks : kd, 1);
            __syncwarp();
          }
#ifdef TCAGG_TIMELINE
          for (int q = 0; q < (int)AB; ++q) { TLW(1, (int)(p0 + b4) + q - 128, 1); TLW(1, (int)(p0 + b4) + q - 128, 2); }
#endif
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (uint32_t q = 0; q < AB; ++q) {
              uint32_t s = s0 + b4 + q, d = d0 + b4 + q;
              if (s >= R) s -= R;
              if (s >= R) s -= R;
              if (d >= DS) d -= DS;
              if (d >= DS) d -= DS;
              issue_pair(s, d, 0u, true);
            }
            *(volatile uint32_t*)&agg_pos = seq_base + b4 + AB;
          }
          __syncwarp();
#ifdef TCAGG_TIMELINE
          for (int q = 0; q < (int)AB; ++q) TLW(1, (int)(p0 + b4) + q - 128, 3);
#endif
        }
      } else {
#pragma unroll 1
        for (int pp = 0; pp < PAIRS; ++pp) do_pair((uint32_t)(PAIRS * t + pp), pp);
      }
      seq_base += (uint32_t)ti.total;
    }
  } else if (warp == PROJ_WARP) {
    // ================================================== projection MMA issue ==================================================
    constexpr uint32_t IP = idesc_of(64, TILE / ZH, 0, 0), ZHI = desc_hi(64, 2);
    const uint32_t zh_lo = ((smem_u32(z_hi) & 0x3FFFFu) >> 4) | (1u << 16), zl_lo = ((smem_u32(z_lo) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const uint32_t b = t & 1;
#pragma unroll 1
      for (int nh = 0; nh < ZH; ++nh) {      // node half: Z rows 16 nh .. (2048 bytes into every 64-column chunk), accumulator columns 16 nh ..
        if (nh == 0) TLW(3, t - 8, 0);
        WAIT_LONG(&zfull[nh], t & 1, 3, TCAGG_NS_Z);
        if (nh == 0) {
          TLW(3, t - 8, 1);
          WAIT(&ofree[b], ((t >> 1) & 1) ^ 1, 4);
          TLW(3, t - 8, 2);
        }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t zrow = (uint32_t)nh * (2048u >> 4), dcol = (uint32_t)nh * (TILE / ZH);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const uint32_t d = tmem + COL_O + 32 * b + dcol + hf * HALF, a0 = tmem + COL_W + hf * HALF;
#pragma unroll 2
            for (int i = 0; i < 18; ++i) {
              const int ks = 18 * hf + i;
              const uint32_t zoff = (uint32_t)(ks >> 2) * (ZCHUNK >> 4) + 2 * (ks & 3) + zrow;
              mma_ts(d, a0 + 8 * i, zh_lo + zoff, ZHI, IP, i > 0);
              mma_ts(d, a0 + 8 * i, zl_lo + zoff, ZHI, IP, 1u);
            }
          }
          mma_commit(&zfree[nh]);
          if (nh == ZH - 1) mma_commit(&ofull[b]);
        }
        __syncwarp();
      }
      TLW(3, t - 8, 3);
    }
  } else if (warp < EPI_WARPS) {
    // ================================================== epilogue ==================================================
    // quadrant q: lanes 0-15 = accumulator of K half 0, lanes 16-31 = K half 1; rows 16 q + l: q = 0, 1 -> W_hi rows of outputs
    // 0-15 / 16-31, q = 2, 3 -> the W_lo rows of the same outputs
    const int o = 16 * (warp & 1) + (lane & 15);
    const float my_bias = bias[o];
    for (int t = 0; t < T; ++t) {
      const uint32_t b = t & 1;
      WAIT_LONG(&ofull[b], (t >> 1) & 1, 5, TCAGG_NS_O);
      if (warp == 0) TLW(3, t - 8, 4);
      tc_fence_after();
      float* sb = stage + (t & 1) * (TILE * C_OUT);
      const int64_t n0 = (t_begin + t) * TILE;
      const int live = (int)(N - n0 < TILE ? N - n0 : TILE);
      // two passes of 16 nodes (32 accumulator registers at once do not fit under the 80-register cap of a 768-thread CTA)
#pragma unroll 1
      for (int hcol = 0; hcol < 2; ++hcol) {
        float v[16];
        tmem_ld16(tmem + COL_O + 32 * b + 16 * hcol + ((uint32_t)(warp * 32) << 16), v);
        if (hcol == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ofree[b]);
        }
#pragma unroll
        for (int n = 0; n < 16; ++n) v[n] += __shfl_xor_sync(0xffffffffu, v[n], 16);
        if (warp >= 2 && lane < 16) {
#pragma unroll
          for (int n = 0; n < 16; ++n) sb[(16 * hcol + n) * C_OUT + o] = v[n];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp < 2 && lane < 16) {
          float* op = out + (n0 + 16 * hcol) * ldo + o;
#pragma unroll
          for (int n = 0; n < 16; ++n) {
            float r = v[n] + sb[(16 * hcol + n) * C_OUT + o] + my_bias;
            r = r > 0.f ? r : r * slope;
            if (16 * hcol + n < live) *op = r;
            op += ldo;
          }
        }
      }
      if (warp == 0) TLW(3, t - 8, 5);
    }
  } else if (warp < AGG_WARP) {
    // ================================================== drain ==================================================
    const int set = (warp - EPI_WARPS) >> 2, qd = warp & 3;
    const int half = lane >> 4, c = 16 * qd + (lane & 15);
    const int npairs = PAIRS * T;
    // this thread's byte offset inside a Z row: head pairs -> 32-bit word of channel c in K block 2 hp + (c >= 32); head 8 -> K block 8.
    // The row of pair pp is (pp / 4) * 8 + (pp % 4) + 4 half, so row & 7 = (pp & 3) + 4 half and the swizzle term splits into a
    // per-thread part and (pp & 3) << 4.
    const uint32_t hconst = (uint32_t)half * 512;                                  // 4 rows of 128 B
    const uint32_t sw_pair = ((((uint32_t)(c & 31) >> 2) ^ (4u * half)) << 4) + (uint32_t)(c & 3) * 4 + (uint32_t)(c >> 5) * ZCHUNK + hconst;
    const uint32_t sw_h8 = ((((uint32_t)c >> 3) ^ (4u * half)) << 4) + (uint32_t)(c & 7) * 2 + 8 * ZCHUNK + hconst;
    const uint32_t lane_t = tmem + COL_D + ((uint32_t)(qd * 32) << 16);
#ifdef TCAGG_LD18
    constexpr int NR = 18, LO = 9;
#else
    constexpr int NR = 32, LO = 16;
#endif
#ifdef TCAGG_DRAIN1
    uint32_t ra[NR];
#else
    uint32_t ra[NR], rb[NR];
#endif
#ifdef TCAGG_LD18
#define DRAIN_LD tmem_ld18_issue
#else
#define DRAIN_LD tmem_ld32_issue
#endif
    auto dwait_ld = [&](int p, uint32_t (&r)[NR]) {
      const uint32_t dslot = (uint32_t)p % DS, dk = (uint32_t)p / DS;
      if (qd == 0) TLW(2, p - 128, 0);
      WAIT(&dfull[dslot >> 1], dk & 1, 6);
      if (qd == 0) TLW(2, p - 128, 1);
      tc_fence_after();
      DRAIN_LD(lane_t + dslot * 32, r);
    };
    auto process = [&](int p, uint32_t (&r)[NR], int p_next, uint32_t (&rn)[NR]) -> bool {
      const int t = p >> 4, pp = p & 15;
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dfree[(uint32_t)p % DS]);
      if (qd == 0) TLW(2, p - 128, 2);
      float z[H];
#pragma unroll
      for (int h = 0; h < H; ++h) z[h] = __uint_as_float(r[h]) + __uint_as_float(r[LO + h]);
      // the next pair's accumulator, if it is already complete: its TMEM read overlaps the split below
      bool pre = false;
#ifndef TCAGG_DRAIN1
      if (p_next < npairs) {
        pre = phase_done(&dfull[((uint32_t)p_next % DS) >> 1], ((uint32_t)p_next / DS) & 1);
        if (pre) {
          tc_fence_after();
          DRAIN_LD(lane_t + ((uint32_t)p_next % DS) * 32, rn);
        }
      }
#endif
      const int zu = pp / ZPAIRS, zp = pp % ZPAIRS;
      if (zp < NSETS && t >= 1) {   // first pair of this warp in this (half) tile: the projection of tile t-1 must have finished reading these Z rows
        WAIT_LONG(&zfree[zu], (t - 1) & 1, 7, TCAGG_NS_Z);
        tc_fence_after();
      }
      if (qd == 0) TLW(2, p - 128, 3);
      const uint32_t rowoff = (uint32_t)(pp >> 2) * 1024 + (uint32_t)(pp & 3) * 128, sx = (uint32_t)(pp & 3) << 4;
      uint8_t* zh = z_hi + rowoff;
      uint8_t* zl = z_lo + rowoff;
      const uint32_t off = sw_pair ^ sx;      // (chunk ^ (row & 7)) << 4: the pair's part of the XOR only touches bits 4-5
#pragma unroll
      for (int hp = 0; hp < 4; ++hp) {
        const float za = z[2 * hp], zb = z[2 * hp + 1];
        const uint32_t h2 = pack_bf16x2(za, zb);
        const float fa = __uint_as_float(h2 << 16), fb = __uint_as_float(h2 & 0xffff0000u);
        const uint32_t l2 = pack_bf16x2(za - fa, zb - fb);
        *reinterpret_cast<uint32_t*>(zh + 2 * hp * ZCHUNK + off) = h2;
        *reinterpret_cast<uint32_t*>(zl + 2 * hp * ZCHUNK + off) = l2;
      }
      {
        const uint32_t off8 = sw_h8 ^ sx;
        const uint32_t h1 = pack_bf16x2(z[8], 0.f);
        const uint32_t l1 = pack_bf16x2(z[8] - __uint_as_float(h1 << 16), 0.f);
        *reinterpret_cast<uint16_t*>(zh + off8) = (uint16_t)h1;
        *reinterpret_cast<uint16_t*>(zl + off8) = (uint16_t)l1;
      }
      if (qd == 0) TLW(2, p - 128, 4);
#ifdef TCAGG_ZARRIVE_PAIR
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&zfull[0]);
#else
      if (zp >= ZPAIRS - NSETS) {   // this warp's last pair of the (half) tile: one fence + arrival per warp and hand-over unit
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&zfull[zu]);
      }
#endif
      if (qd == 0) TLW(2, p - 128, 5);
      if (qd == 0) PROGRESS(34 + set, p + 1);
      return pre;
    };
    // two register sets alternate; `have` = the current set's TMEM read has been issued already
#ifdef TCAGG_DRAIN1
    // one register set, no look-ahead: for builds with more drain sets under a tighter register cap
#pragma unroll 1
    for (int p = set; p < npairs; p += NSETS) {
      dwait_ld(p, ra);
      process(p, ra, npairs, ra);
    }
#else
    bool have = false;
#pragma unroll 1
    for (int p = set; p < npairs; p += 2 * NSETS) {
      if (!have) dwait_ld(p, ra);
      have = process(p, ra, p + NSETS, rb);
      const int p2 = p + NSETS;
      if (p2 < npairs) {
        if (!have) dwait_ld(p2, rb);
        have = process(p2, rb, p2 + NSETS, ra);
      }
    }
#endif
  } else {
    // ================================================== producers ==================================================
    const int g = warp - PROD_WARP0;
    const int half = lane >> 4, sl = lane & 15;
    const uint32_t ring_u32 = smem_u32(ring);
    int* jb = jbuf[g];
    // per-lane constants of the gather: lanes 0-15 copy the 16 chunks (0-7 hi plane, 8-15 lo plane) of the row of slot-lane 2i,
    // lanes 16-31 of slot-lane 2i+1; slot s = 2 (i & 7) + half of node i >> 3:
    //   dst = slot base + (i >> 3) 4096 + ((i >> 2) & 1) 1024 + (i & 3) 256 + half 128 + plane 2048 + ((chunk ^ (s & 7)) << 4)
    const uint32_t ch = (uint32_t)sl;
    const uint32_t lane_c = (ch >> 3) * XT + (uint32_t)half * 128;
    const uint32_t c7h = ((ch & 7) ^ (uint32_t)half) << 4;
    const uint8_t* src0 = Xs + ch * 16;
    asm volatile("" : "+l"(src0));      // keep the pointer in registers (ptxas otherwise re-derives it from the parameter bank per copy)
    // per-lane constants of the q tile row (SWIZZLE_64B): 64 B = chunks {q_hi h0-7, q_hi h8 + zeros, q_lo h0-7, q_lo h8 + zeros}
    const uint32_t qrow = 4 * XT + (uint32_t)half * 1024 + (uint32_t)(sl >> 3) * 512 + (uint32_t)(sl & 7) * 64, qx = (uint32_t)(sl >> 1) & 3u;
    // work items: (tile t, pair pp, round r) with (16 t + pp) % G == g, in order
    struct Item {         // five registers: three items are live across the software pipeline
      uint32_t seq;       // ring sequence number
      uint32_t bits;      // slot (0-7) | xfree parity (8) | valid (9) | round 0 (10) | lane's slot is a real neighbour / self slot (11)
      int j;              // source row of this lane's slot (the node's own row when the slot is empty: finite data under q = 0)
      int self;           // source row of this lane's node
      float inv_d;
#ifdef TCAGG_TIMELINE
      int p;              // pair index
#endif
      __device__ __forceinline__ bool valid() const { return (bits >> 9) & 1u; }
      __device__ __forceinline__ uint32_t slot() const { return bits & 0xffu; }
      __device__ __forceinline__ uint32_t par() const { return (bits >> 8) & 1u; }
      __device__ __forceinline__ bool first() const { return (bits >> 10) & 1u; }
      __device__ __forceinline__ bool has() const { return (bits >> 11) & 1u; }
    };
#ifdef TCAGG_TIMELINE
#define ITEM_P(it) ((it).p)
#else
#define ITEM_P(it) (-1000)
#endif
    const int n_items = PAIRS * T;
    const int tile0 = (int)(t_begin * TILE);             // N < 2^31 (rowptr is int32)
    int it_t = -1, it_pp = 0, it_r = 0, it_rounds = 0;
    uint32_t seq_base = 0, next_base = 0;
    TileInfo ti{};
    int rb_n, re_n;
    load_rowptr(t_begin, rb_n, re_n);
    int p_next = g;
    auto next_item = [&]() {
      Item it{};
      if (it_t >= 0 && it_r + 1 < it_rounds) {
        ++it_r;
      } else {
        if (p_next >= n_items) return it;                // bits = 0: not valid
        while (it_t < (p_next >> 4)) {                   // enter the next tile (every warp walks every tile: the ring sequence numbers are cumulative)
          ++it_t;
          seq_base = next_base;
          ti = tile_info(t_begin + it_t, rb_n, re_n);
          load_rowptr(t_begin + it_t + 1, rb_n, re_n);
          next_base = seq_base + (uint32_t)ti.total;
        }
        it_pp = p_next & 15;
        it_r = 0;
        it_rounds = __shfl_sync(0xffffffffu, ti.rounds, it_pp);
        p_next += G;
      }
      const int row = (it_pp >> 2) * 8 + (it_pp & 3) + 4 * half;
      const int rb = __shfl_sync(0xffffffffu, ti.rb, row), d = __shfl_sync(0xffffffffu, ti.d, row);
      it.inv_d = __shfl_sync(0xffffffffu, ti.inv_d, row);
#ifdef TCAGG_TIMELINE
      it.p = it_r == 0 ? PAIRS * it_t + it_pp : -1000;
#endif
      it.seq = seq_base + (uint32_t)__shfl_sync(0xffffffffu, ti.excl, it_pp) + (uint32_t)it_r;
      const uint32_t k = it.seq / (uint32_t)R;
      const int s = 16 * it_r + sl;
      const bool has = s < d;
      it.bits = (it.seq - k * (uint32_t)R) | (((k & 1u) ^ 1u) << 8) | (1u << 9) | ((it_r == 0 ? 1u : 0u) << 10) | ((has ? 1u : 0u) << 11);
      int self = d > 0 ? tile0 + it_t * TILE + row : 0;
      int j = self;
      if (has && s > 0) j = __ldg(nbr + rb + s - 1);
      if (HAS_MAP) {
        if (d > 0) self = __ldg(row_map + self);
        j = (has && s > 0) ? __ldg(row_map + j) : self;
      }
      it.j = j;
      it.self = self;
      return it;
    };
    // A parity wait cannot tell "the slot's previous occupant has been consumed" from "the occupant before that has been consumed
    // and the previous one not even issued" - which happens once a producer runs two ring laps ahead of the issue warp (pairs that
    // take several rounds spread a warp's items further than R apart).  So the xfree wait of sequence number s is only entered
    // after the issue warp has issued s - R: from then on the barrier is at most one phase behind the one waited for.
    uint32_t prev_seq = 0;
    bool have_prev = false;
    auto occupant_issued = [&](uint32_t seq) { return seq < (uint32_t)R || *(volatile uint32_t*)&agg_pos + (uint32_t)R > seq; };
    // stage 1 of an item: ring slot free -> gathers in flight, P row of this lane's slot in flight
    auto begin = [&](const Item& it) {
      const uint32_t slot = it.slot();
      TLW(0, ITEM_P(it) - 128, 0);
      // the half warp that copies slot-lanes half, half + 2, ... finds their rows contiguous: index (L & 1) 16 + (L >> 1)
      __syncwarp();
      jb[(lane & 1) * 16 + (lane >> 1)] = it.j;
      __syncwarp();
      const unsigned mask = __ballot_sync(0xffffffffu, it.has());
      const int cA = __popc(mask & 0xffffu), cB = __popc(mask >> 16);      // slots fill from 0: counts decide which iterations run
      const int4* jv = reinterpret_cast<const int4*>(jb + half * 16);
      const int4 j0 = jv[0], j1 = jv[1], j2 = jv[2], j3 = jv[3];
      const int jr[16] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w, j2.x, j2.y, j2.z, j2.w, j3.x, j3.y, j3.z, j3.w};
      // fast path: this warp's previous item passed its own xfree wait, so everything up to prev_seq - R has been consumed (the issue
      // warp works in order); if the slot's occupant two laps back (seq - 2R) is among that, the barrier is at most one phase behind
      // and the poll of agg_pos (one shared-memory wavefront per probe: it was a quarter of the kernel's shared traffic) is not needed
      if (!(have_prev && it.seq - prev_seq <= (uint32_t)R))
        while (!occupant_issued(it.seq)) __nanosleep(256);
      prev_seq = it.seq;
      have_prev = true;
      WAIT_LONG(&xfree[slot >> 1], it.par(), 8, TCAGG_NS_X);
      TLW(0, ITEM_P(it) - 128, 1);
      tc_fence_after();
      const uint32_t sbase = ring_u32 + slot * SLOT_BYTES;
      const uint32_t gb = sbase + lane_c;
      const uint32_t gx[4] = {gb + (c7h ^ 0u), gb + (c7h ^ 32u), gb + (c7h ^ 64u), gb + (c7h ^ 96u)};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const bool on = i < 8 ? 2 * i < cA : 2 * (i - 8) < cB;             // warp-uniform
        if (on) cp_async16(gx[i & 3] + (uint32_t)((i >> 3) * (2 * XT) + ((i >> 2) & 1) * 1024 + (i & 3) * 256), src0 + (size_t)(unsigned)jr[i] * 256);
      }
      // the lane's P row (48 B) travels through its own (still unused) q row of the slot: read back in stage 2
      if (it.has()) {
        const uint8_t* pr = reinterpret_cast<const uint8_t*>(P + (size_t)(unsigned)it.j * PROW);
        // chunk c of the row at position c ^ qx (the q tile's own swizzle): the 64-byte pitch alone puts every other lane on the same banks
        const uint32_t pdst = sbase + qrow;
        cp_async16(pdst + ((0u ^ qx) << 4), pr);
        cp_async16(pdst + ((1u ^ qx) << 4), pr + 16);
        cp_async16(pdst + ((2u ^ qx) << 4), pr + 32);
      }
      cp_async_commit();
      TLW(0, ITEM_P(it) - 128, 2);
    };
    // stage 2: soft assignments of the lane's slot -> q tile; the item's gathers have landed -> hand the slot to the MMA warp
    auto finish = [&](const Item& it, bool newer_in_flight) {
      const uint32_t slot = it.slot();
      TLW(0, ITEM_P(it) - 128, 3);
      uint8_t* sl_base = ring + slot * SLOT_BYTES;
      if (newer_in_flight) cp_async_wait<1>();
      else cp_async_wait<0>();
      TLW(0, ITEM_P(it) - 128, 5);
      __syncwarp();            // the self slot's P row (lane 0 / 16 of the node's half warp) was written by another lane's copy
      const uint8_t* pq = sl_base + qrow;
      const float4 Pc0 = *reinterpret_cast<const float4*>(pq + ((0u ^ qx) << 4)), Pc1 = *reinterpret_cast<const float4*>(pq + ((1u ^ qx) << 4));
      const float pc8 = *reinterpret_cast<const float*>(pq + ((2u ^ qx) << 4));
      float4 Pi0, Pi1;
      float pi8;
      if (it.first()) {      // slot 0 of round 0 is the node itself
        const float4* ps = reinterpret_cast<const float4*>(sl_base + 4 * XT + (uint32_t)half * 1024);
        Pi0 = ps[0];
        Pi1 = ps[1];
        pi8 = reinterpret_cast<const float*>(ps)[8];
      } else {
        const float4* pr = reinterpret_cast<const float4*>(P + (size_t)(unsigned)it.self * PROW);
        Pi0 = __ldg(pr);
        Pi1 = __ldg(pr + 1);
        pi8 = __ldg(reinterpret_cast<const float*>(pr) + 8);
      }
      __syncwarp();            // every lane has read the P rows before any q row overwrites them
      const float pj[H] = {Pc0.x, Pc0.y, Pc0.z, Pc0.w, Pc1.x, Pc1.y, Pc1.z, Pc1.w, pc8};
      const float pi[H] = {Pi0.x, Pi0.y, Pi0.z, Pi0.w, Pi1.x, Pi1.y, Pi1.z, Pi1.w, pi8};
      const float4 c0 = *reinterpret_cast<const float4*>(chs), c1 = *reinterpret_cast<const float4*>(chs + 4);
      const float cc[H] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, chs[8]};
      float l[H];
      float m = -INFINITY;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = (pj[h] - pi[h]) + cc[h];
        m = fmaxf(m, l[h]);
      }
      const float mb = -m * 1.4426950408889634f;
      float sum = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(l[h], 1.4426950408889634f, mb)));
        l[h] = e;
        sum += e;
      }
      const float inv = it.has() ? __fdividef(it.inv_d, sum) : 0.f;     // empty slots: q = 0 (their P registers hold finite leftovers)
#pragma unroll
      for (int h = 0; h < H; ++h) l[h] *= inv;
      uint32_t qh[5], ql[5];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t h2 = pack_bf16x2(l[2 * i], l[2 * i + 1]);
        qh[i] = h2;
        ql[i] = pack_bf16x2(l[2 * i] - __uint_as_float(h2 << 16), l[2 * i + 1] - __uint_as_float(h2 & 0xffff0000u));
      }
      qh[4] = pack_bf16x2(l[8], 0.f);
      ql[4] = pack_bf16x2(l[8] - __uint_as_float(qh[4] << 16), 0.f);
      uint8_t* qr = sl_base + qrow;
      *reinterpret_cast<uint4*>(qr + ((0u ^ qx) << 4)) = make_uint4(qh[0], qh[1], qh[2], qh[3]);
      *reinterpret_cast<uint4*>(qr + ((1u ^ qx) << 4)) = make_uint4(qh[4], 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(qr + ((2u ^ qx) << 4)) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
      *reinterpret_cast<uint4*>(qr + ((3u ^ qx) << 4)) = make_uint4(ql[4], 0u, 0u, 0u);
      TLW(0, ITEM_P(it) - 128, 4);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
      TLW(0, ITEM_P(it) - 128, 6);
      PROGRESS(40 + g, it.seq + 1);
    };

    // three-stage software pipeline, one call site per stage (the loop body has to stay inside the instruction cache):
    //   iteration k: index loads of item k | gathers + P loads of item k-1 | soft assignments + hand-over of item k-2
#ifdef TCAGG_SEQ
    // one item at a time per warp: latency is hidden by the number of producer warps instead of a software pipeline
#pragma unroll 1
    for (;;) {
      const Item n = next_item();
      if (!n.valid()) break;
      begin(n);
      finish(n, false);
    }
#else
    Item a{}, b{};                     // a: indices in flight; b: gathers in flight
#pragma unroll 1
    for (;;) {
      TLW(4, ITEM_P(a) - 128, 0);        // stamps of the iteration that begins item a
      const Item n = next_item();
      TLW(4, ITEM_P(a) - 128, 1);
      // a warp never blocks on a ring slot while it holds an item that the MMA warp is waiting for (the slot's previous
      // occupant may be behind that very item in the MMA warp's order when pairs take several rounds)
      const bool early = a.valid() && b.valid() && !(occupant_issued(a.seq) && phase_done(&xfree[a.slot() >> 1], a.par()));
      TLW(4, ITEM_P(a) - 128, 2);
      if (early) TLW(4, ITEM_P(a) - 128, 3);
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        if (a.valid() && pass == (early ? 1 : 0)) begin(a);
        if (b.valid() && pass == 0) finish(b, a.valid() && !early);
      }
      TLW(4, ITEM_P(a) - 128, 4);
      b = a;
      a = n;
      if (!a.valid() && !b.valid()) break;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AGG_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

}  // namespace tcagg

// ---------------------------------------------------------------------------------------------------------------------
struct TcaggWs {
  uint8_t* Xs;
  float* P;
  uint32_t* Wp;
};
template <class C>
static void carve_tcagg(C& c, int64_t n_src, TcaggWs* out) {
  uint8_t* Xs = c.template take<uint8_t>((size_t)n_src * 256);
  float* P = c.template take<float>((size_t)n_src * tcagg::PROW);
  uint32_t* Wp = c.template take<uint32_t>((size_t)36 * 64 * 8);
  if (out) *out = TcaggWs{Xs, P, Wp};
}
struct NullCarverTa {
  Sizer s;
  template <typename T>
  T* take(size_t n) { s.take<T>(n); return nullptr; }
};

size_t feast_fwd_tcagg_ws_bytes(int64_t n_src) {
  NullCarverTa c;
  carve_tcagg(c, n_src, nullptr);
  return c.s.total();
}

bool feast_tcagg_supported(int c_in, int c_out, int64_t ldx, int64_t ldo, const float* x, int64_t n_src) {
  return c_in == tcagg::C_IN && c_out == tcagg::C_OUT && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
         n_src > 0 && n_src < ((int64_t)1 << 24);      // 32-bit byte offsets into the 256-byte split rows
}

int feast_fwd_tcagg(const float* x, int64_t ldx, int64_t N, const int32_t* rowptr, const int32_t* nbr, const int32_t* row_map, int64_t n_src,
                    const float* W, const float* U, const float* c, const float* bias, float act_slope, float* out, int64_t ldo, bool reuse_ws,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < feast_fwd_tcagg_ws_bytes(n_src)) {
    set_error("feast_fwd (tcagg): workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  Carver cv(ws, ws_bytes);
  TcaggWs Wk;
  carve_tcagg(cv, n_src, &Wk);
  if (!reuse_ws) {
    tcagg::prep_w_kernel<<<(36 * 64 * 8 + 255) / 256, 256, 0, st>>>(W, Wk.Wp);
    GEOBI_LAUNCH_OK("tcagg prep_w");
    tcagg::prep_x_kernel<<<(unsigned)cdiv(n_src, tcagg::PREP_ROWS), tcagg::PREP_THREADS, 0, st>>>(x, ldx, n_src, U, Wk.Xs, Wk.P);
    GEOBI_LAUNCH_OK("tcagg prep_x");
  }
  static const int sms = []() -> int {               // thread-safe one-time set-up (see feast_fused.cu)
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(tcagg::feast_tcagg_64_32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcagg::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(tcagg::feast_tcagg_64_32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcagg::SMEM_BYTES) != cudaSuccess)
      return -1;
    return n;
  }();
  if (sms <= 0) {
    (void)cudaGetLastError();
    set_error("feast_tcagg: device query or shared-memory opt-in failed");
    return GEOBI_ERR_CUDA;
  }
  const int64_t n_tiles = (N + tcagg::TILE - 1) / tcagg::TILE;
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  if (row_map)
    tcagg::feast_tcagg_64_32_kernel<true><<<grid, tcagg::THREADS, tcagg::SMEM_BYTES, st>>>(Wk.Xs, Wk.P, N, rowptr, nbr, row_map, c, Wk.Wp, bias,
                                                                                           act_slope, out, ldo);
  else
    tcagg::feast_tcagg_64_32_kernel<false><<<grid, tcagg::THREADS, tcagg::SMEM_BYTES, st>>>(Wk.Xs, Wk.P, N, rowptr, nbr, row_map, c, Wk.Wp, bias,
                                                                                            act_slope, out, ldo);
  GEOBI_LAUNCH_OK("feast_tcagg");
  return GEOBI_OK;
}

#ifdef TCAGG_TIMELINE
extern "C" __attribute__((visibility("default"))) int geobi_debug_tcagg_timeline(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, tcagg::g_tl, sizeof(long long) * 5 * 64 * 8);
}
#endif
#ifdef TCAGG_DEBUG
extern "C" __attribute__((visibility("default"))) int geobi_debug_tcagg(unsigned int* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, tcagg::g_dbg, sizeof(unsigned int) * 64);
}
#endif

}  // namespace geobi
