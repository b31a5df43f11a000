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
// Warp roles (640 threads, 1 CTA / SM, persistent over contiguous tiles of 32 nodes):
//   warps 0-3    epilogue: projection accumulator -> +bias, leaky_relu -> out; also load W into TMEM at start
//   warps 4-11   drain: aggregation accumulators -> Z operand tiles (two sets of 4 quadrant warps, alternating node pairs)
//   warp  12     MMA issue (one elected lane)
//   warps 13-19  producers: indices, cp.async gathers, soft assignments; two node pairs in flight per warp
#include "tc.cuh"

namespace geobi {
namespace tcagg {

using namespace tc;

constexpr int C_IN = 64, C_OUT = 32;
constexpr int TILE = 32;              // nodes per projection tile = MMA N
constexpr int PAIRS = TILE / 2;
constexpr int G = 7;                  // producer warps
constexpr int R = 2 * G;              // ring slots, one node pair (2 x (x_hi, x_lo, q)) each
constexpr int DS = 9;                 // aggregation accumulator slots in TMEM (one node pair each)
constexpr int PIPE = 8;               // pairs of tile t issued ahead of the projection of tile t-1
constexpr int EPI_WARPS = 4, DRAIN_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS + DRAIN_WARPS;
constexpr int PROD_WARP0 = MMA_WARP + 1;
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
constexpr int PROW = 20;                       // floats per row of the head projections: hi[9], 0, lo[9], 0

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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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

// non-blocking probe of an mbarrier phase (warp-uniform answer: lane 0 tests, the result is broadcast)
__device__ __forceinline__ bool phase_done(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return __shfl_sync(0xffffffffu, done, 0) != 0;
}

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
    if ((it & 1023) == 1023) {
      if (*(volatile unsigned int*)&g_dbg[63]) return;
      if (it > (1ll << 18)) {
        if ((threadIdx.x & 31) == 0) g_dbg[(blockIdx.x == 0 ? 0 : 32) + (threadIdx.x >> 5)] = (unsigned)tag | (parity << 16) | 0x80000000u;
        *(volatile unsigned int*)&g_dbg[63] = 1;
        return;
      }
    }
  }
}
#define WAIT(bar, parity, tag) wait_dbg(bar, parity, tag)
#else
#define WAIT(bar, parity, tag) mbar_wait(bar, parity)
#endif

// ---------------------------------------------------------------------------------------------------------------------
// prep: x (fp32 rows) -> Xs[row] = bf16 hi[64] | lo[64]  (256 B)  and  P[row] = {hi[9], 0, lo[9], 0} of the fp64 head projections
// u_h . x_row  (double-float pairs: the soft assignments need P_j - P_i to fp32 accuracy of the DIFFERENCE, feast.cu).
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
  {
    float hi = (float)a0;
    pr[part] = hi; pr[10 + part] = (float)(a0 - (double)hi);
    hi = (float)a1;
    pr[part + 4] = hi; pr[10 + part + 4] = (float)(a1 - (double)hi);
    if (part == 0) {
      hi = (float)a2;
      pr[8] = hi; pr[18] = (float)(a2 - (double)hi);
    } else if (part == 1) {
      pr[9] = 0.f; pr[19] = 0.f;
    }
  }
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
  __shared__ __align__(8) uint64_t full[R], xfree[R], dfull[DS], dfree[DS], zfull, zfree, ofull[2], ofree[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float chs[12];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* z_hi = sm;
  uint8_t* z_lo = sm + ZPLANE;
  uint8_t* ring = sm + 2 * ZPLANE;
  float* stage = reinterpret_cast<float*>(ring + R * SLOT_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < R; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&xfree[i], 1);
    }
    for (int i = 0; i < DS; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dfree[i], 4);
    }
    mbar_init(&zfull, PAIRS * 4);
    mbar_init(&zfree, 1);
    mbar_init(&ofull[0], 1);
    mbar_init(&ofull[1], 1);
    mbar_init(&ofree[0], EPI_WARPS);
    mbar_init(&ofree[1], EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid < H) chs[tid] = cvec[tid];
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
    int rounds, excl;   // lane = pair (lanes 0-15): ring slots of the pair and their exclusive prefix inside the tile
    int total;
  };
  auto tile_info = [&](int64_t tile, int rb_in, int re_in) {
    TileInfo ti;
    const int64_t node = tile * TILE + lane;
    ti.rb = rb_in;
    ti.d = node < N ? re_in - rb_in + 1 : 0;
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

  if (warp == MMA_WARP) {
    // ================================================== MMA issue ==================================================
    constexpr uint32_t IA = idesc_of(64, 32, 1, 1), IB = idesc_of(64, 16, 1, 1), IP = idesc_of(64, 32, 0, 0);
    constexpr uint32_t XHI = desc_hi(64, 2), QHI = desc_hi(32, 4), ZHI = desc_hi(64, 2);
    const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFFu) >> 4;
    const uint32_t zh_lo = ((smem_u32(z_hi) & 0x3FFFFu) >> 4) | (1u << 16), zl_lo = ((smem_u32(z_lo) & 0x3FFFFu) >> 4) | (1u << 16);
    int rb_n, re_n;
    load_rowptr(t_begin, rb_n, re_n);
    uint32_t seq_base = 0;
    TileInfo ti{};
    auto do_pair = [&](uint32_t p, int pp) {
      const uint32_t dslot = p % DS, dk = p / DS;
      const int rounds = __shfl_sync(0xffffffffu, ti.rounds, pp);
      const uint32_t seq0 = seq_base + (uint32_t)__shfl_sync(0xffffffffu, ti.excl, pp);
      for (int r = 0; r < rounds; ++r) {
        const uint32_t seq = seq0 + r, slot = seq % R, k = seq / R;
        WAIT(&full[slot], k & 1, 1);
        if (r == 0) WAIT(&dfree[dslot], (dk & 1) ^ 1, 2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t xa = ring_lo + slot * (SLOT_BYTES >> 4), qa = xa + (4 * XT >> 4);
          const uint32_t dA = tmem + COL_D + dslot * 32, dB = dA + HALF;
          mma_ss(dA, xa, XHI, qa, QHI, IA, r > 0);
          mma_ss(dA, xa + (XT >> 4), XHI, qa, QHI, IB, 1u);
          mma_ss(dB, xa + (2 * XT >> 4), XHI, qa + (1024 >> 4), QHI, IA, r > 0);
          mma_ss(dB, xa + (3 * XT >> 4), XHI, qa + (1024 >> 4), QHI, IB, 1u);
          mma_commit(&xfree[slot]);
          if (r == rounds - 1) mma_commit(&dfull[dslot]);
        }
        __syncwarp();
      }
    };
    auto do_proj = [&](int t) {
      const uint32_t b = t & 1;
      WAIT(&zfull, t & 1, 3);
      WAIT(&ofree[b], ((t >> 1) & 1) ^ 1, 4);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint32_t d = tmem + COL_O + 32 * b + hf * HALF, a0 = tmem + COL_W + hf * HALF;
#pragma unroll 2
          for (int i = 0; i < 18; ++i) {
            const int ks = 18 * hf + i;
            const uint32_t zoff = (uint32_t)(ks >> 2) * (ZCHUNK >> 4) + 2 * (ks & 3);
            mma_ts(d, a0 + 8 * i, zh_lo + zoff, ZHI, IP, i > 0);
            mma_ts(d, a0 + 8 * i, zl_lo + zoff, ZHI, IP, 1u);
          }
        }
        mma_commit(&zfree);
        mma_commit(&ofull[b]);
      }
      __syncwarp();
    };
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        ti = tile_info(t_begin + t, rb_n, re_n);
        load_rowptr(t_begin + t + 1, rb_n, re_n);
        for (int pp = 0; pp < PIPE; ++pp) do_pair((uint32_t)(PAIRS * t + pp), pp);
      }
      if (t >= 1) do_proj(t - 1);
      if (t < T) {
        for (int pp = PIPE; pp < PAIRS; ++pp) do_pair((uint32_t)(PAIRS * t + pp), pp);
        seq_base += (uint32_t)ti.total;
      }
    }
  } else if (warp < EPI_WARPS) {
    // ================================================== epilogue ==================================================
    // quadrant q: lanes 0-15 = accumulator of K half 0, lanes 16-31 = K half 1; rows 16 q + l: q = 0, 1 -> W_hi rows of outputs
    // 0-15 / 16-31, q = 2, 3 -> the W_lo rows of the same outputs
    const int o = 16 * (warp & 1) + (lane & 15);
    const float my_bias = bias[o];
    for (int t = 0; t < T; ++t) {
      const uint32_t b = t & 1;
      WAIT(&ofull[b], (t >> 1) & 1, 5);
      tc_fence_after();
      float v[TILE];
      tmem_ld32(tmem + COL_O + 32 * b + ((uint32_t)(warp * 32) << 16), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ofree[b]);
#pragma unroll
      for (int n = 0; n < TILE; ++n) v[n] += __shfl_xor_sync(0xffffffffu, v[n], 16);
      float* sb = stage + (t & 1) * (TILE * C_OUT);
      if (warp >= 2 && lane < 16) {
#pragma unroll
        for (int n = 0; n < TILE; ++n) sb[n * C_OUT + o] = v[n];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp < 2 && lane < 16) {
        const int64_t n0 = (t_begin + t) * TILE;
#pragma unroll
        for (int n = 0; n < TILE; ++n) {
          if (n0 + n < N) {
            float r = v[n] + sb[n * C_OUT + o] + my_bias;
            r = r > 0.f ? r : r * slope;
            out[(n0 + n) * ldo + o] = r;
          }
        }
      }
    }
  } else if (warp < MMA_WARP) {
    // ================================================== drain ==================================================
    const int set = (warp - EPI_WARPS) >> 2, qd = warp & 3;
    const int half = lane >> 4, c = 16 * qd + (lane & 15);
    const int npairs = PAIRS * T;
    // this thread's byte offset inside a Z row: head pairs -> 32-bit word of channel c in K block 2 hp + (c >= 32); head 8 -> K block 8
    const uint32_t ch16 = (uint32_t)(c & 31) >> 2, inner = (uint32_t)(c & 3) * 4;
    const uint32_t ch16_8 = (uint32_t)c >> 3, inner_8 = (uint32_t)(c & 7) * 2;
    for (int p = set; p < npairs; p += 2) {
      const int t = p >> 4, pp = p & 15;
      const uint32_t dslot = (uint32_t)p % DS, dk = (uint32_t)p / DS;
      WAIT(&dfull[dslot], dk & 1, 6);
      tc_fence_after();
      uint32_t rr[32];
      tmem_ld32_issue(tmem + COL_D + dslot * 32 + ((uint32_t)(qd * 32) << 16), rr);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dfree[dslot]);
      float z[H];
#pragma unroll
      for (int h = 0; h < H; ++h) z[h] = __uint_as_float(rr[h]) + __uint_as_float(rr[16 + h]);
      if (pp < 2 && t >= 1) {       // first pair of this warp in tile t: the projection of tile t-1 must have finished reading Z
        WAIT(&zfree, (t - 1) & 1, 7);
        tc_fence_after();
      }
      const int row = (pp >> 2) * 8 + (pp & 3) + 4 * half;
      const uint32_t rbase = (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128;
      const uint32_t off = (uint32_t)(c >> 5) * ZCHUNK + rbase + ((ch16 ^ (uint32_t)(row & 7)) << 4) + inner;
#pragma unroll
      for (int hp = 0; hp < 4; ++hp) {
        const float a = z[2 * hp], b = z[2 * hp + 1];
        const uint32_t h2 = pack_bf16x2(a, b);
        const float fa = __uint_as_float(h2 << 16), fb = __uint_as_float(h2 & 0xffff0000u);
        const uint32_t l2 = pack_bf16x2(a - fa, b - fb);
        *reinterpret_cast<uint32_t*>(z_hi + 2 * hp * ZCHUNK + off) = h2;
        *reinterpret_cast<uint32_t*>(z_lo + 2 * hp * ZCHUNK + off) = l2;
      }
      {
        const uint32_t off8 = 8 * ZCHUNK + rbase + ((ch16_8 ^ (uint32_t)(row & 7)) << 4) + inner_8;
        const __nv_bfloat16 bh = __float2bfloat16_rn(z[8]);
        const __nv_bfloat16 bl = __float2bfloat16_rn(z[8] - __bfloat162float(bh));
        *reinterpret_cast<__nv_bfloat16*>(z_hi + off8) = bh;
        *reinterpret_cast<__nv_bfloat16*>(z_lo + off8) = bl;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&zfull);
    }
  } else {
    // ================================================== producers ==================================================
    const int g = warp - PROD_WARP0;
    const int half = lane >> 4, sl = lane & 15;
    const uint32_t ring_u32 = smem_u32(ring);
    // work items: (tile t, pair pp, round r) with (16 t + pp) % G == g, in order
    struct Item {
      int valid;
      uint32_t seq;     // ring sequence number
      int r;            // round
      int has;          // this lane's slot is a real neighbour / self slot
      int j;            // source row of this lane's slot
      int self;         // source row of this lane's node
      float inv_d;
    };
    int it_t = -1, it_pp = 0, it_r = 0, it_rounds = 0;
    uint32_t seq_base = 0, next_base = 0;
    TileInfo ti{};
    int rb_n, re_n;
    load_rowptr(t_begin, rb_n, re_n);
    // first own pair: global pair index p = g
    int64_t p_next = g;
    auto next_item = [&]() {
      Item it{};
      if (it_t >= 0 && it_r + 1 < it_rounds) {
        ++it_r;
      } else {
        if (p_next >= (int64_t)PAIRS * T) return it;     // valid = 0
        const int t = (int)(p_next >> 4);
        while (it_t < t) {                               // enter the next tile(s): every warp walks every tile
          ++it_t;
          seq_base = next_base;
          ti = tile_info(t_begin + it_t, rb_n, re_n);
          load_rowptr(t_begin + it_t + 1, rb_n, re_n);
          next_base = seq_base + (uint32_t)ti.total;
        }
        it_pp = (int)(p_next & 15);
        it_r = 0;
        it_rounds = __shfl_sync(0xffffffffu, ti.rounds, it_pp);
        p_next += G;
      }
      const int row = (it_pp >> 2) * 8 + (it_pp & 3) + 4 * half;
      const int rb = __shfl_sync(0xffffffffu, ti.rb, row), d = __shfl_sync(0xffffffffu, ti.d, row);
      const int64_t node = (t_begin + it_t) * TILE + row;
      it.valid = 1;
      it.seq = seq_base + (uint32_t)__shfl_sync(0xffffffffu, ti.excl, it_pp) + (uint32_t)it_r;
      it.r = it_r;
      const int s = 16 * it_r + sl;
      it.has = s < d;
      it.inv_d = d > 0 ? 1.0f / (float)d : 0.f;
      int self = d > 0 ? (int)node : 0;
      int j = self;
      if (it.has && s > 0) j = __ldg(nbr + rb + s - 1);
      if (HAS_MAP) {
        if (d > 0) self = __ldg(row_map + self);
        j = (it.has && s > 0) ? __ldg(row_map + j) : self;
      }
      it.j = j;
      it.self = self;
      return it;
    };
    // stage 1 of an item: ring slot free -> gathers in flight, P row of this lane's slot in flight
    auto begin = [&](const Item& it, float4 (&Pn)[5]) {
      const uint32_t slot = it.seq % R, k = it.seq / R;
      WAIT(&xfree[slot], (k & 1) ^ 1, 8);
      tc_fence_after();
      const uint32_t sbase = ring_u32 + slot * SLOT_BYTES;
      const unsigned mask = __ballot_sync(0xffffffffu, it.has);
      const uint32_t ch = (uint32_t)sl;                   // 16-byte chunk of the 256-byte source row: 0-7 hi plane, 8-15 lo plane
      const uint32_t lane_part = (ch >> 3) * XT;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if ((mask >> (2 * i)) & 3u) {                     // warp-uniform
          const int L = 2 * i + half;                     // slot-lane served by this half warp
          const int jj = __shfl_sync(0xffffffffu, it.j, L);
          if ((mask >> L) & 1u) {
            const uint32_t s = (uint32_t)(L & 15);
            const uint32_t dst = sbase + (uint32_t)(L >> 4) * (2 * XT) + lane_part + (s >> 3) * 1024 + (s & 7) * 128 + (((ch & 7) ^ (s & 7)) << 4);
            cp_async16(dst, Xs + (size_t)(unsigned)jj * 256 + ch * 16);
          }
        }
      }
      cp_async_commit();
      if (it.has) {
        const float4* pr = reinterpret_cast<const float4*>(P + (size_t)(unsigned)it.j * PROW);
#pragma unroll
        for (int i = 0; i < 5; ++i) Pn[i] = __ldg(pr + i);
      }
    };
    // stage 2: soft assignments of the lane's slot -> q tile; the item's gathers have landed -> hand the slot to the MMA warp
    auto finish = [&](const Item& it, const float4 (&Pc)[5], bool newer_in_flight) {
      const uint32_t slot = it.seq % R;
      uint8_t* qt = ring + slot * SLOT_BYTES + 4 * XT + half * 1024;
      float4 Pi[5];
      if (it.r == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          Pi[i].x = __shfl_sync(0xffffffffu, Pc[i].x, half * 16);
          Pi[i].y = __shfl_sync(0xffffffffu, Pc[i].y, half * 16);
          Pi[i].z = __shfl_sync(0xffffffffu, Pc[i].z, half * 16);
          Pi[i].w = __shfl_sync(0xffffffffu, Pc[i].w, half * 16);
        }
      } else {
        const float4* pr = reinterpret_cast<const float4*>(P + (size_t)(unsigned)it.self * PROW);
#pragma unroll
        for (int i = 0; i < 5; ++i) Pi[i] = __ldg(pr + i);
      }
      uint32_t qh[5], ql[5];
      if (it.has) {
        const float pj[PROW] = {Pc[0].x, Pc[0].y, Pc[0].z, Pc[0].w, Pc[1].x, Pc[1].y, Pc[1].z, Pc[1].w, Pc[2].x, Pc[2].y,
                                Pc[2].z, Pc[2].w, Pc[3].x, Pc[3].y, Pc[3].z, Pc[3].w, Pc[4].x, Pc[4].y, Pc[4].z, Pc[4].w};
        const float pi[PROW] = {Pi[0].x, Pi[0].y, Pi[0].z, Pi[0].w, Pi[1].x, Pi[1].y, Pi[1].z, Pi[1].w, Pi[2].x, Pi[2].y,
                                Pi[2].z, Pi[2].w, Pi[3].x, Pi[3].y, Pi[3].z, Pi[3].w, Pi[4].x, Pi[4].y, Pi[4].z, Pi[4].w};
        float l[H];
        float m = -INFINITY;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          l[h] = ((pj[h] - pi[h]) + (pj[10 + h] - pi[10 + h])) + chs[h];
          m = fmaxf(m, l[h]);
        }
        float sum = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          l[h] = __expf(l[h] - m);
          sum += l[h];
        }
        const float inv = it.inv_d / sum;
#pragma unroll
        for (int h = 0; h < H; ++h) l[h] *= inv;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t h2 = pack_bf16x2(l[2 * i], l[2 * i + 1]);
          qh[i] = h2;
          ql[i] = pack_bf16x2(l[2 * i] - __uint_as_float(h2 << 16), l[2 * i + 1] - __uint_as_float(h2 & 0xffff0000u));
        }
        const uint32_t h8 = pack_bf16x2(l[8], 0.f);
        qh[4] = h8;
        ql[4] = pack_bf16x2(l[8] - __uint_as_float(h8 << 16), 0.f);
      } else {
#pragma unroll
        for (int i = 0; i < 5; ++i) qh[i] = ql[i] = 0u;
      }
      // row sl of the SWIZZLE_64B tile: 64 B = chunks {q_hi h0-7, q_hi h8 + zeros, q_lo h0-7, q_lo h8 + zeros}
      uint8_t* qrow = qt + (sl >> 3) * 512 + (sl & 7) * 64;
      const uint32_t x = (uint32_t)(sl >> 1) & 3u;
      *reinterpret_cast<uint4*>(qrow + ((0u ^ x) << 4)) = make_uint4(qh[0], qh[1], qh[2], qh[3]);
      *reinterpret_cast<uint4*>(qrow + ((1u ^ x) << 4)) = make_uint4(qh[4], 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(qrow + ((2u ^ x) << 4)) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
      *reinterpret_cast<uint4*>(qrow + ((3u ^ x) << 4)) = make_uint4(ql[4], 0u, 0u, 0u);
      if (newer_in_flight) cp_async_wait<1>();
      else cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
    };

    Item cur = next_item();            // indices loaded; gathers not yet issued
    Item prev{};
    float4 Pa[5], Pb[5];               // P rows of the item in stage 2 / of the item entering stage 1
#pragma unroll
    for (int i = 0; i < 5; ++i) Pa[i] = Pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    while (cur.valid || prev.valid) {
      Item nxt{};
      if (cur.valid) nxt = next_item();          // index loads of the following item go out first
      // a warp never blocks on a ring slot while it holds an item that the MMA warp is waiting for (the slot's previous
      // occupant may be behind that very item in the MMA warp's order when pairs take several rounds)
      if (cur.valid && prev.valid && !phase_done(&xfree[cur.seq % R], ((cur.seq / R) & 1) ^ 1)) {
        finish(prev, Pa, false);
        prev.valid = 0;
      }
      if (cur.valid) begin(cur, Pb);
      if (prev.valid) finish(prev, Pa, cur.valid != 0);
#pragma unroll
      for (int i = 0; i < 5; ++i) Pa[i] = Pb[i];
      prev = cur;
      cur = nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
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
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    GEOBI_CUDA_OK(cudaGetDevice(&dev));
    GEOBI_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GEOBI_CUDA_OK(cudaFuncSetAttribute(tcagg::feast_tcagg_64_32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcagg::SMEM_BYTES));
    GEOBI_CUDA_OK(cudaFuncSetAttribute(tcagg::feast_tcagg_64_32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcagg::SMEM_BYTES));
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

#ifdef TCAGG_DEBUG
extern "C" __attribute__((visibility("default"))) int geobi_debug_tcagg(unsigned int* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, tcagg::g_dbg, sizeof(unsigned int) * 64);
}
#endif

}  // namespace geobi
