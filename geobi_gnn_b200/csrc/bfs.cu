// BFS face-patch splitter on the device (the reference grows patches with pure-Python loops:
// /root/reference/code/data_util.py:55-84 mesh_get_neighbor_np, dataset.py:156-193 the split; host C++ twin: host_patch.cpp).
//
// Reference semantics, per patch: out = [seed]; ring by ring, for every face of the previous ring in list order, for its three
// corners in order, for the corner's incident faces in vf-row order: append the face if this patch has not taken it yet; stop at
// exactly `neighbor_count` faces.  Between patches the covered faces leave the candidate set and the next seed is the FIRST arg-max
// of the squared centre-to-centroid distance over what is left.
//
// Parallel form with the same discovery order.  Number the (face q of the ring, corner c, row slot t) triples of a ring in that
// nesting order: key = ((q - lo) * 3 + c) * K + t.  A face new to the patch is appended at the position of its SMALLEST key, and
// the serial loop appends the new faces in increasing key order.  So per ring:
//   bfs_claim    every triple whose face is not in the patch yet does atomicMin(claim[face], ring tag | key); the tag (the bitwise
//                complement of a global ring counter, in the high word) makes any value of an earlier ring lose, so `claim` is never
//                cleared;
//   bfs_count    blocks own contiguous ranges of (q, c) pairs; a triple wins if claim[face] holds its own key; winners per block;
//   bfs_scatter  block offset = sum of the earlier blocks' counts, ordered block scan inside, winners go to out[n + rank] while
//                rank stays below the cut, and block 0 writes the next ring's bounds (ping-pong state, so no kernel races with it).
// The vertex stamps of the host version only skip work (a vertex expanded earlier has no untaken face left) and are not needed.
// Rings are launched in batches of 32 (96 launches, no-ops once the patch is complete) with one state read-back per batch; the
// cover / next-seed step is three more kernels.  Everything is integer work or exact fp32 (explicit round-to-nearest operations in
// numpy's order for the distances): the patches are identical to the host splitter's (tests/test_gpu_patches.py).
#include "common.cuh"

namespace geobi {
namespace bfs {

constexpr int GRID = 148;
constexpr int THREADS = 256;
constexpr int RINGS_PER_BATCH = 32;

struct State {
  int lo, hi;          // the ring to expand next: out[lo, hi)
  int done;            // patch complete (cut reached or nothing new)
  uint32_t ring;       // global ring counter (never reset)
  uint32_t epoch;      // patch counter: fstamp[f] == epoch <=> f is in the current patch
  int n_left;          // faces no patch has covered yet
  int next_seed;       // result of the last cover step (-1: everything covered)
  int pad;
};

struct Ws {
  float* d2_left;                  // [F] squared centre distance; -inf once covered
  uint32_t* fstamp;                // [F]
  unsigned long long* claim;       // [F]
  int* block_sums;                 // [GRID]
  float* part_val;                 // [GRID]
  int* part_idx;                   // [GRID]
  int* fresh;                      // [1]
  State* st;                       // [2] ping-pong
};
template <class C>
static void carve(C& c, int64_t F, Ws* out) {
  Ws w;
  w.d2_left = c.template take<float>((size_t)F);
  w.fstamp = c.template take<uint32_t>((size_t)F);
  w.claim = c.template take<unsigned long long>((size_t)F);
  w.block_sums = c.template take<int>(GRID);
  w.part_val = c.template take<float>(GRID);
  w.part_idx = c.template take<int>(GRID);
  w.fresh = c.template take<int>(1);
  w.st = c.template take<State>(2);
  if (out) *out = w;
}
struct NullCarver {
  Sizer s;
  template <typename T>
  T* take(size_t n) { s.take<T>(n); return nullptr; }
};

// ((pts[fv].mean(1) - centroid) ** 2).sum(1) in numpy's fp32 order (dataset.py:165-166): ((a + b) + c) / 3, subtract, square, (x + y) + z
__global__ void face_d2_kernel(const float* __restrict__ pts, const int64_t* __restrict__ fv, int64_t F, float cx, float cy, float cz,
                               float* __restrict__ out) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float* a = pts + 3 * fv[3 * f];
  const float* b = pts + 3 * fv[3 * f + 1];
  const float* c = pts + 3 * fv[3 * f + 2];
  const float cen[3] = {cx, cy, cz};
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float m = __fdiv_rn(__fadd_rn(__fadd_rn(a[k], b[k]), c[k]), 3.0f);
    const float e = __fsub_rn(m, cen[k]);
    d[k] = __fmul_rn(e, e);
  }
  out[f] = __fadd_rn(__fadd_rn(d[0], d[1]), d[2]);
}

// first arg-max (np.argmax: the lowest index among equal maxima): per-block partials, then one block over the partials
__device__ __forceinline__ void better(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__global__ void __launch_bounds__(THREADS) argmax_partial_kernel(const float* __restrict__ x, int64_t F, float* __restrict__ part_val,
                                                                 int* __restrict__ part_idx) {
  __shared__ float sv[THREADS];
  __shared__ int si[THREADS];
  float v = -INFINITY;
  int idx = 0x7fffffff;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < F; i += (int64_t)gridDim.x * THREADS) better(v, idx, x[i], (int)i);
  sv[threadIdx.x] = v;
  si[threadIdx.x] = idx;
  __syncthreads();
  for (int o = THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      float a = sv[threadIdx.x];
      int b = si[threadIdx.x];
      better(a, b, sv[threadIdx.x + o], si[threadIdx.x + o]);
      sv[threadIdx.x] = a;
      si[threadIdx.x] = b;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part_val[blockIdx.x] = sv[0];
    part_idx[blockIdx.x] = si[0];
  }
}
// one block: final arg-max over the partials -> next_seed of BOTH state copies (-1 when every face is covered); `fresh` (may be
// null) = the number of faces the finished patch newly covered, subtracted from n_left first
__global__ void __launch_bounds__(THREADS) argmax_final_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int parts,
                                                               const int* __restrict__ fresh, State* __restrict__ st, int p) {
  __shared__ float sv[THREADS];
  __shared__ int si[THREADS];
  float v = -INFINITY;
  int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < parts; i += THREADS) better(v, idx, part_val[i], part_idx[i]);
  sv[threadIdx.x] = v;
  si[threadIdx.x] = idx;
  __syncthreads();
  for (int o = THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      float a = sv[threadIdx.x];
      int b = si[threadIdx.x];
      better(a, b, sv[threadIdx.x + o], si[threadIdx.x + o]);
      sv[threadIdx.x] = a;
      si[threadIdx.x] = b;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    State s = st[p];
    if (fresh != nullptr) s.n_left -= *fresh;
    s.next_seed = s.n_left > 0 ? si[0] : -1;
    st[p] = s;
    st[p ^ 1] = s;
  }
}

__global__ void begin_patch_kernel(State* __restrict__ st, int p, int seed, uint32_t* __restrict__ fstamp, int* __restrict__ out) {
  State s = st[p];
  s.epoch += 1;
  s.lo = 0;
  s.hi = 1;
  s.done = 0;
  out[0] = seed;
  fstamp[seed] = s.epoch;
  st[p] = s;
}

// ---- one ring
__global__ void __launch_bounds__(THREADS) bfs_claim_kernel(const State* __restrict__ st, int p, const int64_t* __restrict__ fv,
                                                            const int64_t* __restrict__ vf, int K, const int* __restrict__ out,
                                                            const uint32_t* __restrict__ fstamp, unsigned long long* __restrict__ claim) {
  const State s = st[p];
  if (s.done) return;
  const int pairs = (s.hi - s.lo) * 3;
  const unsigned long long tag = (unsigned long long)(~s.ring) << 32;
  for (int r = blockIdx.x * THREADS + threadIdx.x; r < pairs; r += gridDim.x * THREADS) {
    const int64_t v = fv[(int64_t)out[s.lo + r / 3] * 3 + r % 3];
    const int64_t* row = vf + v * K;
    for (int t = 0; t < K; ++t) {
      const int64_t g = row[t];
      if (g < 0) break;
      if (fstamp[g] != s.epoch) atomicMin(claim + g, tag | (unsigned long long)((uint32_t)r * (uint32_t)K + (uint32_t)t));
    }
  }
}

// winners of pair r (bit t set: slot t appends its face)
__device__ __forceinline__ uint32_t pair_winners(const State& s, unsigned long long tag, int r, const int64_t* __restrict__ fv,
                                                 const int64_t* __restrict__ vf, int K, const int* __restrict__ out,
                                                 const uint32_t* __restrict__ fstamp, const unsigned long long* __restrict__ claim) {
  const int64_t v = fv[(int64_t)out[s.lo + r / 3] * 3 + r % 3];
  const int64_t* row = vf + v * K;
  uint32_t mask = 0;
  for (int t = 0; t < K; ++t) {
    const int64_t g = row[t];
    if (g < 0) break;
    if (fstamp[g] != s.epoch && claim[g] == (tag | (unsigned long long)((uint32_t)r * (uint32_t)K + (uint32_t)t))) mask |= 1u << t;
  }
  return mask;
}

__device__ __forceinline__ int block_sum(int v, int* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int w = 0; w < THREADS / 32; ++w) t += sh[w];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(THREADS) bfs_count_kernel(const State* __restrict__ st, int p, const int64_t* __restrict__ fv,
                                                            const int64_t* __restrict__ vf, int K, const int* __restrict__ out,
                                                            const uint32_t* __restrict__ fstamp, const unsigned long long* __restrict__ claim,
                                                            int* __restrict__ block_sums) {
  __shared__ int sh[THREADS / 32];
  const State s = st[p];
  if (s.done) return;
  const int pairs = (s.hi - s.lo) * 3;
  const int per = (pairs + gridDim.x - 1) / gridDim.x;
  const int b0 = blockIdx.x * per, b1 = min(pairs, b0 + per);
  const unsigned long long tag = (unsigned long long)(~s.ring) << 32;
  int cnt = 0;
  for (int r = b0 + threadIdx.x; r < b1; r += THREADS) cnt += __popc(pair_winners(s, tag, r, fv, vf, K, out, fstamp, claim));
  const int total = block_sum(cnt, sh);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(THREADS) bfs_scatter_kernel(State* __restrict__ st, int p, const int64_t* __restrict__ fv,
                                                              const int64_t* __restrict__ vf, int K, int* __restrict__ out,
                                                              uint32_t* __restrict__ fstamp, const unsigned long long* __restrict__ claim,
                                                              const int* __restrict__ block_sums, int cut) {
  __shared__ int sh[THREADS / 32];
  __shared__ int wsum[THREADS / 32];
  const State s = st[p];
  if (s.done) {
    if (blockIdx.x == 0 && threadIdx.x == 0) st[p ^ 1] = s;
    return;
  }
  // offset of this block and the ring's total
  int before = 0, all = 0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += THREADS) {
    const int v = block_sums[b];
    all += v;
    if (b < (int)blockIdx.x) before += v;
  }
  before = block_sum(before, sh);
  all = block_sum(all, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    State n = s;
    const int grown = min(s.hi + all, cut);
    n.lo = s.hi;
    n.hi = grown;
    n.ring = s.ring + 1;
    n.done = (grown >= cut || all == 0) ? 1 : 0;
    st[p ^ 1] = n;
  }
  const int pairs = (s.hi - s.lo) * 3;
  const int per = (pairs + gridDim.x - 1) / gridDim.x;
  const int b0 = blockIdx.x * per, b1 = min(pairs, b0 + per);
  const unsigned long long tag = (unsigned long long)(~s.ring) << 32;
  int running = s.hi + before;                  // position of the first winner of the current tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = b0; t0 < b1; t0 += THREADS) {   // uniform trip count: every thread walks every tile
    const int r = t0 + threadIdx.x;
    const uint32_t mask = r < b1 ? pair_winners(s, tag, r, fv, vf, K, out, fstamp, claim) : 0u;
    const int cnt = __popc(mask);
    int incl = cnt;                              // ordered block scan: pairs of a tile are in thread order
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int wbase = 0, tile = 0;
    for (int w = 0; w < THREADS / 32; ++w) {
      if (w < warp) wbase += wsum[w];
      tile += wsum[w];
    }
    __syncthreads();
    int pos = running + wbase + incl - cnt;
    if (mask) {
      const int64_t v = fv[(int64_t)out[s.lo + r / 3] * 3 + r % 3];
      const int64_t* row = vf + v * K;
      uint32_t m = mask;
      while (m) {
        const int t = __ffs(m) - 1;
        m &= m - 1;
        if (pos < cut) {
          const int g = (int)row[t];
          out[pos] = g;
          fstamp[g] = s.epoch;
        }
        ++pos;
      }
    }
    running += tile;
  }
}

// faces of the finished patch leave the uncovered set
__global__ void __launch_bounds__(THREADS) bfs_cover_kernel(const State* __restrict__ st, int p, const int* __restrict__ out,
                                                            float* __restrict__ d2_left, int* __restrict__ fresh) {
  __shared__ int sh[THREADS / 32];
  const int n = st[p].hi;
  int cnt = 0;
  for (int i = blockIdx.x * THREADS + threadIdx.x; i < n; i += gridDim.x * THREADS) {
    const int f = out[i];
    if (d2_left[f] != -INFINITY) ++cnt;
    d2_left[f] = -INFINITY;
  }
  const int total = block_sum(cnt, sh);
  if (threadIdx.x == 0 && total) atomicAdd(fresh, total);
}

}  // namespace bfs
}  // namespace geobi

using namespace geobi;

extern "C" size_t geobi_bfs_ws_bytes(int64_t n_faces) {
  if (n_faces < 0) return 0;
  bfs::NullCarver c;
  bfs::carve(c, n_faces, nullptr);
  return c.s.total();
}

static int bfs_read_state(const bfs::Ws& w, int p, bfs::State* host, cudaStream_t st) {
  GEOBI_CUDA_OK(cudaMemcpyAsync(host, w.st + p, sizeof(bfs::State), cudaMemcpyDeviceToHost, st));
  GEOBI_CUDA_OK(cudaStreamSynchronize(st));
  return GEOBI_OK;
}

extern "C" int geobi_bfs_begin(const float* points, const int64_t* fv, int64_t n_faces, const float* centroid_host, int64_t* seed_host,
                               void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(points && fv && centroid_host && seed_host && n_faces > 0 && n_faces < (int64_t)1 << 31, "bfs_begin: bad arguments");
  if (!ws || ws_bytes < geobi_bfs_ws_bytes(n_faces)) { set_error("bfs_begin: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver cv(ws, ws_bytes);
  bfs::Ws w;
  bfs::carve(cv, n_faces, &w);
  GEOBI_CUDA_OK(cudaMemsetAsync(w.fstamp, 0, sizeof(uint32_t) * (size_t)n_faces, st));
  GEOBI_CUDA_OK(cudaMemsetAsync(w.claim, 0xff, sizeof(unsigned long long) * (size_t)n_faces, st));
  bfs::State s{};
  s.n_left = (int)n_faces;
  GEOBI_CUDA_OK(cudaMemcpyAsync(w.st, &s, sizeof(s), cudaMemcpyHostToDevice, st));
  GEOBI_CUDA_OK(cudaStreamSynchronize(st));          // `s` is a stack object
  bfs::face_d2_kernel<<<(unsigned)cdiv(n_faces, 256), 256, 0, st>>>(points, fv, n_faces, centroid_host[0], centroid_host[1], centroid_host[2],
                                                                    w.d2_left);
  bfs::argmax_partial_kernel<<<bfs::GRID, bfs::THREADS, 0, st>>>(w.d2_left, n_faces, w.part_val, w.part_idx);
  bfs::argmax_final_kernel<<<1, bfs::THREADS, 0, st>>>(w.part_val, w.part_idx, bfs::GRID, nullptr, w.st, 0);
  GEOBI_LAUNCH_OK("bfs_begin");
  bfs::State h;
  int rc = bfs_read_state(w, 0, &h, st);
  if (rc) return rc;
  *seed_host = h.next_seed;
  return GEOBI_OK;
}

extern "C" int geobi_bfs_grow(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t k, int64_t seed, int64_t neighbor_count,
                              int32_t* out_faces, int64_t* n_out_host, int64_t* next_seed_host, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(fv && vf && out_faces && n_out_host && next_seed_host && n_faces > 0 && n_faces < (int64_t)1 << 31 && k > 0 && k <= 32,
                "bfs_grow: bad arguments (the incidence table may be at most 32 wide)");
  GEOBI_REQUIRE(seed >= 0 && seed < n_faces && neighbor_count > 0, "bfs_grow: bad seed or patch size");
  if (!ws || ws_bytes < geobi_bfs_ws_bytes(n_faces)) { set_error("bfs_grow: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver cv(ws, ws_bytes);
  bfs::Ws w;
  bfs::carve(cv, n_faces, &w);
  const int cut = (int)(neighbor_count < n_faces ? neighbor_count : n_faces);
  const int K = (int)k;
  int p = 0;                                        // both state copies are equal between calls
  bfs::begin_patch_kernel<<<1, 1, 0, st>>>(w.st, p, (int)seed, w.fstamp, out_faces);
  bfs::State h{};
  for (;;) {
    for (int r = 0; r < bfs::RINGS_PER_BATCH; ++r) {
      bfs::bfs_claim_kernel<<<bfs::GRID, bfs::THREADS, 0, st>>>(w.st, p, fv, vf, K, out_faces, w.fstamp, w.claim);
      bfs::bfs_count_kernel<<<bfs::GRID, bfs::THREADS, 0, st>>>(w.st, p, fv, vf, K, out_faces, w.fstamp, w.claim, w.block_sums);
      bfs::bfs_scatter_kernel<<<bfs::GRID, bfs::THREADS, 0, st>>>(w.st, p, fv, vf, K, out_faces, w.fstamp, w.claim, w.block_sums, cut);
      p ^= 1;
    }
    GEOBI_LAUNCH_OK("bfs ring");
    int rc = bfs_read_state(w, p, &h, st);
    if (rc) return rc;
    if (h.done) break;
  }
  // cover the patch, pick the next seed
  GEOBI_CUDA_OK(cudaMemsetAsync(w.fresh, 0, sizeof(int), st));
  bfs::bfs_cover_kernel<<<bfs::GRID, bfs::THREADS, 0, st>>>(w.st, p, out_faces, w.d2_left, w.fresh);
  bfs::argmax_partial_kernel<<<bfs::GRID, bfs::THREADS, 0, st>>>(w.d2_left, n_faces, w.part_val, w.part_idx);
  bfs::argmax_final_kernel<<<1, bfs::THREADS, 0, st>>>(w.part_val, w.part_idx, bfs::GRID, w.fresh, w.st, p);
  GEOBI_LAUNCH_OK("bfs cover");
  int rc = bfs_read_state(w, p, &h, st);
  if (rc) return rc;
  *n_out_host = h.hi;
  *next_seed_host = h.next_seed;
  return GEOBI_OK;
}
