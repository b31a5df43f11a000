// Integer / graph kernels of libgeobi: scan, COO->CSR (coalesce), facet graph, exact parallel
// greedy matching (graclus), cluster relabel / grouping, edge coarsening.  All results are
// bit-exact w.r.t. the oracle (sorted-unique CSR == torch_sparse.coalesce order).
//
// One device primitive carries every "sort + dedup a short adjacency row" need: rows are
// filled as 64-bit keys (primary<<32 | secondary), one warp rank-sorts its row (O(L^2/32),
// L ~ 6..40 on meshes), marks run heads, and a second pass compacts after a scan.
#include "common.cuh"

namespace geobi {

constexpr uint64_t KEY_INVALID = ~0ull;
constexpr int MAX_ROW = 32768;  // rank sort is quadratic; longer rows are rejected (status)

// status words kept in every graph workspace
enum { ST_ERR = 0, ST_AUX = 1, ST_WORDS = 4 };

// ------------------------------------------------------------------------------ scan
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// exclusive scan of one int per thread over a 256-thread block; total -> *total (all threads)
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int wsum[SCAN_THREADS / 32];
  __shared__ int tot;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int s = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;
    if (lane == 31) tot = si;
  }
  __syncthreads();
  const int r = wsum[wid] + incl - v;
  *total = tot;
  __syncthreads();  // wsum/tot reusable by a following call
  return r;
}

// Single-pass exclusive scan with decoupled look-back (Merrill & Garland): one launch instead of reduce / scan-partials /
// final.  Tiles are handed out by a ticket counter, so a tile only ever waits on tiles that are already running.
// desc[t]: bits 63..62 = 0 not ready, 1 tile aggregate, 2 inclusive prefix; low 32 bits = the value.
__global__ void __launch_bounds__(SCAN_THREADS) scan_onepass_kernel(const int* __restrict__ in, int64_t n, int* __restrict__ out,
                                                                    unsigned long long* desc, unsigned* ticket) {
  __shared__ unsigned s_tile;
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  int tot;
  int ex = block_excl_scan(s, &tot);
  if (threadIdx.x < 32) {
    // warp-wide look-back: 32 predecessors per step; stop at the nearest one that already knows its inclusive prefix
    const int lane = threadIdx.x;
    int run = 0;
    if (tile > 0) {
      if (lane == 0) atomicExch(desc + tile, (1ull << 62) | (unsigned)tot);
      for (int hi = (int)tile - 1; hi >= 0; hi -= 32) {
        const int t = hi - lane;
        unsigned long long d = 2ull << 62;             // lanes before tile 0 act as a zero prefix
        if (t >= 0) {
          do {
            d = *reinterpret_cast<volatile unsigned long long*>(desc + t);
          } while ((d >> 62) == 0);
        }
        const unsigned full = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        const int stop = full ? __ffs(full) - 1 : 31;   // nearest predecessor with a full prefix
        int v2 = lane <= stop ? (int)(unsigned)d : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        run += v2;
        if (full) break;
      }
    }
    if (lane == 0) {
      atomicExch(desc + tile, (2ull << 62) | (unsigned)(run + tot));
      s_prefix = run;
    }
  }
  __syncthreads();
  ex += s_prefix;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
    if (i == n - 1) out[n] = ex;
  }
}

size_t scan_ws_bytes(int64_t n) { return align256(16 + (size_t)(cdiv(n > 0 ? n : 1, SCAN_TILE)) * sizeof(unsigned long long)) + 256; }

int scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, cudaStream_t st) {
  GEOBI_REQUIRE(n >= 0 && out != nullptr, "scan: bad arguments");
  if (n == 0) {
    GEOBI_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(int), st));
    return GEOBI_OK;
  }
  if (ws_bytes < scan_ws_bytes(n) || ws == nullptr) {
    set_error("scan: workspace too small (%zu < %zu)", ws_bytes, scan_ws_bytes(n));
    return GEOBI_ERR_WORKSPACE;
  }
  const int64_t nb = cdiv(n, SCAN_TILE);
  unsigned* ticket = static_cast<unsigned*>(ws);
  unsigned long long* desc = reinterpret_cast<unsigned long long*>(static_cast<char*>(ws) + 16);
  GEOBI_CUDA_OK(cudaMemsetAsync(ws, 0, 16 + (size_t)nb * sizeof(unsigned long long), st));
  scan_onepass_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, out, desc, ticket);
  GEOBI_LAUNCH_OK("scan");
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ row sort / compact
// One warp per row.  Row r occupies raw[start, end): start = raw_rowptr[r] (or r*stride).
// Valid keys are unique.  sorted[start + rank] = key;  vcount[r] = #valid;  ucount[r] = #run heads
// (runs = equal high words) when dedup, else #valid.
__global__ void __launch_bounds__(256) sort_rows_kernel(const int* __restrict__ raw_rowptr, int stride, const uint64_t* __restrict__ raw,
                                                        uint64_t* __restrict__ sorted, int* __restrict__ vcount, int* __restrict__ ucount,
                                                        int64_t nrows, int dedup, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nrows) return;
  const int64_t start = raw_rowptr ? (int64_t)raw_rowptr[r] : r * (int64_t)stride;
  const int64_t end = raw_rowptr ? (int64_t)raw_rowptr[r + 1] : start + stride;
  const int L = (int)(end - start);
  if (L > MAX_ROW) {
    if (lane == 0) {
      atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
      vcount[r] = 0;
      ucount[r] = 0;
    }
    return;
  }
  const uint64_t* row = raw + start;
  int nvalid = 0;
  for (int i0 = 0; i0 < L; i0 += 32) {
    const int i = i0 + lane;
    const uint64_t key = i < L ? row[i] : KEY_INVALID;
    const bool valid = key != KEY_INVALID;
    if (valid) {
      int rank = 0;
      for (int k = 0; k < L; ++k) rank += (row[k] < key) ? 1 : 0;
      sorted[start + rank] = key;
    }
    nvalid += __popc(__ballot_sync(0xffffffffu, valid));
  }
  __syncwarp();
  int heads = nvalid;
  if (dedup) {
    heads = 0;
    for (int t0 = 0; t0 < nvalid; t0 += 32) {
      const int t = t0 + lane;
      bool head = false;
      if (t < nvalid) head = (t == 0) || ((sorted[start + t] >> 32) != (sorted[start + t - 1] >> 32));
      heads += __popc(__ballot_sync(0xffffffffu, head));
    }
  }
  if (lane == 0) {
    vcount[r] = nvalid;
    ucount[r] = heads;
  }
}

// Second pass: write the compacted CSR row.  nbr = high word (nbr_in_hi) or low word of the key; the
// other word ("tag") indexes the weight source: widx = tag (+ raw row start if w_rel); tags >= w_mod
// (w_mod > 0) are flipped copies of edge tag - w_mod.
__global__ void __launch_bounds__(256) compact_rows_kernel(const int* __restrict__ raw_rowptr, int stride, const uint64_t* __restrict__ sorted,
                                                           const int* __restrict__ vcount, const int* __restrict__ out_rowptr, int64_t nrows,
                                                           int dedup, int nbr_in_hi, const float* __restrict__ wsrc, int w_rel, int64_t w_mod,
                                                           int w_mean, int32_t* __restrict__ out_nbr, float* __restrict__ out_w,
                                                           int64_t* __restrict__ eid_out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nrows) return;
  const int64_t start = raw_rowptr ? (int64_t)raw_rowptr[r] : r * (int64_t)stride;
  const int nvalid = vcount[r];
  const int64_t ostart = out_rowptr[r];
  int base = 0;
  for (int t0 = 0; t0 < nvalid; t0 += 32) {
    const int t = t0 + lane;
    uint64_t key = 0;
    bool head = false;
    if (t < nvalid) {
      key = sorted[start + t];
      head = !dedup || t == 0 || ((key >> 32) != (sorted[start + t - 1] >> 32));
    }
    const unsigned mask = __ballot_sync(0xffffffffu, head);
    if (head) {
      const int64_t o = ostart + base + __popc(mask & ((1u << lane) - 1u));
      const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
      out_nbr[o] = (int32_t)(nbr_in_hi ? hi : lo);
      const uint32_t tag = nbr_in_hi ? lo : hi;
      if (eid_out) eid_out[o] = (w_mod > 0 && (int64_t)tag >= w_mod) ? -1 - ((int64_t)tag - w_mod) : (int64_t)tag;
      if (out_w) {
        float s = 0.f;
        int cnt = 0;
        for (int q = t; q < nvalid; ++q) {
          const uint64_t kq = sorted[start + q];
          if (q > t && (!dedup || (kq >> 32) != (key >> 32))) break;
          int64_t tg = nbr_in_hi ? (uint32_t)kq : (uint32_t)(kq >> 32);
          if (w_mod > 0 && tg >= w_mod) tg -= w_mod;
          s += wsrc[w_rel ? start + tg : tg];
          ++cnt;
        }
        out_w[o] = w_mean ? s / (float)cnt : s;
      }
    }
    base += __popc(mask);
  }
}

static int rows_sort_compact(const int* raw_rowptr, int stride, const uint64_t* raw, uint64_t* sorted, int* vcount, int* ucount,
                             int64_t nrows, int dedup, int nbr_in_hi, const float* wsrc, int w_rel, int64_t w_mod, int w_mean,
                             int32_t* out_rowptr, int32_t* out_nbr, float* out_w, int64_t* eid_out, int* status, void* scan_ws,
                             size_t scan_bytes, cudaStream_t st) {
  if (nrows > 0) {
    const unsigned blocks = (unsigned)cdiv(nrows * 32, 256);
    sort_rows_kernel<<<blocks, 256, 0, st>>>(raw_rowptr, stride, raw, sorted, vcount, ucount, nrows, dedup, status);
    GEOBI_LAUNCH_OK("sort_rows");
  }
  int rc = scan_i32(ucount, out_rowptr, nrows, scan_ws, scan_bytes, st);
  if (rc) return rc;
  if (nrows > 0) {
    const unsigned blocks = (unsigned)cdiv(nrows * 32, 256);
    compact_rows_kernel<<<blocks, 256, 0, st>>>(raw_rowptr, stride, sorted, vcount, out_rowptr, nrows, dedup, nbr_in_hi, wsrc, w_rel,
                                                w_mod, w_mean, out_nbr, out_w, eid_out);
    GEOBI_LAUNCH_OK("compact_rows");
  }
  return GEOBI_OK;
}

// copies status + rowptr[n] to the host and synchronises; turns device-side errors into return codes
static int finish_sync(const int* status, const int32_t* rowptr, int64_t nrows, int64_t* nnz_host, const char* what, cudaStream_t st) {
  // pinned landing pad (one per host thread): both copies are queued asynchronously and one synchronise covers them; into
  // pageable memory each copy is staged and blocks on its own (two stream drains and ~12 us between them in the timeline)
  static thread_local int* pad = nullptr;
  if (pad == nullptr) GEOBI_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&pad), 64, cudaHostAllocDefault));
  int* h_status = pad;
  int& h_nnz = pad[ST_WORDS];
  h_nnz = 0;
  GEOBI_CUDA_OK(cudaMemcpyAsync(h_status, status, sizeof(int) * ST_WORDS, cudaMemcpyDeviceToHost, st));
  if (rowptr) GEOBI_CUDA_OK(cudaMemcpyAsync(&h_nnz, rowptr + nrows, sizeof(int), cudaMemcpyDeviceToHost, st));
  GEOBI_CUDA_OK(cudaStreamSynchronize(st));
  if (h_status[ST_ERR] != 0) {
    set_error("%s: device-side check failed (index out of range or adjacency row longer than %d)", what, MAX_ROW);
    return h_status[ST_ERR];
  }
  if (nnz_host) *nnz_host = h_nnz;
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ COO -> CSR
struct CooView {
  const int64_t* row;
  const int64_t* col;
  int64_t n_edges;
  int64_t n_nodes;
  int by_col, drop_self, symmetrize;
  // element id t in [0, n_edges * (symmetrize ? 2 : 1)); returns false if skipped
  __device__ __forceinline__ bool get(int64_t t, int64_t& seg, int64_t& other, int* status) const {
    const bool flip = t >= n_edges;
    const int64_t e = flip ? t - n_edges : t;
    int64_t a = row[e], b = col[e];
    if (a < 0 || a >= n_nodes || b < 0 || b >= n_nodes) {
      atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
      return false;
    }
    if (drop_self && a == b) return false;
    if (flip != (by_col != 0)) { int64_t x = a; a = b; b = x; }
    seg = a;
    other = b;
    return true;
  }
};

__global__ void coo_count_kernel(CooView v, int64_t total, int* __restrict__ count, int* status) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t s, o;
    if (v.get(t, s, o, status)) atomicAdd(&count[s], 1);
  }
}

__global__ void coo_fill_kernel(CooView v, int64_t total, const int* __restrict__ raw_rowptr, int* __restrict__ cursor, int sort_nbr,
                                uint64_t* __restrict__ raw, int* status) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t s, o;
    if (!v.get(t, s, o, status)) continue;
    const int pos = raw_rowptr[s] + atomicAdd(&cursor[s], 1);
    raw[pos] = sort_nbr ? (((uint64_t)o << 32) | (uint64_t)(uint32_t)t) : (((uint64_t)(uint32_t)t << 32) | (uint64_t)(uint32_t)o);
  }
}

static unsigned grid_for(int64_t n, int threads) {
  int64_t b = cdiv(n, threads);
  const int64_t cap = 148 * 64;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

struct CooWs {
  int *count, *raw_rowptr, *vcount, *ucount, *status;
  uint64_t *raw, *sorted;
  void* scan;
  size_t scan_bytes;
};
template <class C>
static void carve_coo(C& c, int64_t total, int64_t n, CooWs* out) {
  int* count = c.template take<int>(n + 1);
  int* raw_rowptr = c.template take<int>(n + 1);
  int* vcount = c.template take<int>(n + 1);
  int* ucount = c.template take<int>(n + 1);
  int* status = c.template take<int>(ST_WORDS);
  uint64_t* raw = c.template take<uint64_t>(total + 1);
  uint64_t* sorted = c.template take<uint64_t>(total + 1);
  const size_t sb = scan_ws_bytes(n + 1);
  char* scan = c.template take<char>(sb);
  if (out) *out = CooWs{count, raw_rowptr, vcount, ucount, status, raw, sorted, scan, sb};
}
struct NullCarver {
  Sizer s;
  template <typename T>
  T* take(size_t n) { s.take<T>(n); return nullptr; }
};

}  // namespace geobi

using namespace geobi;

extern "C" size_t geobi_scan_ws_bytes(int64_t n) { return scan_ws_bytes(n); }

extern "C" int geobi_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, void* stream) {
  return scan_i32(in, out, n, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" size_t geobi_csr_from_coo_ws_bytes(int64_t n_edges, int64_t n_nodes, int flags) {
  NullCarver c;
  carve_coo(c, n_edges * ((flags & GEOBI_COO_SYMMETRIZE) ? 2 : 1), n_nodes, nullptr);
  return c.s.total();
}

extern "C" int geobi_csr_from_coo(const int64_t* row, const int64_t* col, const float* w, int64_t n_edges, int64_t n_nodes, int flags,
                                  int32_t* rowptr, int32_t* nbr, float* w_out, int64_t* eid_out, int64_t* nnz_host, void* ws,
                                  size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sym = (flags & GEOBI_COO_SYMMETRIZE) ? 1 : 0;
  const int64_t total = n_edges * (sym ? 2 : 1);
  GEOBI_REQUIRE(n_edges >= 0 && n_nodes >= 0 && rowptr && (nbr || total == 0), "csr_from_coo: bad arguments");
  GEOBI_REQUIRE(total < (int64_t)1 << 31 && n_nodes < (int64_t)1 << 31, "csr_from_coo: sizes exceed int32 indexing");
  GEOBI_REQUIRE(!(flags & GEOBI_COO_DEDUP) || (flags & GEOBI_COO_SORT_NBR), "csr_from_coo: DEDUP needs SORT_NBR");
  GEOBI_REQUIRE((w == nullptr) == (w_out == nullptr), "csr_from_coo: w and w_out must both be given or both be NULL");
  if (ws_bytes < geobi_csr_from_coo_ws_bytes(n_edges, n_nodes, flags) || !ws) {
    set_error("csr_from_coo: workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  Carver c(ws, ws_bytes);
  CooWs W;
  carve_coo(c, total, n_nodes, &W);
  GEOBI_CUDA_OK(cudaMemsetAsync(W.count, 0, sizeof(int) * (n_nodes + 1), st));
  GEOBI_CUDA_OK(cudaMemsetAsync(W.status, 0, sizeof(int) * ST_WORDS, st));
  CooView v{row, col, n_edges, n_nodes, (flags & GEOBI_COO_BY_COL) ? 1 : 0, (flags & GEOBI_COO_DROP_SELF) ? 1 : 0, sym};
  if (total > 0) {
    coo_count_kernel<<<grid_for(total, 256), 256, 0, st>>>(v, total, W.count, W.status);
    GEOBI_LAUNCH_OK("coo_count");
  }
  int rc = scan_i32(W.count, W.raw_rowptr, n_nodes, W.scan, W.scan_bytes, st);
  if (rc) return rc;
  GEOBI_CUDA_OK(cudaMemsetAsync(W.count, 0, sizeof(int) * (n_nodes + 1), st));
  const int sort_nbr = (flags & GEOBI_COO_SORT_NBR) ? 1 : 0;
  if (total > 0) {
    coo_fill_kernel<<<grid_for(total, 256), 256, 0, st>>>(v, total, W.raw_rowptr, W.count, sort_nbr, W.raw, W.status);
    GEOBI_LAUNCH_OK("coo_fill");
  }
  rc = rows_sort_compact(W.raw_rowptr, 0, W.raw, W.sorted, W.vcount, W.ucount, n_nodes, (flags & GEOBI_COO_DEDUP) ? 1 : 0, sort_nbr, w,
                         /*w_rel=*/0, /*w_mod=*/sym ? n_edges : 0, (flags & GEOBI_COO_W_MEAN) ? 1 : 0, rowptr, nbr, w_out, eid_out,
                         W.status, W.scan, W.scan_bytes, st);
  if (rc) return rc;
  if (nnz_host) return finish_sync(W.status, rowptr, n_nodes, nnz_host, "csr_from_coo", st);
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ CSR -> COO
namespace geobi {
__global__ void csr_to_coo_kernel(const int* __restrict__ rowptr, const int* __restrict__ nbr, int64_t n, int64_t nnz,
                                  int64_t* __restrict__ ei) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const int s = rowptr[r], e = rowptr[r + 1];
  for (int k = s + lane; k < e; k += 32) {
    ei[k] = r;
    ei[nnz + k] = nbr[k];
  }
}
}  // namespace geobi

extern "C" int geobi_csr_to_coo(const int32_t* rowptr, const int32_t* nbr, int64_t n_nodes, int64_t nnz, int64_t* edge_index, void* stream) {
  GEOBI_REQUIRE(rowptr && edge_index && n_nodes >= 0 && nnz >= 0, "csr_to_coo: bad arguments");
  if (n_nodes == 0 || nnz == 0) return GEOBI_OK;
  csr_to_coo_kernel<<<(unsigned)cdiv(n_nodes * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rowptr, nbr, n_nodes, nnz, edge_index);
  GEOBI_LAUNCH_OK("csr_to_coo");
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ facet graph
namespace geobi {
__global__ void facet_fill_kernel(const int64_t* __restrict__ fv, const int64_t* __restrict__ vf, int64_t F, int64_t V, int K,
                                  uint64_t* __restrict__ raw, int* status) {
  const int64_t total = F * 3 * K;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = t / (3 * K);
    const int slot = (int)(t - f * 3 * K);
    const int64_t v = fv[f * 3 + slot / K];
    uint64_t key = KEY_INVALID;
    if (v < 0 || v >= V) {
      atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
    } else {
      const int64_t g = vf[v * K + slot % K];
      if (g >= F) atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
      else if (g >= 0) key = ((uint64_t)g << 32) | (uint32_t)slot;
    }
    raw[t] = key;
  }
}
struct FacetWs {
  int *vcount, *ucount, *status;
  uint64_t *raw, *sorted;
  void* scan;
  size_t scan_bytes;
};
template <class C>
static void carve_facet(C& c, int64_t F, int64_t K, FacetWs* out) {
  int* vcount = c.template take<int>(F + 1);
  int* ucount = c.template take<int>(F + 1);
  int* status = c.template take<int>(ST_WORDS);
  uint64_t* raw = c.template take<uint64_t>(F * 3 * K + 1);
  uint64_t* sorted = c.template take<uint64_t>(F * 3 * K + 1);
  const size_t sb = scan_ws_bytes(F + 1);
  char* scan = c.template take<char>(sb);
  if (out) *out = FacetWs{vcount, ucount, status, raw, sorted, scan, sb};
}
}  // namespace geobi

extern "C" size_t geobi_build_facet_graph_ws_bytes(int64_t n_faces, int64_t k) {
  NullCarver c;
  carve_facet(c, n_faces, k, nullptr);
  return c.s.total();
}

extern "C" int geobi_build_facet_graph(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t n_verts, int64_t k, int32_t* rowptr,
                                       int32_t* nbr, int64_t* nnz_host, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(fv && vf && rowptr && nbr && n_faces > 0 && n_verts > 0 && k > 0, "build_facet_graph: bad arguments");
  GEOBI_REQUIRE(n_faces * 3 * k < (int64_t)1 << 31, "build_facet_graph: 3*K*F exceeds int32 indexing");
  if (ws_bytes < geobi_build_facet_graph_ws_bytes(n_faces, k) || !ws) {
    set_error("build_facet_graph: workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  Carver c(ws, ws_bytes);
  FacetWs W;
  carve_facet(c, n_faces, k, &W);
  GEOBI_CUDA_OK(cudaMemsetAsync(W.status, 0, sizeof(int) * ST_WORDS, st));
  facet_fill_kernel<<<grid_for(n_faces * 3 * k, 256), 256, 0, st>>>(fv, vf, n_faces, n_verts, (int)k, W.raw, W.status);
  GEOBI_LAUNCH_OK("facet_fill");
  int rc = rows_sort_compact(nullptr, (int)(3 * k), W.raw, W.sorted, W.vcount, W.ucount, n_faces, /*dedup=*/1, /*nbr_in_hi=*/1, nullptr, 0, 0,
                             0, rowptr, nbr, nullptr, nullptr, W.status, W.scan, W.scan_bytes, st);
  if (rc) return rc;
  if (nnz_host) return finish_sync(W.status, rowptr, n_faces, nnz_host, "build_facet_graph", st);
  return GEOBI_OK;
}

// Facet 1-ring from SORTED incidence rows: a face's row is the union of its three corners' vf rows, each ascending with its -1
// pads at the end (topology.DeviceTriMesh builds them that way): a three-way merge with duplicate removal, one thread per face,
// instead of filling 3K keys per face and rank-sorting every row with a warp.  Pass 0 counts (and verifies the order: the
// merge would silently drop or repeat entries on an unsorted row), pass 1 writes at the scanned offsets.
namespace geobi {
template <bool WRITE>
__global__ void __launch_bounds__(256) facet_merge_kernel(const int64_t* __restrict__ fv, const int64_t* __restrict__ vf, int64_t F, int64_t V, int K,
                                                          int drop_self, const int32_t* __restrict__ rowptr, int32_t* __restrict__ out, int* status) {
  constexpr int64_t INF = (int64_t)1 << 62;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < F; f += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v0 = fv[3 * f], v1 = fv[3 * f + 1], v2 = fv[3 * f + 2];
    if (v0 < 0 || v0 >= V || v1 < 0 || v1 >= V || v2 < 0 || v2 >= V) {
      atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
      if (!WRITE) out[f] = 0;
      continue;
    }
    const int64_t* r0 = vf + v0 * K;
    const int64_t* r1 = vf + v1 * K;
    const int64_t* r2 = vf + v2 * K;
    int i0 = 0, i1 = 0, i2 = 0;
    auto head = [&](const int64_t* r, int i) -> int64_t {
      if (i >= K) return INF;
      const int64_t g = r[i];
      return g < 0 ? INF : g;
    };
    int64_t a = head(r0, 0), b = head(r1, 0), c = head(r2, 0);
    int32_t* dst = WRITE ? out + rowptr[f] : nullptr;
    int n = 0;
    bool bad = false;
    for (;;) {
      const int64_t m = a < b ? (a < c ? a : c) : (b < c ? b : c);
      if (m == INF) break;
      if (m >= F) { bad = true; break; }
      if (!(drop_self && m == f)) {
        if (WRITE) dst[n] = (int32_t)m;
        ++n;
      }
      if (a == m) { a = head(r0, ++i0); bad |= a <= m; }
      if (b == m) { b = head(r1, ++i1); bad |= b <= m; }
      if (c == m) { c = head(r2, ++i2); bad |= c <= m; }
      if (bad) break;
    }
    if (bad) atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
    if (!WRITE) out[f] = n;
  }
}
}  // namespace geobi

extern "C" size_t geobi_build_facet_graph_sorted_ws_bytes(int64_t n_faces) {
  return align256((size_t)(n_faces + 1) * sizeof(int)) + align256(sizeof(int) * ST_WORDS) + scan_ws_bytes(n_faces + 1) + 1024;
}

extern "C" int geobi_build_facet_graph_sorted(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t n_verts, int64_t k, int drop_self,
                                              int32_t* rowptr, int32_t* nbr, int64_t* nnz_host, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(fv && vf && rowptr && nbr && n_faces > 0 && n_verts > 0 && k > 0, "build_facet_graph_sorted: bad arguments");
  GEOBI_REQUIRE(n_faces * 3 * k < (int64_t)1 << 31, "build_facet_graph_sorted: 3*K*F exceeds int32 indexing");
  if (ws_bytes < geobi_build_facet_graph_sorted_ws_bytes(n_faces) || !ws) {
    set_error("build_facet_graph_sorted: workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  Carver c(ws, ws_bytes);
  int* count = c.take<int>(n_faces + 1);
  int* status = c.take<int>(ST_WORDS);
  const size_t sb = scan_ws_bytes(n_faces + 1);
  char* scan = c.take<char>(sb);
  GEOBI_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int) * ST_WORDS, st));
  facet_merge_kernel<false><<<grid_for(n_faces, 256), 256, 0, st>>>(fv, vf, n_faces, n_verts, (int)k, drop_self, nullptr, count, status);
  GEOBI_LAUNCH_OK("facet_merge (count)");
  int rc = scan_i32(count, rowptr, n_faces, scan, sb, st);
  if (rc) return rc;
  facet_merge_kernel<true><<<grid_for(n_faces, 256), 256, 0, st>>>(fv, vf, n_faces, n_verts, (int)k, drop_self, rowptr, nbr, status);
  GEOBI_LAUNCH_OK("facet_merge (fill)");
  if (nnz_host) return finish_sync(status, rowptr, n_faces, nnz_host, "build_facet_graph_sorted", st);
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ graclus (exact parallel greedy)
namespace geobi {
constexpr int M_NONE = -1, M_SINGLE = -2;

// visiting order: u precedes v iff (rank[u], u) < (rank[v], v); ranks may therefore be any int32 keys (an inverse permutation
// reproduces torch_cluster's order exactly, i.i.d. random keys give a uniformly random order without a sort)
__device__ __forceinline__ bool precedes(int rv, int v, int ru, int u) { return rv < ru || (rv == ru && v < u); }

// Asynchronous exact greedy matching, one launch, no grid barriers.
//
// Every node is owned by one resident thread (grid-stride ownership over a co-resident grid); the thread sweeps over its
// still-undecided nodes until all are decided.  A node u may act when (1) every neighbour that precedes it is decided and
// (2) every undecided neighbour of its chosen partner v comes after u; it then claims v with atomicCAS(label[v], -1, min)
// and publishes its own label afterwards.  Labels only ever go from -1 to their final value, so a stale read can only make
// a node wait (it sees a decided node as undecided), never act wrongly; the CAS is the linearisation point of a claim.
// The result is therefore the serial greedy matching for the visiting order, independent of thread timing.
// blk[u] caches the node u is waiting for, so a blocked node costs 3 loads per sweep.
constexpr int GRACLUS_MAX_SWEEPS = 1 << 20;

// One evaluation of node u.  Neighbour rows are read in chunks of GCH with all loads of a chunk (ids -> states, weights)
// issued back to back, so a decision costs ~2 dependent memory latencies per chunk instead of 3 per neighbour.
// GCH = 4 and two register-resident nodes per thread keep the kernel at 32 registers = 2048 resident threads per SM: the
// matcher speeds up with the number of nodes that have their own thread (A/B on the bench graphs: 64 registers / chunks
// of 8 -> 0.58 ms, 32 registers / chunks of 4 -> 0.38 ms on the 512 000-node facet graph; chunks of 16 -> 1.9 ms);
// the critical path of the whole matching is (dependency depth ~18) x (decision latency).
// st[v] = {label, rank} interleaved: one 8-byte gather per neighbour returns both (the kernel is bound by L2 sector
// requests: ~50 scattered reads per evaluation with separate arrays, ~25 with the pair).
constexpr int GCH = 4;
__device__ __forceinline__ int2 ld_state(const int2* st, int v) {
  const long long raw = __ldcg(reinterpret_cast<const long long*>(st) + v);
  return make_int2((int)(raw & 0xffffffffll), (int)(raw >> 32));
}
// `blk` (in/out) is the node u is waiting for, -1 if none; the owner thread keeps it in a register, so a blocked node
// costs ONE load per sweep (its blocker's label) - with the blocker in memory the poll was three dependent loads
// (own state slot -> own label -> blocker's label), ~45 % of the kernel's stall samples.
__device__ __forceinline__ int graclus_try(int u, const int* __restrict__ rowptr, const int* __restrict__ nbr, const float* __restrict__ w,
                                           int2* st, int& blk, int* __restrict__ label_out) {
  int* lab = reinterpret_cast<int*>(st);           // label of node v at lab[2 * v]
  if (blk >= 0 && __ldcg(lab + 2 * blk) < 0) return 0;   // the node we are waiting for is still undecided (fast path)
  const int2 su = ld_state(st, u);
  if (su.x >= 0) { label_out[u] = su.x; return 1; }   // claimed by a partner
  const int ru = su.y;
  const int end = rowptr[u + 1];
  int best = -1, blocker = -1;
  float wmax = 0.f;
  for (int e0 = rowptr[u]; e0 < end && blocker < 0; e0 += GCH) {
    int v[GCH];
    int2 sv[GCH];
    float wt[GCH];
#pragma unroll
    for (int k = 0; k < GCH; ++k) v[k] = e0 + k < end ? nbr[e0 + k] : -1;
#pragma unroll
    for (int k = 0; k < GCH; ++k) {
      sv[k] = v[k] >= 0 ? ld_state(st, v[k]) : make_int2(0, 0);
      wt[k] = (w && v[k] >= 0) ? w[e0 + k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < GCH; ++k) {
      if (v[k] < 0 || sv[k].x >= 0 || blocker >= 0) continue;
      if (precedes(sv[k].y, v[k], ru, u)) { blocker = v[k]; continue; }   // an earlier neighbour is still undecided
      if (!w) { if (best < 0) best = v[k]; }
      else if (wt[k] >= wmax) { best = v[k]; wmax = wt[k]; }
    }
  }
  if (blocker >= 0) { blk = blocker; return 0; }
  // Only earlier neighbours can claim u, and a claimer CASes label[u] *before* it publishes its own label.  All of them
  // are now observed decided, so after this fence a claim on u (if any) is visible; if none, nobody can claim u any more.
  __threadfence();
  const int lu = __ldcg(lab + 2 * u);
  if (lu >= 0) { label_out[u] = lu; return 1; }
  if (best < 0) {
    __stcg(lab + 2 * u, u);              // no free neighbour: singleton
    label_out[u] = u;
    return 1;
  }
  const int bend = rowptr[best + 1];
  for (int e0 = rowptr[best]; e0 < bend; e0 += GCH) {
    int z[GCH];
    int2 sz[GCH];
#pragma unroll
    for (int k = 0; k < GCH; ++k) z[k] = e0 + k < bend ? nbr[e0 + k] : -1;
#pragma unroll
    for (int k = 0; k < GCH; ++k) sz[k] = z[k] >= 0 ? ld_state(st, z[k]) : make_int2(0, 0);
#pragma unroll
    for (int k = 0; k < GCH; ++k)
      if (z[k] >= 0 && sz[k].x < 0 && precedes(sz[k].y, z[k], ru, u)) { blk = z[k]; return 0; }   // it may still claim `best`
  }
  const int l = best < u ? best : u;
  if (atomicCAS(lab + 2 * best, -1, l) != -1) { blk = -1; return 0; }   // lost a race against a stale view: retry
  __threadfence();
  __stcg(lab + 2 * u, l);
  label_out[u] = l;
  return 1;
}

__global__ void graclus_init_kernel(const int* __restrict__ rank, int n, int2* __restrict__ st, int* __restrict__ pos) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  st[u] = make_int2(-1, rank[u]);        // -1 = undecided
  pos[u] = -1;                           // node this one is waiting for (-1 none, -2 done)
}

__global__ void __launch_bounds__(256, 8) graclus_async_kernel(const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                            const float* __restrict__ w, int2* st, int* __restrict__ label, int n, int* pos,
                                                            int* undecided) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  // per-node wait state: -2 done, -1 evaluate, >= 0 blocker.  The first GR_OWN nodes of a thread live in registers
  // (a co-resident grid owns ~2 nodes per thread at the sizes of this path), any further ones in pos[].
  constexpr int GR_OWN = 2;
  int blk[GR_OWN];
  int left = 0;
#pragma unroll
  for (int k = 0; k < GR_OWN; ++k) {
    const long long u = (long long)tid + (long long)k * stride;
    blk[k] = u < n ? -1 : -2;
    left += u < n ? 1 : 0;
  }
  const long long u_mem = (long long)tid + (long long)GR_OWN * stride;
  for (long long u = u_mem; u < n; u += stride) ++left;
  unsigned backoff = 32;
  for (int sweep = 0; left > 0 && sweep < GRACLUS_MAX_SWEEPS; ++sweep) {
    int still = 0;
#pragma unroll
    for (int k = 0; k < GR_OWN; ++k) {
      if (blk[k] == -2) continue;
      const int u = tid + k * stride;
      if (graclus_try(u, rowptr, nbr, w, st, blk[k], label)) blk[k] = -2;
      else ++still;
    }
    for (long long u = u_mem; u < n; u += stride) {
      int b = pos[u];
      if (b == -2) continue;
      if (graclus_try((int)u, rowptr, nbr, w, st, b, label)) b = -2;
      else ++still;
      pos[u] = b;
    }
    if (still == left) {                 // no progress: let the owners of the blocking nodes run
      __nanosleep(backoff);
      if (backoff < 1024) backoff <<= 1;
    } else {
      backoff = 32;
    }
    left = still;
  }
  if (left > 0) {
    atomicAdd(undecided, left);
#pragma unroll
    for (int k = 0; k < GR_OWN; ++k)
      if (blk[k] != -2) label[tid + k * stride] = -1;   // sweep limit: leave a detectable marker (relabel_clusters rejects it)
    for (long long u = u_mem; u < n; u += stride)
      if (pos[u] != -2) label[u] = -1;
  }
}
}  // namespace geobi

// ws: [undecided counter] | pos[N] | st[N] = {label, rank}
extern "C" size_t geobi_graclus_ws_bytes(int64_t n_nodes) {
  return 256 + align256((size_t)(n_nodes + 1) * sizeof(int)) + align256((size_t)(n_nodes + 1) * sizeof(int2)) + 256;
}

extern "C" int geobi_graclus(const int32_t* rowptr, const int32_t* nbr, const float* w, const int32_t* rank, int64_t n_nodes, int32_t* label,
                             int* undecided_host, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(rowptr && rank && label && n_nodes >= 0 && n_nodes < ((int64_t)1 << 31), "graclus: bad arguments");
  if (undecided_host) *undecided_host = 0;
  if (n_nodes == 0) return GEOBI_OK;
  if (ws_bytes < geobi_graclus_ws_bytes(n_nodes) || !ws) {
    set_error("graclus: workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  int* undecided = static_cast<int*>(ws);
  int* pos = reinterpret_cast<int*>(static_cast<char*>(ws) + 256);
  int2* state = reinterpret_cast<int2*>(static_cast<char*>(ws) + 256 + align256((size_t)(n_nodes + 1) * sizeof(int)));
  GEOBI_CUDA_OK(cudaMemsetAsync(undecided, 0, sizeof(int), st));
  graclus_init_kernel<<<(unsigned)cdiv(n_nodes, 256), 256, 0, st>>>(rank, (int)n_nodes, state, pos);
  static int coresident = 0;   // blocks that are guaranteed to be resident together (threads wait on each other's nodes)
  if (coresident == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    GEOBI_CUDA_OK(cudaGetDevice(&dev));
    GEOBI_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GEOBI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, graclus_async_kernel, 256, 0));
    coresident = sms * per_sm;
    GEOBI_REQUIRE(coresident > 0, "graclus: cooperative launch not possible on this device");
  }
  int64_t nb = cdiv(n_nodes, 256);
  if (nb > coresident) nb = coresident;
  if (const char* cap = getenv("GEOBI_GRACLUS_MAX_BLOCKS")) {   // test hook: forces many nodes per thread (register + memory wait state)
    const int64_t c = atoll(cap);
    if (c >= 1 && c < nb) nb = c;
  }
  int n_all = (int)n_nodes;
  void* args[] = {(void*)&rowptr, (void*)&nbr, (void*)&w, (void*)&state, (void*)&label, (void*)&n_all, (void*)&pos, (void*)&undecided};
  // cooperative launch = the runtime refuses the launch unless all blocks are co-resident (no grid barrier is used)
  GEOBI_CUDA_OK(cudaLaunchCooperativeKernel((const void*)graclus_async_kernel, dim3((unsigned)nb), dim3(256), args, 0, st));
  if (undecided_host) {   // optional convergence check (SYNCS); callers on the hot path fold it into geobi_relabel_clusters
    int h_und = 0;
    GEOBI_CUDA_OK(cudaMemcpyAsync(&h_und, undecided, sizeof(int), cudaMemcpyDeviceToHost, st));
    GEOBI_CUDA_OK(cudaStreamSynchronize(st));
    *undecided_host = h_und;
    if (h_und > 0) {
      set_error("graclus: %d of %lld nodes undecided (sweep limit)", h_und, (long long)n_nodes);
      return GEOBI_ERR_NOCONVERGE;
    }
  }
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ relabel / group_by
namespace geobi {
__global__ void mark_labels_kernel(const int* __restrict__ label, int64_t n, int* __restrict__ flag, int* status) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const int l = label[u];
  if (l < 0 || l >= n) atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
  else flag[l] = 1;
}
__global__ void apply_labels_kernel(const int* __restrict__ label, const int* __restrict__ newid, int64_t n, int* __restrict__ cluster) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const int l = label[u];
  cluster[u] = (l >= 0 && l < n) ? newid[l] : 0;
}
__global__ void group_count_kernel(const int* __restrict__ cluster, int64_t n, int64_t nc, int* __restrict__ count, int* status) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const int c = cluster[u];
  if (c < 0 || c >= nc) atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
  else atomicAdd(&count[c], 1);
}
__global__ void group_fill_kernel(const int* __restrict__ cluster, int64_t n, int64_t nc, const int* __restrict__ raw_rowptr,
                                  int* __restrict__ cursor, uint64_t* __restrict__ raw) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const int c = cluster[u];
  if (c < 0 || c >= nc) return;
  const int pos = raw_rowptr[c] + atomicAdd(&cursor[c], 1);
  raw[pos] = (uint64_t)(uint32_t)u << 32;
}
}  // namespace geobi

namespace geobi {
// Matchings give clusters of one or two nodes with label = min(member): the member CSR needs no sort.
__global__ void pair_count_kernel(const int* __restrict__ label, const int* __restrict__ cluster, int64_t n, int* __restrict__ count) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  if (label[u] != (int)u) count[cluster[u]] = 2;      // the larger member marks its cluster as a pair
}
__global__ void pair_fill_kernel(const int* __restrict__ label, const int* __restrict__ cluster, int64_t n, const int* __restrict__ mrowptr,
                                 int* __restrict__ members) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  members[mrowptr[cluster[u]] + (label[u] != (int)u ? 1 : 0)] = (int)u;
}
__global__ void fill_ones_kernel(int* __restrict__ p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 1;
}
}  // namespace geobi

extern "C" size_t geobi_group_pairs_ws_bytes(int64_t n_clusters) { return align256((size_t)(n_clusters + 1) * sizeof(int)) + scan_ws_bytes(n_clusters + 1) + 256; }

extern "C" int geobi_group_pairs(const int32_t* label, const int32_t* cluster, int64_t n, int64_t nc, int32_t* mrowptr, int32_t* members, void* ws,
                                 size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(label && cluster && mrowptr && members && n >= 0 && nc >= 0, "group_pairs: bad arguments");
  if (ws_bytes < geobi_group_pairs_ws_bytes(nc) || !ws) { set_error("group_pairs: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver c(ws, ws_bytes);
  int* count = c.take<int>(nc + 1);
  const size_t sb = scan_ws_bytes(nc + 1);
  char* scan = c.take<char>(sb);
  if (nc > 0) fill_ones_kernel<<<(unsigned)cdiv(nc, 256), 256, 0, st>>>(count, nc);
  if (n > 0) pair_count_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(label, cluster, n, count);
  int rc = scan_i32(count, mrowptr, nc, scan, sb, st);
  if (rc) return rc;
  if (n > 0) pair_fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(label, cluster, n, mrowptr, members);
  GEOBI_LAUNCH_OK("group_pairs");
  return GEOBI_OK;
}

extern "C" size_t geobi_relabel_ws_bytes(int64_t n) {
  Sizer s;
  s.take<int>(n + 1);
  s.take<int>(n + 2);
  s.take<int>(ST_WORDS);
  s.take<char>(scan_ws_bytes(n + 1));
  return s.total();
}

extern "C" int geobi_relabel_clusters(const int32_t* label, int64_t n, int32_t* cluster, int64_t* n_clusters_host, void* ws, size_t ws_bytes,
                                      void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(label && cluster && n_clusters_host && n >= 0, "relabel_clusters: bad arguments");
  if (n == 0) { *n_clusters_host = 0; return GEOBI_OK; }
  if (ws_bytes < geobi_relabel_ws_bytes(n) || !ws) { set_error("relabel_clusters: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver c(ws, ws_bytes);
  int* flag = c.take<int>(n + 1);
  int* newid = c.take<int>(n + 2);
  int* status = c.take<int>(ST_WORDS);
  const size_t sb = scan_ws_bytes(n + 1);
  char* scan = c.take<char>(sb);
  GEOBI_CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int) * (n + 1), st));
  GEOBI_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int) * ST_WORDS, st));
  const unsigned blocks = (unsigned)cdiv(n, 256);
  mark_labels_kernel<<<blocks, 256, 0, st>>>(label, n, flag, status);
  int rc = scan_i32(flag, newid, n, scan, sb, st);
  if (rc) return rc;
  apply_labels_kernel<<<blocks, 256, 0, st>>>(label, newid, n, cluster);
  GEOBI_LAUNCH_OK("relabel");
  return finish_sync(status, newid, n, n_clusters_host, "relabel_clusters", st);
}

extern "C" size_t geobi_group_by_ws_bytes(int64_t n, int64_t nc) {
  Sizer s;
  s.take<int>(nc + 1); s.take<int>(nc + 1); s.take<int>(nc + 1); s.take<int>(nc + 1);
  s.take<int>(ST_WORDS);
  s.take<uint64_t>(n + 1); s.take<uint64_t>(n + 1);
  s.take<char>(scan_ws_bytes(nc + 1));
  return s.total();
}

// ------------------------------------------------------------------------------ vertex 1-ring from the incidence CSR
// Vertex adjacency of a triangle mesh (to_undirected of the unique edges, dataset.py:211, without the self loops) straight from the
// faces around each vertex: vertex v's neighbours are the two other corners of its incident faces - at most 2 * valence candidates,
// sorted and deduplicated in registers.  One thread per vertex, a count pass and a fill pass around a scan; rows come out ascending.
// The general builder (symmetrise 3F half edges, count, fill, per-row rank sort, compact) does the same in six kernels over 6F entries.
constexpr int RING_MAX_VALENCE = 24;
template <bool WRITE>
__global__ void __launch_bounds__(128) vertex_ring_kernel(const int64_t* __restrict__ fv, const int32_t* __restrict__ vf_rowptr,
                                                          const int32_t* __restrict__ corners, int64_t V, int64_t F,
                                                          const int32_t* __restrict__ rowptr, int32_t* __restrict__ out, int* status) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
    const int b = vf_rowptr[v], e = vf_rowptr[v + 1];
    if (e - b > RING_MAX_VALENCE) {
      atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
      if (!WRITE) out[v] = 0;
      continue;
    }
    int32_t cand[2 * RING_MAX_VALENCE];
    int n = 0;
    bool bad = false;
    for (int q = b; q < e; ++q) {
      const int m = corners[q];
      const int64_t f = m / 3;
      const int c = m - 3 * (int)f;
      if (f < 0 || f >= F) { bad = true; break; }
      const int64_t a0 = fv[3 * f + (c + 1) % 3], a1 = fv[3 * f + (c + 2) % 3];
      if (a0 < 0 || a0 >= V || a1 < 0 || a1 >= V) { bad = true; break; }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int32_t x = (int32_t)(t ? a1 : a0);
        if (x == (int32_t)v) continue;             // degenerate face: no self loop
        // sorted insert without duplicates (rows are 6-12 long)
        int pos = n;
        while (pos > 0 && cand[pos - 1] > x) --pos;
        if (pos > 0 && cand[pos - 1] == x) continue;
        for (int r = n; r > pos; --r) cand[r] = cand[r - 1];
        cand[pos] = x;
        ++n;
      }
    }
    if (bad) {
      atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
      n = 0;
    }
    if (WRITE) {
      int32_t* dst = out + rowptr[v];
      for (int r = 0; r < n; ++r) dst[r] = cand[r];
    } else {
      out[v] = n;
    }
  }
}
extern "C" size_t geobi_mesh_vertex_csr_ws_bytes(int64_t n_verts) {
  return align256(sizeof(int) * (size_t)(n_verts + 1)) + align256(sizeof(int) * ST_WORDS) + align256(scan_ws_bytes(n_verts + 1)) + 512;
}
extern "C" int geobi_mesh_vertex_csr(const int64_t* fv, const int32_t* vf_rowptr, const int32_t* corners, int64_t n_verts, int64_t n_faces,
                                     int32_t* rowptr, int32_t* nbr, int64_t* nnz_host, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(fv && vf_rowptr && corners && rowptr && nbr && n_verts > 0 && n_faces > 0, "mesh_vertex_csr: bad arguments");
  GEOBI_REQUIRE(n_faces * 6 < (int64_t)1 << 31, "mesh_vertex_csr: 6 F exceeds int32 indexing");
  if (ws_bytes < geobi_mesh_vertex_csr_ws_bytes(n_verts) || !ws) {
    set_error("mesh_vertex_csr: workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  Carver c(ws, ws_bytes);
  int* count = c.take<int>(n_verts + 1);
  int* status = c.take<int>(ST_WORDS);
  const size_t sb = scan_ws_bytes(n_verts + 1);
  char* scan = c.take<char>(sb);
  GEOBI_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int) * ST_WORDS, st));
  vertex_ring_kernel<false><<<grid_for(n_verts, 128), 128, 0, st>>>(fv, vf_rowptr, corners, n_verts, n_faces, nullptr, count, status);
  GEOBI_LAUNCH_OK("vertex_ring (count)");
  int rc = scan_i32(count, rowptr, n_verts, scan, sb, st);
  if (rc) return rc;
  vertex_ring_kernel<true><<<grid_for(n_verts, 128), 128, 0, st>>>(fv, vf_rowptr, corners, n_verts, n_faces, rowptr, nbr, status);
  GEOBI_LAUNCH_OK("vertex_ring (fill)");
  if (nnz_host) return finish_sync(status, rowptr, n_verts, nnz_host, "mesh_vertex_csr", st);
  return GEOBI_OK;
}

// member CSR -> the padded [n, k] int64 table of the reference's mesh arrays (vf_indices / vv_indices: OpenMesh circulators padded with -1)
__global__ void __launch_bounds__(256) pad_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ members, int64_t n, int k,
                                                       int divisor, int64_t* __restrict__ out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * k; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / k;
    const int c = (int)(t - i * k);
    const int b = rowptr[i], e = rowptr[i + 1];
    out[t] = b + c < e ? (int64_t)(members[b + c] / divisor) : -1;
  }
}
extern "C" int geobi_pad_rows(const int32_t* rowptr, const int32_t* members, int64_t n_rows, int64_t k, int divisor, int64_t* out, void* stream) {
  GEOBI_REQUIRE(rowptr && out && n_rows >= 0 && k > 0 && k < ((int64_t)1 << 31) && divisor > 0 && (members || n_rows == 0), "pad_rows: bad arguments");
  if (n_rows == 0) return GEOBI_OK;
  pad_rows_kernel<<<grid_for(n_rows * k, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rowptr, members, n_rows, (int)k, divisor, out);
  GEOBI_LAUNCH_OK("pad_rows");
  return GEOBI_OK;
}

extern "C" int geobi_group_by(const int32_t* cluster, int64_t n, int64_t nc, int32_t* mrowptr, int32_t* members, void* ws, size_t ws_bytes,
                              void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(cluster && mrowptr && members && n >= 0 && nc >= 0, "group_by: bad arguments");
  if (ws_bytes < geobi_group_by_ws_bytes(n, nc) || !ws) { set_error("group_by: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver c(ws, ws_bytes);
  int* count = c.take<int>(nc + 1);
  int* raw_rowptr = c.take<int>(nc + 1);
  int* vcount = c.take<int>(nc + 1);
  int* ucount = c.take<int>(nc + 1);
  int* status = c.take<int>(ST_WORDS);
  uint64_t* raw = c.take<uint64_t>(n + 1);
  uint64_t* sorted = c.take<uint64_t>(n + 1);
  const size_t sb = scan_ws_bytes(nc + 1);
  char* scan = c.take<char>(sb);
  GEOBI_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int) * (nc + 1), st));
  GEOBI_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int) * ST_WORDS, st));
  const unsigned blocks = (unsigned)cdiv(n > 0 ? n : 1, 256);
  group_count_kernel<<<blocks, 256, 0, st>>>(cluster, n, nc, count, status);
  int rc = scan_i32(count, raw_rowptr, nc, scan, sb, st);
  if (rc) return rc;
  GEOBI_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int) * (nc + 1), st));
  group_fill_kernel<<<blocks, 256, 0, st>>>(cluster, n, nc, raw_rowptr, count, raw);
  GEOBI_LAUNCH_OK("group_by");
  return rows_sort_compact(raw_rowptr, 0, raw, sorted, vcount, ucount, nc, 0, 1, nullptr, 0, 0, 0, mrowptr, members, nullptr, nullptr, status,
                           scan, sb, st);
}

// ------------------------------------------------------------------------------ pool_edges
namespace geobi {
__global__ void pool_count_kernel(const int* __restrict__ rowptr, const int* __restrict__ mrowptr, const int* __restrict__ members,
                                  int64_t nc, int* __restrict__ count) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int s = 0;
  for (int k = mrowptr[c]; k < mrowptr[c + 1]; ++k) {
    const int m = members[k];
    s += rowptr[m + 1] - rowptr[m];
  }
  count[c] = s;
}
// one warp per coarse node: concatenates its members' fine rows, relabelled
__global__ void __launch_bounds__(256) pool_fill_kernel(const int* __restrict__ rowptr, const int* __restrict__ nbr, const float* __restrict__ w,
                                                        const int* __restrict__ cluster, const int* __restrict__ mrowptr,
                                                        const int* __restrict__ members, int64_t nc, const int* __restrict__ raw_rowptr,
                                                        uint64_t* __restrict__ raw, float* __restrict__ raw_w) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nc) return;
  const int64_t start = raw_rowptr[c];
  int pos = 0;
  for (int k = mrowptr[c]; k < mrowptr[c + 1]; ++k) {
    const int m = members[k];
    const int s = rowptr[m], e = rowptr[m + 1];
    for (int q = s + lane; q < e; q += 32) {
      const int cv = cluster[nbr[q]];
      const int p = pos + (q - s);
      raw[start + p] = (cv == (int)c) ? KEY_INVALID : (((uint64_t)(uint32_t)cv << 32) | (uint32_t)p);
      if (raw_w) raw_w[start + p] = w[q];
    }
    pos += e - s;
  }
}
// Fused row builder: one warp per coarse node gathers its members' relabelled fine rows, rank-sorts them, merges
// duplicates (mean weight, summed in key order like compact_rows_kernel) and writes the unique entries to a temporary row
// at the node's raw offset - all out of shared memory for rows of up to PR_ML raw entries (mesh coarsening: ~25-60), out
// of the global scratch arrays beyond that.  Replaces pool_fill + sort_rows + compact_rows (three passes over 64-bit keys
// in global memory) by this kernel and a plain copy once the unique counts have been scanned.
constexpr int PR_ML = 64;
constexpr int PR_W = 16;   // lanes per coarse node: raw rows are 25-50 entries on meshes, a full warp per row idles half its lanes
// `sub`, `half`: lane within / index of the PR_W-lane group; ballots are taken warp-wide and masked to the group.
__device__ __forceinline__ unsigned group_ballot(bool p, int half) { return (__ballot_sync(0xffffffffu, p) >> (half * PR_W)) & ((1u << PR_W) - 1u); }
__device__ __forceinline__ void pool_row(int sub, int half, bool live, int c, int L, int64_t start, const int* __restrict__ rowptr,
                                         const int* __restrict__ nbr, const float* __restrict__ w, const int* __restrict__ cluster, int mb,
                                         int me, const int* __restrict__ members, uint64_t* K, uint64_t* S, float* Wt,
                                         int32_t* __restrict__ tmp_nbr, float* __restrict__ tmp_w, int* __restrict__ ucount) {
  if (live) {
    int pos = 0;
    for (int k = mb; k < me; ++k) {
      const int m = members[k];
      const int s = rowptr[m], e = rowptr[m + 1];
      for (int q = s + sub; q < e; q += PR_W) {
        const int cv = cluster[nbr[q]];
        const int p = pos + (q - s);
        K[p] = (cv == c) ? KEY_INVALID : (((uint64_t)(uint32_t)cv << 32) | (uint32_t)p);
        if (w) Wt[p] = w[q];
      }
      pos += e - s;
    }
  }
  __syncwarp();
  int nvalid = 0;
  for (int i0 = 0; __any_sync(0xffffffffu, live && i0 < L); i0 += PR_W) {
    const int i = i0 + sub;
    const uint64_t key = (live && i < L) ? K[i] : KEY_INVALID;
    const bool valid = key != KEY_INVALID;
    if (valid) {
      int rank = 0;
      for (int k = 0; k < L; ++k) rank += (K[k] < key) ? 1 : 0;
      S[rank] = key;
    }
    nvalid += __popc(group_ballot(valid, half));
  }
  __syncwarp();
  int base = 0;
  for (int t0 = 0; __any_sync(0xffffffffu, t0 < nvalid); t0 += PR_W) {
    const int t = t0 + sub;
    uint64_t key = 0;
    bool head = false;
    if (t < nvalid) {
      key = S[t];
      head = t == 0 || ((key >> 32) != (S[t - 1] >> 32));
    }
    const unsigned mask = group_ballot(head, half);
    if (head) {
      const int64_t o = start + base + __popc(mask & ((1u << sub) - 1u));
      tmp_nbr[o] = (int32_t)(key >> 32);
      if (w) {
        float sum = 0.f;
        int cnt = 0;
        for (int q = t; q < nvalid; ++q) {
          const uint64_t kq = S[q];
          if (q > t && (kq >> 32) != (key >> 32)) break;
          sum += Wt[(uint32_t)kq];
          ++cnt;
        }
        tmp_w[o] = sum / (float)cnt;
      }
    }
    base += __popc(mask);
  }
  if (live && sub == 0) ucount[c] = base;
  __syncwarp();
}

__global__ void __launch_bounds__(256) pool_rows_kernel(const int* __restrict__ rowptr, const int* __restrict__ nbr, const float* __restrict__ w,
                                                        const int* __restrict__ cluster, const int* __restrict__ mrowptr,
                                                        const int* __restrict__ members, int64_t nc, const int* __restrict__ raw_rowptr,
                                                        uint64_t* __restrict__ raw, uint64_t* __restrict__ sorted, float* __restrict__ raw_w,
                                                        int32_t* __restrict__ tmp_nbr, float* __restrict__ tmp_w, int* __restrict__ ucount,
                                                        int* __restrict__ status) {
  constexpr int GROUPS = 256 / PR_W;
  __shared__ uint64_t sk[GROUPS][PR_ML], ss[GROUPS][PR_ML];
  __shared__ float sw[GROUPS][PR_ML];
  const int sub = threadIdx.x % PR_W, grp = threadIdx.x / PR_W, half = grp & (32 / PR_W - 1);
  const int64_t c = (int64_t)blockIdx.x * GROUPS + grp;
  bool live = c < nc;
  int64_t start = 0;
  int L = 0, mb = 0, me = 0;
  if (live) {
    start = raw_rowptr[c];
    L = raw_rowptr[c + 1] - (int)start;
    mb = mrowptr[c];
    me = mrowptr[c + 1];
    if (L > MAX_ROW) {
      if (sub == 0) {
        atomicExch(&status[ST_ERR], GEOBI_ERR_RANGE);
        ucount[c] = 0;
      }
      live = false;
    }
  }
  // the two groups of a warp walk the same code together (warp-wide ballots / barriers); a long row sends both through
  // the global-scratch variant one after the other
  const bool small = !live || L <= PR_ML;
  if (__all_sync(0xffffffffu, small)) {
    pool_row(sub, half, live, (int)c, L, start, rowptr, nbr, w, cluster, mb, me, members, sk[grp], ss[grp], sw[grp], tmp_nbr, tmp_w, ucount);
  } else {
    pool_row(sub, half, live && small, (int)c, L, start, rowptr, nbr, w, cluster, mb, me, members, sk[grp], ss[grp], sw[grp], tmp_nbr, tmp_w,
             ucount);
    pool_row(sub, half, live && !small, (int)c, L, start, rowptr, nbr, w, cluster, mb, me, members, raw + start, sorted + start, raw_w + start,
             tmp_nbr, tmp_w, ucount);
  }
}

// 8 lanes per coarse node: temporary row -> its final place
__global__ void __launch_bounds__(256) pool_copy_kernel(const int* __restrict__ raw_rowptr, const int* __restrict__ out_rowptr, int64_t nc,
                                                        const int32_t* __restrict__ tmp_nbr, const float* __restrict__ tmp_w,
                                                        int32_t* __restrict__ out_nbr, float* __restrict__ out_w) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = gid >> 3;
  if (c >= nc) return;
  const int sub = (int)(gid & 7);
  const int64_t src = raw_rowptr[c], dst = out_rowptr[c];
  const int u = out_rowptr[c + 1] - (int)dst;
  for (int t = sub; t < u; t += 8) {
    out_nbr[dst + t] = tmp_nbr[src + t];
    if (out_w) out_w[dst + t] = tmp_w[src + t];
  }
}

struct PoolWs {
  int *count, *raw_rowptr, *vcount, *ucount, *status;
  uint64_t *raw, *sorted;
  float* raw_w;
  void* scan;
  size_t scan_bytes;
  int32_t* tmp_nbr;
  float* tmp_w;
};
template <class C>
static void carve_pool(C& c, int64_t nnz, int64_t nc, PoolWs* out) {
  int* count = c.template take<int>(nc + 1);
  int* raw_rowptr = c.template take<int>(nc + 1);
  int* vcount = c.template take<int>(nc + 1);
  int* ucount = c.template take<int>(nc + 1);
  int* status = c.template take<int>(ST_WORDS);
  uint64_t* raw = c.template take<uint64_t>(nnz + 1);
  uint64_t* sorted = c.template take<uint64_t>(nnz + 1);
  float* raw_w = c.template take<float>(nnz + 1);
  const size_t sb = scan_ws_bytes(nc + 1);
  char* scan = c.template take<char>(sb);
  int32_t* tmp_nbr = c.template take<int32_t>(nnz + 1);
  float* tmp_w = c.template take<float>(nnz + 1);
  if (out) *out = PoolWs{count, raw_rowptr, vcount, ucount, status, raw, sorted, raw_w, scan, sb, tmp_nbr, tmp_w};
}
}  // namespace geobi

extern "C" size_t geobi_pool_edges_ws_bytes(int64_t nnz_fine, int64_t n_clusters) {
  NullCarver c;
  carve_pool(c, nnz_fine, n_clusters, nullptr);
  return c.s.total();
}

extern "C" int geobi_pool_edges(const int32_t* rowptr, const int32_t* nbr, const float* w, int64_t n_nodes, int64_t nnz_fine, const int32_t* cluster,
                                const int32_t* mrowptr, const int32_t* members, int64_t n_clusters, int32_t* out_rowptr, int32_t* out_nbr,
                                float* out_w, int64_t* nnz_host, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(rowptr && cluster && mrowptr && members && out_rowptr && n_nodes >= 0 && n_clusters >= 0 && nnz_fine >= 0, "pool_edges: bad arguments");
  GEOBI_REQUIRE((w == nullptr) == (out_w == nullptr), "pool_edges: w and out_w must both be given or both be NULL");
  const int64_t h_nnz = nnz_fine;
  if (ws_bytes < geobi_pool_edges_ws_bytes(h_nnz, n_clusters) || !ws) { set_error("pool_edges: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver c(ws, ws_bytes);
  PoolWs W;
  carve_pool(c, h_nnz, n_clusters, &W);
  GEOBI_CUDA_OK(cudaMemsetAsync(W.status, 0, sizeof(int) * ST_WORDS, st));
  if (n_clusters > 0) {
    pool_count_kernel<<<(unsigned)cdiv(n_clusters, 256), 256, 0, st>>>(rowptr, mrowptr, members, n_clusters, W.count);
    GEOBI_LAUNCH_OK("pool_count");
  }
  int rc = scan_i32(W.count, W.raw_rowptr, n_clusters, W.scan, W.scan_bytes, st);
  if (rc) return rc;
  if (getenv("GEOBI_POOL_GENERIC") == nullptr) {
    if (n_clusters > 0) {
      pool_rows_kernel<<<(unsigned)cdiv(n_clusters * PR_W, 256), 256, 0, st>>>(rowptr, nbr, w, cluster, mrowptr, members, n_clusters, W.raw_rowptr,
                                                                            W.raw, W.sorted, W.raw_w, W.tmp_nbr, W.tmp_w, W.ucount, W.status);
      GEOBI_LAUNCH_OK("pool_rows");
    }
    rc = scan_i32(W.ucount, out_rowptr, n_clusters, W.scan, W.scan_bytes, st);
    if (rc) return rc;
    if (n_clusters > 0 && h_nnz > 0) {
      pool_copy_kernel<<<(unsigned)cdiv(n_clusters * 8, 256), 256, 0, st>>>(W.raw_rowptr, out_rowptr, n_clusters, W.tmp_nbr, W.tmp_w, out_nbr, out_w);
      GEOBI_LAUNCH_OK("pool_copy");
    }
    if (nnz_host) return finish_sync(W.status, out_rowptr, n_clusters, nnz_host, "pool_edges", st);
    return GEOBI_OK;
  }
  if (n_clusters > 0) {
    pool_fill_kernel<<<(unsigned)cdiv(n_clusters * 32, 256), 256, 0, st>>>(rowptr, nbr, w, cluster, mrowptr, members, n_clusters, W.raw_rowptr,
                                                                          W.raw, w ? W.raw_w : nullptr);
    GEOBI_LAUNCH_OK("pool_fill");
  }
  rc = rows_sort_compact(W.raw_rowptr, 0, W.raw, W.sorted, W.vcount, W.ucount, n_clusters, 1, 1, w ? W.raw_w : nullptr, /*w_rel=*/1, 0,
                         /*w_mean=*/1, out_rowptr, out_nbr, out_w, nullptr, W.status, W.scan, W.scan_bytes, st);
  if (rc) return rc;
  if (nnz_host) return finish_sync(W.status, out_rowptr, n_clusters, nnz_host, "pool_edges", st);
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ remove_self_loops (order preserving)
namespace geobi {
__global__ void rsl_flag_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t E, int* __restrict__ flag) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) flag[e] = row[e] != col[e] ? 1 : 0;
}
__global__ void rsl_scatter_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, const float* __restrict__ w, int64_t E,
                                   const int* __restrict__ offs, int64_t count, int64_t* __restrict__ out, float* __restrict__ w_out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || row[e] == col[e]) return;
  const int64_t o = offs[e];
  if (o >= count) return;   // the caller's count was too small: never write out of bounds
  out[o] = row[e];
  out[count + o] = col[e];
  if (w_out) w_out[o] = w[e];
}
}  // namespace geobi

extern "C" size_t geobi_remove_self_loops_ws_bytes(int64_t n_edges) {
  return 2 * align256((size_t)(n_edges + 2) * sizeof(int)) + scan_ws_bytes(n_edges + 1) + 256;
}

extern "C" int geobi_remove_self_loops(const int64_t* row, const int64_t* col, const float* w, int64_t n_edges, int64_t count, int64_t* out,
                                       float* w_out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(row && col && n_edges >= 0 && count >= 0 && (out || count == 0), "remove_self_loops: bad arguments");
  GEOBI_REQUIRE((w == nullptr) == (w_out == nullptr), "remove_self_loops: w and w_out must both be given or both be NULL");
  if (n_edges == 0 || count == 0) return GEOBI_OK;
  if (!ws || ws_bytes < geobi_remove_self_loops_ws_bytes(n_edges)) { set_error("remove_self_loops: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver c(ws, ws_bytes);
  int* flag = c.take<int>(n_edges + 2);
  int* offs = c.take<int>(n_edges + 2);
  const size_t sb = scan_ws_bytes(n_edges + 1);
  char* scan = c.take<char>(sb);
  const unsigned blocks = (unsigned)cdiv(n_edges, 256);
  rsl_flag_kernel<<<blocks, 256, 0, st>>>(row, col, n_edges, flag);
  int rc = scan_i32(flag, offs, n_edges, scan, sb, st);
  if (rc) return rc;
  rsl_scatter_kernel<<<blocks, 256, 0, st>>>(row, col, w, n_edges, offs, count, out, w_out);
  GEOBI_LAUNCH_OK("remove_self_loops");
  return GEOBI_OK;
}

// ------------------------------------------------------------------------------ CSR of a coalesced undirected edge list
// When the caller knows its edge list is what to_undirected / coalesce produce (non-loop entries sorted by (row, col), no
// duplicates, (j,i) present for every (i,j)) the CSR needs no counting pass, no atomics and no row sort: dropping the loops
// is an order-preserving compaction and rowptr is the list of row boundaries.  Being symmetric, the same CSR is the
// conv's target-indexed adjacency and the matcher's source-indexed one.  The promise is verified on the device; a list
// that breaks it poisons rowptr[n_nodes] with -1, which the caller sees with the entry count it has to read anyway.
namespace geobi {
__global__ void sorted_scatter_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, const float* __restrict__ w, int64_t E,
                                      int64_t n, const int* __restrict__ offs, int* __restrict__ rows32, int32_t* __restrict__ nbr,
                                      float* __restrict__ w_out, int64_t* __restrict__ ei_out, int* __restrict__ bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t r = row[e], c = col[e];
  if (r < 0 || r >= n || c < 0 || c >= n) {
    *bad = 1;
    return;
  }
  if (r == c) return;
  const int64_t o = offs[e];
  rows32[o] = (int)r;
  nbr[o] = (int)c;
  if (w_out) w_out[o] = w[e];
  if (ei_out) {
    ei_out[o] = r;
    ei_out[E + o] = c;
  }
}
// thread p in [0, nnz]: rowptr entries of the rows that start between entries p-1 and p; order check of the pair
__global__ void sorted_rowptr_kernel(const int* __restrict__ rows32, const int32_t* __restrict__ nbr, const int* __restrict__ offs, int64_t E,
                                     int64_t n, int32_t* __restrict__ rowptr, int* __restrict__ bad) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nnz = offs[E];
  if (p > nnz || *bad) return;            // out-of-range entries (found by the scatter) left holes in rows32: do not walk them
  const int64_t r = p < nnz ? rows32[p] : n;
  const int64_t prev = p > 0 ? rows32[p - 1] : -1;
  if (p > 0 && p < nnz && (prev > r || (prev == r && nbr[p - 1] >= nbr[p]))) *bad = 1;   // not sorted / duplicate
  for (int64_t q = prev + 1; q <= r; ++q) rowptr[q] = (int)p;
}
__global__ void sorted_symmetry_kernel(const int* __restrict__ rows32, const int32_t* __restrict__ nbr, const int* __restrict__ offs, int64_t E,
                                       const int32_t* __restrict__ rowptr, int* __restrict__ bad) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= offs[E] || *bad) return;       // an unsorted list has no usable rowptr to search in
  const int r = rows32[p], c = nbr[p];
  int lo = rowptr[c], hi = rowptr[c + 1];            // find r in row c
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (nbr[mid] < r) lo = mid + 1;
    else hi = mid;
  }
  if (lo >= rowptr[c + 1] || nbr[lo] != r) *bad = 1;
}
// a list that broke its promise leaves an EMPTY graph behind (every row [0, 0), count -1): whatever runs on it before the
// host looks at the count touches no adjacency memory
__global__ void sorted_verdict_kernel(const int* __restrict__ bad, int32_t* __restrict__ rowptr, int64_t n) {
  if (!*bad) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) rowptr[i] = i == n ? -1 : 0;
}
}  // namespace geobi

extern "C" size_t geobi_csr_from_sorted_coo_ws_bytes(int64_t n_edges) {
  return 3 * align256((size_t)(n_edges + 2) * sizeof(int)) + scan_ws_bytes(n_edges + 1) + 512;
}

extern "C" int geobi_csr_from_sorted_coo(const int64_t* row, const int64_t* col, const float* w, int64_t n_edges, int64_t n_nodes, int flags,
                                         int32_t* rowptr, int32_t* nbr, float* w_out, int64_t* ei_out, void* ws, size_t ws_bytes,
                                         void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(n_edges >= 0 && n_nodes >= 0 && rowptr != nullptr, "csr_from_sorted_coo: bad arguments");
  GEOBI_REQUIRE(n_edges == 0 || (row && col && nbr), "csr_from_sorted_coo: null edge arrays");
  GEOBI_REQUIRE((w == nullptr) == (w_out == nullptr), "csr_from_sorted_coo: w and w_out must both be given or both be NULL");
  GEOBI_REQUIRE(n_edges < ((int64_t)1 << 31) && n_nodes < ((int64_t)1 << 31), "csr_from_sorted_coo: int32 CSR overflow");
  if (n_edges == 0) {
    GEOBI_CUDA_OK(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * (size_t)(n_nodes + 1), st));
    return GEOBI_OK;
  }
  if (!ws || ws_bytes < geobi_csr_from_sorted_coo_ws_bytes(n_edges)) { set_error("csr_from_sorted_coo: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver c(ws, ws_bytes);
  int* flag = c.take<int>(n_edges + 2);
  int* offs = c.take<int>(n_edges + 2);
  int* rows32 = c.take<int>(n_edges + 2);
  int* bad = c.take<int>(64);
  const size_t sb = scan_ws_bytes(n_edges + 1);
  char* scan = c.take<char>(sb);
  GEOBI_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), st));
  const unsigned blocks = (unsigned)cdiv(n_edges + 1, 256);
  rsl_flag_kernel<<<blocks, 256, 0, st>>>(row, col, n_edges, flag);
  int rc = scan_i32(flag, offs, n_edges, scan, sb, st);
  if (rc) return rc;
  sorted_scatter_kernel<<<blocks, 256, 0, st>>>(row, col, w, n_edges, n_nodes, offs, rows32, nbr, w_out, ei_out, bad);
  sorted_rowptr_kernel<<<blocks, 256, 0, st>>>(rows32, nbr, offs, n_edges, n_nodes, rowptr, bad);
  if (flags & GEOBI_SORTED_CHECK_SYMMETRIC) sorted_symmetry_kernel<<<blocks, 256, 0, st>>>(rows32, nbr, offs, n_edges, rowptr, bad);
  sorted_verdict_kernel<<<(unsigned)cdiv(n_nodes + 1, 256), 256, 0, st>>>(bad, rowptr, n_nodes);
  GEOBI_LAUNCH_OK("csr_from_sorted_coo");
  return GEOBI_OK;
}
