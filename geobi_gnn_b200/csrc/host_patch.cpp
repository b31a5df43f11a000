// Host-side (CPU, C++) mesh splitting for patch-sharded inference: the reference grows patches with pure-Python loops
// (/root/reference/code/data_util.py:55-84, dataset.py:156-193), which takes minutes on a 10 M-face mesh.  Same
// sequential semantics (ring-by-ring BFS, discovery order, cut at exactly `neighbor_count` faces), linear time.
// Built into libgeobi_host.so with g++ (no CUDA): this is input preparation, not the GPU hot path.
#include <stdint.h>

#include <vector>

extern "C" {

// Returns the number of faces written to out (<= neighbor_count, <= n_faces).  taken: scratch byte array [n_faces], zeroed
// by the caller and restored to zero on return.
__attribute__((visibility("default"))) int64_t geobi_host_grow_patch(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t k,
                                                                      int64_t seed, int64_t neighbor_count, int64_t ring_count,
                                                                      uint8_t* taken, int64_t* out) {
  if (seed < 0 || seed >= n_faces || neighbor_count <= 0) return 0;
  int64_t n = 0;
  out[n++] = seed;
  taken[seed] = 1;
  int64_t lo = 0, hi = 1;
  bool full = n >= neighbor_count && false;  // the reference only checks the count after an insertion
  for (int64_t ring = 0; ring < ring_count && !full; ++ring) {
    for (int64_t q = lo; q < hi && !full; ++q) {
      const int64_t face = out[q];
      for (int c = 0; c < 3 && !full; ++c) {
        const int64_t v = fv[face * 3 + c];
        for (int64_t t = 0; t < k; ++t) {
          const int64_t g = vf[v * k + t];
          if (g < 0) break;
          if (!taken[g]) {
            out[n++] = g;
            taken[g] = 1;
            if (n >= neighbor_count) { full = true; break; }
          }
        }
      }
    }
    lo = hi;
    hi = n;
    if (lo == hi) break;
  }
  for (int64_t i = 0; i < n; ++i) taken[out[i]] = 0;
  return n;
}

// First-appearance re-indexing of the selected faces (data_util.get_submesh, data_util.py:318-336).
// slot: scratch int64 [n_verts] filled with -1 by the caller, restored on return.  Returns the number of vertices.
__attribute__((visibility("default"))) int64_t geobi_host_submesh(const int64_t* fv, const int64_t* select, int64_t n_select, int64_t* slot,
                                                                   int64_t* v_idx, int64_t* faces_out) {
  int64_t nv = 0;
  for (int64_t i = 0; i < n_select; ++i)
    for (int c = 0; c < 3; ++c) {
      const int64_t v = fv[select[i] * 3 + c];
      if (slot[v] < 0) {
        slot[v] = nv;
        v_idx[nv++] = v;
      }
      faces_out[i * 3 + c] = slot[v];
    }
  for (int64_t i = 0; i < nv; ++i) slot[v_idx[i]] = -1;
  return nv;
}
}
