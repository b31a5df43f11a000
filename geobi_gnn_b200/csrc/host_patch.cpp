// Host-side (CPU, C++) mesh splitting for patch-sharded inference: the reference grows patches with pure-Python loops
// (/root/reference/code/data_util.py:55-84, dataset.py:156-193), which takes minutes on a 10 M-face mesh.  Same
// sequential semantics (ring-by-ring BFS, discovery order, cut at exactly `neighbor_count` faces), linear time.
// Built into libgeobi_host.so with g++ (no CUDA): this is input preparation, not the GPU hot path.
#include <stdint.h>

#include <math.h>

#include <algorithm>
#include <thread>
#include <vector>

extern "C" {

// Returns the number of faces written to out (<= neighbor_count, <= n_faces).
// fstamp [n_faces] / vstamp [n_vertices]: scratch stamp arrays owned by the caller; an entry equal to `epoch` means "taken by
// this patch" / "vertex already expanded".  The caller zero-fills them once and passes a new non-zero epoch per call, so
// nothing has to be cleaned up between patches.  A vertex whose incident faces were all offered once cannot contribute
// again (every one of them is already taken), so later visits skip its vf row: the same faces in the same order as the
// reference's loops, with one row read per vertex instead of ~6.
__attribute__((visibility("default"))) int64_t geobi_host_grow_patch(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t k,
                                                                      int64_t seed, int64_t neighbor_count, int64_t ring_count,
                                                                      uint32_t* fstamp, uint32_t* vstamp, uint32_t epoch, int64_t* out) {
  if (seed < 0 || seed >= n_faces || neighbor_count <= 0) return 0;
  int64_t n = 0;
  out[n++] = seed;
  fstamp[seed] = epoch;
  int64_t lo = 0, hi = 1;
  bool full = false;  // the reference only checks the count after an insertion
  for (int64_t ring = 0; ring < ring_count && !full; ++ring) {
    for (int64_t q = lo; q < hi && !full; ++q) {
      const int64_t face = out[q];
      if (q + 8 < hi) __builtin_prefetch(fv + out[q + 8] * 3);
      for (int c = 0; c < 3 && !full; ++c) {
        const int64_t v = fv[face * 3 + c];
        if (vstamp[v] == epoch) continue;
        vstamp[v] = epoch;
        for (int64_t t = 0; t < k; ++t) {
          const int64_t g = vf[v * k + t];
          if (g < 0) break;
          if (fstamp[g] != epoch) {
            out[n++] = g;
            fstamp[g] = epoch;
            if (n >= neighbor_count) { full = true; break; }
          }
        }
      }
    }
    lo = hi;
    hi = n;
    if (lo == hi) break;
  }
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

// Squared distance of every face centre to `centroid`, in the fp32 operation order numpy uses for
// `((pts[fv].mean(1) - centroid) ** 2).sum(1)` (dataset.py:165-166): ((a+b)+c)/3 per coordinate, subtract, square, (x+y)+z.
// The seeds of the patch splitter are arg-maxima of this array, and on a near-spherical mesh they are decided by the last
// bit, so the order is part of the contract (built with -ffp-contract=off; tests compare with numpy bit for bit).
__attribute__((visibility("default"))) void geobi_host_face_d2(const float* pts, const int64_t* fv, int64_t n_faces, const float* centroid,
                                                                float* out, int n_threads) {
  auto body = [=](int64_t lo, int64_t hi) {
    for (int64_t f = lo; f < hi; ++f) {
      const float* a = pts + 3 * fv[3 * f];
      const float* b = pts + 3 * fv[3 * f + 1];
      const float* c = pts + 3 * fv[3 * f + 2];
      float d[3];
      for (int k = 0; k < 3; ++k) {
        const float m = ((a[k] + b[k]) + c[k]) / 3.0f;
        const float e = m - centroid[k];
        d[k] = e * e;
      }
      out[f] = (d[0] + d[1]) + d[2];
    }
  };
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_faces / 65536));
  if (nt == 1) return body(0, n_faces);
  std::vector<std::thread> th;
  const int64_t per = (n_faces + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) th.emplace_back(body, std::min(n_faces, t * per), std::min(n_faces, (t + 1) * per));
  for (auto& t : th) t.join();
}

// Book-keeping between two patches of the splitter (dataset.py:186-192): faces of `sel` leave the "uncovered" set (their
// entry of d2_left drops to -inf), *n_left is decremented by the number that were still uncovered, and the next seed =
// first arg-max of what is left (np.argmax semantics: the lowest index among equal maxima) is returned.
__attribute__((visibility("default"))) int64_t geobi_host_cover_next_seed(float* d2_left, int64_t n_faces, const int64_t* sel, int64_t n_sel,
                                                                           int64_t* n_left, int n_threads) {
  const float ninf = -__builtin_inff();
  int64_t fresh = 0;
  for (int64_t i = 0; i < n_sel; ++i) {
    float& d = d2_left[sel[i]];
    fresh += d != ninf;
    d = ninf;
  }
  *n_left -= fresh;
  if (*n_left <= 0) return -1;
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_faces / 65536));
  std::vector<int64_t> best(nt, 0);
  const int64_t per = (n_faces + nt - 1) / nt;
  auto body = [&](int t) {
    const int64_t lo = std::min(n_faces, t * per), hi = std::min(n_faces, (t + 1) * per);
    int64_t b = lo;
    for (int64_t i = lo + 1; i < hi; ++i)
      if (d2_left[i] > d2_left[b]) b = i;
    best[t] = b;
  };
  if (nt == 1) {
    body(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(body, t);
    for (auto& t : th) t.join();
  }
  int64_t b = best[0];
  for (int t = 1; t < nt; ++t)
    if (best[t] < n_faces && d2_left[best[t]] > d2_left[b]) b = best[t];
  return b;
}

// Edge lengths of the centred mesh, per edge, in numpy's fp32 operation order for
// `q = p - centroid; e = q[ev]; ((e[:, 0] - e[:, 1]) ** 2).sum(1) ** 0.5` (dataset.py:140,151-152); the caller takes numpy's
// own mean of the result, so the normalisation scale is the reference's to the bit while the 15 M-edge temporaries of a
// 10 M-face mesh (2.4 s of numpy) become one threaded pass.
__attribute__((visibility("default"))) void geobi_host_edge_lengths(const float* pts, const float* centroid, const int64_t* ev, int64_t n_edges,
                                                                     float* out, int n_threads) {
  auto body = [=](int64_t lo, int64_t hi) {
    for (int64_t e = lo; e < hi; ++e) {
      const float* a = pts + 3 * ev[2 * e];
      const float* b = pts + 3 * ev[2 * e + 1];
      float d[3];
      for (int k = 0; k < 3; ++k) {
        const float t = (a[k] - centroid[k]) - (b[k] - centroid[k]);
        d[k] = t * t;
      }
      out[e] = sqrtf((d[0] + d[1]) + d[2]);
    }
  };
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_edges / 65536));
  if (nt == 1) return body(0, n_edges);
  std::vector<std::thread> th;
  const int64_t per = (n_edges + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) th.emplace_back(body, std::min(n_edges, t * per), std::min(n_edges, (t + 1) * per));
  for (auto& t : th) t.join();
}
}
