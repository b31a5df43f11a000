/* CPU oracle: serial greedy heavy-edge matching.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates the CPU kernel of torch_cluster.graclus (third-party, un-vendored,
 * unpinned; reached from /root/reference/code/net_util.py:127 through
 * torch_geometric.nn.graclus).  Published algorithm, as recalled in
 * SURVEY.md section 8(c):
 *
 *   for u in perm:                       (perm = randperm(N) upstream; explicit here)
 *     if u already labelled: continue
 *     label[u] = u
 *     scan row u of the CSR; among still-unlabelled neighbours v keep the one
 *     with weight >= running max (running max starts at 0, so later ties win
 *     and non-positive... strictly negative weights never match);
 *     unweighted: take the first unlabelled neighbour.
 *     if found: label[u] = label[v] = min(u, v)
 *
 * PARITY UNPINNED: the reference has no tests or vectors for this (third-party
 * algorithm; cross-checked only against the pure-Python loop in oracle/pyg.py).
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <stdint.h>

int oracle_graclus_greedy(int64_t n, const int64_t* rowptr, const int64_t* col,
                          const float* weight /* may be NULL */,
                          const int64_t* perm, int64_t* label) {
  for (int64_t i = 0; i < n; ++i) label[i] = -1;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t u = perm[k];
    if (u < 0 || u >= n) return -1;
    if (label[u] >= 0) continue;
    label[u] = u;
    int64_t best = -1;
    float wmax = 0.0f;
    for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {
      const int64_t v = col[e];
      if (label[v] >= 0) continue;
      if (!weight) { best = v; break; }
      if (weight[e] >= wmax) { best = v; wmax = weight[e]; }
    }
    if (best >= 0) {
      const int64_t m = u < best ? u : best;
      label[u] = m;
      label[best] = m;
    }
  }
  return 0;
}

/* Incident-face vertex update, the scalar statement of update_position2
 * (/root/reference/code/data_util.py:529-556): n_iter Jacobi sweeps of
 *   p_v += (1/max(deg_v,1)) * sum_{f in vf[v], f>=0} n_f * (n_f . (c_f - p_v)),
 * c_f = mean of the 3 corners, optionally projected on depth_direction.
 * Used to cross-check the tensor restatement and as a timed CPU baseline. */
void oracle_update_position2(int64_t V, int64_t F, int64_t maxval, const float* points_in,
                             const int64_t* fv, const int64_t* vf, const float* fnormal,
                             int n_iter, const float* depth /* may be NULL */,
                             float* cent /* scratch F*3 */, float* p /* out V*3 */,
                             float* q /* scratch V*3 */) {
  for (int64_t i = 0; i < V * 3; ++i) p[i] = points_in[i];
  for (int it = 0; it < n_iter; ++it) {
    for (int64_t f = 0; f < F; ++f)
      for (int c = 0; c < 3; ++c)
        cent[f * 3 + c] = (p[fv[f * 3] * 3 + c] + p[fv[f * 3 + 1] * 3 + c] + p[fv[f * 3 + 2] * 3 + c]) / 3.0f;
    for (int64_t v = 0; v < V; ++v) {
      float acc[3] = {0.f, 0.f, 0.f};
      int64_t cnt = 0;
      for (int64_t k = 0; k < maxval; ++k) {
        const int64_t f = vf[v * maxval + k];
        if (f < 0) continue;
        ++cnt;
        float d = 0.f;
        for (int c = 0; c < 3; ++c) d += fnormal[f * 3 + c] * (cent[f * 3 + c] - p[v * 3 + c]);
        for (int c = 0; c < 3; ++c) acc[c] += fnormal[f * 3 + c] * d;
      }
      if (cnt < 1) cnt = 1;
      for (int c = 0; c < 3; ++c) acc[c] /= (float)cnt;
      if (depth) {
        float s = 0.f;
        for (int c = 0; c < 3; ++c) s += acc[c] * depth[v * 3 + c];
        for (int c = 0; c < 3; ++c) acc[c] = s * depth[v * 3 + c];
      }
      for (int c = 0; c < 3; ++c) q[v * 3 + c] = p[v * 3 + c] + acc[c];
    }
    for (int64_t i = 0; i < V * 3; ++i) p[i] = q[i];
  }
}
