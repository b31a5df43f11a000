// One iteration of PoolingLayer's coarsening loop after the matching (net_util.py:100-140) as a single host call:
//   consecutive_cluster -> [the cluster count comes back: the only sync] -> member CSR -> cluster max/mean of the features
//   (and mean of the positions) -> coarse adjacency (pool_edge: relabel, drop loops, coalesce with mean weights).
// Every output is a capacity-sized buffer the caller allocates BEFORE the call (the cluster count is at most the node
// count), so the kernels that depend on the count are queued from C++ microseconds after the synchronisation returns,
// instead of after the host language has sized and allocated their outputs (~50 us of idle GPU per step in the timeline).
#include "common.cuh"

using namespace geobi;

extern "C" size_t geobi_pool_step_ws_bytes(int64_t n_nodes, int64_t nnz_cap) {
  size_t a = geobi_relabel_ws_bytes(n_nodes);
  const size_t b = geobi_group_pairs_ws_bytes(n_nodes);
  const size_t c = geobi_pool_edges_ws_bytes(nnz_cap, n_nodes);
  if (b > a) a = b;
  if (c > a) a = c;
  return a + 256;
}

extern "C" int geobi_pool_step(const int32_t* rowptr, const int32_t* nbr, const float* w, int64_t n_nodes, int64_t nnz_cap, const int32_t* label,
                               const float* x, int64_t ldx, int channels, int op, const float* pos, int64_t ldp, int pos_channels,
                               int32_t* cluster, int32_t* mrowptr, int32_t* members, float* x_out, int64_t ldxo, float* pos_out, int64_t ldpo,
                               int32_t* out_rowptr, int32_t* out_nbr, float* out_w, int64_t* n_clusters_host, void* ws, size_t ws_bytes,
                               void* stream) {
  GEOBI_REQUIRE(rowptr && label && cluster && mrowptr && members && out_rowptr && n_clusters_host && n_nodes >= 0 && nnz_cap >= 0,
                "pool_step: null argument");
  GEOBI_REQUIRE(nnz_cap == 0 || (nbr && out_nbr), "pool_step: null adjacency");
  GEOBI_REQUIRE((w == nullptr) == (out_w == nullptr) || nnz_cap == 0, "pool_step: w and out_w must both be given or both be NULL");
  GEOBI_REQUIRE((x == nullptr) == (x_out == nullptr) && (pos == nullptr) == (pos_out == nullptr), "pool_step: feature / output mismatch");
  if (!ws || ws_bytes < geobi_pool_step_ws_bytes(n_nodes, nnz_cap)) {
    set_error("pool_step: workspace too small");
    return GEOBI_ERR_WORKSPACE;
  }
  // the three stages run back to back on one stream, so they can share the workspace
  int rc = geobi_relabel_clusters(label, n_nodes, cluster, n_clusters_host, ws, ws_bytes, stream);   // SYNCS: cluster count
  if (rc) return rc;
  const int64_t nc = *n_clusters_host;
  rc = geobi_group_pairs(label, cluster, n_nodes, nc, mrowptr, members, ws, ws_bytes, stream);
  if (rc) return rc;
  if (x && nc > 0) {
    rc = geobi_segment_reduce(x, ldx, channels, mrowptr, members, 0, nc, op, x_out, ldxo, stream);
    if (rc) return rc;
  }
  rc = geobi_pool_edges(rowptr, nbr, w, n_nodes, nnz_cap, cluster, mrowptr, members, nc, out_rowptr, out_nbr, out_w, nullptr, ws, ws_bytes, stream);
  if (rc) return rc;
  if (pos && nc > 0) rc = geobi_segment_reduce(pos, ldp, pos_channels, mrowptr, members, 0, nc, /*mean*/ 0, pos_out, ldpo, stream);
  return rc;
}
