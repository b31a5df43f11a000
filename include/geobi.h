/* geobi.h — C ABI of libgeobi.so: the B200 (sm_100a) kernels behind the GeoBi-GNN
 * dual-domain forward.
 *
 * The reference (zhangyk18/GeoBi-GNN) has NO native / FFI boundary: its seam is the
 * Python module surface imported by train_dual.py / test_dual.py.  This
 * header is therefore the boundary a maintainer would bind underneath that surface;
 * every entry point cites the reference expression(s) it replaces (file:line under
 * /root/reference/code/).  The Python binding is geobi_gnn_b200/_lib.py (ctypes); see
 * INTEGRATION.md.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *  - Features are fp32 row-major with an explicit leading dimension (`ld*`, in floats),
 *    so a conv can write straight into a column block of a wider buffer (no torch.cat).
 *  - Graphs are CSR with int32 indices: rowptr[N+1], nbr[nnz]; no self loops unless said.
 *    Reference-facing edge lists are int64 [2,E] row-major (edge_index) as in PyG.
 *  - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous on that stream
 *    except the ones documented "syncs" (they return a size to the host).
 *  - No allocation inside: scratch comes from the caller (`ws`, `ws_bytes`), sized by the
 *    matching *_ws_bytes() query.  Workspaces need 256-byte alignment.
 *  - Return 0 on success, <0 on error (GEOBI_ERR_*); geobi_last_error() gives the text.
 *  - There is no CPU fallback anywhere in this library.
 */
#ifndef GEOBI_H
#define GEOBI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GEOBI_API __attribute__((visibility("default")))
#else
#define GEOBI_API
#endif

#define GEOBI_OK 0
#define GEOBI_ERR_INVALID (-1)  /* bad argument (shape, null, unsupported channel count) */
#define GEOBI_ERR_CUDA (-2)     /* CUDA runtime error (launch, memcpy)                   */
#define GEOBI_ERR_WORKSPACE (-3) /* ws_bytes too small                                    */
#define GEOBI_ERR_RANGE (-4)    /* index out of range / row too long, found on device    */
#define GEOBI_ERR_NOCONVERGE (-5)

#define GEOBI_HEADS 9           /* FeaStConv(heads=9): network.py:258-268 */

GEOBI_API const char* geobi_last_error(void);
GEOBI_API int geobi_version(void);
/* sm count and compute capability of the current device (host outputs). */
GEOBI_API int geobi_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);

/* ------------------------------------------------------------------ integer / graph */

/* out[i] = sum_{k<i} in[k], i = 0..n  (n+1 outputs).  Building block of every CSR here
 * (replaces torch.cumsum in torch_cluster's rowptr build and torch.unique's counting). */
GEOBI_API size_t geobi_scan_ws_bytes(int64_t n);
GEOBI_API int geobi_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, void* stream);

/* COO (int64 edge_index rows) -> CSR.
 * Replaces: torch_sparse.coalesce (net_util.py:263,294; data_util.py:432,455),
 * torch_geometric.utils.remove_self_loops (net_util.py:163,292), to_undirected
 * (dataset.py:212) and the CSR build inside torch_cluster.graclus (net_util.py:127).
 *   segment key  = row[e] (or col[e] with GEOBI_COO_BY_COL), neighbour = the other end.
 *   GEOBI_COO_DROP_SELF   drop row==col
 *   GEOBI_COO_SORT_NBR    order each CSR row by (neighbour, edge id); otherwise by edge id
 *                         (= a stable sort by segment key, the order graclus sees)
 *   GEOBI_COO_DEDUP       merge equal (key, neighbour) pairs (needs SORT_NBR); weights
 *                         reduced with GEOBI_COO_W_MEAN (pool_edge's op) or summed.
 *   GEOBI_COO_SYMMETRIZE  also insert every (col,row): to_undirected.
 * Outputs: rowptr[N+1], nbr[cap], w_out[cap] (if w), eid_out[cap] (optional: source edge id
 * of each CSR entry, -1-e for a flipped copy), cap >= E (2E with SYMMETRIZE).
 * nnz = rowptr[N] stays on the device; *nnz_host (optional) is filled -> then SYNCS. */
#define GEOBI_COO_BY_COL 1
#define GEOBI_COO_DROP_SELF 2
#define GEOBI_COO_SORT_NBR 4
#define GEOBI_COO_DEDUP 8
#define GEOBI_COO_W_MEAN 16
#define GEOBI_COO_SYMMETRIZE 32
GEOBI_API size_t geobi_csr_from_coo_ws_bytes(int64_t n_edges, int64_t n_nodes, int flags);
GEOBI_API int geobi_csr_from_coo(const int64_t* row, const int64_t* col, const float* w, int64_t n_edges,
                       int64_t n_nodes, int flags, int32_t* rowptr, int32_t* nbr, float* w_out,
                       int64_t* eid_out, int64_t* nnz_host, void* ws, size_t ws_bytes, void* stream);

/* CSR of an edge list the caller knows to be coalesced and undirected — what the reference's dataset builds with
 * to_undirected (+ add_self_loops appended last, dataset.py:211-213) and build_facet_graph (data_util.py:436-456): the
 * non-loop entries are sorted by (row, col) without duplicates and (j,i) is present for every (i,j).  Self loops are
 * dropped (net_util.py:163; FeaStConv re-adds its own) by an order-preserving compaction, rowptr is the list of row
 * boundaries: no counting pass, no atomics, no row sort.  The result serves the conv (target-indexed) and the matcher
 * (source-indexed, entry order == edge order, so w_out lines up with it).  ei_out (optional, int64 [2, n_edges] with row
 * stride n_edges) receives the stripped edge list.  The entry count stays on the device in rowptr[n_nodes]; if the list
 * breaks the promise (out of range, unsorted, duplicate, or - with GEOBI_SORTED_CHECK_SYMMETRIC - a missing reverse edge)
 * rowptr[n_nodes] is set to -1 instead.  Asynchronous. */
#define GEOBI_SORTED_CHECK_SYMMETRIC 1
GEOBI_API size_t geobi_csr_from_sorted_coo_ws_bytes(int64_t n_edges);
GEOBI_API int geobi_csr_from_sorted_coo(const int64_t* row, const int64_t* col, const float* w, int64_t n_edges,
                              int64_t n_nodes, int flags, int32_t* rowptr, int32_t* nbr, float* w_out,
                              int64_t* ei_out, void* ws, size_t ws_bytes, void* stream);

/* torch_geometric.utils.remove_self_loops (net_util.py:163,292) with the surviving count already known to the caller
 * (e.g. the nnz of a DROP_SELF CSR of the same list): order-preserving compaction into out [2,count] (+ w_out), no sync. */
GEOBI_API size_t geobi_remove_self_loops_ws_bytes(int64_t n_edges);
GEOBI_API int geobi_remove_self_loops(const int64_t* row, const int64_t* col, const float* w, int64_t n_edges, int64_t count,
                                      int64_t* out, float* w_out, void* ws, size_t ws_bytes, void* stream);

/* CSR -> int64 edge_index [2, nnz] (row-major sorted, the layout coalesce returns). */
GEOBI_API int geobi_csr_to_coo(const int32_t* rowptr, const int32_t* nbr, int64_t n_nodes, int64_t nnz,
                     int64_t* edge_index, void* stream);

/* Facet 1-ring graph, self included: data_util.build_facet_graph (data_util.py:436-456).
 * fv [F,3], vf [V,K] int64 (pad -1).  nbr capacity 3*K*F.  SYNCS if nnz_host != NULL. */
GEOBI_API size_t geobi_build_facet_graph_ws_bytes(int64_t n_faces, int64_t k);
GEOBI_API int geobi_build_facet_graph(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t n_verts,
                            int64_t k, int32_t* rowptr, int32_t* nbr, int64_t* nnz_host, void* ws,
                            size_t ws_bytes, void* stream);
/* The same graph when every vf row is ASCENDING with its -1 pads at the end (topology.DeviceTriMesh): a three-way merge per face
 * instead of a fill + per-row sort.  An unsorted row is detected on the device (GEOBI_ERR_RANGE at the sync).  drop_self != 0 leaves
 * the self entry out: the loop-free CSR the convolution and the matcher walk, without the list round trip. */
GEOBI_API size_t geobi_build_facet_graph_sorted_ws_bytes(int64_t n_faces);
GEOBI_API int geobi_build_facet_graph_sorted(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t n_verts, int64_t k, int drop_self,
                                             int32_t* rowptr, int32_t* nbr, int64_t* nnz_host, void* ws, size_t ws_bytes, void* stream);

/* Heavy-edge matching identical to torch_cluster.graclus's serial CPU kernel for the visiting
 * order `perm` (net_util.py:127), computed in parallel: a node acts when it precedes all its
 * undecided neighbours and also all undecided neighbours of its chosen partner, claiming the partner with a CAS.
 * w may be NULL (unweighted: first free neighbour).
 * Adjacency must be symmetric (true for every graph on this path).  label[u] = min(u, partner).
 * rank: any int32 priority keys, u is visited before v iff (rank[u], u) < (rank[v], v) — an inverse permutation reproduces
 * torch_cluster's order exactly, i.i.d. random keys give a uniformly random order without a sort.
 * One launch, no grid barriers (resident threads re-evaluate their nodes until decided; the result does not depend on
 * timing); asynchronous on the stream unless undecided_host != NULL (then SYNCS and
 * reports nodes left undecided, 0 on success).  A label < 0 is caught by geobi_relabel_clusters (GEOBI_ERR_RANGE). */
GEOBI_API size_t geobi_graclus_ws_bytes(int64_t n_nodes);
GEOBI_API int geobi_graclus(const int32_t* rowptr, const int32_t* nbr, const float* w, const int32_t* rank,
                            int64_t n_nodes, int32_t* label, int* undecided_host, void* ws, size_t ws_bytes, void* stream);

/* torch_geometric consecutive_cluster (net_util.py:128): dense relabel of labels in [0,N) by
 * ascending label value.  SYNCS: *n_clusters_host. */
GEOBI_API size_t geobi_relabel_ws_bytes(int64_t n_nodes);
GEOBI_API int geobi_relabel_clusters(const int32_t* label, int64_t n_nodes, int32_t* cluster, int64_t* n_clusters_host,
                           void* ws, size_t ws_bytes, void* stream);

/* Members of each cluster as a CSR (ascending node id): the index that turns
 * scatter(x, cluster, reduce=max|mean) (net_util.py:131-134) into a segment reduce. */
GEOBI_API size_t geobi_group_by_ws_bytes(int64_t n_nodes, int64_t n_clusters);
GEOBI_API int geobi_group_by(const int32_t* cluster, int64_t n_nodes, int64_t n_clusters, int32_t* mrowptr,
                   int32_t* members, void* ws, size_t ws_bytes, void* stream);

/* Vertex adjacency of a triangle mesh = to_undirected(unique edges) of dataset.py:211 without the self loops, from the incidence CSR
 * (geobi_group_by over fv.reshape(-1): `corners` are flat corner indices 3 f + c): symmetric, rows ascending, duplicate-free - the CSR
 * geobi_csr_from_coo(SYMMETRIZE | SORT_NBR | DEDUP | DROP_SELF) builds from the 3F half edges, in two kernels + a scan.  Valences
 * above 24 are rejected on the device (GEOBI_ERR_RANGE): use the general builder.  nbr capacity 6 F.  SYNCS if nnz_host != NULL. */
GEOBI_API size_t geobi_mesh_vertex_csr_ws_bytes(int64_t n_verts);
GEOBI_API int geobi_mesh_vertex_csr(const int64_t* fv, const int32_t* vf_rowptr, const int32_t* corners, int64_t n_verts,
                          int64_t n_faces, int32_t* rowptr, int32_t* nbr, int64_t* nnz_host, void* ws, size_t ws_bytes,
                          void* stream);

/* Member CSR -> padded [n_rows, k] int64 table (entry / divisor, -1 pads): OpenMesh's vf_indices / vv_indices layout
 * (dataset.py:204-206) from geobi_group_by's output (divisor 3: corner index -> face id) or from a vertex CSR (divisor 1). */
GEOBI_API int geobi_pad_rows(const int32_t* rowptr, const int32_t* members, int64_t n_rows, int64_t k, int divisor,
                   int64_t* out, void* stream);

/* Same member CSR when the labels come from a matching (clusters of one or two nodes, label = min member, as
 * geobi_graclus emits them): no sort, three elementwise kernels.  label = raw labels, cluster = dense ids. */
GEOBI_API size_t geobi_group_pairs_ws_bytes(int64_t n_clusters);
GEOBI_API int geobi_group_pairs(const int32_t* label, const int32_t* cluster, int64_t n_nodes, int64_t n_clusters,
                                int32_t* mrowptr, int32_t* members, void* ws, size_t ws_bytes, void* stream);

/* net_util.pool_edge (net_util.py:289-295): relabel by cluster, drop loops, coalesce with MEAN
 * weights; emitted directly as the coarse CSR (rows sorted by neighbour = coalesce order).
 * out capacity = nnz of the fine graph.  SYNCS if nnz_host != NULL. */
GEOBI_API size_t geobi_pool_edges_ws_bytes(int64_t nnz_fine, int64_t n_clusters);
GEOBI_API int geobi_pool_edges(const int32_t* rowptr, const int32_t* nbr, const float* w, int64_t n_nodes,
                     int64_t nnz_fine, const int32_t* cluster, const int32_t* mrowptr, const int32_t* members,
                     int64_t n_clusters, int32_t* out_rowptr, int32_t* out_nbr, float* out_w,
                     int64_t* nnz_host, void* ws, size_t ws_bytes, void* stream);

/* One iteration of PoolingLayer.forward's loop after the matching (net_util.py:100-140) in a single call:
 * geobi_relabel_clusters (SYNCS: *n_clusters_host) -> geobi_group_pairs -> geobi_segment_reduce of the features (op) and,
 * if given, of the positions (mean) -> geobi_pool_edges.  `label` must come from a matching (geobi_graclus).  All outputs
 * are caller-allocated at capacity (n_nodes rows / nnz_cap entries; rows [0, n_clusters) are written), so the kernels that
 * depend on the cluster count are queued from inside this call right after the synchronisation.  Results are identical
 * to the five separate calls. */
GEOBI_API size_t geobi_pool_step_ws_bytes(int64_t n_nodes, int64_t nnz_cap);
GEOBI_API int geobi_pool_step(const int32_t* rowptr, const int32_t* nbr, const float* w, int64_t n_nodes, int64_t nnz_cap,
                    const int32_t* label, const float* x, int64_t ldx, int channels, int op, const float* pos,
                    int64_t ldp, int pos_channels, int32_t* cluster, int32_t* mrowptr, int32_t* members,
                    float* x_out, int64_t ldxo, float* pos_out, int64_t ldpo, int32_t* out_rowptr,
                    int32_t* out_nbr, float* out_w, int64_t* n_clusters_host, void* ws, size_t ws_bytes,
                    void* stream);

/* ------------------------------------------------------------------ segment / gather */

/* out[s, :] = reduce_{k in rowptr[s]..rowptr[s+1]} x[idx[k], :]   (op 0 = mean with count
 * clamped to 1, 1 = max with empty -> 0, 2 = sum).  One kernel for: cluster max/mean pool
 * (net_util.py:131-134), pool_pos, vertex->facet corner mean (network.py:335),
 * facet->vertex incident mean (DualFusionLayer.fusion, net_util.py:274-278).
 * rowptr == NULL means fixed-size segments of `fixed` entries (idx is [n_seg, fixed]). */
GEOBI_API int geobi_segment_reduce(const float* x, int64_t ldx, int channels, const int32_t* rowptr,
                         const int32_t* idx, int fixed, int64_t n_seg, int op, float* out, int64_t ldo,
                         void* stream);

/* out[i, :] = x[idx[i], :]  — PoolingLayer.unpooling (net_util.py:242-245). */
GEOBI_API int geobi_gather_rows(const float* x, int64_t ldx, int channels, const int32_t* idx, int64_t n_out,
                      float* out, int64_t ldo, void* stream);

/* Squared feature distance per CSR entry d2[e] = |x_i - x_nbr(e)|^2 and the live Graclus
 * weight of PoolingLayer._get_edge_weight (net_util.py:160-240):
 *   mode 0: w_out = d2;  1: exp(d2/-p);  2: w*exp(d2/-p);  10: w + exp(d2/-2)  (type 10, :226-230) */
GEOBI_API int geobi_edge_weight_feat(const float* x, int64_t ldx, int channels, const int32_t* rowptr,
                           const int32_t* nbr, int64_t n_nodes, const float* w_in, int mode, float param,
                           float* w_out, void* stream);

/* data_util.calc_weight (data_util.py:383-398) over an int64 edge list (self loops count in
 * the mean length, as upstream): w = clamp(n_i.n_j, 1e-3) * exp(l2 / (-2*mean(sqrt(l2)) + 1e-12)). */
GEOBI_API size_t geobi_calc_weight_ws_bytes(int64_t n_edges);
GEOBI_API int geobi_calc_weight(const float* pos, const float* nrm, const int64_t* row, const int64_t* col,
                      int64_t n_edges, float* w_out, void* ws, size_t ws_bytes, void* stream);
/* The same weights for a loop-free int32 CSR (rows = sources), written in CSR entry order (what geobi_graclus reads).  The mean edge
 * length is taken over the reference's list: the CSR entries plus n_loops zero-length self loops (dataset.py:211 appends one per
 * vertex, build_facet_graph keeps one per face).  Workspace: geobi_calc_weight_ws_bytes(n_nodes). */
/* Mean length of the entries of a symmetric loop-free CSR = mean undirected edge length (the 1 / scale of dataset.py:151-152),
 * written as one float on the device: no host round trip.  Workspace: geobi_calc_weight_ws_bytes(n_nodes). */
GEOBI_API int geobi_mean_edge_length_csr(const float* pos, const int32_t* rowptr, const int32_t* nbr, int64_t n_nodes,
                               float* mean_out, void* ws, size_t ws_bytes, void* stream);
GEOBI_API int geobi_calc_weight_csr(const float* pos, const float* nrm, const int32_t* rowptr, const int32_t* nbr, int64_t n_nodes,
                          int64_t n_loops, float* w_out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ FeaSt convolution */

/* torch_geometric.nn.FeaStConv(C_in, C_out, heads=9) forward (call sites network.py:271-299):
 *   out_i = act( 1/(deg_i+1) * sum_{j in N(i) + {i}} sum_h softmax_h(U(x_j-x_i)+c) * (W_h x_j) + b )
 * on a CSR by TARGET node without self loops (the self loop is implicit, as remove_self_loops +
 * add_self_loops make it upstream).  W = lin.weight [9*C_out, C_in] (row h*C_out+o),
 * U = u.weight [9, C_in].  act_slope: 1.0 = none, 0.2 = the leaky_relu after most convs.
 * Evaluation is aggregate-first: Z[i,h,:] = sum_j q_ijh x_j, out = W_flat . Z — no per-edge
 * tensor is ever written to HBM; for C_in=64 -> C_out=32 with BF16X3 the whole layer is ONE persistent kernel
 * (feast_fused.cu) and Z never leaves shared memory.  Supported C_in: 1..128, C_out: multiple of 4 up to 128.
 * row_map (optional, int32 [n_nodes]) fuses PoolingLayer.unpooling (net_util.py:242-245, network.py:289,295) into the conv:
 * node v's features are row row_map[v] of x, which then has n_src (coarse) rows — the [n_nodes, C_in] unpooled copy is never
 * built and U.x is evaluated once per coarse row.  Needs C_in in {32,64,128}, 16-byte aligned rows.  NULL: x has n_nodes rows.
 * precision: GEOBI_PREC_FP32 (all fp32 CUDA cores; parity 1e-5), GEOBI_PREC_BF16 (projection on tcgen05
 * tensor cores, bf16 operands / fp32 accumulate, one pass) or GEOBI_PREC_BF16X3 (same, operands split
 * hi + lo, three passes: fp32-grade results from the tensor cores). */
#define GEOBI_PREC_FP32 0
#define GEOBI_PREC_BF16 1
#define GEOBI_PREC_BF16X3 2   /* tcgen05 with split operands x = hi + lo (both bf16), 3 passes: ~1e-6, fp32-grade */
/* OR-ed into `precision`: the workspace still holds P = X.U^T and the split weights of the previous call on the SAME
 * x / U / W (fused 64->32 kernel only) — skips the two small preparation kernels; bench.py uses it to time the fused
 * kernel alone. */
#define GEOBI_FEAST_REUSE_WS 0x100
GEOBI_API size_t geobi_feast_fwd_ws_bytes(int64_t n_nodes, int c_in, int c_out, int precision);
GEOBI_API int geobi_feast_fwd(const float* x, int64_t ldx, int64_t n_nodes, int c_in, const int32_t* rowptr,
                    const int32_t* nbr, const int32_t* row_map, int64_t n_src, const float* W, const float* U,
                    const float* c, const float* bias, int c_out, float act_slope, float* out, int64_t ldo,
                    int precision, void* ws, size_t ws_bytes, void* stream);

/* ---- training step (train_dual.py:199-218): pieces of the FeaSt / pooling backward that are not plain GEMMs ---- */

/* Forward intermediates for the backward pass: P = X.U^T ([N,9], computed in fp64 and stored as a double-float pair
 * {hi = (float)P, lo = (float)(P - hi)} packed in 8 bytes - opaque to the caller, consumed by geobi_feast_bwd_edges) and the
 * aggregate Z (fp32 [N, 9*C_in],
 * Z[i, h*C_in + c] = mean_{j in N(i)+{i}} q_ijh x_j[c]).  out = act(Z . W_flat^T + b) is then a dense product. */
GEOBI_API int geobi_feast_aggregate(const float* x, int64_t ldx, int64_t n_nodes, int c_in, const int32_t* rowptr,
                                    const int32_t* nbr, const float* U, const float* c, double* P, float* Z, void* stream);

/* Edge part of the FeaSt backward: given dZ [N, 9*C_in] accumulates (atomically, into zero-initialised buffers)
 * dx [N, C_in] (gradient reaching x through the gathered rows), dP [N, 9] (gradient of the head logits' projections,
 * fp32) and dc [9].  The caller finishes with dX += dP.U, dU = dP^T.X (PyG FeaStConv's autograd, network.py:271-299).
 * dx may be NULL when the layer's input needs no gradient (c_in <= 16, 32, 64 or 128). */
GEOBI_API int geobi_feast_bwd_edges(const float* x, int64_t ldx, int64_t n_nodes, int c_in, const int32_t* rowptr,
                                    const int32_t* nbr, const double* P, const float* c, const float* dZ, float* dx,
                                    int64_t lddx, float* dP, float* dc, void* stream);

/* Backward of scatter(x, cluster, reduce='max') (net_util.py:134): routes g[s,:] to the arg-max member of segment s
 * (first member on ties).  dx must be zero-initialised. */
GEOBI_API int geobi_segment_max_bwd(const float* x, int64_t ldx, int channels, const int32_t* rowptr, const int32_t* idx,
                                    int64_t n_seg, const float* g, int64_t ldg, float* dx, int64_t lddx, void* stream);

/* Whole backward of one FeaStConv layer (replaces autograd through torch_geometric's FeaStConv in the training step,
 * train_dual.py:199-218; layer built at network.py:258-268, called at :271-299) in ONE call:
 *   g = g_out * act'(out)  (leaky_relu slope act_slope; out = the layer's forward output, may be NULL when act_slope == 1),
 *   dbias = sum_n g,  dZ = g . W_flat,  dW = g^T . Z  (Z = the forward's aggregate, recomputed),  the edge part
 *   (geobi_feast_bwd_edges),  dx += dP . U,  dU = dP^T . x.
 * The products with a long reduction or a wide output (dZ, dW) run on tcgen05 with operands split hi + lo (three bf16 passes, fp32
 * accumulation in tensor memory: fp32-grade results); dW's reduction over the nodes is a split-K kernel whose operands are TMA boxes
 * of the node-major planes used as MN-major UMMA tiles, with a fixed-order second pass (deterministic).
 * Outputs are OVERWRITTEN: dx [N, c_in] (row stride lddx; NULL when the layer's input needs no gradient), dW [9*c_out, c_in]
 * (lin.weight's layout), dU [9, c_in], dc [9], dbias [c_out].  Rows of g_out / out must be 16-byte aligned; ws 128-byte aligned. */
GEOBI_API size_t geobi_feast_bwd_ws_bytes(int64_t n_nodes, int c_in, int c_out);
GEOBI_API int geobi_feast_bwd(const float* x, int64_t ldx, int64_t n_nodes, int c_in, const int32_t* rowptr, const int32_t* nbr,
                              const float* W, const float* U, const float* c, int c_out, float act_slope, const float* out,
                              int64_t ldo, const float* g_out, int64_t ldg, float* dx, int64_t lddx, float* dW, float* dU,
                              float* dc, float* dbias, void* ws, size_t ws_bytes, void* stream);

/* Backward of one FC head (fc_v1/fc_v2, fc_f1/fc_f2: network.py:324-325,340-341; replaces autograd through the two F.linear calls in
 * the training step): y = W2 . a + b2, a = leaky_relu(h, act_slope), h = W1 . f + b1; given dy = dL/dy [N, c_out] (the gradient
 * BEFORE the head's epilogue - residual / force_depth / normalize stay with the caller):
 *   dW2 = dy^T . a,  db2 = sum dy,  dh = (dy . W2) * act'(h),  dW1 = dh^T . f,  db1 = sum dh,  df = dh . W1.
 * h is recomputed on tcgen05 (split bf16 operands, b1 folded in as an extra operand column); a and dh are written once as bf16
 * hi | lo planes [N, hidden] in the workspace and consumed by the split-K kernel of geobi_feast_bwd (dW2, [dW1; db1]: reduction over
 * the rows) and by the TMA GEMM (df).  c_in must be 32, hidden a multiple of 256 (<= 4096), c_out <= 4.  Outputs are overwritten;
 * df [N, 32] may be NULL.  ws 128-byte aligned. */
GEOBI_API size_t geobi_mlp_head_bwd_ws_bytes(int64_t n, int c_in, int hidden);
GEOBI_API int geobi_mlp_head_bwd(const float* f, int64_t ldf, int64_t n, int c_in, const float* W1, const float* b1, int hidden,
                                 const float* W2, int c_out, float act_slope, const float* dy, int64_t lddy, float* df, int64_t lddf,
                                 float* dW1, float* db1, float* dW2, float* db2, void* ws, size_t ws_bytes, void* stream);

/* Per-node linear layer on the tcgen05 tensor cores: out = act(A . W^T + bias), A fp32 [M,K] rounded to bf16
 * (split hi + lo for BF16X3), W fp32 [N,K] (nn.Linear layout), fp32 accumulation in TMEM.  Any K (zero padded to a
 * multiple of 64), N in {32,64,128,256}, out rows 16-byte aligned.  Replaces F.linear / cuBLAS sgemm for the
 * projections (FeaSt `lin`, network.py:258-268; DualFusionLayer linears, net_util.py:252-256).
 * ws: bf16 planes of W and A. */
GEOBI_API size_t geobi_linear_tc_ws_bytes(int64_t m, int k, int n);
GEOBI_API int geobi_linear_tc(const float* A, int64_t lda, int64_t M, int K, const float* W, int N, const float* bias,
                              float act_slope, float* out, int64_t ldo, int precision, void* ws, size_t ws_bytes,
                              void* stream);

/* The two linear heads of DualGNN (network.py:324-325,340-341) fused so the [N,1024] hidden never
 * reaches HBM:  y = W2 . leaky_relu(W1 . f + b1, 0.2) + b2, then epilogue
 *   0: none | 1: y += res (network.py:332) | 2: y = y * res2 (force_depth, :327; c_out==1
 *   broadcasts) then += res | 3: y = normalize(y) (network.py:343, eps 1e-12).
 * f [N, c_in] (c_in <= 64), W1 [hidden, c_in], W2 [c_out, hidden], c_out <= 4.
 * ws (tensor-core precisions): the split-bf16 copy of W1 that every CTA streams with TMA; 128-byte aligned. */
GEOBI_API size_t geobi_fc_head_ws_bytes(int hidden);
GEOBI_API int geobi_fc_head_fwd(const float* f, int64_t ldf, int64_t n, int c_in, const float* W1, const float* b1,
                      int hidden, const float* W2, const float* b2, int c_out, int epilogue,
                      const float* res, int64_t ldres, const float* res2, int64_t ldres2, float* out,
                      int64_t ldo, int precision, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ dual-domain transfer */

/* data_util.computer_face_normal (data_util.py:182-198): normalize(cross(v1-v0, v2-v0)), eps 1e-12. */
GEOBI_API int geobi_face_normal(const float* points, int64_t ldp, const int64_t* fv, int64_t n_faces, float* out,
                      int64_t ldo, void* stream);

/* Vertex->facet transfer of DualGNN.forward (network.py:335-337) in one pass:
 * out[f] = [ xf[f, 0:cf] | mean of the 3 corner predictions | face normal of the predictions ]. */
GEOBI_API int geobi_v2f_transfer(const float* feat_v, int64_t ldv, const int64_t* fv, const float* xf, int64_t ldxf,
                       int cf, int64_t n_faces, float* out, int64_t ldo, void* stream);

/* Backward of geobi_v2f_transfer for the training step (the autograd of network.py:335-337: corner mean + normalize(cross)):
 * g_out [F, >= 6] = gradients of (corner mean | face normal), i.e. columns cf .. cf+5 of the transfer's output (pass the pointer to
 * column cf); d_feat_v [V, >= 3] is ACCUMULATED into (zero it first).  Covers SURVEY.md 8(b)'s geobi_face_normal_bwd. */
GEOBI_API int geobi_v2f_transfer_bwd(const float* feat_v, int64_t ldv, const int64_t* fv, const float* g_out, int64_t ldg,
                           int64_t n_faces, float* d_feat_v, int64_t lddv, void* stream);

/* ---------------------------------------------------------------- BFS face-patch splitter */
/* The reference's mesh split (dataset.py:156-193 over data_util.mesh_get_neighbor_np, data_util.py:55-84; pure-Python loops
 * upstream) on the device, with the reference's discovery order: seed = first arg-max of the squared face-centre distance to
 * `centroid` over the faces no patch has covered; ring-by-ring growth through faces sharing a vertex (faces of the ring in list
 * order, corners in order, incident faces in vf-row order), cut at exactly `neighbor_count` faces.  Identical patches to the host
 * splitter (geobi_host_grow_patch).  The workspace carries the state between calls:
 *   geobi_bfs_begin   distances (numpy's fp32 operation order), stamps; *seed_host = the first seed.  centroid_host: 3 floats on the HOST.
 *   geobi_bfs_grow    grows the patch of `seed` into out_faces (device int32, at least min(neighbor_count, n_faces) entries),
 *                     removes it from the uncovered set; *n_out_host = its size, *next_seed_host = the next seed or -1 when every
 *                     face is covered.  Synchronises the stream (one state read-back per 32 rings).
 * fv int64 [F,3], vf int64 [V,k] padded with -1 (k <= 32). */
GEOBI_API size_t geobi_bfs_ws_bytes(int64_t n_faces);
GEOBI_API int geobi_bfs_begin(const float* points, const int64_t* fv, int64_t n_faces, const float* centroid_host, int64_t* seed_host,
                              void* ws, size_t ws_bytes, void* stream);
GEOBI_API int geobi_bfs_grow(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t k, int64_t seed, int64_t neighbor_count,
                             int32_t* out_faces, int64_t* n_out_host, int64_t* next_seed_host, void* ws, size_t ws_bytes, void* stream);

/* Backward of geobi_face_normal (data_util.computer_face_normal, data_util.py:182-198: normalize(cross(p1-p0, p2-p0))):
 * d_points (zero-initialised by the caller, [V, >=3]) += the gradient reaching the three corners of every face from
 * g_normal [F, 3].  Same kernel as geobi_v2f_transfer_bwd without the corner-mean part. */
GEOBI_API int geobi_face_normal_bwd(const float* points, int64_t ldp, const int64_t* fv, const float* g_normal, int64_t ldg,
                                    int64_t n_faces, float* d_points, int64_t lddp, void* stream);

/* data_util.update_position2 (data_util.py:529-556; test_dual.py:72 runs 60 iterations):
 * n_iter Jacobi sweeps p_v += mean_{f in vf[v]} n_f (n_f . (c_f - p_v)), optional projection on
 * depth_direction.  vf [V,K] int64 padded -1.  out may not alias points. */
GEOBI_API size_t geobi_update_position_ws_bytes(int64_t n_verts, int64_t n_faces);
GEOBI_API int geobi_update_position(const float* points, const int64_t* fv, const int64_t* vf, int64_t k,
                          const float* face_normals, int n_iter, const float* depth, int64_t n_verts,
                          int64_t n_faces, float* out, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOBI_H */
