/* geobi_host.h - C ABI of geobi_gnn_b200/libgeobi_host.so: the HOST-side (CPU, C++) input preparation either side of the GPU
 * hot path.  No CUDA, no torch types: plain pointers and sizes.  Every entry point cites the
 * reference code it replaces; all of them are bit-exact against that code's numpy / Python arithmetic (tests/test_abi.py,
 * tests/test_gpu_patches.py::test_splitter_*, tests/test_reference_golden.py::test_patch_walk_is_the_references).
 * The GPU kernels' ABI is include/geobi.h.  Bound with ctypes in geobi_gnn_b200/patches.py, dataset.py and meshio.py.
 */
#ifndef GEOBI_HOST_H
#define GEOBI_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- BFS patch splitter: data_util.mesh_get_neighbor_np (/root/reference/code/data_util.py:55-84) ----------------------
 * Ring-by-ring growth from face `seed` over fv [n_faces,3] / vf [n_vertices,k] (-1 padded), faces in the reference's
 * discovery order, cut at exactly `neighbor_count` faces or after `ring_count` rings.  fstamp [n_faces] / vstamp [n_vertices]
 * are caller-owned stamp arrays (zero-filled once; pass a fresh non-zero `epoch` per call).  Returns the number of faces
 * written to `out`. */
int64_t geobi_host_grow_patch(const int64_t* fv, const int64_t* vf, int64_t n_faces, int64_t k, int64_t seed, int64_t neighbor_count,
                              int64_t ring_count, uint32_t* fstamp, uint32_t* vstamp, uint32_t epoch, int64_t* out);

/* data_util.get_submesh (data_util.py:318-336): first-appearance re-indexing of the selected faces.  slot [n_vertices] is
 * scratch filled with -1 (restored on return).  v_idx receives the original vertex ids, faces_out [n_select,3] the local
 * ones.  Returns the number of vertices. */
int64_t geobi_host_submesh(const int64_t* fv, const int64_t* select, int64_t n_select, int64_t* slot, int64_t* v_idx, int64_t* faces_out);

/* Seed rule of the splitter (dataset.py:163-166,186-192): squared distance of every face centre to the centroid in numpy's fp32
 * operation order ... */
void geobi_host_face_d2(const float* pts, const int64_t* fv, int64_t n_faces, const float* centroid, float* out, int n_threads);

/* ... and the book-keeping between two patches: the faces of `sel` leave the uncovered set (d2_left -> -inf, *n_left
 * decremented by the number that were still uncovered); returns the next seed = np.argmax of what is left. */
int64_t geobi_host_cover_next_seed(float* d2_left, int64_t n_faces, const int64_t* sel, int64_t n_sel, int64_t* n_left, int n_threads);

/* data_util.center_and_scale, numpy branch (data_util.py:201-230; dataset.py:140): per-edge lengths of the centred mesh in
 * numpy's fp32 operation order; the caller takes numpy's mean, so the scale is the reference's to the bit. */
void geobi_host_edge_lengths(const float* pts, const float* centroid, const int64_t* ev, int64_t n_edges, float* out, int n_threads);

/* ---- Wavefront .obj: om.read_trimesh / om.write_mesh (dataset.py:134-135; test_dual.py:29,73) --------------------------
 * Records used by the path: `v x y z`, `f a b c ...` (fan-triangulated; `a/b/c` tokens; negative = relative indices).
 * Pass 1: counts[0] = vertices, counts[1] = triangles.  Pass 2 (same buffer, same n_threads) fills points [V,3] float64 and
 * faces [T,3] int64 and returns -1, or the byte offset of the first malformed record. */
int geobi_host_obj_count(const char* buf, int64_t n_bytes, int n_threads, int64_t* counts);
int64_t geobi_host_obj_parse(const char* buf, int64_t n_bytes, int n_threads, double* points, int64_t* faces);

/* `# V vertices, F faces`, `v %.6g %.6g %.6g`, 1-based `f a b c`.  Returns 0, or -1 when the file cannot be written. */
int geobi_host_obj_write(const char* path, const double* points, int64_t n_vertices, const int64_t* faces, int64_t n_faces, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* GEOBI_HOST_H */
