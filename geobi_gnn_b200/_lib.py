"""ctypes binding of libgeobi.so (C ABI declared in include/geobi.h).

The library is built in-tree by ``geobi_gnn_b200/csrc/Makefile`` (nvcc, sm_100a).
There is no fallback: if the shared object is missing, or a call is made
without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GEOBI_LIB_PATH") or os.path.join(_HERE, "libgeobi.so")   # override: A/B kernel builds (profiles/)

_i64, _i32, _f32, _sz, _p = C.c_int64, C.c_int, C.c_float, C.c_size_t, C.c_void_p

# name -> (restype, argtypes); mirrors include/geobi.h one to one
SIGNATURES = {
    "geobi_last_error": (C.c_char_p, []),
    "geobi_version": (_i32, []),
    "geobi_device_info": (_i32, [_p, _p, _p]),
    "geobi_scan_ws_bytes": (_sz, [_i64]),
    "geobi_exclusive_scan_i32": (_i32, [_p, _p, _i64, _p, _sz, _p]),
    "geobi_csr_from_coo_ws_bytes": (_sz, [_i64, _i64, _i32]),
    "geobi_csr_from_coo": (_i32, [_p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "geobi_pool_step_ws_bytes": (_sz, [_i64, _i64]),
    "geobi_pool_step": (_i32, [_p, _p, _p, _i64, _i64, _p, _p, _i64, _i32, _i32, _p, _i64, _i32, _p, _p, _p, _p, _i64, _p, _i64, _p, _p, _p,
                               _p, _p, _sz, _p]),
    "geobi_csr_from_sorted_coo_ws_bytes": (_sz, [_i64]),
    "geobi_csr_from_sorted_coo": (_i32, [_p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "geobi_remove_self_loops_ws_bytes": (_sz, [_i64]),
    "geobi_remove_self_loops": (_i32, [_p, _p, _p, _i64, _i64, _p, _p, _p, _sz, _p]),
    "geobi_csr_to_coo": (_i32, [_p, _p, _i64, _i64, _p, _p]),
    "geobi_build_facet_graph_ws_bytes": (_sz, [_i64, _i64]),
    "geobi_build_facet_graph": (_i32, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "geobi_build_facet_graph_sorted_ws_bytes": (_sz, [_i64]),
    "geobi_build_facet_graph_sorted": (_i32, [_p, _p, _i64, _i64, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "geobi_graclus_ws_bytes": (_sz, [_i64]),
    "geobi_graclus": (_i32, [_p, _p, _p, _p, _i64, _p, _p, _p, _sz, _p]),
    "geobi_relabel_ws_bytes": (_sz, [_i64]),
    "geobi_relabel_clusters": (_i32, [_p, _i64, _p, _p, _p, _sz, _p]),
    "geobi_group_by_ws_bytes": (_sz, [_i64, _i64]),
    "geobi_group_by": (_i32, [_p, _i64, _i64, _p, _p, _p, _sz, _p]),
    "geobi_group_pairs_ws_bytes": (_sz, [_i64]),
    "geobi_group_pairs": (_i32, [_p, _p, _i64, _i64, _p, _p, _p, _sz, _p]),
    "geobi_pool_edges_ws_bytes": (_sz, [_i64, _i64]),
    "geobi_pool_edges": (_i32, [_p, _p, _p, _i64, _i64, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "geobi_segment_reduce": (_i32, [_p, _i64, _i32, _p, _p, _i32, _i64, _i32, _p, _i64, _p]),
    "geobi_gather_rows": (_i32, [_p, _i64, _i32, _p, _i64, _p, _i64, _p]),
    "geobi_edge_weight_feat": (_i32, [_p, _i64, _i32, _p, _p, _i64, _p, _i32, _f32, _p, _p]),
    "geobi_calc_weight_ws_bytes": (_sz, [_i64]),
    "geobi_calc_weight": (_i32, [_p, _p, _p, _p, _i64, _p, _p, _sz, _p]),
    "geobi_mean_edge_length_csr": (_i32, [_p, _p, _p, _i64, _p, _p, _sz, _p]),
    "geobi_mesh_vertex_csr_ws_bytes": (_sz, [_i64]),
    "geobi_mesh_vertex_csr": (_i32, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "geobi_pad_rows": (_i32, [_p, _p, _i64, _i64, _i32, _p, _p]),
    "geobi_calc_weight_csr": (_i32, [_p, _p, _p, _p, _i64, _i64, _p, _p, _sz, _p]),
    "geobi_feast_fwd_ws_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "geobi_feast_fwd": (_i32, [_p, _i64, _i64, _i32, _p, _p, _p, _i64, _p, _p, _p, _p, _i32, _f32, _p, _i64, _i32, _p, _sz, _p]),
    "geobi_feast_aggregate": (_i32, [_p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "geobi_feast_bwd_edges": (_i32, [_p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "geobi_feast_bwd_ws_bytes": (_sz, [_i64, _i32, _i32]),
    "geobi_feast_bwd": (_i32, [_p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _i32, _f32, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "geobi_mlp_head_bwd_ws_bytes": (_sz, [_i64, _i32, _i32]),
    "geobi_mlp_head_bwd": (_i32, [_p, _i64, _i64, _i32, _p, _p, _i32, _p, _i32, _f32, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "geobi_face_normal_bwd": (_i32, [_p, _i64, _p, _p, _i64, _i64, _p, _i64, _p]),
    "geobi_bfs_ws_bytes": (_sz, [_i64]),
    "geobi_bfs_begin": (_i32, [_p, _p, _i64, _p, _p, _p, _sz, _p]),
    "geobi_bfs_grow": (_i32, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "geobi_segment_max_bwd": (_i32, [_p, _i64, _i32, _p, _p, _i64, _p, _i64, _p, _i64, _p]),
    "geobi_linear_tc_ws_bytes": (_sz, [_i64, _i32, _i32]),
    "geobi_linear_tc": (_i32, [_p, _i64, _i64, _i32, _p, _i32, _p, _f32, _p, _i64, _i32, _p, _sz, _p]),
    "geobi_fc_head_ws_bytes": (_sz, [_i32]),
    "geobi_fc_head_fwd": (_i32, [_p, _i64, _i64, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _i32, _p, _sz, _p]),
    "geobi_face_normal": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p]),
    "geobi_v2f_transfer": (_i32, [_p, _i64, _p, _p, _i64, _i32, _i64, _p, _i64, _p]),
    "geobi_v2f_transfer_bwd": (_i32, [_p, _i64, _p, _p, _i64, _i64, _p, _i64, _p]),
    "geobi_update_position_ws_bytes": (_sz, [_i64, _i64]),
    "geobi_update_position": (_i32, [_p, _p, _p, _i64, _p, _i32, _p, _i64, _i64, _p, _p, _sz, _p]),
}

_lib = None


class GeobiError(RuntimeError):
    pass


def load():
    """Load libgeobi.so (once) and attach the signatures.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GeobiError(
                f"{LIB_PATH} is missing: build it with `make -C geobi_gnn_b200/csrc` "
                "(or __graft_entry__.build()).  geobi_gnn_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().geobi_last_error().decode("utf-8", "replace")
        raise GeobiError(f"{what or 'libgeobi'} failed (code {rc}): {msg}")
