"""CPU oracle: graph + feature assembly of the reference dataset.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/code/dataset.py:196-269 (``process_one_submesh`` and
``post_processing``) and the patch stitch of /root/reference/code/test_dual.py:49-61.
OpenMesh is replaced by any object exposing the same index arrays as numpy
(``points, ev, fv, vf, vv, face_normals, vertex_normals`` — see
geobi_gnn_b200/synth.py:TriMesh).  Pinned against the reference's own
dataset.py executed over the same mesh stand-in (tests/test_reference_golden.py); the OpenMesh arrays are unpinned.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_data_util as data_util
from .pyg import Data, add_self_loops, to_undirected


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def process_one_submesh(mesh_n, name="graph", mesh_o=None):
    """dataset.py:196-243 — (graph_v, graph_f) before post_processing."""
    ev, fv = _t(mesh_n.ev, torch.long), _t(mesh_n.fv, torch.long)
    vf = _t(mesh_n.vf, torch.long)
    edge_dual = data_util.build_edge_fv(fv)
    # vertex graph: undirected 1-ring, self loops appended last
    pos_v = _t(mesh_n.points, torch.float32)
    nrm_v = _t(mesh_n.vertex_normals, torch.float32)
    ei_v, _ = add_self_loops(to_undirected(ev.t()))
    w_v = data_util.calc_weight(pos_v, nrm_v, ei_v)
    graph_v = Data(name=f"{name}-v", pos=pos_v, normal=nrm_v, edge_index=ei_v, edge_weight=w_v.float(),
                   depth_direction=F.normalize(pos_v, dim=1), edge_dual=edge_dual[1])
    # facet graph: 1-ring through shared vertices, self included, sorted
    pos_f = pos_v[fv].mean(1)
    nrm_f = _t(mesh_n.face_normals, torch.float32).reshape(-1, 3)
    ei_f = data_util.build_facet_graph(fv, vf)
    w_f = data_util.calc_weight(pos_f, nrm_f, ei_f)
    graph_f = Data(name=f"{name}-f", pos=pos_f, normal=nrm_f, edge_index=ei_f, edge_weight=w_f.float(),
                   fv_indices=fv, edge_dual=edge_dual[0])
    if mesh_o is not None:
        graph_v.y = _t(mesh_o.points, torch.float32)
        graph_f.y = _t(mesh_o.face_normals, torch.float32)
    return graph_v, graph_f


def attach_normalisation(dual_data, points_noisy, ev):
    """dataset.py:140,151-152 — centroid / scale of the WHOLE noisy mesh (numpy path of
    center_and_scale), stored on graph_v."""
    p = np.asarray(points_noisy, dtype=np.float32)
    centroid = p.mean(0, keepdims=True)
    q = p - centroid
    e = q[ev]
    scale = 1 / (((e[:, 0] - e[:, 1]) ** 2).sum(1) ** 0.5).mean()
    dual_data[0].centroid = torch.from_numpy(centroid).float()
    dual_data[0].scale = scale
    return dual_data


def post_processing(dual_data, data_type="Synthetic"):
    """dataset.py:245-269 — x = cat((pos - centroid) * scale, normal); strip the rest."""
    data_v, data_f = dual_data
    data_f.x = torch.cat(((data_f.pos - data_v.centroid) * data_v.scale, data_f.normal), 1)
    data_f.normal = data_f.edge_dual = None
    data_f.pos = None
    data_v.x = torch.cat(((data_v.pos - data_v.centroid) * data_v.scale, data_v.normal), 1)
    data_v.y = None if data_v.y is None else (data_v.y - data_v.centroid) * data_v.scale
    data_v.normal = data_v.centroid = data_v.scale = data_v.edge_dual = None
    data_v.pos = None
    if data_type not in ("Kinect_v1", "Kinect_v2"):
        data_v.depth_direction = None
    return data_v, data_f


def build_dual_data(mesh_n, mesh_o=None, data_type="Synthetic", name="graph"):
    """process_one_data's single-patch branch (dataset.py:144-153) + post_processing."""
    dd = process_one_submesh(mesh_n, name, mesh_o)
    attach_normalisation(dd, mesh_n.points, mesh_n.ev)
    return post_processing(dd, data_type)


def stitch_patches(n_vertices, n_faces, results):
    """test_dual.py:49-61 — results: iterable of (vert_p, norm_p, V_idx, F_idx)."""
    cnt = torch.zeros(n_vertices, 1)
    vp = torch.zeros(n_vertices, 3)
    np_ = torch.zeros(n_faces, 3)
    for vert_p, norm_p, v_idx, f_idx in results:
        cnt[v_idx] += 1
        vp[v_idx] += vert_p
        np_[f_idx] += norm_p
    return vp / cnt, F.normalize(np_, dim=1)
