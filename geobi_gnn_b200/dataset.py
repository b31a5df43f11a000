"""Graph + feature assembly with the reference's surface (/root/reference/code/dataset.py:196-269),
on the device.

``process_one_submesh`` / ``post_processing`` keep their names, argument meaning and the layout
of the returned ``(graph_v, graph_f)`` tuple.  OpenMesh (not installable here) is replaced by
any object exposing its index arrays as numpy: ``points, ev, fv, vf, vv, face_normals,
vertex_normals`` (geobi_gnn_b200/synth.py:TriMesh).  File I/O, the .pt
cache and the Kinect readers are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import data_util
from .data import Data


def _t(a, dtype, device):
    if torch.is_tensor(a):                      # topology.DeviceTriMesh hands over device tensors
        return a.to(device=device, dtype=dtype)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(device)


def process_one_submesh(mesh_n, name="graph", mesh_o=None, device="cuda"):
    """dataset.py:196-243."""
    fv = _t(mesh_n.fv, torch.long, device)
    vf = _t(mesh_n.vf, torch.long, device)
    edge_dual_fv = data_util.build_edge_fv(fv)
    pos_v = _t(mesh_n.points, torch.float32, device)
    normal_v = _t(mesh_n.vertex_normals, torch.float32, device)
    if getattr(mesh_n, "vertex_csr", None) is not None:      # DeviceTriMesh already holds to_undirected(ev) as a CSR
        edge_idx_v = data_util.with_self_loops_appended(mesh_n.vertex_csr)
    else:
        ev = _t(mesh_n.ev, torch.long, device)
        edge_idx_v = data_util.to_undirected_with_self_loops(ev.t().contiguous(), pos_v.size(0))
    edge_wei_v = data_util.calc_weight(pos_v, normal_v, edge_idx_v)
    graph_v = Data(name=f"{name}-v", pos=pos_v, normal=normal_v, edge_index=edge_idx_v, edge_weight=edge_wei_v,
                   depth_direction=F.normalize(pos_v, dim=1), edge_dual=edge_dual_fv[1], coalesced_undirected=True)
    pos_f = pos_v[fv].mean(1)
    normal_f = _t(mesh_n.face_normals, torch.float32, device).reshape(-1, 3)
    edge_idx_f = data_util.build_facet_graph(fv, vf)
    edge_wei_f = data_util.calc_weight(pos_f, normal_f, edge_idx_f)
    graph_f = Data(name=f"{name}-f", pos=pos_f, normal=normal_f, edge_index=edge_idx_f, edge_weight=edge_wei_f,
                   fv_indices=fv, edge_dual=edge_dual_fv[0], coalesced_undirected=True)
    # coalesced_undirected: both builders above emit sorted, duplicate-free, symmetric lists (self loops aside), which
    # lets the network build one CSR per graph without a sort (nn.input_graph); the device verifies the claim
    if mesh_o is not None:
        graph_v.y = _t(mesh_o.points, torch.float32, device)
        graph_f.y = _t(mesh_o.face_normals, torch.float32, device)
    return graph_v, graph_f


def normalisation(points_noisy, ev):
    """(centroid [1,3] fp32, scale) of dataset.py:140,151-152: `q = p - centroid; e = q[ev]; 1 / |e0 - e1|.mean()`.  The per-edge
    lengths come from one threaded C++ pass with numpy's fp32 operation order (bit-equal, tests/test_abi.py) and numpy takes
    the mean, so the numbers are upstream's; the [E,2,3] temporaries of a 10 M-face mesh are not built."""
    import ctypes as C
    import os
    from . import patches
    host = lambda a: a.detach().cpu().numpy() if torch.is_tensor(a) else a      # a DeviceTriMesh hands over device tensors
    p = np.ascontiguousarray(host(points_noisy), dtype=np.float32)
    ev = np.ascontiguousarray(host(ev), dtype=np.int64)
    centroid = np.ascontiguousarray(p.mean(0, keepdims=True))
    length = np.empty(ev.shape[0], dtype=np.float32)
    patches._host().geobi_host_edge_lengths(patches._p(p), patches._p(centroid), patches._p(ev), C.c_int64(ev.shape[0]), patches._p(length),
                                            C.c_int(min(8, os.cpu_count() or 1)))
    return centroid, 1 / length.mean()


def attach_normalisation(dual_data, points_noisy, ev, precomputed=None):
    """dataset.py:140,151-152: centroid / scale of the whole noisy mesh (numpy fp32, as upstream).
    `precomputed` = (centroid tensor, scale) lets a caller that normalises many patches of one mesh pay for it once."""
    if precomputed is not None:
        dual_data[0].centroid, dual_data[0].scale = precomputed
        return dual_data
    centroid, scale = normalisation(points_noisy, ev)
    dual_data[0].centroid = torch.from_numpy(centroid).float().to(dual_data[0].pos.device)
    dual_data[0].scale = float(scale)
    return dual_data


def post_processing(dual_data, data_type="Synthetic", is_plot=False):
    """dataset.py:245-269."""
    data_v, data_f = dual_data
    data_f.x = torch.cat(((data_f.pos - data_v.centroid) * data_v.scale, data_f.normal), 1)
    data_f.normal = data_f.edge_dual = None
    if not is_plot:
        data_f.pos = None
    data_v.x = torch.cat(((data_v.pos - data_v.centroid) * data_v.scale, data_v.normal), 1)
    data_v.y = None if data_v.y is None else (data_v.y - data_v.centroid) * data_v.scale
    data_v.normal = data_v.centroid = data_v.scale = data_v.edge_dual = None
    if not is_plot:
        data_v.pos = None
    else:
        data_v.pos = data_v.y
        data_v.fv_indices = data_f.fv_indices
    if data_type not in ["Kinect_v1", "Kinect_v2"]:
        data_v.depth_direction = None
    return data_v, data_f


def build_dual_data(mesh_n, mesh_o=None, data_type="Synthetic", name="graph", device="cuda"):
    """Single-patch branch of process_one_data (dataset.py:144-153) followed by post_processing."""
    dd = process_one_submesh(mesh_n, name, mesh_o, device)
    attach_normalisation(dd, mesh_n.points, mesh_n.ev)
    return post_processing(dd, data_type)
