"""Tensor half of the reference's data_util (/root/reference/code/data_util.py:182-230,383-556)
on libgeobi kernels.  Same function names and argument meaning; inputs must be CUDA tensors.

Outputs that are graphs come back as the reference's int64 ``edge_index`` [2,E] (bit-identical
to torch_sparse.coalesce's sorted layout) with the int32 CSR they were built from attached, so
the layers never rebuild it.
"""
from __future__ import annotations

import torch

from . import nn as gnn
from . import ops


def computer_face_normal(points, fv_indices):
    """data_util.py:182-198."""
    return ops.face_normal(points, fv_indices)


def center_and_scale(points, ev_indices, s_type=0):
    """data_util.py:201-230 (one-off preprocessing reduction; plain tensor expressions)."""
    centroid = points.mean(0, keepdim=True)
    p = points - centroid
    if s_type == 0:
        e = p[ev_indices]
        scale = ((e[:, 0] - e[:, 1]) ** 2).sum(1).sqrt().mean()
    elif s_type == 1:
        scale = ((p.max(0)[0] - p.min(0)[0]) ** 2).sum().sqrt()
    elif s_type == 2:
        scale = p.abs().max()
    elif s_type == 3:
        scale = (p ** 2).sum(1).max().sqrt()
    else:
        raise ValueError(s_type)
    scale = 1 / scale
    return p * scale, centroid, scale


def calc_weight(node_pos, node_normal, edge_index):
    """data_util.py:383-398 — bilateral Graclus weight."""
    return ops.calc_weight(node_pos, node_normal, edge_index)


def build_facet_graph(fv_indices, vf_indices, vf_sorted=False):
    """data_util.py:436-456 — sorted [2,E] with self entries; CSR (self entries dropped lazily by the
    conv / matcher builders) is attached for reuse.  `vf_sorted` (not upstream): the caller guarantees ascending vf rows."""
    g = ops.build_facet_graph_csr(fv_indices, vf_indices, vf_sorted)
    return g.edge_index()


def build_vertex_graph(ev_indices, vv_indices):
    """data_util.py:407-433 — 2-ring vertex graph (dead code upstream; kept for API parity)."""
    ev, vv = ev_indices.long(), vv_indices.long()
    n = vv.shape[0]
    row = torch.cat([ev[:, 0], ev[:, 1]])
    col = torch.cat([ev[:, 1], ev[:, 0]])
    j = vv[col]
    i = row.unsqueeze(1).expand_as(j)
    keep = j > -1
    pairs = torch.stack([i[keep], j[keep]])
    return ops.csr_from_coo(pairs, n, None, ops.COO_SORT_NBR | ops.COO_DEDUP).edge_index()


def build_edge_vf(vf_indices):
    """data_util.py:459-475 (pure index layout)."""
    vf = vf_indices.long()
    v, k = vf.shape
    i = torch.arange(v, device=vf.device).repeat_interleave(k)
    j = vf.reshape(-1)
    keep = j > -1
    return torch.stack([i[keep], j[keep]])


def build_edge_fv(fv_indices):
    """data_util.py:478-489 (pure index layout)."""
    f = fv_indices.shape[0]
    return torch.stack([torch.arange(f, device=fv_indices.device).repeat_interleave(3), fv_indices.reshape(-1).long()])


def to_undirected_with_self_loops(ev_t, num_nodes=None):
    """dataset.py:211-213: to_undirected (sorted, deduplicated) then add_self_loops (appended last)."""
    n = int(ev_t.max()) + 1 if num_nodes is None else num_nodes
    g = ops.csr_from_coo(ev_t, n, None, ops.COO_SYMMETRIZE | ops.COO_SORT_NBR | ops.COO_DEDUP)
    g.symmetric = True
    return with_self_loops_appended(g)


def with_self_loops_appended(g):
    """[sorted undirected list | (i,i) for every node] from the loop-free symmetric CSR of a vertex graph."""
    n = g.n
    loops = torch.arange(n, device=g.rowptr.device, dtype=torch.int64).unsqueeze(0).repeat(2, 1)
    ei = torch.cat([g.edge_index(), loops], 1)
    gnn.tag_of(ei)["tgt"] = g          # mesh edges carry no loops: the CSR holds exactly the non-loop part
    return ei


def update_position(points, fv_indices, vf_indices, face_normals, n_iter=20, depth_direction=None, lmd=1):
    """data_util.py:492-526 (scatter variant; same arithmetic as update_position2)."""
    return ops.update_position(points, fv_indices, vf_indices, face_normals, n_iter, depth_direction)


def update_position2(points, fv_indices, vf_indices, face_normals, n_iter=20, depth_direction=None):
    """data_util.py:529-556; test_dual.py:72 runs it with n_iter=60."""
    return ops.update_position(points, fv_indices, vf_indices, face_normals, n_iter, depth_direction)
