"""CPU oracle: tensor half of the reference's data_util.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/code/data_util.py:182-230 (face normal, centre/scale) and
:383-556 (bilateral weight, graph / incidence builders, vertex update) in plain
PyTorch on CPU.  Pinned against the reference's own data_util.py, executed over stand-ins for openmesh /
torch_scatter / torch_sparse / matplotlib (tests/golden/make_reference_golden.py, tests/test_reference_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .pyg import coalesce, scatter


def computer_face_normal(points, fv_indices):
    """data_util.py:182-198 — normalize(cross(v1-v0, v2-v0)), eps 1e-12."""
    tri = points[fv_indices]
    n = torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=1)
    return F.normalize(n, dim=1)


def center_and_scale(points, ev_indices, s_type=0):
    """data_util.py:201-230 — centroid, 1/scale with scale chosen by s_type."""
    centroid = points.mean(0, keepdim=True)
    p = points - centroid
    if s_type == 0:
        e = p[ev_indices]
        scale = ((e[:, 0] - e[:, 1]) ** 2).sum(1).sqrt().mean()
    elif s_type == 1:
        scale = ((p.max(0)[0] - p.min(0)[0]) ** 2).sum().sqrt()
    elif s_type == 2:
        scale = p.abs().max()
    else:
        scale = (p ** 2).sum(1).max().sqrt()
    scale = 1 / scale
    return p * scale, centroid, scale


def calc_weight(node_pos, node_normal, edge_index):
    """data_util.py:383-398 — clamp(n_i.n_j, 1e-3) * exp(|p_i-p_j|^2 / (-2*mean_len + 1e-12)),
    mean_len over ALL given edges (self loops included)."""
    p = node_pos[edge_index]
    l2 = ((p[0] - p[1]) ** 2).sum(1)
    mean_len = l2.sqrt().mean()
    nn_ = node_normal[edge_index]
    dn = (nn_[0] * nn_[1]).sum(1)
    return torch.clamp(dn, 0.001) * (l2 / (-2 * mean_len + 1e-12)).exp()


def build_edge_fv(fv_indices):
    """data_util.py:478-489 — facet->vertex incidence COO [2, 3F]."""
    f = fv_indices.shape[0]
    return torch.stack([torch.arange(f).repeat_interleave(3), fv_indices.reshape(-1)])


def build_edge_vf(vf_indices):
    """data_util.py:459-475 — vertex->facet incidence COO, -1 pads dropped."""
    v, k = vf_indices.shape
    i = torch.arange(v).repeat_interleave(k)
    j = vf_indices.long().reshape(-1)
    keep = j > -1
    return torch.stack([i[keep], j[keep]])


def build_facet_graph(fv_indices, vf_indices):
    """data_util.py:436-456 — 1-ring facet graph (self included), coalesced."""
    fv, vf = fv_indices.long(), vf_indices.long()
    f = fv.shape[0]
    j = vf[fv, :].reshape(f, -1)
    i = torch.arange(f).unsqueeze(1).expand_as(j)
    keep = j > -1
    return coalesce(torch.stack([i[keep], j[keep]]), None, f, f)[0]


def build_vertex_graph(ev_indices, vv_indices):
    """data_util.py:407-433 — 2-ring vertex graph (dead code upstream, dataset.py:214)."""
    ev, vv = ev_indices.long(), vv_indices.long()
    n = vv.shape[0]
    row = torch.cat([ev[:, 0], ev[:, 1]])
    col = torch.cat([ev[:, 1], ev[:, 0]])
    j = vv[col]
    i = row.unsqueeze(1).expand_as(j)
    keep = j > -1
    return coalesce(torch.stack([i[keep], j[keep]]), None, n, n)[0]


def update_position(points, fv_indices, vf_indices, face_normals, n_iter=20, depth_direction=None, lmd=1):
    """data_util.py:492-526 — scatter-mean variant."""
    fv, vf = fv_indices.long(), vf_indices.long()
    v_idx, f_idx = build_edge_vf(vf)
    n_adj = face_normals[f_idx]
    for _ in range(n_iter):
        cent = points[fv].mean(1)
        d = cent[f_idx] - points[v_idx]
        step = n_adj * (n_adj * d).sum(1, keepdim=True)
        res = scatter(step, v_idx, dim=0, reduce="mean")
        if depth_direction is not None:
            res = (res * depth_direction).sum(1, keepdim=True) * depth_direction
        points = points + res
    return points


def update_position2(points, fv_indices, vf_indices, face_normals, n_iter=20, depth_direction=None):
    """data_util.py:529-556 — padded-gather variant (pad face -> zero normal, count clamped >= 1)."""
    fv, vf = fv_indices.long(), vf_indices.long()
    cnt = (vf > -1).sum(-1, keepdim=True).clamp(min=1)
    fn = torch.cat([face_normals, face_normals.new_zeros(1, 3)])
    n_adj = fn[vf]
    for _ in range(n_iter):
        cent = points[fv].mean(1)
        d = cent[vf] - points.unsqueeze(1)
        step = n_adj * (n_adj * d).sum(-1, keepdim=True)
        mean = step.sum(1) / cnt
        if depth_direction is not None:
            mean = (mean * depth_direction).sum(1, keepdim=True) * depth_direction
        points = points + mean
    return points
