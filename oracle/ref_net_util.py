"""CPU oracle: pooling / fusion layers of the reference.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/code/net_util.py:56-302 (PoolingLayer with its 12
edge-weight modes, DualFusionLayer, pool_edge, pool_face) in plain PyTorch on
CPU over the shims in oracle/pyg.py.  Pinned against the reference's own net_util.py executed over the same shims
(tests/golden/make_reference_golden.py, tests/test_reference_golden.py); the shims themselves are unpinned.

Two test hooks that the reference does not have (its matching is random, so
parity needs them; SURVEY.md section 8c "parity protocol"):
  * ``PoolingLayer.perm_fn``   callable n -> visiting order for the greedy matcher
                               (default torch.randperm, as upstream);
  * ``PoolingLayer.forced``    list of raw cluster-label tensors to use instead of
                               running the matcher (teacher forcing);
  * ``PoolingLayer.trace``     after forward: per step (edge_index, weight, perm, labels).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.nn import Linear, Parameter, init

from . import pyg
from .pyg import Data, coalesce, consecutive_cluster, graclus, pool_pos, remove_self_loops, scatter


def pool_edge(cluster, edge_index, edge_attr=None, op="mean"):
    """net_util.py:289-295 — relabel, drop loops, coalesce (weights reduced by MEAN)."""
    n = cluster.size(0)
    edge_index = cluster[edge_index.reshape(-1)].view(2, -1)
    edge_index, edge_attr = remove_self_loops(edge_index, edge_attr)
    if edge_index.numel() > 0:
        edge_index, edge_attr = coalesce(edge_index, edge_attr, n, n, op=op)
    return edge_index, edge_attr


def pool_face(cluster, fv_indices):
    """net_util.py:298-302 — relabel faces, drop degenerate ones."""
    f = cluster[fv_indices.reshape(-1)].view(-1, 3)
    bad = (f[:, 0] == f[:, 1]) | (f[:, 0] == f[:, 2]) | (f[:, 1] == f[:, 2])
    return f[~bad]


def _sq_feat_dist(x, edge_index):
    d = x[edge_index[0]] - x[edge_index[1]]
    return (d * d).sum(1)


def _minmax(v):
    return (v - v.min()) / (v.max() - v.min() + 1e-12)


class PoolingLayer(torch.nn.Module):
    def __init__(self, in_channel, pool_type="max", pool_step=2, edge_weight_type=0, wei_param=2):
        super().__init__()
        assert pool_type in ("max", "mean")
        self.pool_type, self.pool_step = pool_type, pool_step
        self.edge_weight_type, self.wei_param = edge_weight_type, wei_param
        if edge_weight_type in (4, 5):
            self.lin = Linear(in_channel, in_channel)
        if edge_weight_type in (3, 4, 5):
            self.att_l = Parameter(torch.empty(1, in_channel))
            self.att_r = Parameter(torch.empty(1, in_channel))
            init.xavier_uniform_(self.att_l.data, gain=1.414)
            init.xavier_uniform_(self.att_r.data, gain=1.414)
        self.unpooling_indices = None
        self.perm_fn = torch.randperm
        self.forced = None
        self.trace = []

    # net_util.py:160-240
    def _get_edge_weight(self, data):
        w = data.edge_weight if "edge_weight" in data else None
        ei, w = remove_self_loops(data.edge_index, w)
        if ei.numel() == 0:
            return None
        data.edge_index, data.edge_weight = ei, w       # written back (net_util.py:166-167)
        t, x = self.edge_weight_type, data.x
        if t == -1:
            return None
        if t == 0:
            return w
        if t == 1:
            return (_sq_feat_dist(x, ei) / (-self.wei_param)).exp()
        if t == 2:
            return w * (_sq_feat_dist(x, ei) / (-self.wei_param)).exp()
        if t in (3, 4, 5):
            h = x if t == 3 else F.leaky_relu(self.lin(x), 0.2)
            al, ar = (h * self.att_l).sum(-1), (h * self.att_r).sum(-1)
            a = (al[ei[0]] + ar[ei[1]]) + (al[ei[1]] + ar[ei[0]])
            s = torch.sigmoid(a)
            return s if t != 5 else (s + w) / 2
        if t == 6:
            return _minmax(w)
        if t == 7:
            return _minmax(-_sq_feat_dist(x, ei))
        if t == 8:
            return _minmax((_sq_feat_dist(x, ei) / (-2)).exp())
        if t == 9:
            return _minmax(w) + _minmax((_sq_feat_dist(x, ei) / (-2)).exp())
        if t == 10:
            return w + (_sq_feat_dist(x, ei) / (-2)).exp()
        raise ValueError(t)

    # net_util.py:76-158
    def forward(self, data, visual=False):
        w = self._get_edge_weight(data)
        x, ei, pos = data.x, data.edge_index, data.pos
        edge_dual = data.edge_dual if "edge_dual" in data else None
        face = data.fv_indices if "fv_indices" in data else None
        clusts, self.trace = [], []
        for step in range(self.pool_step):
            n = x.shape[0]
            if self.forced is not None:
                raw, perm = self.forced[step], None
            else:
                perm = self.perm_fn(n)
                raw = graclus(ei, w, n, perm=perm)
            self.trace.append((ei, w, perm, raw))
            cluster, _ = consecutive_cluster(raw)
            clusts.append(cluster)
            x = scatter(x, cluster, dim=0, reduce=self.pool_type)
            ei, w = pool_edge(cluster, ei, w)
            pos = None if pos is None else pool_pos(cluster, pos)
            edge_dual = None if edge_dual is None else cluster[edge_dual]
            if ei.numel() == 0:
                break
        up = clusts[-1]
        for c in clusts[-2::-1]:
            up = up[c]
        self.unpooling_indices = up
        return Data(x, ei, edge_dual=edge_dual, edge_weight=w, pos=pos, fv_indices=face)

    def unpooling(self, x):
        return x if self.unpooling_indices is None else x[self.unpooling_indices]


class DualFusionLayer(torch.nn.Module):
    """net_util.py:248-278 — not instantiated by DualGNN; kept for API parity."""

    def __init__(self, in_channel):
        super().__init__()
        self.lin_v1 = Linear(in_channel * 2, in_channel)
        self.lin_v2 = Linear(in_channel, in_channel)
        self.lin_f1 = Linear(in_channel * 2, in_channel)
        self.lin_f2 = Linear(in_channel, in_channel)

    @staticmethod
    def fusion(x_i, edge_dual, x_j):
        row, col = edge_dual
        return torch.cat([x_i, scatter(x_j[col], row, dim=0, reduce="mean")], 1)

    def forward(self, data_v, data_f):
        m, n = data_v.x.shape[0], data_f.x.shape[0]
        ed = torch.stack([data_v.edge_dual, data_f.edge_dual])
        if ed.numel() > 0:
            ed, _ = coalesce(ed, None, m, n)
        x_v = self.fusion(data_v.x, ed, data_f.x)
        x_f = self.fusion(data_f.x, ed.flip(0), data_v.x)
        x_v = F.leaky_relu(self.lin_v2(F.leaky_relu(self.lin_v1(x_v), 0.2)), 0.2)
        x_f = F.leaky_relu(self.lin_f2(F.leaky_relu(self.lin_f1(x_f), 0.2)), 0.2)
        return x_v, x_f
