"""Pooling / fusion layers with the reference's surface (/root/reference/code/net_util.py),
running on libgeobi kernels.

* ``PoolingLayer(in_channel, pool_type, pool_step, edge_weight_type, wei_param)``  net_util.py:56-245
* ``DualFusionLayer(in_channel)``                                                  net_util.py:248-278
* ``pool_edge`` / ``pool_face``                                                    net_util.py:289-302

Differences a caller can see: the Graclus matching is deterministic given the node
visiting order (``PoolingLayer.perm_fn``; default = ``torch.randperm`` on the device, as the
upstream CPU kernel draws one), and ``PoolingLayer.forced`` / ``.trace`` allow teacher forcing
and inspection in tests.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch.nn import Linear, Parameter, init

from . import nn as gnn
from . import ops
from .data import Data
from .ops import CSRGraph


def _match_csr(data) -> Optional[CSRGraph]:
    """Source-indexed CSR (stable edge order) + weights, self loops dropped — what graclus sees.
    Also performs the reference's write-back of the stripped edge list (net_util.py:163-167)."""
    n = data.x.size(0)
    if "csr" in data and data.csr.n == n and not data.csr.rejected:   # a PoolingLayer's output: its coarse CSR (weights included) is attached
        g = data.csr
        return None if g.cap == 0 else g
    ei = data.edge_index
    w = data.edge_weight if "edge_weight" in data else None
    tag = gnn.tag_of(ei)
    st = tag.get("sorted")
    if st is not None and st[0].n == n and st[3] is w and not st[0].rejected:   # nn.input_graph built it from this very (edge_index, edge_weight)
        g, ei2, w2, _ = st
        if ei.size(1) == 0:
            return None
        # No sync here: the entry count (which also carries the verdict of the coalesced-list check) is read back
        # asynchronously and looked at behind the first cluster-count sync of the pooling loop (PoolingLayer._forward).
        # The reference's write-back of the stripped list (net_util.py:163-167) happens when somebody reads it.
        g.read_back_async()

        def stripped_edge_index(g=g, ei=ei, ei2=ei2):
            nnz = g.nnz
            if nnz == ei.size(1):
                return ei
            ei_s = ei2[:, :nnz]
            t2 = gnn.tag_of(ei_s)
            t2["tgt"] = t2["src"] = g
            return ei_s

        def stripped_edge_weight(g=g, ei=ei, w=w, w2=w2):
            if w is None:
                return None
            return w if g.nnz == ei.size(1) else w2[:g.nnz]

        data.set_lazy("edge_index", stripped_edge_index)
        if w is not None:
            data.set_lazy("edge_weight", stripped_edge_weight)
        data.csr = g                     # set after the lazies: assigning edge_index / edge_weight drops a stale csr
        return g
    g = tag.get("src")
    if g is not None and g.n == n and not g.rejected:    # built by us: no self loops, CSR order == edge order
        if g.cap == 0:
            return None
        return g.with_weight(w)
    g = ops.csr_from_coo(ei, n, w, ops.COO_DROP_SELF)    # stable sort by source; syncs once (nnz = edges that are not loops)
    if g.nnz == 0:
        return None
    if g.nnz != ei.size(1):                              # strip the loops from the caller-visible list too (no second sync)
        ei2, w2 = ops.remove_self_loops(ei, w, g.nnz)
        if "tgt" in tag:
            gnn.tag_of(ei2)["tgt"] = tag["tgt"]
        data.edge_index, data.edge_weight = ei2, w2
    return g


def _minmax(v):
    return (v - v.min()) / (v.max() - v.min() + 1e-12)


class PoolingLayer(torch.nn.Module):
    def __init__(self, in_channel, pool_type="max", pool_step=2, edge_weight_type=0, wei_param=2):
        super().__init__()
        assert pool_type in ["max", "mean"]
        self.pool_type, self.pool_step = pool_type, pool_step
        self.edge_weight_type, self.wei_param = edge_weight_type, wei_param
        if edge_weight_type in [4, 5]:
            self.lin = Linear(in_channel, in_channel)
        if edge_weight_type in [3, 4, 5]:
            self.att_l = Parameter(torch.empty(1, in_channel))
            self.att_r = Parameter(torch.empty(1, in_channel))
            init.xavier_uniform_(self.att_l.data, gain=1.414)
            init.xavier_uniform_(self.att_r.data, gain=1.414)
        self._clusts, self._unpool_i32_val = None, None
        self._unpool_i64 = None
        self.perm_fn = None          # callable(n) -> permutation (upstream draws torch.randperm); None = random keys on device
        self.forced = None           # list of raw label tensors (teacher forcing)
        self.trace = []

    # ------------------------------------------------------------------ net_util.py:160-240
    def _get_edge_weight(self, data) -> Optional[CSRGraph]:
        g = _match_csr(data)
        if g is None:
            return None
        t, x, w = self.edge_weight_type, data.x.detach(), g.w
        if t == -1:
            nw = None
        elif t == 0:
            nw = w
        elif t == 1:
            nw = ops.edge_weight_feat(x, g, 1, self.wei_param)
        elif t == 2:
            nw = ops.edge_weight_feat(x, g, 2, self.wei_param, w)
        elif t == 10:
            nw = ops.edge_weight_feat(x, g, 10, 2.0, w)
        elif t in (3, 4, 5):
            h = x if t == 3 else F.leaky_relu(self.lin(x), 0.2)
            al, ar = (h * self.att_l).sum(-1), (h * self.att_r).sum(-1)
            ei = g.edge_index()
            s = torch.sigmoid((al[ei[0]] + ar[ei[1]]) + (al[ei[1]] + ar[ei[0]]))
            nw = s if t != 5 else (s + w) / 2
        elif t == 6:
            nw = _minmax(w)
        else:
            d2 = ops.edge_weight_feat(x, g, 0)
            if t == 7:
                nw = _minmax(-d2)
            elif t == 8:
                nw = _minmax((d2 / (-2)).exp())
            elif t == 9:
                nw = _minmax(w) + _minmax((d2 / (-2)).exp())
            else:
                raise ValueError(f"edge_weight_type {t}")
        return g.with_weight(None if nw is None else nw.detach().contiguous())

    # ------------------------------------------------------------------ net_util.py:76-158
    def forward(self, data, visual=False):
        with ops.size_ref(data.x.size(0)):     # size-stable allocations for everything derived from this level
            return self._forward(data, visual)

    def _forward(self, data, visual=False):
        g = self._get_edge_weight(data)
        g_in = data.csr if "csr" in data else None     # the level's CSR as built (it may carry an asynchronous count read-back)
        x, pos = data.x, data.pos
        edge_dual = data.edge_dual if "edge_dual" in data else None
        face = data.fv_indices if "fv_indices" in data else None
        dev = x.device
        if g is None:            # no edges at all: every node is its own cluster at each step
            g = CSRGraph(torch.zeros(x.size(0) + 1, dtype=torch.int32, device=dev),
                         torch.empty(0, dtype=torch.int32, device=dev), x.size(0), 0, None, True)
        empty_in = g.cap == 0
        clusts, self.trace = [], []
        op = ops.OP_MAX if self.pool_type == "max" else ops.OP_MEAN
        for step in range(self.pool_step):
            n = x.size(0)
            if self.forced is not None:
                label, perm = self.forced[step].to(dev).to(torch.int32), None
            else:
                perm = None if self.perm_fn is None else self.perm_fn(n).to(dev)   # None: random priority keys on the device
                label, _ = ops.graclus(g, perm, use_weight=g._w is not None)
            self.trace.append((g, perm, label))
            if self.forced is None:
                # the rest of the step is one library call (same kernels, queued right behind the count read-back); in a training
                # step the pooled features re-enter the autograd graph through PoolStepFn (backward: geobi_segment_max_bwd / gather)
                train = torch.is_grad_enabled() and x.requires_grad
                x_in = x
                cluster, nc, mrowptr, members, x, g, pos = ops.pool_step(g, label, x_in.detach() if train else x_in, op, pos)
                if train:
                    from .autograd import PoolStepFn
                    x = PoolStepFn.apply(x_in, [x], mrowptr, members, nc, op, cluster)
                if g_in is not None and g_in._pin is not None:
                    g_in.nnz             # already on the host (queued before the sync pool_step just did): raises on a bad verdict
                clusts.append(cluster)
                edge_dual = None if edge_dual is None else cluster.long()[edge_dual]
                if empty_in or g.cap == 0:
                    break
                continue
            cluster, nc = ops.relabel_clusters(label)
            if g_in is not None and g_in._pin is not None:
                g_in.nnz
            clusts.append(cluster)
            # matcher output = clusters of one or two nodes: member CSR without a sort; arbitrary (forced) labels: general path
            mrowptr, members = ops.group_by(cluster, nc) if self.forced is not None else ops.group_pairs(label, cluster, nc)
            if torch.is_grad_enabled() and x.requires_grad:
                from .autograd import SegmentMaxFn, SegmentMeanFn
                x = SegmentMaxFn.apply(x, mrowptr, members, nc) if op == ops.OP_MAX else SegmentMeanFn.apply(x, mrowptr, members, nc, cluster)
            else:
                x = ops.segment_reduce(x, mrowptr, members, nc, op)
            g = ops.pool_edges(g, cluster, mrowptr, members, nc)
            pos = None if pos is None else ops.segment_reduce(pos, mrowptr, members, nc, ops.OP_MEAN)
            edge_dual = None if edge_dual is None else cluster.long()[edge_dual]
            # upstream breaks here when no edge is left (net_util.py:139); continuing is equivalent (every node becomes its
            # own cluster: identity pooling) and keeps the edge count on the device, so we only stop when it is known.
            if empty_in or g.cap == 0:
                break
        # composition of the per-step maps (net_util.py:153-156): deferred to the first read (the decoder, several layers
        # later) - right here the GPU queue is empty behind the last count read-back and the next conv should go out first
        self._clusts, self._unpool_i32_val = clusts, None
        self._unpool_i64 = None
        out = Data(x, None, edge_dual=edge_dual, pos=pos, fv_indices=face)

        def coarse_edge_index(g=g):
            ei = g.edge_index()                      # syncs once (entry count), then one csr_to_coo launch
            gnn.attach_symmetric_csr(ei, g, has_self_loops=False)
            return ei

        def coarse_edge_weight(g=g):
            g.nnz
            return g.w

        # the coarse COO list and its weights exist for callers that look at them (net_util.py:158 returns them); the
        # next conv / pooling layer walks `csr` directly and needs neither the list nor the entry count on the host
        out.set_lazy("edge_index", coarse_edge_index)
        out.set_lazy("edge_weight", coarse_edge_weight)
        out.csr = g
        return out

    @property
    def _unpool_i32(self):
        if self._unpool_i32_val is None and self._clusts:
            up = self._clusts[-1]
            for c in self._clusts[-2::-1]:
                up = torch.index_select(up, 0, c)
            self._unpool_i32_val = up.contiguous()
        return self._unpool_i32_val

    @property
    def unpool_map(self):
        """int32 fine-node -> coarse-node map of the last forward (None: identity); FeaStConv(row_map=...) consumes it."""
        return self._unpool_i32

    @property
    def unpooling_indices(self):
        """int64 map as upstream keeps it (net_util.py:153-156); converted from the int32 map on first read."""
        if self._unpool_i32 is None:
            return None
        if self._unpool_i64 is None:
            self._unpool_i64 = self._unpool_i32.long()
        return self._unpool_i64

    def unpooling(self, x, out=None):
        if self._unpool_i32 is None:
            return x
        if torch.is_grad_enabled() and x.requires_grad:
            from .autograd import GatherRowsFn
            return GatherRowsFn.apply(x, self._unpool_i32)
        return ops.gather_rows(x, self._unpool_i32, out=out)


class DualFusionLayer(torch.nn.Module):
    """Not instantiated by DualGNN (and edge_dual is nulled at dataset.py:252,260); kept for API parity."""

    def __init__(self, in_channel):
        super().__init__()
        self.lin_v1 = Linear(in_channel * 2, in_channel)
        self.lin_v2 = Linear(in_channel, in_channel)
        self.lin_f1 = Linear(in_channel * 2, in_channel)
        self.lin_f2 = Linear(in_channel, in_channel)

    @staticmethod
    def fusion(x_i, g: CSRGraph, x_j):
        """cat(x_i, mean over the incident rows of x_j) — the incidence mean is one CSR segment reduce."""
        return torch.cat([x_i, ops.segment_reduce(x_j, g.rowptr, g.nbr, g.n, ops.OP_MEAN)], dim=1)

    def forward(self, data_v, data_f):
        m, n = data_v.x.shape[0], data_f.x.shape[0]
        ed = torch.stack([data_v.edge_dual, data_f.edge_dual], dim=0)
        big = max(m, n)
        g_vf = ops.csr_from_coo(ed, big, None, ops.COO_SORT_NBR | ops.COO_DEDUP)
        g_fv = ops.csr_from_coo(ed, big, None, ops.COO_BY_COL | ops.COO_SORT_NBR | ops.COO_DEDUP)
        g_vf = CSRGraph(g_vf.rowptr[:m + 1], g_vf.nbr, m, g_vf.nnz)
        g_fv = CSRGraph(g_fv.rowptr[:n + 1], g_fv.nbr, n, g_fv.nnz)
        x_v = self.fusion(data_v.x, g_vf, data_f.x)
        x_f = self.fusion(data_f.x, g_fv, data_v.x)
        x_v = F.leaky_relu(self.lin_v2(F.leaky_relu(self.lin_v1(x_v), 0.2)), 0.2)
        x_f = F.leaky_relu(self.lin_f2(F.leaky_relu(self.lin_f1(x_f), 0.2)), 0.2)
        return x_v, x_f


def pool_edge(cluster, edge_index, edge_attr=None, op="mean"):
    """net_util.py:289-295 on int64 edge lists (API parity; the layers above stay in CSR)."""
    if op not in ("mean", "add"):
        raise ValueError(op)
    n = cluster.size(0)
    ei = cluster.long()[edge_index.reshape(-1)].view(2, -1)
    flags = ops.COO_DROP_SELF | ops.COO_SORT_NBR | ops.COO_DEDUP | (ops.COO_W_MEAN if op == "mean" else 0)
    keep = ei[0] != ei[1]
    if int(keep.sum()) == 0:
        return ei[:, keep], (None if edge_attr is None else edge_attr[keep])
    g = ops.csr_from_coo(ei, n, edge_attr, flags)
    return g.edge_index(), g.w


def pool_face(cluster, fv_indices):
    """net_util.py:298-302."""
    face = cluster.long()[fv_indices.reshape(-1)].view(-1, 3)
    bad = (face[:, 0] == face[:, 1]) | (face[:, 0] == face[:, 2]) | (face[:, 1] == face[:, 2])
    return face[~bad]
