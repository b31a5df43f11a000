"""CPU oracle: third-party semantics the reference path reaches.  TEST INFRASTRUCTURE ONLY.

Restated (not copied; the sources are not in /root/reference and the packages
are not installable here) from the published behaviour of

* torch_geometric  — ``Data``, ``remove_self_loops``, ``add_self_loops``,
  ``to_undirected``, ``nn.FeaStConv``, ``nn.graclus``, ``consecutive_cluster``,
  ``pool_pos``                       (imports: /root/reference/code/network.py:7-11,
                                      net_util.py:4-8, dataset.py:6-8)
* torch_scatter    — ``scatter``     (net_util.py:10,132-134,277; data_util.py:521)
* torch_sparse     — ``coalesce``    (net_util.py:9,263,294; data_util.py:432,455)
* torch_cluster    — ``graclus``     (via net_util.py:127)

All unpinned by the reference (no requirements file).  PARITY UNPINNED for this file: the reference's own modules
are executed over these functions to produce tests/golden/reference_*.npz, so a misreading here is shared by both sides.
Everything is plain PyTorch on CPU in the reference's evaluation order (per-edge
FeaSt projection, materialised per-edge tensors, unsorted scatter).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _oracle_lib():
    """liboracle.so built from oracle/graclus.c by `make -C oracle` (or build())."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_graclus_greedy.restype = ctypes.c_int
    return _LIB


# ----------------------------------------------------------------------------- Data
class Data:
    """Attribute bag with torch_geometric.data.Data's observable behaviour on this path:
    assigning None deletes the key, the well-known keys read as None when absent,
    ``num_nodes`` falls back to ``x.size(0)``, ``.to(device)`` moves tensors."""

    _KNOWN = ("x", "edge_index", "edge_attr", "y", "pos", "edge_weight", "normal")

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos, **kwargs).items():
            setattr(self, k, v)

    def __setattr__(self, key, value):
        if value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __getattr__(self, key):
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        if key in Data._KNOWN:
            return None
        raise AttributeError(key)

    def __contains__(self, key):
        return key in self._store

    @property
    def num_nodes(self):
        s = self._store
        if "x" in s:
            return s["x"].size(0)
        if "pos" in s:
            return s["pos"].size(0)
        return int(s["edge_index"].max()) + 1

    def keys(self):
        return list(self._store.keys())

    def to(self, device):
        for k, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[k] = v.to(device)
        return self

    def clone(self):
        d = Data()
        for k, v in self._store.items():
            d._store[k] = v.clone() if torch.is_tensor(v) else v
        return d


# ----------------------------------------------------------------------------- utils
def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loops = torch.arange(n, dtype=edge_index.dtype).unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        edge_attr = torch.cat([edge_attr, edge_attr.new_full((n,) + edge_attr.shape[1:], fill_value)])
    return torch.cat([edge_index, loops], 1), edge_attr


def coalesce(index, value, m, n, op="add"):
    """Sort COO by row*n+col, merge duplicates (value reduced with `op`)."""
    key = index[0] * n + index[1]
    key, perm = torch.sort(key, stable=True)
    uniq, inv = torch.unique_consecutive(key, return_inverse=True)
    out_index = torch.stack([uniq // n, uniq % n])
    if value is None:
        return out_index, None
    value = value[perm]
    out = scatter(value, inv, dim=0, dim_size=uniq.numel(), reduce={"add": "sum"}.get(op, op))
    return out_index, out


def to_undirected(edge_index, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    row = torch.cat([edge_index[0], edge_index[1]])
    col = torch.cat([edge_index[1], edge_index[0]])
    return coalesce(torch.stack([row, col]), None, n, n)[0]


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    """torch_scatter.scatter along dim 0: sum / mean (sum / clamp(count,1)) / max (empty -> 0)."""
    assert dim == 0
    n = (int(index.max()) + 1 if index.numel() else 0) if dim_size is None else dim_size
    shape = (n,) + tuple(src.shape[1:])
    if reduce in ("sum", "add", "mean"):
        out = torch.zeros(shape, dtype=src.dtype).index_add_(0, index, src)
        if reduce == "mean":
            cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
            out = out / cnt.clamp_(min=1).view((n,) + (1,) * (src.dim() - 1))
        return out
    if reduce == "max":
        idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
        out = torch.zeros(shape, dtype=src.dtype)
        return out.scatter_reduce(0, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def consecutive_cluster(src):
    uniq, inv = torch.unique(src, sorted=True, return_inverse=True)
    perm = torch.empty_like(uniq).scatter_(0, inv, torch.arange(inv.numel()))
    return inv, perm


def pool_pos(cluster, pos):
    return scatter(pos, cluster, dim=0, reduce="mean")


def graclus_csr(edge_index, weight, num_nodes):
    """The CSR the torch_cluster wrapper hands to its kernel: self loops dropped,
    edges stably sorted by row.  Returns (rowptr int64 [N+1], col int64, weight or None)."""
    row, col = edge_index
    keep = row != col
    row, col = row[keep], col[keep]
    if weight is not None:
        weight = weight[keep]
    order = torch.sort(row, stable=True)[1]
    row, col = row[order], col[order]
    if weight is not None:
        weight = weight[order].contiguous().float()
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=num_nodes), 0)
    return rowptr, col.contiguous(), weight


def graclus(edge_index, weight=None, num_nodes=None, perm=None):
    """Greedy heavy-edge matching.  `perm` is the node visiting order (upstream draws
    torch.randperm(N); pass it explicitly for reproducible parity).  Labels = min(u,v)."""
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    if perm is None:
        perm = torch.randperm(n)
    rowptr, col, w = graclus_csr(edge_index, weight, n)
    label = torch.empty(n, dtype=torch.int64)
    perm = perm.contiguous().to(torch.int64)
    rc = _oracle_lib().oracle_graclus_greedy(
        ctypes.c_int64(n), ctypes.c_void_p(rowptr.data_ptr()), ctypes.c_void_p(col.data_ptr()),
        ctypes.c_void_p(w.data_ptr() if w is not None else None),
        ctypes.c_void_p(perm.data_ptr()), ctypes.c_void_p(label.data_ptr()))
    assert rc == 0
    return label


def graclus_python(edge_index, weight, num_nodes, perm):
    """Same algorithm in pure Python loops (small cases; cross-checks graclus.c)."""
    rowptr, col, w = graclus_csr(edge_index, weight, num_nodes)
    rowptr, col = rowptr.tolist(), col.tolist()
    w = None if w is None else w.tolist()
    label = [-1] * num_nodes
    for u in perm.tolist():
        if label[u] >= 0:
            continue
        label[u] = u
        best, wmax = -1, 0.0
        for e in range(rowptr[u], rowptr[u + 1]):
            v = col[e]
            if label[v] >= 0:
                continue
            if w is None:
                best = v
                break
            if w[e] >= wmax:
                best, wmax = v, w[e]
        if best >= 0:
            label[u] = label[best] = min(u, best)
    return torch.tensor(label, dtype=torch.int64)


# ----------------------------------------------------------------------------- FeaStConv
class FeaStConv(torch.nn.Module):
    """torch_geometric.nn.FeaStConv(in, out, heads) with aggr='mean', add_self_loops=True,
    bias=True (ctor sites /root/reference/code/network.py:258-268).

        q_ij = softmax_h(u(x_j - x_i) + c);  out_i = mean_{j in N(i) + i} sum_h q_ijh (W_h x_j) + b

    State-dict keys follow PyG 2.x: lin.weight [H*C_out, C_in], u.weight [H, C_in], c [H],
    bias [C_out]; head h / output o is row h*C_out + o of lin.weight.
    Kept in the reference evaluation order: the projection runs per EDGE."""

    def __init__(self, in_channels, out_channels, heads=1):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.lin = torch.nn.Linear(in_channels, heads * out_channels, bias=False)
        self.u = torch.nn.Linear(in_channels, heads, bias=False)
        self.c = torch.nn.Parameter(torch.empty(heads))
        self.bias = torch.nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / self.in_channels ** 0.5
        torch.nn.init.uniform_(self.lin.weight, -bound, bound)
        torch.nn.init.uniform_(self.u.weight, -bound, bound)
        torch.nn.init.normal_(self.c, mean=0.0, std=0.1)
        torch.nn.init.normal_(self.bias, mean=0.0, std=0.1)

    def forward(self, x, edge_index):
        n = x.size(0)
        edge_index, _ = remove_self_loops(edge_index)
        edge_index, _ = add_self_loops(edge_index, num_nodes=n)
        src, dst = edge_index
        x_j, x_i = x[src], x[dst]
        q = F.softmax(self.u(x_j - x_i) + self.c, dim=1)                       # [E, H]
        y = self.lin(x_j).view(-1, self.heads, self.out_channels)              # [E, H, C_out]
        msg = (y * q.unsqueeze(-1)).sum(1)                                     # [E, C_out]
        return scatter(msg, dst, dim=0, dim_size=n, reduce="mean") + self.bias
