"""FeaStConv and graph-tag helpers (host side of the fused conv kernel).

``FeaStConv`` keeps torch_geometric's constructor, parameter names and
state-dict layout (PyG 2.x: ``lin.weight [9*C_out, C_in]``, ``u.weight [9, C_in]``,
``c [9]``, ``bias [C_out]``) so checkpoints written by the reference
(test_dual.py:130) load unchanged; ``load_state_dict`` also accepts the PyG 1.x
keys ``weight [C_in, 9*C_out]`` / ``u [C_in, 9]``.  Ctor sites in the reference:
/root/reference/code/network.py:258-268; call sites :271-299.
"""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import config, ops
from .ops import CSRGraph

_TAG = "_geobi_graphs"


def tag_of(edge_index: torch.Tensor) -> dict:
    tag = getattr(edge_index, _TAG, None)
    if tag is None:
        tag = {}
        setattr(edge_index, _TAG, tag)
    return tag


def attach_symmetric_csr(edge_index: torch.Tensor, g: CSRGraph, has_self_loops: bool):
    """Record that `g` is edge_index's adjacency (minus self loops) and is symmetric, so it
    serves both as the conv's target-CSR and the matcher's source-CSR."""
    tag = tag_of(edge_index)
    tag["tgt"] = g
    if not has_self_loops:
        tag["src"] = g          # CSR entry order == edge order only when nothing was dropped
    return edge_index


def conv_csr(edge_index: Union[torch.Tensor, CSRGraph], n: int) -> CSRGraph:
    """Target-indexed CSR without self loops for FeaStConv (cached on the tensor)."""
    if isinstance(edge_index, CSRGraph):
        return edge_index
    tag = tag_of(edge_index)
    g = tag.get("tgt")
    if g is None or g.n != n or g.rejected:
        g = ops.csr_from_coo(edge_index, n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR, sync=False)
        tag["tgt"] = g
    return g


def input_graph(data, n: int) -> CSRGraph:
    """Conv CSR of an input-level graph.  A Data that carries `coalesced_undirected=True` (dataset.py builds its graphs
    with to_undirected / build_facet_graph, so ours sets it) gets the sort-free builder, whose result also serves the
    matcher of the first pooling layer (net_util._match_csr); anything else goes through conv_csr."""
    if "csr" in data and data.csr.n == n and not data.csr.rejected:      # a PoolingLayer's output carries its coarse CSR; edge_index stays unmaterialised
        return data.csr
    ei = data.edge_index
    tag = tag_of(ei)
    g = tag.get("tgt")
    if g is not None and g.n == n and not g.rejected:
        return g
    if "coalesced_undirected" in data and data.coalesced_undirected:
        w = data.edge_weight if "edge_weight" in data else None
        g, ei2, w2 = ops.csr_from_sorted_coo(ei, n, w)
        tag["tgt"] = g
        tag["sorted"] = (g, ei2, w2, w)
        return g
    return conv_csr(ei, n)


class FeaStConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, heads: int = 1):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.lin = torch.nn.Linear(in_channels, heads * out_channels, bias=False)
        self.u = torch.nn.Linear(in_channels, heads, bias=False)
        self.c = torch.nn.Parameter(torch.empty(heads))
        self.bias = torch.nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / self.in_channels ** 0.5
        with torch.no_grad():
            self.lin.weight.uniform_(-bound, bound)
            self.u.weight.uniform_(-bound, bound)
            self.c.normal_(0.0, 0.1)
            self.bias.normal_(0.0, 0.1)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # PyG 1.x layout -> 2.x layout
        w, u = prefix + "weight", prefix + "u"
        if w in state_dict and prefix + "lin.weight" not in state_dict:
            state_dict[prefix + "lin.weight"] = state_dict.pop(w).t().contiguous()
        if u in state_dict and prefix + "u.weight" not in state_dict:
            state_dict[prefix + "u.weight"] = state_dict.pop(u).t().contiguous()
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, x: torch.Tensor, edge_index: Union[torch.Tensor, CSRGraph], act_slope: float = 1.0,
                out: Optional[torch.Tensor] = None, row_map: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`row_map` (inference only): the conv reads x[row_map[v]] for node v — PoolingLayer.unpooling fused in."""
        if self.heads != 9:
            raise NotImplementedError("libgeobi's fused FeaSt kernel is specialised for heads=9 (network.py:258-268)")
        n_nodes = x.size(0) if row_map is None else row_map.numel()
        g = conv_csr(edge_index, n_nodes)
        if torch.is_grad_enabled() and (x.requires_grad or self.lin.weight.requires_grad):
            if out is not None or row_map is not None:
                raise RuntimeError("FeaStConv(out=..., row_map=...) are inference-only fusions; call it under torch.no_grad()")
            from .autograd import FeaStFn            # training step: same forward kernels + backward kernels
            return FeaStFn.apply(x, self.lin.weight, self.u.weight, self.c, self.bias, g, float(act_slope), config.precision_code())
        return ops.feast_fwd(x, g, self.lin.weight, self.u.weight, self.c, self.bias, act_slope=act_slope, out=out,
                             precision=config.precision_code(), row_map=row_map)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, heads={self.heads}"
