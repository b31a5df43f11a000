"""Disjoint-union batching of dual graphs and patch sharding across ranks.

The reference cannot batch (its Collater returns ``batch[0]``, dataset.py:29-31; batch_size is
gradient accumulation, train_dual.py:211-218).  DualGNN has no cross-graph operator on this path
(no BatchNorm, no global pool), so a batch is exactly the disjoint union of its patches with
offset indices: every conv, matching and pooling step acts per connected component.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .data import Data


def collate_dual(patches: Sequence[Tuple[Data, Data]]) -> Tuple[Data, Data, dict]:
    """Union of post-processed (graph_v, graph_f) patches.  Returns (data_v, data_f, slices) where
    slices['v'] / slices['f'] are the [start, end) node ranges of each patch."""
    vs, fs = [p[0] for p in patches], [p[1] for p in patches]
    v_off, f_off, acc_v, acc_f = [], [], 0, 0
    for dv, df in zip(vs, fs):
        v_off.append(acc_v)
        f_off.append(acc_f)
        acc_v += dv.x.size(0)
        acc_f += df.x.size(0)

    def cat(items, key, offs=None, dim=0):
        ts = [getattr(d, key) for d in items]
        if any(t is None for t in ts):
            return None
        if offs is not None:
            ts = [t + o for t, o in zip(ts, offs)]
        return torch.cat(ts, dim)

    data_v = Data(x=cat(vs, "x"), edge_index=cat(vs, "edge_index", v_off, 1), edge_weight=cat(vs, "edge_weight"), y=cat(vs, "y"))
    if all("depth_direction" in d for d in vs):
        data_v.depth_direction = cat(vs, "depth_direction")
    data_f = Data(x=cat(fs, "x"), edge_index=cat(fs, "edge_index", f_off, 1), edge_weight=cat(fs, "edge_weight"), y=cat(fs, "y"),
                  fv_indices=cat(fs, "fv_indices", v_off))
    # the union of sorted lists with increasing offsets is sorted; it stays symmetric and duplicate-free
    if all("coalesced_undirected" in d and d.coalesced_undirected for d in vs):
        data_v.coalesced_undirected = True
    if all("coalesced_undirected" in d and d.coalesced_undirected for d in fs):
        data_f.coalesced_undirected = True
    slices = dict(v=[(o, o + d.x.size(0)) for o, d in zip(v_off, vs)], f=[(o, o + d.x.size(0)) for o, d in zip(f_off, fs)])
    return data_v, data_f, slices


def shard(items: Sequence, rank: int, world: int) -> List:
    """Patches dealt round-robin to ranks (independent units, no collective)."""
    return [it for i, it in enumerate(items) if i % world == rank]


def fresh_view(data: Data) -> Data:
    """New Data over the same tensors with graph tags dropped: what a caller holding only the
    reference's input layout (x, int64 edge_index, edge_weight, fv_indices) would pass (the `coalesced_undirected`
    flag, a fact about how the dataset built the list, travels with it)."""
    if "csr" in data:                 # CSR-native inputs (device front end): the graph IS the attached CSR, the lists are lazy items
        return data.shallow_copy()
    out = Data()
    for k in data.keys:
        v = getattr(data, k)
        out.__setattr__(k, v.view_as(v) if torch.is_tensor(v) and k == "edge_index" else v)
    return out
