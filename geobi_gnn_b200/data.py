"""Minimal stand-in for ``torch_geometric.data.Data`` (absent from this image).

Only the behaviour the reference path relies on (SURVEY.md 8b "container
contract"): attribute get/set, ``hasattr``, assigning ``None`` removes the key
(dataset.py:252-262), the common keys read as ``None`` when absent,
``num_nodes`` (train_dual.py:246), ``.to(device)``, and the constructor
``Data(x, edge_index, **kwargs)`` used at net_util.py:158.
"""
from __future__ import annotations

import torch

_OPTIONAL = frozenset(("x", "edge_index", "edge_attr", "y", "pos", "edge_weight", "normal"))


class _Lazy:
    """A value computed on first read (Data.set_lazy)."""
    __slots__ = ("thunk",)

    def __init__(self, thunk):
        self.thunk = thunk


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.__dict__["_items"] = {}
        for key, value in (("x", x), ("edge_index", edge_index), ("edge_attr", edge_attr), ("y", y), ("pos", pos)):
            setattr(self, key, value)
        for key, value in kwargs.items():
            setattr(self, key, value)

    def __setattr__(self, key, value):
        if key in ("edge_index", "edge_weight"):
            self._items.pop("csr", None)       # the attached CSRGraph (PoolingLayer output) described the old lists
        if value is None:
            self._items.pop(key, None)
        else:
            self._items[key] = value

    def set_lazy(self, key, thunk):
        """`key` reads as thunk() (evaluated once, on first access).  PoolingLayer uses it for the coarse edge_index /
        edge_weight: the network itself walks the attached CSR (`csr`), so the int64 COO list - and the device sync
        that sizing it needs - only happen if a caller actually looks at it."""
        self._items[key] = _Lazy(thunk)

    def __getattr__(self, key):
        items = self.__dict__["_items"]
        if key in items:
            v = items[key]
            if isinstance(v, _Lazy):
                v = v.thunk()
                if v is None:
                    items.pop(key, None)
                    return None
                items[key] = v
            return v
        if key in _OPTIONAL:
            return None
        raise AttributeError(f"Data has no attribute {key!r}")

    def __delattr__(self, key):
        self._items.pop(key, None)

    def __contains__(self, key):
        return key in self._items

    def __repr__(self):
        parts = []
        for k, v in self._items.items():
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else (f"{k}=<lazy>" if isinstance(v, _Lazy) else f"{k}={v!r}"))
        return "Data(" + ", ".join(parts) + ")"

    @property
    def keys(self):
        return list(self._items)

    def shallow_copy(self):
        """New Data over the same items; lazy items stay lazy (and shared: the first reader of either copy evaluates the thunk)."""
        out = Data()
        out._items.update(self._items)
        return out

    def tensors(self):
        """(key, tensor) of the tensor-valued items that exist NOW: lazy items stay unevaluated."""
        return [(k, v) for k, v in self._items.items() if torch.is_tensor(v)]

    @property
    def num_nodes(self):
        for key in ("x", "pos", "normal"):
            if key in self._items:
                return self._items[key].size(0)
        if "edge_index" in self._items and self.edge_index.numel():
            return int(self.edge_index.max()) + 1
        return 0

    def to(self, device, non_blocking=False):
        for k in list(self._items):
            v = getattr(self, k)
            if torch.is_tensor(v):
                self._items[k] = v.to(device, non_blocking=non_blocking)   # same-device: identity, graph tags survive
        return self

    def clone(self):
        out = Data()
        for k in list(self._items):
            v = getattr(self, k)
            out._items[k] = v.clone() if torch.is_tensor(v) else v
        return out
