"""Tensor-level wrappers over the C ABI (include/geobi.h).

PyTorch is used only as plumbing: device memory, the current stream and dtype
bookkeeping.  Every function here launches hand-written sm_100a kernels from
libgeobi.so; none has a CPU path (CPU tensors raise).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import torch

from . import _lib

PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
FEAST_REUSE_WS = 0x100
COO_BY_COL, COO_DROP_SELF, COO_SORT_NBR, COO_DEDUP, COO_W_MEAN, COO_SYMMETRIZE = 1, 2, 4, 8, 16, 32
OP_MEAN, OP_MAX, OP_SUM = 0, 1, 2

import itertools

_launch_ticks = itertools.count()   # kernel launches made through the ABI (bench.py reports it); next() is atomic under the GIL, so
_launch_base = [0]                  # forwards running on several host threads do not lose counts


def launch_count() -> int:
    # peek without consuming: itertools.count has no read accessor, so take a tick and compensate
    _launch_base[0] -= 1
    return next(_launch_ticks) + _launch_base[0] + 1


def _count(n=1):
    for _ in range(n):
        next(_launch_ticks)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.GeobiError("geobi_gnn_b200 ops need CUDA tensors (there is no CPU fallback)")


_dummy = {}


def _ptr(t: Optional[torch.Tensor]):
    """Device pointer of `t`; empty tensors (data_ptr()==0) get a valid 256-byte dummy so that the ABI's
    NULL-means-absent convention keeps working."""
    if t is None:
        return None
    if t.numel() == 0:
        d = _dummy.get(t.device)
        if d is None:
            d = _dummy[t.device] = torch.zeros(256, dtype=torch.uint8, device=t.device)
        return C.c_void_p(d.data_ptr())
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


_ws_cache = {}


def _ws(nbytes: int, device, slot: int = 0) -> torch.Tensor:
    """Scratch buffer per (device, slot, current stream), grown on demand: ops queued on different streams never share one."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, slot, torch._C._cuda_getCurrentRawStream(idx))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.25), 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def release_workspaces():
    _ws_cache.clear()


# ----------------------------------------------------------------------------- size-stable allocation
# Level sizes depend on the (random) matching and change by a percent or so from one forward to the next.  If every
# intermediate were allocated at its exact size the caching allocator would keep meeting sizes it has no cached block for
# and fall back to cudaMalloc / cudaFree (device-wide syncs: 10-300 ms spikes).  Inside `size_ref(ref)` every variable-size
# buffer is instead carved from a capacity rounded up to ref/16 steps (never above ref), so the request sizes are the same
# on every forward and the allocator always hits its cache.
import threading

_TLS = threading.local()     # per host thread: the forward and a front end running on a helper thread (HostBatchRunner) must not share it


class size_ref:
    def __init__(self, ref: int):
        self.ref = int(ref)

    def __enter__(self):
        self.prev = getattr(_TLS, "ref", None)
        _TLS.ref = self.ref
        return self

    def __exit__(self, *a):
        _TLS.ref = self.prev


def _cap(n: int, ref) -> int:
    if ref is None or n > ref:
        return max(n, 1)
    q = max(ref // 16, 1)
    return min(max(ref, 1), -(-max(n, 1) // q) * q)


def valloc(n: int, tail, dtype, device, ref=None) -> torch.Tensor:
    """Uninitialised [n, *tail] tensor (contiguous) backed by a capacity-bucketed allocation (see above)."""
    ref = getattr(_TLS, "ref", None) if ref is None else ref
    m = 1
    for t in tail:
        m *= t
    flat = torch.empty(_cap(n, ref) * m, dtype=dtype, device=device)
    return flat[:n * m].view((n,) + tuple(tail))


def _rows(x: torch.Tensor):
    """(tensor, ld, channels) of a 2-D fp32 tensor whose rows are contiguous (column slices allowed)."""
    if x.dim() != 2 or x.dtype != torch.float32:
        raise _lib.GeobiError(f"expected a 2-D float32 tensor, got {tuple(x.shape)} {x.dtype}")
    if x.size(1) > 1 and x.stride(1) != 1:
        x = x.contiguous()
    if x.size(0) > 1 and x.stride(0) < x.size(1):
        x = x.contiguous()
    return x, (x.stride(0) if x.size(0) > 1 else max(x.size(1), 1)), x.size(1)


# ----------------------------------------------------------------------------- graph container
_pin_pool = []      # recycled pinned int32[1] landing pads (CSRGraph.read_back_async)


class CSRGraph:
    """int32 CSR adjacency without self loops (+ optional per-entry fp32 weight).

    `nbr` / `w` may be capacity-sized buffers: the exact entry count lives on the device in rowptr[n] and is only
    copied to the host (one stream sync) when somebody asks for `.nnz` — kernels walk rows and never need it."""

    def __init__(self, rowptr, nbr, n, nnz=None, w=None, symmetric=False, _ei=None):
        self.rowptr, self._nbr, self.n, self._nnz, self._w, self.symmetric = rowptr, nbr, n, nnz, w, symmetric
        self._pin = None      # (pinned int32[1], event): an asynchronous read-back of rowptr[n] already in flight
        # weak: the edge_index tensor carries this graph in its tag (nn.tag_of); a strong back-reference would be a cycle
        # that only the cyclic GC frees, i.e. hundreds of MB per forward released late and in bursts
        self._ei_ref = None if _ei is None else weakref.ref(_ei)

    @property
    def cap(self) -> int:
        """Upper bound on the number of entries (buffer length)."""
        return self._nnz if self._nnz is not None else self._nbr.numel()

    @property
    def nnz(self) -> int:
        if self._nnz is None:
            if self._pin is not None:
                pin, ev = self._pin
                ev.synchronize()                  # free when a later sync of the same stream has already happened
                self._nnz = int(pin[0])
                self._pin = None
                _pin_pool.append(pin)
            else:
                self._nnz = int(self.rowptr[self.n].item()) if self.n > 0 else 0      # syncs
        if self._nnz < 0:      # geobi_csr_from_sorted_coo's verdict
            raise _lib.GeobiError("edge list was declared coalesced_undirected but is not: its non-loop entries must be sorted "
                                  "by (row, col), unique, in range, and contain (j,i) for every (i,j)")
        return self._nnz

    @property
    def rejected(self) -> bool:
        """True once the device-side check of a `coalesced_undirected` claim is known to have failed (such a graph is empty
        and must not be served from a tensor's tag cache)."""
        return self._nnz is not None and self._nnz < 0

    @property
    def nbr(self) -> torch.Tensor:
        return self._nbr if self._nnz is None else self._nbr[:self._nnz]

    @property
    def w(self) -> Optional[torch.Tensor]:
        return self._w if (self._w is None or self._nnz is None) else self._w[:self._nnz]

    def read_back_async(self):
        """Queue the copy of the entry count (and, for csr_from_sorted_coo, its verdict) into pinned memory; `.nnz` then
        costs no stream drain once any later synchronisation of this stream has returned."""
        if self._nnz is None and self._pin is None and self.n > 0:
            try:                                   # several host threads may run forwards at once (bench.py, serving)
                pin = _pin_pool.pop()
            except IndexError:
                pin = torch.empty(1, dtype=torch.int32).pin_memory()
            pin.copy_(self.rowptr[self.n:self.n + 1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.rowptr.device))
            self._pin = (pin, ev)

    def with_weight(self, w) -> "CSRGraph":
        return CSRGraph(self.rowptr, self._nbr, self.n, self._nnz, w, self.symmetric, self._ei_ref() if self._ei_ref else None)

    def edge_index(self) -> torch.Tensor:
        """int64 [2, nnz], row-major sorted when rows are sorted (coalesce layout).  Syncs if nnz is not known yet."""
        ei = self._ei_ref() if self._ei_ref is not None else None
        if ei is None:
            nnz = self.nnz
            flat = torch.empty(2 * _cap(nnz, max(self._nbr.numel(), nnz)), dtype=torch.int64, device=self.rowptr.device)
            ei = flat[:2 * nnz].view(2, nnz)
            if nnz:
                lib = _lib.load()
                _lib.check(lib.geobi_csr_to_coo(_ptr(self.rowptr), _ptr(self._nbr), self.n, nnz, _ptr(ei), _stream()), "csr_to_coo")
                _count()
            self._ei_ref = weakref.ref(ei)
        return ei


def exclusive_scan(v: torch.Tensor) -> torch.Tensor:
    _need_cuda(v)
    lib = _lib.load()
    v = v.contiguous().to(torch.int32)
    n = v.numel()
    out = torch.empty(n + 1, dtype=torch.int32, device=v.device)
    ws = _ws(lib.geobi_scan_ws_bytes(n), v.device)
    _lib.check(lib.geobi_exclusive_scan_i32(_ptr(v), _ptr(out), n, _ptr(ws), ws.numel(), _stream()), "exclusive_scan")
    _count(1)
    return out


def csr_from_coo(edge_index: torch.Tensor, n_nodes: int, weight: Optional[torch.Tensor] = None, flags: int = 0,
                 want_eid: bool = False, sync: bool = True):
    """COO int64 [2,E] -> CSRGraph (+ eid).  sync=True reads nnz back (one stream sync) and trims the buffers;
    sync=False leaves nnz on the device (lazy)."""
    _need_cuda(edge_index, weight)
    lib = _lib.load()
    ei = edge_index.contiguous()
    if ei.dtype != torch.int64:
        ei = ei.long()
    e = ei.size(1)
    cap = e * (2 if flags & COO_SYMMETRIZE else 1)
    dev = ei.device
    rowptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
    nbr = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    w = None if weight is None else weight.contiguous().float()
    w_out = None if w is None else torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
    eid = torch.empty(max(cap, 1), dtype=torch.int64, device=dev) if want_eid else None
    ws = _ws(lib.geobi_csr_from_coo_ws_bytes(e, n_nodes, flags), dev)
    nnz = C.c_int64(0)
    row, col = ei[0], ei[1]
    _lib.check(lib.geobi_csr_from_coo(_ptr(row), _ptr(col), _ptr(w), e, n_nodes, flags, _ptr(rowptr), _ptr(nbr), _ptr(w_out),
                                      _ptr(eid), C.byref(nnz) if sync else None, _ptr(ws), ws.numel(), _stream()), "csr_from_coo")
    _count(6)            # count, scan, fill, sort_rows, scan, compact_rows
    if not sync:
        g = CSRGraph(rowptr, nbr, n_nodes, None, w_out)
        return (g, eid) if want_eid else g
    k = int(nnz.value)
    g = CSRGraph(rowptr, nbr[:k], n_nodes, k, None if w_out is None else w_out[:k])
    return (g, eid[:k]) if want_eid else g


def csr_from_sorted_coo(edge_index: torch.Tensor, n_nodes: int, weight: Optional[torch.Tensor] = None, check_symmetric: bool = True):
    """CSR of a coalesced, undirected edge list (self loops allowed anywhere, they are dropped): no sort, no sync.
    Returns (CSRGraph with lazy nnz, stripped edge list buffer int64 [2,E], stripped weight buffer or None); the first
    `.nnz` read raises if the list was not what it was declared to be."""
    _need_cuda(edge_index, weight)
    lib = _lib.load()
    ei = edge_index.contiguous()
    if ei.dtype != torch.int64:
        ei = ei.long()
    e = ei.size(1)
    dev = ei.device
    rowptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
    nbr = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
    w = None if weight is None else weight.contiguous().float()
    w_out = None if w is None else torch.empty(max(e, 1), dtype=torch.float32, device=dev)
    ei_out = torch.empty((2, max(e, 1)), dtype=torch.int64, device=dev)
    ws = _ws(lib.geobi_csr_from_sorted_coo_ws_bytes(e), dev)
    _lib.check(lib.geobi_csr_from_sorted_coo(_ptr(ei[0]), _ptr(ei[1]), _ptr(w), e, n_nodes, 1 if check_symmetric else 0, _ptr(rowptr),
                                             _ptr(nbr), _ptr(w_out), _ptr(ei_out) if e else None, _ptr(ws), ws.numel(), _stream()),
               "csr_from_sorted_coo")
    _count(6 if check_symmetric else 5)
    return CSRGraph(rowptr, nbr, n_nodes, 0 if e == 0 else None, w_out, True), ei_out, w_out


def build_facet_graph_csr(fv: torch.Tensor, vf: torch.Tensor, vf_sorted: bool = False, drop_self: bool = False) -> CSRGraph:
    """Facet 1-ring CSR with the self entry (data_util.build_facet_graph).  Syncs once.
    vf_sorted: every row of vf is ascending with its -1 pads last (topology.DeviceTriMesh) - merge instead of sort.
    drop_self (with vf_sorted): leave the self entries out - the loop-free CSR the network walks."""
    if drop_self and not vf_sorted:
        raise ValueError("build_facet_graph_csr(drop_self=True) needs vf_sorted=True")
    _need_cuda(fv, vf)
    lib = _lib.load()
    fv, vf = fv.contiguous().long(), vf.contiguous().long()
    f, v, k = fv.size(0), vf.size(0), vf.size(1)
    dev = fv.device
    rowptr = torch.empty(f + 1, dtype=torch.int32, device=dev)
    nbr = torch.empty(max(3 * k * f, 1), dtype=torch.int32, device=dev)
    nnz = C.c_int64(0)
    if vf_sorted:
        ws = _ws(lib.geobi_build_facet_graph_sorted_ws_bytes(f), dev)
        _lib.check(lib.geobi_build_facet_graph_sorted(_ptr(fv), _ptr(vf), f, v, k, int(drop_self), _ptr(rowptr), _ptr(nbr), C.byref(nnz), _ptr(ws), ws.numel(),
                                                      _stream()), "build_facet_graph_sorted")
        _count(3)
    else:
        ws = _ws(lib.geobi_build_facet_graph_ws_bytes(f, k), dev)
        _lib.check(lib.geobi_build_facet_graph(_ptr(fv), _ptr(vf), f, v, k, _ptr(rowptr), _ptr(nbr), C.byref(nnz), _ptr(ws), ws.numel(),
                                               _stream()), "build_facet_graph")
        _count(4)
    n = int(nnz.value)
    return CSRGraph(rowptr, nbr[:n].clone(), f, n, None, True)


def graclus(g: CSRGraph, perm: Optional[torch.Tensor] = None, weight: Optional[torch.Tensor] = None, use_weight: bool = True,
            keys: Optional[torch.Tensor] = None, check: bool = False):
    """Exact greedy matching -> raw labels int32 [N] (= min(u, partner)).
    Visiting order: `perm` (torch_cluster's randperm semantics; rank = inverse permutation) or int32 priority `keys`
    (u before v iff (keys[u], u) < (keys[v], v)); with neither, i.i.d. random keys = a uniformly random order, no sort."""
    lib = _lib.load()
    dev = g.rowptr.device
    _need_cuda(g.rowptr, perm, keys)
    if perm is not None:
        rank = valloc(g.n, (), torch.int32, dev)
        rank[perm.to(dev).long()] = torch.arange(g.n, dtype=torch.int32, device=dev)
    elif keys is not None:
        rank = keys.to(torch.int32).contiguous()
    else:
        rank = valloc(g.n, (), torch.int32, dev)
        if g.n:
            torch.randint(0, 2 ** 31 - 1, (g.n,), dtype=torch.int32, device=dev, out=rank)
    w = (g._w if weight is None else weight) if use_weight else None
    label = valloc(g.n, (), torch.int32, dev)
    ws = _ws(lib.geobi_graclus_ws_bytes(g.n), dev, slot=2)
    und = C.c_int(0)
    _lib.check(lib.geobi_graclus(_ptr(g.rowptr), _ptr(g._nbr), _ptr(w), _ptr(rank), g.n, _ptr(label), C.byref(und) if check else None,
                                 _ptr(ws), ws.numel(), _stream()), "graclus")
    _count(2)
    return label, und.value


def relabel_clusters(label: torch.Tensor):
    """consecutive_cluster: (dense int32 cluster ids [N], number of clusters).  Syncs."""
    _need_cuda(label)
    lib = _lib.load()
    label = label.contiguous().to(torch.int32)
    n = label.numel()
    cluster = valloc(n, (), torch.int32, label.device)
    ws = _ws(lib.geobi_relabel_ws_bytes(n), label.device)
    nc = C.c_int64(0)
    _lib.check(lib.geobi_relabel_clusters(_ptr(label), n, _ptr(cluster), C.byref(nc), _ptr(ws), ws.numel(), _stream()), "relabel_clusters")
    _count(3)
    return cluster, int(nc.value)


def group_by(cluster: torch.Tensor, n_clusters: int):
    _need_cuda(cluster)
    lib = _lib.load()
    cluster = cluster.contiguous().to(torch.int32)
    n = cluster.numel()
    dev = cluster.device
    mrowptr = valloc(n_clusters + 1, (), torch.int32, dev)
    members = valloc(max(n, 1), (), torch.int32, dev)
    ws = _ws(lib.geobi_group_by_ws_bytes(n, n_clusters), dev)
    _lib.check(lib.geobi_group_by(_ptr(cluster), n, n_clusters, _ptr(mrowptr), _ptr(members), _ptr(ws), ws.numel(), _stream()), "group_by")
    _count(6)
    return mrowptr, members[:n]


def group_pairs(label: torch.Tensor, cluster: torch.Tensor, n_clusters: int):
    """Member CSR for matching labels (clusters of <= 2 nodes, label = min member): no sort."""
    _need_cuda(label, cluster)
    lib = _lib.load()
    n = cluster.numel()
    dev = cluster.device
    mrowptr = valloc(n_clusters + 1, (), torch.int32, dev)
    members = valloc(max(n, 1), (), torch.int32, dev)
    ws = _ws(lib.geobi_group_pairs_ws_bytes(n_clusters), dev)
    _lib.check(lib.geobi_group_pairs(_ptr(label), _ptr(cluster), n, n_clusters, _ptr(mrowptr), _ptr(members), _ptr(ws), ws.numel(),
                                     _stream()), "group_pairs")
    _count(4)
    return mrowptr, members[:n]


def pool_edges(g: CSRGraph, cluster: torch.Tensor, mrowptr: torch.Tensor, members: torch.Tensor, n_clusters: int) -> CSRGraph:
    """net_util.pool_edge on CSR.  Asynchronous: the coarse nnz stays on the device (CSRGraph.nnz reads it lazily)."""
    lib = _lib.load()
    dev = g.rowptr.device
    cap = g.cap
    out_rowptr = valloc(n_clusters + 1, (), torch.int32, dev)
    out_nbr = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    gw = g._w
    out_w = None if gw is None else torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
    ws = _ws(lib.geobi_pool_edges_ws_bytes(cap, n_clusters), dev)
    _lib.check(lib.geobi_pool_edges(_ptr(g.rowptr), _ptr(g._nbr), _ptr(gw), g.n, cap, _ptr(cluster), _ptr(mrowptr), _ptr(members),
                                    n_clusters, _ptr(out_rowptr), _ptr(out_nbr), _ptr(out_w), None, _ptr(ws), ws.numel(),
                                    _stream()), "pool_edges")
    _count(5)            # pool_count, scan, pool_rows, scan, pool_copy
    return CSRGraph(out_rowptr, out_nbr[:cap] if cap else out_nbr[:0], n_clusters, 0 if cap == 0 else None,
                    None if out_w is None else (out_w[:cap] if cap else out_w[:0]), g.symmetric)


def pool_step(g: CSRGraph, label: torch.Tensor, x: torch.Tensor, op: int, pos: Optional[torch.Tensor] = None):
    """relabel_clusters + group_pairs + segment_reduce(x[, pos]) + pool_edges in one library call (one sync: the cluster
    count).  Outputs are allocated at capacity up front and trimmed to views afterwards.
    Returns (cluster, n_clusters, mrowptr, members, x_coarse, coarse CSRGraph, pos_coarse)."""
    _need_cuda(g.rowptr, label, x, pos)
    lib = _lib.load()
    dev = x.device
    n, cap = g.n, g.cap
    x, ldx, c = _rows(x)
    cluster = valloc(n, (), torch.int32, dev)
    mrowptr = valloc(n + 1, (), torch.int32, dev)
    members = valloc(max(n, 1), (), torch.int32, dev)
    x_out = valloc(n, (c,), torch.float32, dev)
    out_rowptr = valloc(n + 1, (), torch.int32, dev)
    out_nbr = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    gw = g._w
    out_w = None if gw is None else torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
    p = ldp = cp = pos_out = None
    if pos is not None:
        p, ldp, cp = _rows(pos)
        pos_out = valloc(n, (cp,), torch.float32, dev)
    ws = _ws(lib.geobi_pool_step_ws_bytes(n, cap), dev)
    nc = C.c_int64(0)
    _lib.check(lib.geobi_pool_step(_ptr(g.rowptr), _ptr(g._nbr), _ptr(gw), n, cap, _ptr(label), _ptr(x), ldx, c, op, _ptr(p), ldp or 0, cp or 0,
                                   _ptr(cluster), _ptr(mrowptr), _ptr(members), _ptr(x_out), c, _ptr(pos_out), cp or 0, _ptr(out_rowptr),
                                   _ptr(out_nbr), _ptr(out_w), C.byref(nc), _ptr(ws), ws.numel(), _stream()), "pool_step")
    _count(3 + 4 + 1 + 5 + (1 if pos is not None else 0))
    k = int(nc.value)
    gc = CSRGraph(out_rowptr[:k + 1], out_nbr[:cap] if cap else out_nbr[:0], k, 0 if cap == 0 else None,
                  None if out_w is None else (out_w[:cap] if cap else out_w[:0]), g.symmetric)
    return cluster, k, mrowptr[:k + 1], members[:n], x_out[:k], gc, (None if pos_out is None else pos_out[:k])


def remove_self_loops(edge_index: torch.Tensor, weight: Optional[torch.Tensor], count: int):
    """Order-preserving removal of row==col pairs when the surviving `count` is already known (no sync):
    torch_geometric.utils.remove_self_loops (net_util.py:163)."""
    _need_cuda(edge_index, weight)
    lib = _lib.load()
    ei = edge_index.contiguous().long()
    e = ei.size(1)
    dev = ei.device
    out = torch.empty((2, count), dtype=torch.int64, device=dev)
    w = None if weight is None else weight.contiguous().float()
    w_out = None if w is None else torch.empty(count, dtype=torch.float32, device=dev)
    ws = _ws(lib.geobi_remove_self_loops_ws_bytes(e), dev)
    _lib.check(lib.geobi_remove_self_loops(_ptr(ei[0]), _ptr(ei[1]), _ptr(w), e, count, _ptr(out), _ptr(w_out), _ptr(ws), ws.numel(),
                                           _stream()), "remove_self_loops")
    _count(3)
    return out, w_out


# ----------------------------------------------------------------------------- segment / gather
def segment_reduce(x: torch.Tensor, rowptr: Optional[torch.Tensor], idx: torch.Tensor, n_seg: int, op: int, fixed: int = 0,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, idx)
    lib = _lib.load()
    x, ldx, c = _rows(x)
    idx = idx.contiguous().to(torch.int32)
    if out is None:
        out = valloc(n_seg, (c,), torch.float32, x.device)
    o, ldo, _ = _rows(out)
    _lib.check(lib.geobi_segment_reduce(_ptr(x), ldx, c, _ptr(rowptr), _ptr(idx), fixed, n_seg, op, _ptr(o), ldo, _stream()), "segment_reduce")
    _count()
    return out


def gather_rows(x: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, idx)
    lib = _lib.load()
    x, ldx, c = _rows(x)
    idx = idx.contiguous().to(torch.int32)
    n = idx.numel()
    if out is None:
        out = valloc(n, (c,), torch.float32, x.device)
    o, ldo, _ = _rows(out)
    _lib.check(lib.geobi_gather_rows(_ptr(x), ldx, c, _ptr(idx), n, _ptr(o), ldo, _stream()), "gather_rows")
    _count()
    return out


def edge_weight_feat(x: torch.Tensor, g: CSRGraph, mode: int, param: float = 2.0, w_in: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, g.rowptr)
    lib = _lib.load()
    x, ldx, c = _rows(x)
    cap = g.cap
    w_out = torch.empty(max(cap, 1), dtype=torch.float32, device=x.device)[:cap]
    _lib.check(lib.geobi_edge_weight_feat(_ptr(x), ldx, c, _ptr(g.rowptr), _ptr(g._nbr), g.n, _ptr(w_in), mode, float(param), _ptr(w_out),
                                          _stream()), "edge_weight_feat")
    _count()
    return w_out


def calc_weight(pos: torch.Tensor, nrm: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    _need_cuda(pos, nrm, edge_index)
    lib = _lib.load()
    pos, nrm = pos.contiguous().float(), nrm.contiguous().float()
    ei = edge_index.contiguous().long()
    e = ei.size(1)
    w = torch.empty(e, dtype=torch.float32, device=pos.device)
    ws = _ws(lib.geobi_calc_weight_ws_bytes(e), pos.device)
    _lib.check(lib.geobi_calc_weight(_ptr(pos), _ptr(nrm), _ptr(ei[0]), _ptr(ei[1]), e, _ptr(w), _ptr(ws), ws.numel(), _stream()), "calc_weight")
    _count(3)
    return w


RING_MAX_VALENCE = 24


def mesh_vertex_csr(fv: torch.Tensor, vf_rowptr: torch.Tensor, corners: torch.Tensor, n_verts: int) -> CSRGraph:
    """Vertex 1-ring CSR (symmetric, ascending rows, no loops) from the faces-around-vertices CSR (group_by over fv.reshape(-1)).
    Syncs once (entry count).  Valences above RING_MAX_VALENCE raise: use csr_from_coo on the half edges."""
    _need_cuda(fv, vf_rowptr, corners)
    lib = _lib.load()
    fv = fv.contiguous().long()
    f = fv.size(0)
    dev = fv.device
    rowptr = torch.empty(n_verts + 1, dtype=torch.int32, device=dev)
    nbr = torch.empty(max(6 * f, 1), dtype=torch.int32, device=dev)
    ws = _ws(lib.geobi_mesh_vertex_csr_ws_bytes(n_verts), dev)
    nnz = C.c_int64(0)
    _lib.check(lib.geobi_mesh_vertex_csr(_ptr(fv), _ptr(vf_rowptr), _ptr(corners.contiguous()), n_verts, f, _ptr(rowptr), _ptr(nbr), C.byref(nnz),
                                         _ptr(ws), ws.numel(), _stream()), "mesh_vertex_csr")
    _count(3)
    k = int(nnz.value)
    return CSRGraph(rowptr, nbr[:k], n_verts, k, None, True)


def pad_rows(rowptr: torch.Tensor, members: torch.Tensor, k: int, divisor: int = 1) -> torch.Tensor:
    """Member CSR -> padded int64 [n, k] table, entries divided by `divisor`, -1 pads (the reference's vf_indices / vv_indices layout)."""
    _need_cuda(rowptr, members)
    lib = _lib.load()
    n = rowptr.numel() - 1
    out = torch.empty((n, max(int(k), 0)), dtype=torch.int64, device=rowptr.device)
    if n and k:
        _lib.check(lib.geobi_pad_rows(_ptr(rowptr), _ptr(members.contiguous()), n, int(k), int(divisor), _ptr(out), _stream()), "pad_rows")
        _count()
    return out


def mean_edge_length_csr(pos: torch.Tensor, g: CSRGraph) -> torch.Tensor:
    """Mean undirected edge length of a symmetric loop-free CSR as a 0-dim fp32 device tensor (no sync)."""
    _need_cuda(pos, g.rowptr)
    lib = _lib.load()
    pos = pos.contiguous().float()
    out = torch.empty((), dtype=torch.float32, device=pos.device)
    ws = _ws(lib.geobi_calc_weight_ws_bytes(g.n), pos.device)
    _lib.check(lib.geobi_mean_edge_length_csr(_ptr(pos), _ptr(g.rowptr), _ptr(g._nbr), g.n, _ptr(out), _ptr(ws), ws.numel(), _stream()),
               "mean_edge_length_csr")
    _count(3)
    return out


def calc_weight_csr(pos: torch.Tensor, nrm: torch.Tensor, g: CSRGraph, n_loops: int) -> torch.Tensor:
    """calc_weight for the entries of a loop-free CSR, in CSR order; the mean edge length counts `n_loops` zero-length self loops
    as the reference's list does (geobi_calc_weight_csr)."""
    _need_cuda(pos, nrm, g.rowptr)
    lib = _lib.load()
    pos, nrm = pos.contiguous().float(), nrm.contiguous().float()
    w = torch.empty(max(g._nbr.numel(), 1), dtype=torch.float32, device=pos.device)
    ws = _ws(lib.geobi_calc_weight_ws_bytes(g.n), pos.device)
    _lib.check(lib.geobi_calc_weight_csr(_ptr(pos), _ptr(nrm), _ptr(g.rowptr), _ptr(g._nbr), g.n, int(n_loops), _ptr(w), _ptr(ws), ws.numel(),
                                         _stream()), "calc_weight_csr")
    _count(3)
    return w


# ----------------------------------------------------------------------------- conv / heads / transfer
def feast_fwd(x: torch.Tensor, g: CSRGraph, W: torch.Tensor, U: torch.Tensor, c: torch.Tensor, bias: torch.Tensor,
              act_slope: float = 1.0, out: Optional[torch.Tensor] = None, precision: int = PREC_FP32,
              row_map: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FeaStConv forward on a TARGET-indexed CSR without self loops (implicit self loop).
    row_map (int32 [g.n]): node v reads row row_map[v] of x (unpooling fused into the conv); x then has the coarse row count."""
    _need_cuda(x, g.rowptr, W, row_map)
    lib = _lib.load()
    x, ldx, c_in = _rows(x)
    c_out = bias.numel()
    n = g.n if row_map is not None else x.size(0)
    n_src = x.size(0)
    if row_map is not None and not feast_row_map_supported(x):
        x = gather_rows(x, row_map)            # generic fallback: materialise the unpooled rows
        x, ldx, c_in = _rows(x)
        row_map, n_src = None, n
    if out is None:
        out = valloc(n, (c_out,), torch.float32, x.device)
    o, ldo, _ = _rows(out)
    if o.data_ptr() != out.data_ptr():
        raise _lib.GeobiError("feast_fwd: `out` must have contiguous rows")
    ws = _ws(lib.geobi_feast_fwd_ws_bytes(max(n, n_src), c_in, c_out, precision & 0xff), x.device, slot=1)
    _lib.check(lib.geobi_feast_fwd(_ptr(x), ldx, n, c_in, _ptr(g.rowptr), _ptr(g._nbr), _ptr(row_map), n_src, _ptr(W.contiguous()),
                                   _ptr(U.contiguous()), _ptr(c.contiguous()), _ptr(bias.contiguous()), c_out, float(act_slope), _ptr(o), ldo,
                                   precision, _ptr(ws), ws.numel(), _stream()), "feast_fwd")
    # bf16: weight split + projection + aggregation + GEMM; the fused 64->32 kernel merges the last two; fp32 has no weight split
    _count(3 if (precision & 0xff) == PREC_FP32 or ((precision & 0xff) == PREC_BF16X3 and c_in == 64 and c_out == 32) else 4)
    return out


def feast_row_map_supported(x: torch.Tensor) -> bool:
    return x.size(1) in (32, 64, 128) and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and x.stride(1) == 1


def linear_tc(a: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, act_slope: float = 1.0, out: Optional[torch.Tensor] = None,
              precision: int = PREC_BF16X3):
    """act(a @ W.T + bias) on tcgen05 tensor cores (bf16 operands, fp32 accumulate)."""
    _need_cuda(a, W, bias)
    lib = _lib.load()
    a, lda, k = _rows(a)
    n = W.size(0)
    if out is None:
        out = torch.empty((a.size(0), n), dtype=torch.float32, device=a.device)
    o, ldo, _ = _rows(out)
    ws = _ws(lib.geobi_linear_tc_ws_bytes(a.size(0), k, n), a.device)
    _lib.check(lib.geobi_linear_tc(_ptr(a), lda, a.size(0), k, _ptr(W.contiguous()), n, _ptr(bias.contiguous()), float(act_slope),
                                   _ptr(o), ldo, precision, _ptr(ws), ws.numel(), _stream()), "linear_tc")
    _count(3)
    return out


def fc_head_fwd(f: torch.Tensor, W1, b1, W2, b2, epilogue: int = 0, res: Optional[torch.Tensor] = None,
                res2: Optional[torch.Tensor] = None, precision: int = PREC_FP32) -> torch.Tensor:
    _need_cuda(f, W1)
    lib = _lib.load()
    f, ldf, c_in = _rows(f)
    n, hidden, c_out = f.size(0), W1.size(0), W2.size(0)
    oc = 3 if (epilogue == 2 and c_out == 1) else c_out
    out = torch.empty((n, oc), dtype=torch.float32, device=f.device)
    ldres = ldres2 = 0
    if res is not None:
        res, ldres, _ = _rows(res)
    if res2 is not None:
        res2, ldres2, _ = _rows(res2)
    ws = _ws(lib.geobi_fc_head_ws_bytes(hidden), f.device, slot=1)
    _lib.check(lib.geobi_fc_head_fwd(_ptr(f), ldf, n, c_in, _ptr(W1.contiguous()), _ptr(b1.contiguous()), hidden, _ptr(W2.contiguous()),
                                     _ptr(b2.contiguous()), c_out, epilogue, _ptr(res), ldres, _ptr(res2), ldres2, _ptr(out), oc,
                                     precision, _ptr(ws), ws.numel(), _stream()), "fc_head_fwd")
    _count(1 if precision == PREC_FP32 else 2)      # tensor-core path: weight split + fused head
    return out


def face_normal(points: torch.Tensor, fv: torch.Tensor) -> torch.Tensor:
    _need_cuda(points, fv)
    lib = _lib.load()
    p, ldp, _ = _rows(points)
    fv = fv.contiguous().long()
    out = torch.empty((fv.size(0), 3), dtype=torch.float32, device=p.device)
    _lib.check(lib.geobi_face_normal(_ptr(p), ldp, _ptr(fv), fv.size(0), _ptr(out), 3, _stream()), "face_normal")
    _count()
    return out


def v2f_transfer(feat_v: torch.Tensor, fv: torch.Tensor, xf: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[xf | corner mean of feat_v | face normal of feat_v]  (network.py:335-337)."""
    _need_cuda(feat_v, fv, xf)
    lib = _lib.load()
    p, ldv, _ = _rows(feat_v)
    xf_, ldxf, cf = _rows(xf)
    fv = fv.contiguous().long()
    f = fv.size(0)
    if out is None:
        out = torch.empty((f, cf + 6), dtype=torch.float32, device=p.device)
    o, ldo, _ = _rows(out)
    _lib.check(lib.geobi_v2f_transfer(_ptr(p), ldv, _ptr(fv), _ptr(xf_), ldxf, cf, f, _ptr(o), ldo, _stream()), "v2f_transfer")
    _count()
    return out


def update_position(points, fv, vf, face_normals, n_iter=20, depth_direction=None) -> torch.Tensor:
    _need_cuda(points, fv, vf, face_normals)
    lib = _lib.load()
    p = points.contiguous().float()
    fv, vf = fv.contiguous().long(), vf.contiguous().long()
    fn = face_normals.contiguous().float()
    d = None if depth_direction is None else depth_direction.contiguous().float()
    v, f, k = p.size(0), fv.size(0), vf.size(1)
    out = torch.empty_like(p)
    ws = _ws(lib.geobi_update_position_ws_bytes(v, f), p.device)
    _lib.check(lib.geobi_update_position(_ptr(p), _ptr(fv), _ptr(vf), k, _ptr(fn), int(n_iter), _ptr(d), v, f, _ptr(out), _ptr(ws),
                                         ws.numel(), _stream()), "update_position")
    _count(2 * int(n_iter))
    return out
