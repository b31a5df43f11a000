"""Mesh splitting into face patches and stitching of per-patch predictions (host side + two small kernels).

Mirrors the reference's rule exactly (dataset.py:156-193 + data_util.mesh_get_neighbor_np / get_submesh,
data_util.py:55-84,318-336): seed = unvisited face farthest from the centroid, ring-by-ring BFS through faces sharing a
vertex until `sub_size` faces, vertices re-indexed in order of first appearance; predictions of overlapping patches are
averaged (test_dual.py:49-61).  The BFS runs in C++ (libgeobi_host.so) instead of Python loops.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_HOST = None


def _host():
    global _HOST
    if _HOST is None:
        path = os.path.join(_HERE, "libgeobi_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `make -C geobi_gnn_b200/csrc`")
        lib = C.CDLL(path)
        lib.geobi_host_grow_patch.restype = C.c_int64
        lib.geobi_host_submesh.restype = C.c_int64
        lib.geobi_host_obj_parse.restype = C.c_int64
        _HOST = lib
    return _HOST


def _p(a):
    return C.c_void_p(a.ctypes.data)


def mesh_get_neighbor_np(fv_indices, vf_indices, seed_idx, neighbor_count=None, ring_count=None, _stamps=None):
    """data_util.py:55-84 — list of face ids in discovery order."""
    assert neighbor_count is not None or ring_count is not None, "'neighbor_count' and 'ring_count' are both None"
    fv = np.ascontiguousarray(fv_indices, dtype=np.int64)
    vf = np.ascontiguousarray(vf_indices, dtype=np.int64)
    f = fv.shape[0]
    big = np.iinfo(np.int64).max
    nc = big if neighbor_count is None else int(neighbor_count)
    rc = big if ring_count is None else int(ring_count)
    # _stamps = [face stamps u32 [F], vertex stamps u32 [V], epoch]: scratch a caller growing many patches allocates once
    st = [np.zeros(f, dtype=np.uint32), np.zeros(vf.shape[0], dtype=np.uint32), 0] if _stamps is None else _stamps
    st[2] += 1
    out = np.empty(min(nc, f), dtype=np.int64)
    n = _host().geobi_host_grow_patch(_p(fv), _p(vf), C.c_int64(f), C.c_int64(vf.shape[1]), C.c_int64(int(seed_idx)), C.c_int64(nc),
                                      C.c_int64(rc), _p(st[0]), _p(st[1]), C.c_uint32(st[2]), _p(out))
    return out[:n]


def get_submesh(fv_indices, select_faces, _slot=None):
    """data_util.py:318-336 — (V_idx in first-appearance order, faces re-indexed)."""
    fv = np.ascontiguousarray(fv_indices, dtype=np.int64)
    sel = np.ascontiguousarray(select_faces, dtype=np.int64)
    slot = np.full(int(fv.max()) + 1, -1, dtype=np.int64) if _slot is None else _slot
    v_idx = np.empty(min(3 * sel.shape[0], slot.shape[0]), dtype=np.int64)
    faces = np.empty((sel.shape[0], 3), dtype=np.int64)
    nv = _host().geobi_host_submesh(_p(fv), _p(sel), C.c_int64(sel.shape[0]), _p(slot), _p(v_idx), _p(faces))
    return v_idx[:nv].copy(), faces


def get_submesh_device(fv_dev: torch.Tensor, select_faces_dev: torch.Tensor, n_vertices: int):
    """get_submesh (data_util.py:318-336) on the device: (V_idx int64 in first-appearance order, faces re-indexed int64 [Fs,3]).
    First appearance = smallest position in the selected faces' flattened corner list, so the vertex order (and with it every index
    array derived from the patch) equals the host routine's; integer glue on a few tensor ops, no host round trip per patch."""
    faces_g = fv_dev.index_select(0, select_faces_dev.long())
    flat = faces_g.reshape(-1)
    big = flat.numel()
    first = torch.full((n_vertices,), big, dtype=torch.int64, device=fv_dev.device)
    first.scatter_reduce_(0, flat, torch.arange(big, device=fv_dev.device), reduce="amin", include_self=True)
    present = torch.nonzero(first < big).reshape(-1)
    order = torch.argsort(first.index_select(0, present))
    v_idx = present.index_select(0, order)
    remap = torch.empty(n_vertices, dtype=torch.int64, device=fv_dev.device)
    remap[v_idx] = torch.arange(v_idx.numel(), device=fv_dev.device)
    return v_idx, remap[faces_g]


def split_mesh(points, fv_indices, vf_indices, submesh_size, filter_patch_count=0) -> List[Tuple[np.ndarray, int]]:
    """dataset.py:156-193 — [(select_faces, seed), ...]."""
    return list(iter_split_mesh(points, fv_indices, vf_indices, submesh_size, filter_patch_count))


def iter_split_mesh(points, fv_indices, vf_indices, submesh_size, filter_patch_count=0):
    """The same walk as a generator: each (select_faces, seed) is handed out as soon as it is grown, so a consumer can start on
    patch k while patch k+1 is being found (`prefetch` below runs the walk on a helper thread; the library calls release the GIL)."""
    pts = np.asarray(points, dtype=np.float32)
    fv = np.ascontiguousarray(fv_indices, dtype=np.int64)
    vf = np.ascontiguousarray(vf_indices, dtype=np.int64)
    pts = np.ascontiguousarray(pts)
    centroid = np.ascontiguousarray(pts.mean(0, keepdims=True))
    # d2 = ((pts[fv].mean(1) - centroid) ** 2).sum(1) (dataset.py:165-166) in one threaded pass with numpy's fp32 operation
    # order: the seeds are arg-maxima of d2, so it has to match bit for bit (tests/test_abi.py checks it does)
    d2 = np.empty(fv.shape[0], dtype=np.float32)
    lib, nthr = _host(), min(8, os.cpu_count() or 1)
    lib.geobi_host_face_d2(_p(pts), _p(fv), C.c_int64(fv.shape[0]), _p(centroid), _p(d2), C.c_int(nthr))
    stamps = [np.zeros(fv.shape[0], dtype=np.uint32), np.zeros(vf.shape[0], dtype=np.uint32), 0]
    seed = int(np.argmax(d2))
    # upstream rescans `np.where(~flag)` and `d2[left]` after every patch (dataset.py:188-192): O(F) index arrays per patch.
    # Same seeds from a masked copy of d2: covered faces drop to -inf, the next seed is the first arg-max of what is left.
    n_left = C.c_int64(fv.shape[0])
    lib.geobi_host_cover_next_seed.restype = C.c_int64
    while True:
        sel = mesh_get_neighbor_np(fv, vf, seed, neighbor_count=submesh_size, _stamps=stamps)
        nxt = lib.geobi_host_cover_next_seed(_p(d2), C.c_int64(fv.shape[0]), _p(sel), C.c_int64(sel.shape[0]), C.byref(n_left), C.c_int(nthr))
        if len(sel) > filter_patch_count:
            yield sel, seed
        if nxt < 0:
            break
        seed = int(nxt)


def iter_split_mesh_device(points, fv_dev: torch.Tensor, vf_dev: torch.Tensor, submesh_size: int, filter_patch_count: int = 0, centroid=None):
    """The same walk on the device (geobi_bfs_begin / geobi_bfs_grow: ring-parallel BFS with the reference's discovery order,
    dataset.py:156-193, data_util.py:55-84): yields (select_faces int32 DEVICE tensor, seed) - identical patches to `split_mesh`.
    `points`: host array or device tensor [V,3]; `centroid` = points.mean(0) as numpy computes it (fp32 pairwise sum; the seeds are
    arg-maxima decided by the last bit, so it is taken from the host copy unless the caller passes it); fv int64 [F,3] and the padded
    incidence table vf int64 [V,k] on the device."""
    from . import _lib, ops
    lib = _lib.load()
    dev = fv_dev.device
    if centroid is None:
        pts_h = points.detach().cpu().numpy() if torch.is_tensor(points) else np.asarray(points)
        centroid = np.ascontiguousarray(np.asarray(pts_h, dtype=np.float32).mean(0, keepdims=True)).reshape(3)
    centroid = np.ascontiguousarray(centroid, dtype=np.float32).reshape(3)
    pts_d = points if torch.is_tensor(points) and points.is_cuda else torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32)).to(dev)
    pts_d = pts_d.contiguous().float()
    fv_dev, vf_dev = fv_dev.contiguous().long(), vf_dev.contiguous().long()
    f, k = fv_dev.size(0), vf_dev.size(1)
    ws = torch.empty(lib.geobi_bfs_ws_bytes(f), dtype=torch.uint8, device=dev)      # carries the walk's state between calls
    seed, n_out, nxt = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    with torch.cuda.device(dev):
        stream = ops._stream()
        _lib.check(lib.geobi_bfs_begin(ops._ptr(pts_d), ops._ptr(fv_dev), f, C.c_void_p(centroid.ctypes.data), C.byref(seed), ops._ptr(ws),
                                       ws.numel(), stream), "bfs_begin")
        cur = int(seed.value)
        while cur >= 0:
            out = torch.empty(min(int(submesh_size), f), dtype=torch.int32, device=dev)
            _lib.check(lib.geobi_bfs_grow(ops._ptr(fv_dev), ops._ptr(vf_dev), f, k, cur, int(submesh_size), ops._ptr(out), C.byref(n_out),
                                          C.byref(nxt), ops._ptr(ws), ws.numel(), stream), "bfs_grow")
            if int(n_out.value) > filter_patch_count:
                yield out[:int(n_out.value)], cur
            cur = int(nxt.value)


def split_mesh_device(points, fv_dev, vf_dev, submesh_size, filter_patch_count=0, centroid=None):
    return list(iter_split_mesh_device(points, fv_dev, vf_dev, submesh_size, filter_patch_count, centroid))


def prefetch(iterable, depth=2):
    """Runs `iterable` on a helper thread, at most `depth` items ahead of the consumer; exceptions surface at the consumer."""
    import queue
    import threading
    q, end = queue.Queue(maxsize=depth), object()

    def work():
        try:
            for item in iterable:
                q.put((item, None))
            q.put((end, None))
        except BaseException as exc:                  # handed to the consumer, which re-raises it
            q.put((end, exc))

    threading.Thread(target=work, daemon=True).start()
    while True:
        item, exc = q.get()
        if exc is not None:
            raise exc
        if item is end:
            return
        yield item


class Stitcher:
    """Accumulates patch predictions into whole-mesh arrays and averages (test_dual.py:49-61)."""

    def __init__(self, n_vertices, n_faces, device):
        self.sum_v = torch.zeros((n_vertices, 1), dtype=torch.float32, device=device)
        self.vp = torch.zeros((n_vertices, 3), dtype=torch.float32, device=device)
        self.np_ = torch.zeros((n_faces, 3), dtype=torch.float32, device=device)

    def add(self, vert_p, norm_p, v_idx, f_idx):
        # indices are unique inside a patch, so these are plain scatters (index_add_ = one elementwise kernel each)
        v_idx, f_idx = torch.as_tensor(v_idx, device=self.vp.device), torch.as_tensor(f_idx, device=self.vp.device)
        self.sum_v.index_add_(0, v_idx, torch.ones((v_idx.numel(), 1), device=self.vp.device))
        self.vp.index_add_(0, v_idx, vert_p)
        self.np_.index_add_(0, f_idx, norm_p)

    def merge(self, other_sum_v, other_vp, other_np):
        self.sum_v += other_sum_v
        self.vp += other_vp
        self.np_ += other_np

    def finish(self):
        return self.vp / self.sum_v, torch.nn.functional.normalize(self.np_, dim=1)
