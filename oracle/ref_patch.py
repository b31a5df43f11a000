"""CPU oracle: patch growth / submesh extraction / mesh splitting.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/code/data_util.py:55-84 (``mesh_get_neighbor_np``), :318-336 (``get_submesh``) and the
splitting loop of /root/reference/code/dataset.py:156-193 in pure Python / numpy (small cases only).  Pinned: the reference's own
functions produce the same patches (tests/golden/reference_pipeline_ico8.npz, tests/test_reference_golden.py).
"""
from __future__ import annotations

import numpy as np


def mesh_get_neighbor_np(fv_indices, vf_indices, seed_idx, neighbor_count=None, ring_count=None):
    """Ring-by-ring BFS over faces sharing a vertex; stops as soon as `neighbor_count` faces are collected.
    Order of discovery: faces of the current ring in collection order, their 3 vertices in fv order, the incident
    faces of each vertex in vf order (stopping at the first -1 pad)."""
    assert neighbor_count is not None or ring_count is not None
    big = 1 << 62
    neighbor_count = big if neighbor_count is None else neighbor_count
    ring_count = big if ring_count is None else ring_count
    taken = np.zeros(fv_indices.shape[0], dtype=bool)
    out = [int(seed_idx)]
    taken[seed_idx] = True
    lo, hi = 0, 1
    ring = 0
    while ring < ring_count:
        for face in out[lo:hi]:
            for v in fv_indices[face]:
                for g in vf_indices[v]:
                    if g < 0:
                        break
                    if not taken[g]:
                        out.append(int(g))
                        taken[g] = True
                        if len(out) >= neighbor_count:
                            return out
        lo, hi = hi, len(out)
        if lo == hi:
            return out
        ring += 1
    return out


def get_submesh(fv_indices, select_faces):
    """data_util.py:318-336 — V_idx = original vertex ids in order of FIRST APPEARANCE when the selected faces'
    corners are scanned in order; F = the selected faces re-indexed into V_idx."""
    flat = fv_indices[np.asarray(select_faces)].reshape(-1)
    slot = {}
    v_idx, f = [], np.zeros(flat.shape[0], dtype=np.int32)
    for i, v in enumerate(flat.tolist()):
        k = slot.get(v)
        if k is None:
            k = slot[v] = len(v_idx)
            v_idx.append(v)
        f[i] = k
    return np.array(v_idx), f.reshape(-1, 3)


def split_mesh(points, fv_indices, vf_indices, submesh_size, filter_patch_count=0):
    """dataset.py:156-193 — list of (select_faces, seed): seeds are the unvisited face farthest from the centroid
    (np.argmax: first on ties), patches may overlap on faces already visited by an earlier patch's growth."""
    centroid = points.astype(np.float32).mean(0, keepdims=True)
    pts = points.astype(np.float32)
    cent = pts[fv_indices].mean(1)
    d2 = ((cent - centroid) ** 2).sum(1)
    flag = np.zeros(fv_indices.shape[0], dtype=bool)
    seed = int(np.argmax(d2))
    patches = []
    while True:
        sel = mesh_get_neighbor_np(fv_indices, vf_indices, seed, neighbor_count=submesh_size)
        flag[np.asarray(sel)] = True
        if len(sel) > filter_patch_count:
            patches.append((sel, seed))
        left = np.where(~flag)[0]
        if left.size == 0:
            break
        seed = int(left[np.argmax(d2[left])])
    return patches
