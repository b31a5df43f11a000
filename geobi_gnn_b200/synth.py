"""Synthetic noisy meshes + halfedge-free topology arrays (host side, numpy).

The reference ships no mesh data (``/root/reference/.gitignore:3-5``) and reads
its ``.obj`` files through OpenMesh (``/root/reference/code/dataset.py:134-137``,
``:198-204``).  Every config in BASELINE.json is "synthetic noisy icosphere /
patches", so this module provides

* ``icosphere(n)``          geodesic icosphere of frequency n: F = 20 n^2,
                            V = 10 n^2 + 2;
* ``add_normal_noise``      p += sigma * mean_edge_len * N(0,1) * n_v;
* ``TriMesh``               the index arrays the reference pulls out of OpenMesh
                            (``ev/fv/vf/vv_indices``, face / vertex normals),
                            restated from the OpenMesh defaults (SURVEY.md
                            section 8a row A0).  ``vf`` / ``vv`` rows are sorted
                            ascending and padded with -1; OpenMesh emits them in
                            circulation order, but every consumer on the hot
                            path is order-insensitive (set union + coalesce in
                            ``build_facet_graph``; a sum in ``update_position2``).

Nothing here runs on the GPU: it is the input builder for tests and bench.py.
"""
from __future__ import annotations

import numpy as np

__all__ = ["icosphere", "add_normal_noise", "TriMesh", "mean_edge_length"]


def _icosahedron():
    phi = (1.0 + 5.0 ** 0.5) / 2.0
    v = []
    for a in (-1.0, 1.0):
        for b in (-phi, phi):
            v += [(0.0, a, b), (a, b, 0.0), (b, 0.0, a)]
    v = np.asarray(v, dtype=np.float64)
    # faces = all vertex triples at mutual distance 2 (edge length), oriented outward
    d2 = ((v[:, None, :] - v[None, :, :]) ** 2).sum(-1)
    adj = np.abs(d2 - 4.0) < 1e-9
    faces = []
    for a in range(12):
        for b in range(a + 1, 12):
            if not adj[a, b]:
                continue
            for c in range(b + 1, 12):
                if adj[a, c] and adj[b, c]:
                    n = np.cross(v[b] - v[a], v[c] - v[a])
                    faces.append((a, b, c) if np.dot(n, v[a] + v[b] + v[c]) > 0 else (a, c, b))
    faces = np.asarray(faces, dtype=np.int64)
    assert faces.shape == (20, 3)
    return v / np.linalg.norm(v, axis=1, keepdims=True), faces


def icosphere(n: int):
    """Unit geodesic icosphere.  Returns (points float64 [V,3], faces int64 [F,3]).

    Vertex ids are canonical (12 corners, then the 30 base edges' interior
    points, then each base face's interior points row-major), so shared points
    are never duplicated and neighbouring vertices inside a base face are close
    in index (gather locality)."""
    assert n >= 1
    bv, bf = _icosahedron()
    # base edges (a<b) -> id
    eid = {}
    for f in bf:
        for k in range(3):
            a, b = int(f[k]), int(f[(k + 1) % 3])
            key = (min(a, b), max(a, b))
            if key not in eid:
                eid[key] = len(eid)
    assert len(eid) == 30
    n_e = n - 1
    n_i = (n - 1) * (n - 2) // 2
    V = 12 + 30 * n_e + 20 * n_i
    pts = np.zeros((V, 3), dtype=np.float64)
    I, J = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    valid = (I + J) <= n
    all_faces = []

    def edge_ids(x, y, t):
        # ids of the points at parameter t (1..n-1, measured from x) on base edge x-y
        a, b = (x, y) if x < y else (y, x)
        k = t if x < y else n - t
        return 12 + eid[(a, b)] * n_e + (k - 1)

    for fi, (A, B, C) in enumerate(bf):
        A, B, C = int(A), int(B), int(C)
        G = np.full((n + 1, n + 1), -1, dtype=np.int64)
        # interior: i>=1, j>=1, i+j<=n-1, row-major in (i, j)
        inter = (I >= 1) & (J >= 1) & ((I + J) <= n - 1)
        G[inter] = 12 + 30 * n_e + fi * n_i + np.arange(n_i)
        if n >= 2:
            t = np.arange(1, n)
            G[t, 0] = edge_ids(A, B, t)          # j = 0: A -> B, parameter i
            G[0, t] = edge_ids(A, C, t)          # i = 0: A -> C, parameter j
            G[n - t, t] = edge_ids(B, C, t)      # i + j = n: B -> C, parameter j
        G[0, 0], G[n, 0], G[0, n] = A, B, C
        w = (n - I - J)[valid][:, None] * bv[A] + I[valid][:, None] * bv[B] + J[valid][:, None] * bv[C]
        pts[G[valid]] = w / np.linalg.norm(w, axis=1, keepdims=True)
        up = (I + J) <= n - 1
        iu, ju = I[up], J[up]
        ups = np.stack([G[iu, ju], G[iu + 1, ju], G[iu, ju + 1]], 1)
        dn = (I + J) <= n - 2
        i_d, j_d = I[dn], J[dn]
        dns = np.stack([G[i_d + 1, j_d], G[i_d + 1, j_d + 1], G[i_d, j_d + 1]], 1)
        tri = np.concatenate([ups, dns], 0)
        ki = np.concatenate([iu, i_d])
        kj = np.concatenate([ju, j_d])
        kt = np.concatenate([np.zeros_like(iu), np.ones_like(i_d)])
        order = np.lexsort((kt, kj, ki))
        all_faces.append(tri[order])
    faces = np.concatenate(all_faces, 0)
    assert faces.shape[0] == 20 * n * n and faces.min() == 0 and faces.max() == V - 1
    return pts, faces


def mean_edge_length(points: np.ndarray, ev: np.ndarray) -> float:
    d = points[ev[:, 0]] - points[ev[:, 1]]
    return float(np.sqrt((d * d).sum(1)).mean())


class TriMesh:
    """Index arrays of a triangle mesh, as the reference reads them from OpenMesh.

    Attributes (numpy):
      points [V,3] float64, faces/fv [F,3] int64, ev [E,2] int64 (edge creation
      order = first appearance scanning faces, orientation = first halfedge),
      vf [V,maxval] int64 padded -1, vv [V,maxval] int64 padded -1,
      face_normals [F,3] float64 = normalize(cross(p1-p0, p2-p0)),
      vertex_normals [V,3] float64 = normalize(sum of incident face normals).
    """

    def __init__(self, points: np.ndarray, faces: np.ndarray):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.fv = np.ascontiguousarray(faces, dtype=np.int64)
        V, F = self.points.shape[0], self.fv.shape[0]
        self.n_vertices, self.n_faces = V, F
        # --- edges
        h0 = self.fv.reshape(-1)
        h1 = self.fv[:, [1, 2, 0]].reshape(-1)
        key = np.minimum(h0, h1) * V + np.maximum(h0, h1)
        _, first = np.unique(key, return_index=True)
        first.sort()
        self.ev = np.stack([h0[first], h1[first]], 1)
        # --- vertex -> incident faces (ascending face id), padded
        flat = self.fv.reshape(-1)
        order = np.argsort(flat, kind="stable")
        vs, fs = flat[order], order // 3
        cnt = np.bincount(vs, minlength=V)
        start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        slot = np.arange(vs.shape[0]) - start[vs]
        self.vf = np.full((V, int(cnt.max()) if V else 0), -1, dtype=np.int64)
        self.vf[vs, slot] = fs
        # --- vertex -> neighbour vertices (ascending), padded
        a = np.concatenate([self.ev[:, 0], self.ev[:, 1]])
        b = np.concatenate([self.ev[:, 1], self.ev[:, 0]])
        o = np.lexsort((b, a))
        a, b = a[o], b[o]
        cnt = np.bincount(a, minlength=V)
        start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        slot = np.arange(a.shape[0]) - start[a]
        self.vv = np.full((V, int(cnt.max()) if V else 0), -1, dtype=np.int64)
        self.vv[a, slot] = b
        self.update_normals()

    def update_normals(self):
        p = self.points[self.fv]
        n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
        self.face_normals = n / np.clip(np.linalg.norm(n, axis=1, keepdims=True), 1e-300, None)
        vn = np.zeros_like(self.points)
        for k in range(3):
            np.add.at(vn, self.fv[:, k], self.face_normals)
        self.vertex_normals = vn / np.clip(np.linalg.norm(vn, axis=1, keepdims=True), 1e-300, None)


def add_normal_noise(points, faces, sigma=0.2, seed=0):
    """Noisy copy: p += sigma * mean_edge_len * N(0,1) * vertex_normal."""
    m = TriMesh(points, faces)
    rng = np.random.default_rng(seed)
    amp = rng.standard_normal(m.n_vertices) * sigma * mean_edge_length(m.points, m.ev)
    return m.points + amp[:, None] * m.vertex_normals
