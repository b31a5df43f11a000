"""Wavefront .obj in / out - the two calls the reference makes through OpenMesh around the hot path:
`om.read_trimesh(path)` (dataset.py:134-135, 296-300; test_dual.py reads the noisy mesh) and `om.write_mesh(path, mesh)`
(test_dual.py:73 writes the denoised mesh).  Only what those calls use: vertex positions and faces; polygons are fan-
triangulated as OpenMesh's TriMesh reader does, texture / normal indices (`f v/vt/vn`) and negative (relative) indices are
accepted, everything else (`vn`, `vt`, groups, materials) is skipped.  Parsing and formatting run in libgeobi_host.so
(csrc/host_obj.cpp, threaded: 1 M faces read in 0.14 s and written in 0.17 s instead of 3.5 s / 4.4 s in Python); the plain-Python
versions below are the specification the native ones are tested against.  The arrays feed synth.TriMesh / topology.DeviceTriMesh.
"""
from __future__ import annotations

import numpy as np


def _threads():
    import os
    return min(16, os.cpu_count() or 1)


def read_obj(path, n_threads=None):
    """-> (points float64 [V,3], faces int64 [F,3]).  Parsed by libgeobi_host.so (csrc/host_obj.cpp: the file is cut at line
    boundaries and parsed by several threads, 3.5 s -> 0.1 s per million faces); same records, same numbers as `_read_obj_py`."""
    import ctypes as C
    from . import patches
    lib = patches._host()
    buf = np.fromfile(path, dtype=np.uint8)
    counts = np.zeros(2, dtype=np.int64)
    if n_threads is None:
        n_threads = _threads() if buf.size > (1 << 20) else 1
    lib.geobi_host_obj_count(patches._p(buf), C.c_int64(buf.size), C.c_int(n_threads), patches._p(counts))
    points = np.empty((int(counts[0]), 3), dtype=np.float64)
    fv = np.empty((int(counts[1]), 3), dtype=np.int64)
    bad = lib.geobi_host_obj_parse(patches._p(buf), C.c_int64(buf.size), C.c_int(n_threads), patches._p(points), patches._p(fv))
    if bad >= 0:
        line = bytes(buf[bad:bad + 80]).split(b"\n")[0].decode("utf-8", "replace").rstrip()
        raise ValueError(f"{path}: malformed record at byte {bad}: {line!r}")
    if fv.size and (fv.min() < 0 or fv.max() >= points.shape[0]):
        raise ValueError(f"{path}: face index out of range")
    return points, fv


def _read_obj_py(path):
    """The same reader in plain Python: the specification host_obj.cpp is tested against (tests/test_abi.py)."""
    pts, faces = [], []
    with open(path, "r") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                pts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(pts) + i)
                for k in range(1, len(idx) - 1):              # fan triangulation
                    faces.append((idx[0], idx[k], idx[k + 1]))
    points = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    fv = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    if fv.size and (fv.min() < 0 or fv.max() >= points.shape[0]):
        raise ValueError(f"{path}: face index out of range")
    return points, fv


def write_obj(path, points, faces, n_threads=None):
    """Positions with 6 significant decimals (OpenMesh's default stream precision), 1-based faces.  Formatted by
    libgeobi_host.so (threaded; byte-identical to `_write_obj_py`)."""
    import ctypes as C
    import os
    from . import patches
    p = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
    fv = np.ascontiguousarray(np.asarray(faces, dtype=np.int64).reshape(-1, 3))
    if n_threads is None:
        n_threads = _threads() if p.shape[0] + fv.shape[0] > 100000 else 1
    rc = patches._host().geobi_host_obj_write(os.fsencode(os.fspath(path)), patches._p(p), C.c_int64(p.shape[0]), patches._p(fv),
                                              C.c_int64(fv.shape[0]), C.c_int(n_threads))
    if rc != 0:
        raise OSError(f"cannot write {path}")


def _write_obj_py(path, points, faces):
    """The same writer in plain Python (specification / cross-check)."""
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    fv = np.asarray(faces, dtype=np.int64).reshape(-1, 3) + 1
    with open(path, "w") as f:
        f.write(f"# {p.shape[0]} vertices, {fv.shape[0]} faces\n")
        f.write("".join(f"v {a:.6g} {b:.6g} {c:.6g}\n" for a, b, c in p))
        f.write("".join(f"f {a} {b} {c}\n" for a, b, c in fv))


def denoise_obj(net, path_in, path_out, sub_size: int = 20000, data_type: str = "Synthetic", device="cuda", n_iter: int = 60):
    """test_dual.predict_one for one file: read -> predict_mesh (patch split, forward, stitch, 60-sweep vertex update) -> write.
    Returns (updated vertices [V,3], facet normals [F,3]) as numpy."""
    import torch
    from . import inference, synth, topology
    points, fv = read_obj(path_in)
    # whole-mesh index arrays on the device when there is one (the numpy stand-in costs ~1.5 s per million faces)
    mesh = topology.DeviceTriMesh(points, fv, device) if torch.device(device).type == "cuda" else synth.TriMesh(points, fv)
    V, Np, _ = inference.predict_mesh(net, mesh, sub_size, data_type=data_type, device=device, n_iter=n_iter)
    V = V.detach().cpu().numpy()
    write_obj(path_out, V, fv)
    return V, Np.detach().cpu().numpy()
