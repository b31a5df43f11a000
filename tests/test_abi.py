"""CPU: libgeobi.so loads without a GPU and exports exactly what include/geobi.h declares."""
import os
import re

import pytest

from tests import util
from geobi_gnn_b200 import _lib

HEADER = os.path.join(util.ROOT, "include", "geobi.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(geobi_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run `make -C geobi_gnn_b200/csrc` or __graft_entry__.build()"


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in geobi.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.geobi_version() >= 100


def test_queries_work_without_gpu():
    lib = _lib.load()
    assert lib.geobi_scan_ws_bytes(1000) > 0
    assert lib.geobi_csr_from_coo_ws_bytes(1000, 100, 0) > 1000 * 16
    assert lib.geobi_feast_fwd_ws_bytes(1000, 64, 32, 0) >= 1000 * 9 * 64 * 4
    assert lib.geobi_feast_bwd_ws_bytes(1000, 64, 32) >= 1000 * 9 * 64 * 8
    assert lib.geobi_feast_bwd_ws_bytes(1000, 129, 32) == 0
    assert lib.geobi_mlp_head_bwd_ws_bytes(1000, 32, 1024) >= 1000 * 1024 * 8          # a and dh planes
    assert lib.geobi_mlp_head_bwd_ws_bytes(1000, 64, 1024) == 0 and lib.geobi_mlp_head_bwd_ws_bytes(1000, 32, 1000) == 0
    assert lib.geobi_bfs_ws_bytes(1000) >= 1000 * 16 and lib.geobi_bfs_ws_bytes(-1) == 0


def test_ops_refuse_cpu_tensors():
    import torch
    from geobi_gnn_b200 import ops
    with pytest.raises(_lib.GeobiError):
        ops.exclusive_scan(torch.ones(4, dtype=torch.int32))
    with pytest.raises(_lib.GeobiError):
        ops.face_normal(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.long))


def test_module_surface_matches_reference():
    import torch
    from geobi_gnn_b200 import network, net_util, data_util
    net = network.DualGNN()
    want = util.oracle_net(0).state_dict()
    have = net.state_dict()
    assert list(have.keys()) == list(want.keys())
    assert all(have[k].shape == want[k].shape for k in want)
    for fn in ("loss_v", "loss_n", "dual_loss", "error_v", "error_n", "laplacian_loss"):
        assert callable(getattr(network, fn))
    for fn in ("calc_weight", "build_facet_graph", "build_vertex_graph", "build_edge_vf", "build_edge_fv", "computer_face_normal",
               "center_and_scale", "update_position", "update_position2"):
        assert callable(getattr(data_util, fn))
    assert hasattr(net_util, "PoolingLayer") and hasattr(net_util, "DualFusionLayer") and hasattr(net_util, "pool_edge")
    # PyG 1.x checkpoint layout is accepted
    conv = network.FeaStConv(6, 32, 9)
    sd = {"weight": torch.zeros(6, 288), "u": torch.zeros(6, 9), "c": torch.zeros(9), "bias": torch.zeros(32)}
    conv.load_state_dict(sd)
    assert conv.lin.weight.shape == (288, 6)


def test_data_container_lazy_attributes_and_csr_invalidation():
    """geobi_gnn_b200.data.Data: the PyG container contract (None removes a key, optional keys read as None) plus the lazy
    values PoolingLayer uses for the coarse edge_index / edge_weight, and the rule that re-assigning either drops `csr`."""
    import torch
    from geobi_gnn_b200.data import Data
    calls = []
    d = Data(x=torch.zeros(3, 2))
    assert d.edge_index is None and "edge_index" not in d
    d.set_lazy("edge_index", lambda: calls.append(1) or torch.ones(2, 4, dtype=torch.long))
    d.csr = object()
    assert "edge_index" in d and "<lazy>" in repr(d) and calls == []          # nothing evaluated yet
    assert d.edge_index.shape == (2, 4) and calls == [1]
    assert d.edge_index.shape == (2, 4) and calls == [1]                       # evaluated once
    assert "csr" in d
    d.edge_index = torch.zeros(2, 1, dtype=torch.long)                         # a new list invalidates the attached CSR
    assert "csr" not in d
    d.set_lazy("edge_weight", lambda: None)
    assert d.edge_weight is None and "edge_weight" not in d                    # a thunk returning None removes the key
    d.y = None
    assert "y" not in d and d.y is None
    c = d.clone()
    assert torch.equal(c.edge_index, d.edge_index) and c.x is not d.x
    with __import__("pytest").raises(AttributeError):
        d.not_a_key


def test_collate_keeps_the_coalesced_flag_and_size_buckets_are_stable():
    """Host-side logic that needs no GPU: the disjoint-union batch keeps `coalesced_undirected` only if every patch has it,
    fresh_view carries it along, and the capacity buckets of ops.valloc depend on the level's reference size only."""
    import torch
    from geobi_gnn_b200 import batching, ops
    from geobi_gnn_b200.data import Data

    def patch(n, flag):
        ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
        dv = Data(x=torch.zeros(n, 6), edge_index=ei, edge_weight=torch.ones(4))
        df = Data(x=torch.zeros(n, 6), edge_index=ei, edge_weight=torch.ones(4), fv_indices=torch.zeros(n, 3, dtype=torch.long))
        if flag:
            dv.coalesced_undirected = df.coalesced_undirected = True
        return dv, df

    dv, df, sl = batching.collate_dual([patch(3, True), patch(4, True)])
    assert dv.coalesced_undirected and df.coalesced_undirected
    assert dv.edge_index.tolist() == [[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]]      # offsets keep the list sorted
    assert sl["v"] == [(0, 3), (3, 7)] and df.fv_indices.shape == (7, 3)
    assert batching.fresh_view(dv).coalesced_undirected and batching.fresh_view(dv).edge_index is not dv.edge_index
    dv2, df2, _ = batching.collate_dual([patch(3, True), patch(4, False)])
    assert "coalesced_undirected" not in dv2 and "coalesced_undirected" not in df2
    assert batching.shard(list(range(7)), 1, 3) == [1, 4]
    # capacity buckets: sizes that differ by a few percent from one forward to the next land in the same bucket
    assert ops._cap(1000, 4096) == ops._cap(1010, 4096) == 1024
    assert ops._cap(5000, 4096) == 5000 and ops._cap(0, 4096) == 256 and ops._cap(7, None) == 7
    with ops.size_ref(4096):
        t = ops.valloc(1000, (3,), torch.float32, torch.device("cpu"))
    assert t.shape == (1000, 3) and t.untyped_storage().nbytes() == 1024 * 3 * 4


def test_obj_round_trip(tmp_path):
    """meshio: what om.read_trimesh / om.write_mesh do for the path - positions, faces, fan triangulation,
    `v/vt/vn` tokens and negative indices."""
    import numpy as np
    from geobi_gnn_b200 import meshio, synth
    p, f = synth.icosphere(2)
    path = tmp_path / "ico.obj"
    meshio.write_obj(path, p, f)
    p2, f2 = meshio.read_obj(path)
    assert np.array_equal(f2, f) and np.abs(p2 - p).max() < 1e-6
    m = synth.TriMesh(p2, f2)
    assert m.n_faces == 80 and m.n_vertices == 42
    quad = tmp_path / "quad.obj"
    quad.write_text("# a quad and a triangle\nvn 0 0 1\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 2 0 0\nvt 0 0\n"
                    "f 1/1/1 2/1/1 3/1/1 4/1/1\nf -4 -1 -3\n")
    pq, fq = meshio.read_obj(quad)
    assert pq.shape == (5, 3) and fq.tolist() == [[0, 1, 2], [0, 2, 3], [1, 4, 2]]
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with __import__("pytest").raises(ValueError):
        meshio.read_obj(bad)


def test_mesh_normalisation_is_the_oracles_bit_for_bit():
    """dataset.normalisation (threaded C++ edge lengths + numpy mean) against the oracle's numpy restatement of
    dataset.py:140,151-152: same centroid and the same fp32 scale, also above the threading threshold and for torch inputs."""
    import types
    import numpy as np
    import torch
    from geobi_gnn_b200 import dataset, synth
    from oracle import ref_dataset
    for n, seed in ((6, 0), (25, 1), (160, 2)):                       # 160 -> 768 000 edges: several threads
        p, f = synth.icosphere(n)
        rng = np.random.default_rng(seed)
        p = synth.add_normal_noise((p * rng.uniform(0.3, 50) + rng.normal(size=(1, 3)) * 7).astype(np.float32), f, 0.25, seed=seed)
        ev = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        ev = ev[ev[:, 0] < ev[:, 1]]
        holder = [types.SimpleNamespace()]
        ref_dataset.attach_normalisation(holder, p, ev)
        c, s = dataset.normalisation(p, ev)
        assert np.array_equal(c, holder[0].centroid.numpy()) and np.float32(s) == np.float32(holder[0].scale)
        c2, s2 = dataset.normalisation(torch.from_numpy(np.asarray(p, dtype=np.float32)), torch.from_numpy(ev))
        assert np.array_equal(c, c2) and s == s2


def test_native_obj_io_equals_the_python_specification(tmp_path):
    """csrc/host_obj.cpp (threaded .obj reader / writer in libgeobi_host.so) against meshio._read_obj_py / _write_obj_py: the same
    arrays bit for bit and the same bytes, for every thread count (chunk cuts land inside lines, between CR and LF, before
    relative indices) and for the record forms a mesh file can hold."""
    import filecmp
    import numpy as np
    import pytest
    from geobi_gnn_b200 import meshio, synth
    rng = np.random.default_rng(5)
    p, f = synth.icosphere(7)
    p = (p * rng.uniform(1e-4, 1e4, size=(p.shape[0], 1)) * rng.choice([-1.0, 1.0], size=p.shape)).astype(np.float32)
    p[3] = [0.0, -0.0, 1e-30]
    p[4] = [123456789.0, 1e20, -5e-7]
    for T in (1, 3, 8):
        meshio.write_obj(tmp_path / f"c{T}.obj", p, f, n_threads=T)
    meshio._write_obj_py(tmp_path / "py.obj", p, f)
    assert all(filecmp.cmp(tmp_path / f"c{T}.obj", tmp_path / "py.obj", shallow=False) for T in (1, 3, 8))
    want = meshio._read_obj_py(tmp_path / "py.obj")
    for T in (1, 2, 5, 16):
        got = meshio.read_obj(tmp_path / "py.obj", n_threads=T)
        assert got[0].dtype == np.float64 and got[1].dtype == np.int64
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[1], f)
    # a file with everything else in it: CRLF and lone CR line ends, tabs, signs, exponents, v/vt/vn tokens, relative indices,
    # polygons, a `v<TAB>` record (not a vertex for either reader), degenerate `f` records, comments, vn / vt / g / usemtl records, indented (ignored) records, no final newline
    lines = ["# comment", "mtllib x.mtl", "v 0 0 0", "v 1\t0 \t 0   ", "v\t7 7 7", "v +1.5e0 1E+0 -0.0 0.5 0.5 0.5", "vn 0 0 1", "vt 0.5 0.5", "v 0 1 .5", "v 2. 0 1e-3",
             "g group", "f 1 2 3", "f 1/1/1 2/1/1 3/1/1 4/1/1", "f -1 -2 -3", "f +1 2//1 -1 4/2", "f 1 2", "f 5", "  v 9 9 9", " f 1 2 3",
             "usemtl m", "v 5 5 5", "f -1 1 2 3 4 5", "s off"]
    for name, sep, tail in (("lf", "\n", "\n"), ("crlf", "\r\n", "\r\n"), ("cr", "\r", ""), ("mixed", "\n\n\r\n", "")):
        path = tmp_path / f"{name}.obj"
        path.write_bytes((sep.join(lines) + tail).encode())
        want = meshio._read_obj_py(path)
        assert want[0].shape == (6, 3) and want[1].shape[0] == 1 + 2 + 1 + 2 + 0 + 0 + 4
        for T in (1, 2, 3, 7, 32):
            got = meshio.read_obj(path, n_threads=T)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (name, T)
            assert np.array_equal(np.signbit(got[0]), np.signbit(want[0]))
    empty = tmp_path / "empty.obj"
    empty.write_bytes(b"")
    pe, fe = meshio.read_obj(empty)
    assert pe.shape == (0, 3) and fe.shape == (0, 3)
    for text in ("v 0 0 zero\nf 1 1 1\n", "v 0 0\n", "v 0 0 0\nf 1 x 1\n", "v 0 0 0\nf 1 2.5 1\n"):
        bad = tmp_path / "bad.obj"
        bad.write_text(text)
        with pytest.raises(ValueError):
            meshio.read_obj(bad)
        with pytest.raises((ValueError, IndexError)):
            meshio._read_obj_py(bad)
    with pytest.raises(OSError):
        meshio.write_obj(tmp_path / "no_such_dir" / "x.obj", p, f)


def test_host_library_exports_what_its_header_declares():
    """include/geobi_host.h <-> libgeobi_host.so (patch splitter, normalisation, .obj I/O): every declared entry point is
    exported, nothing else is, and the header compiles as C and as C++ next to the definitions' signatures."""
    import ctypes
    import subprocess
    header = os.path.join(util.ROOT, "include", "geobi_host.h")
    src = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(geobi_host_[a-z0-9_]+)\s*\(", src)))
    assert len(names) == 8
    path = os.path.join(util.ROOT, "geobi_gnn_b200", "libgeobi_host.so")
    lib = ctypes.CDLL(path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in geobi_host.h but not exported"
    exported = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    assert sorted(l.split()[-1] for l in exported.splitlines() if " T geobi_" in l) == names
    for compiler, lang in (("gcc", "c"), ("g++", "c++")):
        subprocess.run([compiler, "-x", lang, "-fsyntax-only", "-Wall", "-Werror", header], check=True)
    # the definitions agree with the declarations: compile both sources with the header force-included
    for cpp in ("host_patch.cpp", "host_obj.cpp"):
        subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-include", header, os.path.join(util.ROOT, "geobi_gnn_b200", "csrc", cpp)], check=True)


def test_native_obj_reader_fuzz_against_the_python_specification(tmp_path):
    """400 random line soups (valid and broken records, every line-end style, random thread counts): the native reader and the
    Python specification either return identical arrays or both refuse the file."""
    import numpy as np
    from geobi_gnn_b200 import meshio
    rng = np.random.default_rng(2024)
    numbers = ["0", "1", "-1", "+2", "3.5", "-0.0", "1e3", "1E-3", "+.5", "7.", "1e", "abc", "", "2.5", "nan", "inf", "-inf", "12345678901234567890"]
    indices = ["1", "2", "3", "4", "-1", "-2", "+1", "1/2", "2/1/3", "3//1", "//1", "x", "1.5", "0", "99"]
    seps = [" ", "  ", "\t", " \t "]
    ends = ["\n", "\r\n", "\r", "\n\n"]
    agree_ok = agree_fail = 0
    for case in range(400):
        lines = []
        for _ in range(int(rng.integers(0, 12))):
            kind = rng.choice(["v", "v", "v", "f", "f", "vn", "#", "", " v", "g"], p=None)
            if kind == "v":
                toks = [str(rng.choice(numbers[:10] if rng.random() < 0.9 else numbers)) for _ in range(int(rng.choice([3, 3, 3, 4, 6, 2])))]
                lines.append("v " + str(rng.choice(seps)).join(toks) + (" " if rng.random() < 0.2 else ""))
            elif kind == "f":
                toks = [str(rng.choice(indices[:9] if rng.random() < 0.9 else indices)) for _ in range(int(rng.integers(0, 6)))]
                lines.append("f " + str(rng.choice(seps)).join(toks))
            elif kind in ("vn", "g", " v"):
                lines.append(f"{kind} 1 2 3")
            else:
                lines.append(kind + (" note" if kind else ""))
        text = "".join(l + str(rng.choice(ends)) for l in lines)
        if lines and rng.random() < 0.3:
            text = text.rstrip("\r\n")
        path = tmp_path / "fuzz.obj"
        path.write_bytes(text.encode())
        try:
            want = meshio._read_obj_py(path)
        except (ValueError, IndexError, OverflowError):
            want = None
        for T in (1, int(rng.integers(2, 9))):
            try:
                got = meshio.read_obj(path, n_threads=T)
            except ValueError:
                got = None
            assert (got is None) == (want is None), (case, T, text)
            if got is not None:
                assert np.array_equal(got[0], want[0], equal_nan=True) and np.array_equal(got[1], want[1]), (case, T, text)
        agree_ok += want is not None
        agree_fail += want is None
    assert agree_ok > 100 and agree_fail > 30
