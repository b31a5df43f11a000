"""Patch-split inference (BASELINE configs[3] shape): C++ splitter vs the Python oracle (CPU), and the whole
predict_one pipeline — split, per-patch forward, stitch, de-normalise, 60-sweep update — vs the oracle (GPU)."""
import numpy as np
import pytest
import torch

from tests import util
from oracle import ref_data_util, ref_dataset, ref_patch


def test_splitter_matches_python_oracle():
    from geobi_gnn_b200 import patches
    mesh, _ = util.noisy_icosphere(9)
    for sub in (50, 333, 1000, 5000):
        a = patches.split_mesh(mesh.points, mesh.fv, mesh.vf, sub)
        b = ref_patch.split_mesh(mesh.points, mesh.fv, mesh.vf, sub)
        assert len(a) == len(b)
        covered = np.zeros(mesh.n_faces, bool)
        for (sel, seed), (sel2, seed2) in zip(a, b):
            assert seed == seed2 and np.array_equal(sel, np.asarray(sel2))
            covered[sel] = True
            va, fa = patches.get_submesh(mesh.fv, sel)
            vb, fb = ref_patch.get_submesh(mesh.fv, sel2)
            assert np.array_equal(va, vb) and np.array_equal(fa, fb)
            assert np.array_equal(va[fa], mesh.fv[sel])              # re-indexing round trip
        assert covered.all()
    ring = patches.mesh_get_neighbor_np(mesh.fv, mesh.vf, 7, ring_count=2)
    assert np.array_equal(ring, np.asarray(ref_patch.mesh_get_neighbor_np(mesh.fv, mesh.vf, 7, ring_count=2)))


@pytest.mark.gpu
def test_patched_inference_matches_oracle_pipeline():
    from geobi_gnn_b200 import inference, network, synth
    DEV = "cuda"
    mesh, _ = util.noisy_icosphere(10)
    sub = 700
    ref = util.oracle_net(0)
    util.set_perm_fn(ref, 21)
    # oracle: dataset.py:156-193 + test_dual.py:49-72
    pts32 = mesh.points.astype(np.float32)
    results, forced = [], []
    for sel, seed in ref_patch.split_mesh(mesh.points, mesh.fv, mesh.vf, sub):
        v_idx, faces = ref_patch.get_submesh(mesh.fv, sel)
        dd = ref_dataset.process_one_submesh(synth.TriMesh(mesh.points[v_idx], faces))
        ref_dataset.attach_normalisation(dd, pts32, mesh.ev)
        centroid, scale = dd[0].centroid, dd[0].scale
        dd = ref_dataset.post_processing(dd)
        with torch.no_grad():
            vp, nrm, _ = ref([dd[0], dd[1]])
        forced.append([[t[3] for t in pl.trace] for pl in util.poolings(ref)])
        results.append((vp, nrm, torch.from_numpy(v_idx), torch.tensor(sel)))
    Vp, Np = ref_dataset.stitch_patches(mesh.n_vertices, mesh.n_faces, results)
    Vp = Vp / scale + centroid
    want_V = ref_data_util.update_position2(Vp, torch.from_numpy(mesh.fv), torch.from_numpy(mesh.vf), Np, 60)
    assert len(results) >= 3
    mine = network.DualGNN().to(DEV).eval()
    mine.load_state_dict(ref.state_dict())
    V, Np_m, Vp_m = inference.predict_mesh(mine, mesh, sub, device=DEV, forced=forced)
    assert util.rel_err(Vp_m, Vp) < util.TOL_FP32
    assert util.rel_err(Np_m, Np) < util.TOL_NORMAL
    assert util.rel_err(V, want_V) < util.TOL_FP32
    # unsplit branch (n_faces <= sub_size)
    V1, N1, _ = inference.predict_mesh(mine, mesh, 10 ** 9, device=DEV)
    assert V1.shape == (mesh.n_vertices, 3) and torch.isfinite(V1).all()
    assert util.rel_err(N1.norm(dim=1), torch.ones(mesh.n_faces)) < 1e-5


@pytest.mark.gpu
def test_denoise_obj_file_to_file(tmp_path):
    """meshio.denoise_obj = read .obj -> predict_mesh -> write .obj (test_dual.py:25-87 for one file)."""
    import numpy as np
    import torch
    from geobi_gnn_b200 import inference, meshio, network, synth
    from tests import util
    torch.manual_seed(9)
    net = network.DualGNN().to("cuda").eval()
    for pl in util.poolings(net):
        pl.perm_fn = lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(n))
    p, f = synth.icosphere(5)
    pn = synth.add_normal_noise(p, f, 0.2, seed=3)
    src, dst = tmp_path / "noisy.obj", tmp_path / "out.obj"
    meshio.write_obj(src, pn, f)
    V, Np = meshio.denoise_obj(net, src, dst, sub_size=300, device="cuda")
    p_in, f_in = meshio.read_obj(src)
    want = inference.predict_mesh(net, synth.TriMesh(p_in, f_in), 300, device="cuda")
    assert np.abs(V - want[0].cpu().numpy()).max() < 1e-4 * max(1.0, np.abs(V).max())
    p_out, f_out = meshio.read_obj(dst)
    assert np.array_equal(f_out, f) and np.abs(p_out - V).max() < 1e-5 * max(1.0, np.abs(V).max())
    assert abs(np.linalg.norm(Np, axis=1) - 1).max() < 1e-5


def _open_shuffled_mesh(n, seed):
    """Icosphere with a cap of faces removed (boundary vertices, ragged vf rows) and faces / vertices renumbered at random."""
    from geobi_gnn_b200 import synth
    p, f = synth.icosphere(n)
    rng = np.random.default_rng(seed)
    f = f[p[f].mean(1)[:, 2] < 0.6]                                  # open the mesh
    used = np.unique(f)
    remap = np.full(p.shape[0], -1, dtype=np.int64)
    order = rng.permutation(used.shape[0])
    remap[used] = order
    p2 = np.empty((used.shape[0], 3), dtype=np.float32)
    p2[order] = p[used]
    f = remap[f][rng.permutation(f.shape[0])]
    return synth.TriMesh(synth.add_normal_noise(p2, f, 0.2, seed=seed).astype(np.float32), f)


def test_splitter_open_mesh_random_numbering_and_filter():
    """Boundary vertices (vf rows padded with -1 at different lengths), arbitrary numbering, and filter_patch_count
    (dataset.py:183): still the oracle's patches, in order."""
    from geobi_gnn_b200 import patches
    mesh = _open_shuffled_mesh(10, 3)
    assert (mesh.vf < 0).any()
    for sub, filt in ((200, 0), (777, 0), (300, 299), (10 ** 6, 0)):
        a = patches.split_mesh(mesh.points, mesh.fv, mesh.vf, sub, filt)
        b = ref_patch.split_mesh(mesh.points, mesh.fv, mesh.vf, sub, filt)
        assert len(a) == len(b) and len(a) >= 1
        for (sel, seed), (sel2, seed2) in zip(a, b):
            assert seed == seed2 and np.array_equal(sel, np.asarray(sel2))
    for seed in (0, 5, mesh.n_faces - 1):                           # ring-limited growth, and a scratch reused across calls
        st = [np.zeros(mesh.n_faces, np.uint32), np.zeros(mesh.n_vertices, np.uint32), 0]
        for rings in (1, 3):
            got = patches.mesh_get_neighbor_np(mesh.fv, mesh.vf, seed, ring_count=rings, _stamps=st)
            assert np.array_equal(got, np.asarray(ref_patch.mesh_get_neighbor_np(mesh.fv, mesh.vf, seed, ring_count=rings)))


def test_splitter_seed_arithmetic_is_numpys_bit_for_bit():
    """The seeds are arg-maxima of ((pts[fv].mean(1) - centroid)**2).sum(1) (dataset.py:165-166,192); on a near-sphere the
    winner is decided by the last bit, so the threaded C++ pass has to reproduce numpy's fp32 result exactly, and the
    next-seed scan np.argmax's first-of-equal-maxima rule (sizes above the threading threshold)."""
    import ctypes as C
    from geobi_gnn_b200 import patches, synth
    lib = patches._host()
    lib.geobi_host_cover_next_seed.restype = C.c_int64
    p, f = synth.icosphere(130)                                     # 338 000 faces: several threads
    rng = np.random.default_rng(0)
    p = (p * 37.5 + rng.normal(size=(1, 3)) * 4).astype(np.float32)
    f = np.ascontiguousarray(f, dtype=np.int64)
    cen = np.ascontiguousarray(p.mean(0, keepdims=True))
    want = ((p[f].mean(1) - cen) ** 2).sum(1)
    for nthr in (1, 3, 8):
        got = np.empty(f.shape[0], dtype=np.float32)
        lib.geobi_host_face_d2(patches._p(p), patches._p(f), C.c_int64(f.shape[0]), patches._p(cen), patches._p(got), C.c_int(nthr))
        assert want.dtype == np.float32 and np.array_equal(want, got)
    d2 = np.round(want * 4) / 4                                     # many exact ties
    d2 = d2.astype(np.float32)
    covered = np.zeros(f.shape[0], bool)
    left = C.c_int64(f.shape[0])
    for it in range(6):
        sel = np.ascontiguousarray(rng.choice(f.shape[0], 50000, replace=False).astype(np.int64))
        if it == 5:
            sel = np.arange(f.shape[0], dtype=np.int64)             # everything covered -> -1
        nxt = lib.geobi_host_cover_next_seed(patches._p(d2), C.c_int64(f.shape[0]), patches._p(sel), C.c_int64(sel.shape[0]), C.byref(left),
                                             C.c_int(1 + it))
        covered[sel] = True
        assert left.value == int((~covered).sum())
        if covered.all():
            assert nxt == -1
        else:
            ref = np.where(covered, -np.inf, np.round(want * 4) / 4)
            assert nxt == int(np.argmax(ref)) and not covered[nxt]


def test_streaming_splitter_yields_the_same_patches():
    """patches.iter_split_mesh (generator) and patches.prefetch (helper thread, bounded look-ahead) hand out exactly the list
    split_mesh returns, in order; an exception inside the producer reaches the consumer."""
    from geobi_gnn_b200 import patches
    mesh = _open_shuffled_mesh(12, 3)
    want = patches.split_mesh(mesh.points, mesh.fv, mesh.vf, 200, 60)
    assert len(want) > 5
    it = patches.iter_split_mesh(mesh.points, mesh.fv, mesh.vf, 200, 60)
    first = next(it)
    assert np.array_equal(first[0], want[0][0]) and first[1] == want[0][1]
    rest = list(it)
    assert len(rest) == len(want) - 1 and all(np.array_equal(a[0], b[0]) and a[1] == b[1] for a, b in zip(rest, want[1:]))
    got = list(patches.prefetch(patches.iter_split_mesh(mesh.points, mesh.fv, mesh.vf, 200, 60), depth=2))
    assert len(got) == len(want) and all(np.array_equal(a[0], b[0]) and a[1] == b[1] for a, b in zip(got, want))
    assert list(patches.prefetch(iter(()))) == []

    def broken():
        yield 1
        raise KeyError("inside the producer")
    with pytest.raises(KeyError):
        list(patches.prefetch(broken()))


@pytest.mark.gpu
def test_get_submesh_device_equals_host_routine():
    """The device cut-out (patches.get_submesh_device) reproduces data_util.get_submesh: vertices in first-appearance order, faces
    re-indexed - on BFS patches (discovery order) and on a shuffled face selection."""
    import numpy as np
    from geobi_gnn_b200 import patches, synth
    p, f = synth.icosphere(12)
    mesh = synth.TriMesh(p, f)
    fv_dev = torch.from_numpy(mesh.fv).cuda()
    rng = np.random.default_rng(3)
    sels = [patches.mesh_get_neighbor_np(mesh.fv, mesh.vf, int(s), neighbor_count=700) for s in (0, 123, 2000)]
    sels.append(rng.permutation(f.shape[0])[:900].astype(np.int64))
    for sel in sels:
        v_host, faces_host = patches.get_submesh(mesh.fv, sel)
        v_dev, faces_dev = patches.get_submesh_device(fv_dev, torch.from_numpy(np.ascontiguousarray(sel)).cuda(), mesh.n_vertices)
        assert np.array_equal(v_dev.cpu().numpy(), v_host)
        assert np.array_equal(faces_dev.cpu().numpy(), faces_host)


@pytest.mark.gpu
def test_device_splitter_yields_the_host_splitters_patches():
    """geobi_bfs_begin / geobi_bfs_grow (ring-parallel BFS, atomicMin claims, ordered compaction) against the host C++ splitter -
    itself equal to the Python oracle (tests above): same seeds, same faces in the same discovery order, for closed and open /
    shuffled meshes, patch sizes from a handful of faces to more than the mesh, with and without filter_patch_count."""
    from geobi_gnn_b200 import patches
    DEV = "cuda"
    cases = [(util.noisy_icosphere(9)[0], ((2, 0), (50, 0), (333, 0), (1000, 0), (5000, 0))),
             (_open_shuffled_mesh(10, 3), ((200, 0), (777, 0), (300, 299), (10 ** 6, 0))),
             (util.noisy_icosphere(40)[0], ((4000, 0), (31999, 0)))]
    for mesh, subs in cases:
        fv_d, vf_d = torch.from_numpy(np.ascontiguousarray(mesh.fv)).to(DEV), torch.from_numpy(np.ascontiguousarray(mesh.vf)).to(DEV)
        for sub, filt in subs:
            want = patches.split_mesh(mesh.points, mesh.fv, mesh.vf, sub, filt)
            got = patches.split_mesh_device(mesh.points, fv_d, vf_d, sub, filt)
            assert len(got) == len(want) and len(got) >= 1, (sub, len(got), len(want))
            for k, ((sel, seed), (sel2, seed2)) in enumerate(zip(got, want)):
                assert seed == seed2, (sub, k)
                assert sel.dtype == torch.int32 and np.array_equal(sel.cpu().numpy(), np.asarray(sel2)), (sub, k)
    # points already on the device, centroid handed over
    mesh = cases[0][0]
    pts_d = torch.from_numpy(mesh.points.astype(np.float32)).to(DEV)
    cen = mesh.points.astype(np.float32).mean(0)
    got = patches.split_mesh_device(pts_d, torch.from_numpy(mesh.fv).to(DEV), torch.from_numpy(mesh.vf).to(DEV), 1000, centroid=cen)
    want = patches.split_mesh(mesh.points, mesh.fv, mesh.vf, 1000)
    assert [s for _, s in got] == [s for _, s in want]
    assert all(np.array_equal(a.cpu().numpy(), np.asarray(b)) for (a, _), (b, _) in zip(got, want))


@pytest.mark.gpu
def test_predict_mesh_with_device_partition_and_normalisation():
    """The all-device whole-mesh path (device BFS partition, device normalisation, no host copy of the index arrays: what bench.py's
    configs[3] runs) against the default path of the same function (host splitter, numpy normalisation) on a device-resident mesh."""
    from geobi_gnn_b200 import inference, network, topology
    DEV = "cuda"
    mesh_h, _ = util.noisy_icosphere(12)
    mesh = topology.DeviceTriMesh(mesh_h.points, mesh_h.fv, DEV)
    torch.manual_seed(1)
    net = network.DualGNN().to(DEV).eval()
    sub = 900
    util.set_perm_fn(net, 3)
    want = inference.predict_mesh(net, mesh, sub, device=DEV, return_parts=True)
    norm, cen = inference.device_normalisation(mesh)
    parts = inference.partition(mesh, sub, centroid=cen)
    assert all(torch.is_tensor(sel) and sel.is_cuda for sel, _ in parts)
    host_parts = inference.partition(mesh, sub, host=inference.host_views(mesh))
    assert [s for _, s in parts] == [s for _, s in host_parts]
    assert all(np.array_equal(a.cpu().numpy(), np.asarray(b)) for (a, _), (b, _) in zip(parts, host_parts))
    util.set_perm_fn(net, 3)
    got = inference.predict_mesh(net, mesh, sub, device=DEV, parts=parts, norm=norm, return_parts=True)
    assert got[3] == want[3] == len(parts)
    assert util.rel_err(got[0], want[0]) < util.TOL_FP32 and util.rel_err(got[2], want[2]) < util.TOL_FP32
    assert util.rel_err(got[1], want[1]) < util.TOL_NORMAL
