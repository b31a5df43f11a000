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
    assert util.rel_err(Vp_m, Vp) < 5e-5
    assert util.rel_err(Np_m, Np) < 2e-4
    assert util.rel_err(V, want_V) < 5e-5
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
