"""The reference's OWN code as the yardstick.  tests/golden/reference_ico{3,5}.npz were produced by executing
/root/reference/code/{dataset,data_util,net_util,network}.py unmodified (tests/golden/make_reference_golden.py; only the
un-installable third-party packages are stood in for).  CPU: the oracle's restatement of those files reproduces the vectors.
GPU (`-m gpu`): the CUDA path reproduces them, teacher-forced with the reference's cluster labels.  Neither reads
/root/reference at run time, except the regeneration check, which is skipped where the reference is absent."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import util
from oracle import ref_data_util, ref_dataset, ref_network
from geobi_gnn_b200 import synth

CASES = {3: dict(weight_seed=0, data_type="Synthetic", wei_param=2), 5: dict(weight_seed=7, data_type="Kinect_v1", wei_param=10)}
FLOAT_TOL = 1e-6          # CPU oracle vs the reference on another host's CPU: same operations, possibly another vector width


def _golden(n):
    return np.load(os.path.join(util.GOLDEN, f"reference_ico{n}.npz"))


def _forced(g):
    """Reference call order (network.py:318-343): vertex pooling 1, 2, facet pooling 1, 2 - two matching steps each."""
    assert int(g["n_poolings"]) == 8
    labels = [torch.from_numpy(g[f"labels_{i}"]) for i in range(8)]
    return [labels[0:2], labels[2:4], labels[4:6], labels[6:8]]


def _oracle_net(case):
    torch.manual_seed(case["weight_seed"])
    net = ref_network.DualGNN(force_depth=case["data_type"] != "Synthetic", pool_type="max", wei_param=case["wei_param"])
    return net.eval()


@pytest.mark.parametrize("n", sorted(CASES))
def test_oracle_reproduces_the_reference(n):
    g, case = _golden(n), CASES[n]
    # same constructor order and initialisers: seeded weights are the reference's, bit for bit
    net = _oracle_net(case)
    h = hashlib.sha256()
    for k, v in net.state_dict().items():
        h.update(k.encode())
        h.update(v.numpy().tobytes())
    assert np.array_equal(np.frombuffer(h.digest(), dtype=np.uint8), g["state_sha"])
    # dataset.py:140-153,196-269 - graphs bit-exact, features to FLOAT_TOL
    mesh_n, mesh_o = synth.TriMesh(g["points_noisy"], g["faces"]), synth.TriMesh(g["points_original"], g["faces"])
    dd = ref_dataset.process_one_submesh(mesh_n, "g", mesh_o)
    assert np.array_equal(dd[0].edge_index.numpy(), g["raw_v_edge_index"]) and np.array_equal(dd[1].edge_index.numpy(), g["raw_f_edge_index"])
    assert np.array_equal(dd[0].edge_dual.numpy(), g["edge_dual_v"]) and np.array_equal(dd[1].edge_dual.numpy(), g["edge_dual_f"])
    assert util.rel_err(dd[0].edge_weight, g["raw_v_edge_weight"]) < FLOAT_TOL and util.rel_err(dd[1].edge_weight, g["raw_f_edge_weight"]) < FLOAT_TOL
    ref_dataset.attach_normalisation(dd, g["points_noisy"], mesh_n.ev)
    assert util.rel_err(dd[0].centroid.reshape(-1), g["centroid"].reshape(-1)) < FLOAT_TOL and abs(dd[0].scale / g["scale"] - 1) < FLOAT_TOL
    centroid, scale = dd[0].centroid.clone(), dd[0].scale
    dv, df = ref_dataset.post_processing(dd, case["data_type"])
    for t, k in ((dv.x, "x_v"), (dv.y, "y_v"), (df.x, "x_f"), (df.y, "y_f")):
        assert util.rel_err(t, g[k]) < FLOAT_TOL, k
    assert ("depth_direction" in dv) == (case["data_type"] != "Synthetic")
    y_v, y_f = dv.y, df.y
    # network.py:254-343 + net_util.py:56-302, teacher-forced with the reference's clusters AND free-running on its seed
    for mode in ("forced", "seeded"):
        net = _oracle_net(case)
        if mode == "forced":
            for pl, f in zip(util.poolings(net), _forced(g)):
                pl.forced = f
        else:
            util.set_perm_fn(net, 1000 + case["weight_seed"])
        with torch.no_grad():
            vp, nrm, third = net([_copy(dv), _copy(df)])
        assert third is None
        assert util.rel_err(vp, g["vert_p"]) < 1e-5 and util.rel_err(nrm, g["norm_p"]) < 1e-5
        if mode == "seeded":
            got = [t[3] for pl in util.poolings(net) for t in pl.trace]
            assert all(np.array_equal(a.numpy(), g[f"labels_{i}"]) for i, a in enumerate(got))
    # losses, errors (network.py:347-413) and the 60-sweep vertex update (data_util.py:529-556)
    vp, nrm = torch.from_numpy(g["vert_p"]), torch.from_numpy(g["norm_p"])
    for name, val in (("loss_v_L1", ref_network.loss_v(vp, y_v, "L1")), ("loss_v_L2", ref_network.loss_v(vp, y_v, "L2")),
                      ("loss_n_L1", ref_network.loss_n(nrm, y_f, "L1")), ("loss_n_L2", ref_network.loss_n(nrm, y_f, "L2")),
                      ("error_v", ref_network.error_v(vp, y_v)), ("error_n", ref_network.error_n(nrm, y_f)),
                      ("dual_loss", ref_network.dual_loss(ref_network.loss_v(vp, y_v, "L1"), ref_network.loss_n(nrm, y_f, "L1"), 2.0, 0.5))):
        assert abs(float(val) / float(g[name]) - 1) < 1e-5, name
    fv, vf = torch.from_numpy(mesh_n.fv), torch.from_numpy(mesh_n.vf)
    depth = torch.nn.functional.normalize(torch.from_numpy(g["points_noisy"]), dim=1) if case["data_type"] != "Synthetic" else None
    V = ref_data_util.update_position2(vp / scale + centroid, fv, vf, nrm, 60, depth_direction=depth)
    assert util.rel_err(V, g["updated_vertices"]) < 1e-5
    assert util.rel_err(ref_data_util.computer_face_normal(V, fv), g["updated_normals"]) < 1e-4


def _copy(d):
    out = type(d)()
    for k in d.keys():
        v = getattr(d, k)
        setattr(out, k, v.clone() if torch.is_tensor(v) else v)
    return out


@pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="the reference tree exists only in the build container")
def test_vectors_regenerate_from_the_reference(tmp_path):
    """The committed .npz files are what the generating script produces from /root/reference today."""
    script = os.path.join(util.GOLDEN, "make_reference_golden.py")
    code = ("import sys, runpy, numpy as np; sys.argv=['x']; m = runpy.run_path(%r); ref = m['import_reference']();\n"
            "out, _ = m['reference_case'](ref, **m['CASES'][0]); np.savez(%r, **out)" % (script, str(tmp_path / "again.npz")))
    subprocess.check_call([sys.executable, "-W", "ignore", "-c", code], cwd=util.ROOT)
    again, g = np.load(tmp_path / "again.npz"), _golden(3)
    assert sorted(again.files) == sorted(k for k in g.files if k != "case")
    for k in again.files:                     # integers (graphs, labels, digest) exactly; floats to rounding (thread count may differ)
        if np.issubdtype(again[k].dtype, np.floating):
            assert util.rel_err(again[k], g[k]) < (1e-4 if k == "updated_normals" else 1e-5), k      # normals of sliver faces amplify
        else:
            assert np.array_equal(again[k], g[k]), k


# ------------------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("n", sorted(CASES))
@pytest.mark.parametrize("front_end", ["host", "device"])
def test_cuda_path_reproduces_the_reference(n, front_end):
    """Mesh -> graphs -> DualGNN forward -> losses -> vertex update through the C-ABI kernels against the reference's vectors:
    index arrays bit-exact, floats within 1e-5 max-norm relative (BASELINE.json north_star); unit normals as vectors within util.TOL_NORMAL."""
    from geobi_gnn_b200 import data_util, dataset, network, topology
    DEV = "cuda"
    g, case = _golden(n), CASES[n]
    make = synth.TriMesh if front_end == "host" else (lambda p, f: topology.DeviceTriMesh(p, f, DEV))
    mesh_n, mesh_o = make(g["points_noisy"], g["faces"]), make(g["points_original"], g["faces"])
    dd = dataset.process_one_submesh(mesh_n, "g", mesh_o, DEV)
    assert np.array_equal(dd[0].edge_index.cpu().numpy(), g["raw_v_edge_index"])
    assert np.array_equal(dd[1].edge_index.cpu().numpy(), g["raw_f_edge_index"])
    assert np.array_equal(dd[0].edge_dual.cpu().numpy(), g["edge_dual_v"]) and np.array_equal(dd[1].edge_dual.cpu().numpy(), g["edge_dual_f"])
    assert util.rel_err(dd[0].edge_weight, g["raw_v_edge_weight"]) < 1e-5 and util.rel_err(dd[1].edge_weight, g["raw_f_edge_weight"]) < 1e-5
    dataset.attach_normalisation(dd, g["points_noisy"], synth.TriMesh(g["points_noisy"], g["faces"]).ev)
    assert util.rel_err(dd[0].centroid.reshape(-1), g["centroid"].reshape(-1)) < FLOAT_TOL and abs(dd[0].scale / g["scale"] - 1) < FLOAT_TOL
    centroid, scale = dd[0].centroid.clone(), dd[0].scale
    dv, df = dataset.post_processing(dd, case["data_type"])
    for t, k in ((dv.x, "x_v"), (dv.y, "y_v"), (df.x, "x_f"), (df.y, "y_f")):
        assert util.rel_err(t, g[k]) < 1e-5, k
    y_v, y_f = dv.y, df.y
    net = network.DualGNN(force_depth=case["data_type"] != "Synthetic", pool_type="max", wei_param=case["wei_param"]).to(DEV).eval()
    net.load_state_dict(_oracle_net(case).state_dict())              # == the reference's seeded weights (state_sha, CPU test)
    for pl, f in zip(util.poolings(net), _forced(g)):
        pl.forced = f
    with torch.no_grad():
        vp, nrm, third = net([dv, df])
    assert third is None
    assert util.rel_err(vp, g["vert_p"]) < util.TOL_FP32
    assert util.rel_err(nrm, g["norm_p"]) < util.TOL_NORMAL
    for name, val in (("loss_v_L1", network.loss_v(vp, y_v, "L1")), ("loss_n_L1", network.loss_n(nrm, y_f, "L1")),
                      ("error_v", network.error_v(vp, y_v)), ("error_n", network.error_n(nrm, y_f)),
                      ("dual_loss", network.dual_loss(network.loss_v(vp, y_v, "L1"), network.loss_n(nrm, y_f, "L1"), 2.0, 0.5))):
        assert abs(float(val) / float(g[name]) - 1) < 1e-4, name
    # vertex update from the REFERENCE's network outputs (so this step is compared on its own)
    fv, vf = torch.from_numpy(g["faces"]).to(DEV), torch.as_tensor(mesh_n.vf).to(DEV)
    depth = torch.nn.functional.normalize(torch.from_numpy(g["points_noisy"]).to(DEV), dim=1) if case["data_type"] != "Synthetic" else None
    Vp = torch.from_numpy(g["vert_p"]).to(DEV) / scale + centroid
    V = data_util.update_position2(Vp, fv, vf, torch.from_numpy(g["norm_p"]).to(DEV), 60, depth_direction=depth)
    assert util.rel_err(V, g["updated_vertices"]) < 1e-5
    assert util.rel_err(data_util.computer_face_normal(V, fv), g["updated_normals"]) < 1e-4


# ------------------------------------------------------------------------------------------- patches, pipeline, training
def _per_patch_forced(g):
    labels = [torch.from_numpy(g[f"labels_{i}"]) for i in range(8 * int(g["n_patches"]))]
    return [[labels[8 * k + 2 * j: 8 * k + 2 * j + 2] for j in range(4)] for k in range(int(g["n_patches"]))]


def test_patch_walk_is_the_references():
    """dataset.py:156-193 with data_util.mesh_get_neighbor_np / get_submesh as the reference executes them: the oracle's and
    the product's (host C++) splitters give the same patches - faces in discovery order, vertices in first-use order - and the
    filtered walk registers the same cache names."""
    from geobi_gnn_b200 import patches
    from oracle import ref_patch
    g = np.load(os.path.join(util.GOLDEN, "reference_pipeline_ico8.npz"))
    mesh = synth.TriMesh(g["points_noisy"], g["faces"])
    sub = int(g["sub_size"])
    for split, cut in ((ref_patch.split_mesh, ref_patch.get_submesh), (patches.split_mesh, patches.get_submesh)):
        parts = split(g["points_noisy"], mesh.fv, mesh.vf, sub)
        assert len(parts) == int(g["n_patches"])
        for k, (sel, seed) in enumerate(parts):
            assert np.array_equal(np.asarray(sel), g[f"patch{k}_faces"]) and int(seed) == int(g[f"patch{k}_faces"][0])
            v_idx, faces = cut(mesh.fv, sel)
            assert np.array_equal(v_idx, g[f"patch{k}_vertices"])
            assert np.array_equal(g[f"patch{k}_vertices"][faces], mesh.fv[np.asarray(sel)])
        kept = split(g["points_noisy"], mesh.fv, mesh.vf, sub, int(g["filter_patch_count"]))
        assert [f"noisy-sub{sub}-{seed}.pt" for _, seed in kept] == list(g["filtered_names"])
        assert 0 < len(kept) == len(parts) - 1          # the 80-face component is dropped


def test_oracle_pipeline_and_training_step_reproduce_the_reference():
    from oracle import ref_patch
    # test_dual.predict_one on a mesh that takes the patch branch
    g = np.load(os.path.join(util.GOLDEN, "reference_pipeline_ico8.npz"))
    mesh, mesh_o = synth.TriMesh(g["points_noisy"], g["faces"]), synth.TriMesh(g["points_original"], g["faces"])
    torch.manual_seed(3)
    net = ref_network.DualGNN(force_depth=False, pool_type="max", wei_param=2).eval()
    forced, results = _per_patch_forced(g), []
    for k in range(int(g["n_patches"])):
        v_idx, faces = ref_patch.get_submesh(mesh.fv, g[f"patch{k}_faces"])
        dd = ref_dataset.process_one_submesh(synth.TriMesh(g["points_noisy"][v_idx], faces))
        ref_dataset.attach_normalisation(dd, g["points_noisy"], mesh.ev)
        centroid, scale = dd[0].centroid, dd[0].scale
        dd = ref_dataset.post_processing(dd)
        for pl, f in zip(util.poolings(net), forced[k]):
            pl.forced = f
        with torch.no_grad():
            vp, nrm, _ = net([dd[0], dd[1]])
        results.append((vp, nrm, torch.from_numpy(v_idx), torch.from_numpy(g[f"patch{k}_faces"])))
    Vp, Np = ref_dataset.stitch_patches(mesh.n_vertices, mesh.n_faces, results)
    V = ref_data_util.update_position2(Vp / scale + centroid, torch.from_numpy(mesh.fv), torch.from_numpy(mesh.vf), Np, 60)
    assert util.rel_err(V, g["updated_vertices"]) < 1e-5
    Nt = torch.from_numpy(np.asarray(mesh_o.face_normals, dtype=np.float32))
    assert abs(float(ref_network.error_n(Np, Nt)) / float(g["angle1"]) - 1) < 1e-4
    assert abs(float(ref_network.error_n(ref_data_util.computer_face_normal(V, torch.from_numpy(mesh.fv)), Nt)) / float(g["angle2"]) - 1) < 1e-4
    # one training micro-step (train_dual.py:204-214)
    t = np.load(os.path.join(util.GOLDEN, "reference_train_ico4.npz"))
    (dv, df), _, _ = util.oracle_inputs(4, seed=1)
    torch.manual_seed(2)
    net = ref_network.DualGNN(force_depth=False, pool_type="max", wei_param=2).train()
    labels = [torch.from_numpy(t[f"labels_{i}"]) for i in range(8)]
    for j, pl in enumerate(util.poolings(net)):
        pl.forced = labels[2 * j: 2 * j + 2]
    y_v, y_f = dv.y, df.y
    vp, nrm, _ = net([dv, df])
    loss = ref_network.dual_loss(ref_network.loss_v(vp, y_v, "L1"), ref_network.loss_n(nrm, y_f, "L1"))
    loss.backward()
    assert abs(float(loss) / float(t["loss"]) - 1) < 1e-5
    assert [n for n, _ in net.named_parameters()] == list(t["param_names"])
    for name, prm in net.named_parameters():
        _check_grad(name, prm.grad, t, 1e-4)


def _check_grad(name, grad, t, tol):
    flat = grad.detach().reshape(-1).cpu()
    idx = torch.linspace(0, flat.numel() - 1, min(512, flat.numel())).long()
    want = torch.from_numpy(t[f"gsample/{name}"])
    scale = float(t[f"gnorm/{name}"]) / max(flat.numel(), 1) ** 0.5 + 1e-12          # rms entry of the reference gradient
    assert abs(float(flat.norm()) - float(t[f"gnorm/{name}"])) <= tol * float(t[f"gnorm/{name}"]) + 1e-9, name
    assert float((flat[idx] - want).abs().max()) <= 10 * tol * max(scale, float(want.abs().max())), name


@pytest.fixture
def train_precision(request):
    from geobi_gnn_b200 import config
    old = config.get_precision()
    yield request.param
    config.set_precision(old)


@pytest.mark.gpu
@pytest.mark.parametrize("train_precision", ["fp32", "bf16x3"], indirect=True)
def test_cuda_pipeline_and_training_step_reproduce_the_reference(train_precision):
    """inference.predict_mesh (split -> per-patch forward -> stitch -> 60-sweep update) and one training micro-step through the
    CUDA kernels, teacher-forced with the reference's clusters, against what the reference's own code produced.  The training step
    runs in 'fp32' (library GEMMs in the backward) and in 'bf16x3' (all-native backward: the mode bench.py trains in)."""
    from geobi_gnn_b200 import config, data_util, inference, network
    DEV = "cuda"
    g = np.load(os.path.join(util.GOLDEN, "reference_pipeline_ico8.npz"))
    mesh, mesh_o = synth.TriMesh(g["points_noisy"], g["faces"]), synth.TriMesh(g["points_original"], g["faces"])
    torch.manual_seed(3)
    ref = ref_network.DualGNN(force_depth=False, pool_type="max", wei_param=2)
    net = network.DualGNN(force_depth=False, pool_type="max", wei_param=2).to(DEV).eval()
    net.load_state_dict(ref.state_dict())
    V, Np, Vp, n_patches = inference.predict_mesh(net, mesh, int(g["sub_size"]), device=DEV, forced=_per_patch_forced(g), return_parts=True)
    assert n_patches == int(g["n_patches"])
    assert util.rel_err(V, g["updated_vertices"]) < util.TOL_FP32
    Nt = torch.from_numpy(np.asarray(mesh_o.face_normals, dtype=np.float32)).to(DEV)
    fv = torch.from_numpy(mesh.fv).to(DEV)
    assert abs(float(network.error_n(Np, Nt)) / float(g["angle1"]) - 1) < 1e-3
    assert abs(float(network.error_n(data_util.computer_face_normal(V, fv), Nt)) / float(g["angle2"]) - 1) < 1e-3
    # training micro-step
    t = np.load(os.path.join(util.GOLDEN, "reference_train_ico4.npz"))
    (dv, df), _, _ = util.oracle_inputs(4, seed=1)
    torch.manual_seed(2)
    ref = ref_network.DualGNN(force_depth=False, pool_type="max", wei_param=2)
    net = network.DualGNN(force_depth=False, pool_type="max", wei_param=2).to(DEV).train()
    net.load_state_dict(ref.state_dict())
    config.set_precision(train_precision)
    labels = [torch.from_numpy(t[f"labels_{i}"]) for i in range(8)]
    for j, pl in enumerate(util.poolings(net)):
        pl.forced = labels[2 * j: 2 * j + 2]
    dv_m, df_m = util.data_to(dv, DEV), util.data_to(df, DEV)
    y_v, y_f = dv_m.y, df_m.y
    vp, nrm, _ = net([dv_m, df_m])
    loss = network.dual_loss(network.loss_v(vp, y_v, "L1"), network.loss_n(nrm, y_f, "L1"))
    loss.backward()
    assert abs(float(loss) / float(t["loss"]) - 1) < 1e-4
    assert [n for n, _ in net.named_parameters()] == list(t["param_names"])
    for name, prm in net.named_parameters():
        _check_grad(name, prm.grad, t, 2e-3)


# --------------------------------------------------------------------------------------------- PoolingLayer, every mode
def _pooling_input(g):
    mesh = synth.TriMesh(g["points_noisy"], g["faces"])
    dd = ref_dataset.process_one_submesh(mesh, "g", None)
    ref_dataset.attach_normalisation(dd, g["points_noisy"], mesh.ev)
    pos_f = dd[1].pos.clone()
    _, df = ref_dataset.post_processing(dd, "Synthetic")
    df.pos = pos_f                                       # is_plot=True upstream: the facet positions stay and are pooled too
    return df


def _check_pooled(g, key, layer, pooled, tol):
    assert np.array_equal(pooled.edge_index.cpu().numpy(), g[f"{key}/edge_index"])
    assert np.array_equal(layer.unpooling_indices.cpu().numpy(), g[f"{key}/unpooling_indices"])
    assert util.rel_err(pooled.x, g[f"{key}/x"]) < tol and util.rel_err(layer.unpooling(pooled.x), g[f"{key}/unpooled"]) < tol
    assert util.rel_err(pooled.pos, g[f"{key}/pos"]) < tol
    if f"{key}/edge_weight" in g.files:
        assert util.rel_err(pooled.edge_weight, g[f"{key}/edge_weight"]) < max(tol, 1e-6)
    else:
        assert pooled.edge_weight is None


@pytest.mark.parametrize("pool_type", ["max", "mean"])
def test_oracle_pooling_layer_is_the_references_in_every_mode(pool_type):
    from oracle import ref_net_util
    g = np.load(os.path.join(util.GOLDEN, "reference_pooling_ico3.npz"))
    for t in g["modes"].tolist():
        key = f"t{t}_{pool_type}"
        torch.manual_seed(t + 20)
        layer = ref_net_util.PoolingLayer(6, pool_type, 2, t, 2)
        g_perm = torch.Generator().manual_seed(50 + t)
        layer.perm_fn = lambda n: torch.randperm(n, generator=g_perm)        # free-running on the reference's seed
        with torch.no_grad():
            pooled = layer(_pooling_input(g))
        for i, tr in enumerate(layer.trace):
            assert np.array_equal(tr[3].numpy(), g[f"{key}/labels_{i}"]), (key, i)
            if i == 0 and f"{key}/matcher_weight" in g.files:
                assert util.rel_err(tr[1], g[f"{key}/matcher_weight"]) < 1e-6, key
        _check_pooled(g, key, layer, pooled, 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("pool_type", ["max", "mean"])
def test_cuda_pooling_layer_is_the_references_in_every_mode(pool_type):
    from geobi_gnn_b200 import net_util
    from oracle import ref_net_util
    g = np.load(os.path.join(util.GOLDEN, "reference_pooling_ico3.npz"))
    for t in g["modes"].tolist():
        key = f"t{t}_{pool_type}"
        torch.manual_seed(t + 20)
        ref = ref_net_util.PoolingLayer(6, pool_type, 2, t, 2)               # the reference's seeded parameters (CPU test)
        layer = net_util.PoolingLayer(6, pool_type, 2, t, 2).to("cuda")
        layer.load_state_dict(ref.state_dict())
        layer.forced = [torch.from_numpy(g[f"{key}/labels_{i}"]) for i in range(2)]
        with torch.no_grad():
            pooled = layer(util.data_to(_pooling_input(g), "cuda"))
        _check_pooled(g, key, layer, pooled, 1e-5)


# ------------------------------------------------------------------------------------------------ data_util entry points
def _check_data_util(du, g, dev, tol):
    pn = torch.from_numpy(g["points_noisy"]).to(dev)
    mesh = synth.TriMesh(g["points_noisy"], g["faces"])                       # open mesh: ragged vf / vv rows
    ev, vv = torch.from_numpy(mesh.ev).to(dev), torch.from_numpy(mesh.vv).to(dev)
    fv, vf = torch.from_numpy(mesh.fv).to(dev), torch.from_numpy(mesh.vf).to(dev)
    for s_type in (0, 1, 2, 3):
        q, c, sc = du.center_and_scale(pn, ev, s_type)
        for kind in ("np", "t"):                                              # the reference's numpy and tensor branches agree
            assert util.rel_err(q, g[f"cs_{kind}{s_type}/points"]) < tol and util.rel_err(c, g[f"cs_{kind}{s_type}/centroid"]) < tol
            assert abs(float(sc) / float(g[f"cs_{kind}{s_type}/scale"]) - 1) < tol
    assert np.array_equal(du.build_vertex_graph(ev, vv).cpu().numpy(), g["vertex_graph_2ring"])
    assert np.array_equal(du.build_edge_vf(vf).cpu().numpy(), g["edge_vf"])
    assert np.array_equal(du.build_edge_fv(fv).cpu().numpy(), g["edge_fv"])
    assert np.array_equal(du.build_facet_graph(fv, vf).cpu().numpy(), g["facet_graph"])
    assert util.rel_err(du.computer_face_normal(pn, fv), g["face_normals"]) < tol
    target, depth = torch.from_numpy(g["target_normals"]).to(dev), torch.nn.functional.normalize(pn, dim=1)
    assert util.rel_err(du.update_position(pn, fv, vf, target, 10), g["update_position_10"]) < tol
    assert util.rel_err(du.update_position(pn, fv, vf, target, 10, depth_direction=depth), g["update_position_10_depth"]) < tol
    assert util.rel_err(du.update_position2(pn, fv, vf, target, 10), g["update_position2_10"]) < tol
    assert util.rel_err(du.update_position2(pn, fv, vf, target, 10, depth_direction=depth), g["update_position2_10_depth"]) < tol
    vn = torch.from_numpy(np.asarray(mesh.vertex_normals, dtype=np.float32)).to(dev)
    ei = torch.cat([ev.t(), ev.t().flip(0)], 1).contiguous()
    assert util.rel_err(du.calc_weight(pn, vn, ei), g["calc_weight_vertex"]) < tol


def test_oracle_data_util_is_the_references():
    _check_data_util(ref_data_util, np.load(os.path.join(util.GOLDEN, "reference_data_util_ico4.npz")), "cpu", 1e-6)


@pytest.mark.gpu
def test_cuda_data_util_is_the_references():
    from geobi_gnn_b200 import data_util
    _check_data_util(data_util, np.load(os.path.join(util.GOLDEN, "reference_data_util_ico4.npz")), "cuda", 1e-5)


# --------------------------------------------------------------------------- BASELINE configs[0] at full size (20 480 faces)
def _config0():
    g = np.load(os.path.join(util.GOLDEN, "reference_config0_ico32.npz"))
    p, f = synth.icosphere(32)
    pn = synth.add_normal_noise(p, f, 0.2, seed=0).astype(np.float32)
    sha = hashlib.sha256(pn.tobytes() + f.astype(np.int64).tobytes()).digest()
    assert np.array_equal(np.frombuffer(sha, dtype=np.uint8), g["mesh_sha"]), "synthetic mesh generator changed"
    forced = [torch.from_numpy(g[f"labels_{i}"]).long() for i in range(8)]
    return g, pn, p.astype(np.float32), f, [forced[0:2], forced[2:4], forced[4:6], forced[6:8]]


def _check_config0(g, vp, nrm, y_v, y_f, nw, tol_v, tol_n):
    assert util.rel_err(vp, g["vert_p"]) < tol_v and util.rel_err(nrm, g["norm_p"]) < tol_n
    for name, val in (("loss_v_L1", nw.loss_v(vp, y_v, "L1")), ("loss_n_L1", nw.loss_n(nrm, y_f, "L1")),
                      ("error_v", nw.error_v(vp, y_v)), ("error_n", nw.error_n(nrm, y_f))):
        assert abs(float(val) / float(g[name]) - 1) < 1e-4, name


def test_oracle_config0_full_size_is_the_references():
    g, pn, p, f, forced = _config0()
    dd = ref_dataset.process_one_submesh(synth.TriMesh(pn, f), "g", synth.TriMesh(p, f))
    sha = hashlib.sha256(dd[0].edge_index.numpy().tobytes() + dd[1].edge_index.numpy().tobytes()).digest()
    assert np.array_equal(np.frombuffer(sha, dtype=np.uint8), g["edge_index_sha"])                   # 71 682 + 266 180 pairs, bit-equal
    assert [dd[0].pos.shape[0], dd[1].pos.shape[0], dd[0].edge_index.shape[1], dd[1].edge_index.shape[1]] == g["sizes"].tolist()
    ref_dataset.attach_normalisation(dd, pn, synth.TriMesh(pn, f).ev)
    dv, df = ref_dataset.post_processing(dd, "Synthetic")
    y_v, y_f = dv.y, df.y
    torch.manual_seed(0)
    net = ref_network.DualGNN(force_depth=False, pool_type="max", wei_param=2).eval()
    util.set_perm_fn(net, 4000)                                                                       # free-running on the reference's seed
    with torch.no_grad():
        vp, nrm, _ = net([dv, df])
    labels = [t[3] for pl in util.poolings(net) for t in pl.trace]
    assert all(np.array_equal(a.numpy(), g[f"labels_{i}"]) for i, a in enumerate(labels))
    assert [int(l.max()) + 1 for l in labels] == g["level_sizes"].tolist()
    _check_config0(g, vp, nrm, y_v, y_f, ref_network, 1e-5, 1e-5)


@pytest.mark.gpu
def test_cuda_config0_full_size_is_the_references():
    from geobi_gnn_b200 import dataset, network, topology
    g, pn, p, f, forced = _config0()
    dd = dataset.process_one_submesh(topology.DeviceTriMesh(pn, f, "cuda"), "g", topology.DeviceTriMesh(p, f, "cuda"), "cuda")
    sha = hashlib.sha256(dd[0].edge_index.cpu().numpy().tobytes() + dd[1].edge_index.cpu().numpy().tobytes()).digest()
    assert np.array_equal(np.frombuffer(sha, dtype=np.uint8), g["edge_index_sha"])
    dataset.attach_normalisation(dd, pn, synth.TriMesh(pn, f).ev)
    dv, df = dataset.post_processing(dd, "Synthetic")
    y_v, y_f = dv.y, df.y
    torch.manual_seed(0)
    net = network.DualGNN(force_depth=False, pool_type="max", wei_param=2).to("cuda").eval()
    net.load_state_dict(_seeded_state(0))                      # == the reference's seeded weights (state_sha check above)
    for pl, fl in zip(util.poolings(net), forced):
        pl.forced = fl
    with torch.no_grad():
        vp, nrm, _ = net([dv, df])
    _check_config0(g, vp, nrm, y_v, y_f, network, util.TOL_FP32, util.TOL_NORMAL)


def _seeded_state(seed):
    torch.manual_seed(seed)
    return ref_network.DualGNN(force_depth=False, pool_type="max", wei_param=2).state_dict()


def _check_misc(nw, nu, g, dev, tol):
    """laplacian_loss (network.py:347-361) and DualFusionLayer (net_util.py:248-278) against the reference's values."""
    from oracle import ref_net_util
    mesh = synth.TriMesh(g["points_noisy"], g["faces"])
    pts, moved = torch.from_numpy(g["points_noisy"]).to(dev), torch.from_numpy(g["moved_points"]).to(dev)
    ev = torch.from_numpy(mesh.ev).to(dev)
    ei = torch.cat([ev.t(), ev.t().flip(0), torch.arange(pts.shape[0], device=dev).repeat(2, 1)], 1)
    vn = torch.from_numpy(np.asarray(mesh.vertex_normals, dtype=np.float32)).to(dev)
    assert abs(float(nw.laplacian_loss(moved, pts, ei)) / float(g["laplacian_loss"]) - 1) < tol
    assert abs(float(nw.laplacian_loss(moved, pts, ei, normal=vn)) / float(g["laplacian_loss_normal"]) - 1) < tol
    fv = torch.from_numpy(mesh.fv).to(dev)
    edge_dual = torch.stack([torch.arange(fv.shape[0], device=dev).repeat_interleave(3), fv.reshape(-1)])     # build_edge_fv
    torch.manual_seed(4)
    seeded = ref_net_util.DualFusionLayer(8)                                 # the reference's seeded parameters
    layer = nu.DualFusionLayer(8).to(dev)
    layer.load_state_dict(seeded.state_dict())
    make = (lambda **kw: util.data_to(util.pyg.Data(**kw), dev)) if dev != "cpu" else util.pyg.Data
    data_v = make(x=torch.from_numpy(g["fusion_x_v"]).to(dev), edge_dual=edge_dual[1])
    data_f = make(x=torch.from_numpy(g["fusion_x_f"]).to(dev), edge_dual=edge_dual[0])
    with torch.no_grad():
        out_v, out_f = layer(data_v, data_f)
    assert util.rel_err(out_v, g["fusion_out_v"]) < tol and util.rel_err(out_f, g["fusion_out_f"]) < tol


def test_oracle_laplacian_loss_and_fusion_layer_are_the_references():
    from oracle import ref_net_util
    g = np.load(os.path.join(util.GOLDEN, "reference_data_util_ico4.npz"))
    _check_misc(ref_network, ref_net_util, g, "cpu", 1e-6)
    from geobi_gnn_b200 import network                                     # the product's losses are tensor expressions: CPU-checkable
    mesh = synth.TriMesh(g["points_noisy"], g["faces"])
    pts, moved = torch.from_numpy(g["points_noisy"]), torch.from_numpy(g["moved_points"])
    ev = torch.from_numpy(mesh.ev)
    ei = torch.cat([ev.t(), ev.t().flip(0), torch.arange(pts.shape[0]).repeat(2, 1)], 1)
    vn = torch.from_numpy(np.asarray(mesh.vertex_normals, dtype=np.float32))
    assert abs(float(network.laplacian_loss(moved, pts, ei)) / float(g["laplacian_loss"]) - 1) < 1e-6
    assert abs(float(network.laplacian_loss(moved, pts, ei, normal=vn)) / float(g["laplacian_loss_normal"]) - 1) < 1e-6


@pytest.mark.gpu
def test_cuda_laplacian_loss_and_fusion_layer_are_the_references():
    from geobi_gnn_b200 import net_util, network
    _check_misc(network, net_util, np.load(os.path.join(util.GOLDEN, "reference_data_util_ico4.npz")), "cuda", 1e-5)
