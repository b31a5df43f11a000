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
    for k in again.files:
        assert np.array_equal(again[k], g[k]), k


# ------------------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("n", sorted(CASES))
@pytest.mark.parametrize("front_end", ["host", "device"])
def test_cuda_path_reproduces_the_reference(n, front_end):
    """Mesh -> graphs -> DualGNN forward -> losses -> vertex update through the C-ABI kernels against the reference's vectors:
    index arrays bit-exact, floats within the stated fp32 tolerances (1e-5 op-level; 5e-5 / 2e-4 after ~20 stacked layers)."""
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
    assert util.rel_err(vp, g["vert_p"]) < 5e-5
    assert util.rel_err(nrm, g["norm_p"]) < 2e-4
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
