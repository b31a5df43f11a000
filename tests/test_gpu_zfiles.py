"""GPU: the file-level drivers around the hot path - DualDataset building its cache from .obj
files on the device, the training driver producing run files, and predict_dir consuming them."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from tests import util
from oracle import ref_dataset, ref_patch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _write_split(root, data_type, split, names, n, noise_levels=(1,)):
    from geobi_gnn_b200 import meshio, synth
    base = os.path.join(root, data_type, split)
    for sub in ("original", "noisy"):
        os.makedirs(os.path.join(base, sub), exist_ok=True)
    for i, name in enumerate(names):
        p, f = synth.icosphere(n)
        p = (p * (1.0 + 0.5 * i)).astype(np.float32)
        meshio.write_obj(os.path.join(base, "original", f"{name}.obj"), p, f)
        for lvl in noise_levels:
            meshio.write_obj(os.path.join(base, "noisy", f"{name}_n{lvl}.obj"), synth.add_normal_noise(p, f, 0.1 * lvl, seed=10 * i + lvl), f)
    return base


def _oracle_sample(noisy_obj, original_obj, sel=None):
    """What dataset.py:129-193 + post_processing give for one (sub)mesh, from the same files, on the CPU."""
    from geobi_gnn_b200 import meshio, synth
    pn, f = meshio.read_obj(noisy_obj)
    po, _ = meshio.read_obj(original_obj)
    pn, po = pn.astype(np.float32), po.astype(np.float32)
    whole = synth.TriMesh(pn, f)
    if sel is None:
        dd = ref_dataset.process_one_submesh(whole, "g", synth.TriMesh(po, f))
    else:
        v_idx, faces = ref_patch.get_submesh(whole.fv, sel)
        dd = ref_dataset.process_one_submesh(synth.TriMesh(pn[v_idx], faces), "g", synth.TriMesh(po[v_idx], faces))
    ref_dataset.attach_normalisation(dd, pn, whole.ev)
    return ref_dataset.post_processing(dd, "Synthetic"), whole


def _check_sample(got, want):
    for g, w in zip(got, want):
        assert g.x.is_cuda and torch.equal(g.edge_index.cpu(), w.edge_index)          # integer work: bit-exact
        assert util.rel_err(g.x, w.x) < 1e-5 and util.rel_err(g.y, w.y) < 1e-5
        assert util.rel_err(g.edge_weight, w.edge_weight) < 1e-5
    assert torch.equal(got[1].fv_indices.cpu(), want[1].fv_indices)
    assert "depth_direction" not in got[0] and got[0].pos is None


def test_dataset_builds_and_reuses_its_cache(tmp_path):
    from geobi_gnn_b200 import dataset
    root = str(tmp_path / "dataset")
    base = _write_split(root, "Synthetic", "train", ("a", "b"), 6, noise_levels=(1, 2))
    ds = dataset.DualDataset("Synthetic", "train", root=root, device=DEV)
    assert len(ds) == 4 and [os.path.basename(f) for f in ds.processed_files] == ["a_n1.pt", "a_n2.pt", "b_n1.pt", "b_n2.pt"]
    assert all(os.path.exists(f) for f in ds.processed_files)
    for i in (0, 3):
        want, _ = _oracle_sample(ds.files_noisy[i], ds.files_original[i])
        _check_sample(ds[i], want)
    stamp = {f: os.stat(f).st_mtime_ns for f in ds.processed_files}
    ds2 = dataset.DualDataset("Synthetic", "train", root=root, device=DEV)             # nothing is rebuilt
    assert {f: os.stat(f).st_mtime_ns for f in ds2.processed_files} == stamp
    a, b = ds.get(1), ds2.get(1)
    assert torch.equal(a[0].x, b[0].x) and torch.equal(a[1].x, b[1].x)
    built = dataset.DualDataset.process_one_data(ds.files_noisy[0], 10 ** 9, ds.files_original[0], device=DEV)   # obj=None: no files
    assert len(built) == 1 and built[0][1] is None and built[0][2] is None
    fresh = dataset.post_processing(built[0][0], "Synthetic")
    cached = ds.get(0)
    for g, w in zip(fresh, cached):
        assert torch.equal(g.edge_index, w.edge_index) and util.rel_err(g.x, w.x) < 1e-6 and util.rel_err(g.edge_weight, w.edge_weight) < 1e-6


def test_dataset_splits_big_meshes_into_the_oracles_patches(tmp_path):
    from geobi_gnn_b200 import dataset, meshio
    root = str(tmp_path / "dataset")
    base = _write_split(root, "Synthetic", "test", ("ball",), 9)
    sub, keep = 500, 120
    ds = dataset.DualDataset("Synthetic", "test", submesh_size=sub, filter_patch_count=keep, root=root, device=DEV)
    _, whole = _oracle_sample(ds.files_noisy[0], ds.files_original[0])
    parts = ref_patch.split_mesh(whole.points, whole.fv, whole.vf, sub, keep)
    assert len(ds) == len(parts) > 2
    assert [os.path.basename(f) for f in ds.processed_files] == [f"ball_n1-sub{sub}-{seed}.pt" for _, seed in parts]
    for k in (0, len(parts) - 1):
        want, _ = _oracle_sample(ds.files_noisy[0], ds.files_original[0], sel=parts[k][0])
        _check_sample(ds[k], want)
        pv, fv = meshio.read_obj(ds.processed_files[k][:-3] + ".obj")                  # the patch saved "for visualization"
        assert fv.shape[0] == len(parts[k][0]) and pv.shape[0] == want[0].x.shape[0]


def test_train_driver_then_predict_dir(tmp_path):
    """Two epochs of the training driver on a three-mesh set (gradient accumulation over 2), then predict_dir over the test
    split with the run files it wrote and over a bare directory of .obj files."""
    from geobi_gnn_b200 import checkpoint, inference, meshio, train
    root, logs = str(tmp_path / "dataset"), str(tmp_path / "log")
    _write_split(root, "Synthetic", "train", ("a", "b", "c"), 4)
    test_base = _write_split(root, "Synthetic", "test", ("t",), 8)
    opt = train.parse_arguments(["--data_type=Synthetic", "--flag=unit", "--gpu=0", "--seed=5", "--max_epoch=2", "--batch_size=2",
                                 "--sub_size=600", "--filter_patch_count=50", "--lr_sch=step", "--lr_decay=0.5", "--lr_step", "1"])
    params_file = train.train(opt, dataset_root=root, log_root=logs, tensorboard=False)
    run_dir = os.path.dirname(params_file)
    assert os.path.basename(params_file) == "GeoBi-GNN_Synthetic_params.pth" and os.path.exists(os.path.join(run_dir, "GeoBi-GNN_Synthetic_model.pth"))
    assert "Epoch   0" in open(os.path.join(run_dir, "training_info.txt")).read()
    rows = [json.loads(l) for l in open(os.path.join(run_dir, "train", "scalars.jsonl"))]
    assert len([r for r in rows if r.get("tag") == "dual_loss"]) == 4                     # 2 epochs x ceil(3 / 2) optimiser steps
    assert all(np.isfinite(r["value"]) for r in rows if "value" in r)
    evals = [json.loads(l) for l in open(os.path.join(run_dir, "test", "scalars.jsonl"))]
    assert len([r for r in evals if r.get("tag") == "error_f"]) == 2
    opt2, net = checkpoint.load_run(params_file, DEV)
    assert opt2.flag.startswith("GeoBi-GNN_Synthetic_unit_") and opt2.sub_size == 600 and not net.training
    # test split with ground truth: the 1280-face mesh goes through the patch branch (sub_size 600)
    faces, mean1, mean2 = inference.predict_dir(params_file, dataset_root=root, gpu=0)
    assert faces == 1280 and 0 < mean1 < 180 and 0 < mean2 < 180
    out = glob.glob(os.path.join(test_base, f"result_{opt2.flag}", "*.obj"))
    assert [os.path.basename(o) for o in out] == ["t_n1-60.obj"]
    pv, fv = meshio.read_obj(out[0])
    src_p, src_f = meshio.read_obj(os.path.join(test_base, "noisy", "t_n1.obj"))
    assert np.array_equal(fv, src_f) and pv.shape == src_p.shape and np.isfinite(pv).all()
    # bare directory, no ground truth, unsplit
    bare = tmp_path / "bare"
    bare.mkdir()
    meshio.write_obj(bare / "x.obj", src_p, src_f)
    faces, mean1, mean2 = inference.predict_dir(params_file, data_dir=str(bare), sub_size=10 ** 6, gpu=0)
    assert (faces, mean1, mean2) == (1280, 0.0, 0.0) and os.path.exists(bare / f"result_{opt2.flag}" / "x-60.obj")
