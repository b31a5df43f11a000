"""Host side around the hot path that needs no GPU: cache files, run files, augmentation,
collation and the directory walk of DualDataset over files that are already cached."""
import argparse
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from tests import util
from oracle import ref_dataset


def _raw_dual(n=3, seed=0):
    """(graph_v, graph_f) as process_one_data leaves them in a cache file (before post_processing), oracle-built, on the CPU."""
    mesh_n, mesh_o = util.noisy_icosphere(n, seed=seed)
    dd = ref_dataset.process_one_submesh(mesh_n, "g", mesh_o)
    ref_dataset.attach_normalisation(dd, mesh_n.points, mesh_n.ev)
    return dd, mesh_n, mesh_o


def _same(a, b):
    ka = {k for k in a.keys if getattr(a, k) is not None}
    kb = {k for k in (b.keys() if callable(b.keys) else b.keys) if getattr(b, k) is not None}
    assert ka - {"coalesced_undirected"} == kb - {"coalesced_undirected"}, (ka, kb)
    for k in kb - {"coalesced_undirected"}:
        x, y = getattr(a, k), getattr(b, k)
        if torch.is_tensor(y):
            assert x.dtype == y.dtype and torch.equal(x, y), k
        else:
            assert x == y or np.float32(x) == np.float32(y), k


def test_cache_file_round_trip(tmp_path):
    """save_dual_data / load_dual_data: every tensor bit-equal, scalars kept, no pickled class inside (weights_only load)."""
    from geobi_gnn_b200 import dataset
    dd, _, _ = _raw_dual()
    mine = tuple(util.data_to(d, "cpu") for d in dd)
    path = tmp_path / "g.pt"
    dataset.save_dual_data(mine, path)
    plain = torch.load(path, weights_only=True)
    assert plain["format"] == dataset._PT_FORMAT and set(plain) == {"format", "v", "f"}
    back = dataset.load_dual_data(path)
    for got, want in zip(back, dd):
        _same(got, want)
    assert isinstance(back[0].scale, float)
    # post_processing of the loaded sample == the oracle's post_processing of the original
    got = dataset.post_processing(back, "Synthetic")
    want = ref_dataset.post_processing(dd, "Synthetic")
    for g, w in zip(got, want):
        assert torch.equal(g.x, w.x) and torch.equal(g.y, w.y) and torch.equal(g.edge_index, w.edge_index)
    with pytest.raises(ValueError):
        torch.save({"something": 1}, tmp_path / "other.pt")
        dataset.load_dual_data(tmp_path / "other.pt")


class _FakePyG:
    """Installs throw-away `torch_geometric.data.data` / `.storage` modules so that a pickle with the reference's class paths
    can be written here (torch_geometric itself is not installed), and removes them again."""

    def __enter__(self):
        self.saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.data", "torch_geometric.data.data",
                                                       "torch_geometric.data.storage")}
        mods = {k: types.ModuleType(k) for k in self.saved}

        class Data:                                     # 1.x: attributes live in the instance dict
            pass

        class GlobalStorage:                            # 2.x: Data.__dict__['_store'] -> storage with '_mapping'
            pass

        for cls, mod in ((Data, "torch_geometric.data.data"), (GlobalStorage, "torch_geometric.data.storage")):
            cls.__module__, cls.__qualname__ = mod, cls.__name__
            setattr(mods[mod], cls.__name__, cls)
        sys.modules.update(mods)
        self.Data, self.Storage = Data, GlobalStorage
        return self

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_style_pickles_load_without_torch_geometric(tmp_path):
    """The reference caches `torch.save((Data_v, Data_f))` (dataset.py:154-155): both PyG layouts load into this package's
    Data with the same tensors, and a pickle that names any other global is refused."""
    from geobi_gnn_b200 import dataset
    dd, _, _ = _raw_dual()
    with _FakePyG() as fake:
        old, new = [], []
        for d in dd:
            state = {k: getattr(d, k) for k in d.keys()}
            state["scale"] = np.float32(state["scale"]) if "scale" in state else None      # upstream stores a numpy scalar
            o = fake.Data()
            o.__dict__.update(state)
            old.append(o)
            n, st = fake.Data(), fake.Storage()
            st.__dict__["_mapping"] = {k: v for k, v in state.items() if v is not None}
            n.__dict__["_store"] = st
            new.append(n)
        torch.save(tuple(old), tmp_path / "old.pt")
        torch.save(tuple(new), tmp_path / "new.pt")
    assert "torch_geometric" not in sys.modules
    for name in ("old.pt", "new.pt"):
        back = dataset.load_dual_data(tmp_path / name)
        for got, want in zip(back, dd):
            _same(got, want)
        assert isinstance(back[0].scale, float) and np.float32(back[0].scale) == np.float32(dd[0].scale)

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    torch.save((Evil(), Evil()), tmp_path / "evil.pt")
    with pytest.raises(pickle.UnpicklingError):
        dataset.load_dual_data(tmp_path / "evil.pt")


def test_random_rotate_and_collater():
    """dataset.py:19-69: same random draw as upstream (np.random.uniform(size=3)), z-only when z_rotated, features / targets /
    pos rotated as row vectors; the collater hands a tuple sample through."""
    from geobi_gnn_b200 import dataset
    dd, _, _ = _raw_dual(2, seed=4)
    sample = dataset.post_processing(tuple(util.data_to(d, "cpu") for d in dd), "Kinect_v1", is_plot=True)
    before = [(d.x.clone(), d.y.clone()) for d in sample]
    depth0, pos0 = sample[0].depth_direction.clone(), sample[0].pos.clone()
    for z_only in (True, False):
        for d, (x, y) in zip(sample, before):
            d.x, d.y = x.clone(), y.clone()
        sample[0].depth_direction, sample[0].pos = depth0.clone(), pos0.clone()
        np.random.seed(11)
        a = np.random.uniform(size=3) * 2 * np.pi
        c, s = np.cos(a), np.sin(a)
        rx = np.array([[1, 0, 0], [0, c[0], -s[0]], [0, s[0], c[0]]])
        ry = np.array([[c[1], 0, s[1]], [0, 1, 0], [-s[1], 0, c[1]]])
        rz = np.array([[c[2], -s[2], 0], [s[2], c[2], 0], [0, 0, 1]])
        rot = torch.from_numpy(rz if z_only else rz @ ry @ rx).float()
        np.random.seed(11)
        out = dataset.RandomRotate(z_only)(sample)
        assert out is sample
        for d, (x, y) in zip(out, before):
            assert torch.equal(d.x[:, :3], x[:, :3] @ rot) and torch.equal(d.x[:, 3:], x[:, 3:] @ rot)
            assert torch.equal(d.y, y @ rot)
        assert torch.equal(out[0].depth_direction, depth0 @ rot) and torch.equal(out[0].pos, pos0 @ rot)
        assert out[1].pos is not None                                                # is_plot keeps it, rotated too
        assert util.rel_err(out[0].x[:, 3:].norm(dim=1), before[0][0][:, 3:].norm(dim=1)) < 1e-6
    coll = dataset.Collater([])
    assert coll([sample]) is sample
    assert torch.equal(coll([0.5, 1.5]), torch.tensor([0.5, 1.5]))
    with pytest.raises(TypeError):
        coll([object()])


def test_run_files_round_trip(tmp_path):
    """`*_params.pth` (pickled argparse namespace, train_dual.py:127) and `*_model.pth` (state dict, :276) -> load_run."""
    from geobi_gnn_b200 import checkpoint, network
    opt = argparse.Namespace(data_type="Kinect_v1", flag="GeoBi-GNN_Kinect_v1_t_20260101-000000", force_depth=True, pool_type="max",
                             wei_param=10, sub_size=20000, model_name="GeoBi-GNN_Kinect_v1_model.pth", lr_step=[10], seed=7)
    params = tmp_path / "GeoBi-GNN_Kinect_v1_params.pth"
    checkpoint.save_params(opt, params)
    torch.manual_seed(3)
    net = network.DualGNN(force_depth=True, pool_type="max", wei_param=10)
    checkpoint.save_model(net, tmp_path / opt.model_name)
    opt2, net2 = checkpoint.load_run(str(params), "cpu", sub_size=5000)
    assert vars(opt2) == {**vars(opt), "sub_size": 5000}
    assert not net2.training and net2.force_depth
    for (k, a), (k2, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    # a params file written by plain pickle-based torch.save of a Namespace (what upstream does) loads the same way
    torch.save(opt, tmp_path / "plain.pth")
    assert vars(checkpoint.load_params(tmp_path / "plain.pth")) == vars(opt)


def test_directory_walk_over_cached_files(tmp_path):
    """DualDataset (dataset.py:72-283) on a tree whose samples are already cached: list order, `<name>_n*.obj` matching,
    `data_list_txt`, len / get / transform - all without a GPU (nothing has to be built)."""
    from geobi_gnn_b200 import dataset, meshio
    root = tmp_path / "dataset"
    split = root / "Synthetic" / "train"
    for sub in ("original", "noisy", "processed_data"):
        (split / sub).mkdir(parents=True)
    raws = {}
    for i, name in enumerate(("block", "ant", "unlisted")):
        dd, mesh_n, mesh_o = _raw_dual(2, seed=i)
        meshio.write_obj(split / "original" / f"{name}.obj", mesh_o.points, mesh_o.fv)
        for lvl in (1, 2):
            meshio.write_obj(split / "noisy" / f"{name}_n{lvl}.obj", mesh_n.points, mesh_n.fv)
            dataset.save_dual_data(tuple(util.data_to(d, "cpu") for d in dd), split / "processed_data" / f"{name}_n{lvl}.pt")
        raws[name] = dd
    (root / "Synthetic" / "train_list.txt").write_text("block\n\nant\n")
    ds = dataset.DualDataset("Synthetic", "train", data_list_txt="train_list.txt", root=str(root), device="cpu")
    assert len(ds) == 4 and ds.len() == 4
    assert [os.path.basename(f) for f in ds.processed_files] == ["block_n1.pt", "block_n2.pt", "ant_n1.pt", "ant_n2.pt"]
    assert [os.path.basename(f) for f in ds.files_original] == ["block.obj", "block.obj", "ant.obj", "ant.obj"]
    want = ref_dataset.post_processing(raws["ant"], "Synthetic")
    got = ds[2]
    for g, w in zip(got, want):
        assert torch.equal(g.x, w.x) and torch.equal(g.y, w.y) and torch.equal(g.edge_index, w.edge_index)
        assert torch.equal(g.edge_weight, w.edge_weight)
    assert "depth_direction" not in got[0] and got[0].pos is None and torch.equal(got[1].fv_indices, want[1].fv_indices)
    # without a list file: every original, sorted; transform applied on indexing only
    np.random.seed(0)
    ds2 = dataset.DualDataset("Synthetic", "train", root=str(root), device="cpu", transform=dataset.RandomRotate(False))
    assert len(ds2) == 6 and os.path.basename(ds2.processed_files[0]) == "ant_n1.pt"
    plain, turned = ds2.get(0), ds2[0]
    assert not torch.equal(plain[0].x, turned[0].x)
    assert util.rel_err(turned[0].x.norm(dim=1), plain[0].x.norm(dim=1)) < 1e-5
    assert sum(1 for _ in ds2) == 6


def test_building_a_sample_without_a_gpu_fails_loudly(tmp_path):
    """No CPU fallback on the data path either: a sample that is not cached has to be built by the CUDA graph builders, and
    asking for that on the CPU raises instead of quietly taking another route."""
    from geobi_gnn_b200 import dataset, meshio, synth
    from geobi_gnn_b200._lib import GeobiError
    p, f = synth.icosphere(3)
    meshio.write_obj(tmp_path / "a_n1.obj", p, f)
    with pytest.raises(GeobiError, match="no CPU fallback"):
        dataset.DualDataset.process_one_data(str(tmp_path / "a_n1.obj"), 10 ** 9, None, device="cpu")
