"""Golden vectors from THE REFERENCE'S OWN CODE, executed in the build container.  TEST INFRASTRUCTURE ONLY.

/root/reference/code/{network,net_util,data_util,dataset}.py cannot be imported as they stand: torch_geometric,
torch_scatter, torch_sparse and openmesh are not installable here.  This script installs stand-ins for exactly
those THIRD-PARTY packages in `sys.modules` (built from oracle/pyg.py and the numpy mesh of geobi_gnn_b200/synth.py), puts the
unmodified reference directory on `sys.path` and imports it.  Everything the reference itself wrote on the hot path then runs
as written - `DualDataset.process_one_submesh / post_processing`, `data_util.center_and_scale / calc_weight /
build_facet_graph / build_edge_fv / update_position2 / computer_face_normal`, `net_util.PoolingLayer / DualFusionLayer /
pool_edge`, `network.GNNModule / DualGNN / loss_* / error_*` - and its outputs are stored in
`tests/golden/reference_ico{N}.npz`.

What this pins: the oracle's restatement of the reference's own files (oracle/ref_*.py) and, through tests/test_gpu_*.py, the
CUDA path, against the reference executed here.  What it does NOT pin: the third-party operators (FeaStConv, graclus,
scatter, coalesce, OpenMesh index arrays), which both sides take from the same restatement.

Run (in the container that has /root/reference):   python tests/golden/make_reference_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE = "/root/reference/code"
sys.path.insert(0, ROOT)

from geobi_gnn_b200 import synth  # noqa: E402
from oracle import pyg  # noqa: E402


class _OMTriMesh:
    """The openmesh.TriMesh calls the path makes (dataset.py:134-243, test_dual.py:29-73) over synth.TriMesh."""

    def __init__(self, points=None, face_vertex_indices=None):
        self._m = synth.TriMesh(np.asarray(points), np.asarray(face_vertex_indices))

    def points(self):
        return self._m.points

    def n_faces(self):
        return self._m.n_faces

    def n_vertices(self):
        return self._m.n_vertices

    def ev_indices(self):
        return self._m.ev

    def fv_indices(self):
        return self._m.fv

    def vf_indices(self):
        return self._m.vf

    def vv_indices(self):
        return self._m.vv

    def update_face_normals(self):
        self._m.update_normals()

    def update_vertex_normals(self):
        pass

    def face_normals(self):
        return self._m.face_normals

    def vertex_normals(self):
        return self._m.vertex_normals


def install_third_party_stand_ins():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if not hasattr(np, "bool"):                       # data_util.py:64,245,286 / dataset.py:157 use the alias numpy 1.24 removed
        np.bool = bool
    placeholder = type("Placeholder", (torch.nn.Module,), {"__init__": lambda self, *a, **k: torch.nn.Module.__init__(self)})
    mod("openmesh", TriMesh=_OMTriMesh)
    mod("matplotlib", pyplot=mod("matplotlib.pyplot"))
    mod("torch_scatter", scatter=pyg.scatter)
    mod("torch_sparse", coalesce=pyg.coalesce)
    consecutive = mod("torch_geometric.nn.pool.consecutive", consecutive_cluster=pyg.consecutive_cluster)
    pool_pool = mod("torch_geometric.nn.pool.pool", pool_pos=pyg.pool_pos)
    pool = mod("torch_geometric.nn.pool", consecutive=consecutive, pool=pool_pool)
    nn = mod("torch_geometric.nn", FeaStConv=pyg.FeaStConv, GCNConv=placeholder, GATConv=placeholder, graclus=pyg.graclus, pool=pool)
    utils = mod("torch_geometric.utils", to_undirected=pyg.to_undirected, remove_self_loops=pyg.remove_self_loops,
                add_self_loops=pyg.add_self_loops)
    data = mod("torch_geometric.data", Data=pyg.Data, Batch=type("Batch", (), {}),
               Dataset=type("Dataset", (), {"__init__": lambda self, transform=None: setattr(self, "transform", transform)}))
    mod("torch_geometric", nn=nn, utils=utils, data=data)


def import_reference():
    install_third_party_stand_ins()
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import data_util, dataset, net_util, network       # noqa: E401  (the reference's modules, unmodified)
    assert os.path.dirname(os.path.abspath(network.__file__)) == REFERENCE
    return types.SimpleNamespace(data_util=data_util, dataset=dataset, net_util=net_util, network=network)


class _RecordedGraclus:
    """torch_geometric.nn.graclus as net_util.py:127 calls it, with the random visiting order drawn from a seeded generator and
    the labels kept, so the oracle / CUDA side can be teacher-forced with the reference's clusters."""

    def __init__(self, seed):
        self.gen, self.labels = torch.Generator().manual_seed(seed), []

    def __call__(self, edge_index, weight=None, num_nodes=None):
        n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
        lab = pyg.graclus(edge_index, weight, n, perm=torch.randperm(n, generator=self.gen))
        self.labels.append(lab.clone())
        return lab


def reference_case(ref, n, mesh_seed, weight_seed, data_type, wei_param):
    p, f = synth.icosphere(n)
    pn = synth.add_normal_noise(p, f, 0.2, seed=mesh_seed).astype(np.float32)
    mesh_n, mesh_o = _OMTriMesh(pn, f), _OMTriMesh(p.astype(np.float32), f)
    out = {"points_noisy": pn, "points_original": p.astype(np.float32), "faces": f.astype(np.int64)}
    # ---- dataset.py:140-153 (single-patch branch of process_one_data), :196-269
    _, centroid, scale = ref.data_util.center_and_scale(pn, mesh_n.ev_indices())
    dual = ref.dataset.DualDataset.process_one_submesh(mesh_n, "g", mesh_o)
    out.update(centroid=np.asarray(centroid, dtype=np.float32), scale=np.float32(scale),
               raw_v_edge_index=dual[0].edge_index.numpy(), raw_v_edge_weight=dual[0].edge_weight.numpy(),
               raw_f_edge_index=dual[1].edge_index.numpy(), raw_f_edge_weight=dual[1].edge_weight.numpy(),
               edge_dual_v=dual[0].edge_dual.numpy(), edge_dual_f=dual[1].edge_dual.numpy())
    dual[0].centroid = torch.from_numpy(np.asarray(centroid)).float()
    dual[0].scale = scale
    data_v, data_f = ref.dataset.DualDataset.post_processing(dual, data_type)
    out.update(x_v=data_v.x.numpy(), y_v=data_v.y.numpy(), x_f=data_f.x.numpy(), y_f=data_f.y.numpy())
    # ---- network.py:254-343 with net_util.py:56-302
    torch.manual_seed(weight_seed)
    net = ref.network.DualGNN(force_depth=data_type in ("Kinect_v1", "Kinect_v2"), pool_type="max", wei_param=wei_param)
    net.eval()
    rec = _RecordedGraclus(1000 + weight_seed)
    ref.net_util.graclus = rec                          # the name PoolingLayer looks up (net_util.py:5,127)
    with torch.no_grad():
        vert_p, norm_p, alpha = net([data_v, data_f])
    out.update(vert_p=vert_p.numpy(), norm_p=norm_p.numpy())
    for i, lab in enumerate(rec.labels):
        out[f"labels_{i}"] = lab.numpy()
    out["n_poolings"] = np.int64(len(rec.labels))
    # ---- losses / errors (network.py:347-413) and the vertex update (data_util.py:529-556, test_dual.py:63-72)
    nw = ref.network
    out.update(loss_v_L1=nw.loss_v(vert_p, data_v.y, "L1").numpy(), loss_v_L2=nw.loss_v(vert_p, data_v.y, "L2").numpy(),
               loss_n_L1=nw.loss_n(norm_p, data_f.y, "L1").numpy(), loss_n_L2=nw.loss_n(norm_p, data_f.y, "L2").numpy(),
               error_v=nw.error_v(vert_p, data_v.y).numpy(), error_n=nw.error_n(norm_p, data_f.y).numpy(),
               dual_loss=nw.dual_loss(nw.loss_v(vert_p, data_v.y, "L1"), nw.loss_n(norm_p, data_f.y, "L1"), 2.0, 0.5).numpy())
    Vp = vert_p / scale + torch.from_numpy(np.asarray(centroid)).float()
    fv, vf = torch.from_numpy(mesh_n.fv_indices()).long(), torch.from_numpy(mesh_n.vf_indices()).long()
    depth = torch.nn.functional.normalize(torch.from_numpy(pn), dim=1) if data_type in ("Kinect_v1", "Kinect_v2") else None
    V = ref.data_util.update_position2(Vp, fv, vf, norm_p, 60, depth_direction=depth)
    out.update(updated_vertices=V.numpy(), updated_normals=ref.data_util.computer_face_normal(V, fv).numpy())
    out["state_sha"] = np.frombuffer(_state_digest(net), dtype=np.uint8)
    return out, net


def _state_digest(net):
    import hashlib
    h = hashlib.sha256()
    for k, v in net.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.digest()


CASES = (dict(n=3, mesh_seed=0, weight_seed=0, data_type="Synthetic", wei_param=2),
         dict(n=5, mesh_seed=3, weight_seed=7, data_type="Kinect_v1", wei_param=10))


def main():
    ref = import_reference()
    here = os.path.dirname(os.path.abspath(__file__))
    for case in CASES:
        out, net = reference_case(ref, **case)
        out["case"] = np.array(repr(sorted(case.items())))
        path = os.path.join(here, f"reference_ico{case['n']}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes,", sum(p.numel() for p in net.parameters()), "parameters, keys:", len(out))


if __name__ == "__main__":
    main()
