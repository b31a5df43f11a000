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


GRAD_SAMPLE = 512          # gradient entries kept per parameter tensor (evenly strided), plus every tensor's L2 norm


def grad_sample(t):
    flat = t.detach().reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, min(GRAD_SAMPLE, flat.numel())).long()
    return flat[idx].numpy()


def reference_training_case(ref, n, mesh_seed, weight_seed, data_type, wei_param):
    """One micro-step of train_dual.py:204-214 on the case's mesh: forward, L1 / L1 dual loss, backward - the reference's
    modules under torch autograd.  Stored: the loss, every parameter gradient's norm and a strided sample of its entries."""
    p, f = synth.icosphere(n)
    pn = synth.add_normal_noise(p, f, 0.2, seed=mesh_seed).astype(np.float32)
    mesh_n, mesh_o = _OMTriMesh(pn, f), _OMTriMesh(p.astype(np.float32), f)
    _, centroid, scale = ref.data_util.center_and_scale(pn, mesh_n.ev_indices())
    dual = ref.dataset.DualDataset.process_one_submesh(mesh_n, "g", mesh_o)
    dual[0].centroid = torch.from_numpy(np.asarray(centroid)).float()
    dual[0].scale = scale
    data_v, data_f = ref.dataset.DualDataset.post_processing(dual, data_type)
    y_v, y_f = data_v.y, data_f.y
    torch.manual_seed(weight_seed)
    net = ref.network.DualGNN(force_depth=data_type in ("Kinect_v1", "Kinect_v2"), pool_type="max", wei_param=wei_param)
    net.train()
    rec = _RecordedGraclus(2000 + weight_seed)
    ref.net_util.graclus = rec
    vert_p, norm_p, _ = net([data_v, data_f])
    nw = ref.network
    loss = nw.dual_loss(nw.loss_v(vert_p, y_v, "L1"), nw.loss_n(norm_p, y_f, "L1"), v_scale=1, n_scale=1)
    loss.backward()
    out = {"loss": loss.detach().numpy(), "n_poolings": np.int64(len(rec.labels))}
    for i, lab in enumerate(rec.labels):
        out[f"labels_{i}"] = lab.numpy()
    names = []
    for name, prm in net.named_parameters():
        names.append(name)
        out[f"gnorm/{name}"] = prm.grad.norm().numpy()
        out[f"gsample/{name}"] = grad_sample(prm.grad)
    out["param_names"] = np.array(names)
    return out


def reference_pipeline_case(ref, n, mesh_seed, weight_seed, data_type, wei_param, sub_size, filter_patch_count):
    """test_dual.predict_one (test_dual.py:25-87) as written, on a mesh big enough to take the patch branch: the reference's
    process_one_data (BFS patches with mesh_get_neighbor_np / get_submesh), per-patch forward, stitch, de-normalisation and
    60-sweep update_position2, plus the angle errors against the ground truth.  Mesh files are served from memory."""
    import argparse
    import openmesh
    import test_dual                                   # /root/reference/code/test_dual.py
    p, f = synth.icosphere(n)
    p2, f2 = synth.icosphere(2)                        # + a small separate component (80 faces): a patch the filter drops
    p, f = np.concatenate([p, p2 * 0.3 + np.array([[2.5, 0.0, 0.0]])]).astype(np.float32), np.concatenate([f, f2 + p.shape[0]])
    pn = synth.add_normal_noise(p, f, 0.2, seed=mesh_seed).astype(np.float32)
    files = {"/mem/noisy.obj": (pn, f), "/mem/original.obj": (p.astype(np.float32), f)}
    written = {}
    openmesh.read_trimesh = lambda name: _OMTriMesh(*files[name])
    openmesh.write_mesh = lambda name, mesh: written.__setitem__(name, np.array(mesh.points()))
    # the patch list on its own: (dual_data, V_idx, select_faces) per kept patch, in order (dataset.py:156-193)
    holder = types.SimpleNamespace(filter_patch_count=filter_patch_count, processed_dir="/nonexistent", processed_files=[])
    all_data = ref.dataset.DualDataset.process_one_data("/mem/noisy.obj", sub_size, "/mem/original.obj")
    assert len(all_data) > 2
    out = {"points_noisy": pn, "points_original": p.astype(np.float32), "faces": f.astype(np.int64), "n_patches": np.int64(len(all_data)),
           "sub_size": np.int64(sub_size)}
    for k, (dual, v_idx, sel) in enumerate(all_data):
        out[f"patch{k}_faces"] = np.asarray(sel, dtype=np.int64)
        out[f"patch{k}_vertices"] = np.asarray(v_idx, dtype=np.int64)
    # the filtered variant of the same walk (train_dual.py:141 passes filter_patch_count through the DualDataset object);
    # cache writes are redirected to nowhere: only the list of names it registers is of interest
    torch_save, ref.dataset.torch.save = ref.dataset.torch.save, (lambda *a, **k: None)
    try:
        ref.dataset.DualDataset.process_one_data("/mem/noisy.obj", sub_size, "/mem/original.obj", obj=holder)
    finally:
        ref.dataset.torch.save = torch_save
    out["filtered_names"] = np.array([os.path.basename(x) for x in holder.processed_files])
    out["filter_patch_count"] = np.int64(filter_patch_count)
    written.clear()                                    # the patch .obj files "for visualization" (dataset.py:185-186)
    # the whole prediction
    torch.manual_seed(weight_seed)
    net = ref.network.DualGNN(force_depth=data_type in ("Kinect_v1", "Kinect_v2"), pool_type="max", wei_param=wei_param).eval()
    rec = _RecordedGraclus(3000 + weight_seed)
    ref.net_util.graclus = rec
    opt = argparse.Namespace(sub_size=sub_size, data_type=data_type)
    angle1, angle2, n_faces = test_dual.predict_one(opt, net, torch.device("cpu"), "/mem/noisy.obj", "/mem/result.obj", "/mem/original.obj")
    assert list(written) == ["/mem/result-60.obj"] and len(rec.labels) == 8 * len(all_data)
    out.update(updated_vertices=written["/mem/result-60.obj"].astype(np.float32), angle1=np.float32(angle1), angle2=np.float32(angle2),
               n_faces=np.int64(n_faces))
    for i, lab in enumerate(rec.labels):
        out[f"labels_{i}"] = lab.numpy()
    return out


CASES = (dict(n=3, mesh_seed=0, weight_seed=0, data_type="Synthetic", wei_param=2),
         dict(n=5, mesh_seed=3, weight_seed=7, data_type="Kinect_v1", wei_param=10))


POOLING_MODES = (-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10)


def reference_pooling_cases(ref, n=3, mesh_seed=2):
    """net_util.PoolingLayer (net_util.py:56-245) for every edge_weight_type x pool_type on the facet graph of a small mesh:
    the matcher's input weights (`_get_edge_weight`), the pooled graph and the unpooling indices."""
    p, f = synth.icosphere(n)
    pn = synth.add_normal_noise(p, f, 0.2, seed=mesh_seed).astype(np.float32)
    mesh_n = _OMTriMesh(pn, f)
    _, centroid, scale = ref.data_util.center_and_scale(pn, mesh_n.ev_indices())
    out = {"points_noisy": pn, "faces": f.astype(np.int64), "modes": np.array(POOLING_MODES)}
    for t in POOLING_MODES:
        for pool_type in ("max", "mean"):
            dual = ref.dataset.DualDataset.process_one_submesh(_OMTriMesh(pn, f), "g", None)
            dual[0].centroid, dual[0].scale = torch.from_numpy(np.asarray(centroid)).float(), scale
            _, data_f = ref.dataset.DualDataset.post_processing(dual, "Synthetic", is_plot=True)     # keeps pos (pooled too)
            torch.manual_seed(t + 20)
            layer = ref.net_util.PoolingLayer(6, pool_type, 2, t, 2)
            rec = _RecordedGraclus(50 + t)
            ref.net_util.graclus = rec
            with torch.no_grad():
                w_in = layer._get_edge_weight(data_f)
                pooled = layer(data_f)
            key = f"t{t}_{pool_type}"
            out[f"{key}/x"] = pooled.x.numpy()
            out[f"{key}/edge_index"] = pooled.edge_index.numpy()
            out[f"{key}/pos"] = pooled.pos.numpy()
            out[f"{key}/fv_is_none"] = np.bool_(pooled.fv_indices is None) if "fv_indices" in pooled else np.bool_(True)
            if pooled.edge_weight is not None:
                out[f"{key}/edge_weight"] = pooled.edge_weight.numpy()
            if w_in is not None:
                out[f"{key}/matcher_weight"] = w_in.numpy()
            out[f"{key}/unpooling_indices"] = layer.unpooling_indices.numpy()
            for i, lab in enumerate(rec.labels):
                out[f"{key}/labels_{i}"] = lab.numpy()
            out[f"{key}/unpooled"] = layer.unpooling(pooled.x).numpy()
    return out


def reference_config0_case(ref, n=32, mesh_seed=0, weight_seed=0):
    """BASELINE.json configs[0] at full size: one noisy icosphere of 20 480 faces (10 242 vertices), Synthetic, random-init
    weights, through the reference's dataset assembly and DualGNN.forward.  The mesh is regenerated by the tests from
    synth.icosphere / add_normal_noise (sha-256 of both arrays stored); kept here: outputs, cluster labels (int32), losses."""
    import hashlib
    p, f = synth.icosphere(n)
    pn = synth.add_normal_noise(p, f, 0.2, seed=mesh_seed).astype(np.float32)
    mesh_n, mesh_o = _OMTriMesh(pn, f), _OMTriMesh(p.astype(np.float32), f)
    _, centroid, scale = ref.data_util.center_and_scale(pn, mesh_n.ev_indices())
    dual = ref.dataset.DualDataset.process_one_submesh(mesh_n, "g", mesh_o)
    dual[0].centroid, dual[0].scale = torch.from_numpy(np.asarray(centroid)).float(), scale
    sizes = np.array([dual[0].pos.shape[0], dual[1].pos.shape[0], dual[0].edge_index.shape[1], dual[1].edge_index.shape[1]])
    ei_sha = hashlib.sha256(dual[0].edge_index.numpy().tobytes() + dual[1].edge_index.numpy().tobytes()).digest()
    data_v, data_f = ref.dataset.DualDataset.post_processing(dual, "Synthetic")
    y_v, y_f = data_v.y, data_f.y
    torch.manual_seed(weight_seed)
    net = ref.network.DualGNN(force_depth=False, pool_type="max", wei_param=2).eval()
    rec = _RecordedGraclus(4000 + weight_seed)
    ref.net_util.graclus = rec
    with torch.no_grad():
        vert_p, norm_p, _ = net([data_v, data_f])
    nw = ref.network
    out = {"mesh_sha": np.frombuffer(hashlib.sha256(pn.tobytes() + f.astype(np.int64).tobytes()).digest(), dtype=np.uint8),
           "edge_index_sha": np.frombuffer(ei_sha, dtype=np.uint8), "sizes": sizes, "scale": np.float32(scale),
           "centroid": np.asarray(centroid, dtype=np.float32), "vert_p": vert_p.numpy(), "norm_p": norm_p.numpy(),
           "loss_v_L1": nw.loss_v(vert_p, y_v, "L1").numpy(), "loss_n_L1": nw.loss_n(norm_p, y_f, "L1").numpy(),
           "error_v": nw.error_v(vert_p, y_v).numpy(), "error_n": nw.error_n(norm_p, y_f).numpy(),
           "level_sizes": np.array([int(l.max()) + 1 for l in rec.labels])}
    for i, lab in enumerate(rec.labels):
        out[f"labels_{i}"] = lab.numpy().astype(np.int32)
    return out


def reference_data_util_case(ref, n=4, mesh_seed=6):
    """The remaining data_util entry points of SURVEY.md 8a that the forward cases do not reach: center_and_scale in all four
    scale modes (numpy and tensor inputs), build_vertex_graph (2-ring), build_edge_vf, computer_face_normal, update_position
    (the scatter variant, with and without the depth constraint)."""
    p, f = synth.icosphere(n)
    keep = p[f].mean(1)[:, 2] < 0.7                      # an OPEN mesh: boundary vertices, ragged vf / vv rows
    f = f[keep]
    used = np.unique(f)
    remap = np.full(p.shape[0], -1, dtype=np.int64)
    remap[used] = np.arange(used.shape[0])
    p, f = p[used], remap[f]
    pn = synth.add_normal_noise(p.astype(np.float32), f, 0.2, seed=mesh_seed).astype(np.float32)
    m = _OMTriMesh(pn, f)
    du = ref.data_util
    out = {"points_noisy": pn, "faces": f.astype(np.int64)}
    ev = m.ev_indices()
    for s_type in (0, 1, 2, 3):
        q, c, sc = du.center_and_scale(pn, ev, s_type)
        out[f"cs_np{s_type}/points"], out[f"cs_np{s_type}/centroid"], out[f"cs_np{s_type}/scale"] = q, c, np.float32(sc)
        q, c, sc = du.center_and_scale(torch.from_numpy(pn), torch.from_numpy(ev).long(), s_type)
        out[f"cs_t{s_type}/points"], out[f"cs_t{s_type}/centroid"], out[f"cs_t{s_type}/scale"] = q.numpy(), c.numpy(), sc.numpy()
    ev_t, vv_t = torch.from_numpy(ev).long(), torch.from_numpy(m.vv_indices()).long()
    fv_t, vf_t = torch.from_numpy(m.fv_indices()).long(), torch.from_numpy(m.vf_indices()).long()
    out["vertex_graph_2ring"] = du.build_vertex_graph(ev_t, vv_t).numpy()
    out["edge_vf"] = du.build_edge_vf(vf_t).numpy()
    out["edge_fv"] = du.build_edge_fv(fv_t).numpy()
    out["facet_graph"] = du.build_facet_graph(fv_t, vf_t).numpy()
    pts = torch.from_numpy(pn)
    normals = du.computer_face_normal(pts, fv_t)
    out["face_normals"] = normals.numpy()
    target = torch.nn.functional.normalize(normals + 0.3 * torch.randn(normals.shape, generator=torch.Generator().manual_seed(1)), dim=1)
    out["target_normals"] = target.numpy()
    depth = torch.nn.functional.normalize(pts, dim=1)
    out["update_position_10"] = du.update_position(pts, fv_t, vf_t, target, 10).numpy()
    out["update_position_10_depth"] = du.update_position(pts, fv_t, vf_t, target, 10, depth_direction=depth).numpy()
    out["update_position2_10"] = du.update_position2(pts, fv_t, vf_t, target, 10).numpy()
    out["update_position2_10_depth"] = du.update_position2(pts, fv_t, vf_t, target, 10, depth_direction=depth).numpy()
    vn = torch.from_numpy(np.asarray(m.vertex_normals(), dtype=np.float32))
    ei = torch.cat([ev_t.t(), ev_t.t().flip(0)], 1)
    out["calc_weight_vertex"] = du.calc_weight(pts, vn, ei).numpy()
    # network.laplacian_loss (network.py:347-361) on the vertex graph with self loops, with and without the normal projection
    ei_loops = torch.cat([ei, torch.arange(pts.shape[0]).repeat(2, 1)], 1)
    moved = pts + 0.05 * torch.randn(pts.shape, generator=torch.Generator().manual_seed(2))
    out["laplacian_loss"] = ref.network.laplacian_loss(moved, pts, ei_loops).numpy()
    out["laplacian_loss_normal"] = ref.network.laplacian_loss(moved, pts, ei_loops, normal=vn).numpy()
    out["moved_points"] = moved.numpy()
    # net_util.DualFusionLayer (net_util.py:248-278) on the vertex <-> facet incidence (edge_dual of process_one_submesh)
    edge_dual = du.build_edge_fv(fv_t)
    gen = torch.Generator().manual_seed(3)
    x_v, x_f = torch.randn(pts.shape[0], 8, generator=gen), torch.randn(fv_t.shape[0], 8, generator=gen)
    torch.manual_seed(4)
    fusion = ref.net_util.DualFusionLayer(8)
    data_v, data_f = pyg.Data(x=x_v.clone(), edge_dual=edge_dual[1]), pyg.Data(x=x_f.clone(), edge_dual=edge_dual[0])
    with torch.no_grad():
        f_v, f_f = fusion(data_v, data_f)
    out.update(fusion_x_v=x_v.numpy(), fusion_x_f=x_f.numpy(), fusion_out_v=f_v.numpy(), fusion_out_f=f_f.numpy())
    return out


TRAIN_CASE = dict(n=4, mesh_seed=1, weight_seed=2, data_type="Synthetic", wei_param=2)
PIPELINE_CASE = dict(n=8, mesh_seed=4, weight_seed=3, data_type="Synthetic", wei_param=2, sub_size=300, filter_patch_count=150)


def main():
    ref = import_reference()
    here = os.path.dirname(os.path.abspath(__file__))
    for case in CASES:
        out, net = reference_case(ref, **case)
        out["case"] = np.array(repr(sorted(case.items())))
        path = os.path.join(here, f"reference_ico{case['n']}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes,", sum(p.numel() for p in net.parameters()), "parameters, keys:", len(out))
    for name, fn, case in (("reference_train_ico4.npz", reference_training_case, TRAIN_CASE),
                           ("reference_pipeline_ico8.npz", reference_pipeline_case, PIPELINE_CASE)):
        out = fn(ref, **case)
        out["case"] = np.array(repr(sorted(case.items())))
        np.savez_compressed(os.path.join(here, name), **out)
        print(name, os.path.getsize(os.path.join(here, name)), "bytes, keys:", len(out))
    out = reference_config0_case(ref)
    np.savez_compressed(os.path.join(here, "reference_config0_ico32.npz"), **out)
    print("reference_config0_ico32.npz", os.path.getsize(os.path.join(here, "reference_config0_ico32.npz")), "bytes, sizes:", out["sizes"])
    out = reference_data_util_case(ref)
    np.savez_compressed(os.path.join(here, "reference_data_util_ico4.npz"), **out)
    print("reference_data_util_ico4.npz", os.path.getsize(os.path.join(here, "reference_data_util_ico4.npz")), "bytes, keys:", len(out))
    out = reference_pooling_cases(ref)
    np.savez_compressed(os.path.join(here, "reference_pooling_ico3.npz"), **out)
    print("reference_pooling_ico3.npz", os.path.getsize(os.path.join(here, "reference_pooling_ico3.npz")), "bytes, keys:", len(out))


if __name__ == "__main__":
    main()
