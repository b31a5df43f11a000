"""Graph + feature assembly with the reference's surface (/root/reference/code/dataset.py:196-269),
on the device.

``process_one_submesh`` / ``post_processing`` keep their names, argument meaning and the layout
of the returned ``(graph_v, graph_f)`` tuple.  OpenMesh (not installable here) is replaced by
any object exposing its index arrays as numpy: ``points, ev, fv, vf, vv, face_normals,
vertex_normals`` (geobi_gnn_b200/synth.py:TriMesh, topology.DeviceTriMesh).
The directory data set with its .pt cache, the augmentation and the collater (dataset.py:19-283)
are in the second half of this file; the Kinect depth-map readers are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import data_util, ops
from .data import Data


def _host_array(a):
    return a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)


def _t(a, dtype, device):
    if torch.is_tensor(a):                      # topology.DeviceTriMesh hands over device tensors
        return a.to(device=device, dtype=dtype)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(device)


def process_one_submesh(mesh_n, name="graph", mesh_o=None, device="cuda", csr_native=False):
    """dataset.py:196-243.

    csr_native (not upstream; needs a topology.DeviceTriMesh): both graphs are handed to the network as the loop-free CSRs the
    device front end already holds (`Data.csr`, bilateral weights computed in CSR order by geobi_calc_weight_csr), and the
    reference's int64 `edge_index` (with its self loops) / `edge_weight` become LAZY items, built by the list-based code below only
    if a caller reads them - the same arrangement PoolingLayer uses for the coarse levels.  It saves the CSR -> list -> CSR round
    trip (0.7 of the front end's 2.3 ms of GPU time per million faces); weights agree with the list-based ones to the last bit
    unless the fp64 mean edge length rounds differently to fp32 (a different summation order)."""
    fv = _t(mesh_n.fv, torch.long, device)
    vf = _t(mesh_n.vf, torch.long, device)
    edge_dual_fv = data_util.build_edge_fv(fv)
    pos_v = _t(mesh_n.points, torch.float32, device)
    normal_v = _t(mesh_n.vertex_normals, torch.float32, device)
    pos_f = pos_v[fv].mean(1)
    normal_f = _t(mesh_n.face_normals, torch.float32, device).reshape(-1, 3)
    vf_sorted = bool(getattr(mesh_n, "vf_sorted", False))
    if csr_native:
        if getattr(mesh_n, "vertex_csr", None) is None or not vf_sorted:
            raise ValueError("process_one_submesh(csr_native=True) needs a topology.DeviceTriMesh")
        g_v = mesh_n.vertex_csr
        g_v = g_v.with_weight(ops.calc_weight_csr(pos_v, normal_v, g_v, pos_v.size(0)))
        g_f = ops.build_facet_graph_csr(fv, vf, True, drop_self=True)
        g_f.symmetric = True
        g_f = g_f.with_weight(ops.calc_weight_csr(pos_f, normal_f, g_f, fv.size(0)))
        graph_v = Data(name=f"{name}-v", pos=pos_v, normal=normal_v, depth_direction=F.normalize(pos_v, dim=1), edge_dual=edge_dual_fv[1],
                       coalesced_undirected=True)
        graph_f = Data(name=f"{name}-f", pos=pos_f, normal=normal_f, fv_indices=fv, edge_dual=edge_dual_fv[0], coalesced_undirected=True)

        def lists_v(cache=[]):
            if not cache:
                ei = data_util.with_self_loops_appended(mesh_n.vertex_csr)
                cache.append((ei, data_util.calc_weight(pos_v, normal_v, ei)))
            return cache[0]

        def lists_f(cache=[]):
            if not cache:
                ei = data_util.build_facet_graph(fv, vf, vf_sorted=True)
                cache.append((ei, data_util.calc_weight(pos_f, normal_f, ei)))
            return cache[0]

        for d, lists, g in ((graph_v, lists_v, g_v), (graph_f, lists_f, g_f)):
            d.set_lazy("edge_index", lambda lists=lists: lists()[0])
            d.set_lazy("edge_weight", lambda lists=lists: lists()[1])
            d.csr = g                    # after the lazies: assigning edge_index / edge_weight drops a stale csr
    else:
        if getattr(mesh_n, "vertex_csr", None) is not None:      # DeviceTriMesh already holds to_undirected(ev) as a CSR
            edge_idx_v = data_util.with_self_loops_appended(mesh_n.vertex_csr)
        else:
            ev = _t(mesh_n.ev, torch.long, device)
            edge_idx_v = data_util.to_undirected_with_self_loops(ev.t().contiguous(), pos_v.size(0))
        edge_wei_v = data_util.calc_weight(pos_v, normal_v, edge_idx_v)
        graph_v = Data(name=f"{name}-v", pos=pos_v, normal=normal_v, edge_index=edge_idx_v, edge_weight=edge_wei_v,
                       depth_direction=F.normalize(pos_v, dim=1), edge_dual=edge_dual_fv[1], coalesced_undirected=True)
        edge_idx_f = data_util.build_facet_graph(fv, vf, vf_sorted=vf_sorted)
        edge_wei_f = data_util.calc_weight(pos_f, normal_f, edge_idx_f)
        graph_f = Data(name=f"{name}-f", pos=pos_f, normal=normal_f, edge_index=edge_idx_f, edge_weight=edge_wei_f,
                       fv_indices=fv, edge_dual=edge_dual_fv[0], coalesced_undirected=True)
    # coalesced_undirected: both builders above emit sorted, duplicate-free, symmetric lists (self loops aside), which
    # lets the network build one CSR per graph without a sort (nn.input_graph); the device verifies the claim
    if mesh_o is not None:
        graph_v.y = _t(mesh_o.points, torch.float32, device)
        graph_f.y = _t(mesh_o.face_normals, torch.float32, device)
    return graph_v, graph_f


def normalisation(points_noisy, ev):
    """(centroid [1,3] fp32, scale) of dataset.py:140,151-152: `q = p - centroid; e = q[ev]; 1 / |e0 - e1|.mean()`.  The per-edge
    lengths come from one threaded C++ pass with numpy's fp32 operation order (bit-equal, tests/test_abi.py) and numpy takes
    the mean, so the numbers are upstream's; the [E,2,3] temporaries of a 10 M-face mesh are not built."""
    import ctypes as C
    import os
    from . import patches
    p = np.ascontiguousarray(_host_array(points_noisy), dtype=np.float32)      # a DeviceTriMesh hands over device tensors
    ev = np.ascontiguousarray(_host_array(ev), dtype=np.int64)
    centroid = np.ascontiguousarray(p.mean(0, keepdims=True))
    length = np.empty(ev.shape[0], dtype=np.float32)
    patches._host().geobi_host_edge_lengths(patches._p(p), patches._p(centroid), patches._p(ev), C.c_int64(ev.shape[0]), patches._p(length),
                                            C.c_int(min(8, os.cpu_count() or 1)))
    return centroid, 1 / length.mean()


def attach_normalisation(dual_data, points_noisy, ev, precomputed=None):
    """dataset.py:140,151-152: centroid / scale of the whole noisy mesh (numpy fp32, as upstream).
    `precomputed` = (centroid tensor, scale) lets a caller that normalises many patches of one mesh pay for it once."""
    if precomputed is not None:
        dual_data[0].centroid, dual_data[0].scale = precomputed
        return dual_data
    centroid, scale = normalisation(points_noisy, ev)
    dual_data[0].centroid = torch.from_numpy(centroid).float().to(dual_data[0].pos.device)
    dual_data[0].scale = float(scale)
    return dual_data


def post_processing(dual_data, data_type="Synthetic", is_plot=False):
    """dataset.py:245-269."""
    data_v, data_f = dual_data
    data_f.x = torch.cat(((data_f.pos - data_v.centroid) * data_v.scale, data_f.normal), 1)
    data_f.normal = data_f.edge_dual = None
    if not is_plot:
        data_f.pos = None
    data_v.x = torch.cat(((data_v.pos - data_v.centroid) * data_v.scale, data_v.normal), 1)
    data_v.y = None if data_v.y is None else (data_v.y - data_v.centroid) * data_v.scale
    data_v.normal = data_v.centroid = data_v.scale = data_v.edge_dual = None
    if not is_plot:
        data_v.pos = None
    else:
        data_v.pos = data_v.y
        data_v.fv_indices = data_f.fv_indices
    if data_type not in ["Kinect_v1", "Kinect_v2"]:
        data_v.depth_direction = None
    return data_v, data_f


def build_dual_on_device(mesh_n, mesh_o=None, data_type="Synthetic", name="graph", csr_native=False):
    """build_dual_data for a topology.DeviceTriMesh with the normalisation (dataset.py:140,151-152: centroid, 1 / mean edge length)
    computed on the device as well - no host round trip and no stream synchronisation beyond the mesh's own entry counts.  The fp32
    mean is taken by a device reduction instead of numpy's pairwise sum: centroid / scale agree with `normalisation` to ~1e-7."""
    dev = mesh_n.points.device
    dd = process_one_submesh(mesh_n, name, mesh_o, dev, csr_native=csr_native)
    pts = mesh_n.points
    centroid = pts.mean(0, keepdim=True)
    # mean undirected edge length = mean over the entries of the symmetric vertex CSR (fp64 sum of the fp32 lengths)
    dd[0].centroid, dd[0].scale = centroid, 1.0 / ops.mean_edge_length_csr(pts - centroid, mesh_n.vertex_csr)
    return post_processing(dd, data_type)


def build_dual_data(mesh_n, mesh_o=None, data_type="Synthetic", name="graph", device="cuda"):
    """Single-patch branch of process_one_data (dataset.py:144-153) followed by post_processing."""
    dd = process_one_submesh(mesh_n, name, mesh_o, device)
    attach_normalisation(dd, mesh_n.points, mesh_n.ev)
    return post_processing(dd, data_type)


# ---------------------------------------------------------------------------------------------------------------------
# Directory data set, cache files and augmentation: the host side of train_dual.py / test_dual.py
# around the hot path.  OpenMesh is replaced by meshio.read_obj + topology.DeviceTriMesh (synth.TriMesh without a GPU).
# ---------------------------------------------------------------------------------------------------------------------
import glob
import os
import sys

CODE_DIR = os.path.dirname(os.path.abspath(__file__))
BASE_DIR = os.path.dirname(CODE_DIR)
DATASET_DIR = os.path.join(BASE_DIR, "dataset")      # dataset.py:12-15; override per instance with DualDataset(root=...)
LOG_DIR = os.path.join(BASE_DIR, "log")


class Collater:
    """dataset.py:19-36 for the element type the drivers use: a tuple sample is handed through as it is (the reference
    trains with DataLoader batch_size 1 and accumulates gradients over `opt.batch_size` steps, train_dual.py:211-218)."""

    def __init__(self, follow_batch=()):
        self.follow_batch = follow_batch

    def collate(self, batch):
        elem = batch[0]
        if isinstance(elem, float):
            return torch.tensor(batch, dtype=torch.float)
        if isinstance(elem, tuple):
            return elem
        raise TypeError(f"DataLoader found invalid type: {type(elem)}")

    __call__ = collate


class RandomRotate:
    """dataset.py:39-69.  Same draw (np.random.uniform(size=3) * 2*pi, so a seeded run rotates as upstream does), same
    matrices, and - as upstream - z_rotated=True means "about z only".  The rotation is applied on whatever device the sample
    lives on; the features are rotated in place."""

    def __init__(self, z_rotated=True):
        self.z_rotated = z_rotated

    def __call__(self, data):
        a = np.random.uniform(size=(3)) * 2 * np.pi
        rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
        rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        rot = rz if self.z_rotated else np.dot(rz, np.dot(ry, rx))
        rot = torch.from_numpy(rot).to(data[0].y.dtype).to(data[0].y.device)
        for d in data:
            d.x[:, 0:3] = torch.matmul(d.x[:, 0:3], rot)
            d.x[:, 3:6] = torch.matmul(d.x[:, 3:6], rot)
            d.y[:, 0:3] = torch.matmul(d.y[:, 0:3], rot)
            for key in ("pos", "centroid", "depth_direction"):
                if key in d and getattr(d, key) is not None:
                    setattr(d, key, torch.matmul(getattr(d, key), rot))
        return data


_PT_FORMAT = "geobi-gnn_b200/dual-data/1"


def save_dual_data(dual_data, path):
    """The reference caches `torch.save((Data_v, Data_f), name.pt)` (dataset.py:154-155,183-184): a pickle of PyG objects.
    Written here as plain dicts of tensors / scalars (loads with weights_only=True, no class of this package inside)."""
    def plain(d):
        out = {}
        for k in d.keys:
            v = getattr(d, k)
            if torch.is_tensor(v):
                out[k] = v.detach().cpu()
            elif isinstance(v, (bool, int, float, str)):
                out[k] = v
            elif isinstance(v, np.generic):
                out[k] = v.item()
        return out
    torch.save({"format": _PT_FORMAT, "v": plain(dual_data[0]), "f": plain(dual_data[1])}, path)


def _bag_to_data(obj):
    """PyG `Data` as it comes out of a pickle whose torch_geometric classes were replaced by attribute bags: 1.x keeps the
    attributes in the instance dict, 2.x in `_store` -> `_mapping` [layouts recalled; no file of the authors' to check]."""
    state = dict(getattr(obj, "__dict__", {}))
    store = state.get("_store")
    if store is not None:
        inner = getattr(store, "__dict__", {})
        state = dict(inner.get("_mapping", inner))
    out = Data()
    for k, v in state.items():
        if k.startswith("_") or v is None:
            continue
        setattr(out, k, v)
    return out


def load_dual_data(path, device=None):
    """-> (Data_v, Data_f) from a cache file of this package (save_dual_data) or of the reference (pickled PyG tuple).
    The second kind needs Python's unpickler: torch_geometric classes are mapped to empty attribute bags and every other
    global goes through torch's own allow-list, so a cache file cannot import arbitrary code."""
    import pickle
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError:
        obj = _load_pyg_pickle(path)
    if isinstance(obj, dict) and obj.get("format") == _PT_FORMAT:
        dual = tuple(Data(**obj[k]) for k in ("v", "f"))
    elif isinstance(obj, (tuple, list)) and len(obj) == 2:
        dual = tuple(d if isinstance(d, Data) else _bag_to_data(d) for d in obj)
    else:
        raise ValueError(f"{path}: not a dual-data cache file")
    if "scale" in dual[0] and not isinstance(dual[0].scale, float):                # tensor / numpy scalar in upstream's files
        dual[0].scale = float(dual[0].scale)
    if device is not None:
        for d in dual:
            d.to(device)
    return dual


def _load_pyg_pickle(path):
    import pickle

    class _Bag:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, state):
            if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):      # (dict, slots) form
                state = {**(state[0] or {}), **state[1]}
            self.__dict__.update(state if isinstance(state, dict) else {"_state": state})

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] in ("torch_geometric", "torch_sparse"):
                return type(name, (_Bag,), {})
            if (module, name) in (("collections", "OrderedDict"), ("torch._utils", "_rebuild_tensor_v2"), ("torch", "Size"),
                                  ("torch._utils", "_rebuild_parameter"), ("numpy.core.multiarray", "scalar"),
                                  ("numpy._core.multiarray", "scalar"), ("numpy", "dtype"), ("_codecs", "encode")) or (
                    module == "torch" and name.endswith(("Storage", "Tensor"))) or (module == "torch" and name in _TORCH_DTYPES):
                return super().find_class(module, name)
            raise pickle.UnpicklingError(f"{path}: global {module}.{name} is not allowed in a dual-data cache file")

    class _Shim:                                                     # torch.load's pickle_module protocol
        Unpickler = _Unpickler
        __name__ = "pickle"

        @staticmethod
        def load(f, **kw):
            return _Unpickler(f, **kw).load()

    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_Shim)


_TORCH_DTYPES = frozenset(n for n in dir(torch) if isinstance(getattr(torch, n), torch.dtype))


class DualDataset:
    """dataset.py:72-283: `<root>/<data_type>/<train|test>/{original/<name>.obj, noisy/<name>_n*.obj}` -> one cached
    `processed_data/<name>[-sub<size>-<seed>].pt` per (sub)mesh, `get(i)` = load + post_processing.  `len()` / `get()` /
    `processed_dir` / `process_one_data` keep their names; indexing applies `transform` as the PyG base class does.
    `device`: where the graphs are built (the C-ABI builders need a GPU) and where samples are returned."""

    def __init__(self, data_type, train_or_test="train", data_list_txt=None, filter_patch_count=0, submesh_size=sys.maxsize,
                 transform=None, root=None, device="cuda"):
        self.data_type = data_type
        self.root_dir = os.path.join(DATASET_DIR if root is None else root, data_type)
        self.data_dir = os.path.join(self.root_dir, train_or_test)
        self.filter_patch_count = filter_patch_count
        self.submesh_size = submesh_size
        self.processed_folder = "processed_data"
        self.transform = transform
        self.device = device
        self.processed_files, self.files_noisy, self.files_original = [], [], []
        noisy_dir, original_dir = os.path.join(self.data_dir, "noisy"), os.path.join(self.data_dir, "original")
        if data_list_txt is not None:
            data_list = list(filter(None, [line.strip() for line in open(os.path.join(self.root_dir, data_list_txt))]))
        else:
            data_list = [os.path.basename(d)[:-4] for d in sorted(glob.glob(os.path.join(original_dir, "*.obj")))]
        for name in data_list:
            for name_n in sorted(glob.glob(os.path.join(noisy_dir, f"{name}_n*.obj"))):
                self.files_noisy.append(name_n)
                self.files_original.append(os.path.join(original_dir, f"{name}.obj"))
        self.process_data()

    @property
    def processed_dir(self):
        return os.path.join(self.data_dir, self.processed_folder)

    def process_data(self):
        os.makedirs(self.processed_dir, exist_ok=True)
        for noisyfile, originalfile in zip(self.files_noisy, self.files_original):
            self.process_one_data(noisyfile, self.submesh_size, originalfile=originalfile, obj=self, device=self.device)

    @staticmethod
    def process_one_data(noisyfile, submesh_size, originalfile=None, obj=None, device="cuda"):
        """dataset.py:129-193 -> [(dual_data, V_idx, select_faces), ...] for the (sub)meshes that had to be built (cached ones
        are only registered in obj.processed_files, as upstream)."""
        from . import meshio, patches, synth, topology
        filter_patch_count = 0 if obj is None else obj.filter_patch_count
        on_gpu = torch.device(device).type == "cuda"
        make = (lambda p, f: topology.DeviceTriMesh(p, f, device)) if on_gpu else synth.TriMesh
        points_noisy, fv = meshio.read_obj(noisyfile)
        points_noisy = points_noisy.astype(np.float32)
        points_original = None if originalfile is None else meshio.read_obj(originalfile)[0].astype(np.float32)
        base = os.path.basename(noisyfile)[:-4]
        all_dual_data = []

        mesh_n = make(points_noisy, fv)
        centroid, scale = normalisation(points_noisy, mesh_n.ev)                    # of the WHOLE noisy mesh (dataset.py:140)

        def build(filename, mesh_sub, pts_o, faces, v_idx, sel):
            pro_name = None
            if obj is not None:
                pro_name = os.path.join(obj.processed_dir, f"{filename}.pt")
                obj.processed_files.append(pro_name)
                if os.path.exists(pro_name):
                    return
            dual = process_one_submesh(mesh_sub, filename, None if pts_o is None else make(pts_o, faces), device)
            dual[0].centroid = torch.from_numpy(centroid).float().to(dual[0].pos.device)
            dual[0].scale = float(scale)
            all_dual_data.append((dual, v_idx, sel))
            if pro_name is not None:
                save_dual_data(dual, pro_name)
                if v_idx is not None:                                                # "save for visualization" (dataset.py:185-186)
                    meshio.write_obj(os.path.join(obj.processed_dir, f"{filename}.obj"), _host_array(mesh_sub.points), faces)

        if fv.shape[0] <= submesh_size:
            build(base, mesh_n, points_original, fv, None, None)
            return all_dual_data
        slot = np.full(points_noisy.shape[0], -1, dtype=np.int64)
        for sel, seed in patches.split_mesh(points_noisy, fv, _host_array(mesh_n.vf), submesh_size, filter_patch_count):
            v_idx, faces = patches.get_submesh(fv, sel, _slot=slot)
            build(f"{base}-sub{submesh_size}-{seed}", make(points_noisy[v_idx], faces),
                  None if points_original is None else points_original[v_idx], faces, v_idx, sel)
        return all_dual_data

    def len(self):
        return len(self.processed_files)

    __len__ = len

    def get(self, idx):
        return post_processing(load_dual_data(self.processed_files[idx], self.device), self.data_type)

    def __getitem__(self, idx):
        data = self.get(idx)
        return data if self.transform is None else self.transform(data)

    def __iter__(self):
        return (self[i] for i in range(len(self)))
