"""Per-phase time of one configs[3] patch (10 M-face mesh, 1 M-face BFS patches on the device): cut-out, patch topology, graphs +
features, forward, stitch - each phase synchronised, five patches."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import config, dataset, inference, network, patches, topology
config.set_precision("bf16x3")
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
mesh = bench.noisy_device_mesh(bench.BIG_FREQ, 0, dev)
norm, cen = inference.device_normalisation(mesh)
parts = inference.partition(mesh, bench.BIG_SUB, centroid=cen)
torch.manual_seed(0)
net = network.DualGNN().to(dev).eval()
st = patches.Stitcher(mesh.n_vertices, mesh.n_faces, dev)
tot = {}
def lap(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    tot[name] = tot.get(name, 0.0) + (t - t0)
    return t
for rep in range(2):
    tot.clear()
    for k in range(5):
        sel, seed = parts[k]
        torch.cuda.synchronize(); t = time.perf_counter()
        v_idx, faces = patches.get_submesh_device(mesh.fv, sel, mesh.n_vertices); t = lap("cut-out (get_submesh_device)", t)
        sub = topology.DeviceTriMesh(mesh.points.index_select(0, v_idx), faces, dev); t = lap("patch topology (DeviceTriMesh)", t)
        dual = dataset.process_one_submesh(sub, "p", None, dev, csr_native=True)
        dataset.attach_normalisation(dual, None, None, precomputed=norm)
        dual = dataset.post_processing(dual, "Synthetic"); t = lap("graphs + features", t)
        with torch.no_grad():
            vp, nrm, _ = net([dual[0], dual[1]])
        t = lap("forward", t)
        st.add(vp, nrm, v_idx, sel.long()); t = lap("stitch", t)
print({k: round(1e3 * v / 5, 2) for k, v in tot.items()}, "ms per patch (synchronised phases)")
