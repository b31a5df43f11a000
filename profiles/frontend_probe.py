"""Where the device front end spends its time at 1 M faces (configs[2] shape): topology.DeviceTriMesh stages, dataset.process_one_submesh
stages, normalisation, post_processing.  Each stage synchronised and timed (second run, warm)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geobi_gnn_b200 import data_util, dataset, ops, synth, topology
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 224
p, f = synth.icosphere(n)
pts = torch.from_numpy(p.astype(np.float32)).to(dev); fv = torch.from_numpy(f).to(dev)
def T(label, fn, reps=3):
    out = fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    print(f"{label:46s} {1e3 * (time.perf_counter() - t0) / reps:8.2f} ms", flush=True)
    return out
V, F = pts.size(0), fv.size(0)
mesh = T("DeviceTriMesh (whole)", lambda: topology.DeviceTriMesh(pts, fv, dev))
h = torch.stack([fv.reshape(-1), fv[:, [1, 2, 0]].reshape(-1)])
g = T("  vertex csr_from_coo (symm+sort+dedup)", lambda: ops.csr_from_coo(h, V, None, ops.COO_SYMMETRIZE | ops.COO_SORT_NBR | ops.COO_DEDUP | ops.COO_DROP_SELF))
ei = T("  vertex_csr.edge_index()", lambda: ops.CSRGraph(g.rowptr, g._nbr, g.n, g.nnz).edge_index())
T("  ev = upper triangle", lambda: ei[:, ei[0] < ei[1]].t().contiguous())
mr = T("  group_by (faces around vertices)", lambda: ops.group_by(fv.reshape(-1).to(torch.int32), V))
def pad():
    mrowptr, members = mr
    cnt = (mrowptr[1:] - mrowptr[:-1]).long(); k = int(cnt.max())
    rows = torch.repeat_interleave(torch.arange(V, device=dev), cnt)
    cols = torch.arange(members.numel(), device=dev) - torch.repeat_interleave(mrowptr[:-1].long(), cnt)
    vf = torch.full((V, k), -1, dtype=torch.int64, device=dev); vf[rows, cols] = members.div(3, rounding_mode="floor").long(); return vf
T("  vf padded table (eager glue)", pad)
T("  normals (face + vertex)", mesh.update_normals)
dual = T("process_one_submesh (whole)", lambda: dataset.process_one_submesh(mesh, "m", None, dev))
T("  with_self_loops_appended", lambda: data_util.with_self_loops_appended(mesh.vertex_csr))
T("  build_facet_graph", lambda: data_util.build_facet_graph(mesh.fv, mesh.vf))
T("  calc_weight facet", lambda: data_util.calc_weight(dual[1].pos, dual[1].normal, dual[1].edge_index))
T("  calc_weight vertex", lambda: data_util.calc_weight(dual[0].pos, dual[0].normal, dual[0].edge_index))
T("  build_edge_fv", lambda: data_util.build_edge_fv(mesh.fv))
T("normalisation (host)", lambda: dataset.normalisation(mesh.points, mesh.ev), reps=1)
