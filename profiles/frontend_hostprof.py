import os, sys, time, numpy as np, torch, cProfile, pstats
sys.path.insert(0, os.getcwd())
from geobi_gnn_b200 import data_util, dataset, ops, synth, topology
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
p, f = synth.icosphere(224)
pts = torch.from_numpy(p.astype(np.float32)).to(dev); fv = torch.from_numpy(f).to(dev)
def work():
    mesh = topology.DeviceTriMesh(pts, fv, dev)
    dual = dataset.process_one_submesh(mesh, "m", None, dev)
    torch.cuda.synchronize()
    return dual
for _ in range(3): work()
t0 = time.perf_counter(); work(); print("whole ms", 1e3 * (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable(); work(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
