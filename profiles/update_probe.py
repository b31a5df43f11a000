"""60-sweep vertex update (data_util.update_position2) on the configs[3] mesh (10 M faces) and on a 1 M-face mesh: ms per call."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import data_util
dev = torch.device("cuda", 0)
for freq in (224, bench.BIG_FREQ):
    mesh = bench.noisy_device_mesh(freq, 0, dev)
    fn = mesh.face_normals
    for _ in range(2):
        out = data_util.update_position2(mesh.points, mesh.fv, mesh.vf, fn, 60)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        out = data_util.update_position2(mesh.points, mesh.fv, mesh.vf, fn, 60)
    b.record()
    torch.cuda.synchronize()
    print(f"faces {mesh.n_faces}: update_position2 x60 = {a.elapsed_time(b) / 5:.3f} ms  (K = {mesh.vf.size(1)}, finite {bool(torch.isfinite(out).all())})")
    del mesh, out
    torch.cuda.empty_cache()
