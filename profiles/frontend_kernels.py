"""GPU kernels of the device front end (what HostBatchRunner.upload_mesh queues on the copy stream per configs[2] mesh):
torch.profiler totals per call, warm."""
import os, sys, collections, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, dataset, topology, nn as gnn
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
mesh0 = bench.noisy_device_mesh(bench.MESH_FREQ, 0, dev)
pts, fcs = mesh0.points.clone(), mesh0.fv.to(torch.int32).clone()
def front():
    mesh = topology.DeviceTriMesh(pts, fcs.long(), dev)
    if os.environ.get("LISTS"):      # the round-trip path: reference-layout lists, then sort-free CSRs from them
        dv, df = dataset.build_dual_on_device(mesh, None, "Synthetic")
        dv, df = batching.fresh_view(dv), batching.fresh_view(df)
        for d in (dv, df):
            gnn.input_graph(d, d.x.size(0))
    else:                            # what upload_mesh queues: CSRs handed over directly, lists lazy
        dv, df = dataset.build_dual_on_device(mesh, None, "Synthetic", csr_native=True)
    return dv, df
for _ in range(3):
    front()
torch.cuda.synchronize()
REPS = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(REPS):
        front()
    torch.cuda.synchronize()
tot = collections.Counter(); cnt = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name[:70]] += e.device_time; cnt[e.name[:70]] += 1
all_us = sum(tot.values()) / REPS
print(f"GPU time per front-end call: {all_us:.1f} us over {sum(cnt.values()) // REPS} kernels / copies")
for k, v in tot.most_common(40):
    print(f"{v / REPS:9.1f} us  n={cnt[k] // REPS:3d}  {k}")
