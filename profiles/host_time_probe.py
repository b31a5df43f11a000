"""Host-side (launch) time of one configs[2] forward and of one upload_mesh call, without device synchronisation in between:
the end-to-end step is GPU-bound as long as their sum stays under the GPU time of the step (12.7 ms)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import config, inference, network
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
config.set_precision("bf16x3")
mesh = bench.noisy_device_mesh(bench.MESH_FREQ, 0, dev)
hp = mesh.points.cpu().pin_memory(); hf = mesh.fv.to(torch.int32).cpu().pin_memory()
torch.manual_seed(0)
net = network.DualGNN().to(dev).eval()
runner = inference.HostBatchRunner(net, dev, coalesced_undirected=True)
nxt = runner.upload_mesh(hp, hf)
for _ in range(5):
    cur = nxt; nxt = runner.upload_mesh(hp, hf); runner.run(cur)
torch.cuda.synchronize()
tu, tr = [], []
for _ in range(20):
    cur = nxt
    t0 = time.perf_counter(); nxt = runner.upload_mesh(hp, hf); t1 = time.perf_counter()
    runner.run(cur); t2 = time.perf_counter()
    tu.append(t1 - t0); tr.append(t2 - t1)
torch.cuda.synchronize()
print(f"upload_mesh host time: median {1e3 * np.median(tu):.2f} ms (min {1e3 * min(tu):.2f}); run() host time (includes waiting on the count read-backs): "
      f"median {1e3 * np.median(tr):.2f} ms (min {1e3 * min(tr):.2f})")
