"""Step-time distribution of the bench workload under different caching-allocator settings."""
import os, sys, time
conf = sys.argv[1] if len(sys.argv) > 1 else ""
if conf:
    os.environ["PYTORCH_CUDA_ALLOC_CONF"] = conf
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, config, dataset, network
config.set_precision("bf16x3")
dev = torch.device("cuda")
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(bench.N_PATCHES, 0)]
dv, df, _ = batching.collate_dual(patches)
torch.manual_seed(0)
net = network.DualGNN().to(dev).eval()
ts = []
for i in range(40):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        net([batching.fresh_view(dv), batching.fresh_view(df)])
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
ts = ts[5:]
import statistics
print(f"{conf or 'default':45s} median {statistics.median(ts):6.2f} mean {statistics.mean(ts):6.2f} max {max(ts):6.2f} ms; reserved {torch.cuda.memory_reserved()/1e9:.2f} GB")
# back-to-back (no per-step sync), as bench.py times it
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(10):
        with torch.no_grad():
            net([batching.fresh_view(dv), batching.fresh_view(df)])
    torch.cuda.synchronize(); print(f"back-to-back x10: {(time.perf_counter() - t0) * 100:.2f} ms/step")
