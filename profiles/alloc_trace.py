"""Per-step cudaMalloc activity of the caching allocator during the first forwards (why warm-up takes > 3 steps)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, config, dataset, network
config.set_precision("bf16x3")
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(bench.N_PATCHES, 0)]
dv, df, _ = batching.collate_dual(patches)
torch.manual_seed(0)
net = network.DualGNN().to(dev).eval()
prev = torch.cuda.memory_stats().get("num_device_alloc", 0)
for i in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        net([batching.fresh_view(dv), batching.fresh_view(df)])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    st = torch.cuda.memory_stats()
    cur = st.get("num_device_alloc", 0)
    print(f"step {i:2d} {dt:7.2f} ms  cudaMallocs +{cur - prev}  reserved {st['reserved_bytes.all.current']/1e9:.2f} GB  retries {st['num_alloc_retries']}")
    prev = cur
