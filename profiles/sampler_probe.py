"""Does the NVML clock sampler perturb the timed region?  Alternates 20-step blocks with / without it in one process."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, config, dataset, network
config.set_precision("bf16x3")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))); torch.cuda.set_device(dev)
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(bench.N_PATCHES, 0)]
dv, df, _ = batching.collate_dual(patches)
torch.manual_seed(0)
net = network.DualGNN().to(dev).eval()
def block(k=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k):
        with torch.no_grad():
            net([batching.fresh_view(dv), batching.fresh_view(df)])
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
block(5)
for rep in range(3):
    a = block()
    with bench.ClockSampler(dev.index) as c:
        b = block()
    print(f"rank {dev.index}: without sampler {a:.2f} ms/step, with sampler {b:.2f} ms/step ({len(c.rows)} samples)")
