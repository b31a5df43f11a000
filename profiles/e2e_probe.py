import os, sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from geobi_gnn_b200 import batching, config, dataset, network
from geobi_gnn_b200.data import Data
config.set_precision("bf16x3")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))); torch.cuda.set_device(dev)
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(bench.N_PATCHES, 0)]
dv, df, _ = batching.collate_dual(patches)
torch.manual_seed(0)
net = network.DualGNN().to(dev).eval()
kv, kf = ("x", "edge_index", "edge_weight"), ("x", "edge_index", "edge_weight", "fv_indices")
hv = {k: getattr(dv, k).cpu().pin_memory() for k in kv}
hf = {k: getattr(df, k).cpu().pin_memory() for k in kf}
def step(copy=True):
    t0 = time.perf_counter()
    a = Data(**{k: t.to(dev, non_blocking=True) for k, t in hv.items()}) if copy else batching.fresh_view(dv)
    b = Data(**{k: t.to(dev, non_blocking=True) for k, t in hf.items()}) if copy else batching.fresh_view(df)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    with torch.no_grad():
        vp, nrm, _ = net([a, b])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3
for i in range(4): print('resident', [round(x, 2) for x in step(False)])
for i in range(6): print('e2e', [round(x, 2) for x in step(True)])
print(torch.cuda.memory_stats()['num_alloc_retries'], torch.cuda.memory_reserved() / 1e9)
