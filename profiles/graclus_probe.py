"""Time the matcher (geobi_graclus) on the bench's level-1 graphs with feature edge weights."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, dataset, ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(int(os.environ.get("PATCHES", bench.N_PATCHES)), 0)]
dv, df, _ = batching.collate_dual(patches)
for name, d in (("facet", df), ("vertex", dv)):
    n = d.x.size(0)
    g, _, _ = ops.csr_from_sorted_coo(d.edge_index, n, d.edge_weight)
    g.nnz
    torch.manual_seed(0)
    keys = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, dtype=torch.int64).to(torch.int32)
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        label, _ = ops.graclus(g, None, keys=keys)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    nc = int((label == torch.arange(n, device=dev, dtype=label.dtype)).sum())
    print(f"{name}: n={n} nnz={g.nnz} clusters={nc} checksum={int(label.long().sum())} graclus ms: min {min(ts):.4f} mean {np.mean(ts[2:]):.4f}")
