"""Time feast_fused_64_32_kernel alone (bench shape: facet graph, N=512000, 64->32) and check it against the fp32 path."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, dataset, ops, nn as gnn
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(int(os.environ.get("PATCHES", bench.N_PATCHES)), 0)]
dv, df, _ = batching.collate_dual(patches)
n = df.x.size(0)
g = ops.csr_from_coo(df.edge_index, n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
torch.manual_seed(0)
conv = gnn.FeaStConv(64, 32, 9).to(dev)
x = torch.randn(n, 64, device=dev)
out = torch.empty(n, 32, device=dev)
P = (conv.lin.weight.data, conv.u.weight.data, conv.c.data, conv.bias.data)
ref = ops.feast_fwd(x, g, *P, 0.2, precision=ops.PREC_FP32)
got = ops.feast_fwd(x, g, *P, 0.2, out=out, precision=ops.PREC_BF16X3)
torch.cuda.synchronize()
err = ((got - ref).abs().max() / ref.abs().max()).item()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(12):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.feast_fwd(x, g, *P, 0.2, out=out, precision=ops.PREC_BF16X3 | ops.FEAST_REUSE_WS)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
got2 = out.clone()
print(f"N={n} nnz={g.nnz} rel_err_vs_fp32={err:.2e} same_after_reuse={torch.equal(got, got2)} fused_ms: min {min(ts):.4f} mean {np.mean(ts[2:]):.4f}")
