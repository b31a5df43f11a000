"""Time the first-layer FeaSt conv (C_in = 12 -> 32 on the facet graph, 6 -> 32 on the vertex graph) kernel by kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, dataset, ops, nn as gnn
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(bench.N_PATCHES, 0)]
dv, df, _ = batching.collate_dual(patches)
from torch.profiler import profile, ProfilerActivity
for d, cin in ((df, 12), (dv, 6)):
    n = d.x.size(0)
    g = ops.csr_from_coo(d.edge_index, n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
    torch.manual_seed(0)
    conv = gnn.FeaStConv(cin, 32, 9).to(dev)
    x = torch.randn(n, cin, device=dev)
    P = (conv.lin.weight.data, conv.u.weight.data, conv.c.data, conv.bias.data)
    ref = ops.feast_fwd(x, g, *P, 0.2, precision=ops.PREC_FP32)
    for _ in range(3): got = ops.feast_fwd(x, g, *P, 0.2, precision=ops.PREC_BF16X3)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ops.feast_fwd(x, g, *P, 0.2, precision=ops.PREC_BF16X3); torch.cuda.synchronize()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    print(f"C_in={cin} N={n} nnz={g.nnz} err={err:.2e}: " + "; ".join(f"{e.key.split('(')[0][-40:]} {e.device_time_total:.0f}us" for e in prof.key_averages()))
