"""Standalone timings of the native backward calls at the configs[4] shape (16 x 8000-face patches: 128 000 facet nodes):
geobi_feast_bwd for the level-0 layer shapes and geobi_mlp_head_bwd, CUDA events, L2 flushed between calls.
Run under `ncu -k regex:"dw_splitk|head_bwd_hidden"` for the counters of the two new tcgen05 kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, config, dataset, ops
from geobi_gnn_b200.autograd import feast_bwd, mlp_head_bwd
from geobi_gnn_b200.nn import input_graph
config.set_precision("bf16x3")
dev = torch.device("cuda")
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(16, 0)]
dv, df, _ = batching.collate_dual(patches)
n = df.x.size(0)
g = input_graph(df, n)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print(f"facet graph: N = {n}, nnz = {int(g.rowptr[-1])}")
only = os.environ.get("PROBE_ONLY", "")          # "head": the FC-head backward alone (for a short ncu capture)
for cin, cout in (() if only == "head" else ((12, 32), (64, 32), (32, 64))):
    x = torch.randn(n, cin, device=dev)
    W, U, c = torch.randn(9 * cout, cin, device=dev) * 0.1, torch.randn(9, cin, device=dev) * 0.1, torch.randn(9, device=dev) * 0.1
    out = ops.feast_fwd(x, g, W, U, c, torch.zeros(cout, device=dev), act_slope=0.2, precision=ops.PREC_BF16X3)
    go = torch.randn(n, cout, device=dev)
    ms = timed(lambda: feast_bwd(x, g, W, U, c, out, go, 0.2, cin > 16))
    zbytes = n * (-(-9 * cin // 64) * 64) * 4
    print(f"geobi_feast_bwd {cin:3d} -> {cout:3d}: {ms:.3f} ms  (Z planes {zbytes / 1e6:.0f} MB, dZ {zbytes / 1e6:.0f} MB)")
f = torch.randn(n, 32, device=dev)
W1, b1, W2 = torch.randn(1024, 32, device=dev) * 0.1, torch.randn(1024, device=dev) * 0.1, torch.randn(3, 1024, device=dev) * 0.1
dy = torch.randn(n, 3, device=dev)
ms = timed(lambda: mlp_head_bwd(f, W1, b1, W2, dy))
print(f"geobi_mlp_head_bwd N = {n}: {ms:.3f} ms  (a + dh planes {n * 1024 * 8 / 1e6:.0f} MB written, {n * 1024 * 12 / 1e6:.0f} MB read back)")
