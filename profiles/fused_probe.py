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

import ctypes
lib = ctypes.CDLL(os.environ["GEOBI_LIB_PATH"]) if os.environ.get("GEOBI_LIB_PATH") else None
if lib is not None and hasattr(lib, "geobi_debug_fused_timeline"):
    buf = (ctypes.c_ulonglong * 512)()
    lib.geobi_debug_fused_timeline(buf)
    tl = np.array(buf[:], dtype=np.int64).reshape(64, 8)
    t0 = tl[0, 4]
    print("per-tile timeline of CTA 0 (us, relative): mma: z_ready, issued, done, drained | agg warp0: start, agg_end, mbar_ok, z_stored")
    for r in range(0, 24):
        print(" ".join(f"{(v - t0) / 1e3:8.2f}" for v in tl[r]))
    d = np.diff(tl[:, 4]) / 1e3
    print("tile period us: mean %.2f" % d.mean(), " mma issue->done %.2f" % ((tl[:, 2] - tl[:, 0]).mean() / 1e3), " drain %.2f" % ((tl[:, 3] - tl[:, 2]).mean() / 1e3),
          " agg %.2f" % ((tl[:, 5] - tl[:, 4]).mean() / 1e3), " mbar wait %.2f" % ((tl[:, 6] - tl[:, 5]).mean() / 1e3), " z store %.2f" % ((tl[:, 7] - tl[:, 6]).mean() / 1e3))
    if hasattr(lib, "geobi_debug_fused_warps"):
        b2 = (ctypes.c_ulonglong * 2048)()
        lib.geobi_debug_fused_warps(b2)
        tw = np.array(b2[:], dtype=np.int64).reshape(64, 16, 2)
        dur = (tw[:, :, 1] - tw[:, :, 0]) / 1e3            # per-warp tile duration (start -> rows stored)
        rel = (tw[:, :, 1] - tl[:, 0:1]) / 1e3              # arrival relative to z_ready of the same tile (<= 0)
        print("per-warp mean tile duration us:", np.round(dur.mean(0), 2))
        print("per-warp mean arrival before z_ready us:", np.round(rel.mean(0), 2))
        last = rel.argmax(1)
        print("last-arriving warp histogram:", np.bincount(last, minlength=16))
    if hasattr(lib, "geobi_debug_fused_stages"):
        b3 = (ctypes.c_ulonglong * 1024)()
        lib.geobi_debug_fused_stages(b3)
        ts = np.array(b3[:], dtype=np.int64).reshape(64, 2, 8)[:, :, :7]
        d = np.diff(ts, axis=2) / 1e3
        names = ["idx+first loads", "P loads+softmax", "syncwarp+pair loop", "mbar wait", "z store", "fence+arrive"]
        for w, nm in ((0, "warp 0"), (1, "warp 15")):
            print(nm, " ".join(f"{n}: {v:.2f}" for n, v in zip(names, d[:, w, :].mean(0))), " total %.2f" % d[:, w, :].sum(1).mean())
