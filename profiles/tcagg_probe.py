"""Bring-up / timing probe of feast_tcagg_64_32_kernel (round 2): correctness on small and irregular graphs against the
fp32 CUDA-core path, then the bench shape (facet graph of 64 patches x 8000 faces, N = 512000) timed alone with the L2 flushed.
GEOBI_LIB_PATH=build_variants/libgeobi_tcagg_dbg.so loads the -DTCAGG_DEBUG build (bounded waits: a protocol bug prints
the stuck warps' wait tags instead of hanging the box).  STAGE=small|big|all."""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geobi_gnn_b200 import _lib, batching, dataset, ops, synth, nn as gnn
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
stage = os.environ.get("STAGE", "all")
TAGS = {1: "mma:full", 2: "mma:dfree", 3: "mma:zfull", 4: "mma:ofree", 5: "epi:ofull", 6: "drain:dfull", 7: "drain:zfree", 8: "prod:xfree"}


def debug_dump(label):
    if not hasattr(lib, "geobi_debug_tcagg"):
        return False
    buf = (ctypes.c_uint * 64)()
    torch.cuda.synchronize()
    lib.geobi_debug_tcagg(buf)
    v = np.array(buf[:], dtype=np.uint32)
    if v[63]:
        print(f"[{label}] STUCK WAITS (cta0 warps 0..31 | another cta):")
        for i in range(63):
            if v[i]:
                print(f"   {'cta0' if i < 32 else 'ctaX'} warp {i % 32}: {TAGS.get(int(v[i]) & 0xffff, '?')} parity {(int(v[i]) >> 16) & 1}")
        return True
    return False


def run_case(label, n, ei, x, row_map=None, slope=0.2, seed=0):
    torch.manual_seed(seed)
    conv = gnn.FeaStConv(64, 32, 9).to(dev)
    P = (conv.lin.weight.data, conv.u.weight.data, conv.c.data, conv.bias.data)
    g = ops.csr_from_coo(ei.to(dev), n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
    rm = None if row_map is None else row_map.to(dev).int()
    ref = ops.feast_fwd(x, g, *P, slope, precision=ops.PREC_FP32, row_map=rm)
    got = ops.feast_fwd(x, g, *P, slope, precision=ops.PREC_BF16X3, row_map=rm)
    torch.cuda.synchronize()
    stuck = debug_dump(label)
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    bad = (~torch.isfinite(got)).sum().item()
    rows_bad = ((got - ref).abs().max(1).values > 1e-4 * ref.abs().max()).nonzero().flatten()
    print(f"[{label}] n={n} nnz={g.nnz} maxdeg={int((g.rowptr[1:] - g.rowptr[:-1]).max())} rel_err={err:.3e} nonfinite={bad} bad_rows={rows_bad.numel()}"
          + (f" first bad rows {rows_bad[:12].tolist()}" if rows_bad.numel() else ""), flush=True)
    if stuck:
        sys.exit(3)
    return err, g, P


if stage in ("small", "all"):
    for freq, seed in ((2, 0), (6, 1), (12, 2)):
        p, f = synth.icosphere(freq)
        m = synth.TriMesh(synth.add_normal_noise(p, f, 0.2, seed), f)
        dv, df = dataset.build_dual_data(m, synth.TriMesh(p, f), device=dev)
        for name, d in (("v", dv), ("f", df)):
            n = d.x.size(0)
            torch.manual_seed(seed)
            run_case(f"ico{freq}-{name}", n, d.edge_index, torch.randn(n, 64, device=dev) * 2.0)
    # irregular: hubs of degree 120 / 40 / 17 (multi-round nodes), isolated nodes
    torch.manual_seed(5)
    n = 3000
    src, dst = torch.randint(0, n - 10, (24000,)), torch.randint(0, n - 10, (24000,))
    hubs = torch.cat([torch.full((120,), 0), torch.full((40,), 1), torch.full((17,), 2)])
    ei = torch.stack([torch.cat([src, hubs]), torch.cat([dst, torch.randint(3, n - 10, (177,))])])
    ei = ei[:, ei[0] != ei[1]]
    ei = torch.unique(torch.cat([ei, ei.flip(0)], 1), dim=1)
    run_case("hubs", n, ei, torch.randn(n, 64, device=dev) * 2.0)
    nc = n // 3
    run_case("hubs+row_map", n, ei, torch.randn(nc, 64, device=dev) * 2.0, row_map=torch.randint(0, nc, (n,)), slope=1.0)
    run_case("tiny", 5, torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]]), torch.randn(5, 64, device=dev))

if stage in ("big", "all"):
    import bench
    patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(int(os.environ.get("PATCHES", bench.N_PATCHES)), 0)]
    dv, df, _ = batching.collate_dual(patches)
    for name, d in (("facet", df), ("vertex", dv)):
        n = d.x.size(0)
        torch.manual_seed(0)
        x = torch.randn(n, 64, device=dev)
        err, g, P = run_case(f"bench-{name}", n, d.edge_index, x)
        out = torch.empty(n, 32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        res = {}
        for label, env in (("tcagg", None), ("fp32agg", "1")):
            if env:
                os.environ["GEOBI_NO_TCAGG"] = env
            else:
                os.environ.pop("GEOBI_NO_TCAGG", None)
            ops.feast_fwd(x, g, *P, 0.2, out=out, precision=ops.PREC_BF16X3)
            tk, tl = [], []
            for _ in range(12):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.feast_fwd(x, g, *P, 0.2, out=out, precision=ops.PREC_BF16X3 | ops.FEAST_REUSE_WS)
                b.record()
                torch.cuda.synchronize()
                tk.append(a.elapsed_time(b))
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.feast_fwd(x, g, *P, 0.2, out=out, precision=ops.PREC_BF16X3)
                b.record()
                torch.cuda.synchronize()
                tl.append(a.elapsed_time(b))
            res[label] = (float(np.mean(tk[2:])), float(np.min(tk)), float(np.mean(tl[2:])))
        os.environ.pop("GEOBI_NO_TCAGG", None)
        alg = bench.feast_bytes_alg(n, g.nnz + n, 64, 32)
        for label, (mean, mn, layer) in res.items():
            print(f"[bench-{name}] {label}: kernel mean {mean:.4f} ms (min {mn:.4f}), whole layer {layer:.4f} ms, "
                  f"algorithmic {alg / 1e6:.1f} MB -> {alg / mean / 1e6:.1f} GB/s = {alg / mean / 1e6 / bench.peaks()[0]:.4f} of peak", flush=True)
