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
os.environ["GEOBI_TCAGG"] = "1"      # this probe is about the tcgen05-aggregation kernel (opt-in in the library)
TAGS = {1: "mma:full", 2: "mma:dfree", 3: "mma:zfull", 4: "mma:ofree", 5: "epi:ofull", 6: "drain:dfull", 7: "drain:zfree", 8: "prod:xfree"}


def debug_dump(label):
    if not hasattr(lib, "geobi_debug_tcagg"):
        return False
    buf = (ctypes.c_uint * 64)()
    torch.cuda.synchronize()
    lib.geobi_debug_tcagg(buf)
    v = np.array(buf[:], dtype=np.uint32)
    if v[63]:
        print(f"[{label}] STUCK: cta {int(v[63]) - 1}; per warp (tag, parity, barrier word offset >> 3):")
        for i in range(32):
            if v[i]:
                b = (int(v[i]) >> 8) & 0xffff
                names = [("full", 14), ("xfree", 7), ("dfull", 4), ("dfree", 8), ("zfull", 2), ("zfree", 2), ("ofull", 2), ("ofree", 2)]
                nm = "?"
                for name, cnt in names:
                    if b < cnt:
                        nm = f"{name}[{b}]"
                        break
                    b -= cnt
                print(f"   warp {i}: {TAGS.get(int(v[i]) & 0xff, '?')} {nm} parity {int(v[i]) >> 31}")
        print(f"   cta0 progress: agg last seq+1 {int(v[32]) & 0xffff} (pair {int(v[32]) >> 16}), drain sets pairs done {int(v[34])}, {int(v[35])}, producers last seq+1 finished {[int(x) for x in v[40:50]]}")
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
        for label, env in (("tcagg", "1"), ("fp32agg", "0")):
            os.environ["GEOBI_TCAGG"] = env
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
        os.environ["GEOBI_TCAGG"] = "1"
        alg = bench.feast_bytes_alg(n, g.nnz + n, 64, 32)
        for label, (mean, mn, layer) in res.items():
            print(f"[bench-{name}] {label}: kernel mean {mean:.4f} ms (min {mn:.4f}), whole layer {layer:.4f} ms, "
                  f"algorithmic {alg / 1e6:.1f} MB -> {alg / mean / 1e6:.1f} GB/s = {alg / mean / 1e6 / bench.peaks()[0]:.4f} of peak", flush=True)

if hasattr(lib, "geobi_debug_tcagg_timeline"):
    # the last launch was the vertex graph with the fp32agg switch restored -> rerun one tcagg launch on the facet graph
    n = df.x.size(0)
    g = ops.csr_from_coo(df.edge_index, n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
    x = torch.randn(n, 64, device=dev)
    ops.feast_fwd(x, g, *P, 0.2, precision=ops.PREC_BF16X3)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (5 * 64 * 8))()
    lib.geobi_debug_tcagg_timeline(buf)
    tl = np.array(buf[:], dtype=np.int64).reshape(5, 64, 8)
    t0 = tl[1, 0, 0]
    np.set_printoptions(linewidth=220, suppress=True)
    def d(a, b):
        return (b - a)
    pr, mm, dr, ti = tl[0], tl[1], tl[2], tl[3]
    print("pair period (MMA issue to issue), clk: mean %.0f" % np.diff(mm[:, 3]).mean())
    print("producer per pair, clk: xfree wait %.0f | gather+P issue %.0f | begin->finish gap %.0f | softmax+q %.0f | cp.async wait %.0f | fence+arrive %.0f | total %.0f" % (
        d(pr[:, 0], pr[:, 1]).mean(), d(pr[:, 1], pr[:, 2]).mean(), d(pr[:, 2], pr[:, 3]).mean(), d(pr[:, 3], pr[:, 4]).mean(), d(pr[:, 4], pr[:, 5]).mean(),
        d(pr[:, 5], pr[:, 6]).mean(), d(pr[:, 0], pr[:, 6]).mean()))
    lp = tl[4]
    print("producer loop iteration that begins the pair, clk: next_item %.0f | early check %.0f | begin+finish(prev) %.0f | total %.0f | early taken %.0f %%" % (
        d(lp[:, 0], lp[:, 1]).mean(), d(lp[:, 1], lp[:, 2]).mean(), d(lp[:, 2], lp[:, 4]).mean(), d(lp[:, 0], lp[:, 4]).mean(), 100.0 * (lp[:, 3] != 0).mean()))
    print("mma per pair, clk: full wait %.0f | dfree wait %.0f | issue %.0f ; producer arrive -> mma sees full %.0f" % (
        d(mm[:, 0], mm[:, 1]).mean(), d(mm[:, 1], mm[:, 2]).mean(), d(mm[:, 2], mm[:, 3]).mean(), d(pr[:, 6], mm[:, 1]).mean()))
    print("drain per pair, clk: dfull wait %.0f | ld %.0f | zfree wait %.0f | split+store %.0f | fence+arrive %.0f | total busy %.0f ; mma issue -> drain sees dfull %.0f" % (
        d(dr[:, 0], dr[:, 1]).mean(), d(dr[:, 1], dr[:, 2]).mean(), d(dr[:, 2], dr[:, 3]).mean(), d(dr[:, 3], dr[:, 4]).mean(), d(dr[:, 4], dr[:, 5]).mean(),
        d(dr[:, 1], dr[:, 5]).mean(), d(mm[:, 3], dr[:, 1]).mean()))
    tt = ti[:56]
    print("tile, clk: period %.0f | zfull wait %.0f | ofree wait %.0f | proj issue %.0f | proj issue -> epilogue sees ofull %.0f | epilogue %.0f" % (
        np.diff(tt[:, 3]).mean(), d(tt[:, 0], tt[:, 1]).mean(), d(tt[:, 1], tt[:, 2]).mean(), d(tt[:, 2], tt[:, 3]).mean(), d(tt[:, 3], tt[:, 4]).mean(),
        d(tt[:, 4], tt[:, 5]).mean()))
    print("first 3 tiles of the window, relative clk per pair: [prod begin, xfree ok, issued, fin start, q done, cp done, arrived | mma full ok, dfree ok, issued | drain dfull ok, ld done, zfree ok, stored, arrived]")
    for i in range(48):
        print(i + 128, (pr[i, :7] - t0).tolist(), (mm[i, 1:4] - t0).tolist(), (dr[i, 1:6] - t0).tolist())
    print("tiles: [zfull wait start, zfull ok, ofree ok, proj issued, epi ofull ok, epi done]")
    for i in range(4):
        print(i + 8, (ti[i, :6] - t0).tolist())

if stage == "ncu":
    # short run for `ncu --set full -k regex:feast_tcagg_64`: facet graph of PATCHES patches, three launches of the layer
    import bench
    patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(int(os.environ.get("PATCHES", 32)), 0)]
    dv, df, _ = batching.collate_dual(patches)
    n = df.x.size(0)
    g = ops.csr_from_coo(df.edge_index, n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
    torch.manual_seed(0)
    conv = gnn.FeaStConv(64, 32, 9).to(dev)
    P = (conv.lin.weight.data, conv.u.weight.data, conv.c.data, conv.bias.data)
    x = torch.randn(n, 64, device=dev)
    for _ in range(3):
        out = ops.feast_fwd(x, g, *P, 0.2, precision=ops.PREC_BF16X3)
    torch.cuda.synchronize()
    print("ncu stage done", n, float(out.abs().max()))
