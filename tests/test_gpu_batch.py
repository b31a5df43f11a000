"""GPU: parity of the configurations the bench lines are quoted on.

* BASELINE configs[1]: a disjoint-union batch of 8000-face patches in the benchmarked numeric mode ('bf16x3').  The reference cannot
  batch (/root/reference/code/dataset.py:29-31: its Collater returns batch[0]; train_dual.py:142), so the oracle loops over the
  patches and the product's union forward is compared slice by slice, with the oracle's matchings teacher-forced
  (labels offset by each patch's node range at that pooling step, SURVEY.md 8c protocol item 3).
* One mid-size single mesh (200 000 faces) the same way: sizes between configs[0] (20 480) and configs[2] (1 M).

Tolerances: BASELINE.json north_star - 1e-5 max-norm relative for fp32, 2e-3 for a bf16 GEMM.  Measured worst cases (B200, written to
gpurun_out/parity_worst_cases.jsonl): precision 'fp32' (CUDA cores) stays below 1.5e-6 on every tap at every size; 'bf16x3'
(tensor-core projections, operands split into two bf16 halves: a 16-bit-mantissa representation, |x - hi - lo| <= 2^-18 |x|)
measures 7e-6 .. 9e-6 per layer on the 8000-face patches of the bench shape - inside the fp32 bar - and up to 2.1e-5 in max norm
over the 200 000-face mesh (the maximum runs over 25x more entries).  So 'bf16x3' is held to 1e-5 on the benchmarked shape and to
3e-5 on the large mesh (70x inside the bf16-GEMM bar that formally applies to it), 'fp32' to 1e-5 everywhere.  The unit normals
are compared as vectors (max |a - b|), see util.TOL_NORMAL.
"""
import json
import os

import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda"
TAPS = ("l1", "p1", "l2", "p2", "l3", "l4", "r1", "r2", "r3", "r4")
# unit normals: |d n| ~ |d y| / |y|; over hundreds of thousands of faces the shortest |y| is ~1e-2 of the typical one, which turns
# a 1e-7 absolute error of y into 1e-5 of n.  Measured worst cases (B200): see profiles/r02_NOTES.md section B.
NORMAL_TOL = util.TOL_NORMAL


def _record(name, payload):
    out = os.path.join(util.ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_worst_cases.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **payload}) + "\n")


def _oracle_loop(meshes, seed=0, perm_seed=77):
    """Oracle forward per patch (recording taps and matchings)."""
    from oracle import ref_dataset
    ref = util.oracle_net(seed)
    ref.record = True
    outs = []
    for i, (mn, mo) in enumerate(meshes):
        dv, df = ref_dataset.build_dual_data(mn, mo)
        util.set_perm_fn(ref, perm_seed + i)
        with torch.no_grad():
            vp, nrm, _ = ref([dv, df])
        outs.append(dict(vp=vp, nrm=nrm, taps={g: {k: ref.taps[g][k].clone() for k in TAPS} for g in ("v", "f")},
                         labels=[[t[3].clone() for t in pl.trace] for pl in util.poolings(ref)],
                         unpool=[pl.unpooling_indices.clone() for pl in util.poolings(ref)]))
    return ref, outs


def _union_forced(outs):
    """Per pooling layer, per step: the union's raw labels = each patch's labels shifted by the node offset of that patch at that step."""
    forced = []
    for li in range(4):
        steps = []
        for s in range(len(outs[0]["labels"][li])):
            off, parts = 0, []
            for o in outs:
                lab = o["labels"][li][s]
                parts.append(lab + off)
                off += lab.numel()
            steps.append(torch.cat(parts))
        forced.append(steps)
    return forced


def _run_union(meshes, precision):
    from geobi_gnn_b200 import batching, config, dataset, network
    ref, outs = _oracle_loop(meshes)
    patches = [dataset.build_dual_data(mn, mo, device=DEV) for mn, mo in meshes]
    if len(patches) > 1:
        dv, df, slices = batching.collate_dual(patches)
    else:
        (dv, df), slices = patches[0], dict(v=[(0, patches[0][0].x.size(0))], f=[(0, patches[0][1].x.size(0))])
    mine = network.DualGNN().to(DEV).eval()
    mine.load_state_dict(ref.state_dict())
    for pl, steps in zip(util.poolings(mine), _union_forced(outs)):
        pl.forced = steps
    mine.taps = {}
    config.set_precision(precision)
    try:
        with torch.no_grad():
            vp, nrm, _ = mine([dv, df])
    finally:
        config.set_precision("fp32")
    return outs, mine, vp, nrm, slices


def _check(name, outs, mine, vp, nrm, slices, tap_tol=util.TOL_FP32):
    worst = dict(vert=0.0, normal=0.0, taps={})
    for i, o in enumerate(outs):
        a, b = slices["v"][i]
        worst["vert"] = max(worst["vert"], util.rel_err(vp[a:b], o["vp"]))
        a, b = slices["f"][i]
        worst["normal"] = max(worst["normal"], float((nrm[a:b].cpu() - o["nrm"]).abs().max()))
    for g in ("v", "f"):
        for k in TAPS:
            want = torch.cat([o["taps"][g][k] for o in outs])          # patch order = union order at every level
            worst["taps"][f"{g}.{k}"] = util.rel_err(mine.taps[g][k], want)
    # integer structure: the union's unpooling maps are the patches' maps shifted by the coarse offsets
    for li, pl in enumerate(util.poolings(mine)):
        off_c, parts = 0, []
        for o in outs:
            u = o["unpool"][li]
            parts.append(u + off_c)
            off_c += int(u.max()) + 1
        assert torch.equal(pl.unpooling_indices.cpu(), torch.cat(parts)), li
    _record(name, worst)
    print(name, json.dumps(worst))
    tap_worst = max(worst["taps"].values())
    assert worst["vert"] < util.TOL_FP32, worst
    assert tap_worst < tap_tol, worst
    assert worst["normal"] < NORMAL_TOL, worst


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_union_batch_of_bench_patches_matches_the_oracle_patch_by_patch(precision):
    """8 patches x 8000 faces (the bench shape, configs[1]), per-patch noise seeds, in the benchmarked numeric mode."""
    meshes = [util.noisy_icosphere(20, seed=s) for s in range(8)]
    outs, mine, vp, nrm, slices = _run_union(meshes, precision)
    assert vp.shape[0] == 8 * 4002 and nrm.shape[0] == 8 * 8000
    _check(f"union8x8000-{precision}", outs, mine, vp, nrm, slices)


@pytest.mark.parametrize("precision,tap_tol", [("fp32", util.TOL_FP32), ("bf16x3", 3e-5)])
def test_single_mesh_200k_faces_matches_the_oracle(precision, tap_tol):
    """One graph pair of 200 000 faces (icosphere frequency 100)."""
    meshes = [util.noisy_icosphere(100, seed=3)]
    outs, mine, vp, nrm, slices = _run_union(meshes, precision)
    assert nrm.shape[0] == 200000
    _check(f"single200k-{precision}", outs, mine, vp, nrm, slices, tap_tol)


def test_configs2_full_size_modes_agree_and_runner_reproduces_the_forward():
    """BASELINE configs[2] at FULL size (1 003 520 faces, one graph pair - the shape `bench.py`'s headline is quoted on; the CPU oracle
    would need tens of GB and minutes here), through size-independent properties:
    * the tensor-core mode ('bf16x3', benchmarked) against the CUDA-core mode ('fp32') with the fp32 run's own matchings teacher-forced
      (its PoolingLayer.trace): vertices within 1e-5, every per-layer tap within 1e-4 in max norm, unit normals within 5e-5 on
      average (bars and measured values at the assertions);
    * outputs finite, normals of unit length;
    * the end-to-end path of the bench (HostBatchRunner.upload_mesh: raw points + faces, front end on the copy stream, CSR-native
      inputs) returns the bits of the direct forward on the same matchings."""
    import numpy as np
    from geobi_gnn_b200 import config, dataset, inference, network, synth, topology
    p, f = synth.icosphere(224)
    assert f.shape[0] == 1003520
    clean = topology.DeviceTriMesh(p, f, DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    amp = torch.randn(clean.n_vertices, 1, generator=g, device=DEV) * 0.2 * clean.mean_edge_length()
    pts = (clean.points + amp * clean.vertex_normals).contiguous()
    del clean
    torch.manual_seed(3)
    net = network.DualGNN().to(DEV).eval()
    pls = util.poolings(net)

    def inputs():
        return dataset.build_dual_on_device(topology.DeviceTriMesh(pts, torch.from_numpy(f).to(DEV), DEV), None, csr_native=True)

    outs = {}
    with torch.no_grad():
        config.set_precision("fp32")
        try:
            net.taps = {}
            v, n, _ = net(list(inputs()))
            forced = [[t[2].clone() for t in pl.trace] for pl in pls]
            outs["fp32"] = (v.clone(), n.clone())
            taps32 = {g: {k: net.taps[g][k].clone() for k in TAPS} for g in ("v", "f")}
            for pl, fl in zip(pls, forced):
                pl.forced = fl
            config.set_precision("bf16x3")
            net.taps = {}
            v, n, _ = net(list(inputs()))
            outs["bf16x3"] = (v.clone(), n.clone())
            tap_err = {f"{g}.{k}": util.rel_err(net.taps[g][k], taps32[g][k]) for g in ("v", "f") for k in TAPS}
            del taps32
            net.taps = None
            # the bench's end-to-end path on the same matchings
            runner = inference.HostBatchRunner(net, torch.device(DEV), coalesced_undirected=True)
            hp = pts.cpu().pin_memory()
            hf = torch.from_numpy(f.astype(np.int32)).pin_memory()
            # pipelined as the bench does: the next upload is queued by the helper thread while this thread queues a forward
            h1 = runner.upload_mesh_async(hp, hf)
            h2 = runner.upload_mesh_async(hp, hf)
            rv, rn = runner.run(h1)
            runner.wait()
            rv, rn = rv.clone(), rn.clone()
            rv2, rn2 = runner.run(h2)
            runner.wait()
            assert torch.equal(rv2, rv) and torch.equal(rn2, rn)
        finally:
            config.set_precision("fp32")
            for pl in pls:
                pl.forced = None
    v32, n32 = outs["fp32"]
    v16, n16 = outs["bf16x3"]
    assert torch.isfinite(v16).all() and torch.isfinite(n16).all()
    assert float((n16.norm(dim=1) - 1).abs().max()) < 1e-5
    ev = util.rel_err(v16, v32)
    dn = (n16 - n32).abs().amax(1)
    en_max, en_q, en_mean = float(dn.max()), float(torch.quantile(dn, 0.9999)), float(dn.mean())
    _record("configs2_full_size", {"faces": int(f.shape[0]), "bf16x3_vs_fp32_vertices": ev, "normals_abs_max": en_max,
                                   "normals_abs_q9999": en_q, "normals_abs_mean": en_mean, "faces_above_5e-5": int((dn > 5e-5).sum()),
                                   "taps": tap_err})
    print("configs2_full_size", json.dumps(tap_err), en_max, en_q, en_mean)
    # Measured (B200): per-layer taps 3e-6 .. 7.3e-5 in max norm (the 16-bit-mantissa operand split of 'bf16x3'; the maximum runs over up
    # to 128 M entries per tap - 2.1e-5 at 200k faces), vertices 1.7e-7.  The unit normals differ by 1.7e-5 on average and 1.2e-3 at
    # worst: normalize() is ill-conditioned where the head's output vector is short, and among a million faces of a random-init network
    # some are 100x shorter than typical.  Bars: taps 1e-4 (20x inside the 2e-3 allowance north_star gives a bf16 GEMM; 'fp32' mode is
    # the one held to 1e-5, tests above), vertices 1e-5, normals 5e-5 on average and 1e-3 at the 99.99th percentile.
    assert ev < 1e-5, ev
    assert max(tap_err.values()) < 1e-4, tap_err
    assert en_mean < 5e-5 and en_q < 1e-3, (en_max, en_q, en_mean)
    assert torch.equal(rv, v16.cpu()) and torch.equal(rn, n16.cpu())
