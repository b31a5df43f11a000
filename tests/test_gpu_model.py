"""GPU: the whole dual-domain forward through the reference-facing modules, against the CPU oracle and
the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from tests import util
from tests.util import pyg, ref_network

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_pair(n, seed=0, force_depth=False, perm_seed=77):
    """Oracle forward (recording) + product forward teacher-forced with the oracle's matchings."""
    from geobi_gnn_b200 import network
    data_type = "Kinect_v1" if force_depth else "Synthetic"
    (dv, df), mesh_n, mesh_o = util.oracle_inputs(n, seed=seed, data_type=data_type)
    ref = util.oracle_net(seed, force_depth=force_depth)
    util.set_perm_fn(ref, perm_seed)
    ref.record = True
    mine = network.DualGNN(force_depth=force_depth).to(DEV)
    mine.load_state_dict(ref.state_dict())
    mine.eval()
    dv_m, df_m = util.data_to(dv, DEV), util.data_to(df, DEV)
    with torch.no_grad():
        want = ref([dv, df])
    for a, b in zip(util.poolings(mine), util.poolings(ref)):
        a.forced = [t[3] for t in b.trace]
    mine.taps = {}
    with torch.no_grad():
        got = mine([dv_m, df_m])
    return ref, mine, want, got, (dv, df), (dv_m, df_m)


@pytest.mark.parametrize("n,force_depth", [(3, False), (12, False), (6, True), (32, False)])
def test_dualgnn_forward_matches_oracle_teacher_forced(n, force_depth):
    ref, mine, want, got, d_ref, d_mine = _run_pair(n, force_depth=force_depth)
    assert got[2] is None and got[0].shape == want[0].shape and got[1].shape == want[1].shape
    # integer structure is bit exact at every level
    for a, b in zip(util.poolings(mine), util.poolings(ref)):
        assert torch.equal(a.unpooling_indices.cpu(), b.unpooling_indices)
    # per-layer taps: each compared with the oracle's tensor of the same name
    worst = 0.0
    for gname in ("v", "f"):
        for k in ("l1", "p1", "l2", "p2", "l3", "l4", "r1", "r2", "r3", "r4"):
            e = util.rel_err(mine.taps[gname][k], ref.taps[gname][k])
            worst = max(worst, e)
            assert e < util.TOL_FP32, (gname, k, e)
    assert util.rel_err(mine.taps["xf12"], ref.taps["xf12"]) < util.TOL_FP32
    assert util.rel_err(got[0], want[0]) < util.TOL_FP32
    assert util.rel_err(got[1], want[1]) < util.TOL_NORMAL   # unit normals: see tests/util.py
    # inputs are mutated as upstream: 64-channel x on graph_v, 12->64 on graph_f, self loops stripped
    assert d_mine[0].x.shape[1] == 64 and d_mine[1].x.shape[1] == 64
    assert torch.equal(d_mine[0].edge_index.cpu(), d_ref[0].edge_index)
    assert torch.equal(d_mine[1].edge_index.cpu(), d_ref[1].edge_index)
    # downstream metric agrees
    e_ref = ref_network.error_n(want[1], d_ref[1].y)
    from geobi_gnn_b200 import network
    e_mine = network.error_n(got[1], d_mine[1].y)
    assert abs(float(e_ref) - float(e_mine)) < 1e-2


def test_dualgnn_matches_golden_vectors():
    from geobi_gnn_b200 import network
    from geobi_gnn_b200.data import Data
    g = np.load(os.path.join(util.GOLDEN, "dualgnn_ico3.npz"))
    t = lambda k: torch.from_numpy(g[k]).to(DEV)
    dv = Data(x=t("x_v"), edge_index=t("ei_v"), edge_weight=t("w_v"))
    df = Data(x=t("x_f"), edge_index=t("ei_f"), edge_weight=t("w_f"), fv_indices=t("faces"))
    mine = network.DualGNN().to(DEV)
    mine.load_state_dict(util.oracle_net(0).state_dict())
    for name, pl in zip(("v1", "v2", "f1", "f2"), util.poolings(mine)):
        perms = [torch.from_numpy(g[f"perm_{name}_{s}"]) for s in range(2)]
        pl.perm_fn = lambda n, it=iter(perms): next(it)
    with torch.no_grad():
        vp, nrm, _ = mine([dv, df])
    # free-running (same visiting order): matchings should coincide with the oracle's on this small mesh
    for name, pl in zip(("v1", "v2", "f1", "f2"), util.poolings(mine)):
        for s, tr in enumerate(pl.trace):
            assert np.array_equal(tr[2].cpu().numpy(), g[f"label_{name}_{s}"]), (name, s)
        assert np.array_equal(pl.unpooling_indices.cpu().numpy(), g[f"unpool_{name}"])
    assert util.rel_err(vp, g["vert_p"]) < util.TOL_FP32
    assert util.rel_err(nrm, g["norm_p"]) < util.TOL_NORMAL


def test_dataset_builder_matches_oracle():
    from geobi_gnn_b200 import dataset
    mesh_n, mesh_o = util.noisy_icosphere(9, seed=2)
    (dv, df), _, _ = util.oracle_inputs(9, seed=2)
    mv, mf = dataset.build_dual_data(mesh_n, mesh_o, device=DEV)
    assert torch.equal(mv.edge_index.cpu(), dv.edge_index) and torch.equal(mf.edge_index.cpu(), df.edge_index)
    assert torch.equal(mf.fv_indices.cpu(), df.fv_indices)
    for a, b in ((mv.x, dv.x), (mf.x, df.x), (mv.edge_weight, dv.edge_weight), (mf.edge_weight, df.edge_weight), (mv.y, dv.y), (mf.y, df.y)):
        assert util.rel_err(a, b) < util.TOL_FP32
    assert "pos" not in mv and "normal" not in mv and "depth_direction" not in mv and "edge_dual" not in mf


def test_free_running_forward_invariants():
    """Default path (random visiting order on the device): outputs are finite, normals unit, matching valid."""
    from geobi_gnn_b200 import dataset, network
    mesh_n, mesh_o = util.noisy_icosphere(16)
    dv, df = dataset.build_dual_data(mesh_n, mesh_o, device=DEV)
    torch.manual_seed(0)
    net = network.DualGNN().to(DEV).eval()
    with torch.no_grad():
        vp, nrm, _ = net([dv, df])
    assert torch.isfinite(vp).all() and torch.isfinite(nrm).all()
    assert util.rel_err(nrm.norm(dim=1), torch.ones(nrm.shape[0])) < 1e-5
    for pl in util.poolings(net):
        for g, perm, label in pl.trace:
            lab = label.cpu().long()
            n = lab.numel()
            assert torch.bincount(lab, minlength=n).max() <= 2 and torch.all(lab <= torch.arange(n))
            single = torch.bincount(lab, minlength=n)[lab] == 1
            ei = g.edge_index().cpu()
            assert not (single[ei[0]] & single[ei[1]]).any()


@pytest.mark.gpu
def test_dualgnn_same_bits_with_and_without_the_coalesced_hint():
    """The sort-free input-graph path (Data.coalesced_undirected) must not change a single bit of the forward."""
    from geobi_gnn_b200 import batching, dataset, network, synth
    torch.manual_seed(3)
    net = network.DualGNN().to(DEV).eval()
    patches = []
    for seed in range(2):
        mn, mo = util.noisy_icosphere(3, seed=seed)
        patches.append(dataset.build_dual_data(mn, mo, device=DEV))
    dv, df, _ = batching.collate_dual(patches)
    assert dv.coalesced_undirected and df.coalesced_undirected

    def run(hint):
        a, b = batching.fresh_view(dv), batching.fresh_view(df)
        if not hint:
            a.coalesced_undirected = None
            b.coalesced_undirected = None
        for pl in util.poolings(net):
            pl.perm_fn = lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(n))
        with torch.no_grad():
            out = net([a, b])
        return [o.clone() for o in out if torch.is_tensor(o)], a, b

    o1, a1, b1 = run(True)
    o0, a0, b0 = run(False)
    assert len(o1) == len(o0) and len(o1) > 0
    for x, y in zip(o1, o0):
        assert torch.equal(x, y)
    assert torch.equal(a1.edge_index, a0.edge_index) and torch.equal(b1.edge_index, b0.edge_index)   # stripped write-back
    assert torch.equal(a1.edge_weight, a0.edge_weight)


@pytest.mark.gpu
@pytest.mark.parametrize("packed", [False, True])
def test_host_batch_runner_equals_direct_forward(packed):
    """inference.HostBatchRunner (pinned host batches, upload on a copy stream) returns the bits of a plain forward, from the
    reference's int64 host layout and from the int32-narrowed one (HostBatchRunner.pack)."""
    from geobi_gnn_b200 import batching, dataset, inference, network
    torch.manual_seed(5)
    net = network.DualGNN().to(DEV).eval()
    batches = []
    for seed in range(3):
        mn, mo = util.noisy_icosphere(3, seed=10 + seed)
        dv, df = dataset.build_dual_data(mn, mo, device=DEV)
        batches.append((dv, df))
    for pl in util.poolings(net):
        pl.perm_fn = lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(n))
    want = []
    with torch.no_grad():
        for dv, df in batches:
            vp, nrm, _ = net([batching.fresh_view(dv), batching.fresh_view(df)])
            want.append((vp.cpu(), nrm.cpu()))
    host = [({k: getattr(dv, k).cpu().pin_memory() for k in ("x", "edge_index", "edge_weight")},
             {k: getattr(df, k).cpu().pin_memory() for k in ("x", "edge_index", "edge_weight", "fv_indices")}) for dv, df in batches]
    if packed:
        host = [(inference.HostBatchRunner.pack(hv), inference.HostBatchRunner.pack(hf)) for hv, hf in host]
        assert host[0][0]["edge_index"].dtype == torch.int32 and host[0][1]["fv_indices"].dtype == torch.int32
    runner = inference.HostBatchRunner(net, DEV, coalesced_undirected=True)
    nxt = runner.upload(*host[0])
    for i in range(3):
        cur = nxt
        if i + 1 < 3:
            nxt = runner.upload(*host[i + 1])
        v, n = runner.run(cur)
        runner.wait()
        assert torch.equal(v, want[i][0]) and torch.equal(n, want[i][1])


@pytest.mark.gpu
def test_forward_rejects_a_false_coalesced_undirected_claim():
    """A Data that claims coalesced_undirected but is not must not produce numbers: the device-side verdict surfaces as a
    GeobiError behind the first pooling step's count read-back."""
    from geobi_gnn_b200 import _lib, batching, dataset, network
    torch.manual_seed(6)
    net = network.DualGNN().to(DEV).eval()
    mn, mo = util.noisy_icosphere(3, seed=1)
    dv, df = dataset.build_dual_data(mn, mo, device=DEV)
    a, b = batching.fresh_view(dv), batching.fresh_view(df)
    perm = torch.randperm(a.edge_index.size(1), device=DEV)
    shuffled_ei = a.edge_index[:, perm].contiguous()           # same edge set, no longer sorted
    shuffled_w = a.edge_weight[perm].contiguous()
    a.edge_index, a.edge_weight = shuffled_ei, shuffled_w
    a.coalesced_undirected = True
    with pytest.raises(_lib.GeobiError):
        with torch.no_grad():
            net([a, b])
    # without the claim the same (unsorted) list is fine
    a2, b2 = batching.fresh_view(dv), batching.fresh_view(df)
    a2.edge_index, a2.edge_weight = shuffled_ei, shuffled_w
    a2.coalesced_undirected = None
    with torch.no_grad():
        out = net([a2, b2])
    assert torch.isfinite(out[0]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("helper_thread", [False, True])
def test_host_batch_runner_upload_mesh_equals_direct_forward(helper_thread):
    """upload_mesh (raw points + faces cross PCIe; topology, graphs, weights and features built on the copy stream) gives the bits of
    a forward on inputs built by the same device front end on the main stream, and the oracle's features within the fp32 bar."""
    from geobi_gnn_b200 import batching, dataset, inference, network, topology
    torch.manual_seed(7)
    net = network.DualGNN().to(DEV).eval()
    for pl in util.poolings(net):
        pl.perm_fn = lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(n))
    meshes = [util.noisy_icosphere(6, seed=20 + s)[0] for s in range(3)]
    want = []
    with torch.no_grad():
        for m in meshes:
            dm = topology.DeviceTriMesh(m.points, m.fv, DEV)
            dv, df = dataset.build_dual_on_device(dm, None)
            (ov, of), _, _ = util.oracle_inputs(6, seed=20 + len(want))
            assert util.rel_err(dv.x, ov.x) < util.TOL_FP32 and util.rel_err(df.x, of.x) < util.TOL_FP32
            assert torch.equal(df.edge_index.cpu(), of.edge_index) and util.rel_err(df.edge_weight, of.edge_weight) < util.TOL_FP32
            vp, nrm, _ = net([batching.fresh_view(dv), batching.fresh_view(df)])
            want.append((vp.cpu(), nrm.cpu()))
    host = [(torch.from_numpy(m.points.astype("float32")).pin_memory(), torch.from_numpy(m.fv.astype("int32")).pin_memory()) for m in meshes]
    runner = inference.HostBatchRunner(net, DEV, coalesced_undirected=True)
    up = runner.upload_mesh_async if helper_thread else runner.upload_mesh      # the front end queued by a helper thread: same bits
    nxt = up(*host[0])
    for i in range(3):
        cur = nxt
        if i + 1 < 3:
            nxt = up(*host[i + 1])
        v, n = runner.run(cur)
        runner.wait()
        assert torch.equal(v, want[i][0]) and torch.equal(n, want[i][1])


@pytest.mark.gpu
def test_csr_native_inputs_equal_list_inputs():
    """dataset.build_dual_on_device(csr_native=True): same forward bits as the list-based inputs, and the lazy edge_index / edge_weight
    read as the reference's lists when somebody asks for them (after the forward, which must not have touched them)."""
    from geobi_gnn_b200 import batching, dataset, network, topology
    from geobi_gnn_b200.data import _Lazy
    torch.manual_seed(11)
    net = network.DualGNN().to(DEV).eval()
    for pl in util.poolings(net):
        pl.perm_fn = lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(n))
    m = util.noisy_icosphere(8, seed=5)[0]
    dm = topology.DeviceTriMesh(m.points, m.fv, DEV)
    with torch.no_grad():
        lv, lf = dataset.build_dual_on_device(dm, None)
        want = net([batching.fresh_view(lv), batching.fresh_view(lf)])
        cv, cf = dataset.build_dual_on_device(dm, None, csr_native=True)
        assert torch.equal(cv.x, lv.x) and torch.equal(cf.x, lf.x)        # before the forward: the network overwrites its inputs' x, as upstream
        got = net([cv, cf])
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    for c, l in ((cv, lv), (cf, lf)):
        assert isinstance(c._items["edge_index"], _Lazy) and isinstance(c._items["edge_weight"], _Lazy)     # the network never read them
        assert torch.equal(c.edge_index, l.edge_index) and torch.equal(c.edge_weight, l.edge_weight)


@pytest.mark.gpu
def test_concurrent_forwards_on_replicas_equal_sequential_ones():
    """bench.py runs several forwards at once (one host thread + CUDA stream + module replica each): with fixed visiting orders every
    concurrent forward returns the bits of the same forward run alone - the library's per-stream workspaces, thread-local size hints
    and pinned landing pads do not leak between threads."""
    import copy
    import threading
    from geobi_gnn_b200 import batching, dataset, network, topology
    torch.manual_seed(21)
    net = network.DualGNN().to(DEV).eval()
    meshes = [util.noisy_icosphere(n, seed=30 + n)[0] for n in (10, 12, 14)]
    inputs = []
    for m in meshes:
        dm = topology.DeviceTriMesh(m.points, m.fv, DEV)
        inputs.append(dataset.build_dual_on_device(dm, None, csr_native=True))
    nets = [net] + [copy.deepcopy(net).eval() for _ in range(2)]
    for n_ in nets:
        for pl in util.poolings(n_):
            pl.perm_fn = lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(n))
    with torch.no_grad():
        want = [[t.clone() for t in net([batching.fresh_view(dv), batching.fresh_view(df)])[:2]] for dv, df in inputs]
    torch.cuda.synchronize()
    got = [None] * 3
    errs = []

    def work(i):
        try:
            torch.cuda.set_device(torch.device(DEV).index or 0)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.default_stream())
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(4):              # several rounds each, to interleave
                    dv, df = inputs[i]
                    out = nets[i]([batching.fresh_view(dv), batching.fresh_view(df)])
                got[i] = [out[0].clone(), out[1].clone()]
            s.synchronize()
        except BaseException as e:
            errs.append(e)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for i in range(3):
        assert torch.equal(got[i][0], want[i][0]) and torch.equal(got[i][1], want[i][1]), i
