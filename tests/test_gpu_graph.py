"""GPU: integer / graph kernels through the C ABI, bit-exact against the CPU oracle."""
import numpy as np
import pytest
import torch

from tests import util
from tests.util import pyg
from oracle import ref_data_util, ref_net_util

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from geobi_gnn_b200 import ops
    return ops


@pytest.mark.parametrize("n", [0, 1, 5, 2048, 2049, 100000, 3000001])
def test_exclusive_scan(n):
    ops = _ops()
    v = torch.randint(0, 7, (n,), dtype=torch.int32)
    out = ops.exclusive_scan(v.to(DEV)).cpu()
    want = torch.zeros(n + 1, dtype=torch.int64)
    want[1:] = torch.cumsum(v.long(), 0)
    assert torch.equal(out.long(), want)


def _random_coo(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g), torch.rand(e, generator=g)


@pytest.mark.parametrize("n,e", [(1, 0), (7, 40), (300, 5000), (20000, 300000)])
def test_csr_from_coo_matches_coalesce(n, e):
    ops = _ops()
    ei, w = _random_coo(n, e, 1)
    # coalesce(op=mean) after remove_self_loops  == pool_edge's tail
    g = ops.csr_from_coo(ei.to(DEV), n, w.to(DEV), ops.COO_DROP_SELF | ops.COO_SORT_NBR | ops.COO_DEDUP | ops.COO_W_MEAN)
    e2, w2 = pyg.remove_self_loops(ei, w)
    if e2.numel():
        e2, w2 = pyg.coalesce(e2, w2, n, n, op="mean")
    assert torch.equal(g.edge_index().cpu(), e2)
    assert g.nnz == e2.shape[1]
    if g.nnz:
        assert util.rel_err(g.w, w2) < 1e-6
    # add
    g = ops.csr_from_coo(ei.to(DEV), n, w.to(DEV), ops.COO_SORT_NBR | ops.COO_DEDUP)
    e3, w3 = pyg.coalesce(ei, w, n, n) if e else (ei, w)
    assert torch.equal(g.edge_index().cpu(), e3)
    if e:
        assert util.rel_err(g.w, w3) < 1e-6
    # stable by-source order (what the matcher sees)
    g, eid = ops.csr_from_coo(ei.to(DEV), n, w.to(DEV), ops.COO_DROP_SELF, want_eid=True)
    rowptr, col, ww = pyg.graclus_csr(ei, w, n)
    assert torch.equal(g.rowptr.cpu().long(), rowptr) and torch.equal(g.nbr.cpu().long(), col)
    if g.nnz:
        assert torch.equal(g.w.cpu(), ww)
        assert torch.equal(w[eid.cpu()], ww)
    # by target
    g = ops.csr_from_coo(ei.to(DEV), n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
    rowptr, col, _ = pyg.graclus_csr(ei.flip(0), None, n)
    assert torch.equal(g.rowptr.cpu().long(), rowptr)
    if g.nnz:
        srt = torch.cat([c.sort()[0] for c in torch.split(col, torch.diff(rowptr).tolist())])
        assert torch.equal(g.nbr.cpu().long(), srt)


def test_csr_from_coo_rejects_out_of_range():
    ops = _ops()
    from geobi_gnn_b200._lib import GeobiError
    ei = torch.tensor([[0, 5], [1, 2]])
    with pytest.raises(GeobiError):
        ops.csr_from_coo(ei.to(DEV), 3)


@pytest.mark.parametrize("n", [1, 3, 16])
def test_mesh_graphs_bit_exact(n):
    ops = _ops()
    from geobi_gnn_b200 import data_util
    mesh, _ = util.noisy_icosphere(n)
    ev, fv, vf, vv = (torch.from_numpy(a) for a in (mesh.ev, mesh.fv, mesh.vf, mesh.vv))
    # vertex graph: to_undirected + add_self_loops  (dataset.py:211-213)
    want, _ = pyg.add_self_loops(pyg.to_undirected(ev.t()))
    got = data_util.to_undirected_with_self_loops(ev.t().contiguous().to(DEV), mesh.n_vertices)
    assert torch.equal(got.cpu(), want)
    # facet graph  (data_util.py:436-456)
    want = ref_data_util.build_facet_graph(fv, vf)
    got = data_util.build_facet_graph(fv.to(DEV), vf.to(DEV))
    assert torch.equal(got.cpu(), want)
    # 2-ring vertex graph, incidence lists
    assert torch.equal(data_util.build_vertex_graph(ev.to(DEV), vv.to(DEV)).cpu(), ref_data_util.build_vertex_graph(ev, vv))
    assert torch.equal(data_util.build_edge_vf(vf.to(DEV)).cpu(), ref_data_util.build_edge_vf(vf))
    assert torch.equal(data_util.build_edge_fv(fv.to(DEV)).cpu(), ref_data_util.build_edge_fv(fv))


@pytest.mark.parametrize("n,weighted", [(2, True), (8, True), (8, False), (40, True)])
def test_graclus_equals_serial_greedy(n, weighted):
    ops = _ops()
    (dv, df), _, _ = util.oracle_inputs(n)
    for d, seed in ((dv, 3), (df, 4)):
        N = d.x.shape[0]
        perm = torch.randperm(N, generator=torch.Generator().manual_seed(seed))
        w = d.edge_weight if weighted else None
        want = pyg.graclus(d.edge_index, w, N, perm=perm)
        ei, ww = pyg.remove_self_loops(d.edge_index, w)
        g = ops.csr_from_coo(ei.to(DEV), N, None if ww is None else ww.to(DEV), 0)
        got, rounds = ops.graclus(g, perm.to(DEV), use_weight=weighted)
        assert torch.equal(got.cpu().long(), want), f"rounds={rounds}"
        cluster, nc = ops.relabel_clusters(got)
        want_c, _ = pyg.consecutive_cluster(want)
        assert torch.equal(cluster.cpu().long(), want_c) and nc == int(want_c.max()) + 1


@pytest.mark.parametrize("max_blocks", [1, 2, 5])
def test_graclus_many_nodes_per_thread_equals_serial_greedy(max_blocks):
    """The matcher keeps two nodes per thread in registers and any further ones in memory; production graphs only reach the
    second path beyond ~600 k nodes per GPU, so the grid is capped here (GEOBI_GRACLUS_MAX_BLOCKS) to run both on a small
    graph: 16 000 facets on 256 / 512 / 1280 threads = up to 63 nodes per thread.  Also a random graph with hubs."""
    import os
    ops = _ops()
    (dv, df), _, _ = util.oracle_inputs(20)
    cases = [(df.edge_index, df.edge_weight, df.x.shape[0])]
    torch.manual_seed(17)
    n = 5000
    src, dst = torch.randint(0, n, (30000,)), torch.randint(0, n, (30000,))
    hub = torch.zeros(600, dtype=torch.long)
    ei = torch.stack([torch.cat([src, hub]), torch.cat([dst, torch.randint(1, n, (600,))])])
    ei = pyg.to_undirected(ei[:, ei[0] != ei[1]], n)
    cases.append((ei, torch.rand(ei.size(1)), n))
    os.environ["GEOBI_GRACLUS_MAX_BLOCKS"] = str(max_blocks)
    try:
        for ei, w, N in cases:
            perm = torch.randperm(N, generator=torch.Generator().manual_seed(N))
            ei2, w2 = pyg.remove_self_loops(ei, w)
            want = pyg.graclus(ei2, w2, N, perm=perm)
            g = ops.csr_from_coo(ei2.to(DEV), N, w2.to(DEV), 0)
            got, und = ops.graclus(g, perm.to(DEV), check=True)
            assert und == 0 and torch.equal(got.cpu().long(), want)
    finally:
        del os.environ["GEOBI_GRACLUS_MAX_BLOCKS"]


def test_graclus_ties_and_isolated_nodes():
    ops = _ops()
    # path 0-1-2-3 with equal weights + isolated node 4: `>=` makes the LATER neighbour win
    ei = torch.tensor([[0, 1, 1, 2, 2, 3], [1, 0, 2, 1, 3, 2]])
    w = torch.ones(6)
    for perm in ([1, 0, 2, 3, 4], [4, 3, 2, 1, 0], [2, 4, 0, 1, 3]):
        perm = torch.tensor(perm)
        want = pyg.graclus(ei, w, 5, perm=perm)
        g = ops.csr_from_coo(ei.to(DEV), 5, w.to(DEV), 0)
        got, _ = ops.graclus(g, perm.to(DEV))
        assert got.cpu().tolist() == want.tolist()


@pytest.mark.parametrize("n", [2, 12])
def test_pool_step_matches_oracle(n):
    """One full pooling step: relabel -> group -> max/mean pool -> pool_edge (net_util.py:126-137)."""
    ops = _ops()
    (dv, df), _, _ = util.oracle_inputs(n)
    for d in (dv, df):
        N = d.x.shape[0]
        perm = torch.randperm(N, generator=torch.Generator().manual_seed(5))
        ei, w = pyg.remove_self_loops(d.edge_index, d.edge_weight)
        raw = pyg.graclus(ei, w, N, perm=perm)
        cl, _ = pyg.consecutive_cluster(raw)
        x = torch.randn(N, 32, generator=torch.Generator().manual_seed(6))
        want_ei, want_w = ref_net_util.pool_edge(cl, ei, w)
        g = ops.csr_from_coo(ei.to(DEV), N, w.to(DEV), 0)
        cluster, nc = ops.relabel_clusters(raw.to(DEV).int())
        mrowptr, members = ops.group_by(cluster, nc)
        assert torch.equal(cluster.cpu().long(), cl)
        for name, op in (("max", ops.OP_MAX), ("mean", ops.OP_MEAN)):
            got = ops.segment_reduce(x.to(DEV), mrowptr, members, nc, op)
            want = pyg.scatter(x, cl, dim=0, reduce=name)
            assert util.rel_err(got, want) < (1e-7 if name == "max" else 1e-6)
        g2 = ops.pool_edges(g, cluster, mrowptr, members, nc)
        assert torch.equal(g2.edge_index().cpu(), want_ei)
        assert util.rel_err(g2.w, want_w) < 1e-6
        # second step on the coarse graph, arbitrary (non-matching) labels: clusters of any size
        lab = torch.randint(0, nc, (nc,), generator=torch.Generator().manual_seed(9))
        cl2, _ = pyg.consecutive_cluster(lab)
        want_ei2, want_w2 = ref_net_util.pool_edge(cl2, want_ei, want_w)
        c2, nc2 = ops.relabel_clusters(lab.to(DEV).int())
        m2, mem2 = ops.group_by(c2, nc2)
        g3 = ops.pool_edges(g2, c2, m2, mem2, nc2)
        assert torch.equal(c2.cpu().long(), cl2)
        assert torch.equal(g3.edge_index().cpu(), want_ei2)
        if want_ei2.numel():
            assert util.rel_err(g3.w, want_w2) < 1e-5
        xm = ops.segment_reduce(x.to(DEV)[:nc], m2, mem2, nc2, ops.OP_MAX)
        assert util.rel_err(xm, pyg.scatter(x[:nc], cl2, dim=0, reduce="max")) < 1e-7


@pytest.mark.gpu
def test_csr_from_sorted_coo_matches_general_builder_and_rejects_broken_promises():
    """nn.input_graph's sort-free CSR (dataset lists are coalesced + undirected) == the general builder's, bit for bit;
    a list that is not what it claims poisons the count and the first nnz read raises."""
    import geobi_gnn_b200 as pkg
    from geobi_gnn_b200 import ops, _lib
    (dv, df), _, _ = util.oracle_inputs(5)
    for d in (dv, df):
        ei = d.edge_index.to(DEV)
        n = d.x.shape[0]
        w = torch.rand(ei.size(1), device=DEV)
        want = ops.csr_from_coo(ei, n, w, ops.COO_DROP_SELF)                       # stable by source = matcher CSR
        conv = ops.csr_from_coo(ei, n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
        g, ei2, w2 = ops.csr_from_sorted_coo(ei, n, w)
        assert g.nnz == want.nnz
        assert torch.equal(g.rowptr, want.rowptr) and torch.equal(g.nbr, want.nbr) and torch.equal(g.w, want.w)
        assert torch.equal(g.rowptr, conv.rowptr) and torch.equal(g.nbr, conv.nbr)
        keep = ei[0] != ei[1]
        assert torch.equal(ei2[:, :g.nnz], ei[:, keep]) and torch.equal(w2[:g.nnz], w[keep])
    ei = dv.edge_index.to(DEV)
    n = dv.x.shape[0]
    broken = [ei.flip(1),                                                        # unsorted
              torch.cat([ei[:, :5], ei[:, 4:]], 1),                              # duplicate
              ei[:, (ei[0] != ei[0, 0]) | (ei[1] == ei[0])]]                     # one node's out-edges missing: asymmetric
    for bad in broken:
        g, _, _ = ops.csr_from_sorted_coo(bad.contiguous(), n, None)
        with pytest.raises(_lib.GeobiError):
            g.nnz
    g, _, _ = ops.csr_from_sorted_coo(broken[2].contiguous(), n, None, check_symmetric=False)
    assert g.nnz > 0
    e0 = torch.empty((2, 0), dtype=torch.int64, device=DEV)
    g, _, _ = ops.csr_from_sorted_coo(e0, 7, None)
    assert g.nnz == 0 and torch.equal(g.rowptr, torch.zeros(8, dtype=torch.int32, device=DEV))


@pytest.mark.gpu
def test_pool_edges_fused_rows_equal_generic_pipeline_including_long_rows():
    """pool_rows_kernel (shared-memory rows, global scratch beyond 64 raw entries) == fill + sort_rows + compact_rows."""
    import os
    from geobi_gnn_b200 import ops
    torch.manual_seed(11)
    n = 3000
    # random graph + two hubs (rows of ~400 entries) so both the shared-memory and the scratch path run
    src = torch.randint(0, n, (9000,))
    dst = torch.randint(0, n, (9000,))
    hub = torch.cat([torch.zeros(400, dtype=torch.long), torch.ones(350, dtype=torch.long)])
    spokes = torch.randint(2, n, (750,))
    ei = torch.stack([torch.cat([src, hub]), torch.cat([dst, spokes])]).to(DEV)
    w = torch.rand(ei.size(1), device=DEV)
    g = ops.csr_from_coo(ei, n, w, ops.COO_SYMMETRIZE | ops.COO_SORT_NBR | ops.COO_DEDUP | ops.COO_DROP_SELF | ops.COO_W_MEAN)
    assert int((g.rowptr[1:] - g.rowptr[:-1]).max()) > 64
    label, _ = ops.graclus(g, torch.randperm(n).to(DEV))
    cluster, nc = ops.relabel_clusters(label)
    mrowptr, members = ops.group_pairs(label, cluster, nc)
    fused = ops.pool_edges(g, cluster, mrowptr, members, nc)
    os.environ["GEOBI_POOL_GENERIC"] = "1"
    try:
        generic = ops.pool_edges(g, cluster, mrowptr, members, nc)
    finally:
        del os.environ["GEOBI_POOL_GENERIC"]
    assert fused.nnz == generic.nnz > 0
    assert torch.equal(fused.rowptr, generic.rowptr) and torch.equal(fused.nbr, generic.nbr) and torch.equal(fused.w, generic.w)
    # no weights
    f2 = ops.pool_edges(g.with_weight(None), cluster, mrowptr, members, nc)
    assert f2.nnz == generic.nnz and torch.equal(f2.nbr, generic.nbr) and f2.w is None


@pytest.mark.gpu
def test_facet_graph_from_sorted_incidence_equals_the_general_builder():
    """geobi_build_facet_graph_sorted (three-way merge of ascending vf rows) gives the CSR of the fill + sort builder bit for bit, on a
    closed mesh, an open one (ragged valences) and a mesh with an isolated vertex; an unsorted row is rejected."""
    import numpy as np
    from geobi_gnn_b200 import _lib, ops, synth, topology
    cases = []
    p, f = synth.icosphere(9)
    cases.append((p, f))
    cases.append((p, f[: f.shape[0] // 3]))                      # open mesh: boundary vertices of valence 1..5
    cases.append((np.concatenate([p, [[2.0, 0, 0]]]), f))         # one vertex without faces (an all -1 row)
    for pts, fcs in cases:
        m = topology.DeviceTriMesh(pts, fcs, "cuda")
        a = ops.build_facet_graph_csr(m.fv, m.vf, vf_sorted=True)
        b = ops.build_facet_graph_csr(m.fv, m.vf, vf_sorted=False)
        assert a.nnz == b.nnz and torch.equal(a.rowptr, b.rowptr) and torch.equal(a.nbr, b.nbr)
    m = topology.DeviceTriMesh(p, f, "cuda")
    vf_bad = m.vf.clone()
    vf_bad[5, :2] = vf_bad[5, :2].flip(0)
    with pytest.raises(_lib.GeobiError):
        ops.build_facet_graph_csr(m.fv, vf_bad, vf_sorted=True)


@pytest.mark.gpu
def test_csr_native_front_end_pieces_equal_the_list_based_ones():
    """The loop-free facet CSR of the merge builder (drop_self) = the general builder's list minus its self entries; the bilateral weights
    computed in CSR order (geobi_calc_weight_csr, mean edge length counting the reference's self loops) = calc_weight on the
    reference's lists, entry for entry (closed mesh, open mesh)."""
    import numpy as np
    from geobi_gnn_b200 import data_util, ops, synth, topology
    p, f = synth.icosphere(9)
    rng = np.random.default_rng(3)
    for pts, fcs in ((p + 0.01 * rng.standard_normal(p.shape), f), (p, f[: f.shape[0] // 3])):
        m = topology.DeviceTriMesh(pts, fcs, "cuda")
        # facet graph
        g = ops.build_facet_graph_csr(m.fv, m.vf, vf_sorted=True, drop_self=True)
        ei = data_util.build_facet_graph(m.fv, m.vf)                      # sorted, self entries included
        keep = ei[0] != ei[1]
        assert g.nnz == int(keep.sum()) and torch.equal(g.edge_index(), ei[:, keep])
        pos_f = m.points[m.fv].mean(1)
        w_list = data_util.calc_weight(pos_f, m.face_normals, ei)
        w_csr = ops.calc_weight_csr(pos_f, m.face_normals, g, m.n_faces)[: g.nnz]
        assert torch.equal(w_csr, w_list[keep])
        # vertex graph: the reference appends one self loop per vertex
        gv = m.vertex_csr
        eiv = data_util.with_self_loops_appended(gv)
        w_list = data_util.calc_weight(m.points, m.vertex_normals, eiv)
        w_csr = ops.calc_weight_csr(m.points, m.vertex_normals, gv, m.n_vertices)[: gv.nnz]
        assert torch.equal(w_csr, w_list[: gv.nnz])


@pytest.mark.gpu
def test_vertex_ring_csr_equals_the_general_builder():
    """geobi_mesh_vertex_csr (other corners of each vertex's faces, sorted + deduplicated per row) = csr_from_coo over the symmetrised half
    edges, bit for bit: closed mesh, open mesh (boundary valences), an isolated vertex, a degenerate face; a fan of valence 30 (above the
    kernel's limit) is rejected by the kernel and routed through the general builder by DeviceTriMesh."""
    import numpy as np
    from geobi_gnn_b200 import _lib, ops, synth, topology
    p, f = synth.icosphere(7)
    degenerate = np.concatenate([f, [[3, 3, 9]]])
    cases = [(p, f), (p, f[: f.shape[0] // 3]), (np.concatenate([p, [[2.0, 0, 0]]]), f), (p, degenerate)]
    for pts, fcs in cases:
        fv = torch.as_tensor(np.asarray(fcs), dtype=torch.int64, device="cuda")
        V = len(pts)
        h = torch.stack([fv.reshape(-1), fv[:, [1, 2, 0]].reshape(-1)])
        want = ops.csr_from_coo(h, V, None, ops.COO_SYMMETRIZE | ops.COO_SORT_NBR | ops.COO_DEDUP | ops.COO_DROP_SELF)
        mrowptr, corners = ops.group_by(fv.reshape(-1).to(torch.int32), V)
        got = ops.mesh_vertex_csr(fv, mrowptr, corners, V)
        assert got.nnz == want.nnz and torch.equal(got.rowptr, want.rowptr) and torch.equal(got.nbr, want.nbr)
    # a fan: vertex 0 with 30 incident faces
    n = 30
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    pts = np.concatenate([[[0.0, 0, 0]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)])
    fcs = np.array([[0, 1 + i, 1 + (i + 1) % n] for i in range(n)])
    fv = torch.as_tensor(fcs, dtype=torch.int64, device="cuda")
    mrowptr, corners = ops.group_by(fv.reshape(-1).to(torch.int32), n + 1)
    with pytest.raises(_lib.GeobiError):
        ops.mesh_vertex_csr(fv, mrowptr, corners, n + 1)
    m = topology.DeviceTriMesh(pts, fcs, "cuda")
    assert int((m.vertex_csr.rowptr[1:] - m.vertex_csr.rowptr[:-1]).max()) == n and m.vertex_csr.nnz == 4 * n
