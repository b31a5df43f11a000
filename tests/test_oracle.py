"""CPU: the oracle against its committed golden vectors, its own cross-checks and the published
properties of the algorithms it restates.  (The reference has no tests or vectors — SURVEY.md 4.)"""
import os

import numpy as np
import pytest
import torch

from tests import util
from tests.util import pyg, ref_dataset, ref_network
from oracle import ref_data_util, ref_net_util


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(util.GOLDEN, "dualgnn_ico3.npz"))


def test_golden_forward_reproduces(golden):
    torch.set_num_threads(1)
    (dv, df), _, _ = util.oracle_inputs(3, seed=0)
    assert np.array_equal(dv.edge_index.numpy(), golden["ei_v"]) and np.array_equal(df.edge_index.numpy(), golden["ei_f"])
    assert util.rel_err(dv.x, golden["x_v"]) < 1e-6 and util.rel_err(df.edge_weight, golden["w_f"]) < 1e-6
    net = util.oracle_net(0)
    util.set_perm_fn(net, 1234)
    with torch.no_grad():
        vp, nrm, third = net([dv, df])
    assert third is None
    for name, pl in zip(("v1", "v2", "f1", "f2"), util.poolings(net)):
        for s, tr in enumerate(pl.trace):
            assert np.array_equal(tr[3].numpy(), golden[f"label_{name}_{s}"])
    assert util.rel_err(vp, golden["vert_p"]) < 1e-5
    assert util.rel_err(nrm, golden["norm_p"]) < 1e-5


def test_sizes_match_survey():
    # config 1 -> V=10242, F=20480, E_v=71682, E_f=266180; 939128 parameters
    (dv, df), _, _ = util.oracle_inputs(32)
    assert dv.x.shape == (10242, 6) and df.x.shape == (20480, 6)
    assert dv.edge_index.shape[1] == 71682 and df.edge_index.shape[1] == 266180
    assert sum(p.numel() for p in ref_network.DualGNN().parameters()) == 939128
    assert sum(p.numel() for p in ref_network.DualGNN(force_depth=True).parameters()) == 937078


def test_graclus_c_matches_python_and_is_a_maximal_matching():
    (dv, df), _, _ = util.oracle_inputs(4)
    for d in (dv, df):
        n = d.x.shape[0]
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(7))
        for w in (d.edge_weight, None):
            lab = pyg.graclus(d.edge_index, w, n, perm=perm)
            assert torch.equal(lab, pyg.graclus_python(d.edge_index, w, n, perm))
            cnt = torch.bincount(lab, minlength=n)
            assert cnt.max() <= 2 and torch.all(lab <= torch.arange(n))
            ei, _ = pyg.remove_self_loops(d.edge_index)
            adj = set(map(tuple, ei.t().tolist()))
            for u in torch.nonzero(lab != torch.arange(n)).flatten().tolist():
                assert (u, int(lab[u])) in adj                       # partners are adjacent
            single = cnt[lab] == 1
            assert not (single[ei[0]] & single[ei[1]]).any()         # maximal: no two adjacent singletons


def test_coalesce_and_pool_edge_semantics():
    idx = torch.tensor([[2, 0, 2, 1, 0, 2], [1, 1, 1, 0, 1, 0]])
    val = torch.tensor([1., 2., 3., 4., 6., 8.])
    i, v = pyg.coalesce(idx, val, 3, 3)
    assert i.tolist() == [[0, 1, 2, 2], [1, 0, 0, 1]] and v.tolist() == [8., 4., 8., 4.]
    i, v = pyg.coalesce(idx, val, 3, 3, op="mean")
    assert v.tolist() == [4., 4., 8., 2.]
    cluster = torch.tensor([0, 0, 1, 1, 2])
    ei = torch.tensor([[0, 1, 2, 3, 4, 1], [2, 3, 4, 0, 0, 0]])
    e2, w2 = ref_net_util.pool_edge(cluster, ei, torch.tensor([1., 3., 5., 7., 9., 11.]))
    assert e2.tolist() == [[0, 1, 1, 2], [1, 0, 2, 0]] and w2.tolist() == [2., 7., 5., 9.]


def test_feast_matches_dense_float64_formula():
    torch.manual_seed(0)
    n, cin, cout = 40, 12, 32
    conv = pyg.FeaStConv(cin, cout, 9)
    x = torch.randn(n, cin)
    ei = torch.randint(0, n, (2, 200))
    out = conv(x, ei)
    W = conv.lin.weight.double().view(9, cout, cin)
    U, c, b, xd = conv.u.weight.double(), conv.c.double(), conv.bias.double(), x.double()
    e2, _ = pyg.remove_self_loops(ei)
    ref = torch.zeros(n, cout, dtype=torch.float64)
    for i in range(n):
        js = e2[0][e2[1] == i].tolist() + [i]
        acc = torch.zeros(cout, dtype=torch.float64)
        for j in js:
            q = torch.softmax(U @ (xd[j] - xd[i]) + c, 0)
            acc += torch.einsum("h,hoc,c->o", q, W, xd[j])
        ref[i] = acc / len(js) + b
    assert util.rel_err(out, ref) < 1e-5


def test_update_position_variants_and_c_agree():
    import ctypes
    (dv, df), mesh_n, _ = util.oracle_inputs(3)
    pts = torch.from_numpy(mesh_n.points).float()
    fv, vf = torch.from_numpy(mesh_n.fv), torch.from_numpy(mesh_n.vf)
    fn = torch.from_numpy(mesh_n.face_normals).float()
    a = ref_data_util.update_position(pts, fv, vf, fn, 10)
    b = ref_data_util.update_position2(pts, fv, vf, fn, 10)
    assert util.rel_err(a, b) < 1e-5
    lib = pyg._oracle_lib()
    V, F, K = pts.shape[0], fv.shape[0], vf.shape[1]
    out, cent, tmp = torch.empty_like(pts), torch.empty(F, 3), torch.empty_like(pts)
    lib.oracle_update_position2(ctypes.c_int64(V), ctypes.c_int64(F), ctypes.c_int64(K), ctypes.c_void_p(pts.data_ptr()),
                                ctypes.c_void_p(fv.data_ptr()), ctypes.c_void_p(vf.data_ptr()), ctypes.c_void_p(fn.data_ptr()),
                                ctypes.c_int(10), None, ctypes.c_void_p(cent.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                ctypes.c_void_p(tmp.data_ptr()))
    assert util.rel_err(out, b) < 1e-5


def test_pooling_weight_modes_run():
    (dv, df), _, _ = util.oracle_inputs(2)
    for t in (-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
        torch.manual_seed(0)
        pl = ref_net_util.PoolingLayer(6, "max", 2, t, 2)
        d = dv.clone()
        out = pl(d)
        assert out.x.shape[0] < dv.x.shape[0] and pl.unpooling(out.x).shape[0] == dv.x.shape[0]


def test_stitch_patches_mean():
    r = [(torch.ones(2, 3), torch.tensor([[1., 0, 0], [0, 1., 0]]), torch.tensor([0, 1]), torch.tensor([0, 1])),
         (3 * torch.ones(2, 3), torch.tensor([[0, 1., 0], [0, 0, 1.]]), torch.tensor([1, 2]), torch.tensor([1, 2]))]
    vp, nrm = ref_dataset.stitch_patches(3, 3, r)
    assert vp.tolist() == [[1.] * 3, [2.] * 3, [3.] * 3]
    assert util.rel_err(nrm[1], torch.tensor([0., 1., 0.])) < 1e-6
