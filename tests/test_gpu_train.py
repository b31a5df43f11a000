"""GPU: training step — parameter gradients of the whole DualGNN against the CPU oracle's autograd (teacher-forced
matchings), and the per-op backward kernels."""
import pytest
import torch

from tests import util
from tests.util import pyg, ref_network

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("cin,cout,slope", [(12, 32, 0.2), (64, 32, 1.0), (128, 64, 0.2)])
def test_feast_backward_matches_autograd(cin, cout, slope):
    from geobi_gnn_b200 import ops
    from geobi_gnn_b200.autograd import FeaStFn
    (dv, df), _, _ = util.oracle_inputs(5)
    torch.manual_seed(cin + cout)
    conv = pyg.FeaStConv(cin, cout, 9)
    n = df.x.shape[0]
    x = (torch.randn(n, cin) * 2).requires_grad_()
    gout = torch.randn(n, cout)
    y = conv(x, df.edge_index)
    y = y if slope == 1.0 else torch.nn.functional.leaky_relu(y, slope)
    y.backward(gout)
    g = ops.csr_from_coo(df.edge_index.to(DEV), n, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
    xm = x.detach().to(DEV).requires_grad_()
    P = [t.detach().to(DEV).requires_grad_() for t in (conv.lin.weight, conv.u.weight, conv.c, conv.bias)]
    ym = FeaStFn.apply(xm, *P, g, slope, ops.PREC_FP32)
    ym.backward(gout.to(DEV))
    assert util.rel_err(ym, y) < util.TOL_FP32
    for got, want, name in ((xm.grad, x.grad, "x"), (P[0].grad, conv.lin.weight.grad, "W"), (P[1].grad, conv.u.weight.grad, "U"),
                            (P[2].grad, conv.c.grad, "c"), (P[3].grad, conv.bias.grad, "bias")):
        assert util.rel_err(got, want) < 2e-4, (name, util.rel_err(got, want))


def test_dualgnn_training_step_gradients_match_oracle():
    from geobi_gnn_b200 import network
    (dv, df), _, _ = util.oracle_inputs(6)
    ref = util.oracle_net(0)
    ref.train()
    util.set_perm_fn(ref, 5)
    mine = network.DualGNN().to(DEV)
    mine.load_state_dict(ref.state_dict())
    mine.train()
    dv_m, df_m = util.data_to(dv, DEV), util.data_to(df, DEV)
    vp, nrm, _ = ref([dv, df])
    loss = ref_network.dual_loss(ref_network.loss_v(vp, dv.y, "L1"), ref_network.loss_n(nrm, df.y, "L1"))
    loss.backward()
    for a, b in zip(util.poolings(mine), util.poolings(ref)):
        a.forced = [t[3] for t in b.trace]
    vp_m, nrm_m, _ = mine([dv_m, df_m])
    loss_m = network.dual_loss(network.loss_v(vp_m, dv_m.y, "L1"), network.loss_n(nrm_m, df_m.y, "L1"))
    loss_m.backward()
    assert abs(float(loss_m) - float(loss)) < 1e-4 * abs(float(loss))
    worst = 0.0
    for (name, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, name
        e = util.rel_err(p.grad, q.grad)
        worst = max(worst, e)
        assert e < 2e-3, (name, e)
    # one optimiser step keeps the two models together
    for net in (ref, mine):
        torch.optim.Adam(net.parameters(), lr=1e-3).step()
    for (name, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert util.rel_err(p.data, q.data) < 1e-3, name


def test_v2f_transfer_backward_matches_eager_autograd():
    """geobi_v2f_transfer_bwd (corner mean + normalize(cross)) against autograd of the reference's tensor expression (network.py:335-337),
    on a noisy sphere; a thin face (one corner moved to 5 % of an edge) is compared at a looser bar - its normal is ill-conditioned
    in fp32 whoever computes it."""
    import torch.nn.functional as F
    from geobi_gnn_b200.autograd import V2FTransferFn
    mesh = util.noisy_icosphere(6, seed=2)[0]
    pts = torch.from_numpy(mesh.points.astype("float32")).to(DEV)
    fv = torch.from_numpy(mesh.fv).to(DEV)
    pts[fv[5, 1]] = pts[fv[5, 0]] + 0.05 * (pts[fv[5, 2]] - pts[fv[5, 0]]) + 0.02 * pts[fv[5, 0]]      # a thin face
    xf = torch.randn(fv.size(0), 6, device=DEV)
    gout = torch.randn(fv.size(0), 12, device=DEV)
    a = pts.clone().requires_grad_(True)
    tri = a[fv]
    want_out = torch.cat((xf, tri.mean(1), F.normalize(torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=1), dim=1)), 1)
    want_out.backward(gout)
    b = pts.clone().requires_grad_(True)
    got_out = V2FTransferFn.apply(b, fv, xf)
    got_out.backward(gout)
    assert util.rel_err(got_out.detach(), want_out.detach()) < 1e-5
    assert util.rel_err(b.grad, a.grad) < 1e-4
    keep = torch.ones(pts.size(0), dtype=torch.bool, device=DEV)
    keep[fv[5]] = False                                        # vertices that do not touch the thin face: fp32 bar
    assert util.rel_err(b.grad[keep], a.grad[keep]) < 1e-5
