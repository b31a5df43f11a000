"""CPU oracle: the reference model and losses.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/code/network.py:254-343 (GNNModule, DualGNN) and
:347-413 (losses, metrics) in plain PyTorch fp32 on CPU, reference evaluation
order (FeaSt projection per edge, materialised [N,1024] hidden).  Pinned against the reference's own network.py executed
over oracle/pyg.py (same seeded weights bit for bit, same outputs and gradients: tests/test_reference_golden.py);
the third-party operators in oracle/pyg.py are unpinned.

``DualGNN.forward`` additionally stores every intermediate the parity tests
teacher-force with, in ``self.taps`` (a dict), when ``self.record`` is True.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ref_data_util as data_util
from .pyg import FeaStConv, remove_self_loops, scatter
from .ref_net_util import PoolingLayer


def _act(t):
    return F.leaky_relu(t, 0.2)


class GNNModule(nn.Module):
    """network.py:254-300 — FeaSt U-Net: C0→32 ▸pool▸ 64 ▸pool▸ 128→128 ▸up▸ 64⧺64→64 ▸up▸ 32⧺32→32.
    leaky_relu(0.2) after every conv except r_conv1 and r_conv3."""

    def __init__(self, in_channel=6, pool_type="max", pool_step=2, edge_weight_type=0, wei_param=2):
        super().__init__()
        self.l_conv1 = FeaStConv(in_channel, 32, 9)
        self.pooling1 = PoolingLayer(32, pool_type, pool_step, edge_weight_type, wei_param)
        self.l_conv2 = FeaStConv(32, 64, 9)
        self.pooling2 = PoolingLayer(64, pool_type, pool_step, edge_weight_type, wei_param)
        self.l_conv3 = FeaStConv(64, 128, 9)
        self.l_conv4 = FeaStConv(128, 128, 9)
        self.r_conv1 = FeaStConv(128, 64, 9)
        self.r_conv2 = FeaStConv(128, 64, 9)
        self.r_conv3 = FeaStConv(64, 32, 9)
        self.r_conv4 = FeaStConv(64, 32, 9)
        self.taps = None

    def forward(self, r1, plot_pool=False):
        tap = {} if self.taps is not None else None

        def rec(k, v):
            if tap is not None:
                tap[k] = v.detach().clone()
            return v

        r1.x = rec("l1", _act(self.l_conv1(r1.x, r1.edge_index)))
        r2 = self.pooling1(r1)
        rec("p1", r2.x)
        r2.x = rec("l2", _act(self.l_conv2(r2.x, r2.edge_index)))
        r3 = self.pooling2(r2)
        rec("p2", r3.x)
        r3.x = rec("l3", _act(self.l_conv3(r3.x, r3.edge_index)))
        r3.x = rec("l4", _act(self.l_conv4(r3.x, r3.edge_index)))
        up2 = rec("r1", self.r_conv1(self.pooling2.unpooling(r3.x), r2.edge_index))
        r2.x = torch.cat((r2.x, up2), 1)
        r2.x = rec("r2", _act(self.r_conv2(r2.x, r2.edge_index)))
        up1 = rec("r3", self.r_conv3(self.pooling1.unpooling(r2.x), r1.edge_index))
        r1.x = torch.cat((r1.x, up1), 1)
        out = rec("r4", _act(self.r_conv4(r1.x, r1.edge_index)))
        if tap is not None:
            tap["ei_r1"], tap["ei_r2"], tap["ei_r3"] = r1.edge_index, r2.edge_index, r3.edge_index
            self.taps = tap
        return out


class DualGNN(nn.Module):
    """network.py:303-343."""

    def __init__(self, force_depth=False, pool_type="max", edge_weight_type=10, wei_param=2):
        super().__init__()
        self.force_depth = force_depth
        self.gnn_v = GNNModule(6, pool_type, 2, edge_weight_type, wei_param)
        self.fc_v1 = nn.Linear(32, 1024)
        self.fc_v2 = nn.Linear(1024, 1 if force_depth else 3)
        self.gnn_f = GNNModule(12, pool_type, 2, edge_weight_type, wei_param)
        self.fc_f1 = nn.Linear(32, 1024)
        self.fc_f2 = nn.Linear(1024, 3)
        self.record = False
        self.taps = {}

    def forward(self, dual_data):
        data_v, data_f = dual_data
        xyz = data_v.x[:, :3]
        if self.record:
            self.gnn_v.taps, self.gnn_f.taps = {}, {}
        g_v = self.gnn_v(data_v)
        feat_v = self.fc_v2(_act(self.fc_v1(g_v)))
        if self.force_depth:
            feat_v = feat_v * data_v.depth_direction
        feat_v = feat_v + xyz
        cent = feat_v[data_f.fv_indices].mean(1)
        nrm = data_util.computer_face_normal(feat_v, data_f.fv_indices)
        data_f.x = torch.cat((data_f.x, cent, nrm), 1)
        xf12 = data_f.x
        g_f = self.gnn_f(data_f)
        feat_f = self.fc_f2(_act(self.fc_f1(g_f)))
        if self.record:
            self.taps = dict(g_v=g_v.detach(), feat_v=feat_v.detach(), xf12=xf12.detach().clone(),
                             g_f=g_f.detach(), feat_f=feat_f.detach(), v=self.gnn_v.taps, f=self.gnn_f.taps)
            self.gnn_v.taps = self.gnn_f.taps = None
        return feat_v, F.normalize(feat_f, dim=1), None


# ---------------------------------------------------------------- losses (network.py:347-413)
def _laplacian(v, edge_idx_v, normal=None):
    row, col = edge_idx_v
    lap = scatter(v[row] - v[col], row, dim=0, reduce="mean")
    if normal is not None:
        lap = normal * (lap * normal).sum(1, keepdim=True)
    return lap


def laplacian_loss(vp, v, edge_idx_v, normal=None):
    edge_idx_v, _ = remove_self_loops(edge_idx_v)
    return (_laplacian(vp, edge_idx_v, normal) - _laplacian(v, edge_idx_v, normal)).abs().sum(1).mean()


def loss_v(vp, v, dis="L2", apply_icp=False):
    if dis == "L1":
        return (vp - v).abs().sum(1).mean()
    if dis == "L2":
        return (vp - v).pow(2).sum(1).mean()
    raise NotImplementedError(dis)      # 'CD' / 'EMD' need kaolin upstream (commented import)


def loss_n(np_, n, norm="L1", fc_p=None, fc=None):
    if norm == "L1":
        return (np_ - n).abs().sum(1).mean()
    if norm == "L2":
        return (np_ - n).pow(2).sum(1).mean()
    raise NotImplementedError(norm)     # 'sided' needs kaolin upstream


def dual_loss(loss_v_, loss_n_, v_scale=1, n_scale=1, alpha=None):
    if alpha is None:
        return loss_v_ * v_scale + loss_n_ * n_scale
    return alpha * loss_v_ * v_scale + (1 - alpha) * loss_n_ * n_scale


def error_v(vp, v):
    return (vp - v).pow(2).sum(1).pow(0.5).mean()


def error_n(np_, n):
    val = torch.clamp(1 - (np_ - n).pow(2).sum(1) / 2, min=-1, max=1)
    return (torch.acos(val) * 180 / math.pi).mean()
