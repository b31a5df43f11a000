"""DualGNN / GNNModule with the reference's surface (/root/reference/code/network.py:254-413),
running on libgeobi kernels.

Constructor signatures, child-module names and state-dict keys match the reference, so
``net.load_state_dict(torch.load(model.pth))`` from test_dual.py:130 works unchanged and
``net(data) -> (vert_p [V,3], unit norm_p [F,3], None)`` is a drop-in for
train_dual.py:203 / test_dual.py:21.  As upstream, forward MUTATES its inputs (``data.x``,
``data.edge_index`` stripped of self loops, ``data_f.x`` widened to 12 channels).

What is different underneath: the U-Net's concatenations are free (convs write into column
blocks of one preallocated buffer), each level's CSR is built once and shared by every conv
on it, the two linear heads are fused (the [N,1024] hidden never reaches HBM) together with
the residual add / normalisation, and the vertex->facet transfer is one kernel.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import config, data_util, ops
from .net_util import PoolingLayer, DualFusionLayer  # noqa: F401  (DualFusionLayer re-exported as upstream, network.py:19)
from .nn import FeaStConv, conv_csr, input_graph


class GNNModule(nn.Module):
    """network.py:254-300."""

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
        self.taps = None            # set to {} to record intermediates (tests)

    def forward(self, data_r1, plot_pool=False):
        tap = self.taps

        def rec(key, value):
            if tap is not None:
                tap[key] = value.clone()
            return value

        x0 = data_r1.x
        n1 = x0.size(0)
        g1 = input_graph(data_r1, n1)
        if torch.is_grad_enabled() and (x0.requires_grad or self.l_conv1.lin.weight.requires_grad):
            return self._forward_train(data_r1, g1, rec)
        buf1 = x0.new_empty((n1, 64))                                   # [ l_conv1 | r_conv3 ]  (network.py:298 cat)
        data_r1.x = rec("l1", self.l_conv1(x0, g1, 0.2, out=buf1[:, :32]))
        data_r2 = self.pooling1(data_r1)
        rec("p1", data_r2.x)

        n2 = data_r2.x.size(0)
        g2 = input_graph(data_r2, n2)
        buf2 = ops.valloc(n2, (128,), torch.float32, x0.device, ref=n1)  # [ l_conv2 | r_conv1 ]  (network.py:292 cat)
        data_r2.x = rec("l2", self.l_conv2(data_r2.x, g2, 0.2, out=buf2[:, :64]))
        data_r3 = self.pooling2(data_r2)
        rec("p2", data_r3.x)

        g3 = input_graph(data_r3, data_r3.x.size(0))
        with ops.size_ref(n1):                                          # coarse-level buffers: sizes stable across forwards
            data_r3.x = rec("l3", self.l_conv3(data_r3.x, g3, 0.2))
            data_r3.x = rec("l4", self.l_conv4(data_r3.x, g3, 0.2))
            # unpooling (network.py:289) is fused into the conv: it gathers rows of the coarse features through the map
            rec("r1", self.r_conv1(data_r3.x, g2, 1.0, out=buf2[:, 64:], row_map=self.pooling2.unpool_map))   # no activation (:290)
            data_r2.x = buf2
            data_r2.x = rec("r2", self.r_conv2(buf2, g2, 0.2))

        rec("r3", self.r_conv3(data_r2.x, g1, 1.0, out=buf1[:, 32:], row_map=self.pooling1.unpool_map))       # network.py:295-296
        data_r1.x = buf1
        return rec("r4", self.r_conv4(buf1, g1, 0.2))


    def _forward_train(self, data_r1, g1, rec):
        """Same wiring with autograd-tracked tensors (explicit concatenations instead of the shared output buffers)."""
        data_r1.x = rec("l1", self.l_conv1(data_r1.x, g1, 0.2))
        data_r2 = self.pooling1(data_r1)
        rec("p1", data_r2.x)
        g2 = input_graph(data_r2, data_r2.x.size(0))
        data_r2.x = rec("l2", self.l_conv2(data_r2.x, g2, 0.2))
        data_r3 = self.pooling2(data_r2)
        rec("p2", data_r3.x)
        g3 = input_graph(data_r3, data_r3.x.size(0))
        data_r3.x = rec("l3", self.l_conv3(data_r3.x, g3, 0.2))
        data_r3.x = rec("l4", self.l_conv4(data_r3.x, g3, 0.2))
        up2 = rec("r1", self.r_conv1(self.pooling2.unpooling(data_r3.x), g2, 1.0))
        data_r2.x = torch.cat((data_r2.x, up2), 1)
        data_r2.x = rec("r2", self.r_conv2(data_r2.x, g2, 0.2))
        up1 = rec("r3", self.r_conv3(self.pooling1.unpooling(data_r2.x), g1, 1.0))
        data_r1.x = torch.cat((data_r1.x, up1), 1)
        return rec("r4", self.r_conv4(data_r1.x, g1, 0.2))


class DualGNN(nn.Module):
    """network.py:303-343."""

    def __init__(self, force_depth=False, pool_type="max", edge_weight_type=10, wei_param=2):
        super().__init__()
        self.force_depth = force_depth
        self.gnn_v = GNNModule(6, pool_type, 2, edge_weight_type, wei_param)
        self.fc_v1 = nn.Linear(32, 1024)
        self.fc_v2 = nn.Linear(1024, 1) if force_depth else nn.Linear(1024, 3)
        self.gnn_f = GNNModule(12, pool_type, 2, edge_weight_type, wei_param)
        self.fc_f1 = nn.Linear(32, 1024)
        self.fc_f2 = nn.Linear(1024, 3)
        self.taps = None

    def forward(self, dual_data):
        data_v, data_f = dual_data
        prec = config.precision_code()
        xyz = data_v.x[:, :3]
        if self.taps is not None:
            self.gnn_v.taps, self.gnn_f.taps = {}, {}
        g_v = self.gnn_v(data_v)
        if torch.is_grad_enabled() and g_v.requires_grad:
            return self._forward_train_tail(data_v, data_f, g_v, xyz)
        if self.force_depth:
            feat_v = ops.fc_head_fwd(g_v, self.fc_v1.weight, self.fc_v1.bias, self.fc_v2.weight, self.fc_v2.bias,
                                     epilogue=2, res=xyz, res2=data_v.depth_direction, precision=prec)
        else:
            feat_v = ops.fc_head_fwd(g_v, self.fc_v1.weight, self.fc_v1.bias, self.fc_v2.weight, self.fc_v2.bias,
                                     epilogue=1, res=xyz, precision=prec)
        # vertex -> facet transfer (network.py:335-337): [x_f | corner mean | face normal]
        data_f.x = ops.v2f_transfer(feat_v, data_f.fv_indices, data_f.x)
        xf12 = data_f.x
        g_f = self.gnn_f(data_f)
        norm_f = ops.fc_head_fwd(g_f, self.fc_f1.weight, self.fc_f1.bias, self.fc_f2.weight, self.fc_f2.bias,
                                 epilogue=3, precision=prec)
        if self.taps is not None:
            self.taps = dict(g_v=g_v, feat_v=feat_v, xf12=xf12, g_f=g_f, v=self.gnn_v.taps, f=self.gnn_f.taps)
            self.gnn_v.taps = self.gnn_f.taps = None
        return feat_v, norm_f, None


    def _forward_train_tail(self, data_v, data_f, g_v, xyz):
        """Training step.  Tensor-core precisions: each head is the fused inference kernel with a native backward
        (autograd.HeadFn -> geobi_mlp_head_bwd); precision 'fp32': plain library GEMMs through autograd (cross-check path).
        The transfer runs the inference kernel with a native backward either way."""
        import torch.nn.functional as F
        from .autograd import HeadFn, V2FTransferFn
        prec = config.precision_code()
        native = prec != ops.PREC_FP32 and self.fc_v1.in_features == 32 and self.fc_v1.out_features % 256 == 0

        def head(fc1, fc2, f):
            if native:
                return HeadFn.apply(f, fc1.weight, fc1.bias, fc2.weight, fc2.bias, prec)
            return fc2(F.leaky_relu(fc1(f), 0.2))

        feat_v = head(self.fc_v1, self.fc_v2, g_v)
        if self.force_depth:
            feat_v = feat_v * data_v.depth_direction
        feat_v = feat_v + xyz
        data_f.x = V2FTransferFn.apply(feat_v, data_f.fv_indices, data_f.x)     # geobi_v2f_transfer / geobi_v2f_transfer_bwd
        g_f = self.gnn_f(data_f)
        feat_f = head(self.fc_f1, self.fc_f2, g_f)
        return feat_v, F.normalize(feat_f, dim=1), None


# ---------------------------------------------------------------- losses / metrics (network.py:347-413)
# Elementwise + one reduction over [N,3]; not on the hot path, kept as plain tensor expressions.
def _laplacian(v, edge_idx_v, normal=None):
    row, col = edge_idx_v
    n = v.size(0)
    lap = torch.zeros_like(v).index_add_(0, row, v[row] - v[col])
    cnt = torch.zeros(n, device=v.device, dtype=v.dtype).index_add_(0, row, torch.ones_like(row, dtype=v.dtype))
    lap = lap / cnt.clamp(min=1).unsqueeze(1)
    if normal is not None:
        lap = normal * (lap * normal).sum(1, keepdim=True)
    return lap


def laplacian_loss(vp, v, edge_idx_v, normal=None):
    keep = edge_idx_v[0] != edge_idx_v[1]
    edge_idx_v = edge_idx_v[:, keep]
    return (_laplacian(vp, edge_idx_v, normal) - _laplacian(v, edge_idx_v, normal)).abs().sum(1).mean()


def loss_v(vp, v, dis="L2", apply_icp=False):
    if apply_icp:
        raise NotImplementedError("apply_icp needs pytorch3d upstream (network.py:14-17,365-367)")
    if dis == "L1":
        return (vp - v).abs().sum(1).mean()
    if dis == "L2":
        return (vp - v).pow(2).sum(1).mean()
    raise NotImplementedError(f"loss_v dis={dis!r} needs kaolin upstream (network.py:12-13,369-372)")


def loss_n(np, n, norm="L1", fc_p=None, fc=None):
    if norm == "L1":
        return (np - n).abs().sum(1).mean()
    if norm == "L2":
        return (np - n).pow(2).sum(1).mean()
    raise NotImplementedError(f"loss_n norm={norm!r} needs kaolin upstream (network.py:385-388)")


def dual_loss(loss_v, loss_n, v_scale=1, n_scale=1, alpha=None):
    if alpha is None:
        return loss_v * v_scale + loss_n * n_scale
    return alpha * loss_v * v_scale + (1 - alpha) * loss_n * n_scale


def error_v(vp, v):
    """Euclidean distance."""
    return (vp - v).pow(2).sum(1).pow(0.5).mean()


def error_n(np, n):
    """Intersection angle in degrees."""
    val = torch.clamp(1 - (np - n).pow(2).sum(1) / 2, min=-1, max=1)
    return (torch.acos(val) * 180 / math.pi).mean()
