"""Autograd wrappers for the training step (train_dual.py:199-218).

Forward passes are the same libgeobi kernels as inference.  Backward passes: a FeaStConv layer's whole backward is ONE library
call (`geobi_feast_bwd`: dZ = g.W_flat and dW = g^T.Z on tcgen05 with split bf16 operands, the edge part, dX += dP.U, dU = dP^T.X)
in the tensor-core precisions; with precision 'fp32' the edge part and the max-pool routing are libgeobi kernels
(`geobi_feast_bwd_edges`, `geobi_segment_max_bwd`) and the dense parts are plain library GEMMs (torch.matmul, fp32) - the
cross-check path of tests/test_gpu_train.py.
"""
from __future__ import annotations


import torch

from . import _lib, ops
from .ops import _count, _ptr, _rows, _stream

H = 9


def feast_aggregate(x, g, U, c):
    """(P fp64 [N,9], Z fp32 [N, 9*C_in]) of the forward."""
    lib = _lib.load()
    x, ldx, c_in = _rows(x)
    n = x.size(0)
    P = torch.empty((n, H), dtype=torch.float64, device=x.device)
    Z = torch.empty((n, H * c_in), dtype=torch.float32, device=x.device)
    _lib.check(lib.geobi_feast_aggregate(_ptr(x), ldx, n, c_in, _ptr(g.rowptr), _ptr(g._nbr), _ptr(U.contiguous()), _ptr(c.contiguous()),
                                         _ptr(P), _ptr(Z), _stream()), "feast_aggregate")
    _count(2)
    return P, Z


def feast_bwd_edges(x, g, P, c, dZ, need_dx=True):
    """need_dx=False (the layer's input needs no gradient: the network's two input layers): dx is neither allocated nor accumulated."""
    lib = _lib.load()
    x, ldx, c_in = _rows(x)
    n = x.size(0)
    if not need_dx and not (c_in <= 16 or c_in in (32, 64, 128)):
        need_dx = True                      # the generic kernel always accumulates it
    dx = torch.zeros((n, c_in), dtype=torch.float32, device=x.device) if need_dx else None
    dP = torch.zeros((n, H), dtype=torch.float32, device=x.device)
    dc = torch.zeros(H, dtype=torch.float32, device=x.device)
    _lib.check(lib.geobi_feast_bwd_edges(_ptr(x), ldx, n, c_in, _ptr(g.rowptr), _ptr(g._nbr), _ptr(P), _ptr(c.contiguous()),
                                         _ptr(dZ.contiguous()), _ptr(dx), c_in, _ptr(dP), _ptr(dc), _stream()), "feast_bwd_edges")
    _count()
    return dx, dP, dc


def feast_bwd(x, g, W, U, c, out, go, slope, need_dx=True):
    """geobi_feast_bwd: (dx | None, dW, dU, dc, dbias) of one FeaStConv layer from the gradient `go` of its (activated) output."""
    lib = _lib.load()
    x, ldx, c_in = _rows(x)
    n, c_out = x.size(0), go.size(1)
    go, ldg, _ = _rows(go)
    if ldg % 4 or go.data_ptr() % 16:
        go, ldg = go.contiguous(), c_out
    o, ldo = None, 0
    if slope != 1.0:
        o, ldo, _ = _rows(out)
        if ldo % 4 or o.data_ptr() % 16:
            o, ldo = o.contiguous(), c_out
    dev = x.device
    dx = torch.empty((n, c_in), dtype=torch.float32, device=dev) if need_dx else None
    dW = torch.empty((H * c_out, c_in), dtype=torch.float32, device=dev)
    dU = torch.empty((H, c_in), dtype=torch.float32, device=dev)
    dc = torch.empty(H, dtype=torch.float32, device=dev)
    db = torch.empty(c_out, dtype=torch.float32, device=dev)
    ws = ops._ws(lib.geobi_feast_bwd_ws_bytes(n, c_in, c_out), dev, slot=2)
    _lib.check(lib.geobi_feast_bwd(_ptr(x), ldx, n, c_in, _ptr(g.rowptr), _ptr(g._nbr), _ptr(W.contiguous()), _ptr(U.contiguous()),
                                   _ptr(c.contiguous()), c_out, float(slope), _ptr(o), ldo, _ptr(go), ldg, _ptr(dx), c_in, _ptr(dW),
                                   _ptr(dU), _ptr(dc), _ptr(db), _ptr(ws), ws.numel(), _stream()), "feast_bwd")
    kpad, chunks = -(-H * c_in // 64) * 64, 0
    while kpad > 0:                         # dZ is produced in column chunks of 256 / 128 / 64
        kpad -= 256 if kpad >= 256 else (128 if kpad >= 128 else 64)
        chunks += 1
    _count(8 + chunks)                      # prep g, prep W^T, projection, aggregation, dZ chunks, edges, split-K, reduce, dP.U
    return dx, dW, dU, dc, db


class FeaStFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, U, c, bias, g, act_slope, precision):
        out = ops.feast_fwd(x.detach(), g, W.detach(), U.detach(), c.detach(), bias.detach(), act_slope=act_slope, precision=precision)
        ctx.save_for_backward(x, W, U, c, out)
        ctx.g, ctx.slope, ctx.native = g, act_slope, (precision & 0xff) == ops.PREC_BF16X3
        return out

    @staticmethod
    def backward(ctx, go):
        x, W, U, c, out = ctx.saved_tensors
        g, slope = ctx.g, ctx.slope
        if ctx.native:
            dx, dW, dU, dc, dbias = feast_bwd(x, g, W, U, c, out, go, slope, ctx.needs_input_grad[0])
            return dx, dW, dU, dc, dbias, None, None, None
        c_in, c_out = x.size(1), out.size(1)
        gpre = go if slope == 1.0 else go * torch.where(out > 0, 1.0, slope)   # leaky_relu keeps the sign of its input
        gpre = gpre.contiguous()
        dbias = gpre.sum(0)
        P, Z = feast_aggregate(x, g, U, c)
        Wf = W.view(H, c_out, c_in).permute(1, 0, 2).reshape(c_out, H * c_in)
        dW = (gpre.t() @ Z).view(c_out, H, c_in).permute(1, 0, 2).reshape(H * c_out, c_in)
        dZ = gpre @ Wf
        need_dx = ctx.needs_input_grad[0]
        dx, dP, dc = feast_bwd_edges(x, g, P, c, dZ, need_dx)
        dx = dx + dP @ U if need_dx else None
        dU = dP.t() @ x
        return dx, dW, dU, dc, dbias, None, None, None


def mlp_head_bwd(f, W1, b1, W2, dy, slope=0.2, need_df=True):
    """geobi_mlp_head_bwd: (df | None, dW1, db1, dW2, db2) of y = W2.leaky_relu(W1.f + b1) + b2 from dy = dL/dy."""
    lib = _lib.load()
    f, ldf, c_in = _rows(f)
    if ldf % 4 or f.data_ptr() % 16:
        f, ldf = f.contiguous(), c_in
    dy, lddy, c_out = _rows(dy)
    n, hidden, dev = f.size(0), W1.size(0), f.device
    df = torch.empty((n, c_in), dtype=torch.float32, device=dev) if need_df else None
    dW1 = torch.empty((hidden, c_in), dtype=torch.float32, device=dev)
    db1 = torch.empty(hidden, dtype=torch.float32, device=dev)
    dW2 = torch.empty((c_out, hidden), dtype=torch.float32, device=dev)
    db2 = torch.empty(c_out, dtype=torch.float32, device=dev)
    ws = ops._ws(lib.geobi_mlp_head_bwd_ws_bytes(n, c_in, hidden), dev, slot=2)
    _lib.check(lib.geobi_mlp_head_bwd(_ptr(f), ldf, n, c_in, _ptr(W1.contiguous()), _ptr(b1.contiguous()), hidden, _ptr(W2.contiguous()), c_out,
                                      float(slope), _ptr(dy), lddy, _ptr(df), c_in, _ptr(dW1), _ptr(db1), _ptr(dW2), _ptr(db2), _ptr(ws),
                                      ws.numel(), _stream()), "mlp_head_bwd")
    _count(9 if need_df else 8)     # split f, W1 prep, split dy, hidden planes, 2 x (split-K + reduce), df GEMM
    return df, dW1, db1, dW2, db2


class HeadFn(torch.autograd.Function):
    """fc2(leaky_relu(fc1(f), 0.2)) (network.py:324-325,340-341): the fused inference kernel forward (the [N,1024] hidden never
    reaches HBM), geobi_mlp_head_bwd backward; the head's epilogue (residual / force_depth / normalize) stays with the caller."""

    @staticmethod
    def forward(ctx, f, W1, b1, W2, b2, precision):
        ctx.save_for_backward(f, W1, b1, W2)
        return ops.fc_head_fwd(f.detach(), W1.detach(), b1.detach(), W2.detach(), b2.detach(), epilogue=0, precision=precision)

    @staticmethod
    def backward(ctx, dy):
        f, W1, b1, W2 = ctx.saved_tensors
        df, dW1, db1, dW2, db2 = mlp_head_bwd(f, W1, b1, W2, dy, 0.2, ctx.needs_input_grad[0])
        return df, dW1, db1, dW2, db2, None


class SegmentMaxFn(torch.autograd.Function):
    """scatter(x, cluster, reduce='max') through the cluster-member CSR."""

    @staticmethod
    def forward(ctx, x, mrowptr, members, n_seg):
        out = ops.segment_reduce(x.detach(), mrowptr, members, n_seg, ops.OP_MAX)
        ctx.save_for_backward(x, mrowptr, members)
        ctx.n_seg = n_seg
        return out

    @staticmethod
    def route(x, mrowptr, members, n_seg, go):
        lib = _lib.load()
        xr, ldx, c = _rows(x)
        go = go.contiguous()
        dx = torch.zeros((x.size(0), c), dtype=torch.float32, device=x.device)
        _lib.check(lib.geobi_segment_max_bwd(_ptr(xr), ldx, c, _ptr(mrowptr), _ptr(members), n_seg, _ptr(go), c, _ptr(dx), c,
                                             _stream()), "segment_max_bwd")
        _count()
        return dx

    @staticmethod
    def backward(ctx, go):
        x, mrowptr, members = ctx.saved_tensors
        return SegmentMaxFn.route(x, mrowptr, members, ctx.n_seg, go), None, None, None


class PoolStepFn(torch.autograd.Function):
    """scatter(x, cluster, max | mean) whose forward value was already produced by geobi_pool_step (one library call for the whole
    coarsening step, net_util.py:100-140); this node only routes the gradient: arg-max member (geobi_segment_max_bwd) or 1/count."""

    @staticmethod
    def forward(ctx, x, pooled, mrowptr, members, n_seg, op, cluster):
        ctx.save_for_backward(x, mrowptr, members, cluster)
        ctx.n_seg, ctx.op = n_seg, op
        return pooled[0].view_as(pooled[0])

    @staticmethod
    def backward(ctx, go):
        x, mrowptr, members, cluster = ctx.saved_tensors
        if ctx.op == ops.OP_MAX:
            return SegmentMaxFn.route(x, mrowptr, members, ctx.n_seg, go), None, None, None, None, None, None
        cnt = (mrowptr[1:] - mrowptr[:-1]).clamp(min=1).to(go.dtype)
        return ops.gather_rows(go / cnt.unsqueeze(1), cluster), None, None, None, None, None, None


class SegmentMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mrowptr, members, n_seg, cluster):
        out = ops.segment_reduce(x.detach(), mrowptr, members, n_seg, ops.OP_MEAN)
        ctx.save_for_backward(mrowptr, cluster)
        return out

    @staticmethod
    def backward(ctx, go):
        mrowptr, cluster = ctx.saved_tensors
        cnt = (mrowptr[1:] - mrowptr[:-1]).clamp(min=1).to(go.dtype)
        return ops.gather_rows(go / cnt.unsqueeze(1), cluster), None, None, None, None


class GatherRowsFn(torch.autograd.Function):
    """x[idx] (PoolingLayer.unpooling, net_util.py:242-245); backward = the sum of the gradient rows over each coarse node's fine
    nodes, as a CSR segment sum (geobi_group_by + geobi_segment_reduce: fixed summation order, no atomics)."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.n = x.size(0)
        return ops.gather_rows(x.detach(), idx)

    @staticmethod
    def backward(ctx, go):
        (idx,) = ctx.saved_tensors
        mrowptr, members = ops.group_by(idx, ctx.n)
        return ops.segment_reduce(go.contiguous(), mrowptr, members, ctx.n, ops.OP_SUM), None


class V2FTransferFn(torch.autograd.Function):
    """[xf | corner mean of feat_v | face normal of feat_v] (network.py:335-337) with a native backward for feat_v
    (geobi_v2f_transfer_bwd); xf is an input feature block and gets no gradient."""

    @staticmethod
    def forward(ctx, feat_v, fv, xf):
        ctx.save_for_backward(feat_v, fv)
        ctx.cf = xf.size(1)
        return ops.v2f_transfer(feat_v.detach(), fv, xf)

    @staticmethod
    def backward(ctx, go):
        feat_v, fv = ctx.saved_tensors
        lib = _lib.load()
        p, ldv, _ = _rows(feat_v)
        go = go.contiguous()
        g = go[:, ctx.cf:]
        d = torch.zeros((p.size(0), 3), dtype=torch.float32, device=p.device)
        _lib.check(lib.geobi_v2f_transfer_bwd(_ptr(p), ldv, _ptr(fv.contiguous().long()), _ptr(g), go.stride(0), fv.size(0), _ptr(d), 3,
                                              _stream()), "v2f_transfer_bwd")
        _count()
        return d, None, None
