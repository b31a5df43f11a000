"""Training step with the reference's semantics (train_dual.py:199-218) on N replicas.

The reference accumulates gradients over `batch_size` meshes (one mesh per forward) and then steps Adam.  Here each rank
runs ONE micro-step on its own mesh / patch batch, gradients are summed over ranks with a single flat fp32 all-reduce
(NCCL over NVLink; parallel.allreduce_gradients) and every rank applies the same optimiser step — i.e. `world_size`
plays the role of the reference's `batch_size`, loss pre-scaled by 1/world_size exactly as `train_loss /= opt.batch_size`.
"""
from __future__ import annotations

import torch

from . import network, parallel


def train_step(net, optimizer, dual_data, loss_v="L1", loss_n="L1", v_scale=1.0, n_scale=1.0, world_size=1):
    """One forward / loss / backward / all-reduce / optimiser step.  Returns (loss, error_v, error_n) tensors of this rank."""
    data_v, data_f = dual_data
    y_v, y_f = data_v.y, data_f.y
    optimizer.zero_grad(set_to_none=True)
    vert_p, norm_p, _ = net([data_v, data_f])
    l_v = network.loss_v(vert_p, y_v, loss_v)
    l_f = network.loss_n(norm_p, y_f, loss_n)
    loss = network.dual_loss(l_v, l_f, v_scale=v_scale, n_scale=n_scale)
    (loss / world_size).backward()
    parallel.allreduce_gradients(net.parameters(), average=False)
    optimizer.step()
    with torch.no_grad():
        return loss.detach(), network.error_v(vert_p, y_v), network.error_n(norm_p, y_f)
