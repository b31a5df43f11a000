"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

* Inference shards by patch / mesh with NO data-path collective: `batching.shard` deals the units,
  every rank runs graph build + forward locally, outputs are stitched by the owner of the output buffer.
* Training is replica based: one mesh (or patch batch) per rank per micro-step is exactly the reference's gradient
  accumulation over `batch_size` meshes (train_dual.py:211-218); the only collective is ONE all-reduce over a flat fp32
  bucket of the 939 128 parameters' gradients (3.76 MB — latency bound on NVLink 5, so a single bucket).
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.distributed as dist


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True, group=None) -> None:
    """Sum (or average) .grad over ranks through one flat bucket; no-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
