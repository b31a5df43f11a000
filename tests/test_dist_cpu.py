"""CPU, world_size 2 over gloo: the host-side multi-rank logic (patch sharding with no collective on the data path,
max-over-ranks timing reduction, flat gradient all-reduce) — what bench.py and parallel.py do on N GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geobi_gnn_b200 import batching, parallel
    from oracle import ref_network
    # 1. patches dealt round-robin, every patch owned by exactly one rank
    items = list(range(7))
    mine = batching.shard(items, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # 2. timing reduction used by bench.py: max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # 3. flat-bucket gradient all-reduce == single-process accumulation (train_dual.py:211-218)
    torch.manual_seed(0)
    net = ref_network.DualGNN()
    torch.manual_seed(100 + rank)
    for p in net.parameters():
        p.grad = torch.randn_like(p)
    local = [p.grad.clone() for p in net.parameters()]
    parallel.allreduce_gradients(net.parameters(), average=True)
    all_local = [None] * world
    dist.all_gather_object(all_local, local)
    want = [sum(g[i] for g in all_local) / world for i in range(len(local))]
    err = max(float((p.grad - w).abs().max()) for p, w in zip(net.parameters(), want))
    if rank == 0:
        q.put((gathered, float(t), err))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax, err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(sum(gathered, [])) == list(range(7)) and gathered[0] == [0, 2, 4, 6] and gathered[1] == [1, 3, 5]
    assert tmax == 11.0
    assert err < 1e-6


def test_stitch_matches_oracle_layout():
    """Patch stitch (test_dual.py:49-61) bookkeeping: visit counts and averaging, on CPU tensors via the oracle."""
    from oracle import ref_dataset
    r = [(torch.ones(3, 3), torch.eye(3), torch.tensor([0, 1, 2]), torch.tensor([0, 1, 2])),
         (3 * torch.ones(2, 3), torch.eye(3)[:2], torch.tensor([2, 3]), torch.tensor([2, 3]))]
    vp, nrm = ref_dataset.stitch_patches(4, 4, r)
    assert vp[2].tolist() == [2.0, 2.0, 2.0] and vp[3].tolist() == [3.0, 3.0, 3.0]
