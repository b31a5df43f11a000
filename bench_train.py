#!/usr/bin/env python
"""Config 5 of BASELINE.json: train_dual.py-style training step (fwd + L1 losses + bwd + flat gradient all-reduce + Adam)
on a batch of synthetic patches per rank, at 1/2/4/8 GPUs.  Not the driver's headline bench (that is bench.py); prints
one JSON line with faces/s per training step.

    python bench_train.py [--steps K] [--warmup W] [--patches P]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_train.py ...
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1, help="informational; the world size comes from torchrun (WORLD_SIZE)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--patches", type=int, default=16, help="patches (8000 faces each) per rank per step")
    ap.add_argument("--precision", default="bf16x3")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    from geobi_gnn_b200 import batching, config, dataset, network, train
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    config.set_precision(args.precision)
    patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(args.patches, rank * args.patches)]
    data_v, data_f, _ = batching.collate_dual(patches)
    torch.manual_seed(0)
    net = network.DualGNN().to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)

    def step():
        return train.train_step(net, opt, [batching.fresh_view(data_v), batching.fresh_view(data_f)], world_size=world)

    for _ in range(args.warmup + 4):
        loss, ev, en = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, ev, en = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    faces = data_f.x.size(0) * world
    if rank == 0:
        print(json.dumps({"metric": "mesh faces/sec (GeoBi-GNN training step: fwd+loss+bwd+allreduce+Adam)", "value": round(faces * args.steps / (ms / 1e3), 1),
                          "unit": "faces/s", "n_gpus": world, "steps": args.steps, "ms_per_step": round(ms / args.steps, 3), "scaling": "weak",
                          "config": {"workload": f"configs[4]: {args.patches} patches x 8000 faces per rank per step, Adam lr 1e-3, L1 losses",
                                     "precision": args.precision, "allreduce": "one flat fp32 bucket, 939128 elements, NCCL"},
                          "loss": float(loss), "error_n_deg": float(en)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
