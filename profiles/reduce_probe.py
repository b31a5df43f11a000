"""NCCL reduce of the stitch accumulators' size (20 + 60 + 120 MB) onto rank 0: time per call at this world size."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ts = [torch.ones(n, device=dev) for n in (5_000_000, 15_000_000, 30_000_000)]
for rep in range(4):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for t in ts:
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0:
        print(f"world {world} rep {rep}: reduce of 200 MB in {1e3 * (t1 - t0):.2f} ms; peer access 0->1: {torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None}")
dist.destroy_process_group()
