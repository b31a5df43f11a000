#!/usr/bin/env python
"""bench.py — faces/sec of the GeoBi-GNN dual-domain forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

Workload (BASELINE.json configs[1]): a disjoint-union batch of 64 synthetic noisy icosphere patches
(frequency 20 -> 8000 faces / 4002 vertices each, 512 000 faces per GPU), random-init DualGNN
(torch.manual_seed(0)), one full forward per step from the reference's input layout
(x, int64 edge_index, edge_weight, fv_indices).  Weak scaling: every rank gets its own 64 patches,
no collective on the data path.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# level sizes vary from step to step (random matching order): let the caching allocator grow segments in place instead of
# falling back to cudaMalloc / cudaFree (each one a device-wide sync) when a request does not fit a cached block
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True,roundup_power2_divisions:4")

import numpy as np  # noqa: E402
import torch  # noqa: E402

PATCH_FREQ = 20          # icosphere frequency -> 8000 faces per patch ("Synthetic-set shape, ~8k faces")
N_PATCHES = 64
TRACE = bool(os.environ.get("BENCH_TRACE"))
PRIME_STEPS = 16         # untimed allocator-priming forwards before the W warm-up steps (see run_ours)
print_json = None
METRIC = "mesh faces/sec (GeoBi-GNN dual-domain forward)"
UNIT = "faces/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML while the timed region runs (a thread polling every 20 ms;
    cheaper and less intrusive than spawning nvidia-smi, same counters as the recipe's `nvidia-smi --query-gpu` line)."""

    def __init__(self, index):
        self.index, self.rows, self.stop, self.thread, self.h = index, [], threading.Event(), None, None
        self.period = float(os.environ.get("BENCH_CLOCK_PERIOD", "0.02"))

    def __enter__(self):
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.h = None
        return self

    def _poll(self):
        nv = self.nv
        while not self.stop.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                                  else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception:
                pass
            self.stop.wait(self.period)

    def __exit__(self, *a):
        self.stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted({name for _, r in self.rows for name, b in bits.items() if r & b})
        return {"sm_mhz": float(np.median([c for c, _ in self.rows])), "sm_max_mhz": float(self.smax), "reasons": reasons,
                "samples": len(self.rows), "source": "NVML, 20 ms polling during the timed region"}


# ------------------------------------------------------------------------------------- workload
def patch_meshes(count, first_seed):
    from geobi_gnn_b200 import synth
    p, f = synth.icosphere(PATCH_FREQ)
    clean = synth.TriMesh(p, f)
    return [(synth.TriMesh(synth.add_normal_noise(p, f, 0.2, first_seed + i), f), clean) for i in range(count)]


def feast_bytes_alg(n, e, c_in, c_out):
    """compulsory HBM traffic of one FeaSt layer (int32 CSR, fp32 features)."""
    return 4 * (n * c_in + n * c_out) + 4 * e + 4 * (n + 1) + 4 * (9 * c_in * c_out + 9 * c_in + 9 + c_out)


def run_ours(args, rank, world, local_rank):
    from geobi_gnn_b200 import batching, config, dataset, network, ops
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    config.set_precision(args.precision)

    meshes = patch_meshes(N_PATCHES, first_seed=rank * N_PATCHES)
    patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in meshes]
    data_v, data_f, _ = batching.collate_dual(patches)
    del patches
    faces_per_rank = data_f.x.size(0)
    torch.manual_seed(0)
    net = network.DualGNN().to(dev).eval()

    in_keys_v, in_keys_f = ("x", "edge_index", "edge_weight"), ("x", "edge_index", "edge_weight", "fv_indices")
    host_v = {k: getattr(data_v, k).cpu().pin_memory() for k in in_keys_v}
    host_f = {k: getattr(data_f, k).cpu().pin_memory() for k in in_keys_f}
    h2d_bytes = sum(t.numel() * t.element_size() for t in list(host_v.values()) + list(host_f.values()))

    def step_resident():
        with torch.no_grad():
            return net([batching.fresh_view(data_v), batching.fresh_view(data_f)])

    # end to end: the user-facing runner (inference.HostBatchRunner) takes HOST batches; every step uploads one batch from
    # pinned memory (on a copy stream, overlapping the previous batch's compute), runs the forward and reads the outputs back
    from geobi_gnn_b200 import inference
    runner = inference.HostBatchRunner(net, dev, coalesced_undirected=True)
    pipe = {"next": None}

    def step_e2e():
        if pipe["next"] is None:
            pipe["next"] = runner.upload(host_v, host_f)
        cur = pipe["next"]
        pipe["next"] = runner.upload(host_v, host_f)      # this step's H2D copy (the batch the next step consumes)
        return runner.run(cur)                            # this step's compute + D2H read of (vertices, normals)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        import gc
        gc.collect()
        if not os.environ.get("BENCH_KEEP_GC"):
            gc.disable()       # a generation-2 collection inside the timed region shows up as a 10-40 ms host stall
        try:
            return _timed(fn, steps)
        finally:
            gc.enable()

    def _timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        marks = []
        for _ in range(steps):
            fn()
            if TRACE:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((e, time.perf_counter()))
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        if TRACE:
            prev_e, prev_t = ev0, t0
            for e, t in marks:
                print(f"[trace rank {rank}] gpu {prev_e.elapsed_time(e):8.2f} ms  host-issue {1e3 * (t - prev_t):8.2f} ms", file=sys.stderr)
                prev_e, prev_t = e, t
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    # allocator priming: steps run back to back keep more blocks alive than synchronised ones; let the caching allocator
    # reach its high-water mark (a handful of cudaMallocs) before the W warm-up steps so the timed region sees none
    for _ in range(PRIME_STEPS):
        step_resident()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step_resident()
    if args.profile_step:            # for `ncu --profile-from-start off`: exactly one step between cudaProfilerStart/Stop
        torch.cuda.synchronize()
        l0 = ops.launch_count()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(f"[profile-step] libgeobi kernels counted by the host side in this step: {ops.launch_count() - l0}", file=sys.stderr)
        return
    l0 = ops.launch_count()
    with ClockSampler(local_rank) as clk:
        ms, wall = timed(step_resident, args.steps)
    launches = ops.launch_count() - l0
    for _ in range(PRIME_STEPS + args.warmup):
        step_e2e()
    dev_allocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    ms_e2e, wall_e2e = timed(step_e2e, args.steps)
    dev_allocs_e2e = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - dev_allocs0
    d2h_bytes = sum(t.numel() * t.element_size() for t in runner.out_host.values())

    total_faces = faces_per_rank * world
    value = total_faces * args.steps / (ms / 1e3)
    e2e_value = total_faces * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: the fused FeaSt conv on the largest layer (facet r_conv4: N=F, 64->32)
    roof = None
    if rank == 0:
        g = ops.csr_from_coo(data_f.edge_index, faces_per_rank, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
        conv = net.gnn_f.r_conv4
        x = torch.randn(faces_per_rank, 64, device=dev)
        out = torch.empty(faces_per_rank, 32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
        prec = config.precision_code()
        fused = prec == ops.PREC_BF16X3
        call = lambda p_=prec: ops.feast_fwd(x, g, conv.lin.weight, conv.u.weight, conv.c, conv.bias, 0.2, out=out, precision=p_)
        for _ in range(3):
            call()

        def time_calls(p_):
            ts = []
            for _ in range(10):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                call(p_)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return float(np.mean(ts))

        t_op = time_calls(prec)                                   # whole layer: projection P + weight split + main kernel
        t_ms = time_calls(prec | ops.FEAST_REUSE_WS) if fused else t_op   # the dominant kernel alone
        alg = feast_bytes_alg(faces_per_rank, g.nnz + faces_per_rank, 64, 32)
        peak, how = peaks()
        achieved = alg / (t_ms / 1e3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this layer from the committed ncu --set full capture
        # (profiles/r01_ncu_full_feast_fused_v2.csv): 226.1 MB + 51.7 MB
        traffic = 277_800_960 if fused else None
        roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic,
                "kernel": ("feast_fused_64_32_kernel" if fused else "geobi_feast_fwd (project + aggregate + gemm)") +
                          " on facet r_conv4: N=%d, E=%d incl. self loops, 64->32" % (faces_per_rank, g.nnz + faces_per_rank),
                "alg_bytes_per_launch": alg, "ms_per_launch": round(t_ms, 4), "ms_whole_layer": round(t_op, 4), "peak_source": how,
                "l2": "flushed (256 MB write) between launches", "traffic_source": "ncu --set full, profiles/r01_ncu_full_feast_fused_v2.csv",
                "note": "HBM is the nominal bound of a gather/segment-sum; this kernel's 9-head weighting costs 576 fp32 FMAs per gathered 256-byte "
                        "row, so at 100 % of the FP32 pipe (FFMA2 = 2 clk, profiles/micro/ffma2_bench.cu) it could reach ~0.3 of the HBM peak; "
                        "ncu: fmaheavy pipe 47 % active, issue slots 55 %, DRAM traffic 1.23x the algorithmic bytes"}

    cpu = cpu_baseline(sample_seconds=12.0) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16": "bf16 tensor-core projections (1 pass), f32 elsewhere",
                          "bf16x3": "f32-grade: split-bf16 x3 tensor-core projections with f32 accumulate, f32 elsewhere"}[args.precision], "data": "synthetic",
                "config": {"workload": "configs[1]: 64 noisy icosphere patches x 8000 faces per GPU, disjoint-union batch, DualGNN fwd, random init",
                           "faces_per_gpu": faces_per_rank, "precision": args.precision,
                           "l2": "per-step working set (inputs 150 MB + >2 GB intermediates) exceeds the 126 MB L2",
                           "timing": "CUDA events on the launch stream, max over ranks", "wall_s": round(wall, 4),
                           "priming": f"{PRIME_STEPS} untimed forwards before the {args.warmup} warm-up steps (caching-allocator high-water mark)"},
                "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": round(ms_e2e / args.steps, 4), "cuda_mallocs_in_timed_region": int(dev_allocs_e2e),
                        "path": "inference.HostBatchRunner: pinned host batch -> H2D + input-level CSR build on a copy stream (under the "
                                "previous batch's forward) -> DualGNN forward -> D2H of vertices and normals on a read-back stream; one "
                                "upload + one forward + one read-back per step"},
                "gpu_launches": launches, "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu}
        print_json(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------- CPU legs (oracle)
def _oracle_patch_inputs(count, first_seed=0):
    from oracle import ref_dataset
    return [ref_dataset.build_dual_data(mn, mo) for mn, mo in patch_meshes(count, first_seed)]


def cpu_baseline(sample_seconds):
    """Oracle ('port' of the reference path; the reference itself cannot be imported here) timed on the host cores
    on a bounded sample of the same workload: whole patches, one forward each, until ~sample_seconds elapsed."""
    from oracle import ref_network
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    net = ref_network.DualGNN().eval()
    pool = _oracle_patch_inputs(4)
    with torch.no_grad():
        net([pool[0][0].clone(), pool[0][1].clone()])        # warm-up
        done, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < sample_seconds or done < 3:
            dv, df = pool[done % len(pool)]
            net([dv.clone(), df.clone()])
            done += 1
        dt = time.perf_counter() - t0
    faces = done * 20 * PATCH_FREQ ** 2
    return {"value": round(faces / dt, 1), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} patches of {20 * PATCH_FREQ ** 2} faces, one oracle forward each, {dt:.1f} s"}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU path for this metric.  The reference cannot be imported or installed
    here (torch_geometric / torch_scatter / torch_sparse / torch_cluster / openmesh absent), so this
    times its restatement in oracle/ (kind 'port') with all host threads; each step = a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import ref_network
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    net = ref_network.DualGNN().eval()
    per_step = 2
    pool = _oracle_patch_inputs(per_step)

    def step():
        with torch.no_grad():
            for dv, df in pool:
                net([dv.clone(), df.clone()])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    faces = per_step * 20 * PATCH_FREQ ** 2
    value = faces * args.steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 64 noisy icosphere patches x 8000 faces per GPU, disjoint-union batch, DualGNN fwd, random init",
                       "sample": f"{per_step} of the 64 patches per step (the reference runs one patch per forward, dataset.py:29-31)"},
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} patches x {20 * PATCH_FREQ ** 2} faces per step, {args.steps} steps"},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_json(line)


def main():
    # stdout carries exactly ONE line (the JSON): anything a library prints there (e.g. "NCCL version ...") goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global print_json
    def print_json(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("GEOBI_PRECISION", "bf16x3"), choices=["fp32", "bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true", help="run warm-up then ONE step inside cudaProfilerStart/Stop and exit")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: geobi_gnn_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run --nproc-per-node N (see README)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
