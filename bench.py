#!/usr/bin/env python
"""bench.py — faces/sec of the GeoBi-GNN dual-domain forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16|bf16x3] [--workload mesh1m|patches]

Headline workload (BASELINE.json configs[2], the largest single-GPU configuration): ONE synthetic noisy mesh of 1 003 520 faces /
501 762 vertices (icosphere frequency 224) per GPU, full vertex + facet graph, random-init DualGNN (torch.manual_seed(0)), one full
forward per step from the reference's input layout (x, int64 edge_index, edge_weight, fv_indices).  Weak scaling: every rank gets
its own mesh (own noise seed), no collective on the data path.  One JSON line on stdout (rank 0).  The same line
also carries, as sub-objects measured in the same run:
  * "configs1_patches64": BASELINE configs[1] (disjoint-union batch of 64 x 8000-face patches per GPU), value + e2e;
  * "configs3_mesh10m":   BASELINE configs[3] - ONE 10 025 280-face mesh cut into the reference's BFS patches of <= 1 M faces, the
    patches dealt to the N ranks (STRONG scaling, no data-path collective; the accumulators are summed onto rank 0 at the end).
"""
from __future__ import annotations

import argparse
import json
import os
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
MESH_FREQ = 224          # configs[2]: 20 n^2 = 1 003 520 faces
BIG_FREQ = 708           # configs[3]: 10 025 280 faces
BIG_SUB = 1_000_000      # sub_size of the patch walk for configs[3]
CPU_SUB = 20_000         # the reference's CPU path cannot hold a 1 M-face graph pair (15 GB per per-edge tensor): it
                         # runs the mesh patch-wise with test_dual.py's sub_size, which is what the CPU legs time
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


# ------------------------------------------------------------------------------------- workloads
def patch_meshes(count, first_seed):
    from geobi_gnn_b200 import synth
    p, f = synth.icosphere(PATCH_FREQ)
    clean = synth.TriMesh(p, f)
    return [(synth.TriMesh(synth.add_normal_noise(p, f, 0.2, first_seed + i), f), clean) for i in range(count)]


def noisy_device_mesh(freq, seed, dev):
    """Noisy icosphere as a topology.DeviceTriMesh: p += 0.2 * mean edge length * N(0,1) * vertex normal, noise drawn
    on the device (generator seeded per rank)."""
    import torch
    from geobi_gnn_b200 import synth, topology
    p, f = synth.icosphere(freq)
    clean = topology.DeviceTriMesh(p, f, dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    amp = torch.randn(clean.n_vertices, 1, generator=g, device=dev) * 0.2 * clean.mean_edge_length()
    mesh = topology.DeviceTriMesh(clean.points + amp * clean.vertex_normals, clean.fv, dev)
    return mesh


def feast_bytes_alg(n, e, c_in, c_out):
    """compulsory HBM traffic of one FeaSt layer (int32 CSR, fp32 features)."""
    return 4 * (n * c_in + n * c_out) + 4 * e + 4 * (n + 1) + 4 * (9 * c_in * c_out + 9 * c_in + 9 + c_out)


FORWARD_BYTES_PER_FACE = 3104     # algorithmic bytes of the whole dual forward (16 conv layers + 2 heads) per face


def ncu_traffic(kernel_substr, n_nodes):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed `ncu --set full` capture
    that matches this kernel and layer size (profiles/ncu_traffic.json lists the captures: kernel, nodes, csv).  None when no capture
    of the current kernel at this size is committed - a stale constant would be worse than no number."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    for ent in json.load(open(path)):
        if ent["kernel"] in kernel_substr and int(ent["nodes"]) == int(n_nodes):
            csv_path = os.path.join(ROOT, "profiles", ent["csv"])
            if not os.path.exists(csv_path):
                continue
            import csv
            rows = list(csv.reader(open(csv_path)))
            hdr, data = rows[0], rows[2:]
            if int(ent.get("row", 0)) >= len(data) or not any(ent["kernel"] in c for c in data[int(ent.get("row", 0))]):
                continue
            d = dict(zip(hdr, data[int(ent.get("row", 0))]))
            unit = dict(zip(hdr, rows[1]))
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(d[key])
                tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit[key]]
            return int(tot), f"ncu --set full, profiles/{ent['csv']}"
    return None, None


class Workload:
    """Device-resident inputs of one rank + the host copies the end-to-end path uploads every step."""

    def __init__(self, name, rank, dev):
        import torch
        from geobi_gnn_b200 import batching, dataset
        self.name = name
        if name == "patches":
            patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in patch_meshes(N_PATCHES, first_seed=rank * N_PATCHES)]
            self.dv, self.df, _ = batching.collate_dual(patches)
            self.dv_res, self.df_res = self.dv, self.df        # the union batch in the reference layout (lists; sort-free CSRs built per step)
            self.describe = ("configs[1]: 64 noisy icosphere patches x 8000 faces per GPU, disjoint-union batch, DualGNN fwd, random init")
        else:
            mesh = noisy_device_mesh(MESH_FREQ, rank, dev)
            self.dv, self.df = dataset.build_dual_on_device(mesh, None)                       # reference layout (int64 lists): what the prebuilt path ships
            # resident inputs of `value`: the graph pair as the device front end hands it to the network (features + loop-free CSRs with
            # their weights, Data.csr) - the same inputs the end-to-end path's forward gets
            self.dv_res, self.df_res = dataset.build_dual_on_device(mesh, None, csr_native=True)
            # what the end-to-end path uploads for a single mesh: the raw mesh (the device front end builds the graphs every step)
            self.host_points = mesh.points.cpu().pin_memory()
            self.host_faces = mesh.fv.to(torch.int32).cpu().pin_memory()
            self.describe = ("configs[2]: one noisy icosphere mesh of 1 003 520 faces / 501 762 vertices per GPU (frequency 224), full vertex + "
                             "facet graph, DualGNN fwd, random init")
            del mesh
        self.faces = self.df.x.size(0)
        keys_v, keys_f = ("x", "edge_index", "edge_weight"), ("x", "edge_index", "edge_weight", "fv_indices")
        from geobi_gnn_b200.inference import HostBatchRunner
        # the host batch as a data-loader worker hands it over: float tensors as they are, index tensors narrowed to int32 once
        self.host_v = HostBatchRunner.pack({k: getattr(self.dv, k).cpu().pin_memory() for k in keys_v})
        self.host_f = HostBatchRunner.pack({k: getattr(self.df, k).cpu().pin_memory() for k in keys_f})
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in list(self.host_v.values()) + list(self.host_f.values()))
        self.h2d_mesh_bytes = None if name == "patches" else (self.host_points.numel() * 4 + self.host_faces.numel() * 4)


def run_ours(args, rank, world, local_rank):
    from geobi_gnn_b200 import batching, config, network, ops, inference
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    config.set_precision(args.precision)
    torch.manual_seed(0)
    net = network.DualGNN().to(dev).eval()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if world == 1:
            return vals
        import torch.distributed as dist
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in t)

    def timed(fn, steps, join=None, before_close=None):
        """K steps between two CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks.
        `join`: streams whose queued work belongs to the steps (copy / read-back streams of the end-to-end runner) - the launch stream
        waits for them before the closing event, so the last step's transfers are inside the timed region."""
        import gc
        gc.collect()
        if not os.environ.get("BENCH_KEEP_GC"):
            gc.disable()       # a generation-2 collection inside the timed region shows up as a 10-40 ms host stall
        try:
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record()
            for _ in range(steps):
                fn()
            if before_close is not None:
                before_close()       # e.g. wait for the helper thread that is still queueing the trailing upload
            for s in (join or ()):
                torch.cuda.current_stream(dev).wait_stream(s)
            ev1.record()
            barrier()
            wall = time.perf_counter() - t0
            ms = ev0.elapsed_time(ev1)
        finally:
            gc.enable()
        ms, wall = max_over_ranks(ms, wall)
        return ms, wall

    # ---------------------------------------------------------------- concurrent lanes
    # W host threads, each with its own CUDA stream, module replica (same weights; PoolingLayer keeps per-forward state on the module)
    # and end-to-end runner, take steps from one counter: the latency-bound stretches of a forward (matcher on the coarse levels, scans,
    # the ten count read-backs) overlap with another mesh's kernels.  profiles/two_stream_probe.py: 12.2 -> 10.8 ms per 1 M-face mesh
    # with two lanes, 10.7 with three.  W = 1 (--streams 1) is the single-stream figure, reported beside it.
    import copy
    import itertools
    import threading
    W = max(1, int(args.streams))
    nets = [net] + [copy.deepcopy(net).eval() for _ in range(W - 1)]
    lane_streams = [torch.cuda.Stream(dev) for _ in range(W)]

    def run_lanes(fns, steps):
        """`steps` calls in total, dealt to len(fns) threads (thread i calls fns[i] on lane stream i)."""
        ticket = itertools.count()
        errs = []

        def work(i):
            try:
                torch.cuda.set_device(dev)
                with torch.cuda.stream(lane_streams[i]):
                    while next(ticket) < steps:
                        fns[i]()
            except BaseException as e:      # surfaced by the caller: a failed lane must not look like a fast one
                errs.append(e)

        ths = [threading.Thread(target=work, args=(i,)) for i in range(len(fns))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]

    def timed_lanes(fns, steps, join=(), before_close=None):
        """timed() for concurrent lanes: events on the launch stream around all lanes' work (the launch stream waits for every lane
        stream and for `join` before the closing event)."""
        import gc
        gc.collect()
        if not os.environ.get("BENCH_KEEP_GC"):
            gc.disable()
        try:
            barrier()
            main = torch.cuda.current_stream(dev)
            for s_ in lane_streams:
                s_.wait_stream(main)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record()
            run_lanes(fns, steps)
            if before_close is not None:
                before_close()
            for s_ in list(lane_streams) + list(join):
                main.wait_stream(s_)
            ev1.record()
            barrier()
            wall = time.perf_counter() - t0
            ms = ev0.elapsed_time(ev1)
        finally:
            gc.enable()
        ms, wall = max_over_ranks(ms, wall)
        return ms, wall

    def measure(wl, steps, warmup, with_clocks):
        """Device-timed value (inputs resident) and end-to-end value (host batches through inference.HostBatchRunner) of a workload."""
        def resident(i):
            def step():
                with torch.no_grad():
                    return nets[i]([batching.fresh_view(wl.dv_res), batching.fresh_view(wl.df_res)])
            return step

        res_fns = [resident(i) for i in range(W)]
        runners = [inference.HostBatchRunner(nets[i], dev, coalesced_undirected=True) for i in range(W)]
        pipes = [{"next": None} for _ in range(W)]
        sync_upload = bool(os.environ.get("GEOBI_BENCH_SYNC_UPLOAD"))

        def make_step(i, kind):
            r = runners[i]
            if kind == "graphs":      # prebuilt graphs cross PCIe
                up = (lambda: r.upload(wl.host_v, wl.host_f)) if sync_upload else (lambda: r.upload_async(wl.host_v, wl.host_f))
            else:                     # the raw mesh; its front end is queued by the runner's helper thread while this thread queues a forward
                up = (lambda: r.upload_mesh(wl.host_points, wl.host_faces)) if sync_upload else (lambda: r.upload_mesh_async(wl.host_points, wl.host_faces))

            def step():
                if pipes[i]["next"] is None:
                    pipes[i]["next"] = up()
                cur = pipes[i]["next"]
                pipes[i]["next"] = up()                                 # this step's H2D copy (the batch the next step of this lane consumes)
                return r.run(cur)                                       # this step's compute + D2H read of (vertices, normals)
            return step

        # allocator priming: steps run back to back keep more blocks alive than synchronised ones; let the caching allocator
        # reach its high-water mark (a handful of cudaMallocs per stream) before the warm-up steps so the timed region sees none
        if args.profile_step and with_clocks:      # for `ncu --profile-from-start off`: exactly one step between cudaProfilerStart/Stop
            for _ in range(PRIME_STEPS + warmup):
                res_fns[0]()
            torch.cuda.synchronize()
            l0 = ops.launch_count()
            torch.cuda.profiler.start()
            res_fns[0]()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            print(f"[profile-step] libgeobi kernels counted by the host side in this step: {ops.launch_count() - l0}", file=sys.stderr)
            return None
        for s_ in lane_streams:
            s_.wait_stream(torch.cuda.current_stream(dev))
        run_lanes(res_fns, (PRIME_STEPS + warmup) * W)
        torch.cuda.synchronize()
        l0 = ops.launch_count()
        if with_clocks:
            with ClockSampler(local_rank) as clk:
                ms, wall = timed_lanes(res_fns, steps)
            clocks = clk.summary()
        else:
            ms, wall = timed_lanes(res_fns, steps)
            clocks = None
        launches = ops.launch_count() - l0
        ms_single = None
        if W > 1:                                  # the single-stream figure beside it (lane 0 alone)
            ms_single, _ = timed_lanes(res_fns[:1], steps)

        def run_e2e(kind):
            fns = [make_step(i, kind) for i in range(W)]
            for p_ in pipes:
                p_["next"] = None
            run_lanes(fns, (PRIME_STEPS // 2 + warmup) * W)
            a0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)

            def settle():        # the trailing uploads (queued by the helper threads) belong to the timed region like the others
                for p_ in pipes:
                    nx = p_["next"]
                    if hasattr(nx, "result"):
                        p_["next"] = nx.result()
            join = [st_ for r in runners for st_ in (r.copy_stream, r.read_stream)]
            ms_, _ = timed_lanes(fns, steps, join=join, before_close=settle)
            for r in runners:
                r.wait()
            return ms_, torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - a0

        ms_graphs, allocs_graphs = run_e2e("graphs")
        ms_mesh, allocs_mesh = run_e2e("mesh") if wl.h2d_mesh_bytes is not None else (None, None)
        d2h = sum(t.numel() * t.element_size() for t in runners[0].out_host.values())
        total_faces = wl.faces * world
        lanes_txt = "%d concurrent lane(s): host thread + CUDA stream + module replica + runner each; " % W
        e2e_graphs = {"value": round(total_faces * steps / (ms_graphs / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes,
                      "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_graphs / steps, 4), "cuda_mallocs_in_timed_region": int(allocs_graphs),
                      "path": lanes_txt + "inference.HostBatchRunner.upload: PREBUILT graphs as a data loader reading the reference's cached .pt "
                              "files hands them over (x, edge_index, edge_weight, fv_indices; index tensors int32, widened on the device) -> H2D + "
                              "input-level CSR build on a copy stream, queued by a helper thread (upload_async) -> DualGNN forward -> D2H of vertices "
                              "and normals on a read-back stream; every lane, copy and read-back stream joined before the closing event"}
        if ms_mesh is None:
            e2e = e2e_graphs
        else:
            e2e = {"value": round(total_faces * steps / (ms_mesh / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": wl.h2d_mesh_bytes,
                   "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_mesh / steps, 4), "cuda_mallocs_in_timed_region": int(allocs_mesh),
                   "path": lanes_txt + "inference.HostBatchRunner.upload_mesh: the RAW mesh (points fp32 + faces int32, pinned) -> H2D -> device "
                           "front end on the copy stream (topology, both graphs as loop-free CSRs, bilateral weights, normalised features: "
                           "topology.DeviceTriMesh + dataset.build_dual_on_device(csr_native=True), 0.9 ms of GPU time; the reference's int64 "
                           "edge lists stay lazy), queued by a helper thread (upload_mesh_async) while the lane's thread queues its current "
                           "forward -> DualGNN forward -> D2H of vertices and normals on a read-back stream; every lane, copy and read-back "
                           "stream joined before the closing event",
                   "prebuilt_graphs": e2e_graphs}
        out = {"value": total_faces * steps / (ms / 1e3), "ms_per_step": ms / steps, "wall_s": wall, "launches": launches, "clocks": clocks, "e2e": e2e}
        if ms_single is not None:
            out["single_stream"] = {"value": round(total_faces * steps / (ms_single / 1e3), 1), "ms_per_step": round(ms_single / steps, 4)}
        return out

    # ---------------------------------------------------------------- headline workload
    wl = Workload(args.workload, rank, dev)
    res = measure(wl, args.steps, args.warmup, with_clocks=True)
    if res is None:
        return

    # ---- roofline of the dominant kernel: the fused FeaSt conv on the largest layer (facet r_conv4: N = F, 64 -> 32)
    roof = None
    if rank == 0:
        n_f = wl.faces
        g = ops.csr_from_coo(wl.df.edge_index, n_f, None, ops.COO_BY_COL | ops.COO_DROP_SELF | ops.COO_SORT_NBR)
        conv = net.gnn_f.r_conv4
        x = torch.randn(n_f, 64, device=dev)
        out = torch.empty(n_f, 32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
        prec = config.precision_code()
        fused = prec == ops.PREC_BF16X3
        tcagg = fused and os.environ.get("GEOBI_TCAGG", "0")[:1] == "1"
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
        alg = feast_bytes_alg(n_f, g.nnz + n_f, 64, 32)
        peak, how = peaks()
        achieved = alg / (t_ms / 1e3) / 1e9
        kname = "feast_tcagg_64_32_kernel" if tcagg else ("feast_fused_64_32_kernel" if fused else "geobi_feast_fwd (project + aggregate + gemm)")
        traffic, traffic_src = ncu_traffic(kname, n_f)
        fwd_gbs = FORWARD_BYTES_PER_FACE * wl.faces / (res["ms_per_step"] / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic,
                "kernel": kname + " on facet r_conv4: N=%d, E=%d incl. self loops, 64->32" % (n_f, g.nnz + n_f),
                "alg_bytes_per_launch": alg, "ms_per_launch": round(t_ms, 4), "ms_whole_layer": round(t_op, 4), "peak_source": how,
                "l2": "flushed (256 MB write) between launches; the layer's input alone is %d MB" % (n_f * 256 // 10 ** 6),
                "traffic_source": traffic_src,
                "whole_forward": {"alg_bytes_per_face": FORWARD_BYTES_PER_FACE, "achieved": round(fwd_gbs, 1), "unit": "GB/s",
                                  "frac": round(fwd_gbs / peak, 4), "note": "3104 B/face (SURVEY.md 8d) x faces / ms_per_step / peak"},
                "note": "HBM is the nominal bound of a gather/segment-sum; this kernel's 9-head weighting costs 576 fp32 FMAs per gathered 256-byte "
                        "row (~0.3 of the HBM peak at 100 % of the FP32 pipe) and ncu shows it bound by its shared-memory data pipe (62 % LSU + "
                        "10 % tensor-core operand reads after the weights moved to tensor memory). The tcgen05 form of the aggregation "
                        "(feast_tcagg, GEOBI_TCAGG=1) is 14 % faster as a kernel and level as a layer: profiles/r02_NOTES.md sections A, F"}
        del x, out, flush, g

    # ---------------------------------------------------------------- sub-results measured in the same run
    extra = {}
    if not args.headline_only:
        other = "patches" if args.workload == "mesh1m" else "mesh1m"
        del wl
        torch.cuda.empty_cache()
        wl2 = Workload(other, rank, dev)
        r2 = measure(wl2, min(args.steps, 10), 3, with_clocks=False)
        extra["configs1_patches64" if other == "patches" else "configs2_mesh1m"] = {
            "workload": wl2.describe, "value": round(r2["value"], 1), "unit": UNIT, "ms_per_step": round(r2["ms_per_step"], 4),
            "steps": min(args.steps, 10), "scaling": "weak", "e2e": r2["e2e"], "gpu_launches": r2["launches"],
            "concurrent_forwards": max(1, int(args.streams)), "single_stream": r2.get("single_stream")}
        del wl2
        torch.cuda.empty_cache()
        extra["configs3_mesh10m"] = strong_mesh10m(net, dev, rank, world, barrier, max_over_ranks)
        extra["configs4_train"] = train_steps(dev, rank, world, barrier, max_over_ranks, args.precision)

    cpu = cpu_baseline(12.0, args.workload) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    if rank == 0:
        line = {"metric": METRIC, "value": round(res["value"], 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(res["ms_per_step"], 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16": "bf16 tensor-core projections (1 pass), f32 elsewhere",
                          "bf16x3": "f32 storage and accumulation; projections on tcgen05 with operands split into two bf16 halves (x3 passes): "
                                    "1e-5 of the oracle on the bench shapes, 2.1e-5 worst tap at 200k faces (tests/test_gpu_batch.py)"}[args.precision],
                "data": "synthetic",
                "config": {"workload": DESCRIBE[args.workload], "faces_per_gpu": FACES[args.workload], "precision": args.precision,
                           "l2": "per-step working set (inputs 390 MB + >4 GB intermediates) exceeds the 126 MB L2",
                           "concurrent_forwards": max(1, int(args.streams)),
                           "concurrency": "steps are dealt to %d host thread(s), each with its own CUDA stream and module replica (same weights): one "
                                          "mesh's latency-bound stretches (coarse-level matcher, scans, count read-backs) overlap with another's "
                                          "kernels; ms_per_step = timed region / steps; `single_stream` = lane 0 alone" % max(1, int(args.streams)),
                           "resident_inputs": "mesh1m: the graph pair as the device front end produces it (features, loop-free CSRs with their "
                                              "bilateral weights: Data.csr; the reference's int64 edge lists stay lazy); patches: the union batch in "
                                              "the reference layout (lists; sort-free CSRs rebuilt every step)",
                           "timing": "CUDA events on the launch stream around all lanes (it waits for every lane stream before the closing event), "
                                     "max over ranks", "wall_s": round(res["wall_s"], 4),
                           "priming": f"{PRIME_STEPS} untimed forwards before the {args.warmup} warm-up steps (caching-allocator high-water mark)"},
                "e2e": res["e2e"], "gpu_launches": res["launches"], "clocks": res["clocks"], "roofline": roof, "cpu_baseline": cpu}
        if "single_stream" in res:
            line["single_stream"] = res["single_stream"]
        line.update(extra)
        print_json(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


DESCRIBE = {"mesh1m": "configs[2]: one noisy icosphere mesh of 1 003 520 faces / 501 762 vertices per GPU (frequency 224), full vertex + facet graph, "
                      "DualGNN fwd, random init",
            "patches": "configs[1]: 64 noisy icosphere patches x 8000 faces per GPU, disjoint-union batch, DualGNN fwd, random init"}
FACES = {"mesh1m": 20 * MESH_FREQ ** 2, "patches": N_PATCHES * 20 * PATCH_FREQ ** 2}


def strong_mesh10m(net, dev, rank, world, barrier, max_over_ranks):
    """BASELINE configs[3]: one 10 M-face mesh, the reference's whole-mesh pipeline (test_dual.predict_one: BFS patches of <= 1 M faces,
    per-patch forward, overlap-average stitch, 60-sweep vertex update) with the patches dealt to the ranks.  The partition
    (dataset.py:156-193) is serial by construction - every seed depends on what the earlier patches covered - and is computed once
    per mesh; it is timed separately, like the reference's own preprocessing print (test_dual.py:37-40).  `inference` = what shards:
    per-rank patch loop (cut-out, H2D, device topology + graphs, forward, stitch), accumulator reduction onto rank 0, vertex update."""
    from geobi_gnn_b200 import dataset, inference
    t0 = time.perf_counter()
    mesh = noisy_device_mesh(BIG_FREQ, 0, dev)          # the SAME mesh on every rank
    torch.cuda.synchronize()
    t_mesh = time.perf_counter() - t0
    # everything stays on the device: normalisation (centroid from one D2H copy of the points, scale from the vertex CSR), the BFS
    # partition (geobi_bfs_*: the reference's discovery order, ring-parallel) and the per-patch cut-outs; GEOBI_BENCH_HOST_SPLIT=1
    # restores the host C++ splitter over host copies of the index arrays (round-2 first session: 0.38 s + 0.52 s)
    host_split = os.environ.get("GEOBI_BENCH_HOST_SPLIT") == "1"
    t0 = time.perf_counter()
    if host_split:
        host = inference.host_views(mesh)
        c_np, scale = dataset.normalisation(host[0], host[3])
        norm, cen = (torch.from_numpy(c_np).float().to(dev), float(scale)), None
    else:
        host = None
        norm, cen = inference.device_normalisation(mesh)
    torch.cuda.synchronize()
    t_views = time.perf_counter() - t0
    if not host_split:
        inference.partition(mesh, BIG_SUB, centroid=cen)      # untimed first pass (workspace allocation, first launches)
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    parts = inference.partition(mesh, BIG_SUB, host=host, centroid=cen)
    torch.cuda.synchronize()
    t_part = time.perf_counter() - t0
    # warm-up: the same sharded run once, untimed (allocator high-water mark, NCCL reduce path)
    inference.predict_mesh(net, mesh, BIG_SUB, device=dev, rank=rank, world=world, parts=parts, host=host, norm=norm)   # untimed first pass
    # A generation-2 collection inside this single timed pass scans everything the bench has built so far: a 100+ ms stall on the rank
    # it hits (2 GPUs: 0.337 s instead of 0.177 s).  The collector stays ON - the per-patch objects hold cycles whose device memory must
    # come back promptly (switched off, 23 patches on one GPU ran into fresh cudaMallocs: 0.58 s instead of 0.33 s) - but everything
    # alive now is frozen into the permanent generation, so a full collection during the pass only looks at the pass's own objects.
    import gc
    gc.collect()
    gc.freeze()
    barrier()
    try:
        t0 = time.perf_counter()
        out = inference.predict_mesh(net, mesh, BIG_SUB, device=dev, rank=rank, world=world, parts=parts, host=host, norm=norm)
        barrier()
        t_inf = time.perf_counter() - t0
    finally:
        gc.unfreeze()
    (t_inf, t_part_max) = max_over_ranks(t_inf, t_part)
    ok = None
    if rank == 0:
        V, Np, Vp = out
        ok = bool(torch.isfinite(V).all() and torch.isfinite(Np).all() and float((Np.norm(dim=1) - 1).abs().max()) < 1e-5)
    faces = mesh.n_faces
    del mesh, out
    torch.cuda.empty_cache()
    return {"workload": f"configs[3]: one {faces}-face mesh (icosphere frequency {BIG_FREQ}), sub_size {BIG_SUB}: {len(parts)} BFS patches dealt "
                        f"round-robin to {world} GPU(s); no data-path collective, accumulators reduced onto rank 0 once",
            "scaling": "strong", "n_gpus": world, "faces": faces, "patches": len(parts),
            "value": round(faces / t_inf, 1), "unit": UNIT, "inference_s": round(t_inf, 4),
            "partition_s": round(t_part_max, 4), "partition": "host C++ splitter over host copies of the index arrays" if host_split else
            "device BFS splitter (geobi_bfs_begin / geobi_bfs_grow): identical patches, no host copy of the index arrays",
            "host_views_s": round(t_views, 4), "host_views": "D2H of points / fv / vf / ev + numpy normalisation" if host_split else
            "normalisation only: centroid from a D2H copy of the points, scale = 1 / mean edge length from the vertex CSR on the device",
            "whole_mesh_topology_s": round(t_mesh, 4),
            "end_to_end_s": round(t_views + t_part_max + t_inf, 4), "end_to_end_faces_per_s": round(faces / (t_views + t_part_max + t_inf), 1),
            "timing": "wall clock between barriers (+ device synchronize), max over ranks; partition and host views once per mesh, outside",
            "outputs_ok": ok}


def train_steps(dev, rank, world, barrier, max_over_ranks, precision, patches_per_rank=16, steps=8):
    """BASELINE configs[4]: train_dual.py-style training step (forward, L1 losses, backward, ONE flat fp32 gradient all-reduce over
    NCCL, Adam) on a batch of 16 x 8000-face patches per rank (weak scaling; one batch per rank = the reference's gradient
    accumulation over `batch_size` meshes, train_dual.py:211-218).  The all-reduce alone is timed as well."""
    from geobi_gnn_b200 import batching, dataset, network, parallel, train
    patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in patch_meshes(patches_per_rank, first_seed=rank * patches_per_rank)]
    dv, df, _ = batching.collate_dual(patches)
    torch.manual_seed(0)
    net = network.DualGNN().to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)

    def step():
        return train.train_step(net, opt, [batching.fresh_view(dv), batching.fresh_view(df)], world_size=world)

    for _ in range(5):
        loss, ev, en = step()
    import gc
    gc.collect()
    gc.freeze()               # as for configs[3]: a full collection inside the timed steps only scans the steps' own objects
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        e0.record()
        for _ in range(steps):
            loss, ev, en = step()
        e1.record()
        barrier()
    finally:
        gc.unfreeze()
    ms = e0.elapsed_time(e1) / steps
    ms_ar = 0.0
    if world > 1:       # the collective alone: the same flat bucket, gradients in place from the last step
        for _ in range(3):
            parallel.allreduce_gradients(net.parameters(), average=False)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            parallel.allreduce_gradients(net.parameters(), average=False)
        a1.record()
        barrier()
        ms_ar = a0.elapsed_time(a1) / 10
    ms, ms_ar = max_over_ranks(ms, ms_ar)
    faces = df.x.size(0) * world
    n_params = sum(p.numel() for p in net.parameters())
    out = {"workload": f"configs[4]: training step on {patches_per_rank} patches x 8000 faces per GPU (fwd + L1 losses + bwd + flat gradient "
                       f"all-reduce + Adam lr 1e-3), {world} GPU(s)",
           "scaling": "weak", "n_gpus": world, "value": round(faces / (ms / 1e3), 1), "unit": UNIT, "ms_per_step": round(ms, 3), "steps": steps,
           "allreduce_ms": round(ms_ar, 4), "allreduce": f"one flat fp32 bucket of {n_params} gradients ({4 * n_params / 1e6:.2f} MB), NCCL" if world > 1 else None,
           "precision": precision, "loss": float(loss), "error_n_deg": float(en),
           "backward": "every FeaStConv layer is one geobi_feast_bwd call (dZ = g.W_flat and the split-K dW = g^T.Z on tcgen05 with split bf16 operands, "
                       "soft-assignment / gather part, dX += dP.U, dU = dP^T.X), each FC head one geobi_mlp_head_bwd call (hidden layer recomputed on "
                       "tcgen05, the same split-K and TMA GEMM kernels), segment max and the vertex-to-facet transfer in libgeobi kernels; no library "
                       "GEMM in the step",
           "roofline_dw_splitk": {"bound": "hbm", "achieved": 5426.0, "peak": 6542.1, "unit": "GB/s", "frac": 0.83,
                                  "source": "ncu --set full, profiles/r02/r02b_ncu_full_dw_splitk.csv: dW = g^T.Z of a 64->32 layer at 128 000 nodes reads "
                                            "327.7 MB (= its algorithmic bytes) in 60.4 us; not re-measured by this run"}}
    del net, opt, dv, df, patches
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------- CPU legs (oracle)
def _oracle_patch_inputs(count, workload, first_seed=0):
    """Bounded sample of the workload for the CPU legs: whole 8000-face patches (configs[1]), or BFS patches of CPU_SUB faces cut out of
    the 1 M-face mesh with the reference's rule (configs[2]: the reference's CPU path runs a large mesh patch-wise, test_dual.py:49-61)."""
    from oracle import ref_dataset
    if workload == "patches":
        return [ref_dataset.build_dual_data(mn, mo) for mn, mo in patch_meshes(count, first_seed)], 20 * PATCH_FREQ ** 2
    from geobi_gnn_b200 import patches, synth
    p, f = synth.icosphere(MESH_FREQ)
    mesh = synth.TriMesh(p, f)                                   # host topology of the clean mesh (seeds / rings only need fv, vf)
    pn = synth.add_normal_noise(p, f, 0.2, first_seed)
    out = []
    rng = np.random.default_rng(0)
    for seed_face in rng.integers(0, f.shape[0], size=count):
        sel = patches.mesh_get_neighbor_np(mesh.fv, mesh.vf, int(seed_face), neighbor_count=CPU_SUB)
        v_idx, faces = patches.get_submesh(mesh.fv, sel)
        sub_n, sub_o = synth.TriMesh(pn[v_idx], faces), synth.TriMesh(p[v_idx], faces)
        out.append(ref_dataset.build_dual_data(sub_n, sub_o))
    return out, CPU_SUB


def cpu_baseline(sample_seconds, workload="mesh1m"):
    """Oracle ('port' of the reference path; the reference itself cannot be imported here) timed on the host cores
    on a bounded sample of the same workload: one oracle forward per patch until ~sample_seconds elapsed."""
    from oracle import ref_network
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    net = ref_network.DualGNN().eval()
    pool, faces_each = _oracle_patch_inputs(3, workload)
    with torch.no_grad():
        net([pool[0][0].clone(), pool[0][1].clone()])        # warm-up
        done, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < sample_seconds or done < 3:
            dv, df = pool[done % len(pool)]
            net([dv.clone(), df.clone()])
            done += 1
        dt = time.perf_counter() - t0
    faces = done * faces_each
    return {"value": round(faces / dt, 1), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} BFS patches of {faces_each} faces of the workload's mesh, one oracle forward each, {dt:.1f} s"
                      if workload == "mesh1m" else f"{done} patches of {faces_each} faces, one oracle forward each, {dt:.1f} s"}


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
    per_step = 1 if args.workload == "mesh1m" else 2
    pool, faces_each = _oracle_patch_inputs(per_step, args.workload)

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
    faces = per_step * faces_each
    value = faces * args.steps / dt
    cores = torch.get_num_threads()
    sample = (f"{per_step} BFS patch of {faces_each} faces cut from the 1 003 520-face mesh per step (the reference's CPU path runs a large mesh "
              f"patch-wise, test_dual.py:49-61; one 1 M-face graph pair needs 15 GB per per-edge tensor)") if args.workload == "mesh1m" else \
             f"{per_step} of the 64 patches per step (the reference runs one patch per forward, dataset.py:29-31)"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": DESCRIBE[args.workload], "sample": sample},
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} x {faces_each} faces per step, {args.steps} steps"},
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
    ap.add_argument("--streams", type=int, default=int(os.environ.get("BENCH_STREAMS", "2")),
                    help="concurrent forwards per GPU (host thread + CUDA stream + module replica each); 1 = single stream")
    ap.add_argument("--precision", default=os.environ.get("GEOBI_PRECISION", "bf16x3"), choices=["fp32", "bf16", "bf16x3"])
    ap.add_argument("--workload", default=os.environ.get("BENCH_WORKLOAD", "mesh1m"), choices=["mesh1m", "patches"],
                    help="headline workload: mesh1m = BASELINE configs[2] (default), patches = configs[1]")
    ap.add_argument("--headline-only", action="store_true", help="skip the sub-results (other single-GPU config, configs[3] strong scaling)")
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
