"""Throughput of W concurrent forwards (W host threads, one CUDA stream and one module replica each, inputs resident) against one:
do the latency-bound stretches of a forward (matcher on the coarse levels, scans, count read-backs) overlap with another mesh's kernels?"""
import copy, os, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, config, dataset, network
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
config.set_precision("bf16x3")
mesh = bench.noisy_device_mesh(bench.MESH_FREQ, 0, dev)
dv, df = dataset.build_dual_on_device(mesh, None, csr_native=True)
torch.manual_seed(0)
net0 = network.DualGNN().to(dev).eval()
STEPS = int(os.environ.get("STEPS", 12))

def worker(net, stream, n, out):
    torch.cuda.set_device(dev)
    with torch.cuda.stream(stream), torch.no_grad():
        for _ in range(n):
            v, nrm, _ = net([batching.fresh_view(dv), batching.fresh_view(df)])
        out.append(float(v.sum()))
    stream.synchronize()

for W in (1, 2, 3):
    nets = [net0] + [copy.deepcopy(net0) for _ in range(W - 1)]
    streams = [torch.cuda.Stream(dev) for _ in range(W)]
    for rep in range(2):          # first repetition warms allocator pools of every stream
        outs = []
        ths = [threading.Thread(target=worker, args=(nets[i], streams[i], STEPS, outs)) for i in range(W)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"W = {W}: {W * STEPS} forwards in {dt * 1e3:.1f} ms = {dt * 1e3 / (W * STEPS):.2f} ms per mesh, {W * STEPS * mesh.n_faces / dt / 1e6:.1f} M faces/s; checks {outs[:W]}")
