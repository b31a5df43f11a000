"""One bench-shaped forward under torch.profiler (CUPTI): warm kernel durations, GPU busy vs idle."""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, config, dataset, network
config.set_precision(os.environ.get("GEOBI_PRECISION", "bf16x3"))
dev = torch.device("cuda")
if os.environ.get("WORKLOAD", "mesh1m") == "patches":
    patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(bench.N_PATCHES, 0)]
    dv, df, _ = batching.collate_dual(patches)
else:
    dv, df = dataset.build_dual_on_device(bench.noisy_device_mesh(bench.MESH_FREQ, 0, dev), None)
torch.manual_seed(0)
net = network.DualGNN().to(dev).eval()
def step():
    with torch.no_grad():
        return net([batching.fresh_view(dv), batching.fresh_view(df)])
for _ in range(12): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
busy = sum(e - s for s, e, _ in ks)
span = ks[-1][1] - ks[0][0]
print(f"wall {wall*1e3:.2f} ms; GPU span {span/1e3:.2f} ms; GPU busy {busy/1e3:.2f} ms; idle {100*(1-busy/span):.1f}%; kernels {len(ks)}")
import collections, re
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    n = re.sub(r"\(.*", "", n)[:60]; agg[n][0] += 1; agg[n][1] += e - s
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{t:9.1f} us {100*t/busy:5.1f}% n={c:4d} {n}")
# biggest idle gaps
gaps = sorted(((ks[i+1][0] - ks[i][1], ks[i][2][:40], ks[i+1][2][:40]) for i in range(len(ks)-1)), reverse=True)[:12]
print("largest gaps (us): ", [(round(g,1), a, b) for g, a, b in gaps])
if os.environ.get("TIMELINE_DUMP"):
    with open(os.environ["TIMELINE_DUMP"], "w") as f:
        prev = ks[0][0]
        for s, e, n in ks:
            f.write(f"{s - ks[0][0]:10.1f} gap={s - prev:7.1f} dur={e - s:7.1f} {re.sub(r'\(.*', '', n)[:70]}\n")
            prev = e
