"""One configs[4]-shaped training step (16 x 8000-face patches) under torch.profiler: kernel totals, GPU busy vs idle."""
import collections, os, re, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, config, dataset, network, train
config.set_precision(os.environ.get("GEOBI_PRECISION", "bf16x3"))
dev = torch.device("cuda")
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(16, 0)]
dv, df, _ = batching.collate_dual(patches)
torch.manual_seed(0)
net = network.DualGNN().to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
def step():
    return train.train_step(net, opt, [batching.fresh_view(dv), batching.fresh_view(df)], world_size=1)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print(f"train step {1e3 * (time.perf_counter() - t0) / 5:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
busy = sum(e - s for s, e, _ in ks)
span = ks[-1][1] - ks[0][0]
print(f"wall {wall*1e3:.2f} ms; GPU span {span/1e3:.2f} ms; GPU busy {busy/1e3:.2f} ms; idle {100*(1-busy/span):.1f}%; kernels {len(ks)}")
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    n = re.sub(r"\(.*", "", n)[:70]; agg[n][0] += 1; agg[n][1] += e - s
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{t:9.1f} us {100*t/busy:5.1f}% n={c:4d} {n}")
