"""cProfile of the host side of the configs[4]-shaped training step (what keeps the GPU idle 31 % of the step)."""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geobi_gnn_b200 import batching, config, dataset, network, train
config.set_precision("bf16x3")
dev = torch.device("cuda")
patches = [dataset.build_dual_data(mn, mo, device=dev) for mn, mo in bench.patch_meshes(16, 0)]
dv, df, _ = batching.collate_dual(patches)
torch.manual_seed(0)
net = network.DualGNN().to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
def step():
    return train.train_step(net, opt, [batching.fresh_view(dv), batching.fresh_view(df)], world_size=1)
for _ in range(6): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
print(f"{1e3 * (time.perf_counter() - t0) / 20:.2f} ms per step under cProfile")
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(30)
