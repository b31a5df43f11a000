import sys, time, torch, cProfile, pstats
sys.path.insert(0, '/root/repo')
from geobi_gnn_b200 import synth, dataset, network, batching, config
config.set_precision('bf16x3')
dev = torch.device('cuda')
p, f = synth.icosphere(4)
m = synth.TriMesh(synth.add_normal_noise(p, f), f)
dv, df = dataset.build_dual_data(m, synth.TriMesh(p, f), device=dev)
torch.manual_seed(0)
net = network.DualGNN().to(dev).eval()
def step():
    with torch.no_grad():
        return net([batching.fresh_view(dv), batching.fresh_view(df)])
for _ in range(5): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f'tiny forward (320 faces): {dt*1e3:.2f} ms per step = host overhead floor')
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize(); pr.disable()
ps = pstats.Stats(pr); ps.sort_stats('tottime').print_stats(18)
