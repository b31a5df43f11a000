"""BASELINE.json configs[3]: patch-sharded inference on one ~10 M-face synthetic mesh (icosphere n=708) across the ranks of
a torchrun launch, sub_size = 1 M faces, no collective on the data path (patches dealt round-robin; accumulators summed
onto rank 0 at the end).  Reports host preparation, per-rank patch forwards and the final update separately."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geobi_gnn_b200 import config, inference, network, patches, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 708
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
config.set_precision("bf16x3")
t0 = time.perf_counter()
from geobi_gnn_b200 import topology
p, f = synth.icosphere(n)
t_gen = time.perf_counter() - t0
# whole-mesh topology on the device (what meshio.denoise_obj does); HOST_TOPOLOGY=1 uses the numpy stand-in for OpenMesh
t0 = time.perf_counter()
if os.environ.get("HOST_TOPOLOGY"):
    mesh = synth.TriMesh(synth.add_normal_noise(p, f, 0.2, 0), f)
else:
    clean = topology.DeviceTriMesh(p, f, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    noise = torch.randn(clean.n_vertices, 1, generator=g, device=dev) * 0.2 * float((clean.points[clean.ev[:, 0]] - clean.points[clean.ev[:, 1]]).norm(dim=1).mean())
    mesh = topology.DeviceTriMesh(clean.points + noise * clean.vertex_normals, clean.fv, dev)
    del clean
torch.cuda.synchronize()
t_mesh = time.perf_counter() - t0
torch.manual_seed(0); net = network.DualGNN().to(dev).eval()
# warm-up on a small mesh (kernel loading, allocator)
ps, fs = synth.icosphere(40); inference.predict_mesh(net, synth.TriMesh(ps, fs), 10 ** 9, device=dev)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
phases = {} if os.environ.get("PHASES") else None       # PHASES=1 adds a sync per phase boundary (not per patch)
V, Np, Vp, n_patches = inference.predict_mesh(net, mesh, sub, device=dev, rank=rank, world=world, return_parts=True, timings=phases) if rank == 0 else \
    (inference.predict_mesh(net, mesh, sub, device=dev, rank=rank, world=world) + (None,))
torch.cuda.synchronize()
if world > 1: dist.barrier()
t_total = time.perf_counter() - t0
if rank == 0:
    ok = bool(torch.isfinite(V).all() and torch.isfinite(Np).all())
    print(json.dumps({"config": f"configs[3]: {mesh.n_faces}-face mesh, sub_size {sub}, {n_patches} patches over {world} GPU(s)",
                      "faces": mesh.n_faces, "n_gpus": world, "mesh_generation_s": round(t_gen, 1), "whole_mesh_topology_s": round(t_mesh, 2),
                      "whole_mesh_topology": "host numpy" if os.environ.get("HOST_TOPOLOGY") else "DeviceTriMesh",
                      "phases_s": None if phases is None else {k: round(v, 3) for k, v in phases.items()},
                      "predict_mesh_total_s": round(t_total, 2), "faces_per_s_end_to_end": round(mesh.n_faces / t_total, 1),
                      "finite": ok, "max_abs(|n|-1)": float((Np.norm(dim=1) - 1).abs().max())}))
if world > 1: dist.destroy_process_group()
