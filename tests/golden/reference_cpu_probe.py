"""Build container only (needs /root/reference): the reference's OWN DualGNN.forward (its modules executed over the third-party
stand-ins of tests/golden/make_reference_golden.py) timed next to the oracle port on the same 8000-face patch, same weights,
same threads.  Shows how representative bench.py's `cpu_baseline` (kind "port", the only one that can run on the GPU box) is of
the reference's CPU path.  TEST INFRASTRUCTURE (lives under tests/ because it executes oracle/).
python tests/golden/reference_cpu_probe.py [n_subdiv=20] [repeats=5] > profiles/r01_reference_vs_port_cpu.json
"""
import json
import os
import runpy
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    gold = runpy.run_path(os.path.join(ROOT, "tests", "golden", "make_reference_golden.py"))
    ref = gold["import_reference"]()
    from geobi_gnn_b200 import synth
    from oracle import ref_dataset, ref_network
    p, f = synth.icosphere(n)
    pn = synth.add_normal_noise(p, f, 0.2, seed=0).astype(np.float32)

    def reference_inputs():
        mesh_n, mesh_o = gold["_OMTriMesh"](pn, f), gold["_OMTriMesh"](p.astype(np.float32), f)
        _, centroid, scale = ref.data_util.center_and_scale(pn, mesh_n.ev_indices())
        dual = ref.dataset.DualDataset.process_one_submesh(mesh_n, "g", mesh_o)
        dual[0].centroid, dual[0].scale = torch.from_numpy(np.asarray(centroid)).float(), scale
        return ref.dataset.DualDataset.post_processing(dual, "Synthetic")

    def port_inputs():
        return ref_dataset.build_dual_data(synth.TriMesh(pn, f), synth.TriMesh(p.astype(np.float32), f))

    torch.manual_seed(0)
    net_ref = ref.network.DualGNN(force_depth=False, pool_type="max", wei_param=2).eval()
    torch.manual_seed(0)
    net_port = ref_network.DualGNN(force_depth=False, pool_type="max", wei_param=2).eval()

    def once(net, make_inputs):
        t0 = time.perf_counter()
        dv, df = make_inputs()
        t1 = time.perf_counter()
        with torch.no_grad():
            net([dv, df])
        return time.perf_counter() - t1, t1 - t0

    runs = {"ref": [], "port": []}
    for _ in range(reps + 1):                       # interleaved, first pair dropped (allocator / page-cache warm-up)
        runs["ref"].append(once(net_ref, reference_inputs))
        runs["port"].append(once(net_port, port_inputs))
    fwd_ref, build_ref = (float(np.median([r[k] for r in runs["ref"][1:]])) for k in (0, 1))
    fwd_port, build_port = (float(np.median([r[k] for r in runs["port"][1:]])) for k in (0, 1))
    print(json.dumps({"faces": int(f.shape[0]), "threads": torch.get_num_threads(), "repeats": reps,
                      "reference_forward_s": round(fwd_ref, 4), "port_forward_s": round(fwd_port, 4),
                      "reference_faces_per_s": round(f.shape[0] / fwd_ref, 1), "port_faces_per_s": round(f.shape[0] / fwd_port, 1),
                      "reference_graph_build_s": round(build_ref, 4), "port_graph_build_s": round(build_port, 4),
                      "note": "reference = /root/reference/code executed unmodified over oracle/pyg.py stand-ins for PyG / torch_scatter / "
                              "torch_sparse / torch_cluster (serial greedy graclus in C); port = oracle/ref_*.py"}))


if __name__ == "__main__":
    main()
