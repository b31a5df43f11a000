"""topology arrays built on the device (topology.DeviceTriMesh) against the numpy stand-in for
OpenMesh (synth.TriMesh), and the whole-mesh pipeline with either front end."""
import numpy as np
import pytest
import torch

from tests import util

DEV = "cuda"


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 4, 9])
def test_device_trimesh_matches_host_trimesh(n):
    from geobi_gnn_b200 import synth, topology
    p, f = synth.icosphere(n)
    pn = synth.add_normal_noise(p, f, 0.2, seed=n)
    host = synth.TriMesh(pn, f)
    dev = topology.DeviceTriMesh(pn, f, DEV)
    assert dev.n_vertices == host.n_vertices and dev.n_faces == host.n_faces
    assert np.array_equal(dev.fv.cpu().numpy(), host.fv)
    assert np.array_equal(dev.vf.cpu().numpy(), host.vf)                      # ascending incident faces, -1 padded
    assert np.array_equal(dev.vv.cpu().numpy(), host.vv)
    he = np.sort(host.ev, axis=1)
    he = he[np.lexsort((he[:, 1], he[:, 0]))]
    assert np.array_equal(dev.ev.cpu().numpy(), he)                           # same undirected edge set, canonical order
    assert np.abs(dev.face_normals.cpu().numpy() - host.face_normals).max() < 2e-6
    assert np.abs(dev.vertex_normals.cpu().numpy() - host.vertex_normals).max() < 2e-6
    assert abs(dev.mean_edge_length() - synth.mean_edge_length(host.points.astype(np.float32), host.ev)) < 1e-6


@pytest.mark.gpu
def test_process_one_submesh_same_graphs_from_either_front_end():
    from geobi_gnn_b200 import dataset, synth, topology
    p, f = synth.icosphere(6)
    pn = synth.add_normal_noise(p, f, 0.2, seed=2)
    gv_h, gf_h = dataset.process_one_submesh(synth.TriMesh(pn, f), "h", None, DEV)
    gv_d, gf_d = dataset.process_one_submesh(topology.DeviceTriMesh(pn, f, DEV), "d", None, DEV)
    for a, b in ((gv_h, gv_d), (gf_h, gf_d)):
        assert torch.equal(a.edge_index, b.edge_index)
        assert util.rel_err(b.edge_weight, a.edge_weight) < 1e-5
        assert util.rel_err(b.pos, a.pos) < 1e-6 and util.rel_err(b.normal, a.normal) < 1e-5
        assert a.coalesced_undirected and b.coalesced_undirected
    assert torch.equal(gf_h.fv_indices, gf_d.fv_indices)


@pytest.mark.gpu
def test_predict_mesh_device_topology_equals_host_topology():
    """Patch pipeline (split -> per-patch forward -> stitch -> update) with either topology front end: same result to fp32 noise."""
    from geobi_gnn_b200 import inference, network, synth
    torch.manual_seed(4)
    net = network.DualGNN().to(DEV).eval()
    for pl in util.poolings(net):
        pl.perm_fn = lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(n))
    p, f = synth.icosphere(8)
    mesh = synth.TriMesh(synth.add_normal_noise(p, f, 0.2, seed=5), f)
    out_h = inference.predict_mesh(net, mesh, 500, device=DEV, device_topology=False, return_parts=True)
    out_d = inference.predict_mesh(net, mesh, 500, device=DEV, device_topology=True, return_parts=True)
    assert out_h[3] == out_d[3] > 1                                            # several patches
    assert util.rel_err(out_d[2], out_h[2]) < util.TOL_FP32                             # network vertices
    assert float((out_d[1] - out_h[1]).abs().max()) < 5e-4                     # unit normals
    assert util.rel_err(out_d[0], out_h[0]) < util.TOL_FP32                             # updated vertices


@pytest.mark.gpu
@pytest.mark.parametrize("sub_size", [500, 100000])
def test_predict_mesh_accepts_device_built_whole_mesh(sub_size):
    """The WHOLE mesh handed over as a DeviceTriMesh (what meshio.denoise_obj does on a GPU) gives the result of the host
    TriMesh: same patches, same stitched vertices; sub_size 100000 exercises the single-graph branch (test_dual.py:44-47)."""
    from geobi_gnn_b200 import inference, network, synth, topology
    torch.manual_seed(4)
    net = network.DualGNN().to(DEV).eval()
    for pl in util.poolings(net):
        pl.perm_fn = lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(n))
    p, f = synth.icosphere(8)
    pn = synth.add_normal_noise(p, f, 0.2, seed=5)
    out_h = inference.predict_mesh(net, synth.TriMesh(pn, f), sub_size, device=DEV, return_parts=True)
    out_d = inference.predict_mesh(net, topology.DeviceTriMesh(pn, f, DEV), sub_size, device=DEV, return_parts=True)
    assert out_h[3] == out_d[3] and (out_d[3] > 1) == (sub_size == 500)
    assert util.rel_err(out_d[2], out_h[2]) < util.TOL_FP32
    assert float((out_d[1] - out_h[1]).abs().max()) < 5e-4
    assert util.rel_err(out_d[0], out_h[0]) < util.TOL_FP32
