"""CPU oracle for the GeoBi-GNN dual-domain forward.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (plain PyTorch fp32 + one small C file) of
the reference path ``network.DualGNN.forward`` and everything under it
(``/root/reference/code/network.py:254-413``, ``net_util.py:56-302``,
``data_util.py:182-230,383-556``, ``dataset.py:196-269``) plus the third-party
semantics that path reaches (torch_geometric FeaStConv / utils, torch_scatter,
torch_sparse.coalesce, torch_cluster.graclus).  None of those third-party
packages is vendored in the reference, none is pinned (no requirements file),
and none is installable here, so their behaviour is restated from their
published algorithms; see SURVEY.md section 8(c).

**Parity status: the reference's own files are PINNED, the third-party operators are
UNPINNED.**  The reference has no tests, golden vectors or fixtures (SURVEY.md
section 4).  ``tests/golden/make_reference_golden.py`` therefore EXECUTES the
unmodified reference modules (``dataset``, ``data_util``, ``net_util``, ``network``,
``test_dual``) in the build container, with stand-ins only for the third-party
packages they import, and stores what they produce: graph / feature assembly, a
DualGNN forward on two configurations (Synthetic and Kinect ``force_depth``), losses
and errors, ``update_position2``, the BFS patch walk, a whole ``predict_one`` over
11 patches and the parameter gradients of one training micro-step
(``tests/golden/reference_*.npz``).  ``tests/test_reference_golden.py`` checks this
oracle against those vectors (bit-equal weights and index arrays, floats to 1e-6 /
1e-5) and, under ``-m gpu``, the CUDA path.  What stays unpinned: FeaStConv,
graclus, scatter, coalesce and the OpenMesh index arrays themselves - both sides
take them from ``oracle/pyg.py`` / ``synth.TriMesh``, restated from the published
behaviour.  ``tests/golden/dualgnn_ico3.npz`` (``make_golden.py``) is the older
self-generated vector set that guards against drift.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``geobi_gnn_b200``) never does.
"""
