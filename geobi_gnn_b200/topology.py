"""Mesh -> topology arrays on the device.

The reference takes these arrays from OpenMesh (`dataset.py:196-210`: `ev_indices`, `fv_indices`, `vf_indices`,
`face_normals`, `vertex_normals`) for every mesh and every patch; with a numpy stand-in (synth.TriMesh) that host work
costs ~1.2 s per million faces and bounds whole-mesh inference (BASELINE config 3).  `DeviceTriMesh` derives the same
arrays from `points [V,3]` and `faces [F,3]` with the library's CSR kernels:

  * vertex graph = symmetrised, deduplicated, row-sorted adjacency of the 3F half-edges (`geobi_csr_from_coo`); `ev` is its
    upper triangle (unique undirected edges, sorted by (min, max) - OpenMesh's creation order is not reproduced, every
    consumer on the path is order-independent: `to_undirected` (`dataset.py:211`) and the mean edge length);
  * `vf` = faces grouped by vertex (`geobi_group_by`: ascending face ids, as OpenMesh's circulators give them for the
    builders' purposes), padded with -1 to the maximum valence (`data_util.py:436-456` wants the padded table);
  * face normals `normalize(cross(p1-p0, p2-p0))` (`geobi_face_normal`), vertex normals = normalised sum of the incident
    face normals (`geobi_segment_reduce`, op sum).

Float results are fp32 (synth.TriMesh computes fp64 and casts): they agree to ~1e-7 relative, integer arrays exactly.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops


class DeviceTriMesh:
    def __init__(self, points, faces, device="cuda"):
        dev = torch.device(device)
        self.points = torch.as_tensor(np.asarray(points, dtype=np.float32) if not torch.is_tensor(points) else points,
                                      dtype=torch.float32, device=dev).contiguous()
        self.fv = torch.as_tensor(np.asarray(faces, dtype=np.int64) if not torch.is_tensor(faces) else faces,
                                  dtype=torch.int64, device=dev).contiguous()
        V, Fc = self.points.size(0), self.fv.size(0)
        self.n_vertices, self.n_faces = V, Fc
        # --- faces around each vertex, ascending, padded to the maximum valence
        mrowptr, members = ops.group_by(self.fv.reshape(-1).to(torch.int32), V)
        self.vf_rowptr, self.vf_members = mrowptr, members.div(3, rounding_mode="floor").to(torch.int32)
        k = int((mrowptr[1:] - mrowptr[:-1]).max()) if V else 0
        self.vf = ops.pad_rows(mrowptr, members, k, 3)
        # --- vertex adjacency (no loops, rows sorted, symmetric): the other corners of each vertex's faces, deduplicated per row
        if 0 < k <= ops.RING_MAX_VALENCE and Fc:
            self.vertex_csr = ops.mesh_vertex_csr(self.fv, mrowptr, members, V)
        else:                          # very high valences (or an empty mesh): the general builder over the 3F half edges
            h = torch.stack([self.fv.reshape(-1), self.fv[:, [1, 2, 0]].reshape(-1)])
            self.vertex_csr = ops.csr_from_coo(h, V, None, ops.COO_SYMMETRIZE | ops.COO_SORT_NBR | ops.COO_DEDUP | ops.COO_DROP_SELF)
        self.vertex_csr.symmetric = True
        self._ev = None                # unique undirected edges: built on first read (the device front end itself never needs them)
        self.vf_sorted = True          # rows ascending, pads last: lets build_facet_graph merge instead of sort
        self.update_normals()

    def update_normals(self):
        self.face_normals = ops.face_normal(self.points, self.fv)
        vn = ops.segment_reduce(self.face_normals, self.vf_rowptr, self.vf_members, self.n_vertices, ops.OP_SUM)
        self.vertex_normals = F.normalize(vn, dim=1, eps=1e-30)

    @property
    def ev(self):
        """Unique undirected edges [E,2] (min, max), sorted: the upper triangle of the vertex CSR."""
        if self._ev is None:
            ei = self.vertex_csr.edge_index()
            self._ev = ei[:, ei[0] < ei[1]].t().contiguous()
        return self._ev

    @property
    def vv(self):
        """Neighbour table padded with -1 (only data_util.build_vertex_graph, dead code upstream, reads it)."""
        g = self.vertex_csr
        k = int((g.rowptr[1:] - g.rowptr[:-1]).max()) if self.n_vertices else 0
        return ops.pad_rows(g.rowptr, g.nbr, k, 1)

    def mean_edge_length(self, centroid=None) -> float:
        return float(ops.mean_edge_length_csr(self.points, self.vertex_csr))
