"""geobi_gnn_b200 — B200-native GeoBi-GNN dual-domain forward.

Host side mirrors the reference's module surface (`network`, `net_util`,
`data_util`, `dataset`); compute runs in hand-written sm_100a kernels behind the
C ABI of `libgeobi.so` (include/geobi.h).  There is no CPU fallback: any op
raises if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
