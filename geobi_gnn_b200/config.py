"""Process-wide numeric mode of the projection GEMMs (FeaSt `lin`, FC heads).

'fp32' : everything on CUDA cores in fp32 — parity target 1e-5 (BASELINE.json north_star).
'bf16' : projections on tcgen05 tensor cores, bf16 operands / fp32 accumulate, one pass (fastest; ~2.5e-3 max-norm).
'bf16x3': same tensor-core kernels with operands split hi + lo (both bf16), three passes — fp32-grade (~1e-6).
Aggregation, soft assignments, pooling and all integer work are identical in both modes.
"""
_PRECISION = "fp32"


def set_precision(mode: str):
    global _PRECISION
    if mode not in ("fp32", "bf16", "bf16x3"):
        raise ValueError("precision must be 'fp32', 'bf16' or 'bf16x3'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def precision_code() -> int:
    return {"fp32": 0, "bf16": 1, "bf16x3": 2}[_PRECISION]
