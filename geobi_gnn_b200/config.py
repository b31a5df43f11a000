"""Process-wide numeric mode of the projection GEMMs (FeaSt `lin`, FC heads).

'fp32' : everything on CUDA cores in fp32 — parity target 1e-5 (BASELINE.json north_star).
'bf16' : projections on tcgen05 tensor cores, bf16 operands / fp32 accumulate — parity target 2e-3.
Aggregation, soft assignments, pooling and all integer work are identical in both modes.
"""
_PRECISION = "fp32"


def set_precision(mode: str):
    global _PRECISION
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def precision_code() -> int:
    return 0 if _PRECISION == "fp32" else 1
