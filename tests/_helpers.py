"""Helpers shared by tests and tests/golden/make_golden.py."""
import numpy as np


def bf16_rne(x32: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32, by bit arithmetic (no NaN inputs)."""
    u = np.ascontiguousarray(x32, np.float32).view(np.uint32).astype(np.uint64)
    r = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)
    return (r << np.uint64(16)).astype(np.uint32).view(np.float32).reshape(np.shape(x32))


def stored_bf16_rows(X32: np.ndarray) -> np.ndarray:
    """What a bf16 table stores for fp32 input rows: RNE_bf16(fp32(x * (1/sqrt(canon |x|^2))))
    (csrc/table_ops.cu commit_rows_kernel), returned as fp32."""
    from oracle.cosine_topk import canon_sqnorm
    n2 = canon_sqnorm(X32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(n2 > 0, 1.0 / np.sqrt(n2), 0.0)
    y = (X32.astype(np.float64) * inv[:, None]).astype(np.float32)
    return bf16_rne(y)
