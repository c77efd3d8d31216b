"""pytest configuration: the `gpu` marker, import path, and shared stream generators."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing
    them (the product has no CPU path: smatrix_open returns NULL there)."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Compile oracle/liboracle.so + libsmxdriver.so (and oracle/_ref when the reference
    checkout is present).  gcc only; a few seconds."""
    from oracle import cpu
    cpu.build(ref=True)


def safe_stream(rng: np.random.Generator, n: int, n_rows: int, n_cols: int, op: str,
                col0_rate: float = 0.1, wide_keys: bool = False, max_val: int = 2**32 - 1):
    """A random op stream inside the reference's safe domain: once non-zero,
    column 0 of a row never returns to 0.  For incr/decr that means column-0 deltas are small
    positive numbers that cannot wrap; for set, column-0 values are always non-zero."""
    if wide_keys:
        row_ids = rng.integers(0, 2**32, size=n_rows, dtype=np.uint64).astype(np.uint32)
        col_ids = rng.integers(1, 2**32, size=n_cols, dtype=np.uint64).astype(np.uint32)
    else:
        row_ids = np.arange(n_rows, dtype=np.uint32)
        col_ids = np.arange(1, n_cols + 1, dtype=np.uint32)
    xs = row_ids[rng.integers(0, n_rows, size=n)]
    ys = col_ids[rng.integers(0, n_cols, size=n)]
    vs = rng.integers(0, max_val + 1, size=n, dtype=np.uint64).astype(np.uint32)
    is0 = rng.random(n) < col0_rate
    ys = np.where(is0, np.uint32(0), ys).astype(np.uint32)
    if op == "set":
        vs = np.where(is0, np.maximum(vs, 1), vs).astype(np.uint32)
    else:  # incr / decr on column 0: bounded so partial sums never come back to 0 mod 2^32
        vs = np.where(is0, (vs % 1000).astype(np.uint32), vs).astype(np.uint32)
    return xs, ys, vs
