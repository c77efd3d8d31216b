"""Synthetic workload definitions of SURVEY.md 8(d) that host and device must agree on.

Only definitions (no table code): the inverse-CDF threshold tables the C3 / C4 generators draw from.
The generators themselves are device kernels (smatrix_b200_gen_c3_* / gen_c4_*, include/smatrix_b200.h);
the CPU baseline has its own C copies and must be handed the SAME table.
"""
from __future__ import annotations

import numpy as np


def zipf_thresholds(m: int, s: float) -> np.ndarray:
    """thr[k] = floor(2^64 * CDF(k+1)) for P(k) ~ k^-s, k = 1..m (uint64, last = 2^64 - 1).
    draw(r) = 1 + #{k : thr[k] < r} for a uniform 64-bit r."""
    w = np.arange(1, m + 1, dtype=np.float64) ** (-float(s))
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    thr = np.minimum(cdf * 18446744073709551616.0, 18446744073709549568.0).astype(np.uint64)
    thr[-1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return thr
