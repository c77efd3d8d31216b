"""The synthetic streams of SURVEY.md 8(d) are generated twice — by device kernels (smatrix_b200_gen_*, what
bench.py feeds the GPU with) and by the CPU checker's C generators (what the reference is fed with).  Both must
produce the SAME ops from the same counter-based definition and the same threshold tables, or the CPU baseline
and the parity blocks of bench.py would talk about different workloads."""
import numpy as np
import pytest

from libsmatrix_b200 import SparseMatrix
from libsmatrix_b200.workloads import zipf_thresholds
from oracle import cpu

U32 = np.uint32


def _dev_arrays(m, *arrays):
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a)
        p = m.dev_alloc(max(a.nbytes, 8))
        m.memcpy(p, a.ctypes.data, a.nbytes)
        out.append(p)
    return out


def _down(m, ptr, n, dtype=U32):
    out = np.empty(n, dtype=dtype)
    m.memcpy(out.ctypes.data, ptr, out.nbytes)
    return out


def _check_streams(m):
    n = 50_000
    # c2 / c5
    dx, dy = m.dev_alloc(4 * n), m.dev_alloc(4 * n)
    for rows, ycols, seed, first in ((13_000_000, 256, 2, 0), (52_000_000, 170, 6, 9_999_999_000)):
        m.gen_c2_ops(seed, first, n, rows, ycols, dx, dy)
        hx, hy = cpu.gen_c2_ops(seed, first, n, rows, ycols)
        assert (_down(m, dx, n) == hx).all() and (_down(m, dy, n) == hy).all()
        m.gen_c2_queries(3, seed, first, n, 2_000_000_000, rows, ycols, dx, dy)
        hx, hy = cpu.gen_c2_queries(3, seed, first, n, 2_000_000_000, rows, ycols)
        assert (_down(m, dx, n) == hx).all() and (_down(m, dy, n) == hy).all()
    # c3: Zipf(1.1) baskets of 8
    thr = zipf_thresholds(40_000, 1.1)
    (d_thr,) = _dev_arrays(m, thr)
    for first in (0, 64 * 12345 + 17):
        m.gen_c3_ops(4, first, n, d_thr, len(thr), dx, dy)
        hx, hy = cpu.gen_c3_ops(4, first, n, thr)
        gx, gy = _down(m, dx, n), _down(m, dy, n)
        assert (gx == hx).all() and (gy == hy).all()
        assert gx.min() >= 1 and gx.max() <= len(thr)
        k = (first + np.arange(n)) % 64
        assert ((gy == 0) == (k % 8 == 0)).all()                      # the first op of every item is its column-0 total
        m.gen_c3_queries(7, 4, first, n, 1_000_000, d_thr, len(thr), dx, dy)
        hx, hy = cpu.gen_c3_queries(7, 4, first, n, 1_000_000, thr)
        assert (_down(m, dx, n) == hx).all() and (_down(m, dy, n) == hy).all()
    # c4: Zipf(1.7) row lengths, distinct non-zero columns per row, odd values
    thr4 = zipf_thresholds(5_000, 1.7)
    (d_thr4,) = _dev_arrays(m, thr4)
    rows = 3_000
    d_lens = m.dev_alloc(4 * rows)
    m.gen_c4_lens(5, 0, rows, d_thr4, len(thr4), d_lens)
    lens = _down(m, d_lens, rows)
    assert (lens == cpu.gen_c4_lens(5, 0, rows, thr4)).all() and lens.min() >= 1 and lens.max() <= len(thr4)
    offs = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))]).astype(np.uint64)
    total = int(offs[-1])
    (d_offs,) = _dev_arrays(m, offs)
    dx4, dy4, dv4 = m.dev_alloc(4 * total), m.dev_alloc(4 * total), m.dev_alloc(4 * total)
    m.gen_c4_ops(5, 0, total, d_offs, rows, dx4, dy4, dv4)
    hx, hy, hv = cpu.gen_c4_ops(5, 0, total, offs)
    gx, gy, gv = _down(m, dx4, total), _down(m, dy4, total), _down(m, dv4, total)
    assert (gx == hx).all() and (gy == hy).all() and (gv == hv).all()
    assert (gy != 0).all() and (gv % 2 == 1).all()
    for r in range(0, rows, 97):                                       # columns are distinct inside a row
        seg = gy[int(offs[r]):int(offs[r + 1])]
        assert len(np.unique(seg)) == len(seg)
        assert (gx[int(offs[r]):int(offs[r + 1])] == U32((r * 2654435761) & 0xFFFFFFFF)).all()
    for p in (dx, dy, d_thr, d_thr4, d_lens, d_offs, dx4, dy4, dv4):
        m.dev_free(p)


def test_device_and_host_generators_agree_on_the_simulator():
    from hostsim import build as hb
    m = SparseMatrix(_lib_path=hb.build())
    _check_streams(m)
    m.close()


@pytest.mark.gpu
def test_device_and_host_generators_agree_on_the_gpu():
    m = SparseMatrix(device=0)
    _check_streams(m)
    m.close()


def test_threshold_table_is_a_cdf():
    thr = zipf_thresholds(1000, 1.1)
    assert thr.dtype == np.uint64 and (np.diff(thr.astype(np.float64)) >= 0).all() and thr[-1] == np.uint64(2**64 - 1)
    w = np.arange(1, 1001, dtype=np.float64) ** -1.1
    assert abs(float(thr[0]) / 2.0**64 - w[0] / w.sum()) < 1e-12
