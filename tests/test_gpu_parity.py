"""GPU parity tests proper: the CUDA library (libsmatrix_b200/lib/libsmatrix_b200.so) through its
C-ABI against the CPU checker, bit-exact.  Run with `pytest -m gpu` on the B200 box."""
import os

import numpy as np
import pytest

import parity_suite as ps
from libsmatrix_b200 import SparseMatrix
from oracle import cpu

pytestmark = pytest.mark.gpu
U32 = np.uint32
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(params=["default", "chunk65536", "partitioned", "arena", "partitioned-wide"])
def make(request, monkeypatch):
    if request.param == "partitioned-wide":                # 256 slices for chunks without ops on column 0
        monkeypatch.setenv("SMATRIX_PARTITION_MIN", "64")
        monkeypatch.setenv("SMATRIX_SLICE_LOG2", "6")
        monkeypatch.setenv("SMATRIX_DIR_LOG2", "14")
        monkeypatch.setenv("SMATRIX_WIDE_SLICES", "1")
    if request.param == "arena":                           # slab + directory carved from a reserved arena
        monkeypatch.setenv("SMATRIX_ARENA_GIB", "2")
        monkeypatch.setenv("SMATRIX_DIR_LOG2", "12")
    if request.param == "chunk65536":
        monkeypatch.setenv("SMATRIX_CHUNK", "65536")       # multi-chunk batches
        monkeypatch.setenv("SMATRIX_DIR_LOG2", "12")       # directory growth from 4096 entries
    if request.param == "partitioned":                     # every chunk re-ordered by directory slice
        monkeypatch.setenv("SMATRIX_PARTITION_MIN", "64")
        monkeypatch.setenv("SMATRIX_SLICE_LOG2", "6")
        monkeypatch.setenv("SMATRIX_DIR_LOG2", "10")
        monkeypatch.setenv("SMATRIX_WIDE_SLICES", "0")     # 128 slices + column-0 twins for every chunk
    return lambda: SparseMatrix()


@pytest.fixture
def make_default():
    return lambda: SparseMatrix()


def test_java_cases(make):
    ps.scenario_java_cases(make, grid=1000)


def test_single_op_api(make_default):
    ps.scenario_single_op_api(make_default)


def test_example_program(make):
    ps.scenario_example_program(make)


def test_quirks(make_default):
    ps.scenario_quirks(make_default)


def test_empty_and_ragged(make):
    ps.scenario_empty_and_ragged(make)


@pytest.mark.parametrize("seed", range(8))
def test_random_streams(make, seed):
    ps.scenario_random(make, seed, n=200000, n_rows=3000, n_cols=300, rounds=4)


def test_set_last_writer(make):
    ps.scenario_set_last_writer(make)


def test_hot_keys(make):
    ps.scenario_hot_keys(make)


def test_cf(make):
    ps.scenario_cf(make, n_baskets=20000, n_items=3000)


def test_col0_ordering(make):
    ps.scenario_col0_ordering(make)


def test_recycling_churn(make):
    ps.scenario_recycling_churn(make)


def test_big_row(make):
    ps.scenario_big_row(make, n_cols=300000)


def test_many_rows(make_default):
    ps.scenario_many_rows(make_default, 1_500_000)     # > 2^20 / 2: the default directory must grow


def test_model_crosscheck(make_default):
    ps.scenario_model_crosscheck(make_default)


def test_threads_single_ops(make_default):
    ps.scenario_threads_single_ops(make_default, per_thread=200)


def test_benchmark_pattern(make):
    ps.scenario_benchmark_pattern(make, threads=32, rounds=32)      # 1 036 288 ops per cell at T=32 x 32


def test_uniform_grid_c1b(make):
    ps.scenario_uniform_grid(make, n_ops=8_000_000, side=10000)


def test_read_path_zipf(make):
    ps.scenario_read_path_zipf(make, n_rows=20000, max_len=200000)


def test_batch_out(make):
    ps.scenario_batch_out(make, n=400000)


def test_batch_out_device_pointers():
    import torch
    m, ref = SparseMatrix(), ps.checker()
    rng = np.random.default_rng(61)
    n = 500000
    xs = rng.integers(0, 2000, n).astype(U32); ys = rng.integers(1, 50, n).astype(U32)
    vs = rng.integers(0, 2**32, n, dtype=np.uint64).astype(U32)
    dev = torch.device("cuda", m.device)
    t = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
    got = m.incr_batch_out(t(xs), t(ys), t(vs)).cpu().numpy().view(U32)
    assert (got == ref.apply("incr", xs, ys, vs, want_out=True)).all()
    m.close(); ref.close()


def test_cf_read_side(make_default):
    ps.scenario_cf_read_side(make_default, n_baskets=20000, n_items=2000)


def test_snapshot_interchange(tmp_path):
    import snapshot_suite as ss
    ss.scenario_snapshot_interchange(lambda f: SparseMatrix(f), tmp_path)


def test_snapshot_roundtrip(tmp_path):
    import snapshot_suite as ss
    ss.scenario_snapshot_roundtrip_big(lambda f: SparseMatrix(f), tmp_path, n_rows=200000)


def test_preaggregation_off_is_identical(monkeypatch):
    monkeypatch.setenv("SMATRIX_PREAGG", "0")
    ps.scenario_hot_keys(lambda: SparseMatrix())


def test_golden_fixtures():
    """Committed input/output vectors generated from the unmodified reference
    (tests/golden/make_golden.py)."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    assert files
    for f in files:
        g = np.load(os.path.join(GOLDEN, f))
        m = SparseMatrix()
        for k in range(int(g["n_batches"])):
            op = str(g[f"op{k}"])
            getattr(m, op + "_batch")(g[f"xs{k}"], g[f"ys{k}"], g[f"vs{k}"])
        assert (np.asarray(m.get_batch(g["qx"], g["qy"])) == g["get"]).all(), f
        assert (np.asarray(m.rowlen_batch(g["rows"])) == g["rowlen"]).all(), f
        o, p = m.getrow_batch(g["rows"])
        assert (o == g["offsets"]).all() and (cpu.sort_rows(o, p) == g["pairs_sorted"]).all(), f
        m.close()


def test_device_pointer_batches():
    """Batches that already live in HBM (torch CUDA tensors -> raw device pointers, zero-copy)."""
    import torch
    m, ref = SparseMatrix(), ps.checker()
    rng = np.random.default_rng(5)
    n = 300000
    xs = rng.integers(0, 5000, n).astype(U32); ys = rng.integers(0, 200, n).astype(U32)
    vs = rng.integers(0, 1000, n).astype(U32)
    dev = torch.device("cuda", m.device)
    t = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
    m.incr_batch(t(xs), t(ys), t(vs)); ref.apply("incr", xs, ys, vs)
    m.incr_batch(t(xs), t(ys), None); ref.apply("incr", xs, ys, np.ones(n, U32))
    out = m.get_batch(t(xs), t(ys))
    assert out.is_cuda
    assert (out.cpu().numpy().view(U32) == ref.get_many(xs, ys)).all()
    rl = m.rowlen_batch(t(np.arange(5100, dtype=U32)))
    assert (rl.cpu().numpy().view(U32) == ref.rowlen_many(np.arange(5100, dtype=U32))).all()
    m.close(); ref.close()


def test_chunks_without_column0(make):
    ps.scenario_no_column0(make)


def test_sliced_gets(make):
    """smatrix_get_batch on device arrays: input order, by directory slice when rows repeat, always by slice."""
    ps.scenario_sliced_gets(make)


def test_sliced_gets_many_rows(monkeypatch):
    """The production geometry: a directory of 2^21 entries = 16 slices of 2^17, 2^22 queries over 400 K rows
    (mode 1 picks the slice order by itself), against the reference."""
    monkeypatch.setenv("SMATRIX_DIR_LOG2", "21")
    ps.scenario_sliced_gets(lambda: SparseMatrix(), n_rows=400_000, n_cols=30, n_ops=3_000_000,
                            sizes=(1 << 20, (1 << 22) + 5))


def test_c2_stream_device_generator_matches_host():
    import torch
    m = SparseMatrix()
    dev = torch.device("cuda", m.device)
    n = 100000
    dx = torch.empty(n, dtype=torch.int32, device=dev); dy = torch.empty_like(dx)
    m.gen_c2_ops(2, 12345, n, 13_000_000, 256, dx.data_ptr(), dy.data_ptr())
    hx, hy = cpu.gen_c2_ops(2, 12345, n, 13_000_000, 256)
    assert (dx.cpu().numpy().view(U32) == hx).all() and (dy.cpu().numpy().view(U32) == hy).all()
    m.gen_c2_queries(3, 2, 777, n, 5_000_000, 13_000_000, 256, dx.data_ptr(), dy.data_ptr())
    hx, hy = cpu.gen_c2_queries(3, 2, 777, n, 5_000_000, 13_000_000, 256)
    assert (dx.cpu().numpy().view(U32) == hx).all() and (dy.cpu().numpy().view(U32) == hy).all()
    m.close()


def test_c2_scaled_build_and_gets():
    """BASELINE config 2 at 1/40 scale (325 K rows x 256 columns, 50 M ops) built from the
    device-side stream and checked cell-for-cell against the reference run with pthreads
    (exact for incr), plus the size-independent properties used at full scale."""
    import torch
    rows, ycols, n_ops, seed = 325_000, 256, 50_000_000, 2
    m = SparseMatrix()
    dev = torch.device("cuda", m.device)
    step = 1 << 24
    dx = torch.empty(step, dtype=torch.int32, device=dev); dy = torch.empty_like(dx)
    for first in range(0, n_ops, step):
        cnt = min(step, n_ops - first)
        m.gen_c2_ops(seed, first, cnt, rows, ycols, dx.data_ptr(), dy.data_ptr())
        m.incr_batch(dx[:cnt], dy[:cnt], None)
    ref = ps.checker()
    threads = min(8, os.cpu_count() or 1) if ref.kind == "reference" else 1
    ref.bench_c2_incr(threads, seed, 0, n_ops, rows, ycols)
    nq = 2_000_000
    qx = torch.empty(nq, dtype=torch.int32, device=dev); qy = torch.empty_like(qx)
    m.gen_c2_queries(3, seed, 0, nq, n_ops, rows, ycols, qx.data_ptr(), qy.data_ptr())
    got = m.get_batch(qx, qy).cpu().numpy().view(U32)
    hx, hy = cpu.gen_c2_queries(3, seed, 0, nq, n_ops, rows, ycols)
    want = ref.get_many(hx, hy)
    assert (got == want).all()
    assert (got[1::2] == 0).all() and (got[0::2] > 0).all()          # exactly 50 % hits
    ids = (np.arange(0, rows, 97, dtype=U32) * U32(2654435761))
    ps.compare(m, ref, ids, hx[:100000], hy[:100000])
    # checksum of checksums: sum of all values == number of ops; nnz == distinct keys
    allrows = (np.arange(rows, dtype=U32) * U32(2654435761))
    o, p = m.getrow_batch(allrows)
    assert int(p[:, 1].astype(np.uint64).sum()) == n_ops
    assert m.stat("nnz") == len(p)
    m.close(); ref.close()
