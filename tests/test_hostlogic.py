"""CPU tests of the HOST LOGIC (round loop, growth, chunking, column-0 phases, set resolution,
CSR assembly) — the C shim and the kernel sources compiled by gcc/g++ against the serial
simulator in tests/hostsim/ (warp width 1, one-thread blocks).  This checks logic only; the
parity tests proper are tests/test_gpu_parity.py on the real CUDA build."""
import os

import pytest

import parity_suite as ps
from libsmatrix_b200 import SparseMatrix


@pytest.fixture(scope="module")
def sim():
    from hostsim import build as hb
    return hb.build()


@pytest.fixture(params=[(6, 1 << 26, 0), (10, 3000, 0), (4, 777, 0), (8, 5000, 1), (12, 4000, 2)],
                ids=["dir64", "chunk3000", "dir16-chunk777", "partitioned", "partitioned-wide"])
def make(request, sim, monkeypatch):
    dir_log2, chunk, partitioned = request.param
    monkeypatch.setenv("SMATRIX_DIR_LOG2", str(dir_log2))   # tiny directory: growth on every test
    monkeypatch.setenv("SMATRIX_CHUNK", str(chunk))         # small chunks: multi-chunk batches
    if partitioned:                                         # chunks re-ordered by directory slice
        monkeypatch.setenv("SMATRIX_PARTITION_MIN", "16")
        monkeypatch.setenv("SMATRIX_SLICE_LOG2", "3")
    # chunks without ops on column 0 use 256 slices (the default) or, in the first partitioned variant, 128 + twins
    monkeypatch.setenv("SMATRIX_WIDE_SLICES", "1" if partitioned == 2 else "0")
    return lambda: SparseMatrix(_lib_path=sim)


def test_java_cases(make):
    ps.scenario_java_cases(make, grid=120)


def test_single_op_api(make):
    ps.scenario_single_op_api(make)


def test_example_program(make):
    ps.scenario_example_program(make)


def test_quirks(make):
    ps.scenario_quirks(make)


def test_empty_and_ragged(make):
    ps.scenario_empty_and_ragged(make)


@pytest.mark.parametrize("seed", range(6))
def test_random_streams(make, seed):
    ps.scenario_random(make, seed, n=8000, n_rows=80, n_cols=70, rounds=4)


def test_set_last_writer(make):
    ps.scenario_set_last_writer(make)


def test_hot_keys(make):
    ps.scenario_hot_keys(make)


def test_cf(make):
    ps.scenario_cf(make, n_baskets=600, n_items=150)


def test_col0_ordering(make):
    ps.scenario_col0_ordering(make)


def test_recycling_churn(make):
    ps.scenario_recycling_churn(make)


def test_big_row(make):
    ps.scenario_big_row(make, n_cols=20000)


def test_many_rows(make):
    ps.scenario_many_rows(make, 5000)


def test_model_crosscheck(make):
    ps.scenario_model_crosscheck(make)


def test_threads_single_ops(make):
    ps.scenario_threads_single_ops(make, per_thread=100)


def test_benchmark_pattern(make):
    ps.scenario_benchmark_pattern(make, threads=4, rounds=3)


def test_uniform_grid_c1b(make):
    ps.scenario_uniform_grid(make, n_ops=30000, side=300)


def test_read_path_zipf(make):
    ps.scenario_read_path_zipf(make, n_rows=300, max_len=3000)


def test_batch_out(make):
    ps.scenario_batch_out(make, n=6000)


def test_chunks_without_column0(make):
    ps.scenario_no_column0(make, n=9000, n_rows=2500)


def test_sliced_gets(make, monkeypatch):
    monkeypatch.setenv("SMATRIX_GET_SLICE_MIN", "16")       # mode 1 needs >= this many queries ...
    monkeypatch.setenv("SMATRIX_SLICE_LOG2", "3")           # ... and more than one directory slice
    ps.scenario_sliced_gets(make, n_rows=300, n_cols=40, n_ops=6000, sizes=(1, 2, 7, 8, 9, 15, 16, 17, 599, 600, 5001))


def test_cf_read_side(make):
    ps.scenario_cf_read_side(make, n_baskets=300, n_items=80)


def test_snapshot_interchange(sim, tmp_path, monkeypatch):
    import snapshot_suite as ss
    monkeypatch.setenv("SMATRIX_DIR_LOG2", "6")
    ss.scenario_snapshot_interchange(lambda f: SparseMatrix(f, _lib_path=sim), tmp_path)


def test_snapshot_roundtrip(sim, tmp_path, monkeypatch):
    import snapshot_suite as ss
    monkeypatch.setenv("SMATRIX_DIR_LOG2", "8")
    ss.scenario_snapshot_roundtrip_big(lambda f: SparseMatrix(f, _lib_path=sim), tmp_path, n_rows=3000)


def test_open_unwritable_path_fails(sim):
    import pytest
    with pytest.raises(ValueError):                 # smatrix_open -> NULL (src/smatrix.c:92-96)
        SparseMatrix("/nonexistent-dir/x.smx", _lib_path=sim)


def test_golden_fixtures(make):
    import numpy as np
    from oracle import cpu
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for f in sorted(x for x in os.listdir(gdir) if x.endswith(".npz")):
        g = np.load(os.path.join(gdir, f))
        m = make()
        for k in range(int(g["n_batches"])):
            getattr(m, str(g[f"op{k}"]) + "_batch")(g[f"xs{k}"], g[f"ys{k}"], g[f"vs{k}"])
        assert (np.asarray(m.get_batch(g["qx"], g["qy"])) == g["get"]).all(), f
        assert (np.asarray(m.rowlen_batch(g["rows"])) == g["rowlen"]).all(), f
        o, p = m.getrow_batch(g["rows"])
        assert (o == g["offsets"]).all() and (cpu.sort_rows(o, p) == g["pairs_sorted"]).all(), f
        m.close()


# ---- big-row re-placement tile by tile (k_migrate_tiles): a simulator variant with TINY tiles (8 sectors instead of
# 1024), where runs of full sectors reach a tile's end all the time, so the spill list, its insertion pass and the
# "list too short -> enlarge and run again" loop are exercised on every growth of a big row
@pytest.fixture(scope="module")
def sim_small_tiles():
    from hostsim import build as hb
    return hb.build(defines=["-DMIG_TILE_LOG=3u"], suffix="_smalltiles")


@pytest.mark.parametrize("spill_cap", [2, 1 << 16], ids=["spill-list-regrows", "spill-list-fits"])
def test_tile_replacement_with_spills(sim_small_tiles, monkeypatch, spill_cap):
    monkeypatch.setenv("SMATRIX_DIR_LOG2", "6")
    monkeypatch.setenv("SMATRIX_CHUNK", "30000")
    monkeypatch.setenv("SMATRIX_SPILL_CAP", str(spill_cap))
    spilled = []

    def make():
        m = SparseMatrix(_lib_path=sim_small_tiles)
        close = m.close
        m.close = lambda: (spilled.append(m.stat("spilled")), close())[1]
        return m
    ps.scenario_big_row(make, n_cols=60000)
    ps.scenario_read_path_zipf(make, n_rows=300, max_len=30000)
    ps.scenario_recycling_churn(make, waves=5, rows_per_wave=60)
    assert sum(spilled) > 0, "tiny tiles must spill"


def test_tile_replacement_off_is_identical(sim, monkeypatch):
    monkeypatch.setenv("SMATRIX_MIGRATE_TILES", "0")        # the reference path: one global CAS per cell
    monkeypatch.setenv("SMATRIX_DIR_LOG2", "6")
    ps.scenario_big_row(lambda: SparseMatrix(_lib_path=sim), n_cols=30000)
