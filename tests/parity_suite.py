"""Parity scenarios shared by the CPU host-logic tests (tests/test_hostlogic.py, serial simulator)
and the GPU parity tests proper (tests/test_gpu_parity.py, the CUDA library through the C-ABI).

Every scenario feeds identical op streams to the library under test (`make()` -> SparseMatrix)
and to the checker (`oracle.cpu.CpuMatrix`: the unmodified reference when oracle/_ref exists,
else the C restatement) and compares, bit-exact:
    get over hits + misses, rowlen (the reference's history-dependent counter), and getrow as
    (column, value) pairs sorted by column.
"""
from __future__ import annotations

import threading

import numpy as np

from oracle import cpu
from oracle.model import ModelMatrix
from conftest import safe_stream

U32 = np.uint32


def checker() -> cpu.CpuMatrix:
    return cpu.CpuMatrix("reference" if cpu.have_reference() else "port")


def compare(m, ref, rows, qx, qy):
    rows = np.asarray(rows, dtype=U32)
    got, want = np.asarray(m.get_batch(qx, qy)), ref.get_many(qx, qy)
    assert (got == want).all(), f"get: {int((got != want).sum())} mismatches"
    got, want = np.asarray(m.rowlen_batch(rows)), ref.rowlen_many(rows)
    assert (got == want).all(), f"rowlen: {int((got != want).sum())} mismatches"
    o1, p1 = m.getrow_batch(rows)
    o2, p2 = ref.getrow_many(rows)
    assert (o1 == o2).all(), "getrow: row sizes differ"
    assert (cpu.sort_rows(o1, p1) == cpu.sort_rows(o2, p2)).all(), "getrow: pairs differ"


def apply_both(m, ref, op, xs, ys, vs):
    getattr(m, op + "_batch")(xs, ys, vs)
    ref.apply(op, xs, ys, vs)


# ------------------------------------------------------------------ reference KATs
def scenario_java_cases(make, grid: int = 1000):
    """src/java/test/TestSparseMatrix.java:22-130 (cases 1-7), same order, one shared matrix."""
    m = make()
    m.set(42, 23, 17)
    assert m.get(42, 23) == 17
    m.set(4231, 2634, 0); m.incr(4231, 2634, 1)
    assert m.get(4231, 2634) == 1
    m.set(1231, 2634, 0); m.incr(1231, 2634, 1); m.incr(1231, 2634, 5)
    assert m.get(1231, 2634) == 6
    i, n = np.meshgrid(np.arange(grid, dtype=U32), np.arange(grid, dtype=U32))
    xs, ys = i.ravel(), n.ravel()
    m.set_batch(xs, ys, np.full(xs.shape, 34, U32))          # case 4, batched
    assert (np.asarray(m.get_batch(xs, ys)) == 34).all()
    for x in range(0, grid, max(1, grid // 7)):                # and a few through the single-op API
        assert m.get(x, x // 2) == 34
    xs = np.arange(grid, dtype=U32)
    m.incr_batch(xs, np.full(grid, 42, U32), np.ones(grid, U32))   # case 5
    assert m.getRowLength(42) == grid
    m.incr_batch(xs, np.full(grid, 85, U32), np.ones(grid, U32))   # case 6
    assert len(m.getRow(85)) == grid
    m.incr_batch(xs, np.full(grid, 83, U32), np.ones(grid, U32))   # case 7
    assert len(m.getRow(83, 230)) == min(230, grid)
    m.close()


def scenario_single_op_api(make):
    """The 8-function API one call at a time, against the checker, incl. return values (Q7)."""
    m, ref = make(), checker()
    rng = np.random.default_rng(7)
    for _ in range(400):
        op = ("set", "incr", "decr", "incr")[rng.integers(0, 4)]
        x, y = int(rng.integers(0, 12)), int(rng.integers(1, 40))
        v = int(rng.integers(0, 2**32))
        assert getattr(m, op)(x, y, v) == getattr(ref, op)(x, y, v)
    for x in range(12):
        m.incr(x, 0, x + 1); ref.incr(x, 0, x + 1)
    for x in range(14):
        assert m.getRowLength(x) == ref.rowlen(x)
        assert m.getRow(x) == dict(ref.getrow(x))
        for y in (0, 1, 7, 39, 40):
            assert m.get(x, y) == ref.get(x, y)
        for nbytes in (0, 8, 12, 16, 100, 10000):   # Q4: n = min(live, max(1, ceil(ret_len / 8)))
            assert len(m.getrow_raw(x, nbytes)) == len(ref.getrow_raw(x, nbytes))
    m.close(); ref.close()


def scenario_example_program(make):
    """examples/smatrix_example.c:17-48: 4 threads x 100 x incr over (1..29, 1..49)."""
    m = make()
    n, i = np.meshgrid(np.arange(1, 30, dtype=U32), np.arange(1, 50, dtype=U32), indexing="ij")
    xs, ys = np.tile(n.ravel(), 100), np.tile(i.ravel(), 100)
    threads = [threading.Thread(target=m.incr_batch, args=(xs, ys, np.ones(xs.shape, U32)))
               for _ in range(4)]
    [t.start() for t in threads]; [t.join() for t in threads]
    assert (np.asarray(m.get_batch(n.ravel(), i.ravel())) == 400).all()
    assert m.getRowLength(23) == 49
    m.close()


def scenario_quirks(make):
    m = make()
    m.incr(5, 0, 1)
    m.incr_batch(np.full(9, 5, U32), np.arange(1, 10, dtype=U32), None)
    assert m.getRowLength(5) == 9 and len(m.getRow(5)) == 10              # Q1
    m.incr(5, 10, 1)
    assert m.getRowLength(5) == 11
    m.set(2, 5, 0)
    assert m.getRowLength(2) == 1 and m.getRow(2) == {5: 0}               # Q2
    m.set(8, 0, 0)
    assert m.getRowLength(8) == 0 and m.getRow(8) == {}
    assert m.stat("rows") == 3                                             # the row itself exists
    assert m.get(77, 1) == 0 and m.getRowLength(77) == 0 and m.getRow(77) == {}   # Q6
    assert m.stat("rows") == 3                                             # ... and created nothing
    assert m.decr(9, 10, 3) == 2**32 - 3                                   # Q7
    assert m.incr(9, 10, 5) == 2
    for x, y in ((0, 0), (0, 1), (2**32 - 1, 2**32 - 1), (2**32 - 1, 0), (0, 2**32 - 1)):
        assert m.set(x, y, 0xDEADBEEF) == 0xDEADBEEF and m.get(x, y) == 0xDEADBEEF
    m.close()


def scenario_empty_and_ragged(make):
    m, ref = make(), checker()
    e = np.zeros(0, U32)
    m.incr_batch(e, e, e); m.set_batch(e, e, e); m.decr_batch(e, e, None)
    assert len(m.get_batch(e, e)) == 0 and len(m.rowlen_batch(e)) == 0
    o, p = m.getrow_batch(e)
    assert list(o) == [0] and len(p) == 0
    q = np.array([1, 2, 3], U32)
    assert (np.asarray(m.get_batch(q, q)) == 0).all() and (np.asarray(m.rowlen_batch(q)) == 0).all()
    o, p = m.getrow_batch(q)
    assert list(o) == [0, 0, 0, 0]
    # ragged: rows of length 0 (column-0-only with value 0), 1, 4, 5, 16, 17, 300
    xs, ys = [], []
    for r, ln in enumerate((0, 1, 4, 5, 16, 17, 300)):
        xs += [r] * (ln + 1); ys += [0] + list(range(1, ln + 1))
    xs, ys = np.array(xs, U32), np.array(ys, U32)
    vs = np.where(ys == 0, 0, 7).astype(U32)
    apply_both(m, ref, "set", xs, ys, vs)
    compare(m, ref, np.arange(9), xs, ys)
    m.close(); ref.close()


# ------------------------------------------------------------------ randomized streams
def scenario_random(make, seed: int, n: int = 30000, n_rows: int = 200, n_cols: int = 120,
                    rounds: int = 4):
    rng = np.random.default_rng(seed)
    m, ref = make(), checker()
    wide = seed % 2 == 1
    all_x, all_y = [], []
    for r in range(rounds):
        op = ("incr", "set", "decr", "incr", "set")[(seed + r) % 5]
        xs, ys, vs = safe_stream(rng, n, n_rows, n_cols, op, col0_rate=(0.0, 0.03, 0.25)[seed % 3],
                                 wide_keys=wide, max_val=3 if seed % 4 == 0 else 2**32 - 1)
        if op == "decr":   # never decrement column 0 (could return to 0: outside the safe domain)
            ys = np.where(ys == 0, ys.max(), ys).astype(U32)
        apply_both(m, ref, op, xs, ys, vs)
        all_x.append(xs); all_y.append(ys)
        rows = np.unique(xs)
        qx = np.concatenate([xs[:4000], rng.integers(0, 2**32, 500, dtype=np.uint64).astype(U32)])
        qy = np.concatenate([ys[:4000], rng.integers(0, 2**32, 500, dtype=np.uint64).astype(U32)])
        perm = rng.permutation(len(qx))
        compare(m, ref, np.concatenate([rows, qx[-20:]]), qx, qy[perm])
        compare(m, ref, rows[:50], qx, qy)
    m.close(); ref.close()


def scenario_set_last_writer(make):
    """Duplicate keys inside one set batch: the last one in input order wins."""
    rng = np.random.default_rng(11)
    m, ref = make(), checker()
    n = 50000
    xs = rng.integers(0, 20, n).astype(U32)
    ys = rng.integers(0, 25, n).astype(U32)            # ~100 writes per key, incl. column 0
    vs = rng.integers(1, 2**32, n, dtype=np.uint64).astype(U32)
    apply_both(m, ref, "set", xs, ys, vs)
    compare(m, ref, np.arange(22), xs, ys)
    # all writes to ONE key
    xs = np.full(5000, 3, U32); ys = np.full(5000, 9, U32)
    vs = np.arange(1, 5001, dtype=U32)
    apply_both(m, ref, "set", xs, ys, vs)
    assert m.get(3, 9) == 5000
    m.close(); ref.close()


def scenario_hot_keys(make):
    """Zipf-hot keys: long runs of identical (x, y) — the warp pre-aggregation path — with values
    that wrap mod 2^32."""
    rng = np.random.default_rng(13)
    m, ref = make(), checker()
    n = 60000
    xs = (rng.zipf(1.3, n) % 50).astype(U32)
    ys = (rng.zipf(1.3, n) % 40 + 1).astype(U32)
    runs = rng.random(n) < 0.5                           # make neighbours identical
    xs[1:][runs[1:]] = xs[:-1][runs[1:]]
    ys[1:][runs[1:]] = ys[:-1][runs[1:]]
    vs = rng.integers(2**31, 2**32, n, dtype=np.uint64).astype(U32)
    apply_both(m, ref, "incr", xs, ys, vs)
    apply_both(m, ref, "decr", xs[::2], ys[::2], vs[::2])
    apply_both(m, ref, "incr", xs, ys, None if False else np.ones(n, U32))
    compare(m, ref, np.arange(52), xs, ys)
    m.close(); ref.close()


def cf_stream(rng, n_baskets: int, n_items: int, basket: int = 8):
    """examples/cf_recommender.c:35-47: per basket, for each n: incr(ids[n], 0, 1) then
    incr(ids[n], ids[i], 1) for every i != n (no dedup inside a basket)."""
    ids = (rng.zipf(1.1, (n_baskets, basket)) % n_items + 1).astype(U32)
    pick = np.array([[-1] + [i for i in range(basket) if i != n] for n in range(basket)])
    xs = np.repeat(ids, basket, axis=1)                      # row ids[n], `basket` ops each
    ys = np.where(pick.ravel()[None, :] < 0, U32(0), ids[:, np.maximum(pick.ravel(), 0)])
    return xs.ravel().astype(U32), ys.ravel().astype(U32)


def scenario_cf(make, n_baskets: int = 3000, n_items: int = 400):
    """Co-occurrence build (BASELINE config 3 shape): every row's first write is column 0, so
    rowlen follows the n / n+1 rule — once in a single batch, once split across batches."""
    rng = np.random.default_rng(17)
    xs, ys = cf_stream(rng, n_baskets, n_items)
    for pieces in (1, 7):
        m, ref = make(), checker()
        for part_x, part_y in zip(np.array_split(xs, pieces), np.array_split(ys, pieces)):
            m.incr_batch(part_x, part_y, None)
            ref.apply("incr", part_x, part_y, np.ones(len(part_x), U32))
        compare(m, ref, np.arange(n_items + 2), xs[:20000], ys[:20000])
        m.close(); ref.close()


def scenario_col0_ordering(make):
    """Column 0 turning non-zero INSIDE a batch, at every position relative to the row's new
    columns around the virtual-resize counts 10 and 18 (DESIGN.md "t0")."""
    m, ref = make(), checker()
    xs, ys = [], []
    row = 0
    for pre in (0, 5, 8, 9):                 # columns already present before the batch
        for pos in range(0, 22):            # column-0 op goes after `pos` new columns
            row += 1
            cols = list(range(1, 23))
            if pre:
                apply_both(m, ref, "incr", np.full(pre, row, U32), np.array(cols[:pre], U32),
                           np.ones(pre, U32))
            new = cols[pre:]
            seq = new[:pos] + [0] + new[pos:]
            xs += [row] * len(seq); ys += seq
    xs, ys = np.array(xs, U32), np.array(ys, U32)
    # interleave the rows over the batch while keeping every row's own op order
    rng = np.random.default_rng(3)
    when = rng.random(len(xs))
    for r in np.unique(xs):
        sel = np.flatnonzero(xs == r)
        when[sel] = np.sort(when[sel])
    order = np.argsort(when, kind="stable")
    apply_both(m, ref, "incr", xs[order], ys[order], np.ones(len(xs), U32))
    compare(m, ref, np.arange(row + 2), xs, ys)
    # same shape with duplicates of the new columns before and after column 0, as a set batch
    xs2 = np.concatenate([xs, xs]) + U32(1000)
    ys2 = np.concatenate([ys, ys[::-1]])
    vs2 = np.arange(1, len(xs2) + 1, dtype=U32)
    apply_both(m, ref, "set", xs2, ys2, vs2)
    compare(m, ref, np.arange(1000, 1000 + row + 2), xs2, ys2)
    m.close(); ref.close()


def scenario_big_row(make, n_cols: int = 40000):
    """One row far beyond the grid-wide re-placement threshold (2^13 cells), grown across batches,
    next to many small rows; plus a read of the whole thing."""
    rng = np.random.default_rng(19)
    m, ref = make(), checker()
    cols = rng.permutation(np.arange(1, n_cols + 1, dtype=U32) * U32(2654435761))
    for part in np.array_split(cols, 5):
        xs = np.concatenate([np.full(len(part), 77, U32), rng.integers(100, 400, 2000).astype(U32)])
        ys = np.concatenate([part, rng.integers(1, 30, 2000).astype(U32)])
        vs = rng.integers(1, 2**32, len(xs), dtype=np.uint64).astype(U32)
        perm = rng.permutation(len(xs))
        apply_both(m, ref, "incr", xs[perm], ys[perm], vs[perm])
    assert m.getRowLength(77) == ref.rowlen(77) == n_cols
    compare(m, ref, np.array([77, 100, 101, 399, 5], U32), np.full(3000, 77, U32), cols[:3000])
    m.close(); ref.close()


def scenario_many_rows(make, n_rows: int):
    """More rows than the initial directory holds: directory growth (and the shrink-to-fit after
    an over-estimate caused by duplicate row ids in the batch)."""
    rng = np.random.default_rng(23)
    m, ref = make(), checker()
    ids = (np.arange(n_rows, dtype=U32) * U32(2654435761))
    xs = np.concatenate([ids, ids[rng.integers(0, n_rows, 3 * n_rows)]])
    ys = rng.integers(0, 6, len(xs)).astype(U32)
    vs = np.ones(len(xs), U32)
    apply_both(m, ref, "incr", xs, ys, vs)
    assert m.stat("rows") == n_rows
    assert m.stat("dir_cap") >= 4 * n_rows
    sample = rng.integers(0, len(xs), 20000)
    compare(m, ref, ids[:: max(1, n_rows // 3000)], xs[sample], ys[sample])
    assert m.stat("nnz") == len(np.unique(xs.astype(np.uint64) << np.uint64(32) | ys))
    assert m.stat("value_sum") == len(xs)        # every op added 1
    m.close(); ref.close()


def scenario_model_crosscheck(make):
    """The pure-Python semantic model as a third, compiler-free checker on a small stream."""
    rng = np.random.default_rng(29)
    m, model = make(), ModelMatrix()
    for op in ("incr", "set", "incr"):
        xs, ys, vs = safe_stream(rng, 3000, 30, 50, op, col0_rate=0.15)
        getattr(m, op + "_batch")(xs, ys, vs)
        for x, y, v in zip(xs, ys, vs):
            getattr(model, op)(int(x), int(y), int(v))
    for x in range(31):
        assert m.getRowLength(x) == model.rowlen(x)
        assert list(m.getRow(x).items()) == model.getrow(x)
    m.close()


def scenario_threads_single_ops(make, per_thread: int = 300):
    """README.md:113,120 'all of the methods are threadsafe': concurrent single-op callers."""
    m = make()

    def work(t):
        for k in range(per_thread):
            m.incr(k % 7, 1 + (k % 5), 1)
            m.incr(100 + t, 1, 2)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in threads]; [t.join() for t in threads]
    total = sum(m.get(x, y) for x in range(7) for y in range(1, 6))
    assert total == 4 * per_thread
    assert all(m.get(100 + t, 1) == 2 * per_thread for t in range(4))
    m.close()


def scenario_benchmark_pattern(make, threads: int = 8, rounds: int = 16):
    """BASELINE config 1: the reference benchmark's fixed pattern (src/smatrix_benchmark.c:29-65),
    thread t uses offset o = 42 + t and does incr(n+o, i+o, 1); incr(i+o, n+o, 1) over n < 23,
    i < 22 — a few thousand distinct keys hit over and over (pure pre-aggregation / contention)."""
    m, ref = make(), checker()
    n, i = np.meshgrid(np.arange(23, dtype=U32), np.arange(22, dtype=U32), indexing="ij")
    xs, ys = [], []
    for t in range(threads):
        o = U32(42 + t)
        a = np.stack([n.ravel() + o, i.ravel() + o], axis=1)      # incr(n+o, i+o)
        b = np.stack([i.ravel() + o, n.ravel() + o], axis=1)      # incr(i+o, n+o)
        pat = np.stack([a, b], axis=1).reshape(-1, 2)             # interleaved like the C loop
        xs.append(np.tile(pat[:, 0], rounds)); ys.append(np.tile(pat[:, 1], rounds))
    # threads interleave arbitrarily in the reference; addition commutes, any order gives the same matrix
    xs, ys = np.concatenate(xs), np.concatenate(ys)
    apply_both(m, ref, "incr", xs, ys, np.ones(len(xs), U32))
    hi = 42 + threads + 23
    qx, qy = np.meshgrid(np.arange(40, hi, dtype=U32), np.arange(40, hi, dtype=U32))
    compare(m, ref, np.arange(40, hi), qx.ravel(), qy.ravel())     # benchmark_get_mixed's key space
    assert int(np.asarray(m.get_batch(qx.ravel(), qy.ravel())).astype(np.uint64).sum()) == len(xs)
    m.close(); ref.close()


def splitmix64(z):
    """SURVEY.md 8(d) RNG: r = splitmix64(seed + i), vectorised (uint64 wrap-around intended)."""
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def scenario_uniform_grid(make, n_ops: int = 4_000_000, side: int = 10000, pieces: int = 3):
    """BASELINE config 1 as BASELINE.json words it (SURVEY.md 8d "C1b"): uniform-random incr over a
    side x side grid, x = (r >> 32) % side, y = r % side, seed 1.  Row 0 and column 0 are ordinary
    members of the grid here, so ~n/side ops land on column 0 of random rows at random moments:
    rowlen depends on when (Q1) and must still match the sequential reference.  Exhaustive compare:
    every distinct key of the stream, every row."""
    with np.errstate(over="ignore"):
        r = splitmix64(np.uint64(1) + np.arange(n_ops, dtype=np.uint64))
    xs = ((r >> np.uint64(32)) % np.uint64(side)).astype(U32)
    ys = ((r & np.uint64(0xFFFFFFFF)) % np.uint64(side)).astype(U32)
    m, ref = make(), checker()
    for px, py in zip(np.array_split(xs, pieces), np.array_split(ys, pieces)):
        m.incr_batch(px, py, None)
        ref.apply("incr", px, py, np.ones(len(px), U32))
    keys = np.unique((xs.astype(np.uint64) << np.uint64(32)) | ys.astype(np.uint64))
    qx, qy = (keys >> np.uint64(32)).astype(U32), (keys & np.uint64(0xFFFFFFFF)).astype(U32)
    compare(m, ref, np.arange(side + 3, dtype=U32), np.concatenate([qx, qx]), np.concatenate([qy, qy + U32(side)]))
    assert m.stat("nnz") == len(keys)
    m.close(); ref.close()


def scenario_read_path_zipf(make, n_rows: int = 3000, max_len: int = 20000, seed: int = 31):
    """BASELINE config 4 shape: rows with Zipf lengths P(k) ~ k^-1.7 on [1, max_len], distinct random
    uint32 columns >= 1, random non-zero values; rowlen over all rows, then getrow over all rows,
    compared sorted by column."""
    rng = np.random.default_rng(seed)
    k = np.arange(1, max_len + 1, dtype=np.float64)
    pk = k ** -1.7
    lens = rng.choice(np.arange(1, max_len + 1), size=n_rows, p=pk / pk.sum())
    lens[0] = max_len                                        # make sure the longest row exists
    ids = (np.arange(n_rows, dtype=U32) * U32(2654435761))
    xs = np.repeat(ids, lens)
    ys = np.concatenate([rng.choice(2**32 - 2, size=int(L), replace=False).astype(np.uint64) + 1
                         for L in lens]).astype(U32)
    vs = rng.integers(1, 2**32, len(xs), dtype=np.uint64).astype(U32)
    perm = rng.permutation(len(xs))
    m, ref = make(), checker()
    apply_both(m, ref, "set", xs[perm], ys[perm], vs[perm])
    assert (np.asarray(m.rowlen_batch(ids)) == lens).all()    # no column 0 anywhere: rowlen == length
    sample = rng.integers(0, len(xs), 5000)
    compare(m, ref, np.concatenate([ids, np.array([1, 2, 3], U32)]), xs[sample], ys[sample])
    m.close(); ref.close()


def scenario_cf_read_side(make, n_baskets: int = 2000, n_items: int = 300):
    """SURVEY.md 8f N4 — neighbors_for_item + cf_cosine (examples/cf_recommender.c:50-86) for a batch of
    items against the same arithmetic done on the CPU with the checker's get / getrow; IEEE double
    sqrt, multiply and divide are correctly rounded on both sides, so the scores match bit for bit."""
    rng = np.random.default_rng(53)
    xs, ys = cf_stream(rng, n_baskets, n_items)
    m, ref = make(), checker()
    m.incr_batch(xs, ys, None); ref.apply("incr", xs, ys, np.ones(len(xs), U32))
    items = np.concatenate([np.arange(1, n_items + 1, dtype=U32)[::3], np.array([0, 99999], U32)])
    offsets, ids, scores = m.cf_neighbors_batch(items)
    o2, p2 = ref.getrow_many(items)
    assert (offsets == o2).all()
    for i, a in enumerate(items):
        lo, hi = int(offsets[i]), int(offsets[i + 1])
        got = dict(zip(ids[lo:hi].tolist(), scores[lo:hi].tolist()))
        a_total = ref.get(int(a), 0)
        want = {}
        for b, cc in p2[lo:hi]:
            b_total = ref.get(int(b), 0) or 1
            den = np.sqrt(np.float64(a_total)) * np.sqrt(np.float64(b_total))
            want[int(b)] = 0.0 if (den == 0.0 or np.float64(cc) > den) else float(np.float64(cc) / den)
        assert got == want, f"item {a}"
    m.close(); ref.close()


def scenario_batch_out(make, n: int = 40000):
    """SURVEY.md 8f N1 — per-op return values of batches = the return values of the reference's
    single-op calls applied in input order (src/smatrix.c:230,241,252), with heavy key duplication,
    column 0, wrap-around and growth inside the batch."""
    rng = np.random.default_rng(59)
    m, ref = make(), checker()
    for op in ("incr", "decr", "set", "incr"):
        xs = (rng.zipf(1.4, n) % 60).astype(U32)
        ys = (rng.zipf(1.4, n) % 45).astype(U32)                 # includes column 0
        vs = rng.integers(1, 2**32, n, dtype=np.uint64).astype(U32)
        if op != "set":                                           # keep column 0 inside the safe domain
            vs = np.where(ys == 0, (vs % 1000) + 1, vs).astype(U32)
        if op == "decr":
            ys = np.where(ys == 0, U32(1), ys).astype(U32)
        got = np.asarray(getattr(m, op + "_batch_out")(xs, ys, vs))
        want = ref.apply(op, xs, ys, vs, want_out=True)
        assert (got == want).all(), f"{op}: {int((got != want).sum())} return values differ"
    got = np.asarray(m.incr_batch_out(xs, ys, None))             # vals == NULL: every value is 1
    want = ref.apply("incr", xs, ys, np.ones(n, U32), want_out=True)
    assert (got == want).all()
    compare(m, ref, np.arange(62), xs, ys)
    m.close(); ref.close()


def scenario_recycling_churn(make, waves: int = 6, rows_per_wave: int = 400):
    """Bucket recycling (replaces smatrix_mfree on every resize, src/smatrix.c:383-416): rows are born in
    waves and grow through every size class at different times, so buckets vacated by older rows are
    reused by younger ones — and a reused bucket must never leak a previous owner's cells."""
    rng = np.random.default_rng(71)
    m, ref = make(), checker()
    born = []
    for w in range(waves):
        ids = (np.arange(w * rows_per_wave, (w + 1) * rows_per_wave, dtype=U32) * U32(2654435761))
        born.append(ids)
        xs, ys = [], []
        for age, old in enumerate(reversed(born)):           # older rows get more (and new) columns each wave
            k = 6 * 3 ** min(age, 5)                          # 6, 18, 54, ... columns per row this wave
            xs.append(np.repeat(old, k))
            ys.append(rng.integers(1, 40 * 3 ** min(age, 5) + 40, size=len(old) * k).astype(U32))
        xs, ys = np.concatenate(xs), np.concatenate(ys)
        perm = rng.permutation(len(xs))
        vs = rng.integers(1, 2**32, len(xs), dtype=np.uint64).astype(U32)
        apply_both(m, ref, "incr", xs[perm], ys[perm], vs[perm])
    allrows = np.concatenate(born)
    sample = rng.integers(0, len(xs), 20000)
    compare(m, ref, np.concatenate([allrows, np.array([7, 8], U32)]), xs[sample], ys[sample])
    assert m.stat("recycled") > 0, "younger rows should have reused buckets vacated by older ones"
    assert m.stat("free_bytes") >= 0 and m.stat("bucket_bytes") >= m.stat("live_bucket_bytes")
    m.close(); ref.close()


def scenario_sliced_gets(make, n_rows: int = 3000, n_cols: int = 90, n_ops: int = 60000,
                         sizes=(1, 2, 7, 8, 9, 2047, 2048, 2049, 50001)):
    """Point reads on DEVICE arrays in the three modes of smatrix_b200_set_get_slices (input order /
    by directory slice when rows repeat / always by slice): the answers must be those of smatrix_get
    (src/smatrix.c:174-185) in INPUT order whatever order the look-ups ran in — hits, missing columns,
    missing rows, column 0, duplicates; batch sizes around the partition tile (2048 queries on the GPU,
    8 on the simulator)."""
    from libsmatrix_b200.matrix import DevPtr
    rng = np.random.default_rng(99)
    m, ref = make(), checker()
    xs = (rng.integers(0, n_rows, n_ops).astype(U32) * U32(2654435761))
    ys = rng.integers(0, n_cols, n_ops).astype(U32)                    # column 0 included
    vs = rng.integers(1, 2**32, n_ops, dtype=np.uint64).astype(U32)
    apply_both(m, ref, "incr", xs, ys, vs)
    import os
    can_slice = m.stat("dir_cap") > (1 << int(os.environ.get("SMATRIX_SLICE_LOG2", 17)))   # more than one slice
    part_min = int(os.environ.get("SMATRIX_GET_SLICE_MIN", 1 << 22))

    def up(a):
        p = m.dev_alloc(max(a.nbytes, 8))
        if a.nbytes:
            m.memcpy(p, a.ctypes.data, a.nbytes)
        return p

    for n in sizes:
        pick = rng.integers(0, n_ops, n)
        qx, qy = xs[pick].copy(), ys[pick].copy()
        qy[1::3] += U32(n_cols)                                        # columns nobody wrote
        qx[2::7] += U32(1)                                             # rows nobody wrote
        want = ref.get_many(qx, qy)
        dx, dy, do = up(qx), up(qy), m.dev_alloc(4 * n + 8)
        auto = can_slice and 2 * n >= m.stat("rows") and n >= part_min
        # + the measurement switches: 4 = ordinary L2 priority for bucket sectors, 8 = the write path's slice
        # count, 16 = resident-grid look-up kernel
        for mode, sliced in ((0, False), (2, can_slice), (1, auto), (2 | 4, can_slice), (2 | 8, can_slice),
                             (2 | 16, can_slice), (1 | 4 | 8 | 16, auto)):
            m.set_get_slices(mode)
            before = m.stat("sliced_gets")
            m._lib.smatrix_b200_memset0(m._handle(), do, 4 * n)
            m.get_batch(DevPtr(dx, n), DevPtr(dy, n), DevPtr(do, n))
            got = np.empty(n, U32)
            m.memcpy(got.ctypes.data, do, 4 * n)
            assert (got == want).all(), f"mode {mode}, n {n}: {int((got != want).sum())} mismatches"
            took = m.stat("sliced_gets") - before
            assert took == (n if sliced and n >= 2 else 0), (mode, n, took)
        for p in (dx, dy, do):
            m.dev_free(p)
    m.set_get_slices(1)
    m.close(); ref.close()


def scenario_no_column0(make, n: int = 200000, n_rows: int = 40000, batches: int = 5):
    """Streams that never touch column 0 (configs 2 and 5): with SMATRIX_WIDE_SLICES such a chunk is ordered over
    256 slices without column-0 parts; a chunk in between that does write column 0 falls back to 128 + 128."""
    import os
    rng = np.random.default_rng(12)
    m, ref = make(), checker()
    for k in range(batches):
        xs = (rng.integers(0, n_rows, n).astype(U32) * U32(2654435761))
        ys = rng.integers(1, 60, n).astype(U32)
        if k == 3:
            ys[::17] = 0
        vs = rng.integers(1, 2**32, n, dtype=np.uint64).astype(U32)
        apply_both(m, ref, "incr" if k != 2 else "set", xs, ys, vs)
    rows = np.unique(xs)
    compare(m, ref, np.concatenate([rows, rows + U32(1)]), xs, ys)
    wide = m.stat("wide_chunks")
    if os.environ.get("SMATRIX_WIDE_SLICES", "1") == "1" and os.environ.get("SMATRIX_SLICE_LOG2") in ("3", "6"):
        assert wide >= batches - 1, "chunks without column 0 should have used 256 slices"
    elif os.environ.get("SMATRIX_WIDE_SLICES", "1") != "1":
        assert wide == 0
    m.close(); ref.close()
