"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference
(oracle/_ref/libsmatrix_ref.so, compiled from /root/reference/src/smatrix.c by oracle/Makefile).

    python tests/golden/make_golden.py

Each .npz holds a sequence of batches (op, xs, ys, vs) inside the reference's safe domain and the
reference's answers: get over a query set, rowlen and column-sorted getrow for a set of rows."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import cpu  # noqa: E402
from conftest import safe_stream  # noqa: E402
from parity_suite import cf_stream  # noqa: E402

U32 = np.uint32


def record(name, batches, qx, qy, rows):
    ref = cpu.CpuMatrix("reference")
    out = {"n_batches": len(batches)}
    for k, (op, xs, ys, vs) in enumerate(batches):
        ref.apply(op, xs, ys, vs)
        out[f"op{k}"] = op
        out[f"xs{k}"], out[f"ys{k}"], out[f"vs{k}"] = xs, ys, vs
    out["qx"], out["qy"], out["get"] = qx, qy, ref.get_many(qx, qy)
    out["rows"], out["rowlen"] = rows, ref.rowlen_many(rows)
    o, p = ref.getrow_many(rows)
    out["offsets"], out["pairs_sorted"] = o, cpu.sort_rows(o, p)
    ref.close()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k in ("get", "pairs_sorted")})


def main():
    cpu.build(ref=True)
    rng = np.random.default_rng(2024)
    # 1. mixed incr/set/decr with column-0 traffic, dense ids
    batches = []
    for op in ("incr", "set", "incr", "decr", "set"):
        xs, ys, vs = safe_stream(rng, 6000, 60, 90, op, col0_rate=0.1)
        if op == "decr":
            ys = np.where(ys == 0, U32(1), ys).astype(U32)
        batches.append((op, xs, ys, vs))
    qx = rng.integers(0, 64, 4000).astype(U32); qy = rng.integers(0, 95, 4000).astype(U32)
    record("mixed_dense.npz", batches, qx, qy, np.arange(64, dtype=U32))
    # 2. random 32-bit keys and values
    batches = []
    for op in ("set", "incr", "incr"):
        xs, ys, vs = safe_stream(rng, 6000, 50, 400, op, col0_rate=0.05, wide_keys=True)
        batches.append((op, xs, ys, vs))
    allx = np.concatenate([b[1] for b in batches]); ally = np.concatenate([b[2] for b in batches])
    pick = rng.integers(0, len(allx), 3000)
    qx = np.concatenate([allx[pick], rng.integers(0, 2**32, 500, dtype=np.uint64).astype(U32)])
    qy = np.concatenate([ally[pick], rng.integers(0, 2**32, 500, dtype=np.uint64).astype(U32)])
    record("wide_keys.npz", batches, qx, qy, np.unique(allx))
    # 3. co-occurrence build (examples/cf_recommender.c:35-47), three batches
    xs, ys = cf_stream(rng, 400, 120)
    batches = [("incr", a, b, np.ones(len(a), U32)) for a, b in zip(np.array_split(xs, 3), np.array_split(ys, 3))]
    record("cf_cooccurrence.npz", batches, xs[:5000], ys[:5000], np.arange(122, dtype=U32))
    # 4. the reference's own Java cases 1-7 as batches (src/java/test/TestSparseMatrix.java:22-130), 200 x 200
    g = 200
    i, n = np.meshgrid(np.arange(g, dtype=U32), np.arange(g, dtype=U32))
    ar = np.arange(g, dtype=U32)
    one = lambda x, y, v: (np.array([x], U32), np.array([y], U32), np.array([v], U32))
    batches = [("set", *one(42, 23, 17)), ("set", *one(4231, 2634, 0)), ("incr", *one(4231, 2634, 1)),
               ("set", *one(1231, 2634, 0)), ("incr", *one(1231, 2634, 1)), ("incr", *one(1231, 2634, 5)),
               ("set", i.ravel(), n.ravel(), np.full(g * g, 34, U32)),
               ("incr", ar, np.full(g, 42, U32), np.ones(g, U32)),
               ("incr", ar, np.full(g, 85, U32), np.ones(g, U32)),
               ("incr", ar, np.full(g, 83, U32), np.ones(g, U32))]
    qx = np.concatenate([i.ravel()[::7], np.array([42, 4231, 1231], U32)])
    qy = np.concatenate([n.ravel()[::7], np.array([23, 2634, 2634], U32)])
    record("java_cases.npz", batches, qx, qy, np.array([42, 83, 85, 0, 199, 4231, 1231, 9999], U32))


if __name__ == "__main__":
    main()
