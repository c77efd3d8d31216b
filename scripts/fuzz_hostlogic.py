"""Randomised stress of the host logic on the serial simulator (tests/hostsim): random knob
settings (directory size, chunk size, partition threshold / slice size, pre-aggregation, recycling,
directory pre-sizing, tile re-placement with tiny tiles and short spill lists, arena mode, 256-slice chunks,
point reads in slice order with their measurement switches; a fifth of the runs on the 32-lane simulator) x random
op mixes (incr / decr / set, column-0 rates, key widths, batch sizes, *_batch_out) against the
checker.  Not part of the test suite; run it after touching smx_host.c:   python scripts/fuzz_hostlogic.py [runs] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hostsim import build as hb
from libsmatrix_b200 import SparseMatrix
import parity_suite as ps
from conftest import safe_stream

U32 = np.uint32
sim = hb.build()
sim_small = hb.build(defines=["-DMIG_TILE_LOG=3u"], suffix="_smalltiles")   # tiny re-placement tiles: cells spill all the time
sim32 = hb.build(defines=["-DSMX_SIM_WARP32"], suffix="_warp32")             # 32-lane lock-step warps (slow)
from libsmatrix_b200.matrix import DevPtr


def device_gets(m, qx, qy):
    """smatrix_get_batch on arrays in the simulator's device memory: the path that may order the queries by slice"""
    n = len(qx)
    ptrs = [m.dev_alloc(4 * n + 8) for _ in range(3)]
    m.memcpy(ptrs[0], qx.ctypes.data, 4 * n); m.memcpy(ptrs[1], qy.ctypes.data, 4 * n)
    m.get_batch(DevPtr(ptrs[0], n), DevPtr(ptrs[1], n), DevPtr(ptrs[2], n))
    out = np.empty(n, U32)
    m.memcpy(out.ctypes.data, ptrs[2], 4 * n)
    for p in ptrs:
        m.dev_free(p)
    return out
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else int(time.time())
print("seed0", seed0, flush=True)
for run in range(runs):
    rng = np.random.default_rng(seed0 + run)
    knobs = {"SMATRIX_DIR_LOG2": int(rng.integers(3, 12)), "SMATRIX_CHUNK": int(rng.choice([97, 777, 3000, 20000, 1 << 26])),
             "SMATRIX_PARTITION_MIN": int(rng.choice([16, 500, 1 << 20])), "SMATRIX_SLICE_LOG2": int(rng.integers(2, 8)),
             "SMATRIX_PARTS_LOG2": int(rng.integers(1, 9)), "SMATRIX_PREAGG": int(rng.integers(0, 2)),
             "SMATRIX_STAGE": int(rng.choice([1024, 5000, 1 << 23])), "SMATRIX_RECYCLE": int(rng.integers(0, 2)),
             "SMATRIX_PRESIZE": int(rng.integers(0, 2)), "SMATRIX_MIGRATE_TILES": int(rng.integers(0, 2)),
             "SMATRIX_SPILL_CAP": int(rng.choice([1, 64, 1 << 16])), "SMATRIX_ARENA_GIB": int(rng.choice([0, 0, 1])),
             "SMATRIX_WIDE_SLICES": int(rng.integers(0, 2)), "SMATRIX_GET_SLICE_MIN": int(rng.choice([2, 100, 1 << 22])),
             "SMATRIX_GET_SLICES": int(rng.choice([0, 1, 1, 2, 2 | 4, 2 | 8, 2 | 16, 2 | 32, 1 | 4 | 8 | 16 | 32]))}
    r = rng.random()
    lib = sim32 if r < 0.2 else (sim_small if r < 0.6 else sim)
    for k, v in knobs.items():
        os.environ[k] = str(v)
    m, ref = SparseMatrix(_lib_path=lib), ps.checker()
    n_rows, n_cols = int(rng.choice([3, 40, 400, 5000])), int(rng.choice([2, 30, 300, 3000, 40000]))
    wide = bool(rng.integers(0, 2)); col0 = float(rng.choice([0.0, 0.02, 0.3]))
    desc = f"run {run} seed {seed0 + run} knobs {knobs} rows {n_rows} cols {n_cols} wide {wide} col0 {col0}"
    try:
        seen_x, seen_y = [], []
        for b in range(int(rng.integers(1, 6))):
            op = str(rng.choice(["incr", "incr", "set", "decr"]))
            n = int(rng.choice([1, 50, 3000, 25000]))
            xs, ys, vs = safe_stream(rng, n, n_rows, n_cols, op, col0_rate=col0, wide_keys=wide,
                                     max_val=int(rng.choice([1, 3, 2**32 - 1])))
            if op == "decr":
                ys = np.where(ys == 0, ys.max(), ys).astype(U32)
                if (ys == 0).all():
                    continue
            if rng.random() < 0.3:      # per-op return values (N1)
                got = np.asarray(getattr(m, op + "_batch_out")(xs, ys, vs))
                want = ref.apply(op, xs, ys, vs, want_out=True)
                assert (got == want).all(), f"{op}_batch_out: {int((got != want).sum())} mismatches"
            else:
                ps.apply_both(m, ref, op, xs, ys, vs if rng.random() < 0.8 or op == "set" else vs)
            seen_x.append(xs); seen_y.append(ys)
            ax, ay = np.concatenate(seen_x), np.concatenate(seen_y)
            qx = np.concatenate([ax[-4000:], rng.integers(0, 2**32, 200, dtype=np.uint64).astype(U32)])
            qy = np.concatenate([ay[-4000:], rng.integers(0, 2**32, 200, dtype=np.uint64).astype(U32)])
            ps.compare(m, ref, np.concatenate([np.unique(ax), qx[-10:]]), qx, qy)
            got = device_gets(m, qx, qy)
            assert (got == ref.get_many(qx, qy)).all(), f"device-array get: {int((got != ref.get_many(qx, qy)).sum())} mismatches"
        m.close(); ref.close()
    except Exception as e:      # noqa: BLE001
        print("FAILED:", desc, "->", repr(e), flush=True)
        raise
    if run % 20 == 0:
        print("ok", desc, flush=True)
print("all", runs, "runs ok")
