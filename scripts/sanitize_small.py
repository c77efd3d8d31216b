"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the write
and read paths — growth through all three re-placement kernels (warp / block-per-row / grid-wide),
recycling of vacated buckets, directory pre-sizing and rehash, set resolution, the column-0 split of
partitioned chunks, getrow (inline / warp / chunked), per-op return values, the snapshot export and a
two-rank routed batch on one GPU."""
import os, sys, tempfile, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SMATRIX_DIR_LOG2", "8")
os.environ.setdefault("SMATRIX_PARTITION_MIN", "4096")
os.environ.setdefault("SMATRIX_SLICE_LOG2", "6")
os.environ.setdefault("SMATRIX_SHARD_TIMEOUT", "600")
import numpy as np
from libsmatrix_b200 import SparseMatrix
from libsmatrix_b200.sharded import ShardedSparseMatrix
rng = np.random.default_rng(0)
path = os.path.join(tempfile.mkdtemp(), "sanitize.smx")
m = SparseMatrix(file_path=path, device=0)
n = 60000
xs = rng.integers(0, 3000, n).astype(np.uint32) * np.uint32(2654435761)
ys = rng.integers(0, 150, n).astype(np.uint32)
vs = rng.integers(1, 1000, n).astype(np.uint32)
m.incr_batch(xs, ys, vs)
m.set_batch(xs[:20000], ys[:20000], vs[:20000])
m.decr_batch(xs[:5000], np.maximum(ys[:5000], 1), vs[:5000])
ret = m.incr_batch_out(xs[:8000], ys[:8000], vs[:8000])
for k in (700, 3000, 30000):                                     # mid rows and one big row, grown in steps
    cols = np.arange(1, k + 1, dtype=np.uint32)
    m.incr_batch(np.full(k, 7, np.uint32), cols, None)
    m.incr_batch(np.full(k // 2, 9, np.uint32), cols[: k // 2] * np.uint32(3), None)
out = m.get_batch(xs, ys)
rl = m.rowlen_batch(np.unique(xs))
o, p = m.getrow_batch(np.concatenate([np.unique(xs)[:500], np.array([7, 9, 12345], np.uint32)]))
print("ok", int(out.sum()) & 0xffff, int(rl.sum()), len(p), m.stat("rows"), m.stat("nnz"), m.stat("recycled"), int(ret.sum()) & 0xff)
before = (m.stat("rows"), m.stat("nnz"), m.stat("value_sum"))
m.close()                                                        # snapshot export (k_snap_*)
m = SparseMatrix(file_path=path, device=0)                       # and load
assert (m.stat("rows"), m.stat("nnz"), m.stat("value_sum")) == before, "snapshot round trip differs"
m.close()

def rank_main(r, errs):                                          # the C router, two ranks (threads) on one GPU
    try:
        sm = ShardedSparseMatrix(r, 2, 0, name=f"smx_sanitize_{os.getpid()}")
        sl = slice(r * n // 2, (r + 1) * n // 2)
        sm.incr_batch(xs[sl], ys[sl], vs[sl])
        sm.set_batch(xs[sl][:3000], ys[sl][:3000], vs[sl][:3000])
        sm.get_batch(xs[sl], ys[sl]); sm.rowlen_batch(np.unique(xs)[r::2])
        sm.getrow_batch(np.unique(xs)[r::2][:400])
        sm.close()
    except BaseException as e:
        errs.append(repr(e))
errs = []
ts = [threading.Thread(target=rank_main, args=(r, errs)) for r in range(2)]
[t.start() for t in ts]; [t.join() for t in ts]
assert not errs, errs
print("sanitize workload done")
