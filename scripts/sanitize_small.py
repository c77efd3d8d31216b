"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the write
and read paths — growth through all three re-placement kernels (warp / block-per-row / grid-wide),
recycling of vacated buckets, directory pre-sizing and rehash, set resolution, the column-0 split of
partitioned chunks (128 slices + column-0 twins / 256 slices), point reads in input and in
directory-slice order, getrow (inline / warp / chunked), per-op return values, the snapshot export and a
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
ys1 = np.maximum(ys, 1)                                          # a chunk without column 0: 256 slices, no twins
m.incr_batch(xs, ys1, vs * np.uint32(2) + np.uint32(1))         # (no cell may end at 0: a reload drops zero-valued cells,
                                                                 #  src/smatrix.c:533-538, and the round trip below counts cells)
from libsmatrix_b200.matrix import DevPtr                        # device-array gets in directory-slice order
dq = [m.dev_alloc(4 * n) for _ in range(3)]
m.memcpy(dq[0], xs.ctypes.data, 4 * n); m.memcpy(dq[1], ys1.ctypes.data, 4 * n)
host = np.empty(n, np.uint32)
for mode in (2, 2 | 4 | 8, 2 | 16 | 32, 0):
    m.set_get_slices(mode)
    m.get_batch(DevPtr(dq[0], n), DevPtr(dq[1], n), DevPtr(dq[2], n))
    m.memcpy(host.ctypes.data, dq[2], 4 * n)
    assert mode == 2 or (host == first).all()
    first = host.copy()
m.set_get_slices(1)
assert m.stat("sliced_gets") == 3 * n and m.stat("wide_chunks") >= 1, (m.stat("sliced_gets"), m.stat("wide_chunks"))
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
