"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the write
and read paths, growth, directory rehash, set resolution, partitioned chunks, getrow."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SMATRIX_DIR_LOG2", "8")
os.environ.setdefault("SMATRIX_PARTITION_MIN", "4096")
os.environ.setdefault("SMATRIX_SLICE_LOG2", "6")
import numpy as np
from libsmatrix_b200 import SparseMatrix
rng = np.random.default_rng(0)
m = SparseMatrix(device=0)
n = 60000
xs = rng.integers(0, 3000, n).astype(np.uint32) * np.uint32(2654435761)
ys = rng.integers(0, 150, n).astype(np.uint32)
vs = rng.integers(1, 1000, n).astype(np.uint32)
m.incr_batch(xs, ys, vs)
m.set_batch(xs[:20000], ys[:20000], vs[:20000])
m.decr_batch(xs[:5000], np.maximum(ys[:5000], 1), vs[:5000])
big = np.arange(1, 30001, dtype=np.uint32)
m.incr_batch(np.full(30000, 7, np.uint32), big, None)          # one big row (grid-wide re-placement)
out = m.get_batch(xs, ys)
rl = m.rowlen_batch(np.unique(xs))
o, p = m.getrow_batch(np.concatenate([np.unique(xs)[:500], np.array([7], np.uint32)]))
print("ok", int(out.sum()) & 0xffff, int(rl.sum()), len(p), m.stat("rows"), m.stat("nnz"))
m.close()
