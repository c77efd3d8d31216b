"""BASELINE config 3 shape at scale: Zipf(1.1) item baskets of 8, all-pairs incr + column-0 totals
(examples/cf_recommender.c:35-47).  Times the build through the C-ABI (device-resident batches) and
checks a prefix against the reference."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from libsmatrix_b200 import SparseMatrix
from oracle import cpu
from parity_suite import cf_stream

n_items = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
n_baskets = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
rng = np.random.default_rng(4)
t0 = time.time(); xs, ys = cf_stream(rng, n_baskets, n_items); print(f"stream: {len(xs)/1e6:.1f} M ops in {time.time()-t0:.1f}s", flush=True)
dev = torch.device("cuda", 0)
m = SparseMatrix(device=0)
B = 1 << 25
dx = torch.from_numpy(xs.view(np.int32)).to(dev); dy = torch.from_numpy(ys.view(np.int32)).to(dev)
m.set_kernel_timing(True)
t0 = time.perf_counter()
steps = []
for off in range(0, len(xs), B):
    t1 = time.perf_counter(); m.incr_batch(dx[off:off + B], dy[off:off + B], None); steps.append(round((time.perf_counter() - t1) * 1e3, 1))
secs = time.perf_counter() - t0
print(f"build {len(xs)/secs/1e6:.0f} Mops/s ({secs*1e3:.0f} ms) rows={m.stat('rows')} nnz={m.stat('nnz')} slab={m.stat('slab_bytes')/1e9:.2f} GB rounds={m.stat('rounds')} grows={m.stat('row_grows')} launches={m.stat('launches')}")
print("step ms", steps)
print({k: round(m.stat('ns_' + k) / 1e6, 1) for k in ("partition", "upsert", "grow_plan", "slab", "migrate", "dir")}, "kernel ms", round(m.stat("kernel_ns") / 1e6, 1))
# hottest rows
top = np.arange(1, 6, dtype=np.uint32)
print("rowlen of items 1..5:", m.rowlen_batch(top), "c0:", m.get_batch(top, np.zeros(5, np.uint32)))
t1 = time.perf_counter(); o, p = m.getrow_batch(top); print(f"getrow of the 5 hottest rows: {len(p)} pairs in {(time.perf_counter()-t1)*1e3:.1f} ms")
q = torch.from_numpy(xs[:B].view(np.int32)).to(dev); qy = torch.from_numpy(ys[:B].view(np.int32)).to(dev)
t1 = time.perf_counter(); out = m.get_batch(q, qy); torch.cuda.synchronize(); print(f"get {len(q)/(time.perf_counter()-t1)/1e6:.0f} Mops/s")
# parity on a prefix (the reference manages ~0.2 Mops/s on this workload)
k = 1_500_000
m2, ref = SparseMatrix(device=0), cpu.CpuMatrix("reference" if cpu.have_reference() else "port")
m2.incr_batch(xs[:k], ys[:k], None); t1 = time.time(); ref.apply("incr", xs[:k], ys[:k], np.ones(k, np.uint32)); print(f"reference: {k/(time.time()-t1)/1e6:.3f} Mops/s on the prefix")
rows = np.unique(xs[:k])[:20000]
assert (np.asarray(m2.get_batch(xs[:k], ys[:k])) == ref.get_many(xs[:k], ys[:k])).all()
assert (np.asarray(m2.rowlen_batch(rows)) == ref.rowlen_many(rows)).all()
o1, p1 = m2.getrow_batch(rows[:2000]); o2, p2 = ref.getrow_many(rows[:2000])
assert (o1 == o2).all() and (cpu.sort_rows(o1, p1) == cpu.sort_rows(o2, p2)).all()
print("prefix parity ok")
