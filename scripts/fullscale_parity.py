"""Full-scale parity (SURVEY.md 8d "parity at scale"): BASELINE config 2 in full — 2 013 265 920
uniform incr ops over 13 M rows -> 1.51 B nnz — built by the UNMODIFIED reference on the host CPU
(oracle/_ref, single thread: its fastest setting) and by the CUDA library, then compared:
  * per-row digests (rowlen, #pairs, sum col, sum val, sum col*val) for ALL 13 M rows,
  * N_GETS point gets of the C2 query stream (50 % hits).
Writes profiles/r2_fullscale_parity.json.  ~8 minutes of CPU time, ~35 GB of host RAM."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from libsmatrix_b200 import SparseMatrix
from oracle import cpu

ROWS, YCOLS, SEED, SEED_GET = 13_000_000, 256, 2, 3
B = 1 << 26
K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
N_OPS = K * B
N_GETS = 50_000_000
cpu.build(ref=True)
assert cpu.have_reference(), "needs oracle/_ref/libsmatrix_ref.so"
report = {"workload": f"C2: {N_OPS} incr ops, {ROWS} rows x {YCOLS} columns, seed {SEED}"}

# ---- GPU build
os.environ["SMATRIX_ARENA_GIB"] = "52"
dev = torch.device("cuda", 0)
m = SparseMatrix(device=0)
dx = torch.empty(B, dtype=torch.int32, device=dev); dy = torch.empty_like(dx)
t0 = time.time()
for k in range(K):
    m.gen_c2_ops(SEED, k * B, B, ROWS, YCOLS, dx.data_ptr(), dy.data_ptr())
    m.incr_batch(dx, dy, None)
report["gpu_build_s"] = round(time.time() - t0, 2)
report["gpu_nnz"], report["gpu_rows"], report["gpu_value_sum"] = m.stat("nnz"), m.stat("rows"), m.stat("value_sum")
print("gpu built", report, flush=True)

# ---- reference build (pthread harness with 1 thread = plain sequential loop in C)
ref = cpu.CpuMatrix("reference")
t0 = time.time()
secs = ref.bench_c2_incr(1, SEED, 0, N_OPS, ROWS, YCOLS)
report["reference_build_s"] = round(secs, 1); report["reference_mops"] = round(N_OPS / secs / 1e6, 3)
print("reference built", report["reference_build_s"], "s", flush=True)

# ---- gets
qx, qy = cpu.gen_c2_queries(SEED_GET, SEED, 0, N_GETS, N_OPS, ROWS, YCOLS)
t0 = time.time(); want = ref.get_many(qx, qy); report["reference_get_mops"] = round(N_GETS / (time.time() - t0) / 1e6, 2)
got = np.asarray(m.get_batch(qx, qy))
report["gets_compared"] = N_GETS
report["get_mismatches"] = int((got != want).sum())
report["get_hits"] = int((want != 0).sum())
print("gets", report["get_mismatches"], "mismatches of", N_GETS, flush=True)

# ---- per-row digests for all rows
ids = (np.arange(ROWS, dtype=np.uint32) * np.uint32(2654435761))
d = cpu.driver()
mismatch_rows, checked, pairs_total = 0, 0, 0
STEP = 1 << 20
t0 = time.time()
for lo in range(0, ROWS, STEP):
    xs = np.ascontiguousarray(ids[lo:lo + STEP]); n = len(xs)
    refd = np.zeros((n, 5), dtype=np.uint64)
    d.drv_row_digests(ref.fnptr("rowlen"), ref.fnptr("getrow"), ref.h, xs.ctypes.data_as(C.POINTER(C.c_uint32)),
                      C.c_size_t(n), refd.ctypes.data_as(C.POINTER(C.c_uint64)))
    tx = torch.from_numpy(xs.view(np.int32)).to(dev)
    rl = m.rowlen_batch(tx).to(torch.int64) & 0xFFFFFFFF
    offs = torch.empty(n + 1, dtype=torch.int64, device=dev)
    total = int(m._lib.smatrix_getrow_batch(m._handle(), tx.data_ptr(), n, offs.data_ptr(), None, 0))
    pairs = torch.empty(2 * total, dtype=torch.int32, device=dev)
    m._lib.smatrix_getrow_batch(m._handle(), tx.data_ptr(), n, offs.data_ptr(), pairs.data_ptr(), total)
    col = pairs[0::2].to(torch.int64) & 0xFFFFFFFF; val = pairs[1::2].to(torch.int64) & 0xFFFFFFFF
    def seg(v):
        c = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(v, 0)])
        return c[offs[1:]] - c[offs[:-1]]
    gd = torch.stack([rl, offs[1:] - offs[:-1], seg(col), seg(val), seg(col * val)], dim=1).cpu().numpy().view(np.uint64)
    bad = (gd != refd).any(axis=1)
    mismatch_rows += int(bad.sum()); checked += n; pairs_total += total
    del pairs, col, val
print("digests done in", round(time.time() - t0, 1), "s", flush=True)
report.update({"rows_checked": checked, "row_digest_mismatches": mismatch_rows, "pairs_compared": pairs_total,
               "digest": "rowlen, #pairs, sum(col), sum(val), sum(col*val) mod 2^64 per row",
               "ok": mismatch_rows == 0 and report["get_mismatches"] == 0 and pairs_total == report["gpu_nnz"]})
print(json.dumps(report))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(ROOT, "gpurun_out", "r2_fullscale_parity.json"), "w"), indent=1)
m.close(); ref.close()
