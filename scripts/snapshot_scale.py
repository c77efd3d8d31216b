"""K9 at scale: time smatrix_close (snapshot save: row blocks laid out on the device in the reference's
format) and smatrix_open (load) of a config-2-shaped table.  usage: snapshot_scale.py [fraction] [dir]"""
import json, os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libsmatrix_b200 import SparseMatrix

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
where = sys.argv[2] if len(sys.argv) > 2 else "/dev/shm"
rows, ops, B = int(13_000_000 * frac), int(2_000_000_000 * frac), 1 << 26
path = os.path.join(where, "smx_snapshot_scale.smx")
if os.path.exists(path):
    os.remove(path)
free = shutil.disk_usage(where).free
os.environ["SMATRIX_ARENA_GIB"] = str(max(4, int(56 * frac) + 2))
dev = torch.device("cuda", 0)
m = SparseMatrix(file_path=path, device=0)
xs, ys = torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
t0 = time.perf_counter()
for first in range(0, ops, B):
    cnt = min(B, ops - first)
    m.gen_c2_ops(2, first, cnt, rows, 256, xs.data_ptr(), ys.data_ptr())
    m.incr_batch(xs[:cnt], ys[:cnt], None)
build_s = time.perf_counter() - t0
before = {k: m.stat(k) for k in ("rows", "nnz", "value_sum")}
t0 = time.perf_counter(); m.close(); save_s = time.perf_counter() - t0
size = os.path.getsize(path)
t0 = time.perf_counter(); m2 = SparseMatrix(file_path=path, device=0); load_s = time.perf_counter() - t0
after = {k: m2.stat(k) for k in ("rows", "nnz", "value_sum")}
os.environ.pop("SMATRIX_ARENA_GIB", None)
m2._lib.smatrix_b200_snapshot  # exported
m2.filename = None
# do not write the file again on close: drop it first
os.remove(path)
print(json.dumps({"fraction": frac, "rows": rows, "ops": ops, "dir": where, "free_bytes_before": free, "file_bytes": size,
                  "build_s": round(build_s, 3), "save_s": round(save_s, 3), "save_gbs": round(size / save_s / 1e9, 3),
                  "load_s": round(load_s, 3), "load_gbs": round(size / load_s / 1e9, 3), "before": before, "after": after,
                  "identical": before == after}))
os._exit(0)   # skip the close-time snapshot of the reloaded handle
