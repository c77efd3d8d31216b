"""measurement helper: host-pointer (pinned) batches through the C-ABI vs staging size; raw H2D/D2H rate."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libsmatrix_b200 import SparseMatrix
dev = torch.device("cuda", 0); B = 1 << 26
h = torch.empty(B, dtype=torch.int32, pin_memory=True); d = torch.empty(B, dtype=torch.int32, device=dev)
for _ in range(2): d.copy_(h, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); print(f"H2D 256 MiB pinned: {0.268435456/(time.perf_counter()-t0):.1f} GB/s")
t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); print(f"D2H 256 MiB pinned: {0.268435456/(time.perf_counter()-t0):.1f} GB/s")
stage = os.environ.get("SMATRIX_STAGE", "default")
os.environ["SMATRIX_ARENA_GIB"] = "24"
m = SparseMatrix(device=0)
dx = torch.empty(B, dtype=torch.int32, device=dev); dy = torch.empty_like(dx)
hx = torch.empty(B, dtype=torch.int32, pin_memory=True); hy = torch.empty(B, dtype=torch.int32, pin_memory=True); ho = torch.empty(B, dtype=torch.int32, pin_memory=True)
rows = 13_000_000
ts = []
for k in range(14):
    m.gen_c2_ops(2, k * B, B, rows, 256, dx.data_ptr(), dy.data_ptr()); hx.copy_(dx); hy.copy_(dy); torch.cuda.synchronize()
    t0 = time.perf_counter(); m.incr_batch(hx, hy, None); ts.append((time.perf_counter() - t0) * 1e3)
print(f"stage={stage}: incr host-pointer ms per 2^26 ops:", [round(t, 1) for t in ts], f"-> steady {B/ (sum(ts[10:])/4) / 1e6 * 1e3:.0f} Mops/s")
tg = []
for k in range(4):
    t0 = time.perf_counter(); m.get_batch(hx, hy, ho); tg.append((time.perf_counter() - t0) * 1e3)
print(f"stage={stage}: get host-pointer ms per 2^26:", [round(t, 1) for t in tg], f"-> {B / (sum(tg[1:]) / 3) / 1e6 * 1e3:.0f} Mops/s")
m.close()
