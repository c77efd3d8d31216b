"""measurement helper: cost of cudaMalloc/cudaFree by size, fresh vs. recycled (through the library's
own dev_alloc/dev_free = cudaMalloc/cudaFree)."""
import sys, time
sys.path.insert(0, ".")
from libsmatrix_b200 import SparseMatrix
m = SparseMatrix(device=0)
for rnd in range(2):
    for gib in (1, 2, 4, 8, 16, 32):
        n = gib << 30
        t0 = time.perf_counter(); p = m.dev_alloc(n); t1 = time.perf_counter()
        m.dev_free(p); t2 = time.perf_counter()
        print(f"round {rnd}: {gib:3d} GiB  malloc {1e3*(t1-t0):8.2f} ms ({1e3*(t1-t0)/gib:6.2f} ms/GiB)  free {1e3*(t2-t1):8.2f} ms")
ps = []
t0 = time.perf_counter()
for k in range(8): ps.append(m.dev_alloc(8 << 30))
print(f"8 x 8 GiB held together: {1e3*(time.perf_counter()-t0):.1f} ms")
for p in ps: m.dev_free(p)
m.close()
