"""measurement helper: per-phase times of the sharded router on N GPUs (torchrun)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libsmatrix_b200 import SparseMatrix
from libsmatrix_b200.sharded import ShardedSparseMatrix

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
os.environ["SMATRIX_ARENA_GIB"] = "20"
m = ShardedSparseMatrix(rank, world, local)
os.environ.pop("SMATRIX_ARENA_GIB")
B = 1 << 26; rows = 13_000_000 * world
xs = torch.empty(B, dtype=torch.int32, device=dev); ys = torch.empty_like(xs)
def sync(): torch.cuda.synchronize(); m.local.sync(); m.router.sync()
T = {}
def tick(name, t0):
    sync(); T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
for k in range(6):
    m.gen_c2_ops(2, (k * world + rank) * B, B, rows, 256, xs.data_ptr(), ys.data_ptr())
    m.incr_batch(xs, ys, None, ordered=False)
for w in range(3):   # warm every buffer of every path (three slot generations)
    m.get_batch(xs, ys); m.incr_batch(xs, ys, None, ordered=False)
sync(); dist.barrier()
T.clear(); reps = 4
for k in range(6, 6 + reps + 1):
    if k == 7: T.clear()      # first timed repetition still allocates the profile-only buffers
    m.gen_c2_ops(2, (k * world + rank) * B, B, rows, 256, xs.data_ptr(), ys.data_ptr()); sync(); dist.barrier()
    t0 = time.perf_counter(); send, oxs, oys, ovs, osrc, opos = m._partition(xs, ys, None, False); tick("incr.partition", t0)
    t0 = time.perf_counter(); recv = m._exchange_counts(send); tick("incr.counts", t0)
    t0 = time.perf_counter(); rx = m._a2a(oxs, send, recv, "rxp"); ry = m._a2a(oys, send, recv, "ryp"); tick("incr.a2a", t0)
    t0 = time.perf_counter(); m.local.incr_batch(rx, ry, None); tick("incr.local", t0)
    t0 = time.perf_counter(); m.incr_batch(xs, ys, None, ordered=False); tick("incr.whole_pipelined", t0)
    m.PIPELINE_MIN = 1 << 40
    t0 = time.perf_counter(); m.incr_batch(xs, ys, None, ordered=False); tick("incr.whole_unpipelined", t0)
    m.PIPELINE_MIN = 1 << 23
    # get
    t0 = time.perf_counter(); send, oxs, oys, ovs, osrc, opos = m._partition(xs, ys, None, False, True); tick("get.partition", t0)
    t0 = time.perf_counter(); recv = m._exchange_counts(send); tick("get.counts", t0)
    t0 = time.perf_counter(); rx = m._a2a(oxs, send, recv, "rxp"); ry = m._a2a(oys, send, recv, "ryp"); tick("get.a2a", t0)
    t0 = time.perf_counter(); ans = m.local.get_batch(rx, ry); tick("get.local", t0)
    t0 = time.perf_counter(); back = m._a2a(ans, recv, send); tick("get.a2a_back", t0)
    t0 = time.perf_counter(); out = m._unpermute(back, opos); tick("get.gather", t0)
    t0 = time.perf_counter(); out2 = m.get_batch(xs, ys); tick("get.whole", t0)
    m._use_p2p = False
    t0 = time.perf_counter(); out3 = m.get_batch(xs, ys); tick("get.whole_nccl", t0)
    t0 = time.perf_counter(); m.incr_batch(xs, ys, None, ordered=False); tick("incr.whole_nccl_pipelined", t0)
    m._use_p2p = True
    assert bool((out2 == out3).all())
if rank == 0:
    for k, v in T.items(): print(f"{k:28s} {v / reps:8.3f} ms per 2^26-op batch per rank")
m.close(); dist.destroy_process_group()
