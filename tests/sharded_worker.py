"""Worker for tests/test_sharded_cpu.py: one rank of a world_size-2 gloo group driving the
row-shard router (host logic) on the serial simulator library."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from libsmatrix_b200.sharded import TorchShardedSparseMatrix as ShardedSparseMatrix  # noqa: E402
from oracle import cpu  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    sim = sys.argv[1]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = ShardedSparseMatrix(rank, world, 0, _lib_path=sim)
    ref = cpu.CpuMatrix("port")
    rng = np.random.default_rng(123)                 # same stream on every rank
    n = 40000
    xs = rng.integers(0, 3000, n).astype(np.uint32) * np.uint32(2654435761)
    ys = rng.integers(0, 60, n).astype(np.uint32)
    vs = rng.integers(0, 1000, n).astype(np.uint32)
    ref.apply("incr", xs, ys, vs)                    # the checker sees the WHOLE batch
    ref.apply("decr", xs[::3], ys[::3], vs[::3] // 2)
    mine = slice(rank * n // world, (rank + 1) * n // world)
    t = lambda a: torch.from_numpy(a.view(np.int32).copy())
    m.incr_batch(t(xs[mine]), t(ys[mine]), t(vs[mine]))
    sx, sy, sv = xs[::3], ys[::3], vs[::3] // 2
    k = len(sx)
    part = slice(rank * k // world, (rank + 1) * k // world)
    m.decr_batch(t(sx[part]), t(sy[part]), t(sv[part]))
    # order-free stream (no column 0), routed in overlapped pieces by the helper thread
    m.PIPELINE_MIN, m.PIPELINE_PIECE = 1000, 1500
    fx = rng.integers(0, 3000, 9000).astype(np.uint32) * np.uint32(2654435761)
    fy = rng.integers(1, 60, 9000).astype(np.uint32)
    ref.apply("incr", fx, fy, np.ones(9000, np.uint32))
    fs = slice(rank * 9000 // world, (rank + 1) * 9000 // world)
    m.incr_batch(t(fx[fs]), t(fy[fs]), None, ordered=False)
    # the same kind of stream as HOST arrays: staged piece by piece (3 - 4 pieces here)
    m.PIPELINE_MIN, m.HOST_PIECE = 1 << 62, 1300
    hx = rng.integers(0, 3000, 9000).astype(np.uint32) * np.uint32(2654435761)
    hy = rng.integers(1, 60, 9000).astype(np.uint32)
    ref.apply("incr", hx, hy, np.ones(9000, np.uint32))
    m.incr_batch(hx[fs], hy[fs], None, ordered=False)
    xs = np.concatenate([xs, fx, hx]); ys = np.concatenate([ys, fy, hy])
    # set with duplicate keys across ranks: the last writer in GLOBAL input order must win
    gx = rng.integers(0, 50, 6000).astype(np.uint32) * np.uint32(2654435761)
    gy = rng.integers(1, 20, 6000).astype(np.uint32)
    gv = rng.integers(1, 2**32, 6000, dtype=np.uint64).astype(np.uint32)
    ref.apply("set", gx, gy, gv)
    gs = slice(rank * 6000 // world, (rank + 1) * 6000 // world)
    m.set_batch(t(gx[gs]), t(gy[gs]), t(gv[gs]))
    xs = np.concatenate([xs, gx]); ys = np.concatenate([ys, gy])
    # every rank asks for a different slice of queries and must get input-ordered answers
    qx = np.concatenate([xs[rank::5], rng.integers(0, 2**32, 300, dtype=np.uint64).astype(np.uint32)])
    qy = np.concatenate([ys[rank::5], rng.integers(0, 70, 300).astype(np.uint32)])
    got = m.get_batch(t(qx), t(qy)).numpy().view(np.uint32)
    assert (got == ref.get_many(qx, qy)).all(), f"rank {rank}: sharded get mismatch"
    assert (m.get_batch(qx, qy) == got).all(), f"rank {rank}: host-array get differs"   # piece-wise path
    rows = np.unique(xs)[rank::2]
    got = m.rowlen_batch(t(rows)).numpy().view(np.uint32)
    assert (got == ref.rowlen_many(rows)).all(), f"rank {rank}: sharded rowlen mismatch"
    # getrow across ranks: rows come back in input order, pairs compared sorted by column
    rows = np.concatenate([np.unique(xs)[rank::3], np.array([11, 12], np.uint32)])
    o1, p1 = m.getrow_batch(rows)
    o2, p2 = ref.getrow_many(rows)
    assert (o1 == o2).all() and (cpu.sort_rows(o1, p1) == cpu.sort_rows(o2, p2)).all(), f"rank {rank}: sharded getrow mismatch"
    try:
        m.cf_neighbors_batch
        raise SystemExit("data-path calls that are not sharded must be refused")
    except AttributeError:
        pass
    # the shard holds exactly the rows this rank owns
    owned = np.array([x for x in np.unique(xs) if m._lib.smatrix_b200_owner(int(x), world) == rank], np.uint32)
    assert m.stat("rows") == len(owned)
    tot = torch.tensor([m.stat("rows"), m.stat("nnz")])
    dist.all_reduce(tot)
    assert int(tot[0]) == len(np.unique(xs))
    o, p = ref.getrow_many(np.unique(xs))
    assert int(tot[1]) == len(p)
    m.close(); ref.close()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
