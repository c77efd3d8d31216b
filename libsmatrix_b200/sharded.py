"""ShardedSparseMatrix — one matrix row-sharded over the GPUs of one box.

    owner(x) = mix_owner(x) mod world        (smatrix_b200_owner; independent of the directory hash)

One process per GPU (torch.distributed, backend nccl over NVLink/NVSwitch).  Every rank holds a
complete private table for the rows it owns.  A batch call is collective: each rank passes ITS
slice of the batch; the slice is bucketed by owner on the device (K8, smatrix_b200_partition),
the per-owner counts are exchanged, one all-to-all-v moves the (x, y, value) triples, and each
rank updates only its own shard.  Reads add the reverse all-to-all-v and un-permute the answers
into input order.

Input order.  The collective batch is the concatenation of the ranks' slices in rank order.
Routing permutes ops, and two results depend on order: which duplicate `set` wins, and the
history-dependent rowlen when column 0 turns non-zero in the same batch as new columns (SURVEY.md
Q1).  With `ordered=True` (default) every op travels with its global index (16 B/op instead of
12) and the owner applies the batch through smatrix_b200_apply_ordered, which is bit-exact with
the sequential reference.  `ordered=False` skips the index for incr/decr streams that never
write column 0 (BASELINE configs 2 and 5): values never depend on order (addition mod 2^32
commutes), so the result is identical.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from .matrix import SparseMatrix


class ShardedSparseMatrix:
    def __init__(self, rank: int, world: int, device: int = 0, group=None, _lib_path: str | None = None):
        self.rank, self.world, self.group = rank, world, group
        self.local = SparseMatrix(device=device, _lib_path=_lib_path)
        arena = os.environ.pop("SMATRIX_ARENA_GIB", None)               # the router holds no data:
        try:                                                             # it must not reserve an arena
            self.router = SparseMatrix(device=device, _lib_path=_lib_path)   # K8 only
        finally:
            if arena is not None:
                os.environ["SMATRIX_ARENA_GIB"] = arena
        self._lib = self.local._lib
        self._pool = None
        self._bufs: dict[str, torch.Tensor] = {}
        self._gen = 0
        self._cuda = _lib_path is None
        self.dev = torch.device("cuda", device) if self._cuda else torch.device("cpu")

    # ------------------------------------------------------------------ plumbing
    def _buf(self, n, like=None):
        return torch.empty(n, dtype=torch.int32, device=self.dev)

    def _slot(self, name: str, n: int):
        """Grow-only reusable device buffer (no allocator traffic in steady state)."""
        b = self._bufs.get(name)
        if b is None or b.numel() < n:
            b = torch.empty(max(n + n // 8, 1024), dtype=torch.int32, device=self.dev)
            self._bufs[name] = b
        return b[:n]

    def _sync_torch(self):
        if self._cuda:
            torch.cuda.current_stream(self.dev).synchronize()

    def _partition(self, xs, ys, vals, want_src: bool, want_pos: bool = False):
        n = xs.numel()
        g = self._gen = (self._gen + 1) % 3      # three generations: routed / in flight / being applied
        oxs = self._slot(f"ox{g}", n)
        oys = self._slot(f"oy{g}", n) if ys is not None else None
        ovs = self._slot(f"ov{g}", n) if vals is not None else None
        osrc = self._slot(f"os{g}", n) if want_src else None
        opos = self._slot(f"op{g}", n) if want_pos else None
        counts = np.zeros(self.world, dtype=np.uint64)
        p = lambda t: t.data_ptr() if t is not None else None
        self._sync_torch()   # inputs may have been produced on torch's stream
        # the router has its own handle (own stream + mutex) so that routing the next sub-batch can
        # overlap the update of the current one
        self._lib.smatrix_b200_partition2(self.router._handle(), p(xs), p(ys), p(vals), n, self.world,
                                          counts.ctypes.data, p(oxs), p(oys), p(ovs), p(osrc), p(opos))
        return [int(c) for c in counts], oxs, oys, ovs, osrc, opos

    def _exchange_counts(self, send):
        t = torch.tensor(send, dtype=torch.int64, device=self.dev)
        r = torch.empty_like(t)
        dist.all_to_all_single(r, t, group=self.group)
        return [int(v) for v in r.tolist()]

    def _a2a(self, t, send, recv, name=None):
        out = self._slot(name, sum(recv)) if name else self._buf(sum(recv))
        dist.all_to_all_single(out, t, output_split_sizes=recv, input_split_sizes=send, group=self.group)
        return out

    def _route(self, xs, ys, vals, want_src=False, ordered=False, want_pos=False):
        send, oxs, oys, ovs, osrc, opos = self._partition(xs, ys, vals, want_src or ordered, want_pos)
        recv = self._exchange_counts(send)
        g = self._gen
        rx = self._a2a(oxs, send, recv, f"rx{g}")
        ry = self._a2a(oys, send, recv, f"ry{g}") if oys is not None else None
        rv = self._a2a(ovs, send, recv, f"rv{g}") if ovs is not None else None
        rord = None
        if ordered:   # global index = (ops in the slices of lower ranks) + index inside my slice
            sizes = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
            sizes[self.rank] = xs.numel()
            dist.all_reduce(sizes, group=self.group)
            if int(sizes.sum()) >= 2**32 - 1:
                raise ValueError("ordered collective batches are limited to 2^32 - 2 ops")
            base = int(sizes[: self.rank].sum())
            gidx = (osrc.long() + base).to(torch.int32)   # wraps into uint32 bit pattern
            rord = self._a2a(gidx, send, recv)
        self._sync_torch()   # the library runs on its own stream
        return send, recv, rx, ry, rv, (opos if want_pos else osrc), rord

    PIPELINE_MIN = 1 << 23      # order-free batches at least this big are routed in overlapped pieces
    PIPELINE_PIECE = 1 << 24

    def _write(self, op: int, xs, ys, vals, ordered: bool):
        if ordered:
            _, _, rx, ry, rv, _, rord = self._route(xs, ys, vals, ordered=True)
            if rx.numel():
                p = lambda t: t.data_ptr() if t is not None else None
                self._lib.smatrix_b200_apply_ordered(self.local._handle(), op, p(rx), p(ry), p(rv), p(rord),
                                                     rx.numel())
            return
        apply = (self.local.incr_batch, self.local.decr_batch)[op]
        n = xs.numel()
        # every rank must cut the same number of pieces (the exchanges are collective)
        nmax = torch.tensor([n], dtype=torch.int64, device=self.dev)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=self.group)
        pieces = max(1, -(-int(nmax) // self.PIPELINE_PIECE)) if int(nmax) >= self.PIPELINE_MIN else 1
        if pieces == 1:
            _, _, rx, ry, rv, _, _ = self._route(xs, ys, vals)
            if rx.numel():
                apply(rx, ry, rv)
            return
        # route piece j+1 (router handle + NCCL, helper thread) while piece j updates the shard
        import concurrent.futures as cf
        if self._pool is None:
            self._pool = cf.ThreadPoolExecutor(max_workers=1)
        cut = lambda t, j: None if t is None else t[j * n // pieces:(j + 1) * n // pieces]

        def route(j):
            if self._cuda:
                torch.cuda.set_device(self.dev)
            return self._route(cut(xs, j), cut(ys, j), cut(vals, j))
        fut = self._pool.submit(route, 0)
        for j in range(pieces):
            _, _, rx, ry, rv, _, _ = fut.result()
            if j + 1 < pieces:
                fut = self._pool.submit(route, j + 1)
            if rx.numel():
                apply(rx, ry, rv)

    # ------------------------------------------------------------------ collective batch API
    def incr_batch(self, xs, ys, vals=None, ordered: bool = True):
        self._write(0, xs, ys, vals, ordered)

    def decr_batch(self, xs, ys, vals=None, ordered: bool = True):
        self._write(1, xs, ys, vals, ordered)

    def set_batch(self, xs, ys, vals=None):
        self._write(2, xs, ys, vals, True)     # last writer in GLOBAL input order wins

    def get_batch(self, xs, ys, out=None):
        send, recv, rx, ry, _, opos, _ = self._route(xs, ys, None, want_pos=True)
        ans = self.local.get_batch(rx, ry) if rx.numel() else self._buf(0)
        self._sync_torch()
        back = self._a2a(ans, recv, send)           # answers travel the reverse way
        return self._unpermute(back, opos, out)

    def _unpermute(self, back, opos, out=None):
        """out[i] = back[opos[i]] — answers arrive in routed order."""
        if out is None:
            out = self._buf(opos.numel())
        self._sync_torch()
        self._lib.smatrix_b200_gather(self.router._handle(), out.data_ptr(), back.data_ptr(),
                                      opos.data_ptr(), opos.numel())
        return out

    def rowlen_batch(self, xs):
        send, recv, rx, _, _, opos, _ = self._route(xs, None, None, want_pos=True)
        ans = self.local.rowlen_batch(rx) if rx.numel() else self._buf(0)
        self._sync_torch()
        back = self._a2a(ans, recv, send)
        return self._unpermute(back, opos)

    # ------------------------------------------------------------------ local controls
    def stat(self, name):
        return self.local.stat(name)

    def __getattr__(self, name):
        # timer_start / timer_stop_ms / set_kernel_timing / gen_c2_* / probe_* / sync: the local shard's
        return getattr(self.local, name)

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        self._bufs: dict[str, torch.Tensor] = {}
        self._gen = 0
        self.router.close()
        self.local.close()
