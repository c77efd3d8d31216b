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

import numpy as np
import torch
import torch.distributed as dist

from .matrix import SparseMatrix


class ShardedSparseMatrix:
    def __init__(self, rank: int, world: int, device: int = 0, group=None, _lib_path: str | None = None):
        self.rank, self.world, self.group = rank, world, group
        self.local = SparseMatrix(device=device, _lib_path=_lib_path)
        self._lib = self.local._lib
        self._cuda = _lib_path is None
        self.dev = torch.device("cuda", device) if self._cuda else torch.device("cpu")

    # ------------------------------------------------------------------ plumbing
    def _buf(self, n, like=None):
        return torch.empty(n, dtype=torch.int32, device=self.dev)

    def _sync_torch(self):
        if self._cuda:
            torch.cuda.current_stream(self.dev).synchronize()

    def _partition(self, xs, ys, vals, want_src: bool):
        n = xs.numel()
        oxs, oys = self._buf(n), self._buf(n) if ys is not None else None
        ovs = self._buf(n) if vals is not None else None
        osrc = self._buf(n) if want_src else None
        counts = np.zeros(self.world, dtype=np.uint64)
        p = lambda t: t.data_ptr() if t is not None else None
        self._sync_torch()   # inputs may have been produced on torch's stream
        self._lib.smatrix_b200_partition(self.local._handle(), p(xs), p(ys), p(vals), n, self.world,
                                         counts.ctypes.data, p(oxs), p(oys), p(ovs), p(osrc))
        return [int(c) for c in counts], oxs, oys, ovs, osrc

    def _exchange_counts(self, send):
        t = torch.tensor(send, dtype=torch.int64, device=self.dev)
        r = torch.empty_like(t)
        dist.all_to_all_single(r, t, group=self.group)
        return [int(v) for v in r.tolist()]

    def _a2a(self, t, send, recv):
        out = self._buf(sum(recv))
        dist.all_to_all_single(out, t, output_split_sizes=recv, input_split_sizes=send, group=self.group)
        return out

    def _route(self, xs, ys, vals, want_src=False, ordered=False):
        send, oxs, oys, ovs, osrc = self._partition(xs, ys, vals, want_src or ordered)
        recv = self._exchange_counts(send)
        rx = self._a2a(oxs, send, recv)
        ry = self._a2a(oys, send, recv) if oys is not None else None
        rv = self._a2a(ovs, send, recv) if ovs is not None else None
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
        return send, recv, rx, ry, rv, osrc, rord

    def _write(self, op: int, xs, ys, vals, ordered: bool):
        _, _, rx, ry, rv, _, rord = self._route(xs, ys, vals, ordered=ordered)
        if not rx.numel():
            return
        if ordered:
            p = lambda t: t.data_ptr() if t is not None else None
            self._lib.smatrix_b200_apply_ordered(self.local._handle(), op, p(rx), p(ry), p(rv), p(rord),
                                                 rx.numel())
        else:
            (self.local.incr_batch, self.local.decr_batch)[op](rx, ry, rv)

    # ------------------------------------------------------------------ collective batch API
    def incr_batch(self, xs, ys, vals=None, ordered: bool = True):
        self._write(0, xs, ys, vals, ordered)

    def decr_batch(self, xs, ys, vals=None, ordered: bool = True):
        self._write(1, xs, ys, vals, ordered)

    def set_batch(self, xs, ys, vals=None):
        self._write(2, xs, ys, vals, True)     # last writer in GLOBAL input order wins

    def get_batch(self, xs, ys, out=None):
        send, recv, rx, ry, _, osrc, _ = self._route(xs, ys, None, want_src=True)
        ans = self.local.get_batch(rx, ry) if rx.numel() else self._buf(0)
        self._sync_torch()
        back = self._a2a(ans, recv, send)           # answers travel the reverse way
        if out is None:
            out = self._buf(xs.numel())
        out[osrc.long()] = back                      # un-permute into input order
        return out

    def rowlen_batch(self, xs):
        send, recv, rx, _, _, osrc, _ = self._route(xs, None, None, want_src=True)
        ans = self.local.rowlen_batch(rx) if rx.numel() else self._buf(0)
        self._sync_torch()
        back = self._a2a(ans, recv, send)
        out = self._buf(xs.numel())
        out[osrc.long()] = back
        return out

    # ------------------------------------------------------------------ local controls
    def stat(self, name):
        return self.local.stat(name)

    def __getattr__(self, name):
        # timer_start / timer_stop_ms / set_kernel_timing / gen_c2_* / probe_* / sync: the local shard's
        return getattr(self.local, name)

    def close(self):
        self.local.close()
