"""oracle/model.py — the semantic model of libsmatrix's in-memory path (pure Python).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and never by libsmatrix_b200/).

This is the layout-free statement of what the reference computes on its *safe domain*
(SURVEY.md 8a, Q1-Q7): per row a dict of the non-zero columns, the column-0 scalar ``c0`` and
the reference's running counter pair (S, U) that makes ``rowlen`` history dependent
(src/smatrix.c:212-223 returns ``used``; src/smatrix.c:343-360 and :383-416 maintain it).

It also carries the *closed form* the CUDA table uses instead of (S, U):

    rowlen = L + d,  L = number of columns != 0 ever written,
    d      = 1 iff some virtual resize happened while c0 != 0,
    a virtual resize happens on the insertion that makes L one of 10, 18, 34, 66, ...
    (L = 2**j + 2, j >= 3) as long as d == 0.

``check_closed_form`` asserts both agree after every op; tests/test_oracle.py pins the model
against the compiled reference and the C restatement.
"""
from __future__ import annotations

M32 = 0xFFFFFFFF


def is_resize_count(n: int) -> bool:
    """True when inserting the n-th non-zero column triggers the reference's row growth
    (used > size/2 at src/smatrix.c:346, starting from 16 cells) while column 0 is uncounted."""
    return n >= 10 and ((n - 2) & (n - 3)) == 0


class Row:
    __slots__ = ("S", "U", "vals", "c0", "L", "d")

    def __init__(self):
        self.S = 16
        self.U = 0
        self.vals: dict[int, int] = {}
        self.c0 = 0
        self.L = 0
        self.d = 0


class ModelMatrix:
    """Sequential model; method names follow src/smatrix.h:87-94."""

    def __init__(self, check_closed_form: bool = True):
        self.rows: dict[int, Row] = {}
        self.check = check_closed_form

    # -- internals -------------------------------------------------------------------------
    def _cell_for_write(self, x: int, y: int) -> Row:
        row = self.rows.get(x)
        if row is None:
            row = self.rows[x] = Row()
        if y != 0 and y not in row.vals:
            if row.U > row.S // 2:  # src/smatrix.c:346-348, recount at :397-402
                row.S *= 2
                row.U = len(row.vals) + (1 if row.c0 != 0 else 0)
            row.U += 1
            row.vals[y] = 0
            # closed form
            row.L += 1
            if row.d == 0 and is_resize_count(row.L) and row.c0 != 0:
                row.d = 1
            if self.check:
                assert row.U == row.L + row.d, (x, y, row.U, row.L, row.d)
        return row

    # -- the 8-function API ----------------------------------------------------------------
    def get(self, x: int, y: int) -> int:
        row = self.rows.get(x)
        if row is None:
            return 0
        return row.c0 if y == 0 else row.vals.get(y, 0)

    def _write(self, x: int, y: int, fn) -> int:
        row = self._cell_for_write(x, y)
        if y == 0:
            new = fn(row.c0) & M32
            if row.c0 != 0 and new == 0:
                raise UnsafeDomain(f"column 0 of row {x} returns to 0 (SURVEY.md Q3)")
            row.c0 = new
            return new
        new = fn(row.vals[y]) & M32
        row.vals[y] = new
        return new

    def set(self, x: int, y: int, v: int) -> int:
        return self._write(x, y, lambda old: v)

    def incr(self, x: int, y: int, v: int) -> int:
        return self._write(x, y, lambda old: old + v)

    def decr(self, x: int, y: int, v: int) -> int:
        return self._write(x, y, lambda old: old - v)

    def rowlen(self, x: int) -> int:
        row = self.rows.get(x)
        return 0 if row is None else row.U

    def getrow(self, x: int) -> list[tuple[int, int]]:
        """All live pairs sorted by column (the comparison key the contract uses)."""
        row = self.rows.get(x)
        if row is None:
            return []
        out = sorted(row.vals.items())
        if row.c0 != 0:
            out.insert(0, (0, row.c0))
        return out

    def nnz(self) -> int:
        return sum(len(r.vals) + (1 if r.c0 else 0) for r in self.rows.values())


class UnsafeDomain(Exception):
    """The op stream left the domain on which the reference's result is layout independent."""


def catch_up(slog: int, d: int, live: int, z: bool) -> tuple[int, int]:
    """The lazy form of the (S, d) automaton the CUDA table uses (smx_kernels.cu rowlen_catch_up):
    given the state as of some earlier moment and the current number of columns, with z = (c0 != 0)
    constant in between, return the state the step-by-step reference automaton would be in."""
    if live:
        last = live - 1
        while last + d > (1 << (slog - 1)):
            slog += 1
            d = 1 if z else d
    return slog, d
