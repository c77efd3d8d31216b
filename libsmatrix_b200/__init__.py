"""libsmatrix_b200 — B200-native (sm_100a) implementation of libsmatrix's in-memory hot path.

The product is the C-ABI shared library `lib/libsmatrix_b200.so` (include/smatrix.h,
include/smatrix_batch.h); this package is the thin Python mirror of the reference's binding
classes on top of it.  Importing it never falls back to a CPU implementation.
"""
from .binding import DEFAULT_SO, load
from .matrix import SparseMatrix

__all__ = ["SparseMatrix", "load", "DEFAULT_SO"]
