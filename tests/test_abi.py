"""CPU tests of the drop-in boundary: the CUDA library builds (nvcc cross-compiles sm_100a here),
loads, exports every symbol include/*.h declares and the reference's bindings call, and refuses
to run without a GPU (no CPU fallback).  No compute calls."""
import os
import re
import subprocess

import pytest

import libsmatrix_b200
from libsmatrix_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


@pytest.fixture(scope="module")
def so_path():
    from libsmatrix_b200 import build
    return build.build()


def declared_symbols():
    names = set()
    for h in sorted(os.listdir(INCLUDE)):
        text = open(os.path.join(INCLUDE, h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(smatrix_[a-z0-9_]+)\s*\(", text))
    return names


def test_headers_declare_the_reference_api():
    # reference src/smatrix.h:87-94
    ref_api = {"smatrix_open", "smatrix_close", "smatrix_get", "smatrix_set", "smatrix_incr",
               "smatrix_decr", "smatrix_rowlen", "smatrix_getrow"}
    batch_api = {"smatrix_incr_batch", "smatrix_decr_batch", "smatrix_set_batch", "smatrix_get_batch",
                 "smatrix_rowlen_batch", "smatrix_getrow_batch"}   # BASELINE.json north_star
    assert ref_api | batch_api <= declared_symbols()


def test_library_exports_every_declared_symbol(so_path):
    out = subprocess.run(["nm", "-D", "--defined-only", so_path], capture_output=True, text=True,
                         check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = declared_symbols() - exported
    assert not missing, f"declared but not exported: {sorted(missing)}"


def test_binding_covers_every_declared_symbol(so_path):
    assert declared_symbols() == set(binding.PROTOTYPES)
    binding.load(so_path)      # raises AttributeError if a prototype has no symbol


def test_static_archive_for_bindings(so_path):
    """The reference's JNI/Ruby glue links smatrix.o statically (src/java/Makefile:22-23)."""
    a = os.path.join(os.path.dirname(so_path), "smatrix-static.a")
    out = subprocess.run(["nm", a], capture_output=True, text=True, check=True).stdout
    for sym in ("smatrix_open", "smatrix_getrow", "smatrix_incr_batch"):
        assert re.search(rf" T {sym}\b", out)


def test_kernels_are_sm100a_sass(so_path):
    out = subprocess.run(["cuobjdump", "-lelf", so_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_a_gpu(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ValueError):        # smatrix_open returns NULL -> binding raises
        libsmatrix_b200.SparseMatrix()


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "libsmatrix_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text.replace("smatrix_oracle", "oracle") or f == "__none__", (
                    f"{f} mentions the oracle")
                assert "hostsim" not in text.lower() or f in ("smx_kernels.cu",), f
