"""Build the product library in-tree: nvcc (sm_100a) for the kernels, gcc for the C host shim.

    python -m libsmatrix_b200.build

Outputs (git-ignored, shipped to the GPU box by gpurun):
    libsmatrix_b200/lib/libsmatrix_b200.so   the drop-in shared library (smatrix.h + batch API)
    libsmatrix_b200/lib/smatrix-static.a     same objects, for bindings that link statically
                                             (reference src/java/Makefile:22-23)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SUFFIX = os.environ.get("SMX_LIB_SUFFIX", "")     # A/B builds of compile-time knobs: lib/libsmatrix_b200<suffix>.so
SO = os.path.join(LIBDIR, f"libsmatrix_b200{SUFFIX}.so")
STATIC = os.path.join(LIBDIR, f"smatrix-static{SUFFIX}.a")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = shutil.which("nvcc") or os.path.join(CUDA_HOME, "bin", "nvcc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("SMX_NVCC_EXTRA", "").split()
GCC_FLAGS = ["-O2", "-fPIC", "-Wall", "-Wextra", "-std=gnu11", f"-I{CUDA_HOME}/include"]


def _stale(out: str, deps: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    cu = os.path.join(CSRC, "smx_kernels.cu")
    hc = os.path.join(CSRC, "smx_host.c")
    rc = os.path.join(CSRC, "smx_router.c")
    hdrs = [os.path.join(CSRC, "smx_internal.h")] + [
        os.path.join(HERE, "..", "include", h) for h in ("smatrix.h", "smatrix_batch.h", "smatrix_b200.h",
                                                         "smatrix_shard.h")]
    cu_o = os.path.join(LIBDIR, f"smx_kernels{SUFFIX}.o")
    hc_o = os.path.join(LIBDIR, f"smx_host{SUFFIX}.o")
    rc_o = os.path.join(LIBDIR, f"smx_router{SUFFIX}.o")
    log = []
    if force or _stale(cu_o, [cu] + hdrs):
        r = subprocess.run([NVCC] + NVCC_FLAGS + ["-c", cu, "-o", cu_o], capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed")
        with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
            f.write(r.stderr)
    if force or _stale(hc_o, [hc] + hdrs):
        r = subprocess.run(["gcc"] + GCC_FLAGS + ["-c", hc, "-o", hc_o], capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed")
    if force or _stale(rc_o, [rc] + hdrs):
        r = subprocess.run(["gcc"] + GCC_FLAGS + ["-c", rc, "-o", rc_o], capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed")
    if force or _stale(SO, [cu_o, hc_o, rc_o]):
        # nvcc links the static CUDA runtime, so the .so only needs libcuda from the driver
        r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", cu_o, hc_o, rc_o,
                            "-o", SO, "-Xlinker", "--no-undefined", "-Xlinker", "-Bsymbolic",
                            "-Xlinker", "--exclude-libs,ALL", "-lpthread", "-lrt"],
                           capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        subprocess.run(["ar", "crs", STATIC, cu_o, hc_o, rc_o], check=True)
    if verbose:
        sys.stderr.write("".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
