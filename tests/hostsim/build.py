"""TEST SCAFFOLDING ONLY: compile the host shim and the kernel file with gcc/g++ against the
malloc-backed fake runtime (hostsim.h) into tests/hostsim/libsmatrix_hostsim.so.  Used by the
`-m "not gpu"` tests to check host LOGIC; never built by build(), never loaded by the package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "libsmatrix_b200", "csrc")
SO = os.path.join(HERE, "libsmatrix_hostsim.so")


def build(defines: list[str] | None = None, suffix: str = "") -> str:
    """defines / suffix: a variant of the simulator library with other compile-time knobs (e.g. tiny
    re-placement tiles), written next to the default one."""
    return _build(os.path.join(HERE, f"libsmatrix_hostsim{suffix}.so"), defines or [], suffix)


def _build(SO: str, defines: list[str], suffix: str) -> str:
    srcs = [os.path.join(CSRC, "smx_kernels.cu"), os.path.join(CSRC, "smx_host.c"),
            os.path.join(HERE, "fake_runtime.cpp"), os.path.join(CSRC, "smx_router.c"),
            os.path.join(HERE, "warp_sched.cpp"),      # the fiber scheduler of the 32-lane variant (-DSMX_SIM_WARP32)
            os.path.join(HERE, "hostsim.h"), os.path.join(CSRC, "smx_internal.h")]
    if os.path.exists(SO) and all(os.path.getmtime(s) < os.path.getmtime(SO) for s in srcs):
        return SO
    common = ["-O1", "-g", "-fPIC", "-DSMX_HOSTSIM", f"-I{HERE}", f"-I{CSRC}", "-Wall",
              "-Wno-unknown-pragmas", "-Wno-unused-function"] + defines
    objs = []
    # -fno-gnu-unique: static locals of templates / inline functions (the kernels' "shared memory") would otherwise be
    # STB_GNU_UNIQUE, i.e. ONE object per process shared by every variant of this library that a test session loads —
    # with arrays sized by SMX_BLOCK (1 here, 256 in the 32-lane variant)
    cxx = ["-std=c++17", "-fno-gnu-unique"]
    for src, cc, extra in ((srcs[0], "g++", ["-x", "c++"] + cxx),
                           (srcs[1], "gcc", ["-std=gnu11"]),
                           (srcs[2], "g++", cxx),
                           (srcs[3], "gcc", ["-std=gnu11"]),
                           (srcs[4], "g++", cxx)):
        o = os.path.join(HERE, os.path.basename(src) + suffix + ".o")
        subprocess.run([cc] + common + extra + ["-c", src, "-o", o], check=True)
        objs.append(o)
    # -Bsymbolic: the fake cuda* symbols must bind inside this library even when a real
    # libcudart is already loaded in the process (torch)
    subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", "-o", SO] + objs + ["-lpthread", "-lrt"], check=True)
    return SO


if __name__ == "__main__":
    print(build())
