"""examples/c_host_example.c — the drop-in boundary driven from plain C (no Python in the loop).
CPU: it links against the product library and its checks hold on the serial simulator build.
GPU: the same binary logic against the CUDA library."""
import os
import subprocess

import pytest

from libsmatrix_b200 import build as product_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "c_host_example.c")


def _compile(tmp_path, libdir, libname):
    exe = str(tmp_path / f"c_host_example_{libname}")
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
                        "-L", libdir, f"-l{libname}", f"-Wl,-rpath,{libdir}", "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_example_links_against_the_product_library(tmp_path):
    so = product_build.build()
    _compile(tmp_path, os.path.dirname(so), "smatrix_b200")


def test_c_example_checks_hold_on_the_simulator(tmp_path):
    from hostsim import build as sim_build
    sim = sim_build.build()
    exe = _compile(tmp_path, os.path.dirname(sim), os.path.basename(sim)[3:-3])
    r = subprocess.run([exe, "300000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "c_host_example: OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_example_on_the_gpu(tmp_path):
    so = product_build.build()
    exe = _compile(tmp_path, os.path.dirname(so), "smatrix_b200")
    r = subprocess.run([exe, "20000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "c_host_example: OK" in r.stdout, r.stdout + r.stderr
