"""examples/c_host_example.c — the drop-in boundary driven from plain C (no Python in the loop).
CPU: it links against the product library and its checks hold on the serial simulator build.
GPU: the same binary logic against the CUDA library."""
import os
import subprocess

import pytest

from libsmatrix_b200 import build as product_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "c_host_example.c")


MULTI = os.path.join(ROOT, "examples", "c_multi_gpu_example.c")


def _compile(tmp_path, libdir, libname, src=SRC):
    exe = str(tmp_path / (os.path.basename(src)[:-2] + "_" + libname))
    r = subprocess.run(["gcc", "-O2", "-pthread", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src,
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


def test_c_multi_gpu_example_ranks_as_threads_on_the_simulator(tmp_path):
    """examples/c_multi_gpu_example.c: a C host drives the sharded matrix (include/smatrix_shard.h) with
    one thread per rank — here 3 ranks on the simulator's single device."""
    from hostsim import build as sim_build
    sim = sim_build.build()
    exe = _compile(tmp_path, os.path.dirname(sim), os.path.basename(sim)[3:-3], MULTI)
    r = subprocess.run([exe, "3", "200000", "1"], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, SMATRIX_DIR_LOG2="8", SMATRIX_SHARD_TIMEOUT="60"))
    assert r.returncode == 0 and "c_multi_gpu_example: OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_multi_gpu_example_on_the_gpus(tmp_path):
    """The same C program on the CUDA library: one rank per GPU of the box (2 ranks on one GPU if the
    box has a single one — the router only needs peer access between the ranks' devices)."""
    import torch
    gpus = max(1, torch.cuda.device_count())
    world = max(2, min(gpus, 8))
    so = product_build.build()
    exe = _compile(tmp_path, os.path.dirname(so), "smatrix_b200", MULTI)
    r = subprocess.run([exe, str(world), "8000000", str(min(gpus, world))], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "c_multi_gpu_example: OK" in r.stdout, r.stdout + r.stderr
