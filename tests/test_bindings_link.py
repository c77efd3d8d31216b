"""SURVEY.md 8f N3 — the reference's JNI and Ruby glue compile UNCHANGED against include/smatrix.h
and link against our static archive: every smatrix_* symbol they call resolves inside our library.
Needs the reference checkout (this container); skipped on the GPU box.  No JDK / Ruby here, so
tests/stubs/ provides just enough of jni.h / ruby.h to compile; nothing is executed."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "smatrix_jni.c")),
                                reason="reference checkout not present")


@pytest.fixture(scope="module")
def archive():
    from libsmatrix_b200 import build
    so = build.build()
    return os.path.join(os.path.dirname(so), "smatrix-static.a")


def _link(tmp_path, src, archive, extra):
    out = tmp_path / (os.path.basename(src) + ".so")
    cmd = ["gcc", "-shared", "-fPIC", "-w", f"-I{ROOT}/include", f"-I{ROOT}/tests/stubs", f"-I{REF}",
           src, archive, "-o", str(out), f"-L{CUDA}/lib64", "-lcudart_static", "-lstdc++", "-ldl",
           "-lrt", "-lpthread"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nm = subprocess.run(["nm", str(out)], capture_output=True, text=True, check=True).stdout
    return nm


def test_jni_glue_links_unchanged(tmp_path, archive):
    nm = _link(tmp_path, os.path.join(REF, "smatrix_jni.c"), archive, ["-Wl,--no-undefined"])
    for sym in ("Java_com_paulasmuth_libsmatrix_SparseMatrix_init",
                "Java_com_paulasmuth_libsmatrix_SparseMatrix_getRowNative",
                "smatrix_open", "smatrix_close", "smatrix_get", "smatrix_set", "smatrix_incr",
                "smatrix_decr", "smatrix_rowlen", "smatrix_getrow"):
        assert re.search(rf" [Tt] {sym}\b", nm), sym      # defined inside the linked object
    assert not re.search(r" U smatrix_", nm)


def test_ruby_glue_links_unchanged(tmp_path, archive):
    # like src/ruby/Makefile:18-19 the interpreter symbols (rb_*) stay undefined until load time
    nm = _link(tmp_path, os.path.join(REF, "smatrix_ruby.c"), archive, [])
    for sym in ("Init_smatrix", "smatrix_rb_incr", "smatrix_open", "smatrix_incr", "smatrix_close"):
        assert re.search(rf" [Tt] {sym}\b", nm), sym
    assert not re.search(r" U smatrix_", nm)
    assert re.search(r" U rb_define_method", nm)
