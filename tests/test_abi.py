"""CPU: the C-ABI shared library loads and exports every symbol include/tgtc_b200.h declares; argument
validation and the no-GPU failure path return status codes + messages (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from tgtc_style_b200 import _lib
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tgtc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tgtc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from tgtc_style_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "library does not export " + s
        assert s in _lib.PROTOTYPES, "python binding lacks a prototype for " + s
    assert sorted(_lib.PROTOTYPES) == syms


def test_abi_version(lib):
    assert lib.tgtc_abi_version() == 2   # matches TGTC_ABI_VERSION in include/tgtc_b200.h


def test_workspace_bytes_arithmetic(lib):
    a = lib.tgtc_render_workspace_bytes(1024, 64, 64, 0)
    # rgbsigma coarse + fine, weights_coarse, ts_fine, shared ts row
    assert a >= 1024 * 64 * 16 + 1024 * 128 * 16 + 1024 * 64 * 4 + 1024 * 128 * 4 + 64 * 4
    assert lib.tgtc_render_workspace_bytes(4096, 64, 64, 1024) == a          # chunked passes reuse one pass worth
    assert lib.tgtc_render_workspace_bytes(0, 64, 64, 0) == 0
    assert lib.tgtc_render_frame_workspace_bytes(1024, 64, 64, 0) >= a + 2 * 1024 * 12


def test_null_context_is_an_error(lib):
    rc = lib.tgtc_set_weights(None, 0, None, None)
    assert rc == 1 and b"null context" in lib.tgtc_last_error()
    rc = lib.tgtc_raygen(None, 4, 4, None, None, 1, 1.0, 0, 0, 16, None, None, None)
    assert rc == 1
    assert lib.tgtc_launch_count(None) == 0
    assert lib.tgtc_destroy(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure path")
def test_create_without_gpu_fails_loudly(lib):
    h = ctypes.c_void_p()
    rc = lib.tgtc_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert len(lib.tgtc_last_error()) > 0
    import tgtc_style_b200 as T
    with pytest.raises(T.TgtcError):
        T.NerfRenderer()


def test_product_never_imports_oracle():
    """the product package must not reference oracle/ (a CPU route would void the parity claims)."""
    pkg = os.path.join(ROOT, "tgtc-style_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "render_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_header_is_plain_c():
    """include/tgtc_b200.h is the C ABI a cgo / JNI / ctypes binding would consume: it must compile as C99 and as C++"""
    import shutil
    import subprocess
    import tempfile
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "hdr.c")
        open(src, "w").write('#include "include/tgtc_b200.h"\nint main(void) { return 0; }\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", ROOT, src], check=True)
        subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", ROOT, "-x", "c++", src], check=True)


def test_oracle_files_say_test_infrastructure():
    """every file under oracle/ declares itself test infrastructure (it must never be mistaken for a product path)"""
    odir = os.path.join(ROOT, "oracle")
    for f in sorted(os.listdir(odir)):
        if f.endswith((".py", ".c", ".h")):
            head = open(os.path.join(odir, f)).read(600)
            assert "TEST INFRASTRUCTURE" in head.upper(), f
