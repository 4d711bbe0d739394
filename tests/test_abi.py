"""The C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol declared in
include/pic_b200.h; the ctypes mirrors of the POD structs match the C layout.  (No compute calls: this is a CPU test.)"""
import ctypes
import os
import re

import pytest

from pypic3d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pic_[A-Za-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _lib.build()
    L = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/pic_b200.h but not exported by libpic_b200.so"
    # every declared symbol is bound by the Python layer, and nothing undeclared is bound
    assert set(names) == set(_lib.SIGNATURES)


def test_struct_layouts_match():
    L = ctypes.CDLL(_lib.build())
    L.pic_params_size.restype = ctypes.c_int
    assert L.pic_params_size() == ctypes.sizeof(_lib.PicParams)
    L.pic_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.pic_version()
    assert ctypes.sizeof(_lib.PicSoA) == 6 * 8 + 8 + 8 + 8 + 8
    assert ctypes.sizeof(_lib.PicLeave) == 8 + 27 * 4 * 2


def test_library_targets_sm_100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pypic3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_operators_refuse_cpu_tensors():
    import pytest
    import torch
    from pypic3d_b200 import ops
    with pytest.raises(ops.PicError):
        ops._chk(torch.zeros(3), "x")


def test_xla_ffi_source_compiles_gated(tmp_path):
    """csrc/pic_xla_ffi.cc (the XLA FFI custom-call handlers over the C ABI) is compile-gated on jaxlib's headers: without them
    (this image) it must still be a valid translation unit; with them it defines the handler symbols."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    src = os.path.join(root, "pypic3d_b200", "csrc", "pic_xla_ffi.cc")
    inc = []
    try:
        import jax  # noqa: F401
        inc = ["-I" + jax.ffi.include_dir()]
    except Exception:
        pass
    obj = str(tmp_path / "ffi.o")
    r = subprocess.run([gxx, "-std=c++17", "-c", "-I" + os.path.join(root, "include"), "-I/usr/local/cuda/include"] + inc + [src, "-o", obj],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    syms = subprocess.run(["nm", obj], capture_output=True, text=True).stdout
    assert ("PicUpdateB" in syms) if inc else ("pic_xla_ffi_available" in syms)
