"""The C-ABI library loads and exports every symbol include/subrecon_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from subrecon_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    return native.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "subrecon_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(native.EXPORTS) == syms


def test_every_declaration_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "subrecon_b200.h")).read()
    assert text.count("SubRecon.java") >= 5 and "JointBranchReconstruction.java" in text and "SiteResult.java" in text


def test_compact_struct_layout(built):
    assert ctypes.sizeof(native.Compact) == native.COMPACT_DTYPE.itemsize == 112


def test_no_cpu_fallback(built):
    """Without a CUDA device creation fails loudly with SR_ERR_CUDA; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.SrError) as ei:
        native.Context(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "subrecon_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libsubrecon_oracle" not in src, f


def test_jni_shim_compiles_against_the_abi():
    """No JDK in this image: the shim (jni/subrecon_jni.c) is compiled with -fsyntax-only against a minimal mock of
    <jni.h> (tests/cpp/jni_mock/jni.h), which checks every sr_* call in it against include/subrecon_b200.h."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fsyntax-only",
                           "-I", os.path.join(root, "tests", "cpp", "jni_mock"), "-I", os.path.join(root, "include"),
                           os.path.join(root, "jni", "subrecon_jni.c")])
