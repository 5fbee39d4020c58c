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


JNI_TYPE = {"int": "jint", "long": "jlong", "double": "jdouble", "boolean": "jboolean", "void": "void", "String": "jstring",
            "ByteBuffer": "jobject", "double[]": "jdoubleArray", "int[]": "jintArray"}


def test_java_natives_match_the_jni_shim():
    """No javac here, so the Java side cannot be compiled — but the binding between java/subrecon/recon/NativeRecon.java and
    jni/subrecon_jni.c can be checked statically: every `native` method has exactly one JNI function with the mangled name
    Java_subrecon_recon_NativeRecon_<method>, (JNIEnv*, jclass) in front, and the JNI image of each Java parameter and of
    the return type; and the shim defines no function Java does not declare. Plus: balanced braces / parentheses and the
    package / class names the mangled names imply."""
    java = open(os.path.join(ROOT, "java", "subrecon", "recon", "NativeRecon.java")).read()
    code = re.sub(r"/\*.*?\*/", "", java, flags=re.S)
    code = re.sub(r"//[^\n]*", "", code)
    code = re.sub(r'"(\\.|[^"\\])*"', '""', code)
    code = re.sub(r"'(\\.|[^'\\])'", "' '", code)
    for a, b in ("{}", "()", "[]"):
        assert code.count(a) == code.count(b), f"unbalanced {a}{b} in NativeRecon.java"
    assert "package subrecon.recon;" in code and "public final class NativeRecon" in code
    natives = {}
    for ret, name, args in re.findall(r"static\s+native\s+([\w\[\]]+)\s+(\w+)\s*\(([^)]*)\)\s*;", code):
        params = [" ".join(a.split()[:-1]) for a in args.split(",") if a.strip()]
        natives[name] = (JNI_TYPE[ret], [JNI_TYPE[t] for t in params])
    assert len(natives) >= 12 and {"multiCreate", "multiRunStreamed", "multiSetTree", "hostAlloc", "compactCapFor"} <= set(natives)
    shim = open(os.path.join(ROOT, "jni", "subrecon_jni.c")).read()
    shim = re.sub(r"/\*.*?\*/", "", shim, flags=re.S)
    found = {}
    for ret, name, args in re.findall(r"JNIEXPORT\s+(\w+)\s+JNICALL\s+Java_subrecon_recon_NativeRecon_(\w+)\s*\(([^)]*)\)", shim):
        types = [re.sub(r"\s+", "", re.match(r"^(.*?)(\w+)$", a.strip()).group(1)) for a in args.split(",")]     # "JNIEnv *env" -> "JNIEnv*"
        assert types[:2] == ["JNIEnv*", "jclass"], (name, types)
        found[name] = (ret, types[2:])
    assert found == natives
    # every sr_* entry the Java host needs is reached from the shim
    for sym in ("sr_multi_create", "sr_multi_destroy", "sr_multi_last_error", "sr_multi_count", "sr_multi_set_model", "sr_multi_set_rates",
                "sr_multi_set_tree", "sr_multi_set_option", "sr_multi_run_streamed", "sr_compact_cap_for", "sr_compact_stride",
                "sr_host_alloc", "sr_host_free"):
        assert sym + "(" in shim, sym
    patch = open(os.path.join(ROOT, "java", "SubRecon.run.patch.java")).read()
    pc = re.sub(r"//[^\n]*", "", patch)
    pc = re.sub(r'"(\\.|[^"\\])*"', '""', pc)
    for a, b in ("{}", "()"):
        assert pc.count(a) == pc.count(b), f"unbalanced {a}{b} in SubRecon.run.patch.java"
    # the patch only calls what NativeRecon declares publicly
    public = set(re.findall(r"public\s+(?:static\s+)?(?:native\s+)?[\w\[\]<>]+\s+(\w+)\s*\(", code)) | {"NativeRecon"}
    for call in re.findall(r"\b(?:nr|NativeRecon)\.(\w+)\s*\(", pc):
        assert call in public, call
