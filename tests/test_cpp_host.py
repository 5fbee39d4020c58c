"""The C++ host mirror (include/subrecon_b200_host.hpp): formatting rules on CPU; the batch JointBranchReconstruction on GPU,
printing exactly what the reference prints (compared with the same lines produced from oracle results)."""
import os
import subprocess

import numpy as np
import pytest

from subrecon_b200 import native, recon

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror")


@pytest.fixture(scope="module")
def exe():
    native.build()
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp")
    hdr = os.path.join(ROOT, "include", "subrecon_b200_host.hpp")
    if not os.path.exists(EXE) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(EXE):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", EXE, src,
                               "-L", os.path.join(ROOT, "subrecon_b200"), "-lsubrecon_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "subrecon_b200")])
    return EXE


def test_cpp_formatting_matches_java_rules_and_python_mirror(exe):
    out = subprocess.run([exe, "format"], capture_output=True, text=True, check=True).stdout.splitlines()
    xs = [1.0, 0.99, -15.66, 0.0, 1.2e-4, 0.001, 12345678.9, 1e-10, -0.0, 123.0, 0.5, 9999999.0, 1e7, 0.30000000000000004]
    assert out[:len(xs)] == [recon.java_double_to_string(x) for x in xs]
    assert out[:8] == ["1.0", "0.99", "-15.66", "0.0", "1.2E-4", "0.001", "1.23456789E7", "1.0E-10"]
    assert out[len(xs)] == "-2.0 0.13 -15.66"
    assert out[len(xs) + 1] == "Result\t14\t-7.0\tRK:0.3\tND:0.3\tAA:0.25"
    assert out[len(xs) + 2] == "Result\t14\t-7.0\tAA:0.25\tRK:0.3\tND:0.3"
    assert out[len(xs) + 3] == recon.SiteResult.get_header()
    assert out[len(xs) + 4] == "0.25 0.3"
    assert out[len(xs) + 5].startswith("Total lnL: -7.0040000000") and out[len(xs) + 6].startswith("0 sites have non-identical")


def write_blob(path, p, letters=None, sanity=False):
    ft = p.flat
    with open(path, "wb") as fh:
        np.array([ft.n_nodes, len(ft.child_idx), p.states.shape[0], p.states.shape[1], len(p.rates), ft.root,
                  int(letters is not None), int(sanity)], dtype=np.int64).tofile(fh)
        np.ascontiguousarray(ft.child_off, dtype=np.int32).tofile(fh)
        np.ascontiguousarray(ft.child_idx, dtype=np.int32).tofile(fh)
        np.ascontiguousarray(ft.blen, dtype=np.float64).tofile(fh)
        np.ascontiguousarray(ft.leaf_row, dtype=np.int32).tofile(fh)
        np.ascontiguousarray(p.states if letters is None else letters, dtype=np.uint8).tofile(fh)
        np.concatenate([p.model.Eval, p.model.Evec.ravel(), p.model.Ievc.ravel(), p.model.pi]).astype(np.float64).tofile(fh)
        np.ascontiguousarray(p.rates, dtype=np.float64).tofile(fh)


@pytest.mark.gpu
def test_cpp_host_prints_what_the_reference_prints(exe, example_problem, tmp_path):
    from oracle import oracle
    p = example_problem
    from subrecon_b200 import treeio
    names, seqs = treeio.read_fasta(os.path.join(ROOT, "tests", "golden", "example", "prot.fasta"))
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    # the Java host's route (letters in, compact records out) and the -debug route (encoded states, all 400 entries checked)
    for tag, kw in (("letters", dict(letters=treeio.alignment_chars(seqs))), ("sanity", dict(sanity=True))):
        blob = tmp_path / f"cfg1_{tag}.bin"
        write_blob(blob, p, **kw)
        for thr, verbose in ((0.5, 0), (0.2, 0), (0.125, 1), (0.0, 1)):
            got = subprocess.run([exe, "run", str(blob), repr(thr), str(verbose), "2"], capture_output=True, text=True, check=True).stdout.splitlines()
            want = [recon.SiteResult.get_header()] + recon.run_columns(
                [recon.SiteResult(i, lnl[i], joint[i], thr) for i in range(len(lnl))], bool(verbose), thr)
            assert got[:-1] == want[:-1], (tag, thr)
            assert abs(float(got[-1].split()[-1]) - float(want[-1].split()[-1])) < 1e-8


@pytest.mark.gpu
def test_cpp_host_strict_root(exe, tmp_path):
    """The Java patch switches "strict_root" on: a tip as root child with K > 1 (where the reference fails on every
    column, utils/Utils.java:56-70) is refused once, with a message; K = 1 runs."""
    from oracle import oracle
    from subrecon_b200 import treeio

    class P:
        pass
    p = P()
    p.flat = treeio.flatten_tree(treeio.parse_newick("(a:0.1,(b:0.2,c:0.3):0.1);"))
    p.model = oracle.OracleModel("WAG")
    p.states = treeio.encode_alignment(["A-", "AR", "RR"])
    p.rates = oracle.gamma_rates(4, 0.5)
    blob = tmp_path / "tiproot.bin"
    write_blob(blob, p)
    r = subprocess.run([exe, "run", str(blob), "0.5", "1", "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "child of the root is a tip" in r.stdout
    p.rates = np.array([1.0])
    write_blob(blob, p)
    r = subprocess.run([exe, "run", str(blob), "0.5", "1", "2"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("Result") == 2


@pytest.mark.gpu
def test_cpp_host_error_behaviour(exe, example_problem, tmp_path):
    import copy
    p = copy.copy(example_problem)
    p.rates = np.array([])                              # -k 0: "must be 1 or higher" (SubRecon.java:209-210)
    blob = tmp_path / "bad.bin"
    write_blob(blob, p)
    r = subprocess.run([exe, "run", str(blob), "0.5", "0", "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "rate categories" in r.stdout
