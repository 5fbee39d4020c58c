"""Host logic either side of the path: readers, flatteners, SiteResult formatting (no GPU)."""
import os
import sys

import numpy as np
import pytest

from subrecon_b200 import palmodel, recon, synth, treeio


def test_newick_pal_conventions():
    t = treeio.parse_newick("((a:1,b:2)x:3,(c,d,e)0.95:4);")
    assert t.children[0] == [1, 4] and t.children[4] == [5, 6, 7]      # Newick order kept, polytomy allowed
    assert t.blen[1] == 3.0 and t.blen[5] == 0.0                        # missing length = 0.0 (SimpleNode default)
    assert t.name[2] == "a" and t.name[1] == ""                         # internal labels ignored
    with pytest.raises(ValueError, match="single child"):
        treeio.parse_newick("((a:1):2,b:1);")
    t2 = treeio.parse_newick(" ( a : 1e-3 ,\n b:2.5E-1 ) ; ")
    assert t2.blen[1] == 1e-3 and t2.blen[2] == 0.25


def test_newick_roundtrip_and_depth():
    t = synth.caterpillar_tree(3000)                                     # deep: must not recurse
    t2 = treeio.parse_newick(t.to_newick())
    assert t2.to_newick() == t.to_newick() and t2.n_nodes == t.n_nodes


def test_example_inputs(example_problem):
    p = example_problem
    assert p.states.shape == (19, 130) and p.states.max() < 20          # 20 letters only, no gaps
    assert p.flat.n_nodes == 37 and len(p.tree.children[0]) == 2
    assert len({bytes(p.states[:, c]) for c in range(130)}) == 51      # 51 distinct columns
    a, b = p.tree.children[0]
    assert abs(p.tree.blen[a] + p.tree.blen[b] - 0.108120) < 1e-6


def test_state_encoding():
    s = treeio.encode_alignment(["ARNDCQEGHILKMFPSTWYV", "arndcqeghilkmfpstwyv", "-_?.*XBJZOU1 ", "AC"])
    assert s[0].tolist() == list(range(20)) and s[1].tolist() == list(range(20))
    assert (s[2, :13] == treeio.MISSING).all()
    assert s[3, :2].tolist() == [0, 4] and (s[3, 2:] == treeio.MISSING).all()       # padded with '-' to the longest
    # the letters form handed to the library's "input_chars" mode: same padding, translation left to the device
    seqs = ["ARNDCQEGHILKMFPSTWYV", "arndcqeghilkmfpstwyv", "-_?.*XBJZOU1 ", "AC"]
    letters = treeio.alignment_chars(seqs)
    assert letters.shape == s.shape and bytes(letters[3]) == b"AC" + b"-" * 18
    assert np.array_equal(treeio._STATE_LUT[letters], s)


def test_flatten_unknown_taxon():
    t = treeio.parse_newick("((a:1,b:1):1,(c:1,d:1):1);")
    with pytest.raises(KeyError):
        treeio.flatten_tree(t, {"a": 0, "b": 1, "c": 2})
    f = treeio.flatten_tree(t, {"a": 3, "b": 2, "c": 1, "d": 0})
    assert f.child_off.tolist() == [0, 2, 4, 4, 4, 6, 6, 6] and f.leaf_row.tolist() == [-1, -1, 3, 2, -1, 1, 0]


def test_fasta_and_phylip(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_text(">x one \nAC D\nEF\n>y\nac-\n")
    names, seqs = treeio.read_fasta(fa)
    assert names == ["x one", "y"] and seqs == ["ACDEF", "ac-"]
    ph = tmp_path / "a.phy"
    ph.write_text(" 2 6\nx  ACD\ny  EFG\n\nHIK\nLMN\n")
    names, seqs = treeio.read_phylip(ph)
    assert names == ["x", "y"] and seqs == ["ACDHIK", "EFGLMN"]


def test_synthetic_trees():
    b = synth.balanced_tree(100, 2)
    assert len(b.leaves()) == 100 and all(len(c) in (0, 2) for c in b.children)
    y = synth.yule_tree(1000, 3)
    assert len(y.leaves()) == 1000 and all(len(y.children[c]) == 2 for c in y.children[0])
    c = synth.caterpillar_tree(5000, 4)
    assert len(c.leaves()) == 5000 and all(len(c.children[x]) == 2 for x in c.children[0])
    assert min(c.blen[1:]) >= 1e-3 and max(c.blen) <= 1.0


def test_round_and_format():
    assert recon.round_double(-2.5, 0) == -2.0 and recon.round_double(0.125, 2) == 0.13       # Math.round = floor(x+1/2)
    assert recon.round_double(-15.664999, 2) == -15.66
    f = recon.java_double_to_string
    assert [f(x) for x in (1.0, 0.99, -15.66, 0.0, 1.2e-4, 0.001, 12345678.9, 1e-10)] == \
        ["1.0", "0.99", "-15.66", "0.0", "1.2E-4", "0.001", "1.23456789E7", "1.0E-10"]


def test_site_result_filter_sort_print():
    p = np.zeros((20, 20))
    p[1, 11] = 0.3
    p[0, 0] = 0.25
    p[2, 3] = 0.3
    p[5, 5] = 0.15
    r = recon.SiteResult(13, -7.004, p, threshold=0.2, sort_by_prob=True, sig_digits=2)
    assert r.get_max_ii_prob() == 0.25 and r.get_max_prob() == 0.3
    assert r.subs == ["RK", "ND", "AA"]                                  # stable: RK (canonical first) stays before ND
    assert str(r) == "Result\t14\t-7.0\tRK:0.3\tND:0.3\tAA:0.25"
    r2 = recon.SiteResult(13, -7.004, p, threshold=0.2, sort_by_prob=False)
    assert r2.subs == ["AA", "RK", "ND"]
    assert recon.SiteResult.get_header() == "[HEADER]\tsite\tln[P(D|theta,alpha)]\tP(A=a,B=b|D,theta,alpha)"


def test_run_columns_print_rule():
    p = np.zeros((20, 20)); p[1, 11] = 0.9; p[1, 1] = 0.1
    q = np.zeros((20, 20)); q[4, 4] = 0.97; q[4, 5] = 0.03
    rs = [recon.SiteResult(0, -1.0, p), recon.SiteResult(1, -2.0, q)]
    lines = recon.run_columns(rs)
    assert lines == ["Result\t1\t-1.0\tRK:0.9", "Total lnL: -3.0000000000"]
    lines = recon.run_columns([rs[1]])
    assert lines[0] == "Total lnL: -2.0000000000" and lines[1].startswith("0 sites have non-identical")
    assert len(recon.run_columns(rs, verbose=True)) == 3


def test_cli_frequency_normalisation():
    f = palmodel.normalise_cli_frequencies(["0.1"] * 20)
    assert abs(sum(f) - 1.0) < 1e-15
    with pytest.raises(ValueError):
        palmodel.normalise_cli_frequencies([0.5, 0.5])
    with pytest.raises(ValueError, match="not recognised"):
        palmodel.host_model("lg")


def jar_state_lut():
    """The reference's tip-state lookup for every byte, from the golden made by executing pal.jar (tests/golden/make_states_golden.py)."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aa_states_golden.json")))
    return np.array([s if 0 <= s < 20 else treeio.MISSING for s in g["states"]], dtype=np.uint8), g["states"]


def test_state_lut_is_what_the_jar_computes():
    """`DataTypeTool.getUniverisalAminoAcids().getState(c)` (molevo/AdvancedAlignmentAminoAcid.java:34-40), executed from
    PAL's bytecode for c = 0..255, against the host's encode table (and through it the device's, tests/test_gpu_parity.py::
    test_residue_letters_encoded_on_device). The jar answers 20 (unknown), 21 ('*') and -2 (gap) for non-residues: all of
    them fail `0 <= state < 20` in JointBranchReconstruction.java:199-205 and become missing data."""
    lut, raw = jar_state_lut()
    assert np.array_equal(lut, treeio._STATE_LUT)
    assert sorted(set(raw)) == [-2] + list(range(22)) and raw[ord("-")] == -2 and raw[ord("*")] == 21 and raw[ord("X")] == 20
    assert [raw[ord(c)] for c in treeio.AA_ORDER] == list(range(20)) == [raw[ord(c.lower())] for c in treeio.AA_ORDER]


@pytest.mark.skipif(not os.path.exists("/root/reference/lib/pal.jar"), reason="reference jar only in the build container")
def test_state_golden_regenerates_from_the_jar():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_states_golden
    cls, states, stubbed = make_states_golden.jar_states()
    assert cls == "pal/datatype/SpecificAminoAcids" and stubbed == [] and states == jar_state_lut()[1]


def test_torch_simulator_matches_the_numpy_one():
    """bench.py fills its alignments with `simulate_alignment_torch` (every column distinct, seconds on a GPU); the parity
    tests use the numpy simulator. Same generative model: compare what the kernels' behaviour depends on — state
    frequencies, missing cells, identity between tips, the share of constant columns — and the prefix property."""
    tree = synth.yule_tree(40, seed=7)
    m, r = palmodel.host_model("wag"), palmodel.gamma_rates(4, 0.5)
    n = 30_000
    a = synth.simulate_alignment(tree, m, r, n, seed=1)
    b = synth.simulate_alignment_torch(tree, m, r, n, seed=2, device="cpu", block=8192)
    assert a.shape == b.shape and b.dtype == np.uint8 and set(np.unique(b)) <= set(range(20)) | {treeio.MISSING}
    assert abs((a == 255).mean() - (b == 255).mean()) < 2e-3
    fa = np.bincount(a[a < 20], minlength=20) / np.count_nonzero(a < 20)
    fb = np.bincount(b[b < 20], minlength=20) / np.count_nonzero(b < 20)
    assert np.abs(fa - fb).max() < 0.01
    for i, j in ((0, 1), (3, 30), (10, 39)):
        assert abs(np.mean(a[i] == a[j]) - np.mean(b[i] == b[j])) < 0.015

    def constant(x):
        y = np.where(x == 255, x[0][None, :], x)
        return np.mean((y == y[0][None, :]).all(axis=0))
    assert abs(constant(a) - constant(b)) < 0.02
    short = synth.simulate_alignment_torch(tree, m, r, 10_000, seed=2, device="cpu", block=8192)
    assert np.array_equal(short, b[:, :10_000])                       # a shorter run is a prefix of a longer one


def test_bench_clock_sampler_windows():
    """bench.py reports clocks / power / throttle reasons for the window `value` was timed over (and for the whole run): the
    windowing must pick the right samples, fall back to the nearest one for a region shorter than the sampling period,
    and name any throttle reason seen inside the window."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    s = bench.ClockSampler(0)
    s.nvml = None                                                   # no NVML / GPU here: feed samples by hand
    off = [False] * 4
    s.samples = [(0.0, 1965.0, 1965.0, 300.0, off), (1.0, 1965.0, 1965.0, 640.0, off), (2.0, 1800.0, 1965.0, 990.0, [False, False, False, True]),
                 (3.0, 1965.0, 1965.0, 310.0, off)]
    w = s.window(0.5, 1.5)
    assert w["samples"] == 1 and w["sm_mhz"] == 1965.0 and w["reasons"] == [] and w["power_w_max"] == 640.0
    w = s.window(0.5, 2.5)
    assert w["samples"] == 2 and w["sm_min_mhz"] == 1800.0 and w["reasons"] == ["sw_power_cap"] and w["sm_max_mhz"] == 1965.0
    w = s.window(1.2, 1.3)                                          # shorter than the period: the nearest sample stands in
    assert w["samples"] == 1 and w["power_w_max"] == 640.0
    assert s.window(None, None)["samples"] == 4
    assert bench.ClockSampler.NAMES == ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
