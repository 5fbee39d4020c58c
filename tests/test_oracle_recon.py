"""The oracle's per-column task (oracle/recon_oracle.c) against (a) the committed golden outputs for the reference's
own example, (b) the reference's -debug invariants, (c) independent brute-force / high-precision evaluations."""
import os

import numpy as np
import pytest

from helpers import load_example, rel_err, synthetic
from oracle import oracle
from subrecon_b200 import treeio

HERE = os.path.dirname(os.path.abspath(__file__))


def test_lysozyme_golden_bit_exact(example_problem):
    """tests/golden/lysozyme_golden.npz = PAL bytecode matrices + a Python transliteration of recon()."""
    g = np.load(os.path.join(HERE, "golden", "lysozyme_golden.npz"))
    p = example_problem
    assert np.array_equal(p.rates, g["rates"]) and np.array_equal(p.model.pi, g["pi"])
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    assert np.array_equal(lnl, g["lnl"])
    assert np.array_equal(joint, g["joint"])
    tot = 0.0
    for x in lnl:
        tot += x
    assert "%.10f" % tot == "-673.1980928246"


def test_work_modes_identical(example_problem):
    p = example_problem
    l0, j0 = oracle.run(p.flat, p.states, p.model, p.rates, mode=0, threads=2)
    for mode in (1, 2):
        l, j = oracle.run(p.flat, p.states, p.model, p.rates, mode=mode, threads=4, n=20)
        assert np.array_equal(l, l0[:20]) and np.array_equal(j, j0[:20])


def test_debug_invariants(example_problem):
    """checkSumToOne and checkConditionalsSumToMarginal (JointBranchReconstruction.java:142-166) at 1e-12."""
    p = example_problem
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    assert np.max(np.abs(joint.sum(axis=(1, 2)) - 1.0)) < 1e-12
    ref = plain_pruning_lnl(p)
    assert np.max(np.abs(lnl - ref)) < 1e-11


def plain_pruning_lnl(p, cols=None):
    """Ordinary whole-tree Felsenstein pruning at the root (computeTotalL, :176-184) in numpy, linear space with
    per-node log scaling — shares no code with the oracle."""
    ft, st = p.flat, p.states
    cols = range(st.shape[1]) if cols is None else cols
    K = len(p.rates)
    out = []
    order, stack = [], [ft.root]
    while stack:
        v = stack.pop()
        order.append(v)
        stack.extend(ft.child_idx[ft.child_off[v]:ft.child_off[v + 1]])
    Ps = {(v, k): oracle.transition_probs(p.model.Eval, p.model.Evec, p.model.Ievc, ft.blen[v] * p.rates[k])
          for v in range(ft.n_nodes) for k in range(K) if v != ft.root}
    for c in cols:
        tot = []
        for k in range(K):
            part, logs = {}, 0.0
            for v in reversed(order):
                ch = ft.child_idx[ft.child_off[v]:ft.child_off[v + 1]]
                if len(ch) == 0:
                    s = st[ft.leaf_row[v], c]
                    vec = np.ones(20) if s >= 20 else np.eye(20)[s]
                else:
                    vec = np.ones(20)
                    for x in ch:
                        vec = vec * (Ps[(x, k)] @ part.pop(x))
                m = vec.max()
                logs += np.log(m)
                part[v] = vec / m
            tot.append(np.log(np.dot(p.model.pi, part[ft.root])) + logs)
        tot = np.array(tot)
        out.append(tot.max() + np.log(np.exp(tot - tot.max()).sum()) - np.log(K))
    return np.array(out)


def test_brute_force_five_taxa():
    """Explicit sum over the root state with scipy expm vs the merged-root-branch formula."""
    from scipy.linalg import expm
    tree = treeio.parse_newick("((t1:0.1,t2:0.25):0.05,(t3:0.3,(t4:0.02,t5:0.4):0.15):0.07);")
    flat = treeio.flatten_tree(tree)
    model = oracle.OracleModel("JTT")
    rates = oracle.gamma_rates(4, 0.5)
    states = treeio.encode_alignment(["AR-", "AKX", "NRW", "DRw", "CEY"])
    lnl, joint = oracle.run(flat, states, model, rates)
    R, pi = model.R, model.pi_model
    nA, nB = tree.children[0]

    def cond(v, c, rate):
        if not tree.children[v]:
            s = states[flat.leaf_row[v], c]
            return np.ones(20) if s >= 20 else np.eye(20)[s]
        out = np.ones(20)
        for x in tree.children[v]:
            out *= expm(R * tree.blen[x] * rate) @ cond(x, c, rate)
        return out

    for c in range(3):
        J = np.zeros((20, 20))
        for rate in rates:
            a, b = cond(nA, c, rate), cond(nB, c, rate)
            PA, PB = expm(R * tree.blen[nA] * rate), expm(R * tree.blen[nB] * rate)
            # sum over the root state r: pi_r P_A[r,a] P_B[r,b]
            J += (pi[:, None, None] * PA[:, :, None] * PB[:, None, :]).sum(axis=0) * a[:, None] * b[None, :] / len(rates)
        assert abs(np.log(J.sum()) - lnl[c]) < 1e-11
        assert rel_err(joint[c], J / J.sum()) < 1e-9


def test_high_precision_small_tree():
    """50-digit mpmath evaluation of the same likelihood pins the mathematics independent of rounding."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    tree = treeio.parse_newick("((a:0.2,b:0.01):0.1,(c:0.3,d:0.05,e:0.6):0.02);")      # polytomy below the root
    flat = treeio.flatten_tree(tree)
    model = oracle.OracleModel("WAG")
    rates = np.array([0.3, 1.7])                                                          # -rates hook (CustomRates)
    states = treeio.encode_alignment(["AV", "AV", "L-", "LI", "KI"])
    lnl, joint = oracle.run(flat, states, model, rates)
    Rm = mp.matrix(model.R.tolist())
    pi = [mp.mpf(float(x)) for x in model.pi]

    def cond(v, c, rate):
        if not tree.children[v]:
            s = int(states[flat.leaf_row[v], c])
            return [mp.mpf(1)] * 20 if s >= 20 else [mp.mpf(1 if i == s else 0) for i in range(20)]
        out = [mp.mpf(1)] * 20
        for x in tree.children[v]:
            P = mp.expm(Rm * (tree.blen[x] * float(rate)))
            cx = cond(x, c, rate)
            out = [out[i] * sum(P[i, j] * cx[j] for j in range(20)) for i in range(20)]
        return out

    nA, nB = tree.children[0]
    for c in range(2):
        J = [[mp.mpf(0)] * 20 for _ in range(20)]
        for rate in rates:
            a, b = cond(nA, c, rate), cond(nB, c, rate)
            P = mp.expm(Rm * ((tree.blen[nA] + tree.blen[nB]) * float(rate)))
            for i in range(20):
                for j in range(20):
                    J[i][j] += pi[i] * a[i] * P[i, j] * b[j] / len(rates)
        tot = sum(sum(r) for r in J)
        assert abs(float(mp.log(tot)) - lnl[c]) < 1e-12
        Jn = np.array([[float(x / tot) for x in r] for r in J])
        assert rel_err(joint[c], Jn) < 1e-10


def test_single_rate_and_missing_column():
    tree = treeio.parse_newick("((a:0.1,b:0.2):0.1,(c:0.1,d:0.3):0.2);")
    flat = treeio.flatten_tree(tree)
    model = oracle.OracleModel("Dayhoff")
    states = treeio.encode_alignment(["A-", "A-", "A?", "Ax"])
    lnl, joint = oracle.run(flat, states, model, np.array([1.0]))
    assert abs(lnl[1]) < 1e-9                               # an all-missing column has likelihood 1
    assert abs(joint[1].sum() - 1.0) < 1e-12
    assert np.argmax(joint[0]) == 0                         # AA


def test_leaf_root_child_matches_reference_failure():
    """A tip as root child with K>1: the reference indexes lnValues[-1]. The oracle reports NaN."""
    tree = treeio.parse_newick("(a:0.1,(b:0.2,c:0.3):0.1);")
    flat = treeio.flatten_tree(tree)
    model = oracle.OracleModel("WAG")
    states = treeio.encode_alignment(["A", "A", "R"])
    lnl, joint = oracle.run(flat, states, model, oracle.gamma_rates(4, 0.5))
    assert np.isnan(joint[0]).any()
    lnl1, joint1 = oracle.run(flat, states, model, np.array([1.0]))       # K=1 takes the length-1 shortcut
    assert np.isfinite(lnl1[0]) and abs(joint1[0].sum() - 1.0) < 1e-12


def test_cfg2_subset_invariants():
    p = synthetic(2, 64)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=4)
    assert np.max(np.abs(joint.sum(axis=(1, 2)) - 1.0)) < 1e-12
    ref = plain_pruning_lnl(p, cols=range(8))
    assert np.max(np.abs(lnl[:8] - ref)) < 1e-10
