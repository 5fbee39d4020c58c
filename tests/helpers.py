"""Shared builders for the tests: BASELINE.json's configs at oracle-friendly sizes."""
import os

import numpy as np

from oracle import oracle
from subrecon_b200 import palmodel, synth, treeio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "tests", "golden", "example")


class Problem:
    def __init__(self, name, flat, states, model, rates, tree=None):
        self.name, self.flat, self.states, self.model, self.rates, self.tree = name, flat, states, model, rates, tree


def load_example():
    tree = treeio.read_tree(os.path.join(EX, "root.tre"))
    names, seqs = treeio.read_fasta(os.path.join(EX, "prot.fasta"))
    states = treeio.encode_alignment(seqs)
    flat = treeio.flatten_tree(tree, {n: i for i, n in enumerate(names)})
    freqs = palmodel.normalise_cli_frequencies(open(os.path.join(EX, "freqs")).read().strip().split(","))
    return Problem("cfg1-lysozyme", flat, states, oracle.OracleModel("WAG", freqs), oracle.gamma_rates(4, 0.5), tree)


def synthetic(cfg, n_sites):
    """cfg in {2,3,4,5}: the synthetic BASELINE configs with `n_sites` simulated columns."""
    if cfg == 2:
        tree, mname, K, freqs, seed = synth.balanced_tree(100, seed=2), "JTT", 4, None, 102
    elif cfg == 3:
        tree, mname, K, freqs, seed = synth.yule_tree(1000, seed=3), "WAG", 4, None, 103
    elif cfg == 4:
        rng = np.random.default_rng(4)
        d = rng.dirichlet(5.0 * np.ones(20))
        tree, mname, K, freqs, seed = synth.caterpillar_tree(5000, seed=4), "BLOSUM62", 8, palmodel.normalise_cli_frequencies(d), 104
    elif cfg == 5:
        tree, mname, K, freqs, seed = synth.yule_tree(2000, seed=5), "Dayhoff", 4, None, 105
    else:
        raise ValueError(cfg)
    model = oracle.OracleModel(mname, freqs)
    rates = oracle.gamma_rates(K, 0.5)
    states = synth.simulate_alignment(tree, model, rates, n_sites, seed)
    return Problem(f"cfg{cfg}", treeio.flatten_tree(tree), states, model, rates, tree)


def rel_err(got, want):
    """max relative error; entries of `want` below 1e-300 are compared absolutely."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    big = np.abs(want) > 1e-300
    r = 0.0
    if big.any():
        r = float(np.max(np.abs(got[big] - want[big]) / np.abs(want[big])))
    if (~big).any():
        r = max(r, float(np.max(np.abs(got[~big] - want[~big]))))
    return r


def lnl_err(got, want):
    """|delta| / max(1, |lnL|): a log-likelihood near 0 (all-missing column) has no meaningful relative error."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want)))) if got.size else 0.0


def random_newick(rng, n_tips):
    """Random rooted topology (binary root, occasional polytomies below it, some zero-length branches) for fuzz tests."""
    nodes = [f"t{i}:{rng.uniform(0.01, 0.5):.3f}" for i in range(n_tips)]
    while len(nodes) > 2:
        k = 2 if (rng.random() < 0.85 or len(nodes) < 5) else 3
        idx = set(int(i) for i in rng.choice(len(nodes), size=k, replace=False))
        sub = "(" + ",".join(nodes[i] for i in sorted(idx)) + f"):{rng.uniform(0.0, 0.4):.3f}"
        nodes = [x for i, x in enumerate(nodes) if i not in idx] + [sub]
    return "(" + ",".join(nodes) + ");"
