"""Synthetic trees and alignments of the shapes BASELINE.json names.

There are no datasets in this environment, so `bench.py` and the tests draw trees and simulate
columns down them under the same model and rate classes the path is run with.
"""
from __future__ import annotations

import numpy as np

from .treeio import MISSING, Tree


def _blen(rng, mean):
    return float(min(1.0, max(1e-3, rng.exponential(mean))))


def balanced_tree(n_tips: int, seed: int = 2, mean_blen: float = 0.05) -> Tree:
    """Recursive ceil(n/2)/floor(n/2) split (config #2)."""
    rng = np.random.default_rng(seed)
    t = Tree()
    root = t.add()
    work = [(root, n_tips)]
    tip = 0
    while work:
        v, n = work.pop()
        if n == 1:
            t.name[v] = f"t{tip}"
            tip += 1
            continue
        for m in ((n + 1) // 2, n // 2):
            c = t.add("", _blen(rng, mean_blen))
            t.children[v].append(c)
        for c, m in reversed(list(zip(t.children[v], ((n + 1) // 2, n // 2)))):
            work.append((c, m))
    return t


def yule_tree(n_tips: int, seed: int = 3, mean_blen: float = 0.05) -> Tree:
    """Random joining of uniformly chosen lineage pairs until two remain (configs #3, #5). Re-draws with
    seed+1 if a root child is a tip (the reference cannot handle that for K>1)."""
    while True:
        rng = np.random.default_rng(seed)
        # build bottom-up as nested tuples, then number top-down so that node 0 is the root
        lineages = [("t%d" % i,) for i in range(n_tips)]
        while len(lineages) > 2:
            i, j = rng.choice(len(lineages), size=2, replace=False)
            a, b = lineages[i], lineages[j]
            for k in sorted((i, j), reverse=True):
                lineages.pop(k)
            lineages.append((a, b))
        if all(len(x) == 2 for x in lineages):
            break
        seed += 1
    t = Tree()
    root = t.add()
    work = [(root, (lineages[0], lineages[1]))]
    while work:
        v, sub = work.pop()
        for s in sub:
            c = t.add(s[0] if len(s) == 1 else "", _blen(rng, mean_blen))
            t.children[v].append(c)
            if len(s) == 2:
                work.append((c, s))
    return t


def caterpillar_tree(n_tips: int, seed: int = 4, mean_blen: float = 0.02) -> Tree:
    """Caterpillar rooted on its middle internal edge: both root children are caterpillars (config #4)."""
    rng = np.random.default_rng(seed)
    t = Tree()
    root = t.add()
    tip = 0
    for n in ((n_tips + 1) // 2, n_tips // 2):
        v = t.add("", _blen(rng, mean_blen))
        t.children[root].append(v)
        remaining = n
        while remaining > 2:
            leaf = t.add(f"t{tip}", _blen(rng, mean_blen)); tip += 1
            inner = t.add("", _blen(rng, mean_blen))
            t.children[v] += [inner, leaf]
            v = inner
            remaining -= 1
        for _ in range(2):
            leaf = t.add(f"t{tip}", _blen(rng, mean_blen)); tip += 1
            t.children[v].append(leaf)
    return t


def transition_matrix(Eval, Evec, Ievc, t):
    t = max(t, 1e-9)
    return np.abs(Evec @ (np.exp(t * Eval)[:, None] * Ievc))


def simulate_alignment(tree: Tree, model, rates, n_sites: int, seed: int, gap_frac=0.01, x_frac=0.001) -> np.ndarray:
    """Evolve `n_sites` columns down `tree` under `model` (HostModel) with equiprobable rate classes.
    Returns uint8 states[tip][site], tips in node order; `gap_frac`/`x_frac` of the cells become missing."""
    rng = np.random.Generator(np.random.PCG64(seed))
    K = len(rates)
    cls = rng.integers(0, K, size=n_sites)
    cdf_pi = np.cumsum(model.pi / model.pi.sum())
    state = {0: np.minimum(np.searchsorted(cdf_pi, rng.random(n_sites)), 19).astype(np.uint8)}
    tips = []
    stack = [0]
    while stack:
        v = stack.pop()
        sv = state.pop(v)
        if not tree.children[v]:
            tips.append((v, sv))
            continue
        for c in tree.children[v]:
            sc = np.empty(n_sites, dtype=np.uint8)
            u = rng.random(n_sites)
            for k in range(K):
                sel = np.nonzero(cls == k)[0]
                if sel.size == 0:
                    continue
                P = transition_matrix(model.Eval, model.Evec, model.Ievc, tree.blen[c] * rates[k])
                cdf = np.cumsum(P / P.sum(axis=1, keepdims=True), axis=1)
                rows = cdf[sv[sel]]
                sc[sel] = np.minimum((rows < u[sel, None]).sum(axis=1), 19)
            state[c] = sc
            stack.append(c)
    tips.sort(key=lambda x: x[0])
    out = np.stack([s for _, s in tips])
    if gap_frac or x_frac:
        r = rng.random(out.shape)
        out[r < gap_frac + x_frac] = MISSING
    return out


def _sim_job(job):
    tree, model, rates, n, seed = job
    return simulate_alignment(tree, model, rates, n, seed)


def simulate_alignment_parallel(tree: Tree, model, rates, n_sites: int, seed: int, block=16384, procs=None) -> np.ndarray:
    """`n_sites` DISTINCT simulated columns: blocks of `block` columns, each with its own seed, spread over the host cores."""
    import multiprocessing as mp
    import os
    starts = list(range(0, n_sites, block))
    jobs = [(tree, model, rates, min(block, n_sites - c0), seed + 7919 * (i + 1)) for i, c0 in enumerate(starts)]
    procs = procs or max(1, min(len(jobs), os.cpu_count() or 1))
    if procs == 1 or len(jobs) == 1:
        parts = [_sim_job(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_sim_job, jobs)
    return np.concatenate(parts, axis=1)


def simulate_alignment_torch(tree: Tree, model, rates, n_sites: int, seed: int, device="cuda", gap_frac=0.01, x_frac=0.001,
                             block=1 << 17, out=None) -> np.ndarray:
    """The same generative process as `simulate_alignment` (root state from pi, one rate class per column, a child's state
    drawn from row `parent state` of P(t_child * rate), `gap_frac + x_frac` of the cells missing), vectorised over columns
    with torch ops on `device` — plumbing for synthetic DATA only (bench.py wants 10^6..10^7 distinct columns in seconds;
    the numpy version takes minutes). Different random stream than the numpy version, same distribution
    (tests/test_host.py::test_torch_simulator_matches_the_numpy_one). Columns come in blocks of `block`, block i seeded with
    (seed, i): the first columns of a longer run equal a shorter run with the same seed. Returns / fills uint8
    states[tip][site], tips in node order."""
    import torch
    K = len(rates)
    n_nodes = len(tree.children)
    tips = [v for v in range(n_nodes) if not tree.children[v]]
    row_of = {v: i for i, v in enumerate(tips)}
    if out is None:
        out = np.empty((len(tips), n_sites), dtype=np.uint8)
    # cumulative transition rows per (node, rate class): [n_nodes][K*20][20]
    cdf = np.zeros((n_nodes, K * 20, 20))
    for v in range(1, n_nodes):
        for k in range(K):
            P = transition_matrix(model.Eval, model.Evec, model.Ievc, tree.blen[v] * rates[k])
            cdf[v, k * 20:(k + 1) * 20] = np.cumsum(P / P.sum(axis=1, keepdims=True), axis=1)
    cdf_d = torch.from_numpy(cdf).to(device)
    cdf_pi = torch.from_numpy(np.cumsum(model.pi / model.pi.sum())).to(device)
    g = torch.Generator(device=device)
    miss = gap_frac + x_frac
    for bi, c0 in enumerate(range(0, n_sites, block)):
        n = min(block, n_sites - c0)
        g.manual_seed(int(seed) * 1_000_003 + bi)
        # (always draw a full block so that a ragged last block is a prefix of the full one)
        cls20 = torch.randint(0, K, (block,), generator=g, device=device)[:n] * 20
        root = torch.searchsorted(cdf_pi, torch.rand(block, generator=g, device=device, dtype=torch.float64)[:n]).clamp_(max=19)
        blk = torch.empty((len(tips), n), dtype=torch.uint8, device=device)
        stack = [(0, root)]
        while stack:
            v, sv = stack.pop()
            if not tree.children[v]:
                cell = sv.to(torch.uint8)
                if miss:
                    cell = torch.where(torch.rand(block, generator=g, device=device)[:n] < miss, torch.full_like(cell, MISSING), cell)
                blk[row_of[v]] = cell
                continue
            for c in tree.children[v]:
                rows = cdf_d[c][cls20 + sv]                                          # [n, 20]
                u = torch.rand(block, generator=g, device=device, dtype=torch.float64)[:n]
                stack.append((c, (rows < u[:, None]).sum(dim=1).clamp_(max=19)))
        out[:, c0:c0 + n] = blk.cpu().numpy()
    return out
