"""Host logic of the native path: the fused pruning schedule (subrecon_b200.cu::build_schedule) is executed here by a
small numpy stack machine with the oracle's matrices and must reproduce the oracle's per-column results. No GPU."""
import numpy as np
import pytest

from helpers import load_example, random_newick, rel_err, synthetic
from oracle import oracle
from subrecon_b200 import native, synth, treeio


@pytest.fixture(scope="module", autouse=True)
def built():
    native.build()


def run_schedule(p, cols, rescale_every=8, cherry_tables=True):
    """Interpret the step list exactly as prune_kernel + root_kernel do (same flags, same slot order)."""
    ft = p.flat
    steps, node_slot, depth = native.schedule_dump(ft, rescale_every, cherry_tables)
    slot_node = {int(s): v for v, s in enumerate(node_slot) if s >= 0}
    cherry_node = {-2 - int(s): v for v, s in enumerate(node_slot) if s <= -2}       # collapsed cherries by table index
    assert sorted(slot_node) == list(range(len(slot_node)))                    # every edge exactly once, densely numbered
    assert sorted(cherry_node) == list(range(len(cherry_node))) and (cherry_tables or not cherry_node)
    rc = ft.child_idx[ft.child_off[ft.root]:ft.child_off[ft.root + 1]]
    def kids_of(v):
        return list(ft.child_idx[ft.child_off[v]:ft.child_off[v + 1]])

    def is_leaf(v):
        return ft.child_off[v] == ft.child_off[v + 1]
    n_edges_in_tables = 0
    for v in cherry_node.values():
        kids = kids_of(v)
        assert len(kids) == 2 and v not in rc
        if all(is_leaf(x) for x in kids):
            n_edges_in_tables += 3                                           # cherry: two tip edges + its own
        else:                                                                # pitchfork: a cherry and a tip
            inner = [x for x in kids if not is_leaf(x)]
            assert len(inner) == 1 and all(is_leaf(y) for y in kids_of(inner[0])) and len(kids_of(inner[0])) == 2
            n_edges_in_tables += 5
    # every non-root edge is either in the matrix stream or inside a clade table
    assert len(slot_node) + n_edges_in_tables == ft.n_nodes - 1 - sum(1 for v in rc if ft.child_off[v] != ft.child_off[v + 1])
    row_off, acc_rows = {}, 0
    for cid in sorted(cherry_node):
        v = cherry_node[cid]
        row_off[v] = acc_rows
        acc_rows += 441 if all(is_leaf(x) for x in kids_of(v)) else 9261
    m = p.model
    K = len(p.rates)
    lnl, joint = [], []
    for c in cols:
        tot = np.zeros((20, 20))
        logs = []
        per_rate = []
        for k in range(K):
            def P_of(v):
                return oracle.transition_probs(m.Eval, m.Evec, m.Ievc, ft.blen[v] * p.rates[k])

            def tip(v, row):
                assert ft.leaf_row[v] == row
                s = p.states[row, c]
                if v in rc:                                                # tip directly below the root: identity table
                    return np.ones(20) if s >= 20 else np.eye(20)[s]
                P = P_of(v)
                return P.sum(axis=1) if s >= 20 else P[:, s]
            def clade(v, rows):
                # what K1b tabulates: the clade's own partial carried up its parent edge; `rows` = its tips' alignment
                # rows in table-index order (cherry tips in child order, then the pitchfork's third tip)
                kids = kids_of(v)
                if all(is_leaf(x) for x in kids):
                    a, b = kids
                    return P_of(v) @ (tip(a, rows[0]) * tip(b, rows[1]))
                inner = [x for x in kids if not is_leaf(x)][0]
                t = [x for x in kids if is_leaf(x)][0]
                return P_of(v) @ (clade(inner, rows[:2]) * tip(t, rows[2]))
            A, stack, cnt, slot, maxsp = None, [], 0, 0, 0
            parts = {}
            seen_cherries = 0
            for w, r0, r1, r2, rb, rb2, rc1, rc2, ro1, ro2, _, _ in steps:
                mv, n_pre, post = w & 3, (w >> 2) & 3, (w >> 5) & 1
                ps, ps2, tri1, tri2 = (w >> 10) & 3, (w >> 12) & 3, (w >> 14) & 1, (w >> 15) & 1
                assert ps <= n_pre + post and (ps2 == 0 or ps < ps2 <= n_pre + post)
                rows = [r0, r1, r2]
                mv_node = None
                if mv:
                    mv_node = slot_node[slot]; slot += 1
                li = 0

                def gather():
                    nonlocal slot, li, seen_cherries
                    li += 1
                    if ps == li or ps2 == li:
                        second = ps2 == li
                        v = cherry_node[seen_cherries]                         # tables are numbered in order of use
                        seen_cherries += 1
                        assert (ro2 if second else ro1) == row_off[v]
                        is_tri = not all(is_leaf(x) for x in kids_of(v))
                        assert is_tri == bool(tri2 if second else tri1)
                        return clade(v, [rows[li - 1], rb2 if second else rb, rc2 if second else rc1])
                    v = slot_node[slot]; slot += 1
                    return tip(v, rows[li - 1])
                for i in range(n_pre):
                    t = gather()
                    A = t if (i == 0 and (w & 16)) else A * t
                if mv:
                    A = P_of(mv_node) @ A
                    if mv == 2:
                        A = A * stack.pop()
                    if post:
                        A = A * gather()
                else:
                    assert not post
                if w & 64:
                    e = int(np.floor(np.log2(A.max())))
                    A = A * 2.0 ** (-e)
                    cnt += e
                if w & 128:
                    stack.append(A)
                    maxsp = max(maxsp, len(stack))
                if w & 256:
                    parts[0] = A
                if w & 512:
                    parts[1] = A
            assert not stack and slot == len(slot_node) and maxsp <= depth and seen_cherries == len(cherry_node)
            a, b = rc
            Pr = oracle.transition_probs(m.Eval, m.Evec, m.Ievc, (ft.blen[a] + ft.blen[b]) * p.rates[k])
            per_rate.append((m.pi * parts[0])[:, None] * Pr * parts[1][None, :])
            logs.append(cnt)
        emax = max(logs)
        for k in range(K):
            tot += per_rate[k] * 2.0 ** (logs[k] - emax)
        lnl.append(np.log(tot.sum()) + emax * np.log(2.0) - np.log(K))
        joint.append(tot / tot.sum())
    return np.array(lnl), np.array(joint), steps, depth


def test_schedule_on_example(example_problem):
    p = example_problem
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    cols = [0, 13, 40, 129]
    got_l, got_j, steps, depth = run_schedule(p, cols, cherry_tables=False)
    assert rel_err(got_l, lnl[cols]) < 1e-12 and rel_err(got_j, joint[cols]) < 1e-9
    assert depth == 2 and len(steps) < 34                       # fused: fewer steps than the 34 edges
    mv_steps = int(np.count_nonzero(steps[:, 0] & 3))
    assert mv_steps == 15                                       # one MV per internal edge below the root children (E_int = N-4)
    # with cherry tables every collapsed cherry's contraction is gone, the results are the same
    got_l, got_j, steps_c, depth_c = run_schedule(p, cols)
    assert rel_err(got_l, lnl[cols]) < 1e-12 and rel_err(got_j, joint[cols]) < 1e-9
    w_ = steps_c[:, 0]
    n_saved = int(sum(np.count_nonzero((w_ >> sh) & 3) for sh in (10, 12)) + np.count_nonzero(w_ & (1 << 14)) + np.count_nonzero(w_ & (1 << 15)))
    assert n_saved >= 4 and int(np.count_nonzero(w_ & 3)) == 15 - n_saved and depth_c <= depth


def test_schedule_shapes():
    class P:
        pass
    model = oracle.OracleModel("JTT")
    rates = np.array([0.4, 1.6])
    newicks = [
        "((a:0.1,b:0.2):0.1,(c:0.1,d:0.3):0.2);",                                        # two cherries
        "((a:0.2,b:0.01,f:0.3,(g:0.1,h:0.2):0.0,i:0.1,j:0.2,k:0.3):0.1,(c:0.3,d:0.05,e:0.6):0.02);",   # polytomies
        "(a:0.1,(b:0.2,c:0.3):0.1);",                                                    # tip as root child
        "(((a:0.1,b:0.1):0.1,(c:0.1,d:0.1):0.1):0.1,((e:0.1,f:0.1):0.1,((g:0.1,h:0.1):0.1,(i:0.2,j:0.3):0.1):0.1):0.1);",
    ]
    for nw in newicks:
        tree = treeio.parse_newick(nw)
        p = P()
        p.flat = treeio.flatten_tree(tree)
        n_tips = len(tree.leaves())
        rng = np.random.default_rng(7)
        p.states = rng.integers(0, 21, size=(n_tips, 6)).astype(np.uint8)
        p.states[p.states == 20] = 255
        p.model, p.rates = model, rates
        if "(a:0.1,(b" in nw:
            rates1 = np.array([1.0])
            p.rates = rates1                                   # the reference only handles this shape for K = 1
        lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
        for ct in (True, False):
            got_l, got_j, steps, depth = run_schedule(p, range(6), cherry_tables=ct)
            assert rel_err(got_l, lnl) < 1e-12 and rel_err(got_j, joint) < 1e-9, nw
            got_l, got_j, _, _ = run_schedule(p, range(2), rescale_every=1, cherry_tables=ct)
            assert rel_err(got_l, lnl[:2]) < 1e-12


@pytest.mark.parametrize("kind,n,depth_max", [("balanced", 64, 5), ("yule", 200, 6), ("caterpillar", 300, 0)])
def test_schedule_synthetic(kind, n, depth_max):
    class P:
        pass
    tree = {"balanced": synth.balanced_tree, "yule": synth.yule_tree, "caterpillar": synth.caterpillar_tree}[kind](n, seed=5)
    p = P()
    p.flat = treeio.flatten_tree(tree)
    p.model = oracle.OracleModel("WAG")
    p.rates = oracle.gamma_rates(2, 0.5)
    p.states = synth.simulate_alignment(tree, p.model, p.rates, 3, seed=9)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    got_l, got_j, steps, depth = run_schedule(p, range(3), cherry_tables=False)
    assert rel_err(got_l, lnl) < 1e-12 and rel_err(got_j, joint) < 1e-9
    assert depth <= depth_max
    n_mv = int(np.count_nonzero(steps[:, 0] & 3))
    assert len(steps) <= n_mv + 4 + (0 if kind != "balanced" else n // 2)      # nearly every step carries an MMA burst
    got_l, got_j, steps_c, depth_c = run_schedule(p, range(3))
    assert rel_err(got_l, lnl) < 1e-12 and rel_err(got_j, joint) < 1e-9 and depth_c <= depth
    w_ = steps_c[:, 0]
    # contractions saved: one per cherry table, two per pitchfork table
    n_saved = int(sum(np.count_nonzero((w_ >> sh) & 3) for sh in (10, 12)) + np.count_nonzero(w_ & (1 << 14)) + np.count_nonzero(w_ & (1 << 15)))
    n_mv_c = int(np.count_nonzero(w_ & 3))
    assert n_mv_c == n_mv - n_saved
    assert n_saved >= {"balanced": n // 2 - 2, "yule": n // 4, "caterpillar": 1}[kind]
    print(kind, "steps", len(steps), "->", len(steps_c), "contractions", n_mv, "->", n_mv_c)


def test_schedule_fuzz():
    """Random topologies (polytomies, tips as root children, zero branches), rate-class counts and rescale periods: the
    collapsed (clade tables) and the plain schedule both reproduce the oracle on the CPU stack machine."""
    class P:
        pass
    rng = np.random.default_rng(20261020)
    for it in range(30):
        n = int(rng.integers(3, 36))
        tree = treeio.parse_newick(random_newick(rng, n))
        p = P()
        p.flat = treeio.flatten_tree(tree)
        rc = p.flat.child_idx[p.flat.child_off[p.flat.root]:p.flat.child_off[p.flat.root + 1]]
        leaf_root = any(p.flat.child_off[v] == p.flat.child_off[v + 1] for v in rc)
        K = 1 if leaf_root else int(rng.integers(1, 4))           # the reference only handles a tip below the root for K = 1
        p.states = rng.integers(0, 22, size=(n, 3)).astype(np.uint8)
        p.states[p.states >= 20] = 255
        p.model = oracle.OracleModel(["WAG", "JTT", "Dayhoff", "BLOSUM62"][it % 4])
        p.rates = oracle.gamma_rates(K, 0.7) if K > 1 else np.array([1.0])
        lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
        for ct in (True, False):
            got_l, got_j, _, _ = run_schedule(p, range(3), rescale_every=int(rng.integers(1, 9)), cherry_tables=ct)
            assert np.max(np.abs(got_l - lnl) / np.maximum(1.0, np.abs(lnl))) < 1e-11 and rel_err(got_j, joint) < 1e-9
