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
            for w, r0, r1, r2, rb0, rb1, rb2, rc0, rc1, rc2, ro0, ro1, ro2, _, _, _ in steps:
                mv, n_pre, post = w & 3, (w >> 2) & 3, (w >> 5) & 1
                cl, tri = (w >> 10) & 7, (w >> 13) & 7
                assert cl < (1 << (n_pre + post)) and (tri & ~cl) == 0
                rows, rows_b, rows_c, rows_off = [r0, r1, r2], [rb0, rb1, rb2], [rc0, rc1, rc2], [ro0, ro1, ro2]
                mv_node = None
                if mv:
                    mv_node = slot_node[slot]; slot += 1
                li = 0

                def gather():
                    nonlocal slot, li, seen_cherries
                    li += 1
                    if (cl >> (li - 1)) & 1:
                        v = cherry_node[seen_cherries]                         # tables are numbered in order of use
                        seen_cherries += 1
                        assert rows_off[li - 1] == row_off[v]
                        is_tri = not all(is_leaf(x) for x in kids_of(v))
                        assert is_tri == bool((tri >> (li - 1)) & 1)
                        return clade(v, [rows[li - 1], rows_b[li - 1], rows_c[li - 1]])
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
    n_saved = int(sum(np.count_nonzero((w_ >> sh) & 1) for sh in (10, 11, 12, 13, 14, 15)))
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
    n_saved = int(sum(np.count_nonzero((w_ >> sh) & 1) for sh in (10, 11, 12, 13, 14, 15)))
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


# ------------------------------------------------------------------------------------------------ the records
FRAG_BYTES, TABLE_BYTES, CLADE_ROW_BYTES = 4096, 4032, 192
SHAPES = {0: (0, False, 1, 0), 1: (0, False, 2, 0), 2: (0, False, 1, 1), 3: (2, True, 1, 0), 4: (2, True, 1, 1), 5: (2, True, 2, 0),
          6: (0, False, 2, 1)}                                    # shape id -> (n_pre, pre starts a partial, mv, post)


def run_records(p, cols, rescale_every=8, cherry_tables=True, smem_levels=99):
    """Interpret the two record streams exactly as prune_kernel does (csrc/prune.cuh): the producer record says which
    bytes of the matrix stream and which index rows land in the stage, the consumer record says what to do with them.
    Independent of the step words: only byte offsets, shape ids, flags and stack levels are used."""
    ft = p.flat
    crec, prec, cgs = native.schedule_records(ft, rescale_every, cherry_tables)
    steps, node_slot, depth = native.schedule_dump(ft, rescale_every, cherry_tables)
    assert len(crec) == len(steps)
    rc = ft.child_idx[ft.child_off[ft.root]:ft.child_off[ft.root + 1]]

    def kids_of(v):
        return list(ft.child_idx[ft.child_off[v]:ft.child_off[v + 1]])

    def is_leaf(v):
        return ft.child_off[v] == ft.child_off[v + 1]
    # the matrix stream: slots in order, an internal edge's MMA fragments (4096 B) or a tip's gather table (4032 B)
    stream_node, off = {}, 0
    for slot, v in sorted((int(s), v) for v, s in enumerate(node_slot) if s >= 0):
        stream_node[off] = v
        off += TABLE_BYTES if is_leaf(v) else FRAG_BYTES
    # the clade tables: in order of first use, 441 or 9261 rows of 192 B
    table_node, toff = {}, 0
    for cid, v in sorted((-2 - int(s), v) for v, s in enumerate(node_slot) if s <= -2):
        table_node[toff] = v
        toff += CLADE_ROW_BYTES * (441 if all(is_leaf(x) for x in kids_of(v)) else 9261)
    m, K = p.model, len(p.rates)
    st = np.where(p.states >= 20, 20, p.states).astype(np.int64)            # K0: codes clamped to 0..20
    lnl, joint = [], []
    for c in cols:
        per_rate, logs = [], []
        for k in range(K):
            def P_of(v):
                return oracle.transition_probs(m.Eval, m.Evec, m.Ievc, ft.blen[v] * p.rates[k])

            def tip_table(v):                                               # K1: rows = states 0..19 + missing
                if v in rc:
                    return np.concatenate([np.eye(20), np.ones((1, 20))])
                P = P_of(v)
                return np.concatenate([P.T, P.sum(axis=1)[None, :]])

            def clade_table_row(v, ix):                                     # K1b, row ix
                kids = kids_of(v)
                if all(is_leaf(x) for x in kids):
                    a, b = kids
                    return P_of(v) @ (tip_table(a)[ix // 21] * tip_table(b)[ix % 21])
                inner = [x for x in kids if not is_leaf(x)][0]
                t = [x for x in kids if is_leaf(x)][0]
                a, b = kids_of(inner)
                u = P_of(inner) @ (tip_table(a)[ix // 441] * tip_table(b)[(ix // 21) % 21])
                return P_of(v) @ (u * tip_table(t)[ix % 21])
            A, cnt, stack, parts, consumed = None, 0, {}, {}, 0
            for i in range(len(crec)):
                x, offs = int(crec[i, 0]), [int(v) for v in crec[i, 1:4]]
                s_off, s_bytes, rk, rows = int(prec[i, 0]), int(prec[i, 1]), int(prec[i, 2]), [int(v) for v in prec[i, 3:6]]
                assert s_off == consumed                                    # the steps eat the stream front to back
                consumed += s_bytes
                n_rows, kinds = rk & 3, rk >> 2
                shape = x & 7
                n_pre, set0, mv, post = (x >> 20) & 3, bool(x & (1 << 22)), (x >> 23) & 3, (x >> 25) & 1
                if shape != 7:
                    assert SHAPES[shape] == (n_pre, set0, mv, post)         # the specialised body does what the bits say
                assert n_rows == n_pre + post and kinds == (x >> 7) & 7
                assert s_bytes == (FRAG_BYTES if mv else 0) + TABLE_BYTES * (n_rows - bin(kinds).count("1"))

                def gather(slot):
                    if (x >> (7 + slot)) & 1:
                        ra, rb, rcc, _ = (int(v) for v in cgs[rows[slot]])
                        v = table_node[offs[slot]]
                        ix = st[ra, c] * 21 + st[rb, c]
                        if rcc >= 0:
                            ix = ix * 21 + st[rcc, c]
                        assert (rcc >= 0) == (not all(is_leaf(q) for q in kids_of(v)))
                        return clade_table_row(v, ix)
                    assert offs[slot] < s_bytes and offs[slot] >= (FRAG_BYTES if mv else 0)
                    v = stream_node[s_off + offs[slot]]
                    assert is_leaf(v) and ft.leaf_row[v] == rows[slot]
                    return tip_table(v)[st[rows[slot], c]]
                for g in range(n_pre):
                    t = gather(g)
                    A = t if (g == 0 and set0) else A * t
                if mv:
                    v = stream_node[s_off]
                    assert not is_leaf(v)
                    A = P_of(v) @ A
                    if mv == 2:
                        A = A * stack.pop((x >> 10) & 31)
                    if post:
                        A = A * gather(n_pre)
                if x & 8:
                    e = int(np.floor(np.log2(A.max())))
                    A, cnt = A * 2.0 ** (-e), cnt + e
                if x & 16:
                    lvl = (x >> 15) & 31
                    assert lvl not in stack and lvl < max(depth, 1)
                    stack[lvl] = A
                if x & 32:
                    parts[0] = A
                if x & 64:
                    parts[1] = A
            assert not stack and consumed == off
            a, b = rc
            Pr = oracle.transition_probs(m.Eval, m.Evec, m.Ievc, (ft.blen[a] + ft.blen[b]) * p.rates[k])
            per_rate.append((m.pi * parts[0])[:, None] * Pr * parts[1][None, :])
            logs.append(cnt)
        emax = max(logs)
        tot = sum(per_rate[k] * 2.0 ** (logs[k] - emax) for k in range(K))
        lnl.append(np.log(tot.sum()) + emax * np.log(2.0) - np.log(K))
        joint.append(tot / tot.sum())
    hist = np.bincount(crec[:, 0] & 7, minlength=8)
    return np.array(lnl), np.array(joint), hist


def test_records_reproduce_the_oracle(example_problem):
    p = example_problem
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    cols = [0, 13, 40, 129]
    for ct in (True, False):
        got_l, got_j, hist = run_records(p, cols, cherry_tables=ct)
        assert rel_err(got_l, lnl[cols]) < 1e-12 and rel_err(got_j, joint[cols]) < 1e-9
        assert hist[7] <= 2                                        # nearly every step takes a specialised body


def test_records_fuzz():
    class P:
        pass
    rng = np.random.default_rng(20261021)
    general = total = 0
    for it in range(30):
        n = int(rng.integers(3, 40))
        tree = treeio.parse_newick(random_newick(rng, n))
        p = P()
        p.flat = treeio.flatten_tree(tree)
        rc = p.flat.child_idx[p.flat.child_off[p.flat.root]:p.flat.child_off[p.flat.root + 1]]
        leaf_root = any(p.flat.child_off[v] == p.flat.child_off[v + 1] for v in rc)
        K = 1 if leaf_root else int(rng.integers(1, 4))
        p.states = rng.integers(0, 22, size=(n, 3)).astype(np.uint8)
        p.states[p.states >= 20] = 255
        p.model = oracle.OracleModel(["WAG", "JTT", "Dayhoff", "BLOSUM62"][it % 4])
        p.rates = oracle.gamma_rates(K, 0.7) if K > 1 else np.array([1.0])
        lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
        for ct in (True, False):
            got_l, got_j, hist = run_records(p, range(3), rescale_every=int(rng.integers(1, 9)), cherry_tables=ct)
            assert np.max(np.abs(got_l - lnl) / np.maximum(1.0, np.abs(lnl))) < 1e-11 and rel_err(got_j, joint) < 1e-9
            general += int(hist[7]); total += int(hist.sum())
    assert general < total                                         # both kinds of body are exercised


def test_records_on_the_bench_trees():
    """Shape histogram of the BASELINE trees: the specialised bodies must cover nearly all steps (the general body is
    ~2x the instructions per step)."""
    for tree, min_cover in ((synth.yule_tree(1000, seed=3), 0.9), (synth.caterpillar_tree(5000, seed=4), 0.99),
                            (synth.yule_tree(2000, seed=5), 0.9), (synth.balanced_tree(100, seed=2), 0.9)):
        crec, prec, cgs = native.schedule_records(treeio.flatten_tree(tree))
        hist = np.bincount(crec[:, 0] & 7, minlength=8)
        assert 1.0 - hist[7] / hist.sum() >= min_cover, hist


def test_tree_validation_rejects_non_trees():
    """a cycle or a node with two parents must come back as an error, not abort the process."""
    import ctypes
    L = native.lib()

    def dump(off, idx, n, root=0):
        off = np.asarray(off, dtype=np.int32); idx = np.asarray(idx, dtype=np.int32)
        bl = np.full(n, 0.1); lr = np.arange(n, dtype=np.int32)
        steps = np.zeros((n + 8, 12), dtype=np.int32)
        ns = ctypes.c_int64()
        I32, D = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
        return L.sr_schedule_dump(n, off.ctypes.data_as(I32), idx.ctypes.data_as(I32), bl.ctypes.data_as(D), lr.ctypes.data_as(I32),
                                  root, 8, 1, n + 8, steps.ctypes.data_as(I32), None, ctypes.byref(ns), None)
    assert dump([0, 2, 4, 4, 4, 4], [1, 2, 3, 4], 5) == 0                    # a proper tree
    assert dump([0, 2, 4, 4, 4, 4], [1, 2, 1, 2], 5) == -1                   # cycle 1 -> 1, 2 (the advisor's example)
    assert dump([0, 2, 4, 4, 4, 4], [1, 2, 3, 3], 5) == -1                   # node 3 twice, node 4 orphaned
    assert dump([0, 2, 4, 4, 6, 6, 6], [1, 2, 3, 4, 3, 5], 6) == -1          # offsets longer than n_nodes + 1 / 3 its own child
    assert dump([0, 2, 1, 4, 4, 4], [1, 2, 3, 4], 5) == -1                   # non-monotone offsets
    assert dump([0, 2, 4, 4, 4, 4], [1, 9, 3, 4], 5) == -1                   # child index out of range
