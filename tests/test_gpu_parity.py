"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): site log-likelihoods and all 400 probabilities within 1e-9 relative in FP64
(entries below 1e-300 compared absolutely). At BASELINE's full sizes the oracle is too slow, so those cases use
size-independent properties: posteriors sum to 1, results are independent of where a column sits in the batch /
tile / chunk (bit-identical), streamed == resident, and the compact scan matches numpy on the full matrix.
"""
import os

import numpy as np
import pytest

from helpers import lnl_err, load_example, random_newick, rel_err, synthetic
from oracle import oracle
from subrecon_b200 import native, recon, treeio

pytestmark = pytest.mark.gpu
TOL = 1e-9
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx():
    c = native.Context(0)
    yield c
    c.close()


def setup(ctx, p):
    ctx.set_option("chunk_cols", 1 << 20)
    ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi)
    ctx.set_rates(p.rates)
    ctx.set_tree(p.flat)
    ctx.set_alignment(p.states)


def check(ctx, p, lnl, joint, **kw):
    r = ctx.run(**kw)
    assert rel_err(r["lnl"], lnl) < TOL
    assert rel_err(r["joint"], joint) < TOL
    assert np.max(np.abs(r["joint"].sum(axis=(1, 2)) - 1.0)) < 1e-12       # checkSumToOne, at 1e-12 not 1e-6
    return r


def test_cfg1_lysozyme_vs_oracle_and_golden(ctx, example_problem):
    p = example_problem
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    setup(ctx, p)
    r = check(ctx, p, lnl, joint)
    g = np.load(os.path.join(HERE, "golden", "lysozyme_golden.npz"))       # PAL-bytecode golden
    assert rel_err(r["lnl"], g["lnl"]) < TOL and rel_err(r["joint"], g["joint"]) < TOL
    tot = 0.0
    for x in r["lnl"]:
        tot += x
    assert abs(tot - (-673.1980928246)) < 1e-9


def test_cfg1_printed_output_identical(ctx, example_problem):
    """Identical printed output at the default -sd and threshold, and with -threshold 0.0 -verbose."""
    p = example_problem
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    setup(ctx, p)
    r = ctx.run()
    for thr, verbose in ((0.5, False), (0.0, True)):
        want = recon.run_columns([recon.SiteResult(i, lnl[i], joint[i], thr) for i in range(len(lnl))], verbose, thr)
        got = recon.run_columns([recon.SiteResult(i, r["lnl"][i], r["joint"][i], thr) for i in range(len(lnl))], verbose, thr)
        assert got[:-1] == want[:-1]                                      # every Result line byte-identical
        assert abs(float(got[-1].split()[-1]) - float(want[-1].split()[-1])) < 1e-8   # Total lnL compared numerically
    lines = recon.run_columns([recon.SiteResult(i, r["lnl"][i], r["joint"][i]) for i in range(130)])
    assert [ln.split("\t")[1] for ln in lines if ln.startswith("Result")] == ["14", "21", "23", "41", "50", "62", "87", "101"]


def test_device_built_matrices_match_pal(ctx, example_problem):
    """K1 against `model.getTransitionProbabilities` (JointBranchReconstruction.java:217)."""
    p = example_problem
    ctx.set_option("cherry_tables", 0)                                    # every edge keeps its own matrix in the stream
    try:
        setup(ctx, p)
        _check_matrices(ctx, p)
    finally:
        ctx.set_option("cherry_tables", 1)


def _check_matrices(ctx, p):
    worst = 0.0
    for v in range(1, p.flat.n_nodes):
        if v in p.tree.children[0]:
            continue                                                      # root edges are merged (:99)
        for k in range(len(p.rates)):
            P = ctx.debug_pmatrix(v, k)
            Po = oracle.transition_probs(p.model.Eval, p.model.Evec, p.model.Ievc, p.flat.blen[v] * p.rates[k])
            worst = max(worst, float(np.max(np.abs(P - Po) / Po)))
    a, b = p.tree.children[0]
    for k in range(len(p.rates)):
        Po = oracle.transition_probs(p.model.Eval, p.model.Evec, p.model.Ievc, (p.flat.blen[a] + p.flat.blen[b]) * p.rates[k])
        worst = max(worst, float(np.max(np.abs(ctx.debug_pmatrix(0, k) - Po) / Po)))
    # same operation order as PAL; only CUDA's exp() (<= 1 ulp) can differ. Off-diagonal entries of the 1e-6
    # branches are ~1e-8 and emerge from cancelling O(1) terms, so 1 ulp there is ~1e-10 relative.
    assert worst < 1e-9


def test_cherry_tables_match_pal(ctx, example_problem):
    """K1b: the table of a collapsed clade against the reference's own arithmetic for that clade
    (JointBranchReconstruction.java:208-230 with PAL's matrices): all 21 x 21 (x 21) state combinations, missing data
    included — cherries and pitchforks."""
    p = example_problem
    setup(ctx, p)
    ft = p.flat
    _, node_slot, _ = native.schedule_dump(ft)
    clades = [v for v, s in enumerate(node_slot) if s <= -2]

    def kids(v):
        return list(ft.child_idx[ft.child_off[v]:ft.child_off[v + 1]])

    def leaf(v):
        return ft.child_off[v] == ft.child_off[v + 1]
    worst, kinds = 0.0, set()
    for v in clades:
        with pytest.raises(native.SrError):
            ctx.debug_pmatrix(v, 0)                                       # its matrices are not in the stream any more
        for k in (0, len(p.rates) - 1):
            def P(x):
                return oracle.transition_probs(p.model.Eval, p.model.Evec, p.model.Ievc, ft.blen[x] * p.rates[k])

            def T(x):                                                     # [state or missing][k]
                return np.concatenate([P(x).T, P(x).sum(axis=1)[None, :]])
            got = ctx.debug_cherry_table(v, k)
            if all(leaf(x) for x in kids(v)):
                a, b = kids(v)
                want = np.einsum("ik,ak,bk->abi", P(v), T(a), T(b))
            else:
                ce = [x for x in kids(v) if not leaf(x)][0]
                t = [x for x in kids(v) if leaf(x)][0]
                a, b = kids(ce)
                u = np.einsum("km,am,bm->abk", P(ce), T(a), T(b))
                want = np.einsum("ik,abk,tk->abti", P(v), u, T(t))
            assert got.shape == want.shape
            kinds.add(got.ndim)
            worst = max(worst, float(np.max(np.abs(got - want) / want)))
    assert worst < 1e-9 and kinds == {3, 4}
    # and the ways of pruning agree far inside the tolerance
    with_tables = ctx.run()
    try:
        for opt in ("cherry_pitchforks", "cherry_tables"):
            ctx.set_option(opt, 0)
            without = ctx.run()
            assert rel_err(with_tables["lnl"], without["lnl"]) < 1e-12 and rel_err(with_tables["joint"], without["joint"]) < 1e-10
    finally:
        ctx.set_option("cherry_tables", 1)
        ctx.set_option("cherry_pitchforks", 1)


def test_cfg2_full(ctx):
    p = synthetic(2, 10_000)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    setup(ctx, p)
    check(ctx, p, lnl, joint)


def test_cfg3_subset_and_full_size_properties(ctx):
    p = synthetic(3, 4096)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    setup(ctx, p)
    check(ctx, p, lnl, joint)
    # full size: 1,000 taxa x 1M sites = the 4,096 simulated columns repeated; every copy must be bit-identical
    big = np.tile(p.states, (1, 256))[:, :1_000_000]
    ctx.set_alignment(big)
    r = ctx.run(want_joint=False, want_compact=True, threshold=0.5)
    lnl_big = r["lnl"]
    assert lnl_big.shape == (1_000_000,)
    ref = np.tile(lnl_big[:4096], 256)[:1_000_000]
    assert np.array_equal(lnl_big, ref)
    assert rel_err(lnl_big[:4096], lnl) < TOL
    c = r["compact"]
    assert np.array_equal(c["lnl"], lnl_big)
    assert rel_err(c["max_prob"][:4096], joint.max(axis=(1, 2))) < TOL


def test_cfg4_deep_caterpillar_rescaling(ctx):
    p = synthetic(4, 1024)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    assert lnl.min() < -3000                                             # log-scale sums far beyond double range
    setup(ctx, p)
    check(ctx, p, lnl, joint)
    ctx.set_option("rescale_every", 3)
    ctx.set_tree(p.flat)
    check(ctx, p, lnl, joint)
    ctx.set_option("rescale_every", 8)


@pytest.mark.parametrize("shape", ["caterpillar", "balanced"])
def test_adversarial_underflow(ctx, shape):
    """The fastest a partial can shrink: every branch at the 1e-9 clamp, neighbouring tips always in conflicting states,
    and those states carry floor-level frequencies — each merge costs ~20 orders of magnitude. The power-of-two rescale
    every `rescale_every` edges must keep every partial inside FP64 (DESIGN.md, K2 'Rescaling': with the default of 8 there
    is head-room; 12 still passes this case, 16 does not)."""
    from subrecon_b200 import synth
    n = 96
    tree = (synth.caterpillar_tree if shape == "caterpillar" else synth.balanced_tree)(n, seed=1)
    for v in range(len(tree.blen)):
        tree.blen[v] = 0.0                                               # -> clamped to 1e-9 times the rate
    flat = treeio.flatten_tree(tree)
    freqs = np.full(20, 1.0)
    freqs[[3, 7, 11]] = 1e-12                                            # PAL's clean-up lifts these to its floor
    model = oracle.OracleModel("WAG", frequencies=freqs / freqs.sum())
    rates = oracle.gamma_rates(2, 0.5)
    rng = np.random.default_rng(5)
    states = np.empty((n, 300), dtype=np.uint8)
    rare = np.array([3, 7, 11], dtype=np.uint8)
    for col in range(300):
        states[:, col] = rare[(np.arange(n) + rng.integers(0, 3)) % 3] if col % 2 else rare[rng.integers(0, 3, size=n)]
    lnl, joint = oracle.run(flat, states, model, rates)
    assert lnl.min() < -1500 and np.isfinite(lnl).all()
    ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
    ctx.set_rates(rates)
    ctx.set_tree(flat)
    ctx.set_alignment(states)
    r = ctx.run()
    assert np.isfinite(r["lnl"]).all() and np.isfinite(r["joint"]).all()
    assert lnl_err(r["lnl"], lnl) < TOL and rel_err(r["joint"], joint) < TOL


def test_cfg4_full_size_properties(ctx):
    """BASELINE config #4 at full size (5,000-taxon caterpillar x 200k sites, K=8): 1,024 simulated columns repeated;
    every copy bit-identical, the first block equal to the oracle, posteriors sum to 1."""
    p = synthetic(4, 1024)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    setup(ctx, p)
    big = np.tile(p.states, (1, 196))[:, :200_000]
    ctx.set_alignment(big)
    for _ in range(4):      # repeated: this shape (matrix stream larger than L2) is the one that exposed the ring-stage races
        r = ctx.run(want_joint=False, want_compact=True, threshold=0.5)
        assert r["lnl"].shape == (200_000,)
        assert np.array_equal(r["lnl"], np.tile(r["lnl"][:1024], 196)[:200_000])
    assert rel_err(r["lnl"][:1024], lnl) < TOL
    assert rel_err(r["compact"]["max_prob"][:1024], joint.max(axis=(1, 2))) < TOL
    tail = ctx.run(site0=199_000, n=1000)
    assert np.max(np.abs(tail["joint"].sum(axis=(1, 2)) - 1.0)) < 1e-12
    assert np.array_equal(tail["lnl"], r["lnl"][199_000:])


def test_cfg5_full_size_properties(ctx):
    """BASELINE config #5 at full size (2,000 taxa x 8M sites, Dayhoff): streamed from host memory in column chunks
    with the compact result; 2,048 simulated columns repeated, every copy bit-identical, first block equal to the oracle."""
    import psutil
    if psutil.virtual_memory().available < 40 * 2**30:
        pytest.skip("needs ~17 GB of host memory for the 2,000 x 8M state matrix")
    p = synthetic(5, 2048)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    setup(ctx, p)
    L = 8_000_000
    big = np.empty((p.states.shape[0], L), dtype=np.uint8)
    for c0 in range(0, L, 2048):
        w = min(2048, L - c0)
        big[:, c0:c0 + w] = p.states[:, :w]
    r = ctx.run_streamed(big, want_joint=False, want_compact=True, threshold=0.5)
    del big
    assert r["lnl"].shape == (L,)
    ref = np.tile(r["lnl"][:2048], L // 2048 + 1)[:L]
    assert np.array_equal(r["lnl"], ref)
    assert rel_err(r["lnl"][:2048], lnl) < TOL
    assert np.array_equal(r["compact"]["lnl"], r["lnl"])
    assert rel_err(r["compact"]["max_ii_prob"][:2048], np.diagonal(joint, axis1=1, axis2=2).max(axis=1)) < TOL


def test_cfg5_subset(ctx):
    p = synthetic(5, 2048)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    setup(ctx, p)
    check(ctx, p, lnl, joint)


@pytest.mark.parametrize("opts", [dict(k_fastest=1), dict(cherry_tables=0), dict(cherry_tables=0, k_fastest=1), dict(cherry_table_mb=1),
                                  dict(cherry_table_mb=40), dict(cherry_pairs=0), dict(cherry_pitchforks=0),
                                  dict(cherry_pitchforks=0, cherry_pairs=0), dict(rescale_every=1), dict(rescale_every=12)])
def test_kernel_variants_agree(ctx, opts):
    """Every schedule / work-order option of the pruning kernel (sr_set_option) gives oracle-exact results."""
    p = synthetic(3, 512)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    setup(ctx, p)
    try:
        for k, v in opts.items():
            ctx.set_option(k, v)
        r = ctx.run()
        assert rel_err(r["lnl"], lnl) < TOL and rel_err(r["joint"], joint) < TOL
    finally:
        for k, v in dict(k_fastest=0, cherry_tables=1, cherry_table_mb=4096, cherry_pairs=1, cherry_pitchforks=1, rescale_every=8).items():
            ctx.set_option(k, v)


def test_ragged_offsets_and_chunks(ctx, example_problem):
    p = example_problem
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates)
    setup(ctx, p)
    for s0, n in ((0, 1), (129, 1), (37, 5), (1, 129), (16, 17), (0, 0)):
        r = ctx.run(site0=s0, n=n)
        assert r["lnl"].shape == (n,)
        if n:
            assert rel_err(r["lnl"], lnl[s0:s0 + n]) < TOL and rel_err(r["joint"], joint[s0:s0 + n]) < TOL
    # chunked pipeline (chunks of 256 columns) and the streamed path give bit-identical results to one launch
    big = np.tile(p.states, (1, 9))
    ctx.set_alignment(big)
    one = ctx.run()
    ctx.set_option("chunk_cols", 256)
    many = ctx.run()
    st = ctx.run_streamed(big)
    part = ctx.run_streamed(big, site0=131, n=700)
    ctx.set_option("chunk_cols", 1 << 20)
    assert np.array_equal(one["lnl"], many["lnl"]) and np.array_equal(one["joint"], many["joint"])
    assert np.array_equal(one["lnl"], st["lnl"]) and np.array_equal(one["joint"], st["joint"])
    assert np.array_equal(part["lnl"], one["lnl"][131:831]) and np.array_equal(part["joint"], one["joint"][131:831])
    assert rel_err(one["lnl"], np.tile(lnl, 9)) < TOL
    # the streamed pipeline's ramp of short chunks at both ends (needs >= 8 kernel waves per chunk to engage) and an
    # odd-sized middle remainder: same bits as uniform chunks and as the resident path
    wide = np.tile(p.states, (1, 1901))[:, :247001]
    try:
        ctx.set_option("stream_chunk_cols", 60000)
        tapered = ctx.run_streamed(wide, want_joint=True, want_compact=True, threshold=0.3)
        ctx.set_option("stream_taper", 0)
        uniform = ctx.run_streamed(wide, want_joint=True, want_compact=True, threshold=0.3)
    finally:
        ctx.set_option("stream_taper", 1)
        ctx.set_option("stream_chunk_cols", 1 << 17)
    for k in ("lnl", "joint"):
        assert np.array_equal(tapered[k], uniform[k])
        assert np.array_equal(tapered[k][:1170], one[k]) and np.array_equal(tapered[k][-131:], tapered[k][:131])
    assert tapered["compact"].tobytes() == uniform["compact"].tobytes()


def test_residue_letters_encoded_on_device(ctx, example_problem):
    """with "input_chars" the library takes the alignment's letters and applies PAL's char -> state map
    (App. B.6) on the device. Same bits as the host-encoded path, resident and streamed, for every possible byte."""
    p = example_problem
    names, seqs = treeio.read_fasta(os.path.join(HERE, "golden", "example", "prot.fasta"))
    letters = treeio.alignment_chars(seqs)
    assert np.array_equal(treeio.encode_alignment(seqs), p.states)
    setup(ctx, p)
    ref = ctx.run()
    # every byte value, both cases, gaps and ambiguity codes: 256 extra columns cycling through 0..255 on each row
    rng = np.random.default_rng(11)
    junk = np.stack([rng.permutation(256).astype(np.uint8) for _ in range(letters.shape[0])])
    lower = np.frombuffer(b"".join(s.lower().encode("latin-1") for s in seqs), dtype=np.uint8).reshape(letters.shape)
    wide = np.concatenate([letters, junk, lower], axis=1)
    lut = treeio._STATE_LUT
    ctx.set_alignment(lut[wide])
    want = ctx.run()
    try:
        ctx.set_option("input_chars", 1)
        ctx.set_alignment(wide)
        got = ctx.run()
        st = ctx.run_streamed(wide)
        part = ctx.run_streamed(wide, site0=100, n=300)
    finally:
        ctx.set_option("input_chars", 0)
    for k in ("lnl", "joint"):
        assert np.array_equal(got[k], want[k]) and np.array_equal(st[k], want[k]) and np.array_equal(part[k], want[k][100:400])
        assert np.array_equal(got[k][:130], ref[k]) and np.array_equal(got[k][-130:], ref[k])
    lnl, _ = oracle.run(p.flat, np.ascontiguousarray(lut[wide]), p.model, p.rates)
    assert lnl_err(got["lnl"], lnl) < TOL             # (junk columns can be all-missing: lnL ~ 0)


def test_column_pattern_compression(ctx, example_problem):
    """compute each distinct column once. Results must be bit-identical to the plain path."""
    p = example_problem
    setup(ctx, p)
    plain = ctx.run(want_joint=True, want_compact=True, threshold=0.3)
    try:
        ctx.set_option("compress_patterns", 1)
        ctx.profile(reset=True)
        comp = ctx.run(want_joint=True, want_compact=True, threshold=0.3)
        prof = ctx.profile(reset=True)
        assert prof["columns_in"] == 130 and prof["columns_unique"] == 51          # 51 distinct columns
        for k in ("lnl", "joint"):
            assert np.array_equal(plain[k], comp[k])
        assert plain["compact"].tobytes() == comp["compact"].tobytes()
        # ragged window + a big batch of repeats + the streamed pipeline
        w = ctx.run(site0=7, n=101)
        assert np.array_equal(w["lnl"], plain["lnl"][7:108]) and np.array_equal(w["joint"], plain["joint"][7:108])
        big = np.tile(p.states, (1, 300))
        ctx.set_alignment(big)
        ctx.profile(reset=True)
        r = ctx.run(want_joint=False)
        prof = ctx.profile(reset=True)
        assert prof["columns_unique"] == 51 and prof["columns_in"] == 39000
        assert np.array_equal(r["lnl"], np.tile(plain["lnl"], 300))
        st = ctx.run_streamed(big)
        assert np.array_equal(st["lnl"], r["lnl"]) and np.array_equal(st["joint"][:130], plain["joint"])
        # Equal hashes are verified against the bytes. Mode 2 hashes only the first two taxa, so different columns collide:
        # the byte compare must notice and the block is recomputed column by column — same bits, nothing saved.
        ctx.set_option("compress_patterns", 2)
        ctx.set_alignment(p.states)
        ctx.profile(reset=True)
        coll = ctx.run(want_joint=True, want_compact=True, threshold=0.3)
        prof = ctx.profile(reset=True)
        assert len({bytes(c) for c in p.states[:2].T}) < 51                      # the weak hash really merges distinct columns
        assert prof["columns_in"] == 130 and prof["columns_unique"] == 130
        assert np.array_equal(coll["lnl"], plain["lnl"]) and np.array_equal(coll["joint"], plain["joint"])
        assert coll["compact"].tobytes() == plain["compact"].tobytes()
    finally:
        ctx.set_option("compress_patterns", 0)


def test_pinned_buffers(ctx, example_problem):
    p = example_problem
    setup(ctx, p)
    n = p.states.shape[1]
    hl, hj = native.PinnedArray(n, np.float64), native.PinnedArray((n, 20, 20), np.float64)
    hs = native.PinnedArray(p.states.shape, np.uint8)
    hs.array[:] = p.states
    r = ctx.run_streamed(hs.array, out={"lnl": hl.array, "joint": hj.array})
    ref = ctx.run()
    assert np.array_equal(r["lnl"], ref["lnl"]) and np.array_equal(hj.array, ref["joint"])
    for h in (hl, hj, hs):
        h.free()


def test_polytomy_missing_data_custom_rates(ctx):
    tree = treeio.parse_newick("((a:0.2,b:0.01,f:0.3,(g:0.1,h:0.2):0.0):0.1,(c:0.3,d:0.05,e:0.6):0.02);")   # zero branch -> 1e-9 clamp
    flat = treeio.flatten_tree(tree)
    model = oracle.OracleModel("BLOSUM62")
    states = treeio.encode_alignment(["AV-W", "AV-W", "L--C", "LI-C", "KI-Y", "K?-y", "xV-W", "AB-W"])
    # K = 16 is the most the root kernel stages in shared memory, K = 20 takes its global-memory variant
    for rates in (np.array([1.0]), np.array([0.3, 1.7]), np.array([0.0, 0.5, 2.5]), oracle.gamma_rates(8, 0.5),
                  oracle.gamma_rates(16, 0.7), oracle.gamma_rates(20, 1.3)):
        lnl, joint = oracle.run(flat, states, model, rates)
        ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
        ctx.set_rates(rates)
        ctx.set_tree(flat)
        ctx.set_alignment(states)
        r = ctx.run()
        assert lnl_err(r["lnl"], lnl) < TOL and rel_err(r["joint"], joint) < TOL
    assert abs(r["lnl"][2]) < 1e-9                                       # all-missing column: likelihood 1


CLADE_SHAPES = [
    "(((a:0.1,b:0.2):0.1,c:0.3):0.2,((d:0.1,(e:0.2,f:0.1):0.3):0.1,g:0.2):0.1);",                    # pitchforks, tip first / last
    "((((a:0.1,b:0.2):0.1,c:0.3):0.2,((d:0.1,e:0.2):0.3,f:0.1):0.1):0.05,((g:0.2,h:0.1):0.2,(i:0.3,j:0.1):0.1):0.3);",  # sibling pitchforks, sibling cherries
    "((a:0.1,b:0.2):0.1,((c:0.3,d:0.1):0.2,e:0.4):0.3);",                                             # cherry / pitchfork AS root children: kept
    "(((((a:0.1,b:0.1):0.1,c:0.1):0.1,d:0.1):0.1,e:0.2):0.1,((f:0.1,g:0.3):0.2,((h:0.1,i:0.1):0.2,(j:0.2,(k:0.1,l:0.3):0.1):0.1):0.2):0.1);",
    "((a:0.2,b:0.01,(g:0.1,h:0.2):0.0,(i:0.1,(j:0.2,k:0.3):0.1):0.2):0.1,((c:0.3,d:0.05):0.1,(e:0.6,f:0.1):0.02,m:0.3):0.02);",   # below polytomies
]


@pytest.mark.parametrize("newick", CLADE_SHAPES)
def test_clade_table_shapes(ctx, newick):
    """Every way a cherry or pitchfork table can enter a step (before / after the contraction, two per step, next to real
    tips, under polytomies, not at the root), with missing data, against the oracle; and against the table-free path."""
    tree = treeio.parse_newick(newick)
    flat = treeio.flatten_tree(tree)
    n_tips = len(tree.leaves())
    rng = np.random.default_rng(len(newick))
    states = rng.integers(0, 22, size=(n_tips, 700)).astype(np.uint8)
    states[states >= 20] = 255
    model = oracle.OracleModel("Dayhoff")
    rates = oracle.gamma_rates(3, 0.8)
    lnl, joint = oracle.run(flat, states, model, rates)
    _, node_slot, _ = native.schedule_dump(flat)
    ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
    ctx.set_rates(rates)
    ctx.set_tree(flat)
    ctx.set_alignment(states)
    r = ctx.run()
    assert lnl_err(r["lnl"], lnl) < TOL and rel_err(r["joint"], joint) < TOL
    if "AS root" not in newick and newick != CLADE_SHAPES[2]:
        assert np.count_nonzero(node_slot <= -2) >= 2
    try:
        ctx.set_option("cherry_tables", 0)
        r0 = ctx.run()
    finally:
        ctx.set_option("cherry_tables", 1)
    assert lnl_err(r0["lnl"], lnl) < TOL and rel_err(r["joint"], r0["joint"]) < 1e-10
    st = ctx.run_streamed(states)
    assert np.array_equal(st["lnl"], r["lnl"]) and np.array_equal(st["joint"], r["joint"])


def test_random_trees_fuzz(ctx):
    """Random topologies (polytomies, zero branches), rate-class counts, models and missing data through the ABI."""
    rng = np.random.default_rng(77)
    for it in range(16):
        n = int(rng.integers(4, 60))
        tree = treeio.parse_newick(random_newick(rng, n))
        flat = treeio.flatten_tree(tree)
        rc = flat.child_idx[flat.child_off[flat.root]:flat.child_off[flat.root + 1]]
        leaf_root = any(flat.child_off[v] == flat.child_off[v + 1] for v in rc)
        K = 1 if leaf_root else int(rng.integers(1, 6))
        states = rng.integers(0, 22, size=(n, 257 + 13 * it)).astype(np.uint8)
        states[states >= 20] = 255
        model = oracle.OracleModel(["WAG", "JTT", "Dayhoff", "BLOSUM62"][it % 4])
        rates = oracle.gamma_rates(K, 0.6) if K > 1 else np.array([1.0])
        lnl, joint = oracle.run(flat, states, model, rates)
        ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
        ctx.set_rates(rates)
        ctx.set_tree(flat)
        ctx.set_alignment(states)
        r = ctx.run()
        assert lnl_err(r["lnl"], lnl) < TOL and rel_err(r["joint"], joint) < TOL, (it, tree.to_newick())


def test_leaf_as_root_child(ctx):
    """K=1 works in the reference (length-1 shortcut, utils/Utils.java:56-58); K>1 makes the reference throw for
    that column — the native path returns the well-defined limit instead (documented)."""
    tree = treeio.parse_newick("(a:0.1,(b:0.2,c:0.3):0.1);")
    flat = treeio.flatten_tree(tree)
    model = oracle.OracleModel("WAG")
    states = treeio.encode_alignment(["A-", "AR", "RR"])
    lnl1, joint1 = oracle.run(flat, states, model, np.array([1.0]))
    ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
    ctx.set_rates(np.array([1.0]))
    ctx.set_tree(flat)
    ctx.set_alignment(states)
    r = ctx.run()
    assert rel_err(r["lnl"], lnl1) < TOL and rel_err(r["joint"], joint1) < TOL
    ctx.set_rates(oracle.gamma_rates(4, 0.5))
    r = ctx.run()
    assert np.all(np.isfinite(r["lnl"])) and np.max(np.abs(r["joint"].sum(axis=(1, 2)) - 1.0)) < 1e-12
    assert np.count_nonzero(r["joint"][0].sum(axis=1) > 0) == 1         # state at A is the observed tip state


def test_degenerate_shapes(ctx):
    """Two-taxon tree (both root children are tips, K=1 as the reference requires), K=16 rate classes, a single column."""
    model = oracle.OracleModel("JTT")
    tree = treeio.parse_newick("(a:0.1,b:0.2);")
    flat = treeio.flatten_tree(tree)
    states = treeio.encode_alignment(["AR-W", "AK-C"])
    lnl, joint = oracle.run(flat, states, model, np.array([1.0]))
    ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
    ctx.set_rates(np.array([1.0]))
    ctx.set_tree(flat)
    ctx.set_alignment(states)
    r = ctx.run()
    assert lnl_err(r["lnl"], lnl) < TOL and rel_err(r["joint"], joint) < TOL
    p = synthetic(2, 64)
    rates = oracle.gamma_rates(16, 0.3)
    lnl, joint = oracle.run(p.flat, p.states, p.model, rates)
    setup(ctx, p)
    ctx.set_rates(rates)
    r = ctx.run()
    assert rel_err(r["lnl"], lnl) < TOL and rel_err(r["joint"], joint) < TOL
    one = ctx.run(site0=63, n=1)
    assert rel_err(one["lnl"], lnl[63:]) < TOL and rel_err(one["joint"], joint[63:]) < TOL
    ctx.set_rates(p.rates)


def test_compact_matches_site_result_scan(ctx, example_problem):
    p = example_problem
    setup(ctx, p)
    full = ctx.run()
    try:
        for thr in (0.5, 0.2, 0.0):
            cap = native.compact_cap_for(thr)
            ctx.set_option("compact_cap", cap)
            c = ctx.run(want_joint=False, want_compact=True, threshold=thr)["compact"]
            for i in range(len(c)):
                J = full["joint"][i]
                sr = recon.SiteResult(i, full["lnl"][i], J, thr, sort_by_prob=False)
                assert c["max_prob"][i] == sr.get_max_prob() and c["max_ii_prob"][i] == sr.get_max_ii_prob()
                assert c["n_kept"][i] == len(sr.probs) <= cap
                assert bool(c["interesting"][i]) == (sr.get_max_ii_prob() <= 1.0 - thr and sr.get_max_prob() >= thr)
                assert c["prob"][i][:len(sr.probs)].tolist() == sr.probs
                assert str(recon.SiteResult.from_compact(i, c[i])) == str(recon.SiteResult(i, full["lnl"][i], J, thr))
    finally:
        ctx.set_option("compact_cap", 8)


def test_compact_capacity_is_lossless_or_refused(ctx, example_problem):
    """A compact record holds `compact_cap` entries (default 8). A threshold that could keep more (floor(1/t) > cap) is
    refused instead of truncated; with the capacity raised to sr_compact_cap_for(t) every kept entry is there."""
    p = example_problem
    setup(ctx, p)
    full = ctx.run()
    assert [native.compact_cap_for(t) for t in (0.5, 0.34, 0.2, 0.125, 0.1, 0.01, 0.0)] == [2, 2, 5, 8, 10, 100, 400]
    with pytest.raises(native.SrError, match="compact_cap"):
        ctx.run(want_joint=False, want_compact=True, threshold=0.1)
    with pytest.raises(native.SrError, match="compact_cap"):
        ctx.run_streamed(p.states, want_joint=False, want_compact=True, threshold=0.0)
    try:
        for thr in (0.5, 0.1, 0.03, 0.0):
            cap = native.compact_cap_for(thr)
            ctx.set_option("compact_cap", cap)
            assert native.compact_dtype(cap).itemsize == native.lib().sr_compact_stride(cap)
            for runner in (lambda: ctx.run(want_joint=False, want_compact=True, threshold=thr),
                           lambda: ctx.run_streamed(p.states, want_joint=False, want_compact=True, threshold=thr)):
                c = runner()["compact"]
                assert c.dtype.itemsize == native.compact_dtype(cap).itemsize and np.array_equal(c["lnl"], full["lnl"])
                for i in range(len(c)):
                    sr = recon.SiteResult(i, full["lnl"][i], full["joint"][i], thr, sort_by_prob=False)
                    assert c["n_kept"][i] == len(sr.probs) <= cap
                    assert c["prob"][i][:len(sr.probs)].tolist() == sr.probs
                    assert [recon.AA_ORDER[a // 20] + recon.AA_ORDER[a % 20] for a in c["pair"][i][:len(sr.probs)]] == sr.subs
                    assert np.all(c["pair"][i][len(sr.probs):] == -1) and np.all(c["prob"][i][len(sr.probs):] == 0.0)
    finally:
        ctx.set_option("compact_cap", 8)
    with pytest.raises(native.SrError):
        ctx.set_option("compact_cap", 401)


@pytest.mark.parametrize("cfg,n", [(1, 130), (2, 10_000), (3, 4096)])
def test_printed_output_via_compact_equals_via_full_joint(ctx, example_problem, cfg, n):
    """The Java host builds its SiteResults from compact records (java/subrecon/recon/NativeRecon.java); what it prints
    must be what the 400-probability route prints — thresholds 0.5 (default), 0.2, 0.125, sorted and -nosort, -verbose."""
    p = example_problem if cfg == 1 else synthetic(cfg, n)
    setup(ctx, p)
    full = ctx.run()
    try:
        for thr in (0.5, 0.2, 0.125):
            ctx.set_option("compact_cap", native.compact_cap_for(thr))
            c = ctx.run(want_joint=False, want_compact=True, threshold=thr)["compact"]
            for sort in (True, False):
                a = [recon.SiteResult(i, full["lnl"][i], full["joint"][i], thr, sort) for i in range(n)]
                b = [recon.SiteResult.from_compact_like_java(i, c[i], thr, sort) for i in range(n)]
                d = [recon.SiteResult.from_compact(i, c[i], sort) for i in range(n)]
                for verbose in (False, True):
                    want = recon.run_columns(a, verbose, thr)
                    assert recon.run_columns(b, verbose, thr) == want and recon.run_columns(d, verbose, thr) == want
                assert [x.get_max_prob() for x in b] == [x.get_max_prob() for x in a]
                assert [x.get_max_ii_prob() for x in b] == [x.get_max_ii_prob() for x in a]
            assert np.array_equal(c["interesting"] != 0, [x.get_max_ii_prob() <= 1.0 - thr and x.get_max_prob() >= thr for x in a])
    finally:
        ctx.set_option("compact_cap", 8)


@pytest.mark.parametrize("cfg,n", [(2, 10_000), (3, 4096), (4, 1024), (5, 2048)])
def test_printed_output_identical_on_synthetic_configs(ctx, cfg, n):
    """north_star: "identical printed output at the default -sd and threshold" — on the oracle subsets of BASELINE configs
    #2-#5, not only on the example. What is printed is roundDouble(x, 2) = floor(100 x + 1/2) / 100 of lnL and of every kept
    probability (utils/Utils.java:28-31, recon/SiteResult.java:124-139), so the GPU and the reference can only print
    different text where a value sits within their (~1e-12 relative) difference of an x.xx5 boundary, of the threshold, or of
    another kept value (sort order). Counted over EVERY value of the subset (vectorised), then whole lines are compared:
    default run (-threshold 0.5), -threshold 0.2, -verbose, -nosort, and -threshold 0.0 -verbose on the leading columns.
    The counts and the closest approach to a boundary go to gpurun_out/r02_printed_identity_cfg<N>.json."""
    import json
    p = synthetic(cfg, n)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    setup(ctx, p)
    r = ctx.run()
    g_lnl, g_joint = r["lnl"], r["joint"]
    rep = {"config": cfg, "columns": n, "sd": 2}

    def rounded(x):
        return np.floor(x * 100.0 + 0.5)

    def margin(x):      # distance of 100 x from the nearest rounding boundary (k + 1/2), in units of the printed last digit
        y = x * 100.0
        return float(np.min(np.abs(y - np.floor(y) - 0.5)))

    rep["lnl_rounding_flips"] = int(np.count_nonzero(rounded(lnl) != rounded(g_lnl)))
    rep["prob_rounding_flips"] = int(np.count_nonzero(rounded(joint) != rounded(g_joint)))
    rep["lnl_closest_to_boundary"] = margin(lnl)
    printable = joint >= 0.005                           # below that both print 0.0 whatever the bits
    rep["prob_closest_to_boundary"] = margin(joint[printable])
    rep["max_abs_diff_x100"] = float(max(np.max(np.abs(lnl - g_lnl)), np.max(np.abs(joint - g_joint))) * 100.0)
    for thr in (0.5, 0.2, 0.125):
        rep[f"threshold_{thr}_membership_flips"] = int(np.count_nonzero((joint >= thr) != (g_joint >= thr)))
    ii_o = np.diagonal(joint, axis1=1, axis2=2).max(axis=1)
    ii_g = np.diagonal(g_joint, axis1=1, axis2=2).max(axis=1)
    rep["print_rule_flips"] = int(np.count_nonzero(((ii_o <= 0.5) & (joint.max(axis=(1, 2)) >= 0.5)) != ((ii_g <= 0.5) & (g_joint.max(axis=(1, 2)) >= 0.5))))

    def lines(l, j, thr, verbose, sort, cols):
        return recon.run_columns([recon.SiteResult(i, l[i], j[i], thr, sort) for i in range(cols)], verbose, thr)

    compared = 0
    for thr, verbose, sort, cols in ((0.5, False, True, n), (0.5, True, True, n), (0.2, False, True, n), (0.2, True, False, n),
                                     (0.0, True, False, min(n, 256)), (0.0, True, True, min(n, 48))):
        want, got = lines(lnl, joint, thr, verbose, sort, cols), lines(g_lnl, g_joint, thr, verbose, sort, cols)
        body_w = [x for x in want if x.startswith("Result")]
        body_g = [x for x in got if x.startswith("Result")]
        if thr == 0.0 and sort:
            # 400 entries per line, most of them printed as 0.0 and many below 1e-300, where the reference's log-space
            # arithmetic and the native linear-space arithmetic flush to zero at different points: the ORDER of entries
            # that print identically may differ; every token and every printed value must still agree
            for a, b in zip(body_w, body_g):
                if a != b:
                    ta, tb = a.split("\t"), b.split("\t")
                    assert ta[:3] == tb[:3] and sorted(ta[3:]) == sorted(tb[3:])
                    moved = [x for x, y in zip(ta[3:], tb[3:]) if x != y]
                    assert all(x.endswith(":0.0") for x in moved), moved[:4]
                    rep["order_only_differences_at_threshold_0_sorted"] = rep.get("order_only_differences_at_threshold_0_sorted", 0) + 1
        else:
            assert body_g == body_w, (thr, verbose, sort)
        assert len(body_g) == len(body_w)
        tot_w, tot_g = float(want[len(body_w)].split()[-1]), float(got[len(body_g)].split()[-1])
        assert abs(tot_g - tot_w) < 1e-9 * max(1.0, abs(tot_w))            # Total lnL: compared numerically
        assert got[len(body_g) + 1:] == want[len(body_w) + 1:]             # the "0 sites have ..." notice, if any
        compared += len(body_w)
    rep["result_lines_compared"] = compared
    try:
        out_dir = os.path.join(os.path.dirname(HERE), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        json.dump(rep, open(os.path.join(out_dir, f"r02_printed_identity_cfg{cfg}.json"), "w"), indent=1)
    except OSError:
        pass
    assert rep["lnl_rounding_flips"] == 0 and rep["prob_rounding_flips"] == 0 and rep["print_rule_flips"] == 0, rep
    assert all(rep[f"threshold_{t}_membership_flips"] == 0 for t in (0.5, 0.2, 0.125)), rep


def test_multi_device_entry(ctx, example_problem):
    """sr_multi_*: one host process, one context + one host thread per device, contiguous column blocks, results in place.
    On a one-GPU box the "devices" are three contexts on cuda:0; with more GPUs every visible device takes part."""
    import torch
    p = example_problem
    setup(ctx, p)
    ref = ctx.run()
    big = np.tile(p.states, (1, 37))[:, :4801]                                # odd size: ragged last block
    devices = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0, 0]
    m = native.MultiContext(devices)
    try:
        assert m.count == len(devices)
        m.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi)
        m.set_rates(p.rates)
        m.set_tree(p.flat)
        pin = native.PinnedArray(big.shape, np.uint8)
        pin.array[:] = big
        for src in (big, pin.array):
            r = m.run_streamed(src, want_joint=True, want_compact=True)
            assert np.array_equal(r["lnl"], np.tile(ref["lnl"], 37)[:4801])
            assert np.array_equal(r["joint"], np.tile(ref["joint"], (37, 1, 1))[:4801])
            assert np.array_equal(r["compact"]["lnl"], r["lnl"])
        part = m.run_streamed(pin.array, site0=77, n=1000, want_joint=False, want_compact=True)
        assert np.array_equal(part["lnl"], r["lnl"][77:1077]) and part["compact"].tobytes() == r["compact"][77:1077].tobytes()
        one = m.run_streamed(pin.array, site0=4800, n=1)                       # -site n: one column, most devices idle
        assert np.array_equal(one["lnl"], r["lnl"][4800:])
        m.set_option("compact_cap", 2)
        r2 = m.run_streamed(pin.array, want_joint=False, want_compact=True, threshold=0.5)
        assert r2["compact"].dtype.itemsize == 56 and np.array_equal(r2["compact"]["prob"], r["compact"]["prob"][:, :2])
        with pytest.raises(native.SrError, match="device .*compact_cap"):
            m.run_streamed(pin.array, want_joint=False, want_compact=True, threshold=0.2)
        with pytest.raises(native.SrError, match="unknown option"):
            m.set_option("no_such_option", 1)
        pin.free()
    finally:
        m.close()
    with pytest.raises(native.SrError):
        native.MultiContext([0, 99])


def test_strict_root_option(ctx):
    """a tip as root child with K > 1 makes the reference throw on every column (utils/Utils.java:56-70).
    With "strict_root" the input is rejected once, from whichever of sr_set_tree / sr_set_rates completes the combination."""
    flat = treeio.flatten_tree(treeio.parse_newick("(a:0.1,(b:0.2,c:0.3):0.1);"))
    ok = treeio.flatten_tree(treeio.parse_newick("((a:0.1,d:0.1):0.1,(b:0.2,c:0.3):0.1);"))
    model = oracle.OracleModel("WAG")
    c = native.Context(0)
    try:
        c.set_option("strict_root", 1)
        c.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
        c.set_rates(oracle.gamma_rates(4, 0.5))
        with pytest.raises(native.SrError, match="child of the root is a tip") as ei:
            c.set_tree(flat)
        assert ei.value.code == -5
        c.set_tree(ok)
        c.set_rates(np.array([1.0]))
        c.set_tree(flat)                                                   # K = 1: the reference handles it
        with pytest.raises(native.SrError, match="child of the root is a tip"):
            c.set_rates(oracle.gamma_rates(2, 0.5))
        c.set_alignment(treeio.encode_alignment(["A-", "AR", "RR"]))
        assert np.isfinite(c.run()["lnl"]).all()                           # still K = 1
    finally:
        c.close()


def test_host_mirror_interface(ctx, example_problem):
    p = example_problem
    j = recon.JointBranchReconstruction(p.states, p.flat, p.model, p.model.pi, p.rates, sanity_check=True, ctx=ctx)
    res = j.call()
    assert len(res) == 130 and res[13].subs[0] == "RK" and str(res[13]).startswith("Result\t14\t-7.0\tRK:1.0")


def test_single_process_multi_context_sharding(ctx, example_problem):
    """One process, several contexts (one per device in production; three on cuda:0 here), each on its own thread."""
    from subrecon_b200 import sharding
    p = example_problem
    setup(ctx, p)
    ref = ctx.run()
    ctxs = [native.Context(0) for _ in range(3)]
    try:
        for c in ctxs:
            c.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi)
            c.set_rates(p.rates)
            c.set_tree(p.flat)
        big = np.tile(p.states, (1, 40))
        r = sharding.run_multi_context(ctxs, big, want_joint=True, want_compact=True)
        assert np.array_equal(r["lnl"], np.tile(ref["lnl"], 40)) and np.array_equal(r["joint"][:130], ref["joint"])
        assert np.array_equal(r["compact"]["lnl"], r["lnl"])
    finally:
        for c in ctxs:
            c.close()


def test_error_behaviour(ctx, example_problem):
    p = example_problem
    fresh = native.Context(0)
    with pytest.raises(native.SrError) as ei:
        fresh.run(0, 1)
    assert ei.value.code == -4                                           # SR_ERR_STATE
    with pytest.raises(native.SrError, match="rate categories"):
        fresh.set_rates([])
    bad = treeio.flatten_tree(treeio.parse_newick("(a:1,b:1,c:1);"))
    with pytest.raises(native.SrError, match="more than two descendents"):
        fresh.set_tree(bad)
    setup(fresh, p)
    with pytest.raises(native.SrError) as ei:
        fresh.run(100, 31)
    assert ei.value.code == -1
    with pytest.raises(native.SrError, match="taxa"):
        fresh.set_alignment(p.states[:5])
        fresh.run()
    with pytest.raises(native.SrError, match="unknown option"):
        fresh.set_option("tile_m", 4)                                     # round-1 tuning variants are gone
    with pytest.raises(native.SrError):
        fresh.set_option("rescale_every", 13)                             # beyond the tested underflow bound
    fresh.close()
    with pytest.raises(native.SrError):
        native.Context(99)


def test_profile_and_roofline_hooks(ctx, example_problem):
    p = example_problem
    setup(ctx, p)
    ctx.set_option("profile", 1)
    ctx.profile(reset=True)
    ctx.run()
    prof = ctx.profile(reset=True)
    ctx.set_option("profile", 0)
    assert prof["n_prune"] == 1 and prof["n_root"] == 1 and prof["ms_prune"] > 0 and prof["columns_pruned"] == 130
    info = ctx.schedule_info()
    assert info["n_steps"] == len(native.schedule_dump(p.flat)[0]) < 34              # fused: fewer steps than the 34 edges
    assert info["flops_per_column_rate"] == 55520 / 4                                 # 55 520 flops/column
    assert 25.0 < ctx.measure_fp64_peak(50) < 45.0
