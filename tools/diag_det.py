"""Where do results depend on position / run? (development diagnostic for the rewritten prune_kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import load_example, synthetic
from subrecon_b200 import native

def report(tag, l, ref, S=192):
    bad = np.nonzero(l != ref)[0]
    msg = f"{tag}: {bad.size} mismatches of {l.size}"
    if bad.size:
        d = np.abs(l[bad] - ref[bad]) / np.abs(ref[bad])
        msg += f" rel min/med/max {d.min():.2e} {np.median(d):.2e} {d.max():.2e}; first {bad[:8]} col%192 {bad[:8] % S} warp {(bad[:8] % S)//16}; tiles {np.unique(bad // S).size}; warps hist {np.bincount((bad % S)//16, minlength=12)}"
    print(msg, flush=True)

ctx = native.Context(0)
for a in sys.argv[1:]:
    k, v = a.split("="); ctx.set_option(k, int(v))
p = load_example()
ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi); ctx.set_rates(p.rates); ctx.set_tree(p.flat)
n0 = p.states.shape[1]
wide = np.tile(p.states, (1, 1901))[:, :247001]
ctx.set_alignment(wide)
r1 = ctx.run(want_joint=False)["lnl"]
r2 = ctx.run(want_joint=False)["lnl"]
ref = np.tile(r1[:n0], 1901)[:247001]
report("example resident run1 vs tiled", r1, ref)
report("example resident run2 vs run1", r2, r1)
ctx.set_option("stream_chunk_cols", 60000)
t = ctx.run_streamed(wide, want_joint=False)["lnl"]
ctx.set_option("stream_taper", 0)
u = ctx.run_streamed(wide, want_joint=False)["lnl"]
ctx.set_option("stream_taper", 1); ctx.set_option("stream_chunk_cols", 1 << 17)
report("example streamed tapered vs tiled", t, ref)
report("example streamed uniform vs tiled", u, ref)
for cfg, nb, L in ((3, 4096, 1_000_000), (4, 1024, 200_000)):
    p = synthetic(cfg, nb)
    ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi); ctx.set_rates(p.rates); ctx.set_tree(p.flat)
    big = np.tile(p.states, (1, L // nb + 1))[:, :L]
    ctx.set_alignment(big)
    for trial in range(3):
        l = ctx.run(want_joint=False)["lnl"]
        report(f"cfg{cfg} resident trial {trial} vs tiled", l, np.tile(l[:nb], L // nb + 1)[:L])
