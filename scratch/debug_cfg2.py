import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from helpers import synthetic, rel_err, load_example
from oracle import oracle
from subrecon_b200 import native
ctx = native.Context(0)
ex = load_example()
for cfg, n in ((2, 10000), (3, 2048)):
    p = synthetic(cfg, n)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    for tm, w in ((2, 8), (2, 11), (2, 15), (4, 8)):
        for trial in range(3):
            # interleave with a small problem, as the test-suite does
            if trial == 1:
                ctx.set_option("tile_m", 2); ctx.set_model(ex.model.Eval, ex.model.Evec, ex.model.Ievc, ex.model.pi); ctx.set_rates(ex.rates); ctx.set_tree(ex.flat); ctx.set_alignment(ex.states); ctx.run()
            ctx.set_option("tile_m", tm); ctx.set_option("warps", w)
            ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi); ctx.set_rates(p.rates); ctx.set_tree(p.flat); ctx.set_alignment(p.states)
            r = ctx.run()
            bad = np.nonzero(np.abs(r["lnl"] - lnl) / np.abs(lnl) > 1e-9)[0]
            print(f"cfg{cfg} tile_m={tm} warps={w} trial={trial}: lnl err {rel_err(r['lnl'], lnl):.2e} joint err {rel_err(r['joint'], joint):.2e} bad cols {bad.size} first {bad[:8]}", flush=True)
