import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from helpers import synthetic
from subrecon_b200 import native
p = synthetic(3, 4096)
ctx = native.Context(0)
for tm in (2,):
    ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi); ctx.set_rates(p.rates); ctx.set_tree(p.flat)
    for reps in (4, 64, 256):
        big = np.tile(p.states, (1, reps))
        ctx.set_alignment(big)
        for trial in range(2):
            r = ctx.run(want_joint=False)
            l = r["lnl"]
            ref = np.tile(l[:4096], reps)
            bad = np.nonzero(l != ref)[0]
            print(f"tile_m={tm} reps={reps} trial={trial}: {bad.size} mismatches of {l.size}")
            if bad.size:
                d = np.abs(l[bad] - ref[bad]) / np.abs(ref[bad])
                print("   rel diff min/median/max", d.min(), np.median(d), d.max())
                print("   first bad cols", bad[:20], " mod 128:", (bad[:20] % 128), " tile:", bad[:20] // 128)
                S = 192
                print("   distinct tiles with errors:", np.unique(bad // S).size, "of", l.size // S, "; hist of col-in-tile//8:", np.bincount((bad % S) // 8, minlength=S // 8))
                # is the base block itself consistent between trials?
        r2 = ctx.run(want_joint=False)
        print("   run-to-run identical:", np.array_equal(r["lnl"], r2["lnl"]), " base block equal across runs:", np.array_equal(r["lnl"][:4096], r2["lnl"][:4096]))
