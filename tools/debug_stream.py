import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import synthetic
from subrecon_b200 import native
p = synthetic(5, 2048)
ctx = native.Context(0)
ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi); ctx.set_rates(p.rates); ctx.set_tree(p.flat)
for L in (1_000_000, 4_000_000):
    big = np.empty((p.states.shape[0], L), dtype=np.uint8)
    for c0 in range(0, L, 2048):
        w = min(2048, L - c0); big[:, c0:c0 + w] = p.states[:, :w]
    for mode in ("streamed", "streamed-pinned", "resident"):
        if mode == "resident":
            ctx.set_alignment(big); r = ctx.run(want_joint=False)
        elif mode == "streamed-pinned":
            h = native.PinnedArray(big.shape, np.uint8); h.array[:] = big
            r = ctx.run_streamed(h.array, want_joint=False, want_compact=True); h.free()
        else:
            r = ctx.run_streamed(big, want_joint=False, want_compact=True)
        l = r["lnl"]; ref = np.tile(l[:2048], L // 2048 + 1)[:L]
        bad = np.nonzero(l != ref)[0]
        print(L, mode, "mismatches", bad.size, "first", bad[:6], "chunk idx", (bad[:6] // 130240), "offset in chunk", (bad[:6] % 130240),
              "maxrel", (np.abs(l[bad]-ref[bad])/np.abs(ref[bad])).max() if bad.size else 0, flush=True)
