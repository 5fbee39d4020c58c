"""Stress the big-tree configs for run-to-run / position-dependent differences (pageable vs pinned upload)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import synthetic
from subrecon_b200 import native
ctx = native.Context(0)
for a in sys.argv[1:]:
    k, v = a.split("="); ctx.set_option(k, int(v))
print("options", sys.argv[1:], flush=True)
for cfg, nbase, L in ((4, 1024, 200_000), (5, 2048, 2_000_000)):
    p = synthetic(cfg, nbase)
    ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi); ctx.set_rates(p.rates); ctx.set_tree(p.flat)
    big = np.tile(p.states, (1, L // nbase + 1))[:, :L].copy()
    pin = native.PinnedArray(big.shape, np.uint8); pin.array[:] = big
    for mode in ("resident-pinned", "streamed-pageable"):
        for trial in range(8):
            src = pin.array if "pinned" in mode else big
            if mode.startswith("resident"):
                ctx.set_alignment(src); r = ctx.run(want_joint=False, want_compact=True)
            else:
                r = ctx.run_streamed(src, want_joint=False, want_compact=True)
            l = r["lnl"]; ref = np.tile(l[:nbase], L // nbase + 1)[:L]
            bad = np.nonzero(l != ref)[0]
            msg = ""
            if bad.size:
                msg = f" first {bad[:5]} mod176 {bad[:5] % 176} rel {np.abs((l[bad]-ref[bad])/ref[bad]).max():.2e} base-block-bad {np.count_nonzero(bad < nbase)}"
            print(f"cfg{cfg} {mode} trial {trial}: mismatches {bad.size}{msg}", flush=True)
    pin.free()
