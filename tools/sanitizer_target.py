"""Small run for compute-sanitizer: config #1 + a 100-taxon synthetic config through every entry point (resident, streamed,
compact, pattern compression). Usage: compute-sanitizer --tool memcheck python tools/sanitizer_target.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_example, rel_err, synthetic  # noqa: E402
from oracle import oracle  # noqa: E402
from subrecon_b200 import native  # noqa: E402

ctx = native.Context(0)
for p in (load_example(), synthetic(2, 600)):
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=4)
    ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi)
    ctx.set_rates(p.rates)
    ctx.set_tree(p.flat)
    ctx.set_alignment(p.states)
    r = ctx.run(want_joint=True, want_compact=True)
    s = ctx.run_streamed(p.states)
    ctx.set_option("compress_patterns", 1)
    c = ctx.run()
    ctx.set_option("compress_patterns", 0)
    w = ctx.run(site0=3, n=77)
    print(p.name, "lnl err %.2e joint err %.2e" % (rel_err(r["lnl"], lnl), rel_err(r["joint"], joint)),
          np.array_equal(s["joint"], r["joint"]), np.array_equal(c["joint"], r["joint"]), np.array_equal(w["lnl"], r["lnl"][3:80]))
ctx.close()
print("done")
