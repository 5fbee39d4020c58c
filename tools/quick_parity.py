import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from helpers import synthetic, rel_err, load_example
from oracle import oracle
from subrecon_b200 import native
ctx = native.Context(0)
opts = dict(a.split('=') for a in sys.argv[1:])
for cfg, n in ((1, 0), (2, 3000), (3, 1024), (4, 256)):
    p = load_example() if cfg == 1 else synthetic(cfg, n)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    for k, v in opts.items(): ctx.set_option(k, int(v))
    ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi); ctx.set_rates(p.rates); ctx.set_tree(p.flat); ctx.set_alignment(p.states)
    r = ctx.run(); r2 = ctx.run()
    print(f"cfg{cfg} {opts}: lnl err {rel_err(r['lnl'], lnl):.2e} joint err {rel_err(r['joint'], joint):.2e} deterministic {np.array_equal(r['joint'], r2['joint'])}", flush=True)
