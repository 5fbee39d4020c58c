"""Where does the streamed (host-to-host) path lose time against the resident one? Sums the per-kernel event times of one
`sr_run_streamed` call and compares them with the call's wall time and with a resident run of the same columns."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from subrecon_b200 import native, palmodel, synth, treeio
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
opts = dict(a.split("=") for a in sys.argv[2:])
tree = synth.yule_tree(1000, seed=3)
model = palmodel.host_model("wag"); rates = palmodel.gamma_rates(4, 0.5)
base = synth.simulate_alignment(tree, model, rates, 8192, seed=103)
hs = native.PinnedArray((1000, L), np.uint8)
for c0 in range(0, L, 8192):
    w = min(8192, L - c0); hs.array[:, c0:c0 + w] = base[:, :w]
hl = native.PinnedArray(L, np.float64); hj = native.PinnedArray((L, 20, 20), np.float64)
ctx = native.Context(0)
for k, v in opts.items(): ctx.set_option(k, int(v))
ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi); ctx.set_rates(rates); ctx.set_tree(treeio.flatten_tree(tree))
out = {"lnl": hl.array, "joint": hj.array}
for mode in (os.environ.get("E2E_MODES", "streamed full joint,streamed lnl only,resident").split(",")):
    for rep in range(3):
        ctx.set_option("profile", 1); ctx.profile(reset=True)
        t0 = time.perf_counter()
        if mode == "resident":
            if rep == 0: ctx.set_alignment(hs.array)
            t0 = time.perf_counter()
            ctx.run(want_joint=False)
        elif mode == "streamed lnl only":
            ctx.run_streamed(hs.array, want_joint=False, out={"lnl": hl.array})
        else:
            ctx.run_streamed(hs.array, out=out)
        dt = (time.perf_counter() - t0) * 1e3
        p = ctx.profile(reset=True)
    print(f"{mode:22s} wall {dt:7.2f} ms | prune sum {p['ms_prune']:7.2f} ms over {p['n_prune']} launches, root sum {p['ms_root']:6.2f}, K1 {p['ms_pbuild']:5.2f}")
