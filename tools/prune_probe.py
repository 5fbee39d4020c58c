"""Small fixed workload for ncu / timing experiments: config #3 tree, `--cols` columns resident, a few launches.
Usage: python tools/prune_probe.py [--cols 131072] [--reps 3] [--taxa 1000] [--tree yule|balanced|caterpillar] [--tables 0|1]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from subrecon_b200 import native, palmodel, synth, treeio  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cols", type=int, default=131072)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--rescale", type=int, default=0)
ap.add_argument("--k-fastest", type=int, default=-1)
ap.add_argument("--tables", type=int, default=1)
ap.add_argument("--taxa", type=int, default=1000)
ap.add_argument("--tree", default="yule")
ap.add_argument("--ncat", type=int, default=4)
ap.add_argument("--uniform", action="store_true", help="uniform random states instead of simulated columns")
ap.add_argument("--distinct", action="store_true", help="simulate every column down the tree (on the GPU, like bench.py) instead of tiling 4096")
ap.add_argument("--model", default="wag")
ap.add_argument("--const", action="store_true", help="every tip in state 0 (no shared-memory bank conflicts in the tip gathers)")
a = ap.parse_args()

tree = {"yule": synth.yule_tree, "balanced": synth.balanced_tree, "caterpillar": synth.caterpillar_tree}[a.tree](a.taxa, seed=3)
model = palmodel.host_model(a.model)
rates = palmodel.gamma_rates(a.ncat, 0.5)
if a.const:
    base = np.zeros((a.taxa, 4096), dtype=np.uint8)
elif a.uniform:
    base = np.random.default_rng(1).integers(0, 20, size=(a.taxa, 4096), dtype=np.uint8)
elif a.distinct:
    base = None
else:
    base = synth.simulate_alignment(tree, model, rates, 4096, seed=103)
if base is None:
    states = synth.simulate_alignment_torch(tree, model, rates, a.cols, seed=103, device="cuda")      # what bench.py feeds the path by default
else:
    states = np.tile(base, (1, (a.cols + 4095) // 4096))[:, :a.cols]
ctx = native.Context(0)
if a.rescale:
    ctx.set_option("rescale_every", a.rescale)
if a.k_fastest >= 0:
    ctx.set_option("k_fastest", a.k_fastest)
ctx.set_option("cherry_tables", a.tables)
ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
ctx.set_rates(rates)
ctx.set_tree(treeio.flatten_tree(tree))
ctx.set_alignment(states)
info = ctx.schedule_info()
ctx.run(want_joint=False)
ctx.set_option("profile", 1)
ctx.profile(reset=True)
for _ in range(a.reps):
    ctx.run(want_joint=False)
p = ctx.profile(reset=True)
ms = p["ms_prune"] / p["n_prune"]
tf = info["flops_per_column_rate"] * len(rates) * a.cols / ms * 1e-9
print(f"tree={a.tree} taxa={a.taxa} cols={a.cols} distinct={int(a.distinct)} tables={a.tables} k_fastest={a.k_fastest} K={a.ncat} depth={info['stack_depth']} steps={info['n_steps']}: "
      f"prune {ms:.3f} ms/launch  {a.cols / ms * 1e-3:.3f} Mcol/s  {tf:.2f} TFLOP/s  root {p['ms_root'] / p['n_root']:.3f} ms")
