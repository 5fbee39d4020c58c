"""Max relative error of the CUDA path against the oracle on the oracle-sized subsets of BASELINE configs 1-5 (the numbers
behind BASELINE.md §5's last two columns; the tests assert < 1e-9, this prints the values).
Usage: python tools/parity_report.py [--out gpurun_out/r02_parity_report.json]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_example, rel_err, lnl_err, synthetic  # noqa: E402
from oracle import oracle  # noqa: E402
from subrecon_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
a = ap.parse_args()
ctx = native.Context(0)
rep = {}
for cfg, n in ((1, 130), (2, 10_000), (3, 4096), (4, 1024), (5, 2048)):
    p = load_example() if cfg == 1 else synthetic(cfg, n)
    lnl, joint = oracle.run(p.flat, p.states, p.model, p.rates, threads=os.cpu_count())
    ctx.set_model(p.model.Eval, p.model.Evec, p.model.Ievc, p.model.pi)
    ctx.set_rates(p.rates)
    ctx.set_tree(p.flat)
    ctx.set_alignment(p.states)
    r = ctx.run()
    c = ctx.run(want_joint=False, want_compact=True, threshold=0.5)
    rep[f"config_{cfg}"] = {"columns": n, "taxa": int(p.states.shape[0]), "rate_classes": len(p.rates),
                            "max_rel_err_lnl": rel_err(r["lnl"], lnl), "max_rel_err_joint": rel_err(r["joint"], joint),
                            "max_abs_err_lnl": float(np.max(np.abs(r["lnl"] - lnl))),
                            "max_abs_dev_of_posterior_sum_from_1": float(np.max(np.abs(r["joint"].sum(axis=(1, 2)) - 1.0))),
                            "min_lnl": float(lnl.min()), "total_lnl_oracle": float(np.sum(lnl)), "total_lnl_gpu": float(np.sum(r["lnl"])),
                            "compact_lnl_bit_identical_to_full": bool(np.array_equal(c["lnl"], r["lnl"]))}
    print(cfg, rep[f"config_{cfg}"])
ctx.close()
if a.out:
    json.dump(rep, open(a.out, "w"), indent=1)
