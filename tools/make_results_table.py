"""BASELINE.md §5 from the bench lines and the parity report under profiles/ (prints the markdown table).
usage: python tools/make_results_table.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def line(name):
    return json.loads(open(os.path.join(P, name)).read().strip().splitlines()[-1])


par = json.load(open(os.path.join(P, "r02_parity_report.json")))
rows = [("#1 lysozyme 19×130", 1, "r02_bench_cfg1_n1.json", 1), ("#2 100×10k JTT", 1, "r02_bench_cfg2_n1.json", 2),
        ("#3 1k×1M WAG", 1, "r02_bench_cfg3_n1.json", 3), ("#3 1k×1M WAG, 1M columns per GPU (weak)", 8, "r02_bench_cfg3_n8_weak.json", 3),
        ("#3 1k×1M WAG, 1M columns in total (strong)", 8, "r02_bench_cfg3_n8_strong.json", 3),
        ("#4 5k caterpillar×200k BLOSUM62 K=8", 1, "r02_bench_cfg4_n1.json", 4),
        ("#5 2k×8M Dayhoff", 1, "r02_bench_cfg5_n1.json", 5), ("#5 2k×8M Dayhoff (strong)", 2, "r02_bench_cfg5_n2_strong.json", 5),
        ("#5 2k×8M Dayhoff (strong)", 4, "r02_bench_cfg5_n4_strong.json", 5), ("#5 2k×8M Dayhoff (strong)", 8, "r02_bench_cfg5_n8_strong.json", 5)]
cpu = {}
print("| config | GPUs | columns/s kernel-only (`value`) | columns/s end-to-end, compact result (`e2e`) | e2e, all 400 probabilities | FP64 TFLOP/s achieved (algorithmic) | % of measured DMMA roof (algorithmic / executed) | CPU (A) columns/s @T | CPU (C) columns/s @T | e2e speed-up vs (A) | max rel err lnL | max rel err joint |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for label, n, f, cfg in rows:
    try:
        d = line(f)
    except OSError:
        continue
    cb = d.get("cpu_baseline")
    if cb:
        cpu[cfg] = (cb["variants"]["A_jar_faithful"]["columns_per_s"], cb["variants"]["C_hoisted_honest_cpu"]["columns_per_s"], cb["cores"])
    a, c, t = cpu.get(cfg, (None, None, None))
    r = d["roofline"]
    pr = par[f"config_{cfg}"]
    e2e = d["e2e"]["value"] if d.get("e2e") else None
    full = d["e2e_full_joint"]["value"] if d.get("e2e_full_joint") else None
    fm = lambda x: "—" if x is None else (f"{x / 1e6:.2f} M" if x >= 1e5 else f"{x:,.0f}")
    print(f"| {label} | {n} | {fm(d['value'])} | {fm(e2e)} | {fm(full)} | {r['achieved']:.1f}{' per GPU' if n > 1 else ''} | "
          f"{100 * r['frac']:.0f} % / {100 * r['frac_executed']:.0f} % | {'—' if a is None else f'{a:.1f} @{t}'} | {'—' if c is None else f'{c:,.0f} @{t}'} | "
          f"{'—' if a is None or e2e is None else f'{e2e / a:,.0f}×'} | {pr['max_rel_err_lnl']:.1e} | {pr['max_rel_err_joint']:.1e} |")
