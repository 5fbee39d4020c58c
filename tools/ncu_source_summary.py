"""Per-opcode instruction counts and stall samples per consumer warp-step from an `ncu --page source --csv` export.
usage: python tools/ncu_source_summary.py src.csv <warp_steps> [--top N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ws = float(sys.argv[2])
ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); isam = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = collections.Counter(); samples = collections.Counter(); tot = 0; tots = 0
stalls = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) <= ia or not r[ia]: continue
    n = int(r[ia]); sm = int(r[isam] or 0)
    op = r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1]
    op = op.split(".")[0]
    ops[op] += n; samples[op] += sm; tot += n; tots += sm
    for i in stall_cols:
        if r[i]: stalls[hdr[i]] += int(r[i])
    lines.append((sm, n, r[0], r[isrc]))
print(f"instructions per warp-step: {tot / ws:.1f}   (samples {tots})")
for op, n in ops.most_common(32):
    print(f"  {op:10s} {n / ws:8.2f} instr/step   {100.0 * samples[op] / max(1, tots):5.1f} % of samples")
print("stall samples:", {k: round(100.0 * v / max(1, tots), 1) for k, v in stalls.most_common(10)})
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
print("hottest instructions by samples:")
for sm, n, addr, src in sorted(lines, reverse=True)[:top]:
    print(f"  {100.0 * sm / tots:5.2f} %  exec/step {n / ws:6.2f}  {addr[-5:]}  {src[:90]}")
