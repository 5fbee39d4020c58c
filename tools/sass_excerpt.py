"""SASS evidence of the shipped prune_kernel (no GPU needed): opcode histogram plus the sites that show the Blackwell-era
mechanisms — DMMA.8x8x4 bursts between BAR.SYNC/BAR.ARV hand-offs, UBLKCP (TMA bulk copies), SYNCS (mbarriers),
USETMAXREG (register re-allocation). Usage: python tools/sass_excerpt.py > profiles/r02_prune_kernel_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "subrecon_b200", "libsubrecon_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, ln in enumerate(sass) if "Function :" in ln and "prune_kernel" in ln)
end = next((i for i in range(start + 1, len(sass)) if "Function :" in sass[i]), len(sass))
body = sass[start:end]
pat = re.compile(r"^\s+/\*([0-9a-f]{4})\*/\s+((?:@!?U?P\d+\s+)?)([A-Za-z0-9_.]+)(.*?);")
ins = []          # (addr, pred, op, rest, line)
for ln in body:
    m = pat.match(ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip(), m.group(3), m.group(4), ln.split("/*")[1].split("*/")[0] + "  " + ln.split("*/", 1)[1].split("/*")[0].rstrip()))
ops = collections.Counter(i[2] for i in ins)
print("# prune_kernel, sm_100a cubin inside subrecon_b200/libsubrecon_b200.so (cuobjdump -sass); built by __graft_entry__.build()")
for i, ln in enumerate(res):
    if "prune_kernel" in ln:
        print("#", res[i + 1].strip())
print(f"# {len(ins)} instructions = {16 * len(ins)} bytes of SASS (instruction cache: 128 KB)")
print("# opcode histogram (static):")
for op, n in ops.most_common():
    print(f"#   {op:34s} {n}")
print()


def show(title, idxs, before=2, after=2):
    print("=" * 100)
    print(title)
    last = -10
    for i in idxs:
        lo, hi = max(0, i - before), min(len(ins), i + after + 1)
        if lo > last + 1:
            print("    ...")
        for j in range(max(lo, last + 1), hi):
            print("   ", ins[j][4])
        last = hi - 1
    print()


find = lambda pred: [i for i, x in enumerate(ins) if pred(x)]
show("register re-allocation (setmaxnreg.dec 32 in the producer's warpgroup, setmaxnreg.inc 160 in the consumers)", find(lambda x: x[2].startswith("USETMAXREG")))
show("TMA bulk copies issued by the producer lane (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes)", find(lambda x: x[2].startswith("UBLKCP")), 3, 1)
show("mbarrier sites (SYNCS): expect_tx + arrive of the producer, try_wait / test_wait of the consumers, the stage release", find(lambda x: x[2].startswith("SYNCS")), 1, 1)
# one complete burst: from the first BAR.SYNC that precedes a DMMA to the BAR.ARV after the 30th DMMA
syncs = find(lambda x: x[2].startswith("BAR.SYNC") and "0x40" in x[3])
arvs = find(lambda x: x[2].startswith("BAR.ARV"))
best = None
for s0 in syncs:
    a1 = min((i for i in arvs if i > s0), default=None)
    if a1 is not None:
        n = sum(1 for x in ins[s0:a1] if x[2].startswith("DMMA"))
        if best is None and n >= 27:
            best = (n, s0, a1)
b0, b1 = max(0, best[1] - 8), min(len(ins) - 1, best[2] + 8)    # (ptxas lets the last DMMAs trail the hand-off by a few slots)
print("=" * 100)
print(f"one 30-DMMA burst of a specialised step body, from the rotation barrier to the hand-off ({b1 - b0 + 1} instructions, "
      f"{sum(1 for x in ins[b0:b1 + 1] if x[2].startswith('DMMA'))} DMMA.8x8x4; 8 instructions of context either side)")
for j in range(b0, b1 + 1):
    print("   ", ins[j][4])
