"""Per-warp clock stamps of the pruning kernel's step phases on one SM (instrumented build, -DSR_TIMELINE).

Build here (no GPU needed):  python tools/timeline.py --build
Run on the GPU box:          SUBRECON_B200_LIB=tools/_build/libsubrecon_tl.so python tools/timeline.py [--tables 0]
Stamps per step and warp: 0 step begins, 1 stage landed, 2 ready to burst, 3 burst starts (after the rotation barrier),
4 last DMMA issued, 5 stage released, 6 epilogue issued, 7 step ends.
"""
import argparse
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TL_LIB = os.path.join(ROOT, "tools", "_build", "libsubrecon_tl.so")
ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--tables", type=int, default=1)
ap.add_argument("--cols", type=int, default=148 * 192 * 2)
ap.add_argument("--first", type=int, default=200, help="first step of the analysed window")
ap.add_argument("--steps", type=int, default=600, help="steps in the analysed window")
ap.add_argument("--dump", default="")
a = ap.parse_args()
if a.build:
    os.makedirs(os.path.dirname(TL_LIB), exist_ok=True)
    sys.path.insert(0, ROOT)
    from subrecon_b200 import native as _n
    subprocess.check_call(["nvcc", *_n.NVCC_FLAGS, "-DSR_TIMELINE", "-I", os.path.join(ROOT, "include"), "-o", TL_LIB,
                           os.path.join(ROOT, "subrecon_b200", "csrc", "subrecon_b200.cu")])
    print("built", TL_LIB)
    sys.exit(0)
os.environ.setdefault("SUBRECON_B200_LIB", TL_LIB)
sys.path.insert(0, ROOT)
from subrecon_b200 import native, palmodel, synth, treeio  # noqa: E402

tree = synth.yule_tree(1000, seed=3)
model = palmodel.host_model("wag")
rates = palmodel.gamma_rates(4, 0.5)
base = synth.simulate_alignment(tree, model, rates, 4096, seed=103)
states = np.tile(base, (1, (a.cols + 4095) // 4096))[:, :a.cols]
ctx = native.Context(0)
ctx.set_option("cherry_tables", a.tables)
ctx.set_model(model.Eval, model.Evec, model.Ievc, model.pi)
ctx.set_rates(rates)
ctx.set_tree(treeio.flatten_tree(tree))
ctx.set_alignment(states)
ctx.run(want_joint=False)
ctx.run(want_joint=False)
L = native.lib()
TLS = 1024
buf = np.zeros((16, TLS, 8), dtype=np.uint32)
L.sr_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
rc = L.sr_debug_timeline(ctx._h, buf.ctypes.data, buf.size)
assert rc == 0, rc
steps, _, _ = native.schedule_dump(treeio.flatten_tree(tree))
n_steps = len(steps)
words = steps[:, 0]
if a.dump:
    np.save(a.dump, buf)
t = buf.astype(np.int64)
nw = int((t[:, 10, 0] != 0).sum())
lo, hi = a.first, min(a.first + a.steps, TLS - 1)
print(f"consumer warps seen: {nw}; window steps [{lo},{hi})")
t0 = t[:nw, lo:hi, :]
nxt = t[:nw, lo + 1:hi + 1, 0]
period = (t[:nw, hi, 0] - t[:nw, lo, 0]) / (hi - lo)
print("mean step period per warp (cycles):", np.round(period, 1))
names = ["wait stage (0->1)", "pre-burst loads/products (1->2)", "rotation wait (2->3)", "burst (3->4)", "release+probe (4->5)",
         "epilogue (5->6)", "tail: rescale/push/store (6->7)", "loop overhead (7->next 0)"]
d = np.stack([t0[:, :, 1] - t0[:, :, 0], t0[:, :, 2] - t0[:, :, 1], t0[:, :, 3] - t0[:, :, 2], t0[:, :, 4] - t0[:, :, 3],
              t0[:, :, 5] - t0[:, :, 4], t0[:, :, 6] - t0[:, :, 5], t0[:, :, 7] - t0[:, :, 6], nxt - t0[:, :, 7]], axis=0)
mvmask = (words[(np.arange(lo, hi)) % n_steps] & 3) != 0
print("phase: mean cycles over warps and MV steps  |  per sub-partition (warp & 3)")
for i, nm in enumerate(names):
    x = d[i][:, mvmask]
    per_sp = [x[[w for w in range(nw) if (w & 3) == sp]].mean() for sp in range(4)]
    print(f"  {nm:36s} {x.mean():8.1f}   | " + " ".join(f"{v:8.1f}" for v in per_sp))
# tensor-pipe view per sub-partition: union of burst intervals
for sp in range(4):
    ws = [w for w in range(nw) if (w & 3) == sp]
    iv = sorted((int(t[w, s, 3]), int(t[w, s, 4])) for w in ws for s in range(lo, hi) if (words[s % n_steps] & 3))
    busy = sum(e - b for b, e in iv)
    span = iv[-1][1] - iv[0][0]
    gaps = [iv[i + 1][0] - iv[i][1] for i in range(len(iv) - 1)]
    print(f"sub-partition {sp}: {len(ws)} warps, bursts cover {100.0 * busy / span:.1f} % of the span; mean burst {busy / len(iv):.0f} cycles, "
          f"mean gap between bursts {np.mean(gaps):.0f} (median {np.median(gaps):.0f}, p90 {np.percentile(gaps, 90):.0f})")
# by step kind
kinds = {}
for s in range(lo, hi):
    w_ = int(words[s % n_steps])
    key = (w_ & 3, (w_ >> 2) & 3, (w_ >> 5) & 1, (w_ >> 6) & 1, (w_ >> 7) & 1)
    kinds.setdefault(key, []).append(s - lo)
print("step kind (mv, n_pre, n_post, rescale, push): count, mean own time outside rotation wait (cycles), burst")
for k_, idx in sorted(kinds.items(), key=lambda kv: -len(kv[1])):
    own = (t0[:, idx, 7] - t0[:, idx, 0]) - d[2][:, idx] - d[0][:, idx]
    print(f"  {k_}: {len(idx):4d}  {own.mean():8.1f}  {d[3][:, idx].mean():8.1f}")
