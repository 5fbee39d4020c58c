"""Host-side mirror of the reference's interface around the hot path.

Same names, argument meaning and error behaviour as the Java classes the native path plugs into:

* `SiteResult`                — /root/reference/src/subrecon/recon/SiteResult.java:52-139 (scan, stable
                                descending bubble sort, tab-delimited `toString`, `roundDouble` + Java `Double.toString`)
* `JointBranchReconstruction` — /root/reference/src/subrecon/recon/JointBranchReconstruction.java:49-80; here ONE
                                object covers a batch of columns (`call()` returns a list of `SiteResult`), because
                                the per-column `Callable`s + thread pool are what the GPU path replaces
* `run_columns`               — the submit / ordered-join / print loop of SubRecon.java:109-150

All arithmetic of the path happens in `libsubrecon_b200.so` (CUDA); this module only flattens inputs
and formats outputs. There is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np

from . import native
from .treeio import AA_ORDER, FlatTree

DELIM = "\t"             # Constants.java:35
SUB_PROB_DELIM = ":"     # Constants.java:36
DEFAULT_PRINT_THRESHOLD = 0.5
DEFAULT_SIG_DIGITS = 2


def round_double(x: float, dec_places: int) -> float:
    """utils/Utils.java:28-31: (double) Math.round(x * (long)10^d) / (long)10^d; Math.round = floor(x + 1/2)."""
    ten = int(math.pow(10.0, float(dec_places)))
    y = x * ten
    if y != y:
        r = 0
    elif math.isinf(y):
        r = (1 << 63) - 1 if y > 0 else -(1 << 63)
    else:
        r = max(-(1 << 63), min((1 << 63) - 1, math.floor(y + 0.5)))
    return float(r) / ten


def java_double_to_string(x: float) -> str:
    """`Double.toString`: shortest repr; plain decimal for 1e-3 <= |x| < 1e7, else d.dddE±n."""
    if x != x:
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    a = abs(x)
    if 1e-3 <= a < 1e7:
        s = repr(a)
        if "e" in s or "E" in s:
            s = f"{a:.17f}".rstrip("0")
            if s.endswith("."):
                s += "0"
        if "." not in s:
            s += ".0"
    else:
        m, e = f"{a:.16e}".split("e")
        digits = repr(a)
        # shortest digits from repr, re-laid-out in Java's scientific form
        mant = digits.replace(".", "").lstrip("0")
        if "e" in mant or "E" in mant:
            mant = mant.split("e")[0].split("E")[0]
        mant = mant.rstrip("0") or "0"
        exp = int(e)
        s = mant[0] + "." + (mant[1:] or "0") + "E" + str(exp)
    return ("-" + s) if x < 0 else s


class SiteResult:
    """recon/SiteResult.java. `branch_probs` is the 20x20 matrix (row = state at node A, column = at node B)."""

    def __init__(self, site, marginal_lnl, branch_probs, threshold=DEFAULT_PRINT_THRESHOLD, sort_by_prob=True,
                 sig_digits=DEFAULT_SIG_DIGITS):
        self.site = int(site)
        self.marginal_lnl = float(marginal_lnl)
        self.sig_digits = sig_digits
        p = np.asarray(branch_probs, dtype=np.float64).reshape(20, 20)
        self.max_prob = max(-1.0, float(p.max()))                     # SiteResult.java:64,75
        self.max_ii_prob = max(-1.0, float(np.diag(p).max()))         # :65,72-74
        ii, jj = np.nonzero(p >= threshold)                           # canonical a-major order, :77-80
        self.probs = [float(p[i, j]) for i, j in zip(ii, jj)]
        self.subs = [AA_ORDER[i] + AA_ORDER[j] for i, j in zip(ii, jj)]
        if sort_by_prob:
            self._sort_by_prob()

    @classmethod
    def from_compact(cls, site, rec, sort_by_prob=True, sig_digits=DEFAULT_SIG_DIGITS):
        """Build from one `sr_compact` record (device-side scan of SiteResult.java:69-83)."""
        self = cls.__new__(cls)
        self.site = int(site)
        self.marginal_lnl = float(rec["lnl"])
        self.sig_digits = sig_digits
        self.max_prob = float(rec["max_prob"])
        self.max_ii_prob = float(rec["max_ii_prob"])
        n = int(rec["n_kept"])
        if n > len(rec["pair"]):
            raise ValueError(f"{n} entries above the threshold; sr_compact keeps {len(rec['pair'])} — rerun with the full joint output")
        self.probs = [float(rec["prob"][i]) for i in range(n)]
        self.subs = [AA_ORDER[int(rec["pair"][i]) // 20] + AA_ORDER[int(rec["pair"][i]) % 20] for i in range(n)]
        if sort_by_prob:
            self._sort_by_prob()
        return self

    @classmethod
    def from_compact_like_java(cls, site, rec, threshold, sort_by_prob=True, sig_digits=DEFAULT_SIG_DIGITS):
        """What the Java host does with a compact record (java/subrecon/recon/NativeRecon.java::reconstruct): SiteResult only
        has the matrix constructor, so the kept entries are put back into an otherwise empty 20x20 matrix, plus carriers for
        maxIIProb / maxProb when those are below the threshold, and the reference's own scan runs on it."""
        n, cap = int(rec["n_kept"]), len(rec["pair"])
        if n > cap:
            raise ValueError(f"compact record overflow: {n} entries kept, capacity {cap}")
        m = np.zeros(400)
        for i in range(n):
            m[int(rec["pair"][i])] = float(rec["prob"][i])
        max_ii, max_p = float(rec["max_ii_prob"]), float(rec["max_prob"])
        if max_ii < threshold:
            m[0] = max(m[0], max_ii)
        if max_p < threshold and max_p > max_ii:
            m[1] = max(m[1], max_p)
        return cls(site, float(rec["lnl"]), m, threshold, sort_by_prob, sig_digits)

    def _sort_by_prob(self):                                           # :96-108, stable descending bubble sort
        pr, sb = self.probs, self.subs
        for i in range(len(pr) - 1):
            for j in range(1, len(pr) - i):
                if pr[j - 1] < pr[j]:
                    pr[j - 1], pr[j] = pr[j], pr[j - 1]
                    sb[j - 1], sb[j] = sb[j], sb[j - 1]

    def get_site(self):
        return self.site

    def get_marginal_lnl(self):
        return self.marginal_lnl

    def get_max_ii_prob(self):
        return self.max_ii_prob

    def get_max_prob(self):
        return self.max_prob

    @staticmethod
    def get_header():                                                  # :119-121
        return DELIM.join(["[HEADER]", "site", "ln[P(D|theta,alpha)]", "P(A=a,B=b|D,theta,alpha)"])

    def __str__(self):                                                 # :124-139
        parts = ["Result", str(self.site + 1), java_double_to_string(round_double(self.marginal_lnl, self.sig_digits))]
        for sub, p in zip(self.subs, self.probs):
            parts.append(sub + SUB_PROB_DELIM + java_double_to_string(round_double(p, self.sig_digits)))
        return DELIM.join(parts)


class JointBranchReconstruction:
    """Batch form of recon/JointBranchReconstruction.java:49-80 on one GPU.

    Arguments keep the reference's meaning: `alignment` = uint8 states[taxon][site], `tree` = FlatTree (root
    with exactly two children A, B), `model` = object with Eval/Evec/Ievc (PAL's MatrixExponential fields),
    `pi` (SubRecon.java:222), `rates` (RateDistribution.getRate), `threshold`, `sig_digits`, `sort_by_prob`,
    `sanity_check` (the hidden -debug flag: checkSumToOne, JointBranchReconstruction.java:157-166).
    """

    def __init__(self, alignment, tree: FlatTree, model, pi, rates, threshold=DEFAULT_PRINT_THRESHOLD,
                 sig_digits=DEFAULT_SIG_DIGITS, sort_by_prob=True, sanity_check=False, device=0, ctx=None):
        self.ctx = ctx or native.Context(device)
        self.threshold, self.sig_digits, self.sort_by_prob, self.sanity_check = threshold, sig_digits, sort_by_prob, sanity_check
        self.ctx.set_model(model.Eval, model.Evec, model.Ievc, pi)
        self.ctx.set_rates(rates)
        self.ctx.set_tree(tree)
        self.ctx.set_alignment(alignment)
        self.n_sites = int(np.asarray(alignment).shape[1])

    def recon(self, site0=0, n=None, compact=False):
        """Raw arrays for columns [site0, site0+n): dict(lnl, joint | compact)."""
        r = self.ctx.run(site0, n, want_joint=not compact, want_compact=compact, threshold=self.threshold)
        if self.sanity_check and r["joint"] is not None:
            s = r["joint"].sum(axis=(1, 2))
            bad = np.nonzero((s < 1.0 - 1e-6) | (s > 1.0 + 1e-6))[0]
            if bad.size:
                raise RuntimeError("ERROR: Failed sanity check. Sum of posterior probs != 1.0. sum=%r" % s[bad[0]])
        return r

    def call(self, site0=0, n=None):
        """-> list[SiteResult] in column order (what `Future.get()` yields, SubRecon.java:123-126)."""
        r = self.recon(site0, n)
        return [SiteResult(site0 + i, r["lnl"][i], r["joint"][i], self.threshold, self.sort_by_prob, self.sig_digits)
                for i in range(len(r["lnl"]))]


def run_columns(results, verbose=False, threshold=DEFAULT_PRINT_THRESHOLD, single_site=False):
    """The print loop of SubRecon.java:102-150 over `SiteResult`s; returns the stdout lines."""
    lines = []
    printing = False
    total = 0.0
    for res in results:
        total += res.get_marginal_lnl()                                   # sequential, column order (:137)
        if verbose or (res.get_max_ii_prob() <= 1.0 - threshold and res.get_max_prob() >= threshold):
            printing = True
            lines.append(str(res))
    if not single_site:
        lines.append("Total lnL: %.10f" % total)                         # :143
    if not printing:
        lines.append("0 sites have non-identical substitution probabilities greater than threshold value "
                     "(threshold=%.5f)" % threshold)
        lines.append("The options -threshold, -nosort and -verbose can be used to control output detail")
    return lines
