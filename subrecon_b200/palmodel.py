"""Host-side stand-in for the PAL objects the reference's Java host hands to the hot path.

In the real drop-in the host stays Java and PAL supplies these values (INTEGRATION.md shows the
reflection/JNI stub). This module exists so the Python host mirror, `bench.py` and the examples can
run without a JVM: it produces the *inputs* of the native path — equilibrium frequencies `pi`, the
eigensystem `(Eval, Evec, Ievc)` of the normalised rate matrix and the discrete-gamma rates — with
the same conventions:

* `check_frequencies`  — `AbstractRateMatrix::checkFrequencies` (floor 1e-10, residual into the largest, tie break)
* `rate_matrix`        — `fromQToR` / `makeValid` / `normalize`
* `eigensystem`        — LAPACK `eig` of R (PAL uses EISPACK elmhes/hqr2; any exact eigensystem of R gives the same
                         P(t) up to rounding; the parity tests feed the oracle's PAL-faithful one to both sides)
* `gamma_rates`        — median discretisation with PAL's 5e-7-tolerance AS-91 quantile (`GammaDistribution::pointChi2`)

Model tables: `data/pal_models.json` (constants of PAL 1.51's WAG/JTT/Dayhoff/BLOSUM62 classes).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "pal_models.json")
MODEL_IDS = {"wag": "WAG", "jtt": "JTT", "dayhoff": "Dayhoff", "blosum62": "BLOSUM62"}   # Constants.java:26-29


def model_tables():
    with open(_DATA) as fh:
        return json.load(fh)


def check_frequencies(freq) -> np.ndarray:
    f = np.array(freq, dtype=np.float64)
    eps = 1e-10
    s = 0.0
    mx, mi = 0.0, 0
    for i in range(20):
        v = f[i]
        if v < eps:
            f[i] = eps
        if v > mx:
            mx, mi = v, i
        s += f[i]
    f[mi] += 1.0 - s
    for i in range(19):
        for j in range(i + 1, 20):
            if f[i] == f[j]:
                f[i] += eps
                f[j] -= eps
    return f


def rate_matrix(q_upper_rows, freq) -> np.ndarray:
    r = np.zeros((20, 20))
    for i in range(19):
        for jj, q in enumerate(q_upper_rows[i]):
            j = i + 1 + jj
            r[i, j] = q * freq[j]
            r[j, i] = q * freq[i]
    for i in range(20):
        r[i, i] = -(r[i].sum() - r[i, i])
    subst = float(np.sum(-np.diag(r) * freq))
    return r / subst


def eigensystem(R):
    w, v = np.linalg.eig(R)
    w = np.real(w)
    v = np.real(v)
    return np.ascontiguousarray(w), np.ascontiguousarray(v), np.ascontiguousarray(np.linalg.inv(v))


# ---- discrete gamma, PAL conventions ---------------------------------------------------------
def _ln_gamma(x):
    f = 0.0
    if x < 7:
        f = 1.0
        z = x - 1.0
        while True:
            z += 1.0
            if not z < 7.0:
                break
            f *= z
        x = z
        f = -math.log(f)
    z = 1.0 / (x * x)
    return (f + (x - 0.5) * math.log(x) - x + 0.918938533204673 +
            (((-0.000595238095238 * z + 0.000793650793651) * z - 0.002777777777778) * z + 0.083333333333333) / x)


def _incomplete_gamma(x, alpha, lga):
    accurate, overflow = 1e-8, 1e30
    if x == 0.0:
        return 0.0
    factor = math.exp(alpha * math.log(x) - x - lga)
    if x > 1 and x >= alpha:
        a = 1 - alpha
        b = a + x + 1
        term = 0.0
        pn = [1.0, x, x + 1, x * b, 0.0, 0.0]
        gin = pn[2] / pn[3]
        while True:
            a += 1
            b += 2
            term += 1
            an = a * term
            pn[4] = b * pn[2] - an * pn[0]
            pn[5] = b * pn[3] - an * pn[1]
            if pn[5] != 0:
                rn = pn[4] / pn[5]
                dif = abs(gin - rn)
                if dif <= accurate and dif <= accurate * rn:
                    break
                gin = rn
            pn[0:4] = pn[2:6]
            if abs(pn[4]) >= overflow:
                pn[0:4] = [v / overflow for v in pn[0:4]]
        return 1 - factor * gin
    gin = term = 1.0
    rn = alpha
    while True:
        rn += 1
        term *= x / rn
        gin += term
        if not term > accurate:
            break
    return gin * (factor / alpha)


def _point_normal(prob):
    a0, a1, a2, a3, a4 = -0.322232431088, -1.0, -0.342242088547, -0.0204231210245, -0.453642210148e-4
    b0, b1, b2, b3, b4 = 0.0993484626060, 0.588581570495, 0.531103462366, 0.103537752850, 0.0038560700634
    p1 = prob if prob < 0.5 else 1 - prob
    y = math.sqrt(math.log(1 / (p1 * p1)))
    z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) / ((((y * b4 + b3) * y + b2) * y + b1) * y + b0)
    return -z if prob < 0.5 else z


def _point_chi2(p, v):
    e, aa = 0.5e-6, 0.6931471805
    if p < 0.000002 or p > 0.999998 or v <= 0:
        raise ValueError("Arguments out of range")
    g = _ln_gamma(v / 2)
    xx = v / 2
    c = xx - 1
    if v < -1.24 * math.log(p):
        ch = math.pow(p * xx * math.exp(g + xx * aa), 1 / xx)
        if ch - e < 0:
            return ch
    elif v > 0.32:
        x = math.sqrt(2.0) * (_point_normal(0.5 * (2.0 * p - 1.0) + 0.5) / math.sqrt(2.0))
        p1 = 0.222222 / v
        ch = v * math.pow(x * math.sqrt(p1) + 1 - p1, 3.0)
        if ch > 2.2 * v + 6:
            ch = -2 * (math.log(1 - p) - c * math.log(0.5 * ch) + g)
    else:
        ch = 0.4
        a = math.log(1 - p)
        while True:
            q = ch
            p1 = 1 + ch * (4.67 + ch)
            p2 = ch * (6.73 + ch * (6.66 + ch))
            t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2
            ch -= (1 - math.exp(a + g + 0.5 * ch + c * aa) * p2 / p1) / t
            if not abs(q / ch - 1) - 0.01 > 0:
                break
    while True:
        q = ch
        p1 = 0.5 * ch
        t = _incomplete_gamma(p1, xx, g)
        p2 = p - t
        t = p2 * math.exp(xx * aa + g + p1 - c * math.log(ch))
        b = t / ch
        a = 0.5 * t - b * c
        s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420
        s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520
        s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520
        s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040
        s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520
        s6 = (120 + c * (346 + 127 * c)) / 5040
        ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))))
        if not abs(q / ch - 1) > e:
            return ch


def gamma_rates(K: int, alpha: float) -> np.ndarray:
    """`new GammaRates(K, alpha)` (SubRecon.java:231): median method, renormalised to mean 1."""
    r = np.array([0.5 * (1.0 / alpha) * _point_chi2((2.0 * i + 1.0) / (2.0 * K), 2.0 * alpha) for i in range(K)])
    mean = 0.0
    for v in r:
        mean += v
    return r / (mean / K)


@dataclass
class HostModel:
    """What `SubRecon.init` derives before scheduling columns (SubRecon.java:198,222-223,113)."""
    name: str
    pi: np.ndarray        # reportModel.getEquilibriumFrequencies(): first clean-up pass
    log_pi: np.ndarray
    pi_model: np.ndarray  # frequencies inside the per-column model: second clean-up pass
    q_upper: np.ndarray   # [190]
    Eval: np.ndarray
    Evec: np.ndarray
    Ievc: np.ndarray


def host_model(model_id: str, frequencies=None) -> HostModel:
    tabs = model_tables()
    key = MODEL_IDS.get(model_id.lower())
    if key is None:
        raise ValueError("ERROR: Model identifier not recognised")      # SubRecon.java:297
    tab = tabs[key]
    f0 = np.array(tab["freq"] if frequencies is None else frequencies, dtype=np.float64)
    pi = check_frequencies(f0)
    with np.errstate(divide="ignore"):
        log_pi = np.log(pi)
    pi_model = check_frequencies(pi)
    R = rate_matrix(tab["q_upper_rows"], pi_model)
    w, v, iv = eigensystem(R)
    q = np.array([x for row in tab["q_upper_rows"] for x in row])
    return HostModel(key, pi, log_pi, pi_model, q, w, v, iv)


def normalise_cli_frequencies(values):
    """`-pi` handling of cli/CommandArgs.java:91-111: sequential sum, every value divided by it."""
    vals = [float(v) for v in values]
    if len(vals) != 20:
        raise ValueError("Error: if specifying amino acid frequencies, exactly 20 values must be provided "
                         "and delimited by comma with no spaces")
    s = 0.0
    for v in vals:
        s += v
    return [v / s for v in vals]
