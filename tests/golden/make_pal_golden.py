"""Generate tests/golden/pal_golden.json by EXECUTING the reference's own PAL 1.51 bytecode.

Run in the build container only (needs /root/reference/lib/pal.jar):
    python tests/golden/make_pal_golden.py
The jar is interpreted by oracle/jvm/interp.py (no JVM exists in this image). Doubles are stored as
C99 hex strings so the file round-trips bit-exactly.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.jvm.interp import VM  # noqa: E402

JAR = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/lib/pal.jar"


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    vm = VM(JAR)
    out = {"_provenance": "outputs of lib/pal.jar (PAL 1.51) bytecode run under oracle/jvm/interp.py; "
                          "Math.exp/log/pow/sqrt = glibc libm", "gamma": [], "models": []}
    for K, alpha in [(1, 0.5), (2, 0.5), (3, 0.7), (4, 0.5), (4, 1.0), (4, 2.3), (5, 0.05), (6, 0.1), (8, 0.5),
                     (8, 3.0), (16, 0.25), (3, 10.0), (4, 0.16), (4, 0.01)]:
        gr = vm.construct("pal/substmodel/GammaRates", "(ID)V", K, alpha)
        out["gamma"].append({"K": K, "alpha": alpha, "rates": hx([vm.call(gr, "getRate", "(I)D", i) for i in range(K)])})
    from subrecon_b200.palmodel import normalise_cli_frequencies
    rng = np.random.default_rng(4)
    dirichlet = rng.dirichlet(5.0 * np.ones(20))
    dirichlet = normalise_cli_frequencies(dirichlet)
    mle = normalise_cli_frequencies(open(os.path.join(ROOT, "tests/golden/example/freqs")).read().strip().split(","))  # CommandArgs.java:91-111
    tiny = [1e-12] * 3 + [1.0 / 17] * 17                         # exercises the 1e-10 floor + tie breaking
    cases = [("WAG", None), ("JTT", None), ("Dayhoff", None), ("BLOSUM62", None), ("WAG", mle),
             ("BLOSUM62", dirichlet), ("JTT", tiny)]
    ts = [0.0, 5e-10, 2.9e-8, 1e-6, 1e-3, 0.0123, 0.1, 0.75, 2.765, 10.0]
    for name, freqs in cases:
        cls = "pal/substmodel/" + name
        f_in = vm.call(cls, "getOriginalFrequencies", "()[D") if freqs is None else list(freqs)
        report = vm.construct(cls, "([D)V", list(f_in))
        pi = list(vm.call(report, "getEquilibriumFrequencies", "()[D"))          # SubRecon.java:222
        model = vm.construct(cls, "([D)V", list(pi))                              # SubRecon.java:113
        vm.call(model, "setDistance", "(D)V", 0.1)                                # builds the eigensystem (B.2)
        me = model.fields["matrixExp_"]
        rec = {"model": name, "freq_in": hx(f_in), "pi": hx(pi), "pi_model": hx(model.fields["frequency"]),
               "R": hx(model.fields["rate"]), "Eval": hx(me.fields["Eval"]), "Evec": hx(me.fields["Evec"]),
               "Ievc": hx(me.fields["Ievc"]), "P": []}
        for t in ts:
            vm.call(me, "setDistance", "(D)V", t)                                 # MatrixExponential::setDistance only
            P = [[0.0] * 20 for _ in range(20)]
            vm.call(me, "getTransitionProbabilities", "([[D)V", P)
            rec["P"].append({"t": float(t).hex(), "P": hx(P)})
        out["models"].append(rec)
        print(name, "custom" if freqs is not None else "default", "done", vm.n_bytecodes, "bytecodes so far")
    out["_stubbed_host_calls"] = sorted(vm.stubbed)
    with open(os.path.join(ROOT, "tests/golden/pal_golden.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote pal_golden.json")


if __name__ == "__main__":
    main()
