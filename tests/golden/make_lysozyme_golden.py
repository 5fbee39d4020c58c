"""Golden outputs for BASELINE config #1 (example/ primate lysozyme, WAG + mle.wag.freqs, K=4, alpha=0.5).

Independent of the C oracle: the transition matrices, frequencies and gamma rates come from EXECUTING
the reference's PAL 1.51 bytecode (oracle/jvm/interp.py), and the per-column arithmetic is a line-by-line
Python transliteration of recon()/downTreeMarginal()/getLnSumComponents()
(/root/reference/src/subrecon/recon/JointBranchReconstruction.java:82-140,191-247,
 /root/reference/src/subrecon/utils/Utils.java:47-80) in IEEE doubles, same operation order.
Run in the build container only:  python tests/golden/make_lysozyme_golden.py
Writes tests/golden/lysozyme_golden.npz (lnl[130], joint[130,20,20], rates[4], pi[20]).
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.jvm.interp import VM  # noqa: E402
from subrecon_b200 import treeio  # noqa: E402

EX = os.path.join(ROOT, "tests/golden/example/")
TINY = 1e-10


def ln(x):
    return -math.inf if x == 0.0 else math.log(x)


def ln_sum_components(v):
    if len(v) == 1:
        return v[0]
    imax, mx = -1, -sys.float_info.max
    for i, x in enumerate(v):
        if x > mx:
            mx, imax = x, i
    r = v[imax]
    s = 1.0
    for i, x in enumerate(v):
        if i == imax:
            continue
        s += math.exp(x - v[imax])
    return r + math.log(s)


def main():
    vm = VM("/root/reference/lib/pal.jar")
    tree = treeio.read_tree(EX + "root.tre")
    names, seqs = treeio.read_fasta(EX + "prot.fasta")
    row = {n: i for i, n in enumerate(names)}
    from subrecon_b200.palmodel import normalise_cli_frequencies
    freqs = normalise_cli_frequencies(open(EX + "freqs").read().strip().split(","))   # CommandArgs.java:91-111
    report = vm.construct("pal/substmodel/WAG", "([D)V", list(freqs))
    pi = list(vm.call(report, "getEquilibriumFrequencies", "()[D"))
    log_pi = [ln(x) for x in pi]
    gr = vm.construct("pal/substmodel/GammaRates", "(ID)V", 4, 0.5)
    rates = [vm.call(gr, "getRate", "(I)D", i) for i in range(4)]
    K = len(rates)
    model = vm.construct("pal/substmodel/WAG", "([D)V", list(pi))
    vm.call(model, "setDistance", "(D)V", 0.1)
    me = model.fields["matrixExp_"]          # PAL recomputes the identical eigensystem on every call

    cache = {}

    def P_of(t):
        if t not in cache:
            vm.call(me, "setDistance", "(D)V", t)
            P = [[0.0] * 20 for _ in range(20)]
            vm.call(me, "getTransitionProbabilities", "([[D)V", P)
            cache[t] = P
        return cache[t]

    # the reference's own tip lookup, executed from the jar (molevo/AdvancedAlignmentAminoAcid.java:34-40)
    aa_type = vm.call("pal/datatype/DataTypeTool", "getUniverisalAminoAcids", "()Lpal/datatype/MolecularDataType;")
    state_cache = {}

    def state_of(name, site):
        c = seqs[row[name]][site]
        if c not in state_cache:
            state_cache[c] = vm.call(aa_type, "getState", "(C)I", ord(c))
        return state_cache[c]

    def down(v, site, count, rate):
        cond = [0.0] * 20
        if not tree.children[v]:
            st = state_of(tree.name[v], site)
            if 0 <= st < 20:
                cond[st] = 1.0
            else:
                cond = [1.0] * 20
        else:
            cond = [1.0] * 20
            for c in tree.children[v]:
                P = P_of(tree.blen[c] * rate)
                cc = down(c, site, count, rate)
                for i in range(20):
                    sm = 0.0
                    for j in range(20):
                        sm += P[i][j] * cc[j]
                    cond[i] *= sm
        big = 0.0
        for i in range(20):
            big = max(big, cond[i])
        for i in range(20):
            cond[i] /= (big + TINY)
        count[0] += math.log(big + TINY)
        return cond

    nA, nB = tree.children[0]
    L = len(seqs[0])
    lnl = np.zeros(L)
    joint = np.zeros((L, 20, 20))
    for site in range(L):
        mix = [0.0] * (400 * K)
        for k in range(K):
            rate = rates[k]
            ca, cb = [0.0], [0.0]
            la = [ln(x) for x in down(nA, site, ca, rate)]
            lb = [ln(x) for x in down(nB, site, cb, rate)]
            corr = ca[0] + cb[0]
            P = P_of((tree.blen[nA] + tree.blen[nB]) * rate)
            for a in range(20):
                lat = log_pi[a] + la[a]
                for b in range(20):
                    scaled = lat + ln(P[a][b]) + lb[b]
                    mix[a * 20 * K + b * K + k] = scaled + corr
        ls = ln_sum_components(mix)
        for a in range(20):
            for b in range(20):
                z = a * 20 * K + b * K
                joint[site, a, b] = math.exp(ln_sum_components(mix[z:z + K]) - ls)
        lnl[site] = ls - math.log(K)
    tot = 0.0
    for x in lnl:
        tot += x
    print("Total lnL: %.10f" % tot, " distinct P matrices from the jar:", len(cache))
    np.savez_compressed(os.path.join(ROOT, "tests/golden/lysozyme_golden.npz"), lnl=lnl, joint=joint,
                        rates=np.array(rates), pi=np.array(pi))


if __name__ == "__main__":
    main()
