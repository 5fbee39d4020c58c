"""Pin the oracle's PAL restatement (oracle/pal_numerics.c) against golden vectors produced by EXECUTING the
reference's pal.jar bytecode (tests/golden/make_pal_golden.py). Bit-exact unless stated."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from subrecon_b200 import palmodel

HERE = os.path.dirname(os.path.abspath(__file__))


def unhex(xs, shape=None):
    a = np.array([float.fromhex(x) for x in xs], dtype=np.float64)
    return a.reshape(shape) if shape else a


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "pal_golden.json")) as fh:
        return json.load(fh)


def test_gamma_rates_bit_exact(golden):
    for g in golden["gamma"]:
        want = unhex(g["rates"])
        got = oracle.gamma_rates(g["K"], g["alpha"])
        assert np.array_equal(got, want), (g["K"], g["alpha"], got, want)
        # the host-side Python stand-in follows the same algorithm
        assert np.array_equal(palmodel.gamma_rates(g["K"], g["alpha"]), want)


def test_survey_rate_values(golden):
    # SURVEY.md §8d quotes these (K=4 / K=8, alpha=0.5)
    assert oracle.gamma_rates(4, 0.5).tolist() == [0.02907775442778477, 0.28071453392572127, 0.9247730548197041, 2.76543465682679]
    assert oracle.gamma_rates(8, 0.5).tolist() == [0.00665652751376567, 0.060913729499308324, 0.1751742809983289,
                                                   0.3631060336461068, 0.6526397102566377, 1.1043643793752917,
                                                   1.8806851663352815, 3.7564601723752795]


def test_gamma_rejects_out_of_range():
    with pytest.raises(ValueError):
        oracle.gamma_rates(4, 0.0)          # alpha = 0 passes SubRecon's check but throws in pointChi2


def test_model_pipeline_bit_exact(golden):
    tabs = oracle.model_tables()
    for m in golden["models"]:
        q = np.array([x for row in tabs[m["model"]]["q_upper_rows"] for x in row])
        f_in = unhex(m["freq_in"])
        pi = oracle.check_frequencies(f_in)
        assert np.array_equal(pi, unhex(m["pi"])), m["model"]
        pim = oracle.check_frequencies(pi)
        assert np.array_equal(pim, unhex(m["pi_model"]))
        assert np.array_equal(palmodel.check_frequencies(f_in), pi)
        R = oracle.rate_matrix(q, pim)
        assert np.array_equal(R, unhex(m["R"], (20, 20)))
        w, v, iv = oracle.eigensystem(R)
        assert np.array_equal(w, unhex(m["Eval"]))
        assert np.array_equal(v, unhex(m["Evec"], (20, 20)))
        assert np.array_equal(iv, unhex(m["Ievc"], (20, 20)))
        for rec in m["P"]:
            t = float.fromhex(rec["t"])
            assert np.array_equal(oracle.transition_probs(w, v, iv, t), unhex(rec["P"], (20, 20))), (m["model"], t)


def test_oracle_model_wrapper_matches_golden(golden):
    m = golden["models"][0]
    om = oracle.OracleModel("WAG")
    assert np.array_equal(om.pi, unhex(m["pi"])) and np.array_equal(om.Eval, unhex(m["Eval"]))


def test_host_eigensystem_gives_same_P(golden):
    # the Python host stand-in uses LAPACK instead of PAL's EISPACK port: same P(t) to rounding
    for name, mid in (("WAG", "wag"), ("JTT", "jtt"), ("Dayhoff", "dayhoff"), ("BLOSUM62", "blosum62")):
        om, hm = oracle.OracleModel(name), palmodel.host_model(mid)
        assert np.array_equal(om.pi, hm.pi) and np.array_equal(om.pi_model, hm.pi_model)
        for t in (1e-3, 0.1, 2.0):
            P1 = oracle.transition_probs(om.Eval, om.Evec, om.Ievc, t)
            P2 = oracle.transition_probs(hm.Eval, hm.Evec, hm.Ievc, t)
            assert np.max(np.abs(P1 - P2)) < 5e-15


def test_tables_match_between_oracle_and_host():
    a, b = oracle.model_tables(), palmodel.model_tables()
    for k in ("WAG", "JTT", "Dayhoff", "BLOSUM62"):
        assert a[k] == b[k]
    # SURVEY Appendix C checksums
    sums = {"WAG": 205.391182, "JTT": 1117.0655370335, "Dayhoff": 687.7114296451, "BLOSUM62": 215.9034713181}
    for k, s in sums.items():
        assert abs(sum(sum(r) for r in a[k]["q_upper_rows"]) - s) < 1e-9


@pytest.mark.skipif(not os.path.exists("/root/reference/lib/pal.jar"), reason="reference jar only in the build container")
def test_tables_come_from_the_jar():
    from oracle.jvm import extract_pal_tables as ex
    from oracle.jvm.classfile import Jar
    jar = Jar("/root/reference/lib/pal.jar")
    tabs = oracle.model_tables()
    for k in ("WAG", "JTT", "Dayhoff", "BLOSUM62"):
        assert ex.extract(jar, "pal/substmodel/" + k) == tabs[k]


@pytest.mark.skipif(not os.path.exists("/root/reference/lib/pal.jar"), reason="reference jar only in the build container")
def test_interpreter_runs_the_jar_live():
    from oracle.jvm.interp import VM
    vm = VM("/root/reference/lib/pal.jar")
    gr = vm.construct("pal/substmodel/GammaRates", "(ID)V", 5, 0.8)
    got = [vm.call(gr, "getRate", "(I)D", i) for i in range(5)]
    assert got == oracle.gamma_rates(5, 0.8).tolist()
