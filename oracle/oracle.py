"""ctypes front end of the CPU oracle (`oracle/libsubrecon_oracle.so`). TEST INFRASTRUCTURE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module; the CUDA product path never does (it fails loudly without its own
extension instead).
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsubrecon_oracle.so")
_D = ctypes.POINTER(ctypes.c_double)
_I32 = ctypes.POINTER(ctypes.c_int32)
_U8 = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("pal_numerics.c", "recon_oracle.c", "oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsubrecon_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Problem(ctypes.Structure):
    _fields_ = [("n_nodes", ctypes.c_int), ("child_off", _I32), ("child_idx", _I32), ("blen", _D),
                ("leaf_row", _I32), ("root", ctypes.c_int), ("n_taxa", ctypes.c_int),
                ("n_sites", ctypes.c_int64), ("states", _U8), ("K", ctypes.c_int), ("rates", _D),
                ("q_upper", _D), ("pi_model", _D), ("log_pi", _D)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.pal_check_frequencies.argtypes = [_D]
        _lib.pal_rate_matrix.argtypes = [_D, _D, _D]
        _lib.pal_eigensystem.argtypes = [_D, _D, _D, _D]
        _lib.pal_transition_probs.argtypes = [_D, _D, _D, ctypes.c_double, _D]
        _lib.pal_gamma_rates.argtypes = [ctypes.c_int, ctypes.c_double, _D]
        _lib.sro_run.argtypes = [ctypes.POINTER(_Problem), ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                 ctypes.c_int, _D, _D]
    return _lib


def _dp(a):
    return a.ctypes.data_as(_D)


def model_tables():
    with open(os.path.join(_HERE, "pal_models.json")) as fh:
        return json.load(fh)


def check_frequencies(f):
    out = np.array(f, dtype=np.float64)
    lib().pal_check_frequencies(_dp(out))
    return out


def rate_matrix(q_upper, freq):
    q = np.ascontiguousarray(q_upper, dtype=np.float64)
    f = np.ascontiguousarray(freq, dtype=np.float64)
    R = np.zeros((20, 20))
    lib().pal_rate_matrix(_dp(q), _dp(f), _dp(R))
    return R


def eigensystem(R):
    R = np.ascontiguousarray(R, dtype=np.float64)
    w, v, iv = np.zeros(20), np.zeros((20, 20)), np.zeros((20, 20))
    rc = lib().pal_eigensystem(_dp(R), _dp(w), _dp(v), _dp(iv))
    if rc:
        raise ArithmeticError(f"pal_eigensystem rc={rc}")
    return w, v, iv


def transition_probs(w, v, iv, t):
    P = np.zeros((20, 20))
    lib().pal_transition_probs(_dp(np.ascontiguousarray(w)), _dp(np.ascontiguousarray(v)),
                               _dp(np.ascontiguousarray(iv)), float(t), _dp(P))
    return P


def gamma_rates(K, alpha):
    r = np.zeros(K)
    if lib().pal_gamma_rates(int(K), float(alpha), _dp(r)):
        raise ValueError("Arguments out of range")
    return r


class OracleModel:
    """SubRecon.init's model bookkeeping with PAL-faithful numerics (SubRecon.java:198,222-223,113)."""

    def __init__(self, name, frequencies=None):
        tab = model_tables()[name]
        self.name = name
        self.q_upper = np.array([x for row in tab["q_upper_rows"] for x in row], dtype=np.float64)
        f0 = np.array(tab["freq"] if frequencies is None else frequencies, dtype=np.float64)
        self.pi = check_frequencies(f0)                 # reportModel.getEquilibriumFrequencies()
        with np.errstate(divide="ignore"):
            self.log_pi = np.log(self.pi)               # Utils.getLnValues(pi)
        self.pi_model = check_frequencies(self.pi)      # per-column `new <Model>(pi)`
        self.R = rate_matrix(self.q_upper, self.pi_model)
        self.Eval, self.Evec, self.Ievc = eigensystem(self.R)


def run(flat_tree, states, model: OracleModel, rates, site0=0, n=None, mode=0, threads=1, want_joint=True):
    """-> (lnl[n], joint[n,20,20] or None) for columns [site0, site0+n)."""
    states = np.ascontiguousarray(states, dtype=np.uint8)
    n_taxa, n_sites = states.shape
    if n is None:
        n = n_sites - site0
    rates = np.ascontiguousarray(rates, dtype=np.float64)
    keep = [np.ascontiguousarray(flat_tree.child_off, dtype=np.int32),
            np.ascontiguousarray(flat_tree.child_idx, dtype=np.int32),
            np.ascontiguousarray(flat_tree.blen, dtype=np.float64),
            np.ascontiguousarray(flat_tree.leaf_row, dtype=np.int32)]
    p = _Problem(flat_tree.n_nodes, keep[0].ctypes.data_as(_I32), keep[1].ctypes.data_as(_I32), _dp(keep[2]),
                 keep[3].ctypes.data_as(_I32), flat_tree.root, n_taxa, n_sites, states.ctypes.data_as(_U8),
                 len(rates), _dp(rates), _dp(model.q_upper), _dp(model.pi), _dp(model.log_pi))
    lnl = np.zeros(n)
    joint = np.zeros((n, 20, 20)) if want_joint else None
    rc = lib().sro_run(ctypes.byref(p), site0, n, mode, threads, _dp(lnl), _dp(joint) if want_joint else None)
    if rc:
        raise RuntimeError(f"sro_run rc={rc}")
    return lnl, joint
