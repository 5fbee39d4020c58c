"""ctypes binding of `libsubrecon_b200.so` (the C ABI in `include/subrecon_b200.h`).

This is the same boundary a JNI stub binds. There is no CPU fallback: if the CUDA
extension is missing or no B200 is present, construction raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# SUBRECON_B200_LIB: an instrumented build of the same sources (tools/timeline.py); never a different implementation
LIB_PATH = os.environ.get("SUBRECON_B200_LIB") or os.path.join(_PKG, "libsubrecon_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

_D = ctypes.POINTER(ctypes.c_double)
_I32 = ctypes.POINTER(ctypes.c_int32)
_U8 = ctypes.POINTER(ctypes.c_uint8)

EXPORTS = ["sr_create", "sr_destroy", "sr_last_error", "sr_set_model", "sr_set_rates", "sr_set_tree",
           "sr_set_alignment", "sr_run", "sr_run_device", "sr_run_streamed", "sr_host_alloc", "sr_host_free",
           "sr_set_option", "sr_get_profile", "sr_profile_reset", "sr_schedule_info", "sr_measure_fp64_peak",
           "sr_debug_pmatrix", "sr_debug_cherry_table", "sr_schedule_dump", "sr_schedule_records", "sr_compact_cap_for",
           "sr_compact_stride", "sr_multi_create", "sr_multi_destroy", "sr_multi_last_error", "sr_multi_count", "sr_multi_ctx",
           "sr_multi_set_model", "sr_multi_set_rates", "sr_multi_set_tree", "sr_multi_set_option", "sr_multi_run_streamed"]


class SrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[sr_status {code}] {msg}")
        self.code = code


class Compact(ctypes.Structure):
    _fields_ = [("lnl", ctypes.c_double), ("max_ii_prob", ctypes.c_double), ("max_prob", ctypes.c_double),
                ("n_kept", ctypes.c_int32), ("interesting", ctypes.c_int32),
                ("pair", ctypes.c_int16 * 8), ("prob", ctypes.c_double * 8)]


COMPACT_DTYPE = np.dtype([("lnl", "<f8"), ("max_ii_prob", "<f8"), ("max_prob", "<f8"), ("n_kept", "<i4"),
                          ("interesting", "<i4"), ("pair", "<i2", (8,)), ("prob", "<f8", (8,))])
assert COMPACT_DTYPE.itemsize == ctypes.sizeof(Compact)


def compact_dtype(cap=8):
    """numpy view of a compact record of capacity `cap` (include/subrecon_b200.h::sr_compact): 32-byte head, int16 pair[cap]
    padded to 8 bytes, double prob[cap]. cap = 8 is `sr_compact` / COMPACT_DTYPE."""
    cap = int(cap)
    pb = (2 * cap + 7) // 8 * 8
    return np.dtype({"names": ["lnl", "max_ii_prob", "max_prob", "n_kept", "interesting", "pair", "prob"],
                     "formats": ["<f8", "<f8", "<f8", "<i4", "<i4", ("<i2", (cap,)), ("<f8", (cap,))],
                     "offsets": [0, 8, 16, 24, 28, 32, 32 + pb], "itemsize": 32 + pb + 8 * cap})


def compact_cap_for(threshold):
    """Entries a compact record must hold so that `threshold` loses nothing: min(400, floor(1/threshold))."""
    return int(lib().sr_compact_cap_for(float(threshold)))


class Profile(ctypes.Structure):
    _fields_ = [("ms_pbuild", ctypes.c_double), ("n_pbuild", ctypes.c_int64),
                ("ms_prune", ctypes.c_double), ("n_prune", ctypes.c_int64),
                ("ms_root", ctypes.c_double), ("n_root", ctypes.c_int64),
                ("columns_pruned", ctypes.c_int64), ("columns_in", ctypes.c_int64), ("columns_unique", ctypes.c_int64),
                ("ms_index", ctypes.c_double), ("n_index", ctypes.c_int64)]


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_PKG, "csrc", "subrecon_b200.cu")
    deps = [src, os.path.join(_PKG, "csrc", "kernels.cuh"), os.path.join(_ROOT, "include", "subrecon_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PATH) for d in deps):
        return LIB_PATH
    subprocess.check_call(["nvcc", *NVCC_FLAGS, "-o", LIB_PATH, src])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SrError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp = ctypes.c_void_p
        L.sr_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
        L.sr_destroy.argtypes = [vp]
        L.sr_destroy.restype = None
        L.sr_last_error.argtypes = [vp]
        L.sr_last_error.restype = ctypes.c_char_p
        L.sr_set_model.argtypes = [vp, _D, _D, _D, _D]
        L.sr_set_rates.argtypes = [vp, ctypes.c_int, _D]
        L.sr_set_tree.argtypes = [vp, ctypes.c_int32, _I32, _I32, _D, _I32, ctypes.c_int32]
        L.sr_set_alignment.argtypes = [vp, ctypes.c_int32, ctypes.c_int64, vp, ctypes.c_int64]
        L.sr_run.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, vp, vp, vp, ctypes.c_double]
        L.sr_run_device.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, vp, vp, vp, ctypes.c_double, vp]
        L.sr_run_streamed.argtypes = [vp, ctypes.c_int32, ctypes.c_int64, vp, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int64, vp, vp, vp, ctypes.c_double]
        L.sr_host_alloc.argtypes = [ctypes.c_size_t]
        L.sr_host_alloc.restype = vp
        L.sr_host_free.argtypes = [vp]
        L.sr_host_free.restype = None
        L.sr_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
        L.sr_get_profile.argtypes = [vp, ctypes.POINTER(Profile)]
        L.sr_profile_reset.argtypes = [vp]
        L.sr_schedule_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32),
                                       ctypes.POINTER(ctypes.c_double)]
        L.sr_measure_fp64_peak.argtypes = [vp, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        L.sr_debug_pmatrix.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, _D]
        L.sr_debug_cherry_table.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, _D, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32)]
        L.sr_schedule_dump.argtypes = [ctypes.c_int32, _I32, _I32, _D, _I32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int64, _I32, _I32, ctypes.POINTER(ctypes.c_int64),
                                       ctypes.POINTER(ctypes.c_int32)]
        L.sr_compact_cap_for.argtypes = [ctypes.c_double]
        L.sr_compact_stride.argtypes = [ctypes.c_int]
        L.sr_compact_stride.restype = ctypes.c_size_t
        L.sr_multi_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, _I32]
        L.sr_multi_destroy.argtypes = [vp]
        L.sr_multi_destroy.restype = None
        L.sr_multi_last_error.argtypes = [vp]
        L.sr_multi_last_error.restype = ctypes.c_char_p
        L.sr_multi_count.argtypes = [vp]
        L.sr_multi_ctx.argtypes = [vp, ctypes.c_int]
        L.sr_multi_ctx.restype = vp
        L.sr_multi_set_model.argtypes = [vp, _D, _D, _D, _D]
        L.sr_multi_set_rates.argtypes = [vp, ctypes.c_int, _D]
        L.sr_multi_set_tree.argtypes = [vp, ctypes.c_int32, _I32, _I32, _D, _I32, ctypes.c_int32]
        L.sr_multi_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
        L.sr_multi_run_streamed.argtypes = [vp, ctypes.c_int32, ctypes.c_int64, vp, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int64, vp, vp, vp, ctypes.c_double]
        L.sr_schedule_records.argtypes = [ctypes.c_int32, _I32, _I32, _D, _I32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_int64, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), _I32,
                                          ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        _lib = L
    return _lib


STEP_INTS = 16


def schedule_dump(ft, rescale_every=8, cherry_tables=True):
    """Host-only: the fused pruning schedule for a FlatTree -> (steps[n,16] int32, node_slot[n_nodes], stack_depth).
    Step columns: word, alignment rows of gathers 0..2; for a clade gather its second row (columns 4..6), its third row
    (7..9) and its table's first row (10..12); 3 unused (include/subrecon_b200.h::sr_schedule_dump)."""
    off = np.ascontiguousarray(ft.child_off, dtype=np.int32)
    idx = np.ascontiguousarray(ft.child_idx, dtype=np.int32)
    bl = np.ascontiguousarray(ft.blen, dtype=np.float64)
    lr = np.ascontiguousarray(ft.leaf_row, dtype=np.int32)
    cap = int(ft.n_nodes) + 8
    steps = np.zeros((cap, STEP_INTS), dtype=np.int32)
    node_slot = np.zeros(ft.n_nodes, dtype=np.int32)
    n = ctypes.c_int64()
    depth = ctypes.c_int32()
    rc = lib().sr_schedule_dump(int(ft.n_nodes), off.ctypes.data_as(_I32), idx.ctypes.data_as(_I32), bl.ctypes.data_as(_D),
                                lr.ctypes.data_as(_I32), int(ft.root), int(rescale_every), int(bool(cherry_tables)), cap,
                                steps.ctypes.data_as(_I32), node_slot.ctypes.data_as(_I32), ctypes.byref(n), ctypes.byref(depth))
    if rc:
        raise SrError(rc, "sr_schedule_dump failed")
    return steps[:n.value].copy(), node_slot, depth.value


def schedule_records(ft, rescale_every=8, cherry_tables=True):
    """Host-only: the record streams `prune_kernel` reads (csrc/prune.cuh) -> (crec[n,4] uint32, prec[n,8] uint32,
    clade_gathers[m,4] int32)."""
    off = np.ascontiguousarray(ft.child_off, dtype=np.int32)
    idx = np.ascontiguousarray(ft.child_idx, dtype=np.int32)
    bl = np.ascontiguousarray(ft.blen, dtype=np.float64)
    lr = np.ascontiguousarray(ft.leaf_row, dtype=np.int32)
    cap = int(ft.n_nodes) + 8
    crec = np.zeros((cap, 4), dtype=np.uint32)
    prec = np.zeros((cap, 8), dtype=np.uint32)
    cg = np.zeros((3 * cap, 4), dtype=np.int32)
    n, m = ctypes.c_int64(), ctypes.c_int64()
    u32 = ctypes.POINTER(ctypes.c_uint32)
    rc = lib().sr_schedule_records(int(ft.n_nodes), off.ctypes.data_as(_I32), idx.ctypes.data_as(_I32), bl.ctypes.data_as(_D),
                                   lr.ctypes.data_as(_I32), int(ft.root), int(rescale_every), int(bool(cherry_tables)), cap,
                                   crec.ctypes.data_as(u32), prec.ctypes.data_as(u32), cg.ctypes.data_as(_I32),
                                   ctypes.byref(n), ctypes.byref(m))
    if rc:
        raise SrError(rc, "sr_schedule_records failed")
    return crec[:n.value].copy(), prec[:n.value].copy(), cg[:m.value].copy()


class PinnedArray:
    """numpy view over `sr_host_alloc` memory (page-locked; DMA'd without a bounce copy)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = lib().sr_host_alloc(max(n, 1))
        if not self.ptr:
            raise SrError(-3, f"sr_host_alloc({n}) failed")
        buf = (ctypes.c_uint8 * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().sr_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU's worth of the hot path (replaces the reference's thread pool, SubRecon.java:109)."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        rc = lib().sr_create(ctypes.byref(self._h), int(device))
        if rc:
            raise SrError(rc, (lib().sr_last_error(None) or b"").decode())
        self.device = device
        self.compact_cap = 8

    def close(self):
        if self._h:
            lib().sr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise SrError(rc, (lib().sr_last_error(self._h) or b"").decode())

    @staticmethod
    def _d(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return a, a.ctypes.data_as(_D)

    def set_model(self, Eval, Evec, Ievc, pi):
        k = [self._d(x) for x in (Eval, Evec, Ievc, pi)]
        if k[0][0].size != 20 or k[1][0].size != 400 or k[2][0].size != 400 or k[3][0].size != 20:
            raise SrError(-1, "model arrays must have 20/400/400/20 elements")
        self._ck(lib().sr_set_model(self._h, k[0][1], k[1][1], k[2][1], k[3][1]))

    def set_rates(self, rates):
        r, rp = self._d(rates)
        self._ck(lib().sr_set_rates(self._h, int(r.size), rp))

    def set_tree(self, ft):
        off = np.ascontiguousarray(ft.child_off, dtype=np.int32)
        idx = np.ascontiguousarray(ft.child_idx, dtype=np.int32)
        bl = np.ascontiguousarray(ft.blen, dtype=np.float64)
        lr = np.ascontiguousarray(ft.leaf_row, dtype=np.int32)
        self._ck(lib().sr_set_tree(self._h, int(ft.n_nodes), off.ctypes.data_as(_I32), idx.ctypes.data_as(_I32),
                                   bl.ctypes.data_as(_D), lr.ctypes.data_as(_I32), int(ft.root)))

    def set_alignment(self, states):
        st = np.ascontiguousarray(states, dtype=np.uint8)
        if st.ndim != 2:
            raise SrError(-1, "states must be [taxon][site]")
        self._ck(lib().sr_set_alignment(self._h, st.shape[0], st.shape[1], st.ctypes.data, st.shape[1]))
        self.n_sites = st.shape[1]

    def set_option(self, name, value):
        self._ck(lib().sr_set_option(self._h, name.encode(), int(value)))
        if name == "compact_cap":
            self.compact_cap = int(value)

    def run(self, site0=0, n=None, want_joint=True, want_compact=False, threshold=0.5, out=None):
        """-> dict(lnl=[n], joint=[n,20,20]|None, compact=structured[n]|None); `out` may supply (pinned) arrays."""
        if n is None:
            n = self.n_sites - site0
        out = out or {}
        lnl = out.get("lnl") if out.get("lnl") is not None else np.empty(n, dtype=np.float64)
        joint = (out.get("joint") if out.get("joint") is not None else np.empty((n, 20, 20), dtype=np.float64)) if want_joint else None
        comp = (out.get("compact") if out.get("compact") is not None else np.zeros(n, dtype=compact_dtype(self.compact_cap))) if want_compact else None
        self._ck(lib().sr_run(self._h, int(site0), int(n), lnl.ctypes.data, joint.ctypes.data if want_joint else None,
                              comp.ctypes.data if want_compact else None, float(threshold)))
        return {"lnl": lnl, "joint": joint, "compact": comp}

    def run_streamed(self, states, site0=0, n=None, want_joint=True, want_compact=False, threshold=0.5, out=None):
        st = states
        if not (isinstance(st, np.ndarray) and st.dtype == np.uint8 and st.ndim == 2 and st.strides[1] == 1):
            st = np.ascontiguousarray(states, dtype=np.uint8)
        if n is None:
            n = st.shape[1] - site0
        out = out or {}
        lnl = out.get("lnl") if out.get("lnl") is not None else np.empty(n, dtype=np.float64)
        joint = (out.get("joint") if out.get("joint") is not None else np.empty((n, 20, 20), dtype=np.float64)) if want_joint else None
        comp = (out.get("compact") if out.get("compact") is not None else np.zeros(n, dtype=compact_dtype(self.compact_cap))) if want_compact else None
        self._ck(lib().sr_run_streamed(self._h, st.shape[0], st.shape[1], st.ctypes.data, st.strides[0], int(site0), int(n),
                                       lnl.ctypes.data, joint.ctypes.data if want_joint else None,
                                       comp.ctypes.data if want_compact else None, float(threshold)))
        return {"lnl": lnl, "joint": joint, "compact": comp}

    def run_device(self, site0, n, d_lnl, d_joint=None, d_compact=None, threshold=0.5, stream=None):
        """Device pointers (ints), enqueued on `stream` (int cudaStream_t) without synchronising."""
        self._ck(lib().sr_run_device(self._h, int(site0), int(n), d_lnl, d_joint, d_compact, float(threshold), stream))

    def profile(self, reset=False):
        p = Profile()
        self._ck(lib().sr_get_profile(self._h, ctypes.byref(p)))
        d = {f[0]: getattr(p, f[0]) for f in Profile._fields_}
        if reset:
            self._ck(lib().sr_profile_reset(self._h))
        return d

    def schedule_info(self):
        n = ctypes.c_int64()
        d = ctypes.c_int32()
        f = ctypes.c_double()
        self._ck(lib().sr_schedule_info(self._h, ctypes.byref(n), ctypes.byref(d), ctypes.byref(f)))
        return {"n_steps": n.value, "stack_depth": d.value, "flops_per_column_rate": f.value}

    def measure_fp64_peak(self, ms=200.0):
        t = ctypes.c_double()
        self._ck(lib().sr_measure_fp64_peak(self._h, float(ms), ctypes.byref(t)))
        return t.value

    def debug_pmatrix(self, node, rate_class):
        out = np.zeros((20, 20))
        self._ck(lib().sr_debug_pmatrix(self._h, int(node), int(rate_class), out.ctypes.data_as(_D)))
        return out

    def debug_cherry_table(self, node, rate_class):
        """The vector collapsed clade `node` hands to its parent per state combination of its tips:
        [21, 21, 20] for a cherry, [21, 21, 21, 20] for a pitchfork (cherry tips first, then the third tip)."""
        out = np.zeros(21 * 21 * 21 * 20)
        n = ctypes.c_int32()
        self._ck(lib().sr_debug_cherry_table(self._h, int(node), int(rate_class), out.ctypes.data_as(_D), out.size, ctypes.byref(n)))
        return out[:n.value * 20].reshape((21, 21, 20) if n.value == 441 else (21, 21, 21, 20))


class MultiContext:
    """One host process driving every GPU of the box (include/subrecon_b200.h::sr_multi): the single-JVM route."""

    def __init__(self, devices=None):
        self._h = ctypes.c_void_p()
        if devices is None:
            rc = lib().sr_multi_create(ctypes.byref(self._h), 0, None)
        else:
            ids = np.ascontiguousarray(list(devices), dtype=np.int32)
            rc = lib().sr_multi_create(ctypes.byref(self._h), int(ids.size), ids.ctypes.data_as(_I32))
        if rc:
            raise SrError(rc, (lib().sr_multi_last_error(None) or b"").decode())
        self.compact_cap = 8

    @property
    def count(self):
        return int(lib().sr_multi_count(self._h))

    def close(self):
        if self._h:
            lib().sr_multi_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise SrError(rc, (lib().sr_multi_last_error(self._h) or b"").decode())

    def set_model(self, Eval, Evec, Ievc, pi):
        k = [Context._d(x) for x in (Eval, Evec, Ievc, pi)]
        self._ck(lib().sr_multi_set_model(self._h, k[0][1], k[1][1], k[2][1], k[3][1]))

    def set_rates(self, rates):
        r, rp = Context._d(rates)
        self._ck(lib().sr_multi_set_rates(self._h, int(r.size), rp))

    def set_tree(self, ft):
        off = np.ascontiguousarray(ft.child_off, dtype=np.int32)
        idx = np.ascontiguousarray(ft.child_idx, dtype=np.int32)
        bl = np.ascontiguousarray(ft.blen, dtype=np.float64)
        lr = np.ascontiguousarray(ft.leaf_row, dtype=np.int32)
        self._ck(lib().sr_multi_set_tree(self._h, int(ft.n_nodes), off.ctypes.data_as(_I32), idx.ctypes.data_as(_I32),
                                         bl.ctypes.data_as(_D), lr.ctypes.data_as(_I32), int(ft.root)))

    def set_option(self, name, value):
        self._ck(lib().sr_multi_set_option(self._h, name.encode(), int(value)))
        if name == "compact_cap":
            self.compact_cap = int(value)

    def run_streamed(self, states, site0=0, n=None, want_joint=True, want_compact=False, threshold=0.5, out=None):
        st = states
        if not (isinstance(st, np.ndarray) and st.dtype == np.uint8 and st.ndim == 2 and st.strides[1] == 1):
            st = np.ascontiguousarray(states, dtype=np.uint8)
        if n is None:
            n = st.shape[1] - site0
        out = out or {}
        lnl = out.get("lnl") if out.get("lnl") is not None else np.empty(n, dtype=np.float64)
        joint = (out.get("joint") if out.get("joint") is not None else np.empty((n, 20, 20), dtype=np.float64)) if want_joint else None
        comp = (out.get("compact") if out.get("compact") is not None else np.zeros(n, dtype=compact_dtype(self.compact_cap))) if want_compact else None
        self._ck(lib().sr_multi_run_streamed(self._h, st.shape[0], st.shape[1], st.ctypes.data, st.strides[0], int(site0), int(n),
                                             lnl.ctypes.data, joint.ctypes.data if want_joint else None,
                                             comp.ctypes.data if want_compact else None, float(threshold)))
        return {"lnl": lnl, "joint": joint, "compact": comp}
