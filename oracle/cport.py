"""TEST INFRASTRUCTURE ONLY. ctypes view of oracle/libppm_oracle.so (oracle/ppm_oracle.c)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libppm_oracle.so")


class _Model(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("n_leaves", C.c_int), ("n_rep", C.c_int), ("n_obs", C.c_int),
                ("left", C.c_void_p), ("right", C.c_void_p), ("bl", C.c_void_p), ("height", C.c_void_p),
                ("order", C.c_void_p), ("sub_off", C.c_void_p), ("sub_idx", C.c_void_p), ("obs", C.c_void_p),
                ("counts", C.c_void_p), ("mrca", C.c_void_p), ("freq", C.c_void_p), ("H", C.c_double)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libppm_oracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "ppm_oracle.c")):
            build()
        L = C.CDLL(LIB_PATH)
        L.ppm_oracle_eval.restype = C.c_double
        L.ppm_oracle_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
        L.ppm_oracle_set_threads.argtypes = [C.c_int]
        L.ppm_oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


class Oracle:
    """Evaluates the reference's logLik / computeI2 / likRepTot components / assignCat categories on a
    flat model dict from oracle/model.py."""

    def __init__(self, model):
        self.m = model
        self._keep = {}
        s = _Model()
        for k in ("n_nodes", "n_leaves", "n_rep", "n_obs"):
            setattr(s, k, int(model[k]))
        for k, dt in (("left", np.int32), ("right", np.int32), ("bl", np.float64), ("height", np.float64),
                      ("order", np.int32), ("sub_off", np.int32), ("sub_idx", np.int32), ("obs", np.int8),
                      ("counts", np.int32), ("mrca", np.int32), ("freq", np.int32)):
            a = np.ascontiguousarray(model[k], dt)
            self._keep[k] = a
            setattr(s, k, a.ctypes.data)
        s.H = float(model["H"])
        self.s = s

    def eval(self, params, components=False, cats=False, first_rep=0, n_eval=None):
        p = np.ascontiguousarray(params, np.float64)
        nrep = self.m["n_rep"]
        n_eval = nrep - first_rep if n_eval is None else n_eval
        i2 = C.c_double()
        parts = np.zeros(4)
        comp = np.full((nrep, 4), np.nan) if components else None
        cat = np.full(nrep, -1, np.int32) if cats else None
        ll = lib().ppm_oracle_eval(C.addressof(self.s), p.ctypes.data, first_rep, n_eval, C.addressof(i2),
                                   parts.ctypes.data, comp.ctypes.data if components else None,
                                   cat.ctypes.data if cats else None)
        return dict(ll=float(ll), i2=float(i2.value), parts=parts, comp=comp, cat=cat)

    def loglik(self, params):
        return self.eval(params)["ll"]


def set_threads(n):
    lib().ppm_oracle_set_threads(int(n))


def max_threads():
    return int(lib().ppm_oracle_max_threads())
