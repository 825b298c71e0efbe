"""TEST INFRASTRUCTURE ONLY. ctypes view of oracle/_ref/libppm_ref.so (the unmodified reference,
built by oracle/Makefile from /root/reference/source). Importable only by tests/, bench.py's
reference arm and oracle/gen_golden.py."""
import ctypes as C
import os
import tempfile
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libppm_ref.so")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_dims.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_get_int.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_get_double.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_get_pattern.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_subtrees.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_subtrees.restype = C.c_int
        L.ref_loglik.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_loglik.restype = C.c_double
        L.ref_compute_i2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_compute_i2.restype = C.c_double
        L.ref_components.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_assign_cat.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_char_p]
        L.ref_assign_cat.restype = C.c_int
        L.ref_proba_cat.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_char_p]
        L.ref_proba_cat.restype = C.c_int
        L.ref_expected_time1.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_expected_time2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_expected_time2.restype = C.c_int
        L.ref_expected_gains.argtypes = [C.c_void_p, C.c_double]
        L.ref_expected_gains.restype = C.c_double
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_max_threads.restype = C.c_int
        _lib = L
    return _lib


class Reference:
    """The reference's Inference object (functions.cpp:76-156) on a tree file + matrix file."""

    def __init__(self, tree_path, pa_path, compress=False):
        self.h = lib().ref_create(tree_path.encode(), pa_path.encode(), int(compress))
        if not self.h:
            raise RuntimeError("reference could not read " + tree_path)
        di = np.zeros(4, np.int32)
        dd = np.zeros(3, np.float64)
        lib().ref_dims(self.h, di.ctypes.data, dd.ctypes.data)
        self.nRep, self.nObs, self.nNodes, self.nLeaves = (int(x) for x in di)
        self.H, self.L, self.meanGenes = (float(x) for x in dd)

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def _int(self, which, n):
        a = np.zeros(n, np.int32)
        lib().ref_get_int(self.h, which, a.ctypes.data)
        return a

    def _dbl(self, which, n):
        a = np.zeros(n, np.float64)
        lib().ref_get_double(self.h, which, a.ctypes.data)
        return a

    mrca = property(lambda s: s._int(0, s.nRep))
    freq = property(lambda s: s._int(1, s.nRep))
    order = property(lambda s: s._int(2, s.nNodes))
    counts = property(lambda s: s._int(3, s.nRep))
    left = property(lambda s: s._int(4, s.nNodes))
    right = property(lambda s: s._int(5, s.nNodes))
    height = property(lambda s: s._dbl(0, s.nNodes))
    branch_length = property(lambda s: s._dbl(1, s.nNodes))

    def pattern(self, rep):
        a = np.zeros(self.nNodes, np.int32)
        lib().ref_get_pattern(self.h, rep, a.ctypes.data)
        return a

    def subtrees(self, rank):
        a = np.zeros(self.nNodes, np.int32)
        n = lib().ref_subtrees(self.h, rank, a.ctypes.data)
        return a[:n].copy()

    def loglik(self, params):
        p = np.ascontiguousarray(params, np.float64)
        return float(lib().ref_loglik(self.h, p.ctypes.data))

    def compute_i2(self, params, want_parts=False):
        p = np.ascontiguousarray(params, np.float64)
        parts = np.zeros(4, np.float64)
        i2 = float(lib().ref_compute_i2(self.h, p.ctypes.data, parts.ctypes.data if want_parts else None))
        return (i2, parts) if want_parts else i2

    def components(self, params):
        p = np.ascontiguousarray(params, np.float64)
        out = np.zeros((self.nRep, 4), np.float64)
        lib().ref_components(self.h, p.ctypes.data, out.ctypes.data)
        return out

    def assign_cat(self, params, i2):
        p = np.ascontiguousarray(params, np.float64)
        out = np.zeros(self.nRep, np.int32)
        fd, tmp = tempfile.mkstemp(prefix="ppm_cat_")
        os.close(fd)
        n = lib().ref_assign_cat(self.h, p.ctypes.data, float(i2), out.ctypes.data, tmp.encode())
        assert n == self.nRep
        return out


def _proba_cat(self, params, i2):
    p = np.ascontiguousarray(params, np.float64)
    out = np.zeros((self.nRep, 3), np.float64)
    fd, tmp = tempfile.mkstemp(prefix="ppm_pc_")
    os.close(fd)
    n = lib().ref_proba_cat(self.h, p.ctypes.data, float(i2), out.ctypes.data, tmp.encode())
    assert n == self.nRep
    return out


Reference.proba_cat = _proba_cat


def set_threads(n):
    lib().ref_set_threads(int(n))


def max_threads():
    return int(lib().ref_max_threads())


def _expected_time1(self, params):
    """expectedTime1 (functions.cpp:637-690) per pattern."""
    p = np.ascontiguousarray(params, np.float64)
    out = np.zeros(self.nRep, np.float64)
    lib().ref_expected_time1(self.h, p.ctypes.data, out.ctypes.data)
    return out


def _expected_time2(self, params):
    """(times[S], dist[nRep][S+1]): expectedTime2_times, expectedTime2_aux + exp(likRep2) (functions.cpp:709-757)."""
    p = np.ascontiguousarray(params, np.float64)
    S = lib().ref_expected_time2(self.h, p.ctypes.data, None, None)
    times = np.zeros(S, np.float64)
    dist = np.zeros((self.nRep, S + 1), np.float64)
    lib().ref_expected_time2(self.h, p.ctypes.data, times.ctypes.data, dist.ctypes.data)
    return times, dist


def _expected_gains(self, g2):
    return float(lib().ref_expected_gains(self.h, float(g2)))


Reference.expected_time1 = _expected_time1
Reference.expected_time2 = _expected_time2
Reference.expected_gains = _expected_gains
