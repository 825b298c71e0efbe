"""Python view of the C-ABI engine (include/ppm_b200.h), mirroring the reference's `Inference` class
(reference source/functions.hpp:17-104): same method names, argument order and error behaviour for the
likelihood path.  ctypes only -- no torch types cross the boundary.  There is no CPU fallback: if the
CUDA library is missing or no sm_100 device is present, construction raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


class PPMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ppm_b200 error %d: %s" % (code, msg))
        self.code = code


class _Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("left", C.c_void_p), ("right", C.c_void_p), ("branch_len", C.c_void_p)]


class Info(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_leaves", C.c_int32), ("n_genes", C.c_int64), ("n_patterns", C.c_int64),
                ("n_obs", C.c_int64), ("n_words", C.c_int32), ("n_slices", C.c_int32), ("n_terms", C.c_int64),
                ("H", C.c_double), ("L", C.c_double), ("mean_genes", C.c_double), ("device", C.c_int32),
                ("shard_rank", C.c_int32), ("shard_world", C.c_int32), ("shard_first", C.c_int64),
                ("shard_count", C.c_int64), ("group_size", C.c_int32)]


SYMBOLS = ["ppm_create", "ppm_create_from_files", "ppm_create_group", "ppm_create_group_from_files", "ppm_destroy", "ppm_last_error", "ppm_get_info",
           "ppm_get_int_array", "ppm_get_double_array", "ppm_loglik", "ppm_loglik_device", "ppm_finalize_ll",
           "ppm_finalize_ll_device", "ppm_compute_i2", "ppm_assign_cat", "ppm_components", "ppm_proba_cat", "ppm_expected_time1",
           "ppm_expected_time2", "ppm_expected_gains", "ppm_get_counters", "ppm_get_plan", "ppm_simulate_genes",
           "ppm_measure_fp64_peak", "ppm_write_ccc", "ppm_set_sparse"]


def lib():
    """Loads ppmmodelpangenome_b200/libppm_b200.so (building it with nvcc if the sources are newer)."""
    global _LIB
    if _LIB is None:
        path = _build.LIB
        try:
            path = _build.build_lib()
        except Exception as e:
            # a stale library would silently run old kernels: only a box without nvcc may use the prebuilt one
            if not os.path.exists(path) or os.path.exists(_build.NVCC):
                raise
            import warnings
            warnings.warn("ppm_b200: could not rebuild %s (%r); loading the prebuilt library" % (path, e))
        L = C.CDLL(path)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.ppm_create.argtypes = [vp, vp, vp, vp, i64, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ppm_create_from_files.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ppm_create_group.argtypes = [vp, vp, vp, vp, i64, C.c_int, vp, i32]
        L.ppm_create_group_from_files.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, vp, i32]
        L.ppm_destroy.argtypes = [vp]
        L.ppm_destroy.restype = None
        L.ppm_last_error.argtypes = [vp]
        L.ppm_last_error.restype = C.c_char_p
        L.ppm_get_info.argtypes = [vp, vp]
        L.ppm_write_ccc.argtypes = [vp, C.c_char_p]
        L.ppm_set_sparse.argtypes = [vp, C.c_int]
        L.ppm_get_int_array.argtypes = [vp, C.c_int, vp, i64]
        L.ppm_get_double_array.argtypes = [vp, C.c_int, vp, i64]
        L.ppm_loglik.argtypes = [vp, vp, i32, vp, vp]
        L.ppm_loglik_device.argtypes = [vp, vp, i32, vp, vp, vp]
        L.ppm_finalize_ll.argtypes = [vp, vp, i32]
        L.ppm_finalize_ll_device.argtypes = [vp, vp, vp, i32, vp]
        L.ppm_compute_i2.argtypes = [vp, vp, i32, vp, vp]
        L.ppm_assign_cat.argtypes = [vp, vp, dbl, vp]
        L.ppm_components.argtypes = [vp, vp, dbl, vp]
        L.ppm_proba_cat.argtypes = [vp, vp, dbl, vp]
        L.ppm_expected_time1.argtypes = [vp, vp, vp]
        L.ppm_expected_time2.argtypes = [vp, vp, vp, vp]
        L.ppm_expected_gains.argtypes = [vp, dbl, vp]
        L.ppm_get_counters.argtypes = [vp, vp]
        L.ppm_get_plan.argtypes = [vp, vp]
        L.ppm_simulate_genes.argtypes = [vp, i64, C.c_uint64, vp, vp, dbl, dbl, i32, vp, vp]
        L.ppm_measure_fp64_peak.argtypes = [C.c_int, vp]
        _LIB = L
    return _LIB


def measure_fp64_peak(device=0):
    v = C.c_double()
    rc = lib().ppm_measure_fp64_peak(device, C.byref(v))
    if rc:
        raise PPMError(rc, "DFMA micro-benchmark failed")
    return v.value


def simulate_genes(left, right, branch_len, n_genes, seed, frac=(0.25, 0.35, 0.40), rates_L=(1.0, 20.0, 50.0, 2500.0),
                   s_10=0.01, s_01=0.01, threads=0):
    """C++ PPM gene simulator (ppm_simulate_genes): bit-packed rows [n_genes][ceil(n/32)] for ppm_create + the truth dict.
    Host code; used by bench.py for the 1M-gene workloads (the numpy simulator in simulate.py serves the small ones)."""
    left = np.ascontiguousarray(left, np.int32); right = np.ascontiguousarray(right, np.int32)
    bl = np.ascontiguousarray(branch_len, np.float64)
    t = _Tree(len(left), left.ctypes.data, right.ctypes.data, bl.ctypes.data)
    nw = ((len(left) + 1) // 2 + 31) // 32
    rows = np.zeros((int(n_genes), nw), np.uint32)
    fr = np.ascontiguousarray(frac, np.float64); rt = np.ascontiguousarray(rates_L, np.float64)
    truth = np.zeros(10)
    rc = lib().ppm_simulate_genes(C.byref(t), int(n_genes), int(seed), fr.ctypes.data, rt.ctypes.data, float(s_10), float(s_01),
                                  int(threads), rows.ctypes.data, truth.ctypes.data)
    if rc:
        raise PPMError(rc, "ppm_simulate_genes failed")
    keys = ("N0", "l0", "i1", "l1", "g2", "l2", "s_10", "s_01", "H", "L")
    return rows, dict(zip(keys, truth.tolist()))


def _params(p, P=None):
    a = np.ascontiguousarray(p, np.float64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.shape[1] != 8:
        raise ValueError("parameter vectors are [N0 l0 i1 l1 g2 l2 s_10 s_01]")
    return a


class Inference:
    """Drop-in for the likelihood side of the reference's Inference object.

    Inference(tree_path, pa_path)                 reads the reference's input files
    Inference.from_arrays(left, right, bl, rows)  flat preorder tree + bit-packed patterns
    """

    def __init__(self, tree_path=None, pa_path=None, dedupe=True, device=0, shard_rank=0, shard_world=1, _handle=None,
                 devices=None):
        """devices=[ids]: a GROUP handle, one pattern shard per listed GPU behind one object (ppm_create_group*)."""
        self._h = C.c_void_p()
        L = lib()
        if _handle is not None:
            self._h = _handle
        elif devices is not None:
            ids = np.ascontiguousarray(devices, np.int32)
            rc = L.ppm_create_group_from_files(C.byref(self._h), os.fsencode(tree_path), os.fsencode(pa_path), int(dedupe),
                                               ids.ctypes.data, len(ids))
            if rc:
                raise PPMError(rc, L.ppm_last_error(None).decode())
        else:
            rc = L.ppm_create_from_files(C.byref(self._h), os.fsencode(tree_path), os.fsencode(pa_path), int(dedupe),
                                         int(device), int(shard_rank), int(shard_world))
            if rc:
                raise PPMError(rc, L.ppm_last_error(None).decode())
        self.info = Info()
        L.ppm_get_info(self._h, C.byref(self.info))
        i = self.info
        # the reference's field names (functions.hpp:78-94)
        self.nRep, self.nObs, self.nNodes, self.nLeaves = i.n_patterns, i.n_obs, i.n_nodes, i.n_leaves
        self.nGenes = i.n_genes
        self.H, self.L, self.meanGenes = i.H, i.L, i.mean_genes

    @classmethod
    def from_arrays(cls, left, right, branch_len, packed_rows, counts=None, dedupe=True, device=0, shard_rank=0,
                    shard_world=1, devices=None):
        left = np.ascontiguousarray(left, np.int32)
        right = np.ascontiguousarray(right, np.int32)
        bl = np.ascontiguousarray(branch_len, np.float64)
        rows = np.ascontiguousarray(packed_rows, np.uint32)
        t = _Tree(len(left), left.ctypes.data, right.ctypes.data, bl.ctypes.data)
        cnt = None if counts is None else np.ascontiguousarray(counts, np.int32)
        h = C.c_void_p()
        L = lib()
        if devices is not None:
            ids = np.ascontiguousarray(devices, np.int32)
            rc = L.ppm_create_group(C.byref(h), C.byref(t), rows.ctypes.data, None if cnt is None else cnt.ctypes.data,
                                    rows.shape[0], int(dedupe), ids.ctypes.data, len(ids))
        else:
            rc = L.ppm_create(C.byref(h), C.byref(t), rows.ctypes.data, None if cnt is None else cnt.ctypes.data,
                              rows.shape[0], int(dedupe), int(device), int(shard_rank), int(shard_world))
        if rc:
            raise PPMError(rc, L.ppm_last_error(None).decode())
        return cls(_handle=h)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if self._h:
            lib().ppm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _n_out(self):
        """patterns covered by the per-pattern outputs: this shard's slice, or everything for a group handle"""
        return self.info.n_patterns if self.info.group_size > 1 else self.info.shard_count

    def _ck(self, rc):
        if rc:
            raise PPMError(rc, lib().ppm_last_error(self._h).decode())

    def _ints(self, which, n):
        a = np.zeros(max(int(n), 1), np.int32)
        self._ck(lib().ppm_get_int_array(self._h, which, a.ctypes.data, a.size))
        return a[:n]

    def _dbls(self, which, n):
        a = np.zeros(max(int(n), 1), np.float64)
        self._ck(lib().ppm_get_double_array(self._h, which, a.ctypes.data, a.size))
        return a[:n]

    mrca = property(lambda s: s._ints(0, s.nRep))
    freq = property(lambda s: s._ints(1, s.nRep))
    counts = property(lambda s: s._ints(2, s.nRep))
    order = property(lambda s: s._ints(3, s.nNodes))
    gene_to_pattern = property(lambda s: s._ints(4, s.nGenes))
    nsub = property(lambda s: s._ints(5, s.nNodes - 1))
    alive = property(lambda s: s._ints(6, s.nNodes - 1))
    height = property(lambda s: s._dbls(0, s.nNodes))
    branchLength = property(lambda s: s._dbls(1, s.nNodes))

    # ------------------------------------------------------------------ the objective
    def logLik(self, N0, l0=None, i1=None, l1=None, g2=None, l2=None, s=None):
        """logLik(N0,l0,i1,l1,g2,l2,{s,s,s,s_01}) of functions.cpp:359 -- or logLik(params[8]) / logLik(params[P][8])."""
        if l0 is not None:
            return float(self.loglik_batch([[N0, l0, i1, l1, g2, l2, s[0], s[3]]])[0])
        p = _params(N0)
        out = self.loglik_batch(p)
        return float(out[0]) if np.ndim(N0) == 1 else out

    def loglik_batch(self, params, return_i2=False):
        p = _params(params)
        ll = np.zeros(p.shape[0])
        i2 = np.zeros(p.shape[0])
        self._ck(lib().ppm_loglik(self._h, p.ctypes.data, p.shape[0], ll.ctypes.data, i2.ctypes.data))
        return (ll, i2) if return_i2 else ll

    def computeI2(self, params, return_parts=False):
        p = _params(params)
        i2 = np.zeros(p.shape[0])
        parts = np.zeros((p.shape[0], 4))
        self._ck(lib().ppm_compute_i2(self._h, p.ctypes.data, p.shape[0], i2.ctypes.data, parts.ctypes.data))
        if np.ndim(params) == 1:
            return (float(i2[0]), parts[0]) if return_parts else float(i2[0])
        return (i2, parts) if return_parts else i2

    def assignCat(self, params8, i2=None):
        """Per-gene argmax categories (functions.cpp:489-520), in input-gene order."""
        p = _params(params8)
        if i2 is None:
            i2 = self.computeI2(p[0])
        out = np.zeros(max(self.nGenes, 1), np.int8)
        self._ck(lib().ppm_assign_cat(self._h, p.ctypes.data, float(i2), out.ctypes.data))
        return out[:self.nGenes]

    def components(self, params8, i2=None):
        p = _params(params8)
        if i2 is None:
            i2 = self.computeI2(p[0])
        n = self._n_out()
        out = np.zeros((max(n, 1), 3))
        self._ck(lib().ppm_components(self._h, p.ctypes.data, float(i2), out.ctypes.data))
        return out[:n]

    def computeProbaCat(self, params8, i2=None):
        """Log posterior category probabilities per input gene (functions.cpp:556-581), shape [nGenes, 3]."""
        p = _params(params8)
        if i2 is None:
            i2 = self.computeI2(p[0])
        out = np.zeros((max(self.nGenes, 1), 3))
        self._ck(lib().ppm_proba_cat(self._h, p.ctypes.data, float(i2), out.ctypes.data))
        return out[:self.nGenes]

    def expectedTime1(self, params8):
        """Expected introduction time of every pattern of this shard taken as a Private gene (functions.cpp:637-708)."""
        p = _params(params8)
        n = self._n_out()
        out = np.zeros(max(n, 1))
        self._ck(lib().ppm_expected_time1(self._h, p.ctypes.data, out.ctypes.data))
        return out[:n]

    def expectedTime2(self, params8):
        """(times[S], dist[nPatterns, S+1]): slice mid-points (expectedTime2_times), the Mobile integrand per slice
        (expectedTime2_aux) and, last column, exp(likRep2) (functions.cpp:709-757)."""
        p = _params(params8)
        n, S = self._n_out(), self.info.n_slices
        times = np.zeros(max(S, 1))
        dist = np.zeros((max(n, 1), S + 1))
        self._ck(lib().ppm_expected_time2(self._h, p.ctypes.data, times.ctypes.data, dist.ctypes.data))
        return times[:S], dist[:n]

    def expectedGains(self, g2):
        """Expected number of Mobile gains along the tree (functions.cpp:593-615); host model, no GPU needed."""
        v = C.c_double()
        self._ck(lib().ppm_expected_gains(self._h, float(g2), C.byref(v)))
        return v.value

    # device-pointer entry points (bench / multi-GPU plumbing)
    STREAM_OWN = C.c_void_p(-1)  # PPM_STREAM_OWN: the handle's private stream; None / 0 is the legacy default stream

    def loglik_device(self, d_params_ptr, P, d_out_ptr, d_i2_ptr, stream_ptr=None):
        self._ck(lib().ppm_loglik_device(self._h, d_params_ptr, P, d_out_ptr, d_i2_ptr, stream_ptr))

    def finalize_device(self, d_i2_ptr, d_ll_ptr, P, stream_ptr=None):
        self._ck(lib().ppm_finalize_ll_device(self._h, d_i2_ptr, d_ll_ptr, P, stream_ptr))

    def set_sparse(self, on=True):
        """Sparse evaluation of the Mobile integral on band-kernel trees (ppm_set_sparse)."""
        self._ck(lib().ppm_set_sparse(self._h, int(bool(on))))

    def write_ccc(self, path):
        """Unique patterns + counts in the reference's compressed matrix format (ppm_write_ccc)."""
        self._ck(lib().ppm_write_ccc(self._h, str(path).encode()))

    def plan(self):
        """Which kernels serve this handle (ppm_get_plan)."""
        a = np.zeros(16, np.int32)
        self._ck(lib().ppm_get_plan(self._h, a.ctypes.data))
        keys = ("band", "band_R", "band_G", "last_kind", "last_mode", "rows_in_place", "deep", "lean", "levels", "band_tiles",
                "blocks", "slots", "leaf_u", "band_stream", "sweep_lp", "band_lp")
        return dict(zip(keys, (int(x) for x in a)))

    def counters(self):
        a = np.zeros(4)
        self._ck(lib().ppm_get_counters(self._h, a.ctypes.data))
        return dict(launches=int(a[0]), sweep_ms=float(a[1]), flops_per_eval=float(a[2]), fp64_instr_per_eval=float(a[3]))
