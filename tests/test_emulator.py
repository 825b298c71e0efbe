"""CPU tier: the kernel bodies (ppm_device_fns.cuh) compiled by g++ as a single-lane emulation
(tests/emu/emu_sweep.cpp, test infrastructure) reproduce the reference's golden outputs.  This checks the
linear-domain arithmetic, the prepass tables and the sweep program on the CPU; the CUDA instantiation of the
same source is checked on the GPU by tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel

EMU_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libppm_emu.so"))
    L.emu_error.restype = C.c_char_p
    L.emu_eval.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6

    def run(tree, pa, dedupe, params, R=8, want_comp=True, n_rows=0):
        p = np.ascontiguousarray(params, np.float64).reshape(-1, 8)
        P = len(p)
        ll, i2, parts = np.zeros(P), np.zeros(P), np.zeros((P, 4))
        npat = C.c_longlong()
        comp = np.zeros((max(n_rows, 1), 3)); cat = np.zeros(max(n_rows, 1), np.int8)
        rc = L.emu_eval(tree.encode(), pa.encode(), int(dedupe), R, p.ctypes.data, P, ll.ctypes.data, i2.ctypes.data,
                        parts.ctypes.data, comp.ctypes.data if want_comp else None, cat.ctypes.data if want_comp else None,
                        C.byref(npat))
        if rc:
            raise RuntimeError(L.emu_error().decode())
        return ll, i2, parts, comp[:npat.value], cat[:npat.value]

    run.lib = L
    return run


@pytest.mark.parametrize("name", ["ref_test", "syn12", "syn100", "cat40", "syn300"])
# < 0: the deep-prefetch path of the global-memory variant (default there: 16); |R| > 100: the leaf-power program (the alive
# leaves of a slice as one exp2 of count tables, first-order in the leaf heights: the golden trees carry ape's 10-digit lengths)
@pytest.mark.parametrize("R", [8, 4, 12, 16, -8, -16, 108, 104, 116, -108, -116])
def test_kernel_bodies_reproduce_reference_golden(golden, emu, name, R):
    g = golden(name)
    try:
        ll, i2, parts, comp, cat = emu(g.tree, g.pa, False, g.params, R=R, n_rows=g.meta["nRep"])
    except RuntimeError as e:
        if abs(R) > 100 and "refused" in str(e):
            pytest.skip("leaf powers refused for this tree: " + str(e))
        raise
    for k, r in enumerate(g.results):
        assert rel(ll[k], r["ll"]) <= 1e-12, (name, k, ll[k], r["ll"])
        assert rel(i2[k], r["i2"]) <= 1e-12
        assert np.allclose(parts[k], r["parts"], rtol=1e-12, atol=1e-13)  # p1 = 1-exp(.) cancels: absolute bound
    ref3 = g.ref3(0)
    fin = np.isfinite(ref3)
    assert np.allclose(comp[fin], ref3[fin], rtol=1e-10, atol=1e-10)
    ok = np.isfinite(g.z["components"][0][:, 3])   # reference lik2 = -inf where its exp(f_aux) underflows
    assert np.array_equal(cat[ok], g.z["cats"][0][ok])


def test_deduplicated_input_gives_the_same_loglik(golden, emu):
    g = golden("ref_test")
    a = emu(g.tree, g.pa, False, g.params[:2], want_comp=False)[0]
    b = emu(g.tree, g.pa, True, g.params[:2], want_comp=False)[0]
    assert np.allclose(a, b, rtol=1e-13, atol=0)


@pytest.mark.parametrize("name", ["ref_test", "syn12", "syn100", "cat40"])
@pytest.mark.parametrize("R", [8, -8, 108, -108])
def test_posterior_outputs_reproduce_reference(golden, emu, name, R):
    """expectedTime1 / expectedTime2_aux / expectedGains (functions.cpp:582-757) from the kernel bodies against
    tests/golden/posterior_*.npz (written by the unmodified reference, oracle/gen_golden_posterior.py)."""
    g = golden(name)
    z = np.load(os.path.join(os.path.dirname(g.tree), "posterior_" + name + ".npz"))
    n, S = g.meta["nRep"], len(z["times"])
    et1, dist, gains = np.zeros(n), np.zeros((n, S + 1)), C.c_double()
    for g2, want in zip(z["gains_g2"], z["gains"]):
        emu.lib.emu_set_posterior_outputs.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        emu.lib.emu_set_posterior_outputs(et1.ctypes.data, dist.ctypes.data, float(g2), C.byref(gains))
        try:
            emu(g.tree, g.pa, False, z["params"], R=R, want_comp=False)
        except RuntimeError as e:
            if abs(R) > 100 and "refused" in str(e):
                pytest.skip("leaf powers refused for this tree: " + str(e))
            raise
        assert rel(gains.value, want) <= 1e-10, (g2, gains.value, want)
    assert np.allclose(et1, z["et1"], rtol=1e-9, atol=0)
    assert np.allclose(dist, z["dist"], rtol=1e-9, atol=1e-300)
