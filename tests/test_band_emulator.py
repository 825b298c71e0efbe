"""CPU tier: the band program (csrc/band_program.cpp: tiles of 32*R slices, FULL / per-register / masked record lists, one
flat descriptor list per warp) and the band prepass tables, executed by the scalar interpreter in tests/emu/emu_sweep.cpp,
reproduce the reference's golden outputs and the C oracle on a seeded 600-genome tree.  The CUDA band kernel runs the same
program with lanes as slices (tests/test_gpu_band.py)."""
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

    def run(tree, pa, dedupe, params, code, n_rows=0):
        p = np.ascontiguousarray(params, np.float64).reshape(-1, 8)
        P = len(p)
        ll, i2, parts = np.zeros(P), np.zeros(P), np.zeros((P, 4))
        npat = C.c_longlong()
        comp = np.zeros((max(n_rows, 1), 3))
        rc = L.emu_eval(tree.encode(), pa.encode(), int(dedupe), code, p.ctypes.data, P, ll.ctypes.data, i2.ctypes.data,
                        parts.ctypes.data, comp.ctypes.data if n_rows else None, None, C.byref(npat))
        if rc:
            raise RuntimeError(L.emu_error().decode())
        return ll, i2, comp[:npat.value]

    return run


@pytest.mark.parametrize("name", ["syn12", "syn100", "cat40", "syn300"])
# 1000 + 100 * (warps per item) + (slices per lane); 3000 + ...: stream layout (band_stream.cu), work units per item
# 5000 + ...: stream layout with leaf powers (the alive leaves of a slice as one exp2; the golden trees carry ape's 10-digit
# branch lengths, so the first-order height correction is exercised)
@pytest.mark.parametrize("code", [1104, 1102, 1108, 1204, 1404, 1802, 2604, 3104, 3204, 3302, 3404, 5104, 5204, 5302, 5404])
def test_band_program_reproduces_reference_golden(golden, emu, name, code):
    g = golden(name)
    ll, i2, comp = emu(g.tree, g.pa, False, g.params, code, n_rows=g.meta["nRep"])
    for k, r in enumerate(g.results):
        assert rel(ll[k], r["ll"]) <= 1e-12, (name, k, ll[k], r["ll"])
        assert rel(i2[k], r["i2"]) <= 1e-12
    ref3 = g.ref3(0)
    fin = np.isfinite(ref3)
    assert np.allclose(comp[fin], ref3[fin], rtol=1e-10, atol=1e-10)


def test_band_program_refuses_what_it_cannot_serve(golden, emu):
    """The shipped 20-genome data at the Appendix-C vectors: (g2+l2) H / 2 >= 600 or pi1 > 0.95 for some of them -- such
    parameter vectors are not eligible (SC_BAND = 0; on the device they take the lane-per-pattern sweep)."""
    g = golden("ref_test")
    with pytest.raises(RuntimeError, match="not eligible"):
        emu(g.tree, g.pa, False, g.params, 1104)


def test_band_program_against_oracle_600_genomes(emu, tmp_path):
    from oracle import cport, model
    from ppmmodelpangenome_b200 import simulate
    rng = np.random.default_rng(600)
    nwk = simulate.sim_tree(600, rng)
    names, M, truth = simulate.sim_genes(nwk, 24, rng)
    tp, pp = str(tmp_path / "t.nwk"), str(tmp_path / "m.txt")
    open(tp, "w").write(nwk + "\n")
    simulate.write_matrix(pp, names, M)
    orc = cport.Oracle(model.model_from_files(tp, pp))
    P = np.stack([simulate.start_params(truth, 24), simulate.start_params(truth, 24, rng)])
    for code in (1104, 1404, 3304):
        ll, i2, _ = emu(tp, pp, True, P, code)
        for k in range(2):
            o = orc.eval(P[k], cats=False)
            assert rel(ll[k], o["ll"]) <= 1e-11, (code, k, ll[k], o["ll"])
            assert rel(i2[k], o["i2"]) <= 1e-11
