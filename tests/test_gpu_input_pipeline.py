"""GPU tier of the input pipeline: the device dedupe (csrc/pattern_kernels.cu: hash, stable radix sort,
verified runs) against the reference's own compress (goldens from the unmodified PAmatrix::compress + Inference
constructor: counts, mrca, freq in first-occurrence order -- bit-exact) and against the host path on a seeded matrix
with many duplicates; the device bit permutation / transposition of the pattern rows; the "ccc" writer round trip."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ppm():
    import ppmmodelpangenome_b200 as p
    return p


class _Env:
    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name", ["ref_test", "syn12", "syn100", "syn300", "cat40"])
def test_device_dedupe_equals_reference_compress(ppm, golden, name):
    g = golden(name + "_ccc")  # outputs of the reference after PAmatrix::compress
    with _Env(PPM_GPU_DEDUPE_MIN=0):
        inf = ppm.Inference(golden(name).tree, golden(name).pa, dedupe=True, device=0)
    assert inf.nRep == g.meta["nRep"] and inf.nObs == g.meta["nObs"]
    assert np.array_equal(inf.counts, g.z["counts"])
    assert np.array_equal(inf.mrca, g.z["mrca"])
    assert np.array_equal(inf.freq, g.z["freq"])
    assert inf.meanGenes == g.meta["meanGenes"]
    r = g.results[0]
    ll, i2 = inf.loglik_batch(np.array([r["params"]]), return_i2=True)
    assert abs(ll[0] / r["ll"] - 1) <= 1e-9 and abs(i2[0] / r["i2"] - 1) <= 1e-9
    inf.close()


def test_device_dedupe_equals_host_dedupe_on_many_duplicates(ppm):
    from ppmmodelpangenome_b200 import simulate
    rng = np.random.default_rng(5)
    nwk = simulate.sim_tree(333, rng)
    names, M, truth = simulate.sim_genes(nwk, 20000, rng)
    left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
    rows = np.concatenate([rows, rows[::3], rows[:50]])  # every third row again, far from its first occurrence
    counts = rng.integers(1, 5, len(rows)).astype(np.int32)
    out = {}
    for how, env in (("gpu", dict(PPM_GPU_DEDUPE_MIN=0)), ("host", dict(PPM_GPU_DEDUPE_MIN=10**12))):
        with _Env(**env):
            inf = ppm.Inference.from_arrays(left, right, bl, rows, counts=counts, dedupe=True, device=0)
        out[how] = (int(inf.nRep), int(inf.nObs), inf.counts.copy(), inf.mrca.copy(), inf.freq.copy(), inf.gene_to_pattern.copy(),
                    float(inf.meanGenes))
        inf.close()
    a, b = out["gpu"], out["host"]
    assert a[0] == b[0] and a[1] == b[1] and a[6] == b[6]
    for k in (2, 3, 4, 5):
        assert np.array_equal(a[k], b[k])
    assert a[0] < len(rows)


def test_ccc_writer_round_trip(ppm, golden, tmp_path):
    g = golden("syn100")
    a = ppm.Inference(g.tree, g.pa, dedupe=True, device=0)
    out = str(tmp_path / "compressed.txt")
    a.write_ccc(out)
    assert open(out).readline().startswith("ccc,")
    b = ppm.Inference(g.tree, out, dedupe=False, device=0)  # already unique: read as it is, counts from the header
    assert b.nRep == a.nRep and b.nObs == a.nObs
    assert np.array_equal(a.counts, b.counts) and np.array_equal(a.mrca, b.mrca) and np.array_equal(a.freq, b.freq)
    p = np.array([g.results[0]["params"]])
    assert np.array_equal(a.loglik_batch(p), b.loglik_batch(p))
    a.close(); b.close()
