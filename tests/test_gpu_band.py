"""GPU parity of the band kernel (slice-parallel sweep of big trees, csrc/band_kernel.cu) through the C ABI:
against the reference's golden outputs (forced on the 100/300-leaf goldens), against the C oracle on seeded
1,000- and 5,000-genome trees, for every (slices per lane R, warps per item G) variant, with parameter vectors the
band kernel must hand to the lane-per-pattern sweep in the same launch, and sharded.  Tolerance 1e-9 relative."""
import os

import numpy as np
import pytest

from conftest import rel

pytestmark = pytest.mark.gpu
RTOL = 1e-9


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
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name", ["syn100", "syn300", "cat40"])
@pytest.mark.parametrize("V,R,G", [(2, 4, 1), (2, 2, 1), (2, 4, 2), (2, 2, 5), (2, 4, 3),
                                   (1, 4, 1), (1, 2, 1), (1, 8, 1), (1, 4, 2), (1, 4, 4), (1, 2, 8), (1, 4, 16)])
def test_band_kernel_against_reference_golden(ppm, golden, name, V, R, G):
    """V = 2: stream layout (prune / integral / mixture kernels, G work units per item); V = 1: one CTA of G warps per item."""
    g = golden(name)
    with _Env(PPM_FORCE_BAND=1, PPM_BAND_R=R, PPM_BAND_G=G, PPM_BAND_V=V):
        inf = ppm.Inference(g.tree, g.pa, dedupe=True, device=0)
    pl = inf.plan()
    assert pl["band"] == 1 and pl["band_R"] == R and pl["band_stream"] == (V == 2)
    assert pl["band_G"] == (min(G, pl["band_tiles"]) if V == 2 else G)
    ll, i2 = inf.loglik_batch(g.params, return_i2=True)
    assert inf.plan()["last_kind"] == 1
    for k, r in enumerate(g.results):
        assert rel(ll[k], r["ll"]) <= RTOL, (name, k, ll[k], r["ll"])
        assert rel(i2[k], r["i2"]) <= RTOL
    assert np.array_equal(ll, inf.loglik_batch(g.params))  # reproducible
    inf.close()
    # the same handle without the band kernel agrees to summation-order accuracy
    with _Env(PPM_NO_BAND=1):
        old = ppm.Inference(g.tree, g.pa, dedupe=True, device=0)
    assert old.plan()["band"] == 0
    assert np.allclose(old.loglik_batch(g.params), ll, rtol=1e-12, atol=0)
    old.close()


def _seeded(ppm, tmp_path, n, genes, seed):
    from ppmmodelpangenome_b200 import simulate
    rng = np.random.default_rng(seed)
    nwk = simulate.sim_tree(n, rng)
    names, M, truth = simulate.sim_genes(nwk, genes, rng)
    tp, pp = str(tmp_path / "t.nwk"), str(tmp_path / "m.txt")
    open(tp, "w").write(nwk + "\n")
    simulate.write_matrix(pp, names, M)
    P = np.stack([simulate.start_params(truth, genes), simulate.start_params(truth, genes, rng), simulate.start_params(truth, genes, rng)])
    return tp, pp, P, truth


@pytest.mark.parametrize("n,genes", [(1500, 80), (5000, 16)])
def test_band_kernel_big_trees_against_oracle(ppm, tmp_path, n, genes):
    """Default plan: the band kernel is chosen on its own for these trees (the oracle on a sample it can hold)."""
    from oracle import cport, model
    tp, pp, P, truth = _seeded(ppm, tmp_path, n, genes, 4000 + n)
    orc = cport.Oracle(model.model_from_files(tp, pp))
    inf = ppm.Inference(tp, pp, dedupe=True, device=0)
    assert inf.plan()["band"] == 1, inf.plan()
    # + one vector the band kernel cannot serve (pi1 > 0.95: factored leaf pairs): the fallback launch of the same call
    if n > 2000:
        P = P[:1]  # (the oracle needs ~2 s per pattern-eval and core at 5,000 genomes)
    Pm = np.vstack([P, P[:1]])
    Pm[-1, 4], Pm[-1, 5] = 30.0, 1.0
    ll, i2 = inf.loglik_batch(Pm, return_i2=True)
    assert inf.plan()["last_kind"] == 1
    for k in range(len(Pm)):
        o = orc.eval(Pm[k], cats=False)
        assert rel(i2[k], o["i2"]) <= RTOL
        assert rel(ll[k], o["ll"]) <= RTOL, (n, k, ll[k], o["ll"])
    # every vector alone (host hint: all served by the band kernel, no fallback launch) agrees bitwise with the batch
    for k in range(len(P)):
        assert inf.logLik(Pm[k]) == ll[k]
    # per-pattern Mobile components come from the lane-per-pattern sweep: the two kernels agree per pattern
    inf.close()


def test_band_kernel_shards_add_up(ppm, tmp_path):
    tp, pp, P, truth = _seeded(ppm, tmp_path, 1000, 200, 77)
    with _Env(PPM_FORCE_BAND=1):
        one = ppm.Inference(tp, pp, dedupe=True, device=0)
    ll, i2 = one.loglik_batch(P, return_i2=True)
    tot = np.zeros(len(P))
    for rk in range(3):
        with _Env(PPM_FORCE_BAND=1):
            s = ppm.Inference(tp, pp, dedupe=True, device=0, shard_rank=rk, shard_world=3)
        assert s.plan()["band"] == 1
        part, i2s = s.loglik_batch(P, return_i2=True)
        assert np.array_equal(i2s, i2)
        tot += part
        s.close()
    tot = np.where(i2 < 0, -1e9 + i2, tot)
    assert np.allclose(tot, ll, rtol=1e-13, atol=0)
    with _Env(PPM_FORCE_BAND=1):
        grp = ppm.Inference(tp, pp, dedupe=True, devices=[0, 0])
    assert np.allclose(grp.loglik_batch(P), ll, rtol=1e-13, atol=0)
    grp.close(); one.close()


def test_band_kernel_penalty_and_zero_error_rates(ppm, tmp_path):
    """i2 < 0 (penalty branch, functions.cpp:362-364), s_10 = 0 / s_01 = 0 (exact zero factors), g2 = 0."""
    from oracle import cport, model
    tp, pp, P, truth = _seeded(ppm, tmp_path, 400, 60, 5)
    orc = cport.Oracle(model.model_from_files(tp, pp))
    with _Env(PPM_FORCE_BAND=1):
        inf = ppm.Inference(tp, pp, dedupe=True, device=0)
    assert inf.plan()["band"] == 1
    Q = np.repeat(P[:1], 4, axis=0)
    Q[0, 0] *= 50           # N0 far too large: i2 < 0
    Q[1, 6] = 0.0           # s_10 = 0
    Q[2, 7] = 0.0           # s_01 = 0
    Q[3, 4] = 0.0           # g2 = 0
    ll, i2 = inf.loglik_batch(Q, return_i2=True)
    assert i2[0] < 0 and ll[0] == -1e9 + i2[0]
    for k in range(4):
        o = orc.eval(Q[k], cats=False)
        assert rel(i2[k], o["i2"]) <= RTOL
        if np.isfinite(o["ll"]):
            assert rel(ll[k], o["ll"]) <= RTOL, (k, ll[k], o["ll"])
    inf.close()


def test_sparse_evaluation_shards_add_up(ppm, tmp_path):
    tp, pp, P, truth = _seeded(ppm, tmp_path, 1400, 200, 91)
    one = ppm.Inference(tp, pp, dedupe=True, device=0)
    ll, i2 = one.loglik_batch(P, return_i2=True)
    tot = np.zeros(len(P))
    for rk in range(2):
        with _Env(PPM_SPARSE=1):
            s = ppm.Inference(tp, pp, dedupe=True, device=0, shard_rank=rk, shard_world=2)
        tot += s.loglik_batch(P)
        s.close()
    assert np.allclose(np.where(i2 < 0, -1e9 + i2, tot), ll, rtol=1e-12, atol=0)
    with _Env(PPM_SPARSE=1):
        grp = ppm.Inference(tp, pp, dedupe=True, devices=[0, 0, 0])
    assert np.allclose(grp.loglik_batch(P), ll, rtol=1e-12, atol=0)
    grp.close(); one.close()


def test_band_stream_batches_are_invisible(ppm, tmp_path):
    """Stream layout: the scratch budget only sets how many items one prune/integral/mixture round serves; results are
    bitwise the same with one batch and with many (PPM_BAND_SCRATCH_MB=1: a few 32-item blocks per batch)."""
    tp, pp, P, truth = _seeded(ppm, tmp_path, 1300, 150, 11)
    one = ppm.Inference(tp, pp, dedupe=True, device=0)
    assert one.plan()["band"] == 1 and one.plan()["band_stream"] == 1
    ll = one.loglik_batch(P)
    n1 = one.counters()["launches"]
    with _Env(PPM_BAND_SCRATCH_MB=1):
        many = one.loglik_batch(P)
        assert one.counters()["launches"] > n1
    assert np.array_equal(ll, many)
    one.close()


@pytest.mark.parametrize("n,genes", [(1500, 300), (5000, 40)])
def test_sparse_evaluation_equals_dense(ppm, tmp_path, n, genes):
    """ppm_set_sparse: dirty lineages only, as a ratio against the all-zero row -- the dense result to rounding,
    for parameter vectors the band kernels serve, for the penalty branch and for one they hand to the sweep."""
    tp, pp, P, truth = _seeded(ppm, tmp_path, n, genes, 600 + n)
    inf = ppm.Inference(tp, pp, dedupe=True, device=0)
    assert inf.plan()["band_stream"] == 1
    Q = np.vstack([P, P[:1], P[:1]])
    Q[-2, 0] *= 50                      # i2 < 0
    Q[-1, 4], Q[-1, 5] = 30.0, 1.0      # pi1 > 0.95: lane-per-pattern sweep
    dense = inf.loglik_batch(Q)
    inf.set_sparse(True)
    sparse = inf.loglik_batch(Q)
    assert np.array_equal(sparse, inf.loglik_batch(Q))  # reproducible
    inf.set_sparse(False)
    assert np.array_equal(dense, inf.loglik_batch(Q))
    assert np.allclose(sparse, dense, rtol=1e-11, atol=0), (sparse, dense)
    inf.close()


@pytest.mark.parametrize("name", ["syn100", "syn300", "cat40"])
def test_sparse_evaluation_against_reference_golden(ppm, golden, name):
    g = golden(name)
    with _Env(PPM_FORCE_BAND=1, PPM_SPARSE=1):
        inf = ppm.Inference(g.tree, g.pa, dedupe=True, device=0)
    ll, i2 = inf.loglik_batch(g.params, return_i2=True)
    for k, r in enumerate(g.results):
        assert rel(ll[k], r["ll"]) <= RTOL, (name, k, ll[k], r["ll"])
    inf.close()
