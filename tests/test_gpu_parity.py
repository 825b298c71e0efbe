"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libppm_b200.so), against
(1) the committed golden outputs of the unmodified reference, (2) the C oracle on seeded inputs, and
(3) size-independent properties at larger sizes.  Tolerances: log-likelihood / i2 1e-9 relative (north star);
integers (counts, mrca, freq, categories) bit-exact."""
import os

import numpy as np
import pytest

from conftest import CASES, rel

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def ppm():
    import ppmmodelpangenome_b200 as p
    return p


def _engine(ppm, g, dedupe, **kw):
    return ppm.Inference(g.tree, g.pa, dedupe=dedupe, device=0, **kw)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dedupe", [False, True])
def test_loglik_i2_against_reference_golden(ppm, golden, name, dedupe):
    g = golden(name)
    inf = _engine(ppm, g, dedupe)
    ll, i2 = inf.loglik_batch(g.params, return_i2=True)
    i2b, parts = inf.computeI2(g.params, return_parts=True)
    for k, r in enumerate(g.results):
        assert rel(ll[k], r["ll"]) <= RTOL, (name, k, ll[k], r["ll"])
        assert rel(i2[k], r["i2"]) <= RTOL
        assert rel(i2b[k], r["i2"]) <= RTOL
        assert np.allclose(parts[k], r["parts"], rtol=RTOL, atol=1e-13)  # p1 = 1-exp(.) cancels: absolute bound
    # one-at-a-time evaluation must agree bitwise with the batched one (same kernels, same order)
    # (every vector on its own: launches whose parameter vectors are all factored, pi1 > 0.95, or all not)
    for k in range(len(g.params)):
        assert inf.logLik(g.params[k]) == ll[k], (name, k)
    inf.close()


@pytest.mark.parametrize("name", CASES)
def test_components_and_categories_against_reference_golden(ppm, golden, name):
    g = golden(name)
    inf = _engine(ppm, g, dedupe=False)
    for k, r in enumerate(g.results):
        if r["i2"] < 0:
            continue
        comp = inf.components(g.params[k], r["i2"])
        ref3 = g.ref3(k)
        fin = np.isfinite(ref3)
        # where the reference's exp(f_aux) underflowed (lik2 = -inf) the scaled engine is finite: skip those
        assert np.allclose(comp[fin], ref3[fin], rtol=1e-9, atol=1e-9), (name, k)
        cat = inf.assignCat(g.params[k], r["i2"])
        ok = np.isfinite(g.z["components"][k][:, 3])
        assert np.array_equal(cat[ok], g.z["cats"][k][ok]), (name, k)
    inf.close()


def test_integer_model_fields_bit_exact(ppm, golden):
    for name in CASES:
        for dd in (False, True):
            g = golden(name + "_ccc" if dd else name)
            inf = _engine(ppm, g, dedupe=dd)
            assert inf.nRep == g.meta["nRep"] and inf.nObs == g.meta["nObs"]
            assert np.array_equal(inf.counts, g.z["counts"])
            assert np.array_equal(inf.mrca, g.z["mrca"])
            assert np.array_equal(inf.freq, g.z["freq"])
            inf.close()


def test_dedupe_scatter_and_sharding(ppm, golden):
    g = golden("syn100")
    full = _engine(ppm, g, dedupe=False)
    dd = _engine(ppm, g, dedupe=True)
    p = g.params[:3]
    ll_full = full.loglik_batch(p)
    ll_dd = dd.loglik_batch(p)
    assert np.allclose(ll_full, ll_dd, rtol=1e-12, atol=0)
    i2 = dd.computeI2(p[0])
    assert np.array_equal(full.assignCat(p[0], i2), dd.assignCat(p[0], i2))  # scattered back to input genes
    # shards: partial sums add up to the whole; penalty applied after the sum
    for world in (2, 3, 8):
        tot = np.zeros(len(p))
        for rk in range(world):
            sh = _engine(ppm, g, dedupe=True, shard_rank=rk, shard_world=world)
            tot += sh.loglik_batch(p)
            sh.close()
        assert np.allclose(tot, ll_dd, rtol=1e-12, atol=0), world
    full.close(); dd.close()


def test_seeded_random_inputs_against_oracle(ppm, tmp_path):
    from oracle import cport, model
    from ppmmodelpangenome_b200 import simulate
    for seed, n, genes, cat in [(101, 7, 40, False), (102, 33, 150, False), (103, 64, 100, True), (104, 130, 60, False)]:
        rng = np.random.default_rng(seed)
        nwk = simulate.sim_tree(n, rng, caterpillar=cat)
        names, M, truth = simulate.sim_genes(nwk, genes, rng)
        tp, pp = str(tmp_path / ("t%d.nwk" % seed)), str(tmp_path / ("m%d.txt" % seed))
        open(tp, "w").write(nwk + "\n")
        simulate.write_matrix(pp, names, M)
        m = model.model_from_files(tp, pp)
        orc = cport.Oracle(m)
        inf = ppm.Inference(tp, pp, dedupe=True, device=0)
        P = np.stack([simulate.start_params(truth, genes, rng) for _ in range(4)])
        ll = inf.loglik_batch(P)
        for k in range(len(P)):
            o = orc.eval(P[k], cats=True)
            assert rel(ll[k], o["ll"]) <= RTOL, (seed, k, ll[k], o["ll"])
            if o["i2"] > 0:
                assert np.array_equal(inf.assignCat(P[k], o["i2"]), o["cat"])
        # the flat-array entry point gives the same answer as the file entry point
        left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
        inf2 = ppm.Inference.from_arrays(left, right, bl, rows, dedupe=True, device=0)
        assert np.array_equal(inf2.loglik_batch(P), ll)
        inf.close(); inf2.close()


def test_large_batch_and_bigger_tree_properties(ppm, tmp_path):
    """Sizes the oracle cannot finish quickly: check invariances instead.
    (a) permuting the genes leaves logLik unchanged to 1e-12; (b) duplicating every gene doubles nObs and the
    per-gene terms consistently: ll(2x data, 2x N0, 2x i1) = 2 ll(data) - 0 (mixture weights are ratios);
    (c) a batch of P vectors equals P single calls bitwise."""
    from ppmmodelpangenome_b200 import simulate
    rng = np.random.default_rng(7)
    nwk = simulate.sim_tree(400, rng)
    names, M, truth = simulate.sim_genes(nwk, 3000, rng)
    left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
    P = np.stack([simulate.start_params(truth, 3000, rng) for _ in range(16)])
    a = ppm.Inference.from_arrays(left, right, bl, rows, dedupe=True, device=0)
    ll = a.loglik_batch(P)
    assert np.all(np.isfinite(ll))
    perm = rng.permutation(len(rows))
    b = ppm.Inference.from_arrays(left, right, bl, rows[perm], dedupe=True, device=0)
    assert np.allclose(b.loglik_batch(P), ll, rtol=1e-12, atol=0)
    c = ppm.Inference.from_arrays(left, right, bl, np.concatenate([rows, rows]), dedupe=True, device=0)
    P2 = P.copy(); P2[:, 0] *= 2; P2[:, 2] *= 2
    ll2, i22 = c.loglik_batch(P2, return_i2=True)
    _, i21 = a.loglik_batch(P, return_i2=True)
    assert np.allclose(i22, 2 * i21, rtol=1e-9)
    assert np.allclose(ll2, 2 * ll, rtol=1e-9)
    for k in (0, 7, 15):
        assert a.logLik(P[k]) == ll[k]
    a.close(); b.close(); c.close()


def test_penalty_and_errors(ppm, golden, tmp_path):
    g = golden("ref_test")
    inf = _engine(ppm, g, dedupe=True)
    r = g.results[8]  # i2 < 0 (Appendix C last row)
    ll, i2 = inf.loglik_batch([r["params"]], return_i2=True)
    assert i2[0] < 0 and ll[0] == -1e9 + i2[0]
    assert rel(ll[0], r["ll"]) <= 1e-12
    inf.close()
    with pytest.raises(ppm.PPMError):
        ppm.Inference(str(tmp_path / "missing.nwk"), g.pa)
    bad = tmp_path / "bad.nwk"
    bad.write_text("((a:1,b:1,c:1):1,d:2);\n")
    with pytest.raises(ppm.PPMError):
        ppm.Inference(str(bad), g.pa)


def test_posterior_category_probabilities(ppm, golden):
    """ppm_proba_cat == computeProbaCat (functions.cpp:556-581): normalised log posteriors per input gene."""
    from oracle import cport, model
    for name, dedupe in (("ref_test", True), ("syn100", False)):
        g = golden(name)
        k = 2 if name == "ref_test" else 0
        p, i2 = g.params[k], g.results[k]["i2"]
        inf = _engine(ppm, g, dedupe)
        got = inf.computeProbaCat(p, i2)
        c = cport.Oracle(model.model_from_files(g.tree, g.pa)).eval(p, components=True)["comp"]
        c3 = np.stack([c[:, 0], np.logaddexp(c[:, 1], c[:, 2]), c[:, 3]], axis=1)
        lp = c3 - np.logaddexp(np.logaddexp(c3[:, 0], c3[:, 1]), c3[:, 2])[:, None]
        fin = np.isfinite(c3).all(axis=1)
        assert got.shape == lp.shape
        assert np.allclose(got[fin], lp[fin], rtol=1e-9, atol=1e-9)
        assert np.allclose(np.exp(got).sum(axis=1), 1.0, atol=1e-12)
        inf.close()


@pytest.mark.parametrize("name", ["ref_test", "syn12", "syn100", "cat40"])
def test_posterior_accessory_outputs_against_reference_golden(ppm, golden, name):
    """ppm_expected_time1 / ppm_expected_time2 / ppm_expected_gains == expectedTime1, expectedTime2_times / _aux +
    exp(likRep2), expectedGains (functions.cpp:582-757); goldens by the unmodified reference
    (oracle/gen_golden_posterior.py)."""
    g = golden(name)
    z = np.load(os.path.join(os.path.dirname(g.tree), "posterior_" + name + ".npz"))
    inf = _engine(ppm, g, dedupe=False)
    et1 = inf.expectedTime1(z["params"])
    assert np.allclose(et1, z["et1"], rtol=RTOL, atol=0)
    times, dist = inf.expectedTime2(z["params"])
    assert np.allclose(times, z["times"], rtol=1e-12, atol=1e-15)
    assert dist.shape == z["dist"].shape
    assert np.allclose(dist, z["dist"], rtol=RTOL, atol=1e-300)
    # the integral column is the weighted sum of the integrand columns (functions.cpp:273-275)
    for g2, want in zip(z["gains_g2"], z["gains"]):
        assert rel(inf.expectedGains(g2), want) <= 1e-10
    # deduplicated handle: one row per unique pattern, same values
    d = _engine(ppm, g, dedupe=True)
    g2p = d.gene_to_pattern
    assert np.array_equal(d.expectedTime1(z["params"])[g2p], et1)
    dd = d.expectedTime2(z["params"])[1][g2p]
    assert np.array_equal(dd[:, :-1], dist[:, :-1])            # per-slice integrands: same arithmetic per lane
    assert np.allclose(dd[:, -1], dist[:, -1], rtol=1e-13, atol=0)  # the sum's order depends on the neighbouring lanes
    d.close()
    inf.close()


def test_posterior_outputs_global_memory_variant(ppm, golden):
    """syn300 runs the global-memory sweep variant: its per-slice integrand must sum to its own integral column and
    the integral must match the Mobile component of ppm_components."""
    g = golden("syn300")
    inf = _engine(ppm, g, dedupe=False)
    p = g.params[0]
    times, dist = inf.expectedTime2(p)
    hgt = np.zeros(inf.info.n_nodes)
    from ppmmodelpangenome_b200.engine import lib
    assert lib().ppm_get_double_array(inf._h, 0, hgt.ctypes.data, len(hgt)) == 0
    assert np.all(np.diff(times) > 0) and times[-1] < hgt.max()
    i2 = inf.computeI2(p)
    comp = inf.components(p, i2)
    with np.errstate(divide="ignore"):
        lik2 = np.log(i2 / inf.nObs) + np.log(dist[:, -1])
    fin = np.isfinite(lik2) & (dist[:, -1] > 1e-290)
    assert fin.sum() > 0
    assert np.allclose(lik2[fin], comp[fin, 2], rtol=1e-12, atol=1e-12)
    inf.close()


def test_group_handle_matches_unsharded_handle(ppm, golden):
    """ppm_create_group*: three pattern shards behind one handle (device ids may repeat, so this also runs on a
    one-GPU box; on a multi-GPU box the same code shards over NVLink peers).  The objective equals the unsharded one
    to summation-order accuracy, penalties included; per-pattern outputs are identical."""
    import torch
    ndev = torch.cuda.device_count()
    devs = [k % ndev for k in range(3)]
    for name in ("ref_test", "syn300"):
        g = golden(name)
        one = _engine(ppm, g, dedupe=True)
        grp = ppm.Inference(g.tree, g.pa, dedupe=True, devices=devs)
        assert grp.info.group_size == 3 and grp.nRep == one.nRep
        ll1, i21 = one.loglik_batch(g.params, return_i2=True)
        ll3, i23 = grp.loglik_batch(g.params, return_i2=True)
        assert np.array_equal(i21, i23)
        assert np.allclose(ll3, ll1, rtol=1e-13, atol=0)
        for k, r in enumerate(g.results):
            assert rel(ll3[k], r["ll"]) <= RTOL
        assert np.array_equal(ll3, grp.loglik_batch(g.params))  # reproducible
        k = 2 if name == "ref_test" else 0
        p, i2 = g.params[k], g.results[k]["i2"]
        assert np.array_equal(grp.assignCat(p, i2), one.assignCat(p, i2))
        # (the integral's summation order depends on which patterns share a warp: not bitwise)
        assert np.allclose(grp.components(p, i2), one.components(p, i2), rtol=1e-13, atol=0)
        assert np.array_equal(grp.expectedTime1(p), one.expectedTime1(p))
        assert np.allclose(grp.computeProbaCat(p, i2), one.computeProbaCat(p, i2), rtol=1e-12, atol=1e-12)
        grp.close(); one.close()


def test_group_with_more_shards_than_patterns_and_bad_device(ppm, golden, tmp_path):
    """Edge cases of group handles: empty shards (more GPUs than patterns) still give the full objective; an invalid
    device ordinal is an error, not a crash."""
    tree = tmp_path / "t.nwk"
    tree.write_text("((a:1,b:1):1,(c:1.5,d:1.5):0.5);\n")
    pa = tmp_path / "pa.txt"
    pa.write_text("genome,g1,g2,g3\na,1,0,1\nb,1,0,0\nc,0,1,1\nd,0,1,1\n")
    p = np.array([[2.0, 0.05, 1.0, 0.5, 0.5, 2.0, 0.01, 0.02]])
    one = ppm.Inference(str(tree), str(pa), dedupe=True, device=0)
    grp = ppm.Inference(str(tree), str(pa), dedupe=True, devices=[0] * 5)   # 3 patterns over 5 shards
    a, b = one.loglik_batch(p, return_i2=True), grp.loglik_batch(p, return_i2=True)
    assert np.allclose(a[0], b[0], rtol=1e-13, atol=0) and np.array_equal(a[1], b[1])
    assert np.array_equal(one.assignCat(p[0], a[1][0]), grp.assignCat(p[0], a[1][0]))
    assert np.array_equal(one.expectedTime1(p[0]), grp.expectedTime1(p[0]))
    assert np.array_equal(one.expectedTime2(p[0])[1][:, :-1], grp.expectedTime2(p[0])[1][:, :-1])
    one.close(); grp.close()
    with pytest.raises(ppm.PPMError):
        ppm.Inference(str(tree), str(pa), devices=[0, 99])


@pytest.mark.parametrize("n,genes,cat", [(1000, 120, False), (6200, 16, True)])
def test_big_trees_against_oracle(ppm, tmp_path, n, genes, cat):
    """The global-memory sweep variant (1,000 genomes: records streamed through the per-warp rings, small tables staged)
    and its big-tree form (6,200-leaf caterpillar: pattern rows read in place, deeper pair pipeline) against the C oracle
    on inputs small enough for it: log-likelihood, i2 and categories."""
    from oracle import cport, model
    from ppmmodelpangenome_b200 import simulate
    rng = np.random.default_rng(1000 + n)
    nwk = simulate.sim_tree(n, rng, caterpillar=cat)
    names, M, truth = simulate.sim_genes(nwk, genes, rng)
    tp, pp = str(tmp_path / "t.nwk"), str(tmp_path / "m.txt")
    open(tp, "w").write(nwk + "\n")
    simulate.write_matrix(pp, names, M)
    orc = cport.Oracle(model.model_from_files(tp, pp))
    inf = ppm.Inference(tp, pp, dedupe=True, device=0)
    P = np.stack([simulate.start_params(truth, genes), simulate.start_params(truth, genes, rng)])
    ll, i2 = inf.loglik_batch(P, return_i2=True)
    for k in range(len(P)):
        o = orc.eval(P[k], cats=(k == 0))
        assert rel(ll[k], o["ll"]) <= RTOL, (n, k, ll[k], o["ll"])
        assert rel(i2[k], o["i2"]) <= RTOL
        if k == 0 and o["i2"] > 0:
            assert np.array_equal(inf.assignCat(P[k], o["i2"]), o["cat"])
    inf.close()
