"""CPU tier: the oracle (oracle/ppm_oracle.c, a plain-C restatement of the reference's log-domain algorithm)
is pinned against what the UNMODIFIED reference produced (tests/golden, oracle/gen_golden.py) and -- where the
reference library exists (this container) -- against the live reference on fresh seeded inputs."""
import numpy as np
import pytest

from conftest import CASES, rel
from oracle import cport, model, ref


@pytest.mark.parametrize("name", CASES + ["ref_test_ccc", "syn100_ccc"])
def test_python_model_builder_matches_reference_fields(golden, name):
    g = golden(name)
    m = model.model_from_files(g.tree, g.pa, dedupe=g.meta["compress"])
    z, meta = g.z, g.meta
    assert (m["n_rep"], m["n_obs"], m["n_nodes"], m["n_leaves"]) == (meta["nRep"], meta["nObs"], meta["nNodes"], meta["nLeaves"])
    assert m["H"] == meta["H"] and m["L"] == meta["L"] and m["mean_genes"] == meta["meanGenes"]
    for k in ("mrca", "freq", "counts", "left", "right"):
        assert np.array_equal(m[k], z[k]), k
    assert np.array_equal(m["height"], z["height"]) and np.array_equal(m["bl"], z["bl"])
    nl = m["n_leaves"]
    # std::sort leaves ties among (equal-height) leaves unspecified: compare the part the integral uses
    assert np.array_equal(np.sort(m["order"][:nl]), np.sort(z["order"][:nl]))
    assert np.array_equal(m["order"][nl:], z["order"][nl:])
    sizes = np.diff(m["sub_off"])
    assert np.array_equal(sizes[nl - 1:], z["subtree_sizes"][nl - 1:])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(golden, name):
    g = golden(name)
    orc = cport.Oracle(model.model_from_files(g.tree, g.pa))
    for k, r in enumerate(g.results):
        e = orc.eval(g.params[k], components=r["i2"] > 0, cats=r["i2"] > 0)
        assert rel(e["ll"], r["ll"]) <= 1e-13, (name, k)
        assert rel(e["i2"], r["i2"]) <= 1e-13
        assert np.allclose(e["parts"], r["parts"], rtol=1e-13, atol=0)
        if r["i2"] > 0:
            ref_c = g.z["components"][k]
            same = (e["comp"] == ref_c) | (np.abs(e["comp"] - ref_c) <= 1e-12 * np.abs(ref_c))
            assert same.all()
            assert np.array_equal(e["cat"], g.z["cats"][k])


def test_oracle_known_answers_appendix_c(golden):
    """the nine (computeI2, logLik) pairs of the shipped test data."""
    g = golden("ref_test")
    assert g.meta["nRep"] == 1410 and g.meta["nNodes"] == 39
    assert g.results[0]["ll"] == -13657.35321059162
    assert g.results[8]["ll"] == -1000003003.8057114
    orc = cport.Oracle(model.model_from_files(g.tree, g.pa, dedupe=True))
    assert orc.m["n_rep"] == 626 and list(orc.m["counts"][:12]) == [168, 6, 6, 2, 4, 3, 5, 1, 1, 7, 3, 2]
    assert rel(orc.loglik(g.params[0]), -13657.353210591615) <= 1e-14


def test_oracle_pattern_slices_add_up(golden):
    g = golden("syn100")
    orc = cport.Oracle(model.model_from_files(g.tree, g.pa))
    full = orc.eval(g.params[0])["ll"]
    n = orc.m["n_rep"]
    parts = [orc.eval(g.params[0], first_rep=a, n_eval=b - a)["ll"] for a, b in ((0, 100), (100, 101), (101, n))]
    assert rel(sum(parts), full) <= 1e-13


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the compiled reference) is not present")
def test_oracle_against_live_reference_on_seeded_inputs(tmp_path):
    from ppmmodelpangenome_b200 import simulate
    for seed, n, genes, cat in [(1, 5, 30, False), (2, 24, 90, False), (3, 31, 50, True)]:
        rng = np.random.default_rng(seed)
        nwk = simulate.sim_tree(n, rng, caterpillar=cat)
        names, M, truth = simulate.sim_genes(nwk, genes, rng)
        tp, pp = str(tmp_path / "t.nwk"), str(tmp_path / "m.txt")
        open(tp, "w").write(nwk + "\n")
        simulate.write_matrix(pp, names, M)
        R = ref.Reference(tp, pp)
        m = model.model_from_files(tp, pp)
        assert np.array_equal(m["mrca"], R.mrca) and np.array_equal(m["freq"], R.freq)
        assert np.array_equal(m["height"], R.height)
        orc = cport.Oracle(m)
        for _ in range(3):
            p = simulate.start_params(truth, genes, rng)
            assert rel(orc.loglik(p), R.loglik(p)) <= 1e-13
            assert rel(orc.eval(p)["i2"], R.compute_i2(p)) <= 1e-13
        R.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the compiled reference) is not present")
def test_posterior_category_probabilities_match_reference_file(golden):
    """computeProbaCat (functions.cpp:556-581) writes log posteriors at 6 significant digits."""
    g = golden("ref_test")
    p = g.params[2]
    R = ref.Reference(g.tree, g.pa)
    i2 = R.compute_i2(p)
    pc = R.proba_cat(p, i2)
    c = cport.Oracle(model.model_from_files(g.tree, g.pa)).eval(p, components=True)["comp"]
    c3 = np.stack([c[:, 0], np.logaddexp(c[:, 1], c[:, 2]), c[:, 3]], axis=1)
    lp = c3 - np.logaddexp(np.logaddexp(c3[:, 0], c3[:, 1]), c3[:, 2])[:, None]
    assert np.allclose(pc, lp, rtol=2e-5, atol=1e-12)
    R.close()
