"""CPU tier: the C++ PPM gene simulator (csrc/simulator.cpp, ppm_simulate_genes; input generation for the benchmarks):
thread-count independence, no all-zero gene, category order, and agreement of the per-category frequency profile with the
numpy restatement of the reference's R simulators (ppmmodelpangenome_b200/simulate.py)."""
import numpy as np

from ppmmodelpangenome_b200 import Inference, simulate, simulate_genes


def _freq(rows):
    f = np.zeros(len(rows), int)
    for w in range(rows.shape[1]):
        x = rows[:, w].copy()
        while x.any():
            f += (x & 1).astype(int)
            x >>= 1
    return f


def test_simulator_is_deterministic_and_matches_the_numpy_model():
    rng = np.random.default_rng(3)
    nwk = simulate.sim_tree(120, rng)
    left, right, bl, labels = simulate.parse_newick(nwk)
    G = 12000
    rows, truth = simulate_genes(left, right, bl, G, 11, threads=1)
    rows4, _ = simulate_genes(left, right, bl, G, 11, threads=4)
    assert np.array_equal(rows, rows4)                      # one random stream per gene: threads do not matter
    head, _ = simulate_genes(left, right, bl, 3000, 11)     # gene g has its own stream: the first 750 are Persistent in both runs
    assert np.array_equal(head[:750], rows[:750])
    assert not np.array_equal(rows, simulate_genes(left, right, bl, G, 12)[0])
    f = _freq(rows)
    assert f.min() >= 1 and f.max() <= 120                  # all-zero genes are redrawn; padding bits stay clear
    n0, n1 = round(G * 0.25), round(G * 0.35)
    names, M, tr = simulate.sim_genes(nwk, G, np.random.default_rng(5))
    fn = M.sum(axis=0)
    for a, b in ((f[:n0], fn[:n0]), (f[n0:n0 + n1], fn[n0:n0 + n1]), (f[n0 + n1:], fn[n0 + n1:])):
        assert abs(a.mean() - b.mean()) <= 0.08 * b.mean() + 0.3, (a.mean(), b.mean())
    assert f[:n0].mean() > 100 > f[n0:].mean()              # Persistent genes nearly everywhere, the rest sparse
    for k in ("l0", "l1", "g2", "l2", "H", "L"):
        assert abs(truth[k] - tr[k]) <= 1e-12 * abs(tr[k])
    # the rows are what ppm_create expects
    inf = Inference.from_arrays(left, right, bl, rows, dedupe=True, device=-1)
    assert inf.nGenes == G and inf.nObs == G and 1 < inf.nRep < G
    assert np.array_equal(np.sort(inf.freq), np.sort(_freq(rows)[np.unique(inf.gene_to_pattern, return_index=True)[1]]))
    inf.close()
