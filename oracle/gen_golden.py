"""TEST INFRASTRUCTURE ONLY. Generates tests/golden/* by running the UNMODIFIED reference
(oracle/_ref/libppm_ref.so, built by oracle/Makefile from /root/reference/source) in this container.
Run:  python oracle/gen_golden.py      (needs /root/reference; the outputs are committed)."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref, model  # noqa: E402
from ppmmodelpangenome_b200 import simulate  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_TEST = "/root/reference/test"

# SURVEY Appendix C parameter vectors [N0 l0 i1 l1 g2 l2 s_10 s_01]
APPENDIX_C = [
    [262, 0.0432618, 114, 8.01477, 1.60832, 159.089, 0.00586003, 0.00260816],
    [565.296351, 0.14830002, 233.17362, 20.3563992, 5.70230914, 514.059084, 0.0212420619, 0.00988912441],
    [327.99137692283216, 0.026236241587262604, 119.57435770257749, 4.2407983237765023, 25.933361027166828,
     279.11027279866653, 0.017071299071589354, 0.13992835635797601],
    [300, 0, 100, 5, 20, 300, 0.02, 0.1],
    [300, 0, 100, 5, 20, 300, 0, 0.1],
    [300, 0.05, 100, 5, 0, 300, 0.02, 0.1],
    [300, 0.05, 100, 5, 20, 300, 0.02, 0],
    [300, 0.05, 100, 5, 20, 0, 0.02, 0.1],
    [526.399022, 0.00189648595, 353.079156, 0.599166937, 42.7556724, 682.477483, 0.0757327973, 0.199456455],
]


def dump_case(name, tree_path, pa_path, vectors, compress):
    R = ref.Reference(tree_path, pa_path, compress)
    arrays = dict(mrca=R.mrca, freq=R.freq, counts=R.counts, order=R.order, height=R.height,
                  bl=R.branch_length, left=R.left, right=R.right)
    nsub = [len(R.subtrees(r)) for r in range(R.nNodes - 1)]
    arrays["subtree_sizes"] = np.array(nsub, np.int32)
    res = []
    comps, cats = [], []
    for v in vectors:
        i2, parts = R.compute_i2(v, True)
        ll = R.loglik(v)
        comps.append(R.components(v))
        cats.append(R.assign_cat(v, i2) if i2 > 0 else np.full(R.nRep, -1, np.int32))
        res.append(dict(params=[float(x) for x in v], ll=ll, i2=i2, parts=[float(x) for x in parts]))
    arrays["components"] = np.stack(comps)
    arrays["cats"] = np.stack(cats).astype(np.int8)
    meta = dict(name=name, compress=bool(compress), nRep=R.nRep, nObs=R.nObs, nNodes=R.nNodes, nLeaves=R.nLeaves,
                H=R.H, L=R.L, meanGenes=R.meanGenes, results=res)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrays)
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    R.close()
    print(name, "nRep", meta["nRep"], "ll0", res[0]["ll"])


def synth(name, n, genes, seed, caterpillar=False, nvec=3):
    rng = np.random.default_rng(seed)
    nwk = simulate.sim_tree(n, rng, caterpillar=caterpillar)
    names, M, truth = simulate.sim_genes(nwk, genes, rng)
    tpath = os.path.join(GOLD, name + ".nwk")
    ppath = os.path.join(GOLD, name + ".pa.txt")
    with open(tpath, "w") as f:
        f.write(nwk + "\n")
    simulate.write_matrix(ppath, names, M)
    vecs = [simulate.start_params(truth, genes)] + [simulate.start_params(truth, genes, rng) for _ in range(nvec - 1)]
    # edge cases on synthetic trees too
    e = list(vecs[0]); e[1] = 0.0; vecs.append(e)            # l0 == 0
    e = list(vecs[0]); e[4] = 0.0; vecs.append(e)            # g2 == 0
    dump_case(name, tpath, ppath, vecs, compress=False)
    dump_case(name + "_ccc", tpath, ppath, vecs[:1], compress=True)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    import shutil
    # the reference's only shipped data pair; small (57 KB), kept as fixture input
    shutil.copy(os.path.join(REF_TEST, "tree.nwk"), os.path.join(GOLD, "ref_test.nwk"))
    shutil.copy(os.path.join(REF_TEST, "pa_matrix.txt"), os.path.join(GOLD, "ref_test.pa.txt"))
    dump_case("ref_test", os.path.join(GOLD, "ref_test.nwk"), os.path.join(GOLD, "ref_test.pa.txt"), APPENDIX_C, False)
    dump_case("ref_test_ccc", os.path.join(GOLD, "ref_test.nwk"), os.path.join(GOLD, "ref_test.pa.txt"), APPENDIX_C[:3], True)
    synth("syn12", 12, 80, seed=11)
    synth("syn100", 100, 300, seed=12)
    synth("cat40", 40, 120, seed=13, caterpillar=True)
    synth("syn300", 300, 60, seed=14, nvec=2)
