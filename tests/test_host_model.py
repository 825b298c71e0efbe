"""CPU tier: the C-ABI library loads, exports every symbol include/ppm_b200.h declares, and its HOST side
(Newick -> flat arrays, matrix -> bit-packed deduplicated patterns, mrca/freq/counts, integral grid) reproduces
the reference's integer fields bit-exactly.  No compute call is made here (there is no CPU compute path)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import CASES, ROOT
import ppmmodelpangenome_b200 as ppm
from ppmmodelpangenome_b200 import simulate


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ppm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ppm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 14
    lib = ppm.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(declared) == sorted(ppm.engine.SYMBOLS)


def _host(g, dedupe, **kw):
    return ppm.Inference(g.tree, g.pa, dedupe=dedupe, device=-1, **kw)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dedupe", [False, True])
def test_host_model_matches_reference_fields(golden, name, dedupe):
    g = golden(name + "_ccc" if dedupe else name)
    inf = _host(g, dedupe)
    meta, z = g.meta, g.z
    assert (inf.nRep, inf.nObs, inf.nNodes, inf.nLeaves) == (meta["nRep"], meta["nObs"], meta["nNodes"], meta["nLeaves"])
    assert inf.H == meta["H"] and inf.L == meta["L"] and inf.meanGenes == meta["meanGenes"]
    assert np.array_equal(inf.counts, z["counts"])
    assert np.array_equal(inf.mrca, z["mrca"])
    assert np.array_equal(inf.freq, z["freq"])
    assert np.array_equal(inf.height, z["height"]) and np.array_equal(inf.branchLength, z["bl"])
    nl = inf.nLeaves
    assert np.array_equal(inf.order[nl:], z["order"][nl:])
    assert np.array_equal(inf.alive[nl - 1:], z["subtree_sizes"][nl - 1:])
    # W = sum nSub * alive (functions.cpp:250-262) agrees with the python restatement
    from oracle import model
    m = model.model_from_files(g.tree, g.pa, dedupe=dedupe)
    h = m["height"][m["order"]]
    w = np.diff(h)
    nsub = np.where(w > 0, np.maximum(1, np.ceil(w * 100 / m["H"])), 0).astype(int)
    nsub[:nl - 1] = 0
    assert np.array_equal(inf.nsub, nsub)
    assert inf.info.n_slices == nsub.sum() and inf.info.n_terms == int((nsub * np.diff(m["sub_off"])).sum())
    inf.close()


def test_flat_array_entry_point_and_dedupe_scatter(golden):
    g = golden("syn100")
    from oracle import model
    names, M, counts = model.read_matrix(g.pa)
    nwk = open(g.tree).readline()
    left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
    a = ppm.Inference.from_arrays(left, right, bl, rows, dedupe=True, device=-1)
    b = _host(g, True)
    assert a.nRep == b.nRep and np.array_equal(a.counts, b.counts) and np.array_equal(a.mrca, b.mrca)
    g2p = a.gene_to_pattern
    assert len(g2p) == M.shape[1] and np.array_equal(np.bincount(g2p), a.counts)
    # rows mapped to one pattern are identical, first occurrence order (objects.cpp:254-273)
    first = {}
    for gene, p in enumerate(g2p):
        first.setdefault(int(p), gene)
        assert np.array_equal(rows[gene], rows[first[int(p)]])
    assert list(first.values()) == sorted(first.values())
    a.close(); b.close()


def test_shard_partition_covers_all_patterns(golden):
    g = golden("cat40")
    for world in (1, 2, 3, 8):
        spans = []
        for r in range(world):
            inf = _host(g, True, shard_rank=r, shard_world=world)
            spans.append((inf.info.shard_first, inf.info.shard_count))
            n = inf.nRep
            inf.close()
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (a, c), (b, _) in zip(spans, spans[1:]):
            assert a + c == b
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_parser_edge_cases(tmp_path):
    pa = tmp_path / "m.txt"
    pa.write_text("g1,g2,g3\nA,1,0,1\nB,0,1,1\nC,0,0,1\n")

    def mk(text, **kw):
        t = tmp_path / "t.nwk"
        t.write_text(text)
        return ppm.Inference(str(t), str(pa), device=-1, **kw)

    inf = mk("((A:1,B:1)lab:0.5,C:1.5):0;\nsecond line ignored\n")   # internal label skipped, first line only
    assert inf.nNodes == 5 and list(inf.freq) == [1, 1, 3] and list(inf.mrca) == [2, 3, 0]
    assert inf.H == 1.5 and inf.L == 4.0
    inf.close()
    inf = mk("(C:1.5,(A:1,B:1):0.5);")                           # no root length, other leaf order
    assert list(inf.mrca) == [3, 4, 0]
    inf.close()
    for bad in ["((A:1,B:1,C:1):1);", "((A:1,B:1):0.5,C);", "((A:1,B:1):0.5,D:1.5);", "(A:1,(B:1,C:1):0.5"]:
        with pytest.raises(ppm.PPMError):
            mk(bad)
    with pytest.raises(ppm.PPMError) as e:
        ppm.Inference(str(tmp_path / "nope.nwk"), str(pa), device=-1)
    assert e.value.code == 2  # PPM_EIO
    pa.write_text("g1,g2\nA,1,2\nB,0,1\nC,0,0\n")
    with pytest.raises(ppm.PPMError):
        mk("((A:1,B:1):0.5,C:1.5);")
    # ccc header carries multiplicities (objects.cpp:175-187)
    pa.write_text("ccc,5,2\nA,1,0\nB,0,1\nC,0,0\n")
    inf = mk("((A:1,B:1):0.5,C:1.5);", dedupe=False)
    assert list(inf.counts) == [5, 2] and inf.nObs == 7
    inf.close()
    # host-only handles refuse to compute; there is no CPU fallback
    inf = mk("((A:1,B:1):0.5,C:1.5);")
    with pytest.raises(ppm.PPMError) as e:
        inf.logLik([1, .1, 1, 1, 1, 1, .01, .01])
    assert e.value.code == 5
    inf.close()


def test_deep_caterpillar_parses_without_recursion():
    rng = np.random.default_rng(3)
    nwk = simulate.sim_tree(3000, rng, caterpillar=True)
    left, right, bl, _ = simulate.parse_newick(nwk)
    rows = np.zeros((2, (3000 + 31) // 32), np.uint32)
    rows[0, 0] = 1; rows[1, -1] = 1 << ((3000 - 1) % 32); rows[1, 0] = 1
    inf = ppm.Inference.from_arrays(left, right, bl, rows, device=-1)
    assert inf.nLeaves == 3000 and inf.info.n_slices >= 2999
    assert inf.mrca[1] == 0 or inf.freq[1] == 2
    inf.close()


def test_ccc_writer_round_trip_on_the_host(golden, tmp_path):
    """ppm_write_ccc: unique patterns + counts in the reference's compressed format (objects.cpp:175-187); reading the file
    back (dedupe off: the header carries the multiplicities) gives the same model, and the oracle's own reader agrees."""
    g = golden("syn100")
    a = ppm.Inference(g.tree, g.pa, dedupe=True, device=-1)
    out = str(tmp_path / "compressed.txt")
    a.write_ccc(out)
    first = open(out).readline().strip().split(",")
    assert first[0] == "ccc" and [int(x) for x in first[1:]] == list(a.counts)
    b = ppm.Inference(g.tree, out, dedupe=False, device=-1)
    assert (b.nRep, b.nObs, b.meanGenes) == (a.nRep, a.nObs, a.meanGenes)
    assert np.array_equal(a.counts, b.counts) and np.array_equal(a.mrca, b.mrca) and np.array_equal(a.freq, b.freq)
    from oracle import model
    names, M, counts = model.read_matrix(out)
    assert M.shape[1] == a.nRep and list(counts) == list(a.counts)
    a.close(); b.close()


def test_matrix_reader_packed_fast_path_and_general_path_agree(tmp_path):
    """The reader packs "v,v,v,v," eight text bytes at a time and transposes 32 x 32 bit tiles; rows it cannot take that way
    (leading zeros, CRLF) go through strtol.  Sizes that are multiples of neither 4 nor 32, genome lines in shuffled order."""
    rng = np.random.default_rng(3)
    n, genes = 37, 1003
    from ppmmodelpangenome_b200 import simulate
    nwk = simulate.sim_tree(n, rng)
    names = ["g%d" % (k + 1) for k in range(n)]
    left, right, bl, labels = simulate.parse_newick(nwk)
    leaves = [labels[v] for v in range(len(left)) if left[v] < 0]
    M = (rng.random((n, genes)) < 0.3).astype(np.uint8)
    M[:, M.sum(axis=0) == 0] = 1
    order = rng.permutation(n)
    tp, pp = tmp_path / "t.nwk", tmp_path / "m.txt"
    tp.write_text(nwk + "\n")
    with open(pp, "w") as f:
        f.write("genome," + ",".join("c%d" % j for j in range(genes)) + "\n")
        for k in order:
            vals = [str(v) for v in M[k]]
            if k % 5 == 0:
                vals[7] = "0" + vals[7]            # a leading zero: the general path takes over mid-line
            f.write(leaves[k] + "," + ",".join(vals) + ("\r\n" if k % 3 == 0 else "\n"))
    inf = ppm.Inference(str(tp), str(pp), dedupe=False, device=-1)
    assert inf.nRep == genes and inf.nLeaves == n
    assert np.array_equal(inf.freq, M.sum(axis=0))
    uniq = ppm.Inference(str(tp), str(pp), dedupe=True, device=-1)
    cols = {}
    for j in range(genes):
        cols.setdefault(M[:, j].tobytes(), len(cols))
    assert uniq.nRep == len(cols)
    assert np.array_equal(uniq.gene_to_pattern, [cols[M[:, j].tobytes()] for j in range(genes)])
    inf.close(); uniq.close()
