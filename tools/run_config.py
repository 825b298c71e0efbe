#!/usr/bin/env python
"""Runs one synthetic BASELINE.json-style configuration through the C ABI on cuda:0 and prints a JSON line:
timings, achieved FP64 rate and size-independent sanity properties (finite values, permutation invariance,
shard partial sums adding up).  Example:  python tools/run_config.py --genomes 1000 --genes 20000 --params 16"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppmmodelpangenome_b200 import Inference, measure_fp64_peak, simulate  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=1000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--params", type=int, default=16)
    ap.add_argument("--caterpillar", action="store_true")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--checks", action="store_true")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    nwk = simulate.sim_tree(a.genomes, rng, caterpillar=a.caterpillar)
    names, M, truth = simulate.sim_genes(nwk, a.genes, rng)
    left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
    t_sim = time.time() - t0
    t0 = time.time()
    inf = Inference.from_arrays(left, right, bl, rows, dedupe=True, device=0)
    t_create = time.time() - t0
    P = np.stack([simulate.start_params(truth, a.genes, rng) for _ in range(a.params)])
    ll, i2 = inf.loglik_batch(P, return_i2=True)  # warm-up
    times, sweeps = [], []
    for _ in range(a.reps):
        t0 = time.time()
        ll, i2 = inf.loglik_batch(P, return_i2=True)
        times.append(time.time() - t0)
        sweeps.append(inf.counters()["sweep_ms"])
    c = inf.counters()
    n_unique = int(inf.nRep)
    sweep_s = float(np.median(sweeps)) * 1e-3
    out = dict(genomes=a.genomes, genes=a.genes, caterpillar=a.caterpillar, unique_patterns=n_unique, P=a.params,
               n_slices=int(inf.info.n_slices), n_terms=int(inf.info.n_terms), flops_per_eval=c["flops_per_eval"],
               sim_s=round(t_sim, 2), create_s=round(t_create, 2), call_s=float(np.median(times)), sweep_s=sweep_s,
               pattern_evals_per_s=n_unique * a.params / float(np.median(times)),
               tflops_algorithmic=n_unique * a.params * c["flops_per_eval"] / sweep_s / 1e12,
               tflops_fp64_issued=n_unique * a.params * c["fp64_instr_per_eval"] * 2 / sweep_s / 1e12,
               fp64_peak_tflops=measure_fp64_peak(0), ll_finite=bool(np.all(np.isfinite(ll))), i2_positive=int((i2 > 0).sum()),
               ll0=float(ll[0]), i20=float(i2[0]))
    if a.checks:
        perm = rng.permutation(len(rows))
        b = Inference.from_arrays(left, right, bl, rows[perm], dedupe=True, device=0)
        out["perm_invariance_rel"] = float(np.max(np.abs(b.loglik_batch(P[:2]) / ll[:2] - 1)))
        b.close()
        tot = np.zeros(2)
        for rk in range(4):
            s = Inference.from_arrays(left, right, bl, rows, dedupe=True, device=0, shard_rank=rk, shard_world=4)
            part, i2s = s.loglik_batch(P[:2], return_i2=True)
            tot += part
            s.close()
        tot = np.where(i2[:2] < 0, -1e9 + i2[:2], tot)
        out["shard4_sum_rel"] = float(np.max(np.abs(tot / ll[:2] - 1)))
        cat = inf.assignCat(P[0], i2[0]) if i2[0] > 0 else None
        out["cat_hist"] = None if cat is None else np.bincount(cat, minlength=3).tolist()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
