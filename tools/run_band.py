#!/usr/bin/env python
"""Times the objective on one synthetic configuration with and without the band kernel (one B200) and prints one JSON
line per variant.  Example: python tools/run_band.py --genomes 1000 --genes 20000 --params 32 --variants default,noband"""
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
    ap.add_argument("--params", type=int, default=32)
    ap.add_argument("--caterpillar", action="store_true")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--variants", default="default,noband")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    nwk = simulate.sim_tree(a.genomes, rng, caterpillar=a.caterpillar)
    names, M, truth = simulate.sim_genes(nwk, a.genes, rng)
    left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
    P = np.stack([simulate.start_params(truth, a.genes, rng) for _ in range(a.params)])
    peak = measure_fp64_peak(0)
    ref = None
    for v in a.variants.split(","):
        env = {}
        if v == "noband":
            env["PPM_NO_BAND"] = "1"
        elif v == "sparse":
            env["PPM_SPARSE"] = "1"
        elif v != "default":
            for kv in v.split("+"):  # e.g. R=2+G=4+CTAS=3
                k, val = kv.split("=")
                env["PPM_BAND_" + k] = val
        for k in list(os.environ):
            if k.startswith("PPM_BAND_") or k in ("PPM_NO_BAND", "PPM_SPARSE"):
                del os.environ[k]
        os.environ.update(env)
        t0 = time.time()
        inf = Inference.from_arrays(left, right, bl, rows, dedupe=True, device=0)
        t_create = time.time() - t0
        ll, i2 = inf.loglik_batch(P, return_i2=True)
        times, sweeps = [], []
        for _ in range(a.reps):
            t0 = time.time()
            ll, i2 = inf.loglik_batch(P, return_i2=True)
            times.append(time.time() - t0)
            sweeps.append(inf.counters()["sweep_ms"])
        c = inf.counters()
        n_unique = int(inf.nRep)
        sweep_s = float(np.median(sweeps)) * 1e-3
        if ref is None:
            ref = ll
        out = dict(variant=v, plan=inf.plan(), genomes=a.genomes, genes=a.genes, unique=n_unique, P=a.params,
                   create_s=round(t_create, 2), call_s=float(np.median(times)), sweep_s=sweep_s,
                   pattern_evals_per_s=n_unique * a.params / float(np.median(times)),
                   tflops_algorithmic=n_unique * a.params * c["flops_per_eval"] / sweep_s / 1e12,
                   tflops_fp64_issued=n_unique * a.params * c["fp64_instr_per_eval"] * 2 / sweep_s / 1e12,
                   peak=peak, frac=n_unique * a.params * c["flops_per_eval"] / sweep_s / 1e12 / peak,
                   max_rel_vs_first=float(np.max(np.abs(ll / ref - 1))), i2_pos=int((i2 > 0).sum()))
        print(json.dumps(out), flush=True)
        inf.close()


if __name__ == "__main__":
    main()
