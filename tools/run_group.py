#!/usr/bin/env python
"""One process, several GPUs: times ppm_loglik on a GROUP handle (ppm_create_group: one pattern shard per GPU, partial
sums gathered over NVLink peer copies) against the unsharded handle on GPU 0, on one synthetic configuration.
Example:  python tools/run_group.py --genomes 1000 --genes 100000 --params 32 --gpus 1,2,4,8"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppmmodelpangenome_b200 import Inference, simulate  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=1000)
    ap.add_argument("--genes", type=int, default=100000)
    ap.add_argument("--params", type=int, default=32)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    nwk = simulate.sim_tree(a.genomes, rng)
    names, M, truth = simulate.sim_genes(nwk, a.genes, rng)
    left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
    P = np.stack([simulate.start_params(truth, a.genes, rng) for _ in range(a.params)])
    base = None
    for n in [int(x) for x in a.gpus.split(",")]:
        inf = Inference.from_arrays(left, right, bl, rows, dedupe=True, devices=list(range(n))) if n > 1 else \
            Inference.from_arrays(left, right, bl, rows, dedupe=True, device=0)
        ll = inf.loglik_batch(P)  # warm-up
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            ll = inf.loglik_batch(P)
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        if base is None:
            base = (t, ll.copy())
        out = dict(genomes=a.genomes, genes=a.genes, unique_patterns=int(inf.nRep), P=a.params, gpus=n, call_s=t,
                   pattern_evals_per_s=int(inf.nRep) * a.params / t, speedup_vs_first=base[0] / t,
                   max_rel_diff_vs_first=float(np.max(np.abs(ll / base[1] - 1))))
        print(json.dumps(out), flush=True)
        inf.close()


if __name__ == "__main__":
    main()
