#!/usr/bin/env python
"""Full MLE fit (the CLI, seed 1) on the bench workload (BASELINE.json configs[1]: 100 genomes x 20,000 genes) and,
for the same inputs, the reference's time per objective evaluation on the host cores (from which the reference's
fit time follows: its optimiser is serial in the evaluations).  Prints one JSON line."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from ppmmodelpangenome_b200 import simulate

import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=0, help="0 = the bench workload (configs[1]); else a seeded simulation of this size")
ap.add_argument("--genes", type=int, default=20000)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--no-reference", action="store_true")
args = ap.parse_args()
if args.genomes:
    rng = np.random.default_rng(3)
    nwk = simulate.sim_tree(args.genomes, rng)
    names, M, truth = simulate.sim_genes(nwk, args.genes, rng)
    params = np.stack([simulate.start_params(truth, args.genes, rng) for _ in range(4)])
    label = "%d genomes x %d genes (seed 3)" % (args.genomes, args.genes)
else:
    nwk, names, M, truth, params = bench.make_workload(0, 4)
    label = "configs[1] 100 genomes x 20000 genes"
tmp = tempfile.mkdtemp(prefix="ppm_fit_")
tp, pp = os.path.join(tmp, "tree.nwk"), os.path.join(tmp, "pa.txt")
open(tp, "w").write(nwk + "\n")
simulate.write_matrix(pp, names, M)
cli = os.path.join(ROOT, "ppmmodelpangenome_b200", "inference")
t0 = time.time()
r = subprocess.run([cli, "1", tp, pp, os.path.join(tmp, "mle.txt"), os.path.join(tmp, "cat.txt")], capture_output=True, text=True,
                   env=dict(os.environ, PPM_VERBOSE="1", PPM_GPUS=str(args.gpus)))
wall = time.time() - t0
out = r.stdout.splitlines()
res = dict(workload=label, gpus=args.gpus, cli_wall_s=wall, returncode=r.returncode,
           evaluations=[l for l in out if l.startswith("Number of function evaluation")],
           message=[l for l in out if l.startswith("Message")], loglik=[l for l in out if l.startswith("logLik value")],
           found=[l for l in out if l.startswith("Found minimum")], launches=r.stderr.strip().splitlines()[-1:] ,
           mle_param=open(os.path.join(tmp, "mle.txt")).read().strip() if r.returncode == 0 else None,
           truth={k: float(v) for k, v in truth.items()})
try:
    from oracle import ref as oref
    if oref.available() and not args.no_reference:
        uniq, counts = np.unique(M.T, axis=0, return_counts=True)
        cp = os.path.join(tmp, "pa_ccc.txt")
        simulate.write_matrix(cp, names, np.ascontiguousarray(uniq.T), counts)
        oref.set_threads(len(os.sched_getaffinity(0)))
        R = oref.Reference(tp, cp)
        ts = []
        for k in range(3):
            t0 = time.time(); R.loglik(params[k]); ts.append(time.time() - t0)
        res["reference_s_per_evaluation"] = float(np.median(ts))
        res["reference_threads"] = oref.max_threads()
except Exception as e:
    res["reference_error"] = repr(e)
print(json.dumps(res))
