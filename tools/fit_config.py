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

nwk, names, M, truth, params = bench.make_workload(0, 4)
tmp = tempfile.mkdtemp(prefix="ppm_fit_")
tp, pp = os.path.join(tmp, "tree.nwk"), os.path.join(tmp, "pa.txt")
open(tp, "w").write(nwk + "\n")
simulate.write_matrix(pp, names, M)
cli = os.path.join(ROOT, "ppmmodelpangenome_b200", "inference")
t0 = time.time()
r = subprocess.run([cli, "1", tp, pp, os.path.join(tmp, "mle.txt"), os.path.join(tmp, "cat.txt")], capture_output=True, text=True,
                   env=dict(os.environ, PPM_VERBOSE="1"))
wall = time.time() - t0
out = r.stdout.splitlines()
res = dict(workload="configs[1] 100 genomes x 20000 genes", cli_wall_s=wall, returncode=r.returncode,
           evaluations=[l for l in out if l.startswith("Number of function evaluation")],
           message=[l for l in out if l.startswith("Message")], loglik=[l for l in out if l.startswith("logLik value")],
           found=[l for l in out if l.startswith("Found minimum")], launches=r.stderr.strip().splitlines()[-1:] ,
           mle_param=open(os.path.join(tmp, "mle.txt")).read().strip() if r.returncode == 0 else None,
           truth={k: float(v) for k, v in truth.items()})
try:
    from oracle import ref as oref
    if oref.available():
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
