#!/usr/bin/env python
"""How does the sweep kernel scale with resident warps?  Runs a tree small enough for 16 warps of lineage slots to fit
in shared memory with PPM_MAX_WARPS = 4..16 (one subprocess each) and prints ms per sweep."""
import json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np
    from ppmmodelpangenome_b200 import Inference, simulate
    n, genes, P = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rng = np.random.default_rng(1)
    nwk = simulate.sim_tree(n, rng)
    names, M, truth = simulate.sim_genes(nwk, genes, rng)
    left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
    inf = Inference.from_arrays(left, right, bl, rows, dedupe=True, device=0)
    Pm = np.stack([simulate.start_params(truth, genes, rng) for _ in range(P)])
    for _ in range(3):
        inf.loglik_batch(Pm)
    ms = []
    for _ in range(5):
        inf.loglik_batch(Pm); ms.append(inf.counters()["sweep_ms"])
    c = inf.counters()
    print(json.dumps(dict(warps=os.environ.get("PPM_MAX_WARPS"), n=n, unique=int(inf.nRep), sweep_ms=float(np.median(ms)),
                          tflops_alg=inf.nRep * P * c["flops_per_eval"] / (np.median(ms) * 1e-3) / 1e12,
                          tflops_fp64=inf.nRep * P * c["fp64_instr_per_eval"] * 2 / (np.median(ms) * 1e-3) / 1e12)))
else:
    for n in (40, 100):
        for w in (4, 6, 8, 10, 11, 12, 14, 16):
            env = dict(os.environ, PPM_MAX_WARPS=str(w))
            r = subprocess.run([sys.executable, __file__, "child", str(n), "20000", "128"], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr[-300:])
