#!/usr/bin/env python
"""bench.py -- gene-likelihood evaluations per second of the PPM log-likelihood (BASELINE.json metric), and wall-seconds
per full MLE fit.

A "step" = one call of the objective for a batch of P parameter vectors over every unique gene pattern of the workload.

PRIMARY workload, the same at every N (so that the driver's 1/2/4/8-GPU numbers are one strong-scaling curve):
  BASELINE.json configs[3] -- a simulated 5,000-genome ultrametric tree x 1,000,000 genes (C++ PPM simulator, seeded;
  ~679 k unique patterns), P = 1.  With N ranks (torchrun) the unique patterns are SHARDED over the ranks of ONE joint
  matrix (ppm_create(shard_rank, shard_world): global n_obs, global i2), every rank evaluates the same parameter vector
  and ONE all-reduce of the P partial log-likelihoods per step (NCCL via torch.distributed) joins them: fixed total work,
  "scaling": "strong".  The all-reduced objective is checked once against an unsharded evaluation.
SECONDARY (N = 1 only, same JSON line, key "secondary"): configs[1] (100 genomes x 20,000 genes, P = 256: the round-1
  headline, with its own roofline and end-to-end number), configs[2] (1,000 genomes x 200,000 genes, P = 1,024, full size),
  configs[4] (10,000-leaf caterpillar x 100,000 genes) and "fit": wall-seconds of full MLE fits through the CLI
  (shipped 20-genome data seed 1 next to the reference CLI; configs[1] data single start and 64 starts).
Inputs are generated once per box and cached under /tmp/ppm_bench_cache.

  python bench.py --gpus 1 --steps 20 --warmup 3           our arm  (CUDA through the C ABI)
  python bench.py --impl reference --steps 3 --warmup 1     reference arm (the reference's own CPU code, bounded sample)
  python bench.py --workload "configs[1]"                   one workload only, as the primary line
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gene-likelihood evals/s (patterns x param-sets/s)"
UNIT = "pattern-evals/s"
CACHE = os.environ.get("PPM_BENCH_CACHE", "/tmp/ppm_bench_cache")

WORKLOADS = {
    # name: genomes, genes, parameter vectors per step, caterpillar, generator
    "configs[1]": dict(genomes=100, genes=20000, P=256, caterpillar=False, gen="numpy", tree_seed=1, gene_seed=1000,
                       text="configs[1]: simulated 100-genome ultrametric tree x 20,000 genes (numpy PPM simulation, seed 1)"),
    "configs[2]": dict(genomes=1000, genes=200000, P=1024, caterpillar=False, gen="cxx", tree_seed=1, gene_seed=7,
                       text="configs[2]: simulated 1,000-genome ultrametric tree x 200,000 genes, 1,024 parameter vectors per launch"),
    "configs[3]": dict(genomes=5000, genes=1000000, P=1, caterpillar=False, gen="cxx", tree_seed=1, gene_seed=7,
                       text="configs[3]: simulated 5,000-genome ultrametric tree x 1,000,000 genes, unique patterns sharded over the GPUs"),
    "configs[4]": dict(genomes=10000, genes=100000, P=2, caterpillar=True, gen="cxx", tree_seed=1, gene_seed=7,
                       text="configs[4]: 10,000-genome caterpillar tree x 100,000 genes (deepest tree, underflow-scaling path)"),
}


# ----------------------------------------------------------------------------------------------- workloads
def _wait_for(path, timeout=1800):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise RuntimeError("timed out waiting for " + path)
        time.sleep(0.2)


def make_workload(name, rank=0, n_params=None):
    """Returns dict(nwk, left, right, bl, rows [genes][words], truth, params [P][8], names).  Rank 0 generates and caches
    (atomic rename); the other ranks wait for the cache file and map it."""
    from ppmmodelpangenome_b200 import simulate, simulate_genes
    w = WORKLOADS[name]
    P = n_params or w["P"]
    os.makedirs(CACHE, exist_ok=True)
    key = "%s_%d_%d_%d_%d_%s" % (w["gen"], w["genomes"], w["genes"], w["tree_seed"], w["gene_seed"], "cat" if w["caterpillar"] else "beta")
    base = os.path.join(CACHE, key)
    meta_p, rows_p = base + ".json", base + ".rows.npy"
    t0 = time.time()
    if rank == 0 and not (os.path.exists(meta_p) and os.path.exists(rows_p)):
        nwk = simulate.sim_tree(w["genomes"], np.random.default_rng(w["tree_seed"]), caterpillar=w["caterpillar"])
        if w["gen"] == "numpy":
            names, M, truth = simulate.sim_genes(nwk, w["genes"], np.random.default_rng(w["gene_seed"]))
            left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
        else:
            left, right, bl, labels = simulate.parse_newick(nwk)
            rows, truth = simulate_genes(left, right, bl, w["genes"], w["gene_seed"])
        tmp = base + ".%d.tmp.npy" % os.getpid()
        np.save(tmp, rows)
        os.replace(tmp, rows_p)
        with open(meta_p + ".tmp", "w") as f:
            json.dump({"nwk": nwk, "truth": truth}, f)
        os.replace(meta_p + ".tmp", meta_p)
    _wait_for(meta_p)
    with open(meta_p) as f:
        meta = json.load(f)
    rows = np.load(rows_p, mmap_mode="r")
    left, right, bl, labels = simulate.parse_newick(meta["nwk"])
    names = [labels[v] for v in range(len(left)) if left[v] < 0]  # leaves in preorder = bit order of the rows
    rng_p = np.random.default_rng(77)
    params = np.stack([simulate.start_params(meta["truth"], w["genes"], rng_p) for _ in range(P)])
    return dict(name=name, nwk=meta["nwk"], left=left, right=right, bl=bl, rows=rows, truth=meta["truth"], params=params,
                names=names, gen_s=time.time() - t0, spec=w)


def unpack_genes(rows, names, sel):
    """M[genome][gene] (uint8) of the selected genes from the bit-packed rows (bit k = k-th leaf in preorder)."""
    sub = np.ascontiguousarray(rows[sel])
    n = len(names)
    M = np.zeros((n, len(sel)), np.uint8)
    for k in range(n):
        M[k] = (sub[:, k >> 5] >> np.uint32(k & 31)) & 1
    return M


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------- CPU reference
def cpu_reference_time(wl, sample_genes, repeats):
    """The reference's own implementation (oracle/_ref, unmodified sources) -- or, if that library did not travel, the C
    port -- timed on the host cores on a bounded sample of the workload's genes, deduplicated ('ccc' input, the reference's
    best case).  Times exactly what the objective does: clearLikValues()+logLik() (functions.cpp:406-409)."""
    from ppmmodelpangenome_b200 import simulate
    n_genes = wl["rows"].shape[0]
    sel = np.sort(np.random.default_rng(5).choice(n_genes, min(sample_genes, n_genes), replace=False))
    sub = unpack_genes(wl["rows"], wl["names"], sel)
    params = np.array(wl["params"], dtype=np.float64, copy=True)
    params[:, 0] *= len(sel) / n_genes  # N0 and i1 are gene counts/rates: scale them to the sample
    params[:, 2] *= len(sel) / n_genes
    uniq, counts = np.unique(sub.T, axis=0, return_counts=True)
    tmp = tempfile.mkdtemp(prefix="ppm_bench_")
    tp, pp = os.path.join(tmp, "tree.nwk"), os.path.join(tmp, "pa.txt")
    with open(tp, "w") as f:
        f.write(wl["nwk"] + "\n")
    simulate.write_matrix(pp, wl["names"], np.ascontiguousarray(uniq.T), counts)
    from oracle import ref as oref
    times = []
    if oref.available():
        kind = "reference"
        oref.set_threads(len(os.sched_getaffinity(0)))  # all host threads, whatever OMP_NUM_THREADS torchrun set
        cores = oref.max_threads()
        R = oref.Reference(tp, pp)
        for k in range(repeats):
            t0 = time.perf_counter()
            R.loglik(params[k % len(params)])
            times.append(time.perf_counter() - t0)
        R.close()
    else:
        kind = "port"
        from oracle import cport, model
        cport.set_threads(len(os.sched_getaffinity(0)))
        cores = cport.max_threads()
        orc = cport.Oracle(model.model_from_files(tp, pp))
        for k in range(repeats):
            t0 = time.perf_counter()
            orc.loglik(params[k % len(params)])
            times.append(time.perf_counter() - t0)
    return len(uniq), times, kind, cores, len(sel)


def reference_sample_size(name):
    # about 3-10 s of CPU work per call on 16 threads: ~85-140 ns per (slice x lineage) term and core
    return {"configs[1]": 2000, "configs[2]": 400, "configs[3]": 32, "configs[4]": 16}[name]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload, 0, max(args.steps + args.warmup, 1))
    sample = reference_sample_size(args.workload)
    if args.warmup > 0:
        cpu_reference_time(wl, sample, 1)
    n_unique, times, kind, cores, n_sel = cpu_reference_time(wl, sample, args.steps)
    total = float(np.sum(times))
    value = n_unique * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["spec"]["text"],
                       "step": "one clearLikValues()+logLik() call (1 parameter vector) on a bounded sample of the workload's genes"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d random genes of the workload = %d unique patterns, 1 parameter vector per step" % (n_sel, n_unique)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
class Dist:
    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t)
        return float(t.item())


def kernel_name(inf):
    pl = inf.plan()
    if pl["last_kind"] == 1 and pl.get("band_stream"):
        # kernel_ms covers the whole band sequence of a step: per batch prune -> integral -> mixture (the integral is ~98 % of it)
        return "ppm::band_integral_kernel<%d,%s> (+ ppm::band_prune_kernel, ppm::band_mixture_kernel per batch)" % (
            pl["band_R"], "true" if pl.get("band_lp") else "false"), pl
    if pl["last_kind"] == 1:
        return "ppm::band_kernel<%d,%s>" % (pl["band_R"], "true" if pl["band_G"] > 1 else "false"), pl
    mode = pl["last_mode"]
    hot = "true" if (pl["lean"] and pl.get("sweep_lp") and mode != 2 and not pl["rows_in_place"]) else "false"
    return "ppm::sweep_kernel<8,1,%d,false,%s,%s>" % (mode, "true" if pl["rows_in_place"] else "false", hot), pl


def load_traffic():
    """dram bytes per launch of the dominant kernel by workload, from the committed ncu summaries (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def time_workload(D, wl, K, W, peak_tf, e2e_reps=3, check_unsharded=False, sampler=None):
    """Times K steps of the objective on workload `wl` (device-resident: CUDA events on the launching stream, max over ranks)
    and the end-to-end path through ppm_loglik with host buffers.  Returns the fields of a bench line."""
    torch = D.torch
    from ppmmodelpangenome_b200 import Inference
    P = len(wl["params"])
    params = np.ascontiguousarray(wl["params"])
    t0 = time.perf_counter()
    inf = Inference.from_arrays(wl["left"], wl["right"], wl["bl"], wl["rows"], dedupe=True, device=D.local,
                                shard_rank=D.rank, shard_world=D.world)
    create_s = time.perf_counter() - t0
    n_unique_total, n_genes = int(inf.nRep), int(inf.nGenes)
    n_shard = int(inf.info.shard_count)

    stream = torch.cuda.Stream()  # a real (non-default) stream: kernels, events and NCCL all go here
    torch.cuda.set_stream(stream)
    sp = ctypes.c_void_p(stream.cuda_stream)
    d_params = torch.from_numpy(params).cuda()
    d_ll = torch.zeros(P, dtype=torch.float64, device="cuda")
    d_i2 = torch.zeros(P, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_device():
        inf.loglik_device(d_params.data_ptr(), P, d_ll.data_ptr(), d_i2.data_ptr(), sp)
        if D.dist:
            D.dist.all_reduce(d_ll)  # the one collective of the path: P partial log-likelihoods
        inf.finalize_device(d_i2.data_ptr(), d_ll.data_ptr(), P, sp)

    for _ in range(W):
        step_device()
    D.barrier()
    if sampler is not None:
        sampler.rows.clear()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sweep_ms = []
    D.barrier()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()  # L2 flush between timed iterations (not timed)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        if k == K - 1 or K <= 64:
            torch.cuda.synchronize()
            sweep_ms.append(inf.counters()["sweep_ms"])
    D.barrier()
    t_wall = time.perf_counter() - t_wall0
    total_ms = D.max(float(np.sum([a.elapsed_time(b) for a, b in ev])))
    cnt = inf.counters()
    kname, plan = kernel_name(inf)
    pattern_evals = float(n_unique_total) * P      # whole job: every unique pattern x every parameter vector, once
    value = pattern_evals * K / (total_ms * 1e-3)

    # end to end through the public API with HOST buffers (params in, ll+i2 out), copies inside the timed region
    ll_host = np.zeros(P)
    t_e2e = []
    D.barrier()
    for k in range(e2e_reps):
        flush.zero_()
        torch.cuda.synchronize()
        if D.dist:
            D.dist.barrier()
        t0 = time.perf_counter()
        part, i2h = inf.loglik_batch(params, return_i2=True)  # ppm_loglik: H2D params, kernels, D2H ll and i2
        if D.dist:
            tt = torch.from_numpy(part).cuda()
            D.dist.all_reduce(tt)
            part = tt.cpu().numpy()
            part = np.where(i2h < 0, -1e9 + i2h, part)
        ll_host = part
        t_e2e.append(time.perf_counter() - t0)
    e2e_value = pattern_evals / D.max(float(np.median(t_e2e)))
    assert np.allclose(d_ll.cpu().numpy(), ll_host, rtol=1e-12, atol=0), "device/host path mismatch"
    checked = None
    if check_unsharded and D.world > 1:
        # once: the all-reduced objective of the sharded job equals an unsharded evaluation of the same joint matrix
        if D.rank == 0:
            whole = Inference.from_arrays(wl["left"], wl["right"], wl["bl"], wl["rows"], dedupe=True, device=D.local)
            ref = whole.loglik_batch(params[:1])
            whole.close()
            checked = float(abs(ll_host[0] / ref[0] - 1))
            assert checked <= 1e-12, ("sharded objective differs from the unsharded one", ll_host[0], ref[0])
        D.barrier()

    F = cnt["flops_per_eval"]  # 5W + 28(n-1) + 2S algorithmic flops per pattern-eval
    sweep = float(np.median(sweep_ms)) if sweep_ms else None
    ach = n_shard * P * F / (sweep * 1e-3) / 1e12 if sweep else None
    pipe = n_shard * P * cnt["fp64_instr_per_eval"] * 2 / (sweep * 1e-3) / 1e12 if sweep else None
    traffic = load_traffic().get(wl["name"]) if D.world == 1 else None
    out = {
        "value": value, "ms_per_step": total_ms / K, "steps": K, "warmup": W,
        "config": {"workload": wl["spec"]["text"], "param_vectors_per_step": P, "unique_patterns": n_unique_total,
                   "unique_patterns_this_rank": n_shard, "genes": n_genes, "genomes": wl["spec"]["genomes"],
                   "parallelism": ("unique patterns of one joint matrix sharded over %d GPUs (ppm_create shard_rank/shard_world), one all-reduce of %d double(s) per step"
                                   % (D.world, P)) if D.world > 1 else "1 GPU",
                   "l2": "flushed between timed steps (256 MB memset); each step timed with CUDA events on the launching stream"},
        "gene_evals_per_s": float(n_genes) * P * K / (total_ms * 1e-3),
        "wall_s_timed_region": t_wall, "create_s": create_s, "generate_or_load_s": wl["gen_s"],
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(P * 8 * 8), "d2h_bytes_per_step": int(P * 8 * 2)},
        "gpu_launches": int(cnt["launches"] + 1) * K,  # kernels of ppm_loglik_device + finalize
        "roofline": {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": (ach / peak_tf) if ach else None,
                     "traffic": traffic, "kernel": kname, "kernel_ms": sweep, "plan": plan,
                     "peak_source": "DFMA micro-benchmark measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry)",
                     "algorithmic_flops_per_pattern_eval": F, "fp64_pipe_tflops_issued": pipe,
                     "fp64_pipe_frac": (pipe / peak_tf) if pipe else None,
                     "frac_note": "achieved = ALGORITHMIC flops (SURVEY 8d: F = 5W + 28(n-1) + 2S per pattern-eval, the reference's dense "
                                  "formulation) / kernel time; frac > 1 means the kernels do less arithmetic than F counts (leaf powers: "
                                  "one exp2 per slice instead of one term per alive leaf) -- fp64_pipe_frac (modelled FP64 instructions "
                                  "issued x 2 / peak) and ncu sm__inst_executed_pipe_fp64 in profiles/ are the pipe utilisation"},
    }
    if checked is not None:
        out["sharded_vs_unsharded_rel"] = checked
    if not plan.get("band_stream") and wl["spec"]["genomes"] >= 256:
        # below 1,200 genomes the dense evaluation stays on the lane-per-pattern sweep; a handle created with PPM_SPARSE=1 takes
        # the band kernels (they carry the sparse evaluation) from 256 genomes
        inf.close()
        os.environ["PPM_SPARSE"] = "1"
        try:
            inf = Inference.from_arrays(wl["left"], wl["right"], wl["bl"], wl["rows"], dedupe=True, device=D.local,
                                        shard_rank=D.rank, shard_world=D.world)
        finally:
            del os.environ["PPM_SPARSE"]
        inf.set_sparse(False)
        plan = inf.plan()
    if plan.get("band_stream"):
        # SEPARATE metric: the same objective with ppm_set_sparse -- dirty lineages only, as a ratio against the
        # all-zero row; the dense evaluation above stays the roofline contract.  Same inputs, same timing method.
        dense_ll = d_ll.cpu().numpy().copy()
        inf.set_sparse(True)
        t0 = time.perf_counter()
        step_device()  # (the first call builds the pattern-only term lists, once per handle)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        step_device()
        Ks = max(1, min(K, 5))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Ks)]
        D.barrier()
        for k in range(Ks):
            flush.zero_()
            evs[k][0].record(stream)
            step_device()
            evs[k][1].record(stream)
        D.barrier()
        sp_ms = D.max(float(np.sum([a.elapsed_time(b) for a, b in evs]))) / Ks
        sparse_ll = d_ll.cpu().numpy()
        out["sparse"] = {"value": pattern_evals / (sp_ms * 1e-3), "unit": UNIT, "ms_per_step": sp_ms, "steps": Ks,
                         "speedup_vs_dense": (total_ms / K) / sp_ms, "first_call_s": build_s,
                         "max_rel_vs_dense": float(np.max(np.abs(sparse_ll / dense_ll - 1))),
                         "note": "ppm_set_sparse(1): patterns present (absent) in <= 1/8 of the genomes are evaluated from term lists of their dirty "
                                 "lineages against the all-zero (all-ones) row (band_sparse_kernel); the others stay dense; not a roofline number"}
        inf.set_sparse(False)
    inf.close()
    del flush, d_params, d_ll, d_i2
    torch.cuda.empty_cache()
    return out


def cpu_baseline_block(wl, repeats=3):
    try:
        nu, times, kind, cores, n_sel = cpu_reference_time(wl, reference_sample_size(wl["name"]), repeats)
        return {"value": nu * len(times) / float(np.sum(times)), "unit": UNIT, "cores": cores, "kind": kind,
                "sample": "%d random genes of the workload = %d unique patterns, 1 parameter vector per call, %d calls" % (n_sel, nu, len(times))}
    except Exception as e:  # the baseline is a report, never a reason to lose the bench line
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}


def fit_block():
    """Wall-seconds per full MLE fit through the drop-in CLI (process start and CUDA initialisation included), next to the
    reference CLI where it travelled (oracle/_ref/inference_ref): shipped 20-genome data, seed 1; and on the configs[1]
    data one start and 64 starts (PPM_STARTS)."""
    from ppmmodelpangenome_b200 import simulate
    out = {}
    cli = os.path.join(ROOT, "ppmmodelpangenome_b200", "inference")
    g = os.path.join(ROOT, "tests", "golden")
    tmp = tempfile.mkdtemp(prefix="ppm_fit_")

    def run(exe, seed, tree, pa, env=None, tag="x"):
        e = dict(os.environ)
        e.update(env or {})
        mle, cat = os.path.join(tmp, tag + ".mle.txt"), os.path.join(tmp, tag + ".cat.txt")
        t0 = time.perf_counter()
        r = subprocess.run([exe, str(seed), tree, pa, mle, cat], capture_output=True, text=True, env=e, timeout=900)
        dt = time.perf_counter() - t0
        if r.returncode:
            raise RuntimeError(r.stderr[-400:])
        lines = open(mle).read().strip().splitlines()
        evals = [int(x.split(":")[1]) for x in r.stdout.splitlines() if x.startswith("Number of function evaluation")]
        return dt, lines, evals

    try:
        tree, pa = os.path.join(g, "ref_test.nwk"), os.path.join(g, "ref_test.pa.txt")
        # process start + CUDA initialisation dominate this fit (0.3 s of kernels) and vary with the driver's state on the box
        # (0.5 s warm, 2.5-3.7 s when the previous client has just released the GPU): three launches, the median is reported
        runs = [run(cli, 1, tree, pa, tag="a") for _ in range(3)]
        dt, lines, evals = sorted(runs, key=lambda r: r[0])[1]
        out["configs[0]_seed1"] = {"fit_wall_s": dt, "fit_wall_s_runs": [r[0] for r in runs], "evaluations": evals[0],
                                   "loglik": float(lines[-1].split()[-1])}
        ref_cli = os.path.join(ROOT, "oracle", "_ref", "inference_ref")
        if os.path.exists(ref_cli):
            dtr, lr, er = run(ref_cli, 1, tree, pa, env={"OMP_NUM_THREADS": str(len(os.sched_getaffinity(0)))}, tag="r")
            out["configs[0]_seed1"].update({"reference_fit_wall_s": dtr, "reference_evaluations": er[0],
                                            "reference_loglik": float(lr[-1].split()[-1]), "reference_threads": len(os.sched_getaffinity(0))})
        wl = make_workload("configs[1]", 0, 1)
        tp, pp = os.path.join(tmp, "c1.nwk"), os.path.join(tmp, "c1.pa.txt")
        with open(tp, "w") as f:
            f.write(wl["nwk"] + "\n")
        simulate.write_matrix(pp, wl["names"], unpack_genes(wl["rows"], wl["names"], np.arange(wl["rows"].shape[0])))
        dt1, l1, e1 = run(cli, 1, tp, pp, tag="b")
        out["configs[1]_seed1"] = {"fit_wall_s": dt1, "evaluations": e1[0], "loglik": float(l1[-1].split()[-1])}
        dt64, l64, e64 = run(cli, 1, tp, pp, env={"PPM_STARTS": "64"}, tag="c")
        out["configs[1]_64_starts"] = {"fit_wall_s": dt64, "starts": len(l64), "evaluations_total": int(np.sum(e64)),
                                       "best_loglik": max(float(x.split()[-1]) for x in l64),
                                       "seed1_line_identical_to_single_start": l64[0] == l1[-1]}
    except Exception as e:
        out["error"] = repr(e)[:300]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="configs[3]", choices=sorted(WORKLOADS))
    ap.add_argument("--params", type=int, default=0, help="parameter vectors per step (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    from ppmmodelpangenome_b200 import measure_fp64_peak
    D = Dist()
    if D.world != args.gpus and D.world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    K, W = args.steps, max(args.warmup, 3)
    wl = make_workload(args.workload, D.rank, args.params or None)
    peak_tf = measure_fp64_peak(D.local)  # DFMA micro-benchmark on this very GPU
    sampler = ClockSampler(D.local)
    sampler.start()
    t_wait = time.perf_counter()
    while not sampler.rows and time.perf_counter() - t_wait < 5.0:
        time.sleep(0.01)  # nvidia-smi is up before the timed region starts
    res = time_workload(D, wl, K, W, peak_tf, check_unsharded=True, sampler=sampler)
    sampler.stop_flag = True  # sampled through the device-resident AND the end-to-end timed loops

    if D.rank == 0:
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": D.world, "steps": K, "warmup": W,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic"}
        for k in ("config", "gene_evals_per_s", "wall_s_timed_region", "create_s", "generate_or_load_s", "e2e", "gpu_launches", "roofline"):
            line[k] = res[k]
        for k in ("sharded_vs_unsharded_rel", "sparse"):
            if k in res:
                line[k] = res[k]
        line["clocks"] = sampler.summary()
        if D.world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(wl)
        if D.world == 1 and not args.no_secondary and args.workload == "configs[3]":
            sec = []
            for name, k2, w2 in (("configs[1]", min(K, 20), 3), ("configs[2]", 2, 1), ("configs[4]", 2, 1)):
                try:
                    w = make_workload(name, 0)
                    r = time_workload(D, w, k2, w2, peak_tf, e2e_reps=1 if name != "configs[1]" else 3)
                    r["workload"] = name
                    r["metric"], r["unit"] = METRIC, UNIT
                    sec.append(r)
                except Exception as e:
                    sec.append({"workload": name, "error": repr(e)[:300]})
            line["secondary"] = sec
            line["fit"] = fit_block()
        print(json.dumps(line))
    if D.dist:
        D.dist.destroy_process_group()


if __name__ == "__main__":
    main()
