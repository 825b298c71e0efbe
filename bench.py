#!/usr/bin/env python
"""bench.py -- gene-likelihood evaluations per second of the PPM log-likelihood (BASELINE.json metric).

A "step" = one call of the objective for a batch of P parameter vectors over every unique gene pattern of the
workload (what one launch of the multi-start / speculative Nelder-Mead driver asks for).  Workload at N=1 is
BASELINE.json configs[1]: a simulated 100-genome ultrametric tree x 20,000 genes (seeded PPM simulation,
ppmmodelpangenome_b200/simulate.py).  With N>1 ranks (torchrun) each rank holds its own 20,000-gene shard of an
N x 20,000-gene matrix on the same tree (weak scaling): patterns are sharded, every rank evaluates the same P
vectors, and ONE all-reduce of the P partial log-likelihoods per step (NCCL via torch.distributed) joins them.

  python bench.py --gpus 1 --steps 20 --warmup 3          our arm  (CUDA through the C ABI)
  python bench.py --impl reference --steps 3 --warmup 1    reference arm (the reference's own CPU code)
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

N_GENOMES, N_GENES, TREE_SEED = 100, 20000, 1
METRIC = "gene-likelihood evals/s (patterns x param-sets/s)"
UNIT = "pattern-evals/s"


def make_workload(rank, P):
    """Seeded synthetic PPM data: same tree on every rank, rank-specific genes, same P parameter vectors."""
    from ppmmodelpangenome_b200 import simulate
    rng_t = np.random.default_rng(TREE_SEED)
    nwk = simulate.sim_tree(N_GENOMES, rng_t)
    rng_g = np.random.default_rng(1000 + rank)
    names, M, truth = simulate.sim_genes(nwk, N_GENES, rng_g)
    rng_p = np.random.default_rng(77)
    params = np.stack([simulate.start_params(truth, N_GENES, rng_p) for _ in range(P)])
    return nwk, names, M, truth, params


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
            time.sleep(0.02)

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


def cpu_reference_time(nwk, names, M, params, sample_patterns, repeats):
    """The reference's own implementation (oracle/_ref, unmodified sources) -- or, if that library did not
    travel, the C port -- timed on the host cores on a bounded sample of the same workload.
    Times exactly what the objective does: clearLikValues()+logLik() (functions.cpp:406-409)."""
    from ppmmodelpangenome_b200 import simulate
    sel = np.sort(np.random.default_rng(5).choice(M.shape[1], sample_patterns, replace=False))
    sub = M[:, sel]
    params = np.array(params, dtype=np.float64, copy=True)
    params[:, 0] *= sample_patterns / M.shape[1]  # N0 and i1 are gene counts/rates: scale them to the sample
    params[:, 2] *= sample_patterns / M.shape[1]
    # unique patterns of the sample (the reference is handed the deduplicated 'ccc' input, its best case)
    uniq, counts = np.unique(sub.T, axis=0, return_counts=True)
    tmp = tempfile.mkdtemp(prefix="ppm_bench_")
    tp, pp = os.path.join(tmp, "tree.nwk"), os.path.join(tmp, "pa.txt")
    with open(tp, "w") as f:
        f.write(nwk + "\n")
    simulate.write_matrix(pp, names, np.ascontiguousarray(uniq.T), counts)
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
    n_unique = len(uniq)
    return n_unique, times, kind, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nwk, names, M, truth, params = make_workload(0, max(args.steps + args.warmup, 1))
    sample = 2000
    n_unique, _, kind, cores = cpu_reference_time(nwk, names, M, params, sample, args.warmup)
    n_unique, times, kind, cores = cpu_reference_time(nwk, names, M, params, sample, args.steps)
    total = float(np.sum(times))
    value = n_unique * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: simulated %d-genome ultrametric tree x %d genes" % (N_GENOMES, N_GENES),
                       "step": "one clearLikValues()+logLik() call (1 parameter vector) on a bounded sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d random genes of the workload = %d unique patterns, 1 parameter vector per step" % (sample, n_unique)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--params", type=int, default=256, help="parameter vectors per step (P)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from ppmmodelpangenome_b200 import Inference, measure_fp64_peak, simulate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, K, W = args.params, args.steps, max(args.warmup, 3)

    nwk, names, M, truth, params = make_workload(rank, P)
    left, right, bl, rows = simulate.pack_patterns(nwk, names, M)
    inf = Inference.from_arrays(left, right, bl, rows, dedupe=True, device=local)
    n_unique, n_genes = int(inf.nRep), int(inf.nGenes)

    stream = torch.cuda.Stream()          # a real (non-default) stream: kernels, events and NCCL all go here
    torch.cuda.set_stream(stream)
    sp = ctypes.c_void_p(stream.cuda_stream)
    d_params = torch.from_numpy(params).cuda()
    d_ll = torch.zeros(P, dtype=torch.float64, device="cuda")
    d_i2 = torch.zeros(P, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_device():
        inf.loglik_device(d_params.data_ptr(), P, d_ll.data_ptr(), d_i2.data_ptr(), sp)
        if world > 1:
            dist.all_reduce(d_ll)  # the one collective of the path: P partial log-likelihoods
        inf.finalize_device(d_i2.data_ptr(), d_ll.data_ptr(), P, sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    for _ in range(W):
        step_device()
    barrier()
    peak_tf = measure_fp64_peak(local)  # DFMA micro-benchmark on this very GPU
    sampler.start()
    t_wait = time.perf_counter()
    while not sampler.rows and time.perf_counter() - t_wait < 5.0:
        time.sleep(0.01)              # nvidia-smi is up before the timed region starts
    sampler.rows.clear()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sweep_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()  # L2 flush between timed iterations (not timed)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        if k == K - 1 or K <= 64:
            torch.cuda.synchronize()
            sweep_ms.append(inf.counters()["sweep_ms"])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([float(np.sum(step_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    cnt = inf.counters()
    units_per_step = torch.tensor([float(n_unique) * P, float(n_genes) * P], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(units_per_step)
    pattern_evals, gene_evals = (float(x) for x in units_per_step.tolist())
    value = pattern_evals * K / (total_ms * 1e-3)

    # end to end through the public API with HOST buffers (pinned params in, ll+i2 out), copies inside the timed region
    ll_host = np.zeros(P)
    t_e2e = []
    barrier()
    for k in range(max(3, min(K, 10))):
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        part, i2h = inf.loglik_batch(params, return_i2=True)  # ppm_loglik: H2D params, kernels, D2H ll and i2
        if world > 1:
            tt = torch.from_numpy(part).cuda()
            dist.all_reduce(tt)
            part = tt.cpu().numpy()
            part = np.where(i2h < 0, -1e9 + i2h, part)
        ll_host = part
        t_e2e.append(time.perf_counter() - t0)
    sampler.stop_flag = True  # sampled through the device-resident AND the end-to-end timed loops
    e2e_t = torch.tensor([float(np.median(t_e2e))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = pattern_evals / float(e2e_t.item())

    # sanity: device-resident and host paths agree
    assert np.allclose(d_ll.cpu().numpy(), ll_host, rtol=1e-12, atol=0), "device/host path mismatch"

    if rank == 0:
        F = cnt["flops_per_eval"]          # 5W + 28(n-1) + 2S algorithmic flops per pattern-eval
        sweep = float(np.median(sweep_ms)) if sweep_ms else None
        ach = n_unique * P * F / (sweep * 1e-3) / 1e12 if sweep else None
        pipe = n_unique * P * cnt["fp64_instr_per_eval"] * 2 / (sweep * 1e-3) / 1e12 if sweep else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: simulated %d-genome ultrametric tree x %d genes per GPU (PPM simulation, seed %d)"
                                   % (N_GENOMES, N_GENES, TREE_SEED),
                       "param_vectors_per_step": P, "unique_patterns_rank0": n_unique, "genes_per_gpu": n_genes,
                       "parallelism": "patterns sharded over %d GPU(s), one all-reduce of %d doubles per step" % (world, P) if world > 1
                                      else "1 GPU",
                       "l2": "flushed between timed steps (256 MB memset); each step timed with CUDA events on the launching stream"},
            "gene_evals_per_s": gene_evals * K / (total_ms * 1e-3),
            "wall_s_timed_region": t_wall,
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(P * 8 * 8), "d2h_bytes_per_step": int(P * 8 * 2)},
            "gpu_launches": int(cnt["launches"] + 1) * K,  # 7 kernels in ppm_loglik_device + finalize
            "roofline": {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (ach / peak_tf) if ach else None,
                         "traffic": 4.99e7 if (P == 256 and world == 1) else None,  # dram read+write bytes per launch, ncu --set full (profiles/)
                         "kernel": "ppm::sweep_kernel<8,1,0>", "kernel_ms": sweep,
                         "peak_source": "DFMA micro-benchmark measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry)",
                         "algorithmic_flops_per_pattern_eval": F,
                         "fp64_pipe_tflops_issued": pipe, "fp64_pipe_frac": (pipe / peak_tf) if pipe else None},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                nu, times, kind, cores = cpu_reference_time(nwk, names, M, params, 2000, 3)
                line["cpu_baseline"] = {"value": nu * len(times) / float(np.sum(times)), "unit": UNIT, "cores": cores, "kind": kind,
                                        "sample": "2000 random genes of the workload = %d unique patterns, 1 parameter vector per call, %d calls" % (nu, len(times))}
            except Exception as e:  # the baseline is a report, never a reason to lose the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
