"""CPU tier, world_size 2 over gloo: the multi-rank plumbing of the path.  Patterns are sharded contiguously by
the C-ABI library's own partition (host-only handles expose it); every rank evaluates ITS slice (here with the
C oracle standing in for the device kernels, which cannot run without a GPU), one all_reduce(sum) of the P
partial log-likelihoods joins them, and the i2 < 0 penalty is applied after the reduction (ppm_finalize_ll) --
exactly the sequence bench.py runs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLD, ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes
    import ppmmodelpangenome_b200 as ppm
    from oracle import cport, model
    tree, pa = os.path.join(GOLD, "syn100.nwk"), os.path.join(GOLD, "syn100.pa.txt")
    h = ppm.Inference(tree, pa, dedupe=True, device=-1, shard_rank=rank, shard_world=world)
    first, count = h.info.shard_first, h.info.shard_count
    m = model.model_from_files(tree, pa, dedupe=True)
    assert m["n_rep"] == h.nRep and np.array_equal(m["counts"], h.counts)
    orc = cport.Oracle(m)
    import json
    params = np.array([r["params"] for r in json.load(open(os.path.join(GOLD, "syn100.json")))["results"]][:3])
    params = np.concatenate([params, [[5000, 0.02, 130, 0.4, 1.0, 50.0, 0.01, 0.01]]])  # infeasible: i2 < 0
    part = np.zeros(len(params)); i2 = np.zeros(len(params))
    for k, p in enumerate(params):
        e = orc.eval(p, first_rep=int(first), n_eval=int(count))
        i2[k] = e["i2"]
        part[k] = 0.0 if e["i2"] < 0 else e["ll"]        # the shard's partial sum; skipped work when infeasible
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)                                   # the one collective of the path
    ll = t.numpy().copy()
    rc = ppm.lib().ppm_finalize_ll(i2.ctypes.data_as(ctypes.c_void_p), ll.ctypes.data_as(ctypes.c_void_p), len(params))
    assert rc == 0
    if rank == 0:
        full = np.array([orc.eval(p)["ll"] for p in params])
        ret["ll"], ret["full"], ret["i2"] = ll, full, i2
        ret["spans"] = (int(first), int(count), int(h.nRep))
    dist.destroy_process_group()


def test_two_ranks_shard_allreduce_finalize():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29541, ret), nprocs=2, join=True)
    ll, full, i2 = ret["ll"], ret["full"], ret["i2"]
    assert np.allclose(ll[:3], full[:3], rtol=1e-12, atol=0)
    assert i2[3] < 0 and ll[3] == -1e9 + i2[3] == full[3]
    first, count, n = ret["spans"]
    assert first == 0 and count == n // 2
