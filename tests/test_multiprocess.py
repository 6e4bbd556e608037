"""CPU, world_size 2, gloo: the N>1 host logic — shard bounds, all-gather of the
256-byte partial records, merge — with the oracle standing in for the GPU engine
of each rank (the exchange + merge code is the code bench.py runs over NCCL)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, m, n, seed, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import enumcpu
    from simplexmethod_b200 import _abi, dist as edist, lpgen
    A, b, c, mx = lpgen.dense_lp(m, n, seed)
    lo, hi = edist.shard_bounds(m, n, rank, world)
    r, _ = enumcpu.solve(A, b, c, mx, rank_begin=lo, rank_end=hi)
    p = _abi.Partial()
    p.key = r.key if r.status == 0 else float("inf")
    p.best_rank = r.best_rank
    p.n_bases, p.n_singular, p.n_infeasible, p.n_feasible = r.n_bases, r.n_singular, r.n_infeasible, r.n_feasible
    p.objective, p.m = r.objective, m
    for i in range(m):
        p.x_B[i], p.basis[i] = r.x_B[i], r.basis[i]
    part = torch.frombuffer(bytearray(bytes(p)), dtype=torch.uint8).clone()
    gathered = torch.zeros(world * edist.RECORD_BYTES, dtype=torch.uint8)
    edist.all_gather_records(part, gathered, world)
    res = edist.merge_records(gathered.numpy().tobytes(), world)
    q.put((rank, lo, hi, res.status, res.best_rank, list(res.basis)[:m], res.objective,
           res.n_bases, res.n_singular, res.n_infeasible, res.n_feasible))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_enumeration_matches_single_process(oracle, world):
    from simplexmethod_b200 import lpgen
    m, n, seed = 6, 16, 21
    A, b, c, mx = lpgen.dense_lp(m, n, seed)
    full, _ = oracle.solve(A, b, c, mx)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs[0][1] == 0 and outs[-1][2] == 8008                      # shards tile [0, C(16,6))
    for a, bb in zip(outs[:-1], outs[1:]):
        assert a[2] == bb[1]
    want = (full.status, full.best_rank, list(full.basis)[:m], full.objective,
            full.n_bases, full.n_singular, full.n_infeasible, full.n_feasible)
    for o in outs:                                                       # every rank holds the same merged result
        assert tuple(o[3:]) == want


def test_shard_bounds_properties():
    from simplexmethod_b200 import dist as edist, lib
    total = lib().enumgpu_binomial(40, 12)
    for world in (1, 2, 4, 8, 7):
        cuts = [edist.shard_bounds(12, 40, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1
    assert edist.shard_bounds(6, 16, 1, 2, rank_begin=100, rank_end=101) == (100, 101)
