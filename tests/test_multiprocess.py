"""CPU, world_size 2 and 3, gloo: the N>1 host logic — interleaved shards (window k belongs to rank k mod
WORLD, the dealing scheme of enumgpu_options.shard_index/shard_count), all-gather of the 256-byte partial
records, merge — with the oracle standing in for the GPU engine of each rank (the exchange + merge code is
the code bench.py runs over NCCL).  The LP is degenerate on purpose: exact ties at the optimum land in
different shards and the merge must still pick the lowest rank.  The window arithmetic the shared kernel
itself uses on its weight axis is checked by simplexmethod_b200/csrc/tests/test_weights.cu (tests/test_abi.py)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _to_partial(_abi, r, m):
    p = _abi.Partial()
    p.key = r.key if r.status == 0 else float("inf")
    p.best_rank = r.best_rank if r.status == 0 else _abi.UINT64_MAX
    p.n_bases, p.n_singular, p.n_infeasible, p.n_feasible = r.n_bases, r.n_singular, r.n_infeasible, r.n_feasible
    p.objective, p.m = r.objective, m
    for i in range(m):
        p.x_B[i], p.basis[i] = r.x_B[i], r.basis[i]
    return p


def _worker(rank, world, port, lp, scheme, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import enumcpu
    from simplexmethod_b200 import _abi, dist as edist, lib, lpgen
    A, b, c, mx = getattr(lpgen, lp[0])(*lp[1:])
    m, n = A.shape
    total = lib().enumgpu_binomial(n, m)
    if scheme == "interleaved":
        wins = edist.interleaved_windows(total, rank, world, window=997)
    else:
        wins = [edist.shard_bounds(m, n, rank, world)]
    p = None
    for (wlo, whi) in wins:                          # this rank's windows, merged like the device merges its units
        r, _ = enumcpu.solve(A, b, c, mx, rank_begin=wlo, rank_end=whi)
        pw = _to_partial(_abi, r, m)
        if p is None:
            p = pw
        else:
            lib().enumgpu_merge_partial(C.byref(p), C.byref(pw))
    if p is None:
        p = _abi.Partial(); p.key = float("inf"); p.best_rank = _abi.UINT64_MAX; p.m = m
    lo, hi = (wins[0][0], wins[-1][1]) if wins else (0, 0)
    part = torch.frombuffer(bytearray(bytes(p)), dtype=torch.uint8).clone()
    gathered = torch.zeros(world * edist.RECORD_BYTES, dtype=torch.uint8)
    edist.all_gather_records(part, gathered, world)
    res = edist.merge_records(gathered.numpy().tobytes(), world)
    q.put((rank, lo, hi, res.status, res.best_rank, list(res.basis)[:m], res.objective,
           res.n_bases, res.n_singular, res.n_infeasible, res.n_feasible, sum(h - l for l, h in wins)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scheme,lp", [("interleaved", ("small_degenerate_lp",)), ("interleaved", ("dense_lp", 6, 16, 21)),
                                       ("contiguous", ("dense_lp", 6, 16, 21))])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_enumeration_matches_single_process(oracle, world, scheme, lp):
    from simplexmethod_b200 import lpgen
    A, b, c, mx = getattr(lpgen, lp[0])(*lp[1:])
    m = A.shape[0]
    full, _ = oracle.solve(A, b, c, mx)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world + 7 * len(lp) + (3 if scheme == "contiguous" else 0)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, lp, scheme, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(o[-1] for o in outs) == full.n_bases                     # the shards tile the rank space
    if scheme == "contiguous":
        assert outs[0][1] == 0 and outs[-1][2] == full.n_bases
        for a, bb in zip(outs[:-1], outs[1:]):
            assert a[2] == bb[1]
    want = (full.status, full.best_rank, list(full.basis)[:m], full.objective,
            full.n_bases, full.n_singular, full.n_infeasible, full.n_feasible)
    for o in outs:                                                       # every rank holds the same merged result
        assert tuple(o[3:-1]) == want


def test_interleaved_windows_tile_the_range():
    from simplexmethod_b200 import dist as edist
    for total, world, window, begin in ((8008, 3, 997, 0), (100, 8, 7, 13), (5, 4, 10, 0), (31824, 2, 1, 31000)):
        seen = []
        for r in range(world):
            seen += edist.interleaved_windows(total, r, world, window, begin)
        seen.sort()
        assert seen[0][0] == begin and seen[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(seen[:-1], seen[1:]))
    with pytest.raises(ValueError):
        edist.interleaved_windows(10, 3, 3, 5)


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
