"""One-process-per-GPU plumbing for the enumeration path (torch.distributed).

The rank space [0, C(n,m)) is cut into WORLD contiguous shards
(``enumgpu_shard_begin``); every rank enumerates its shard and contributes one
256-byte ``enumgpu_partial`` record.  The only exchange step of the path is an
all-gather of those records (NCCL on GPUs, gloo in the CPU tests) followed by
the associative merge ``enumgpu_merge_partial`` (lexicographic min on
(key, rank), sums of counters), so every rank ends with the same result and the
result does not depend on WORLD.
"""
import ctypes as C

from . import _abi
from ._lib import lib

RECORD_BYTES = C.sizeof(_abi.Partial)


def shard_bounds(m: int, n: int, rank: int, world: int, rank_begin: int = 0, rank_end: int = 0):
    L = lib()
    if rank_begin == 0 and rank_end == 0:
        rank_end = L.enumgpu_binomial(n, m)
    return (L.enumgpu_shard_begin(m, n, rank_begin, rank_end, rank, world),
            L.enumgpu_shard_begin(m, n, rank_begin, rank_end, rank + 1, world))


def merge_records(raw: bytes, world: int) -> _abi.Result:
    """raw = WORLD concatenated enumgpu_partial records -> merged result struct."""
    L = lib()
    recs = [_abi.Partial.from_buffer_copy(raw[i * RECORD_BYTES:(i + 1) * RECORD_BYTES]) for i in range(world)]
    for r in recs[1:]:
        L.enumgpu_merge_partial(C.byref(recs[0]), C.byref(r))
    res = _abi.Result()
    L.enumgpu_partial_to_result(C.byref(recs[0]), C.byref(res))
    return res


def all_gather_records(part, gathered, world: int):
    """part: uint8[256] tensor (device of the backend); gathered: uint8[world*256]."""
    import torch.distributed as dist
    if world > 1:
        dist.all_gather_into_tensor(gathered, part)
    else:
        gathered.copy_(part)
    return gathered
