"""One-process-per-GPU plumbing for the enumeration path (torch.distributed).

Process r of WORLD enumerates shard r: the INTERLEAVED windows r, r+WORLD, ... of
the rank space (``enumgpu_options.shard_index / shard_count``: windows of equal
estimated cost for the shared kernel, block-sized rank windows for the
independent one — the cost per basis varies along the rank axis, one contiguous
range per GPU was 13 % imbalanced at 2 GPUs) and contributes one 256-byte
``enumgpu_partial`` record.  The only exchange step of the path is an all-gather
of those records (NCCL on GPUs, gloo in the CPU tests) followed by the
associative, commutative merge ``enumgpu_merge_partial`` (lexicographic min on
(key, rank), sums of counters), so every rank ends with the same result and the
result does not depend on WORLD or on how the windows are cut.

``interleaved_windows`` is the same dealing scheme on plain rank windows, for
callers (and the CPU tests) that drive an engine taking [rank_begin, rank_end);
``shard_bounds`` is the contiguous alternative (``enumgpu_shard_begin``), kept for
callers that checkpoint by range.
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


def interleaved_windows(total: int, rank: int, world: int, window: int, rank_begin: int = 0):
    """Rank windows [lo, hi) of shard `rank` of `world`: windows of `window` ranks over [rank_begin, total), dealt
    round-robin — window k belongs to shard k mod world.  The shards of all ranks tile the range exactly."""
    if world < 1 or not 0 <= rank < world or window < 1:
        raise ValueError("bad shard / window")
    out = []
    lo = rank_begin + rank * window
    while lo < total:
        out.append((lo, min(total, lo + window)))
        lo += world * window
    return out


class ShardedEnumeration:
    """The host-buffer solve of one process of a one-process-per-GPU job (the multi-process twin of
    ``EnumerationSolver.enumerate``): pack A|b|c into pinned memory, H2D, enumerate this rank's interleaved shard
    (``enumgpu_enqueue_host_h`` on the handle's stream — one copy, one kernel launch), all-gather the 256-byte
    records on the DEVICE (NCCL), one D2H copy of the gathered records into pinned memory, merge
    (``enumgpu_merge_records``).  Every rank returns the same result.  All buffers, the handle and the stream
    outlive the call."""

    def __init__(self, device_index: int, rank: int, world: int, algo: int = _abi.ALGO_AUTO):
        import torch
        self._torch = torch
        self.rank, self.world, self.algo = int(rank), int(world), int(algo)
        self.dev = torch.device("cuda", device_index)
        self.part = torch.zeros(RECORD_BYTES, dtype=torch.uint8, device=self.dev)
        self.gathered = torch.zeros(world * RECORD_BYTES, dtype=torch.uint8, device=self.dev)
        self.h_gathered = torch.zeros(world * RECORD_BYTES, dtype=torch.uint8).pin_memory()
        self.handle = C.c_void_p()
        if lib().enumgpu_create(device_index, C.byref(self.handle)) != 0:
            raise RuntimeError(lib().enumgpu_last_error().decode())
        # copies, the enumeration launch and the all-gather are all ordered on ONE stream, owned by torch and handed to
        # the library per call (options.stream): torch's allocators remember the streams their buffers were used on,
        # so the stream must outlive the tensors — a stream owned by the handle would die with close()
        self.stream = torch.cuda.Stream(device=self.dev)
        self.h2d_bytes = 0
        self.d2h_bytes = world * RECORD_BYTES

    def close(self):
        if self.handle:
            self.stream.synchronize()
            lib().enumgpu_destroy(self.handle)
            self.handle = C.c_void_p()

    def solve(self, A, b, c, maximize: bool, rank_begin: int = 0, rank_end: int = 0) -> _abi.Result:
        """A: (m, n) float64 in column-major (Fortran) order, b, c: float64 vectors (what Canonical's getters return)."""
        torch = self._torch
        m, n = A.shape
        ps = _abi.Problem(m, n, A.strides[1] // 8 if n > 1 else m, int(bool(maximize)), A.ctypes.data, b.ctypes.data, c.ctypes.data)
        self.h2d_bytes = (m * n + m + n) * 8
        stream = self.stream
        sharded = self.world > 1
        opt = _abi.Options(-1.0, -1.0, rank_begin, rank_end, 0, self.algo, None, stream.cuda_stream,
                           self.rank if sharded else 0, self.world if sharded else 0)
        # pack into pinned memory + H2D + the enumeration kernel, one C call, nothing synchronised
        if lib().enumgpu_enqueue_host_h(self.handle, C.byref(ps), C.byref(opt), self.part.data_ptr(), None) != 0:
            raise RuntimeError(lib().enumgpu_last_error().decode())
        with torch.cuda.stream(stream):
            all_gather_records(self.part, self.gathered, self.world)        # device side (NCCL)
            self.h_gathered.copy_(self.gathered, non_blocking=True)         # one D2H copy into pinned memory
        stream.synchronize()
        res = _abi.Result()
        lib().enumgpu_merge_records(self.h_gathered.data_ptr(), self.world, C.byref(res))
        return res


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
