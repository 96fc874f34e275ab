"""Expert sharding across GPUs: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

GRBCM experts are independent until the precision-weighted sums over the expert axis
(gr_bcm.py:133,143-144), so the local experts are split into contiguous blocks, one per rank, and the only
data-path collective is an all-reduce(sum, float64) of the [3, m] partial sums per prediction batch (and of
[1+P] NLML/gradient scalars when hyper-parameters are shared).  The communication expert is replicated.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def world(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if not is_dist():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_bounds(nc: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the nc local experts owned by `rank`: a balanced split, floor(nc/G) each and one more
    on the first nc % G ranks (255 over 8 -> 32 x 7 + 31, the split of SURVEY 8e; 9 over 8 -> 2,1,1,...: no empty rank
    while nc >= G).  A rank with lo == hi owns nothing and contributes zero partial sums."""
    base, extra = divmod(nc, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


_cabi_comm = None   # (handle, world) of the library's own NCCL communicator, created on first use when options.comm == "cabi"


def cabi_comm():
    """The C ABI's communicator over the default process group (gpr_comm_init): rank 0's NCCL id travels through torch.distributed
    once, after that the all-reduces go through gpr_comm_allreduce_f64 — the path a consumer without torch would use."""
    global _cabi_comm
    if _cabi_comm is None:
        import ctypes
        from . import _lib
        lib = _lib.load()
        rank, ws = dist.get_rank(), dist.get_world_size()
        nbytes = int(lib.gpr_comm_unique_id_bytes())
        buf = (ctypes.c_char * nbytes)()
        if rank == 0:
            _lib.call("gpr_comm_unique_id", ctypes.cast(buf, ctypes.c_void_p))
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=0)
        buf = (ctypes.c_char * nbytes).from_buffer_copy(box[0])
        handle = ctypes.c_void_p()
        _lib.call("gpr_comm_init", ctypes.cast(ctypes.byref(handle), ctypes.c_void_p), ws, rank, ctypes.cast(buf, ctypes.c_void_p))
        _cabi_comm = (handle, ws)
    return _cabi_comm[0]


def cabi_comm_destroy() -> None:
    global _cabi_comm
    if _cabi_comm is not None:
        from . import _lib
        _lib.call("gpr_comm_destroy", _cabi_comm[0])
        _cabi_comm = None


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks (no-op for a single process).  options.comm == "cabi" routes float64 CUDA tensors on the default
    group through the library's own NCCL wrapper (gpr_comm_allreduce_f64) instead of torch.distributed."""
    if is_dist() and dist.get_world_size(group) > 1:
        from .config import options
        if options.comm == "cabi" and group is None and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous():
            from . import _lib
            _lib.call("gpr_comm_allreduce_f64", cabi_comm(), t.data_ptr(), t.numel(), torch.cuda.current_stream(t.device).cuda_stream)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_gather_rows(t: torch.Tensor, counts, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks [n_r, m] (n_r = counts[r]) in rank order."""
    if not (is_dist() and dist.get_world_size(group) > 1):
        return t
    mx = max(counts)
    pad = torch.zeros(mx, *t.shape[1:], dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, counts)], dim=0)
