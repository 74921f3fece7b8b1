"""Multi-GPU plumbing: shard the bra excitations, one all-reduce of the result.

One process per GPU (torchrun).  Sharding is OPT-IN: the reference functions are purely local, so a
host application that uses ``torch.distributed`` for its own purposes must not get sharded partial
sums out of ``overlaps_cache`` just because a default process group exists.  A program that wants
the N GPUs of a node to share one call says so once::

    from pywfo_b200 import dist
    dist.init_from_env()      # torchrun environment -> process group + dist.enable()
    # or, with a process group the program created itself:  dist.enable()

and from then on EVERY rank must call the entry points with IDENTICAL inputs; each rank computes
the partial matrix of its slice of the bra excitations and all ranks return the full matrix.
The sum of the ``n_bra x n_ket`` partials is the path's only collective.  On a
GPU box it runs inside the CUDA library (``wfo_comm_init``: one-shot peer-memory all-reduce over
NVLink, csrc/comm.cuh); ``torch.distributed`` (NCCL, or gloo in the CPU tests) is the rendezvous and
the fallback.
"""
from __future__ import annotations

import os
import uuid

import numpy as np

from .excitations import shard_bounds

_enabled = False
_comm_ctx = None     # the _lib.Context that holds the in-library communicator


def enabled() -> bool:
    return _enabled


def _group():
    try:
        import torch.distributed as dist
    except Exception:  # torch missing: single process
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def world():
    """(rank, world_size) the entry points shard over: the default process group once sharding has
    been enabled with ``enable()`` / ``init_from_env()``, else (0, 1)."""
    if not _enabled:
        return 0, 1
    d = _group()
    if d is None:
        return 0, 1
    return d.get_rank(), d.get_world_size()


def enable(device=None, in_library: bool = True):
    """Opt in to sharding over the initialised default process group (collective: all ranks call it).
    On a CUDA box this also sets up the library's own communicator on ``device`` (default: the
    current CUDA device) unless ``in_library`` is False."""
    global _enabled, _comm_ctx
    d = _group()
    if d is None or d.get_world_size() == 1:
        _enabled = d is not None
        return world()
    _enabled = True
    if in_library and d.get_backend() == "nccl":
        import torch
        from . import _lib
        tag = [uuid.uuid4().hex if d.get_rank() == 0 else None]
        d.broadcast_object_list(tag, src=0)
        ctx = _lib.context(torch.cuda.current_device() if device is None else device)
        ctx.comm_init(d.get_rank(), d.get_world_size(), f"{tag[0]}")
        _comm_ctx = ctx
        d.barrier()
    return world()


def disable():
    """Back to purely local calls (collective if a communicator exists)."""
    global _enabled, _comm_ctx
    if _comm_ctx is not None:
        _comm_ctx.comm_destroy()
        _comm_ctx = None
    _enabled = False


def has_comm(ctx) -> bool:
    """Does ``ctx`` hold the library's communicator for the current world?"""
    rank, ws = world()
    return ws > 1 and ctx is _comm_ctx and ctx.comm_world() == ws


def init_from_env(backend: str | None = None, in_library: bool = True):
    """Initialise the default group from torchrun's environment and enable sharding (no-op for 1 rank)."""
    import torch
    import torch.distributed as dist

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1:
        return 0, 1
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend, rank=int(os.environ["RANK"]), world_size=ws)
    return enable(in_library=in_library)


def sum_over_ranks(part, device=None):
    """The one collective of the path: all-reduce(sum) of the per-rank partial matrices.
    A CUDA tensor is reduced in place on its device (library communicator if there is one, else
    NCCL); a numpy array goes through the group's backend and comes back as numpy."""
    rank, ws = world()
    if ws == 1:
        return part
    import torch
    import torch.distributed as dist

    was_numpy = isinstance(part, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(part)) if was_numpy else part
    if dist.get_backend() == "nccl" and not t.is_cuda:
        t = t.cuda(device if device is not None else torch.cuda.current_device())
    if (t.is_cuda and _comm_ctx is not None and t.dtype == torch.float64 and t.is_contiguous()
            and t.device.index == _comm_ctx.device and t.numel() * 8 <= (1 << 20)):
        _comm_ctx.allreduce_dev(t, torch.cuda.current_stream(t.device).cuda_stream)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy() if was_numpy else t


def _raise_together(exc):
    """Make a rank-local failure a failure on every rank: all-reduce a status word BEFORE the data
    collective, so that no rank waits in it for a rank that has already raised."""
    rank, ws = world()
    if ws == 1:
        if exc is not None:
            raise exc
        return
    import torch
    import torch.distributed as dist

    flag = torch.tensor([0 if exc is None else 1], dtype=torch.int32)
    if dist.get_backend() == "nccl":
        flag = flag.cuda()
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if exc is not None:
        raise exc
    if int(flag.item()):
        raise RuntimeError("another rank failed in this sharded call (see its traceback)")


def sharded_state_overlaps(partial_fn, K_bra: int, n_bra: int, n_ket: int, device=None):
    """Run ``partial_fn(k_begin, k_end) -> (n_bra, n_ket)`` on this rank's slice of the
    bra excitations and sum the partial matrices over all ranks.

    ``partial_fn`` returns a numpy array (host path) or a torch tensor (device
    path); the all-reduce runs on whatever device the group's backend needs.
    The product passes a closure over the CUDA entry point; the CPU tests pass
    the oracle so that the sharding logic is exercised without a GPU.
    """
    rank, ws = world()
    lo, hi = shard_bounds(K_bra, ws, rank)
    part, exc = None, None
    try:
        part = partial_fn(lo, hi)
        assert tuple(part.shape) == (n_bra, n_ket)
    except Exception as e:   # noqa: BLE001 -- re-raised on every rank
        exc = e
    _raise_together(exc)
    return sum_over_ranks(part, device)


def deal_pairs(n_pairs: int, ws: int, rank: int) -> np.ndarray:
    """Indices of the trajectory pairs rank ``rank`` of ``ws`` computes (round-robin)."""
    return np.arange(rank, n_pairs, ws)


def sharded_trajectory(compute_fn, n_pairs: int, n_states: int, device=None):
    """Trajectory pairs over the ranks.  ``compute_fn(pair_indices, rank, world)`` returns the
    ``(len(pair_indices), n_states, n_states)`` matrices of those pairs, computed on ``world`` = 1
    excitation shards when the pairs are dealt out (at least one pair per rank), else every rank
    takes ALL pairs with its shard ``rank`` of ``world`` of the bra excitations.  Either way the
    zero-padded / partial results are summed with the path's one all-reduce.  A failure on one rank
    (an empty state in one pair, a singular MO matrix of one cycle) is raised on every rank."""
    rank, ws = world()
    out, exc = None, None
    try:
        if ws > 1 and n_pairs >= ws:
            mine = deal_pairs(n_pairs, ws, rank)
            out = np.zeros((n_pairs, n_states, n_states))
            if mine.size:
                out[mine] = compute_fn(mine, 0, 1)
        else:
            out = np.asarray(compute_fn(np.arange(n_pairs), rank, ws))
        assert tuple(out.shape) == (n_pairs, n_states, n_states)
    except Exception as e:   # noqa: BLE001 -- re-raised on every rank
        exc = e
    _raise_together(exc)
    return sum_over_ranks(out, device)
