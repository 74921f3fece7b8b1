"""Multi-GPU plumbing: shard the bra excitations, one all-reduce of the result.

One process per GPU (torchrun); ``torch.distributed`` is used for nothing but
the rendezvous and the single ``all_reduce(sum)`` of the ``n_bra x n_ket``
partial matrices (NCCL over NVLink on the GPU box, gloo in the CPU tests).
The pair space needs no other exchange.
"""
from __future__ import annotations

import os

import numpy as np

from .excitations import shard_bounds


def world():
    """(rank, world_size) of the initialised default process group, else (0, 1)."""
    try:
        import torch.distributed as dist
    except Exception:  # torch missing: single process
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str | None = None):
    """Initialise the default group from torchrun's environment (no-op for 1 rank)."""
    import torch
    import torch.distributed as dist

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or dist.is_initialized():
        return world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend, rank=int(os.environ["RANK"]), world_size=ws)
    return world()


def sum_over_ranks(part, device=None):
    """The one collective of the path: all-reduce(sum) of the per-rank partial matrices."""
    rank, ws = world()
    if ws == 1:
        return part
    import torch
    import torch.distributed as dist

    was_numpy = isinstance(part, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(part)) if was_numpy else part
    if dist.get_backend() == "nccl" and not t.is_cuda:
        t = t.cuda(device if device is not None else torch.cuda.current_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy() if was_numpy else t


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
    part = partial_fn(lo, hi)
    assert tuple(part.shape) == (n_bra, n_ket)
    return sum_over_ranks(part, device)


def deal_pairs(n_pairs: int, ws: int, rank: int) -> np.ndarray:
    """Indices of the trajectory pairs rank ``rank`` of ``ws`` computes (round-robin)."""
    return np.arange(rank, n_pairs, ws)


def sharded_trajectory(compute_fn, n_pairs: int, n_states: int, device=None):
    """Trajectory pairs over the ranks.  ``compute_fn(pair_indices, rank, world)`` returns the
    ``(len(pair_indices), n_states, n_states)`` matrices of those pairs, computed on ``world`` = 1
    excitation shards when the pairs are dealt out (at least one pair per rank), else every rank
    takes ALL pairs with its shard ``rank`` of ``world`` of the bra excitations.  Either way the
    zero-padded / partial results are summed with the path's one all-reduce."""
    rank, ws = world()
    if ws > 1 and n_pairs >= ws:
        mine = deal_pairs(n_pairs, ws, rank)
        out = np.zeros((n_pairs, n_states, n_states))
        if mine.size:
            out[mine] = compute_fn(mine, 0, 1)
    else:
        out = np.asarray(compute_fn(np.arange(n_pairs), rank, ws))
    assert tuple(out.shape) == (n_pairs, n_states, n_states)
    return sum_over_ranks(out, device)
