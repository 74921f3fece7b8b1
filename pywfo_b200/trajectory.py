"""Trajectory driver: state overlaps between the steps of a trajectory.

Upstream, pysisyphus calls ``pywfo.main.overlaps_cache`` once per pair of geometries
(pywfo/tests/ref_cytosin holds a 9-cycle example: ``mo_coeffs (9,137,137)``, ``ci_coeffs
(9,2,29,108)``); every call re-derives S_AO from the MO matrix and walks the CI vectors again.
``trajectory_overlaps`` hands the whole trajectory to the CUDA library in ONE call
(``wfo_trajectory_ci_host``): each cycle's MO matrix and CI tensor is uploaded once and stays
resident in HBM, the inverse behind S_AO is formed once per cycle, and per pair only S_AO, S_MO,
the device-side thresholding and the fused determinant/contraction kernels run.

With several ranks (torchrun) the pairs are dealt round-robin to the ranks when there are at least
as many pairs as ranks (no exchange but the final sum of the zero-padded results); otherwise every
pair shards its bra excitations like a single ``overlaps_cache`` call does.
"""
from __future__ import annotations

import numpy as np

from . import _lib, dist


def consecutive_pairs(n_cycles: int, stride: int = 1) -> np.ndarray:
    """(i, i + stride) for every cycle that has a successor: the overlaps a surface-hopping or
    root-following driver asks for."""
    i = np.arange(0, max(n_cycles - stride, 0))
    return np.stack([i, i + stride], axis=1).astype(np.int32)


def referenced_cycles(pairs: np.ndarray):
    """The cycles a set of pairs touches and the pairs renumbered into that subset: a rank that is dealt a few
    pairs of a long trajectory uploads (and inverts) only those cycles."""
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    cycles, inverse = np.unique(pairs, return_inverse=True)
    return cycles.astype(np.int64), inverse.reshape(-1, 2).astype(np.int32)


def trajectory_overlaps(mo_coeffs, ci_coeffs, occ, ci_thresh=1e-4, ao_ovlps="ket", pairs=None, semantics="cache",
                        tier="auto", device=None):
    """``out[p] = overlaps_cache(mo[i], mo[j], ci[i], ci[j], occ, ci_thresh, ao_ovlps)`` for
    ``(i, j) = pairs[p]`` (default: consecutive cycles); ``semantics="minus"`` gives
    ``pywfo.main.overlaps`` (main.py:81-199) instead of ``overlaps_cache`` (main.py:520-627).

    mo_coeffs (n_cycles, n_mo, n_mo), ci_coeffs (n_cycles, n_states, occ, n_mo - occ).
    Returns (n_pairs, n_states, n_states) float64.
    """
    mos = np.asarray(mo_coeffs, dtype=np.float64)
    ci = np.asarray(ci_coeffs, dtype=np.float64)
    if mos.ndim != 3 or ci.ndim != 4 or mos.shape[0] != ci.shape[0]:
        raise ValueError("mo_coeffs (n_cycles, n_mo, n_mo) and ci_coeffs (n_cycles, n_states, occ, virt) expected")
    if ao_ovlps not in ("bra", "ket"):
        raise Exception("Invalid AO overlaps!")   # pywfo/main.py:308
    if semantics not in ("cache", "minus"):
        raise ValueError("semantics must be 'cache' or 'minus'")
    pr = consecutive_pairs(mos.shape[0]) if pairs is None else np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    ns = ci.shape[1]
    ctx = _lib.context(device)
    side = _lib.AO_BRA if ao_ovlps == "bra" else _lib.AO_KET
    sem = 0 if semantics == "cache" else 1

    def compute(idx, rank, world):
        try:
            cycles, local = referenced_cycles(pr[idx])
            if cycles.size == mos.shape[0]:
                sub_mos, sub_ci = mos, ci
            else:   #  only what this rank's pairs need goes to the device
                sub_mos, sub_ci = np.ascontiguousarray(mos[cycles]), np.ascontiguousarray(ci[cycles])
            return ctx.trajectory_ci_host(sub_mos, sub_ci, occ, local, ci_thresh, sem, tier, side, rank=rank, world=world)[0]
        except _lib.WfoError as exc:
            if exc.code != _lib.ERR_NOCONV:
                raise
            raise np.linalg.LinAlgError("Singular matrix") from exc

    return dist.sharded_trajectory(compute, pr.shape[0], ns, device=ctx.device)
