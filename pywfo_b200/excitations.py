"""Host-side index / coefficient builder feeding the determinant kernels.

Turns the reference's per-state, per-call thresholding into the form the CUDA
path consumes: one excitation table ``exc (K, 2) int32`` per side (the union of
the excitations any state keeps) and one coefficient matrix ``C (n_states, K)``
with exact zeros where a state drops an excitation, so that

    S = Cb ( D00 * D + sigma * dk0 d0l^T ) Ck^T

reproduces each reference entry point term for term.  The sign conventions of
the three reference behaviours are folded into the coefficients here, so the
kernels always work in pywfo's appended ordering (pywfo/main.py:565-576).

    semantics   reference                      threshold  sigma  coefficient factor
    "cache"     main.py:520-627 (also naive,   |c| >= thr   +1   1
                most_naive main.py:314-517)
    "minus"     main.py:81-199                 |c| >  thr   -1   ((-1)^(occ-f+1))^occ   (main.py:154,176)
    "wfoverlap" dets.py:6-39,159-168,201-285   |c| >  thr   +1   (-1)^(occ-f), whole CI column,
                                                                  one entry per (state,f,t) hit
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Tables:
    exc: np.ndarray      # (K, 2) int32, (from, to)
    coeff: np.ndarray    # (n_states, K) float64
    n_occ: int
    n_virt: int

    @property
    def K(self) -> int:
        return int(self.exc.shape[0])


def _mask(ci: np.ndarray, thresh: float, strict: bool) -> np.ndarray:
    a = np.abs(ci)
    return a > thresh if strict else a >= thresh


def locality_order(exc: np.ndarray, block: int = 32) -> np.ndarray:
    """Permutation that sorts excitations by (to // block, from, to).

    The rank-update kernel gathers ``W[t_bra, t_ket]``; walking the ket list in
    this order keeps a CTA's working set of W to a (128 x block) window at a
    time.  The sum over excitations does not depend on the order."""
    if exc.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    f, t = exc[:, 0].astype(np.int64), exc[:, 1].astype(np.int64)
    return np.lexsort((t, f, t // block))


def build_tables(ci: np.ndarray, thresh: float, semantics: str = "cache", order: str | None = None) -> Tables:
    ci = np.asarray(ci, dtype=np.float64)
    if ci.ndim != 3:
        raise ValueError("CI coefficients must have shape (n_states, occ, virt)")
    n_states, n, v = ci.shape
    if semantics in ("cache", "minus"):
        keep = _mask(ci, thresh, strict=(semantics == "minus"))          # per state
        f, t = np.nonzero(keep.any(axis=0))                               # row-major union
        coeff = np.where(keep[:, f, t], ci[:, f, t], 0.0)
        if semantics == "minus":
            # det(sign * M) = sign^occ det(M): main.py:154 and :176
            q = ((-1.0) ** (n - f + 1)) ** n
            coeff = coeff * q[None, :]
    elif semantics == "wfoverlap":
        # np.nonzero over (state, f, t): an excitation kept by m states appears m times,
        # each time with the whole CI column (dets.py:11, 21-22)
        _, f, t = np.nonzero(_mask(ci, thresh, strict=True))
        coeff = ci[:, f, t] * ((-1.0) ** (n - f))[None, :]
    else:
        raise ValueError(f"unknown semantics {semantics!r}")
    exc = np.stack([f, t], axis=1).astype(np.int32).reshape(-1, 2)
    coeff = np.ascontiguousarray(coeff, dtype=np.float64).reshape(n_states, -1)
    if order == "locality" and exc.shape[0] > 1:
        perm = locality_order(exc)
        exc = np.ascontiguousarray(exc[perm])
        coeff = np.ascontiguousarray(coeff[:, perm])
    return Tables(exc=exc, coeff=coeff, n_occ=n, n_virt=v)


def check_minus_inputs(bra_ci: np.ndarray, ket_ci: np.ndarray, thresh: float):
    """``overlaps`` (main.py:186) unpacks ``zip(*combinations)`` per state pair and
    so raises ValueError when either state keeps no excitation."""
    for ci in (bra_ci, ket_ci):
        kept = (np.abs(ci) > thresh).reshape(ci.shape[0], -1).any(axis=1)
        if not kept.all():
            raise ValueError("not enough values to unpack (expected 2, got 0)")


def exc_index_lists(exc: np.ndarray, n_occ: int) -> np.ndarray:
    """(K, n_occ) appended-order MO index lists (pywfo/main.py:565-576); host twin of
    the device kernel ``exc_lists_kernel``, used for explicit-list calls."""
    exc = np.asarray(exc).reshape(-1, 2)
    K = exc.shape[0]
    base = np.arange(n_occ, dtype=np.int32)
    out = np.tile(base, (K, 1))
    for k in range(K):
        f, t = int(exc[k, 0]), int(exc[k, 1])
        if f >= 0:
            out[k, f:n_occ - 1] = base[f + 1:]
            out[k, n_occ - 1] = n_occ + t
    return out


def shard_bounds(K: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced slice of the bra excitations for one rank
    (every pair costs the same, so equal counts balance)."""
    base, rem = divmod(K, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
