"""Synthetic inputs, host side (numpy), as the reference's generator makes them.

The MOs follow pywfo/helpers.py:4-33 call for call (same global-RNG draws, same
QR), so a seeded run here reproduces the reference's matrices bit for bit; the
CI vectors follow the recipe of SURVEY.md section 8(d).  Pure input generation:
nothing here is on the timed path.
"""
from __future__ import annotations

import numpy as np


def get_dummy_mos(num, seed=None):
    """Orthonormal dummy MOs, one per row (pywfo/helpers.py:4-15)."""
    if seed is not None:
        np.random.seed(seed)
    q, _ = np.linalg.qr(np.random.rand(num, num), mode="complete")
    return q.T


def perturb_mat(mat, scale=2e-1):
    """pywfo/helpers.py:18-21."""
    return mat + np.random.rand(*mat.shape) * scale


def get_bra_ket_dummy_mos(num, occ, seed=None):
    """pywfo/helpers.py:24-33.  ``seed`` is accepted and ignored, as upstream."""
    virt = num - occ
    bra_mos = get_dummy_mos(num)
    ket_mos, _ = np.linalg.qr(perturb_mat(bra_mos.T), mode="complete")
    return virt, bra_mos, ket_mos.T


def random_ci(seed, n_states, occ, virt, top_k=None, shared_k=None):
    """Normalised random CIS vectors: standard normal, optional truncation, unit
    Frobenius norm per state.  ``top_k`` keeps each state's K
    largest magnitudes; ``shared_k`` keeps one random subset of K excitations
    shared by all states (the determinant-count sweep)."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((n_states, occ, virt))
    if shared_k is not None and shared_k < occ * virt:
        keep = np.zeros(occ * virt, dtype=bool)
        keep[rng.choice(occ * virt, size=shared_k, replace=False)] = True
        c = np.where(keep.reshape(1, occ, virt), c, 0.0)
    elif top_k is not None and top_k < occ * virt:
        flat = np.abs(c.reshape(n_states, -1))
        kth = np.partition(flat, -top_k, axis=1)[:, -top_k][:, None]
        c = np.where(flat >= kth, c.reshape(n_states, -1), 0.0).reshape(c.shape)
    c /= np.linalg.norm(c.reshape(n_states, -1), axis=1)[:, None, None]
    return c


def synthetic_case(seed, n_mo, occ, n_bra, n_ket, top_k=None, shared_k=None):
    """MOs + CI vectors of one benchmark configuration."""
    np.random.seed(seed)
    virt, bra_mos, ket_mos = get_bra_ket_dummy_mos(n_mo, occ)
    bra_ci = random_ci(seed, n_bra, occ, virt, top_k, shared_k)
    ket_ci = random_ci(seed + 1000, n_ket, occ, virt, top_k, shared_k)
    return bra_mos, ket_mos, bra_ci, ket_ci
