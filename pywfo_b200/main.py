"""Drop-in for the hot path of ``pywfo/main.py``: same names, same arguments.

    overlaps | overlaps_naive | overlaps_most_naive | overlaps_cache
        (bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh=1e-4, ao_ovlps="ket")
    moovlp | moovlp_expl | moovlp_dots (mos1, mos2, S_AO)
    get_S_AO(bra_mos, ket_mos, ao_ovlps)

The host only checks shapes; everything the reference computes goes to the CUDA
library through ``include/wfo_b200.h``: the inverse of the MO matrix for
``ao_ovlps in {"bra","ket"}`` (pywfo/main.py:303; Newton-Schulz on the FP64
tensor-core GEMM), S_AO = Cinv Cinv^T, S_MO, the
thresholding of the CI vectors (main.py:539-540, 112, 122-123), every
determinant pair and the CI-weighted reduction.  Extra keyword arguments
(``tier``, ``device``, ``c_inv``) are a superset of the reference signature.

Calls are purely local, like the reference's, unless the program opts in to
sharding with ``pywfo_b200.dist.enable()`` / ``init_from_env()``: then every rank
(one per GPU) must call with identical inputs, computes the partial matrix of its
shard of the bra excitations, and the partials are summed with one all-reduce
(inside the CUDA library over NVLink peer memory, dist.py / csrc/comm.cuh).
"""
from __future__ import annotations

import logging

import numpy as np

from . import _lib, dist
from .excitations import build_tables, check_minus_inputs

logger = logging.getLogger("pywfo_b200")
logger.addHandler(logging.NullHandler())


def log(message, level="info"):
    """pywfo/main.py:23: kept so that callers' ``log`` imports keep working; quiet."""
    getattr(logger, level)(message)


def _mo_inverse(mos):
    """``np.linalg.inv(mos)`` of pywfo/main.py:303 on the device (Newton-Schulz on the FP64
    tensor-core GEMM).  A matrix the iteration cannot invert raises ``LinAlgError`` like numpy
    does for a singular one; there is no host fallback."""
    try:
        return _lib.context().inverse_host(mos)
    except _lib.WfoError as exc:
        if exc.code != _lib.ERR_NOCONV:
            raise
        raise np.linalg.LinAlgError("Singular matrix") from exc


def get_S_AO(bra_mos, ket_mos, ao_ovlps):
    """pywfo/main.py:301-311: inverse and the product ``inv inv^T`` both on the GPU."""
    if isinstance(ao_ovlps, str):
        inv = _mo_inverse(np.asarray({"bra": bra_mos, "ket": ket_mos}[ao_ovlps], dtype=np.float64))
        # inv @ I @ inv^T through the DMMA GEMM chain
        return _lib.context().smo_host(inv, np.eye(inv.shape[1]), inv)
    if isinstance(ao_ovlps, np.ndarray):  # upstream's isinstance(.., np.array) raises TypeError
        return ao_ovlps
    raise Exception("Invalid AO overlaps!")


def moovlp(mos1, mos2, S_AO):
    """<phi_p|phi_q> = sum_uv C_pu C'_qv S_uv (pywfo/main.py:27-43) on the GPU."""
    return _lib.context().smo_host(mos1, S_AO, mos2)


def moovlp_expl(mos1, mos2, S_AO):
    """pywfo/main.py:46-68: same contraction (upstream: explicit Python loops)."""
    return moovlp(mos1, mos2, S_AO)


def moovlp_dots(mos1, mos2, S_AO):
    """pywfo/main.py:71-78: same contraction (upstream: numba-jitted dots)."""
    return moovlp(mos1, mos2, S_AO)


def _state_overlaps(semantics, sigma, bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh, ao_ovlps,
                    tier="auto", device=None, c_inv=None):
    bra_mos = np.asarray(bra_mos, dtype=np.float64)
    ket_mos = np.asarray(ket_mos, dtype=np.float64)
    bra_ci = np.asarray(bra_ci, dtype=np.float64)
    ket_ci = np.asarray(ket_ci, dtype=np.float64)
    occ = int(occ)
    if bra_ci.ndim != 3 or ket_ci.ndim != 3:
        raise ValueError("CI coefficients must have shape (n_states, occ, virt)")
    if bra_ci.shape[1] != occ or ket_ci.shape[1] != occ:
        raise ValueError("CI coefficients do not match occ")
    ctx = _lib.context(device)
    n_bra, n_ket = bra_ci.shape[0], ket_ci.shape[0]
    # upstream only ever indexes S_MO with occ + to (main.py:565-576), so CI tensors that cover FEWER virtual
    # orbitals than the MO matrices have (frozen / truncated virtual space, different widths on the two sides)
    # are fine there: widen them with zero coefficients, which select nothing and weigh nothing.
    n_virt = bra_mos.shape[0] - occ
    if n_virt >= 0 and (bra_ci.shape[2] < n_virt or ket_ci.shape[2] < n_virt):
        def widen(ci):
            if ci.shape[2] >= n_virt:
                return ci
            wide = np.zeros((ci.shape[0], occ, n_virt))
            wide[:, :, :ci.shape[2]] = ci
            return wide
        bra_ci, ket_ci = widen(bra_ci), widen(ket_ci)
    if isinstance(ao_ovlps, str):
        # everything on the device: the inverse behind S_AO (main.py:303), S_AO, S_MO, the selection
        # of excitations above the threshold, every determinant pair and the CI-weighted reduction.
        # ``c_inv`` (optional input): an inverse of the chosen MO matrix the caller already has.
        if ao_ovlps not in ("bra", "ket"):
            raise KeyError(ao_ovlps)   # upstream: dict lookup, main.py:303
        side = _lib.AO_BRA if ao_ovlps == "bra" else _lib.AO_KET
        if c_inv is not None:
            side = _lib.AO_GIVEN   # the caller's inverse of that MO matrix (e.g. LAPACK's, exactly as upstream)
        rank, world = dist.world()
        in_library = dist.has_comm(ctx)   # the library all-gathers the ket side and all-reduces the result itself
        if world > 1 and not in_library and ctx.comm_world():
            raise RuntimeError("the context's communicator does not match the enabled process group")
        try:
            part, _ = ctx.overlaps_ci_host(bra_mos, ket_mos, c_inv, occ, bra_ci, ket_ci, ci_thresh,
                                           ao_side=side, semantics=0 if semantics == "cache" else 1,
                                           tier=tier, rank=rank, world=world)
        except _lib.WfoError as exc:
            if exc.code != _lib.ERR_NOCONV:
                raise
            raise np.linalg.LinAlgError("Singular matrix") from exc   # what np.linalg.inv raises upstream
        return part if in_library else dist.sum_over_ranks(part, device=ctx.device)
    if not isinstance(ao_ovlps, np.ndarray):
        raise Exception("Invalid AO overlaps!")
    # explicit AO overlap matrix (documented upstream, main.py:305): tables built on the host
    if semantics == "minus":
        check_minus_inputs(bra_ci, ket_ci, ci_thresh)
    bra_t = build_tables(bra_ci, ci_thresh, semantics)
    ket_t = build_tables(ket_ci, ci_thresh, semantics, order="locality")
    if bra_t.K == 0 or ket_t.K == 0:
        return np.zeros((n_bra, n_ket))
    s_mo = ctx.smo_host(bra_mos, ao_ovlps, ket_mos)

    def partial(lo, hi):
        return ctx.state_overlaps_host(s_mo, occ, bra_t.exc, bra_t.coeff, ket_t.exc, ket_t.coeff,
                                       sigma=sigma, tier=tier, k_begin=lo, k_end=hi)

    return dist.sharded_state_overlaps(partial, bra_t.K, n_bra, n_ket, device=ctx.device)


def overlaps_cache(bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh=1e-4, ao_ovlps="ket", **kw):
    """Singlet->singlet state overlaps, pywfo/main.py:520-627."""
    return _state_overlaps("cache", 1.0, bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh, ao_ovlps, **kw)


def overlaps_naive(bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh=1e-4, ao_ovlps="ket", **kw):
    """pywfo/main.py:314-419 (the live definition): same numbers as ``overlaps_cache``."""
    return _state_overlaps("cache", 1.0, bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh, ao_ovlps, **kw)


def overlaps_most_naive(bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh=1e-4, ao_ovlps="ket", **kw):
    """pywfo/main.py:422-517: same numbers as ``overlaps_cache``."""
    return _state_overlaps("cache", 1.0, bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh, ao_ovlps, **kw)


def overlaps(bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh=1e-4, ao_ovlps="ket", **kw):
    """The "minus" (Ms=0 triplet) formula, pywfo/main.py:81-199, with upstream's
    assertions (main.py:90-91), strict threshold and sign^occ behaviour."""
    bra_mos = np.asarray(bra_mos)
    ket_mos = np.asarray(ket_mos)
    bra_ci = np.asarray(bra_ci)
    ket_ci = np.asarray(ket_ci)
    assert bra_mos.shape == ket_mos.shape
    assert bra_ci.shape == ket_ci.shape
    if not isinstance(ao_ovlps, (str, np.ndarray)):
        raise Exception("Invalid AO overlaps!")
    return _state_overlaps("minus", -1.0, bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh, ao_ovlps, **kw)
