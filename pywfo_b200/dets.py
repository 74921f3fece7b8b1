"""Drop-in for ``pywfo.dets.wfoverlap`` (pywfo/dets.py:201-285).

Upstream builds wfoverlap-style occupation strings, sorts them into alpha/beta
"blocks", precomputes every block x block determinant and contracts
``Bci^T (alpha_ovl * beta_ovl) Kci``.  For the closed-shell CIS determinants
it generates, every alpha/beta block is either the ground state or one single
excitation, so the same numbers come out of the fused pair kernel with the
string parity folded into the coefficients (excitations.py, "wfoverlap").
``precompute`` exposes the general block x block determinant table
(dets.py:191-198) for arbitrary index lists.
"""
from __future__ import annotations

import numpy as np

from . import _lib, dist
from .excitations import build_tables


def _blocks_of(blocked):
    """Index lists from the reference's argument (a ``BlockResult`` with ``.blocks``, pywfo/dets.py:114-156) or from
    a plain (n_blocks, n) array of index lists."""
    return np.asarray(getattr(blocked, "blocks", blocked), dtype=np.int32)


def precompute(bra_blocked, ket_blocked, mo_ovlps):
    """Determinants of ``mo_ovlps[bra_block][:, ket_block]`` for all block pairs
    (pywfo/dets.py:191-198), one pivoted LU per pair on the GPU.  Arguments as upstream (``BlockResult`` tuples
    from ``block_dets``) or plain index-list arrays."""
    return _lib.context().det_table_host(mo_ovlps, _blocks_of(bra_blocked), _blocks_of(ket_blocked))


def wfoverlap(bra_mos, ket_mos, bra_ci, ket_ci, ci_thresh, S_AO, closed_shell=True, debug=True,
              tier="auto", device=None):
    assert closed_shell   # dets.py:204
    bra_ci = np.asarray(bra_ci, dtype=np.float64)
    ket_ci = np.asarray(ket_ci, dtype=np.float64)
    np.testing.assert_allclose(bra_ci.shape[1:], ket_ci.shape[1:])   # dets.py:206
    occ = bra_ci.shape[1]
    ctx = _lib.context(device)
    bra_t = build_tables(bra_ci, ci_thresh, "wfoverlap")
    ket_t = build_tables(ket_ci, ci_thresh, "wfoverlap", order="locality")
    n_bra, n_ket = bra_ci.shape[0], ket_ci.shape[0]
    if bra_t.K == 0 or ket_t.K == 0:
        return np.zeros((n_bra, n_ket))
    s_mo = ctx.smo_host(bra_mos, S_AO, ket_mos)   # dets.py:264

    def partial(lo, hi):
        return ctx.state_overlaps_host(s_mo, occ, bra_t.exc, bra_t.coeff, ket_t.exc, ket_t.coeff,
                                       sigma=1.0, tier=tier, k_begin=lo, k_end=hi)

    return dist.sharded_state_overlaps(partial, bra_t.K, n_bra, n_ket, device=ctx.device)
