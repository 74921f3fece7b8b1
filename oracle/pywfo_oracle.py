"""numpy restatement of pywfo's CI wavefunction-overlap path (TEST INFRASTRUCTURE).

Every function cites the reference lines (relative to /root/reference) whose
arithmetic it restates.  Two families live here:

* ``ref_*``  -- faithful restatements of the reference entry points, loop for
  loop (per-state thresholding, memoised ``np.linalg.det`` of gathered minors,
  Python double loop).  These are what ``bench.py`` times as the CPU baseline
  ("port") and what the golden vectors pin.
* ``pair_det_table`` / ``state_overlaps_*`` -- the same mathematics in the
  union-of-excitations / coefficient-matrix form the CUDA kernels use
  (``S = Cb (D00*D + sigma*dk0 d0l^T) Ck^T``), brute force and closed form.
  They are the per-shard checkers of the GPU path.

Conventions: MOs are rows; ``ci[state, from, to]`` with
``to`` relative to ``occ``; an excitation (f, t) maps to the index list
``[0..occ-1] minus f, then occ+t appended`` ("appended ordering",
pywfo/main.py:565-576).  The ground state is the excitation ``(-1, -1)``.
"""
from __future__ import annotations

import itertools
from math import sqrt

import numpy as np

GS = (-1, -1)


# --------------------------------------------------------------------------
# stage 1: AO overlap and MO overlap
# --------------------------------------------------------------------------
def get_S_AO(bra_mos, ket_mos, ao_ovlps="ket"):
    """pywfo/main.py:301-311: ``S_AO = C^-1 C^-T`` of the bra or ket MOs.

    The reference's ndarray branch (main.py:305) raises TypeError because
    ``np.array`` is not a type; the oracle accepts an ndarray (superset).
    """
    if isinstance(ao_ovlps, str):
        if ao_ovlps not in ("bra", "ket"):
            raise KeyError(ao_ovlps)
        inv = np.linalg.inv({"bra": bra_mos, "ket": ket_mos}[ao_ovlps])
        return inv.dot(inv.T)
    if isinstance(ao_ovlps, np.ndarray):
        return ao_ovlps
    raise Exception("Invalid AO overlaps!")


def mo_overlap(bra_mos, ket_mos, S_AO):
    """pywfo/main.py:532 (== main.py:27-43, 71-78): ``C_bra S_AO C_ket^T``."""
    return bra_mos.dot(S_AO).dot(ket_mos.T)


def mo_overlap_loops(mos1, mos2, S_AO):
    """pywfo/main.py:46-68 (``moovlp_expl``): the same contraction as two
    explicit reductions, first over u then over v."""
    pv = np.array([[np.sum(S_AO[:, v] * mos1[p, :]) for v in range(S_AO.shape[1])]
                   for p in range(mos1.shape[0])])
    return np.array([[np.sum(pv[p, :] * mos2[q, :]) for q in range(mos2.shape[0])]
                     for p in range(mos1.shape[0])])


# --------------------------------------------------------------------------
# stage 2: determinant of one pair of Slater determinants
# --------------------------------------------------------------------------
def exc_index_list(exc, occ):
    """pywfo/main.py:565-576 (``form_exc_inds``): appended-order MO index list."""
    f, t = exc
    if f < 0:
        return list(range(occ))
    inds = [i for i in range(occ) if i != f]
    inds.append(occ + t)
    return inds


def sd_pair_det(S_MO, occ, bra_exc, ket_exc):
    """pywfo/main.py:579-585 (``sd_ovlp``): LAPACK determinant of the gathered
    ``occ x occ`` minor of S_MO."""
    rows = exc_index_list(bra_exc, occ)
    cols = exc_index_list(ket_exc, occ)
    return np.linalg.det(S_MO[rows][:, cols])


def index_lists(excs, occ):
    """(K, occ) int array of appended-order index lists for a list of excitations."""
    return np.array([exc_index_list(e, occ) for e in excs], dtype=np.int64).reshape(len(excs), occ)


def pair_det_table(S_MO, occ, bra_excs, ket_excs, chunk=4096):
    """All-pairs table ``D[k, l] = det(S_MO[rows_k][:, cols_l])`` (main.py:583-584
    applied to every pair), batched through LAPACK exactly like ``np.linalg.det``."""
    R = index_lists(bra_excs, occ)
    C = index_lists(ket_excs, occ)
    D = np.empty((len(R), len(C)))
    for k0 in range(0, len(R), max(1, chunk // max(1, len(C)))):
        k1 = min(len(R), k0 + max(1, chunk // max(1, len(C))))
        minors = S_MO[R[k0:k1, None, :, None], C[None, :, None, :]]
        D[k0:k1] = np.linalg.det(minors)
    return D


def det_from_lists(S_MO, rows, cols):
    """Determinants for explicit (K_b, n) row lists x (K_k, n) column lists
    (pywfo/dets.py:191-198, ``precompute``)."""
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    minors = S_MO[rows[:, None, :, None], cols[None, :, None, :]]
    return np.linalg.det(minors)


# --------------------------------------------------------------------------
# stage 3, faithful restatements of the reference entry points
# --------------------------------------------------------------------------
def _above(ci_state, thresh, strict=False):
    """main.py:539-540 (``>=``) / main.py:122-123 (``>``), row-major order."""
    mask = np.abs(ci_state) > thresh if strict else np.abs(ci_state) >= thresh
    f, t = np.nonzero(mask)
    return [(int(a), int(b)) for a, b in zip(f, t)]


def ref_overlaps_cache(bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh=1e-4, ao_ovlps="ket"):
    """pywfo/main.py:520-627 (``overlaps_cache``): per state pair, every
    excitation above threshold yields the spin determinants (alpha exc, beta GS)
    and (alpha GS, beta exc) with prefactor c/sqrt(2) (main.py:554-560); the
    state overlap is sum_k' p_k' sum_l' p_l' D_alpha D_beta (main.py:613-626)
    with memoised pair determinants (main.py:579-585)."""
    S_MO = mo_overlap(bra_mos, ket_mos, get_S_AO(bra_mos, ket_mos, ao_ovlps))
    pref = 1 / sqrt(2)
    memo = {}

    def det(k, l):
        key = (k, l)
        if key not in memo:
            memo[key] = sd_pair_det(S_MO, occ, k, l)
        return memo[key]

    def spin_dets(ci_state):
        out = []
        for e in _above(ci_state, ci_thresh):
            p = pref * ci_state[e]
            out.append((p, e, GS))
            out.append((p, GS, e))
        return out

    ovlps = np.zeros((bra_ci.shape[0], ket_ci.shape[0]))
    for i, ci_i in enumerate(bra_ci):
        bra_sds = spin_dets(ci_i)
        for j, ci_j in enumerate(ket_ci):
            ket_sds = spin_dets(ci_j)
            pk = []
            incs = []
            for p_k, ka, kb in bra_sds:
                pk.append(p_k)
                inc = 0.0
                for p_l, la, lb in ket_sds:
                    inc += p_l * det(ka, la) * det(kb, lb)
                incs.append(inc)
            ovlps[i, j] = (np.array(pk) * np.array(incs)).sum() if pk else 0.0
    return ovlps, len(memo)


def ref_overlaps_naive(bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh=1e-4, ao_ovlps="ket"):
    """pywfo/main.py:314-419 (second, live ``overlaps_naive``) and main.py:422-517
    (``overlaps_most_naive``): the arithmetic of ``overlaps_cache`` without the
    memo -- two determinants per (k', l') term.  Only the summation order differs
    between the three, so one restatement serves both."""
    return ref_overlaps_cache(bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh, ao_ovlps)[0]


def ref_overlaps_minus(bra_mos, ket_mos, bra_ci, ket_ci, occ, ci_thresh=1e-4, ao_ovlps="ket"):
    """pywfo/main.py:81-199 (``overlaps``): sum_kl c_k c'_l [D00 D(k,l) - D(k,0) D(0,l)]
    (main.py:194) with its quirks: per-state lists use strict ``>`` (main.py:122-123),
    the gathered minor is scaled by sign_k*sign_l with sign = (-1)**(occ-from+1)
    before the determinant (main.py:154, 176), bra and ket CI must have equal
    shape (main.py:91), and a state pair without excitations raises ValueError
    (main.py:186, unpacking ``zip(*())``)."""
    assert bra_mos.shape == ket_mos.shape
    assert bra_ci.shape == ket_ci.shape
    S_MO = mo_overlap(bra_mos, ket_mos, get_S_AO(bra_mos, ket_mos, ao_ovlps))

    def sign(e):
        return 1 if e[0] < 0 else (-1) ** (occ - e[0] + 1)

    memo = {}

    def det(k, l):
        key = (k, l)
        if key not in memo:
            rows = exc_index_list(k, occ)
            cols = exc_index_list(l, occ)
            mat = S_MO[rows][:, cols] * (sign(k) * sign(l))
            memo[key] = np.linalg.det(mat)
        return memo[key]

    d00 = det(GS, GS)
    out = []
    for ci_i, ci_j in itertools.product(bra_ci, ket_ci):
        bra_exc = _above(ci_i, ci_thresh, strict=True)
        ket_exc = _above(ci_j, ci_thresh, strict=True)
        if not bra_exc or not ket_exc:
            raise ValueError("not enough values to unpack (expected 2, got 0)")
        tot = 0.0
        terms = []
        for k in bra_exc:
            for l in ket_exc:
                terms.append(ci_i[k] * ci_j[l] * (d00 * det(k, l) - det(k, GS) * det(GS, l)))
        tot = np.array(terms).sum()
        out.append(tot)
    return np.reshape(out, (len(bra_ci), -1))


def _wfoverlap_dets(ci, occ, ci_thresh):
    """pywfo/dets.py:6-39 (``get_dets``) + dets.py:42-96 (``expand_det_string``)
    + dets.py:159-168 (``prepare_data``) without strings.

    Every (state, f, t) with |c| > thresh -- duplicates across states kept, as
    ``np.nonzero`` over the state axis produces them (dets.py:11) -- emits two
    determinants whose coefficient rows are ``+c[:, f, t]/sqrt2`` (alpha excited,
    string 'b' at f and 'a' at t) and ``-c[:, f, t]/sqrt2`` (beta excited),
    multiplied by the parity of sorting the string to alpha...beta order.

    Parity (dets.py:73-92): walking the string left to right with ``na`` alphas
    seen, a 'b' adds ``nalpha - na`` and a 'd' adds ``nalpha - (na+1)``.
    """
    n_states, n, v = ci.shape
    st, fr, to = np.nonzero(np.abs(ci) > ci_thresh)
    coeffs, alpha, beta = [], [], []
    base = list(range(n))
    for f, t in zip(fr, to):
        f = int(f)
        t = int(t)
        for spin, fac in (("a", 1.0), ("b", -1.0)):
            # occupation string over n+v orbitals
            chars = ["d"] * n + ["e"] * v
            if spin == "a":
                chars[f], chars[n + t] = "b", "a"
            else:
                chars[f], chars[n + t] = "a", "b"
            na = 0
            perms = 0
            a_inds, b_inds = [], []
            for i, ch in enumerate(chars):
                if ch == "a":
                    a_inds.append(i)
                    na += 1
                elif ch == "b":
                    b_inds.append(i)
                    perms += n - na
                elif ch == "d":
                    a_inds.append(i)
                    b_inds.append(i)
                    na += 1
                    perms += n - na
            coeffs.append(ci[:, f, t] * (fac / 2 ** 0.5) * (-1) ** perms)
            alpha.append(a_inds)
            beta.append(b_inds)
    if not coeffs:
        return np.zeros((0, n_states)), np.zeros((0, n), int), np.zeros((0, n), int)
    return np.array(coeffs), np.array(alpha), np.array(beta)


def ref_wfoverlap(bra_mos, ket_mos, bra_ci, ket_ci, ci_thresh, S_AO):
    """pywfo/dets.py:201-285 (``wfoverlap``): ``wfo = Bci^T (alpha_ovl * beta_ovl) Kci``
    (dets.py:279-283) over the spin-adapted determinant lists of ``get_dets``,
    including the reference's duplicate-entry behaviour for several states."""
    n = bra_ci.shape[1]
    bc, ba, bb = _wfoverlap_dets(bra_ci, n, ci_thresh)
    kc, ka, kb = _wfoverlap_dets(ket_ci, n, ci_thresh)
    S_MO = mo_overlap(bra_mos, ket_mos, S_AO)
    if len(bc) == 0 or len(kc) == 0:
        return np.zeros((bra_ci.shape[0], ket_ci.shape[0]))
    SS = det_from_lists(S_MO, ba, ka) * det_from_lists(S_MO, bb, kb)
    return bc.T @ SS @ kc


# --------------------------------------------------------------------------
# the same mathematics in kernel form (union excitation tables + coefficient
# matrices); used as the checker of the CUDA path and of the sharded reduction
# --------------------------------------------------------------------------
def state_overlaps_bruteforce(S_MO, occ, bra_excs, Cb, ket_excs, Ck, sigma=1.0,
                              k_begin=0, k_end=None):
    """``S = Cb[:, k0:k1] (D00*D + sigma * dk0 d0l^T)[k0:k1] Ck^T`` with every
    determinant from LAPACK (main.py:583-584 per pair; the collapse of the four
    spin terms of main.py:615-626 / main.py:194 is SURVEY.md section 0.4-0.5).
    ``k_begin:k_end`` selects a shard of the bra excitations (partial sum)."""
    Kb = len(bra_excs)
    k_end = Kb if k_end is None else k_end
    sl = slice(k_begin, k_end)
    bra = list(bra_excs[k_begin:k_end])
    D = pair_det_table(S_MO, occ, bra, list(ket_excs)) if bra else np.zeros((0, len(ket_excs)))
    d00 = sd_pair_det(S_MO, occ, GS, GS)
    dk0 = pair_det_table(S_MO, occ, bra, [GS])[:, 0] if bra else np.zeros(0)
    d0l = pair_det_table(S_MO, occ, [GS], list(ket_excs))[0]
    M = d00 * D + sigma * np.outer(dk0, d0l)
    return np.asarray(Cb)[:, sl] @ M @ np.asarray(Ck).T


def appended_sign(exc, occ):
    """``D_appended = s_k s_l D_inplace`` with ``s = (-1)**(occ-1-f)``."""
    return 1.0 if exc[0] < 0 else float((-1) ** (occ - 1 - exc[0]))


def state_overlaps_closed_form(S_MO, occ, bra_excs, Cb, ket_excs, Ck, sigma=1.0):
    """Rank-2 determinant-lemma form of the same sum:
    ``D(k,l) = s_k s_l det(A) (X[t,f] Y[f',t'] + Ai[f',f] W[t,t'])``."""
    n = occ
    A = S_MO[:n, :n]
    B = S_MO[:n, n:]
    C = S_MO[n:, :n]
    Dv = S_MO[n:, n:]
    Ai = np.linalg.inv(A)
    detA = np.linalg.det(A)
    X = C @ Ai
    Y = Ai @ B
    W = Dv - C @ Y
    fb = np.array([e[0] for e in bra_excs])
    tb = np.array([e[1] for e in bra_excs])
    fk = np.array([e[0] for e in ket_excs])
    tk = np.array([e[1] for e in ket_excs])
    sb = np.array([appended_sign(e, n) for e in bra_excs])
    sk = np.array([appended_sign(e, n) for e in ket_excs])
    x = X[tb, fb]
    y = Y[fk, tk]
    Dkl = detA * (np.outer(x, y) + Ai[fk[None, :], fb[:, None]] * W[tb[:, None], tk[None, :]])
    Dkl *= np.outer(sb, sk)
    dk0 = detA * x * sb
    d0l = detA * y * sk
    M = detA * Dkl + sigma * np.outer(dk0, d0l)
    return np.asarray(Cb) @ M @ np.asarray(Ck).T


# --------------------------------------------------------------------------
# synthetic inputs, as the reference's generator makes them
# --------------------------------------------------------------------------
def dummy_mos(num):
    """pywfo/helpers.py:4-15: rows of the Q of a QR of ``np.random.rand`` (global RNG)."""
    q, _ = np.linalg.qr(np.random.rand(num, num), mode="complete")
    return q.T


def bra_ket_dummy_mos(num, occ):
    """pywfo/helpers.py:24-33 (``seed`` is ignored there, so it is absent here)."""
    bra = dummy_mos(num)
    ket, _ = np.linalg.qr(bra.T + np.random.rand(num, num) * 2e-1, mode="complete")
    return num - occ, bra, ket.T


def state_overlaps_factored(S_MO, occ, bra_ci_signed, ket_ci_signed, sigma=1.0):
    """Fully factored closed form for dense (n_states, occ, virt) CI tensors with the
    appended-order signs already folded in (c~ = c * s):
    ``S_ij = det(A)^2 [ (1+sigma) <c~_i, X^T> <c~'_j, Y> + <c~'_j, Ai^T... >]`` written as
    ``sum_{f,t,f',t'} c~_i[f,t] c~'_j[f',t'] (X[t,f] Y[f',t'](1+sigma) + Ai[f',f] W[t,t'])``.
    O(n_states * (n^2 v + n v^2)) -- usable at the full benchmark sizes."""
    n = occ
    A, B, C, Dv = S_MO[:n, :n], S_MO[:n, n:], S_MO[n:, :n], S_MO[n:, n:]
    Ai = np.linalg.inv(A)
    detA = np.linalg.det(A)
    X = C @ Ai
    Y = Ai @ B
    W = Dv - C @ Y
    bx = np.einsum("ift,tf->i", bra_ci_signed, X)
    ky = np.einsum("jft,ft->j", ket_ci_signed, Y)
    # T_i[f', t'] = sum_{f,t} Ai[f', f] c_i[f, t] W[t, t']
    T = np.einsum("gf,ift,tu->igu", Ai, bra_ci_signed, W, optimize=True)
    cross = np.einsum("igu,jgu->ij", T, ket_ci_signed)
    return detA * detA * ((1.0 + sigma) * np.outer(bx, ky) + cross)


def signed_ci(ci, occ):
    """c~[s, f, t] = c[s, f, t] * (-1)**(occ-1-f): appended-order signs folded in."""
    s = (-1.0) ** (occ - 1 - np.arange(occ))
    return ci * s[None, :, None]


def ref_overlap_from_determinants(S_MO, bra_alpha, bra_beta, bra_coeffs, ket_alpha, ket_beta, ket_coeffs):
    """General determinant lists: ``wfo = Bc^T (det_alpha * det_beta) Kc`` with every block
    determinant from LAPACK (pywfo/dets.py:191-198 and :279-283; coefficients already carry
    the string parity of dets.py:42-96)."""
    da = det_from_lists(S_MO, bra_alpha, ket_alpha)
    db = det_from_lists(S_MO, bra_beta, ket_beta)
    return np.asarray(bra_coeffs).T @ (da * db) @ np.asarray(ket_coeffs)
