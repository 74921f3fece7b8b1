"""On-disk formats of the wfoverlap tool chain that pywfo talks to, and the general
determinant-list overlap they feed.

* determinant files ``"<nstates> <n_mo> <ndets>"`` + one ``d/a/b/e`` string and one coefficient
  per state per line (written by pywfo/dets.py:99-110, samples in pywfo/tests/ref/ and
  pywfo/tests/ref_cytosin/),
* Turbomole ``$scfmo ... format(4d20.14)`` MO files (pywfo/tests/ref/a_mo),
* AO-overlap text files ``"n m"`` + matrix (pywfo/tests/ref/ao_ovlp),
* the ``key=value`` input deck of the Fortran program (pywfo/tests/ref_cytosin/ciovl_*.in).

``overlap_from_determinants`` evaluates the state overlaps of two ARBITRARY determinant lists
the way pywfo.dets.wfoverlap does for its own strings (dets.py:119-156 blocks, :191-198
block x block determinants, :279-283 contraction): unique alpha and beta occupation lists are
"blocks", every bra block x ket block determinant comes from the batched pivoted-LU kernel
(``wfo_det_table_host``), the CI contraction is a small dense product.
"""
from __future__ import annotations

import os
import re

import numpy as np


# ------------------------------------------------------------------ determinant strings
def expand_det_string(det: str, n_alpha: int):
    """Occupation string -> (parity prefactor, alpha MO indices, beta MO indices).

    Same convention as pywfo/dets.py:42-96: sorting the spin orbitals to alpha...beta order
    costs one transposition per alpha electron still to come when a beta electron is met."""
    na = 0
    perms = 0
    alpha, beta = [], []
    for i, ch in enumerate(det):
        if ch == "a":
            alpha.append(i)
            na += 1
        elif ch == "b":
            beta.append(i)
            perms += n_alpha - na
        elif ch == "d":
            alpha.append(i)
            beta.append(i)
            na += 1
            perms += n_alpha - na
        elif ch != "e":
            raise Exception(f"Invalid det string. Expected one of 'deab' but got {ch} at index {i}!")
    return (-1) ** perms, alpha, beta


def dets_from_ci(ci: np.ndarray, ci_thresh: float):
    """Spin-adapted determinant strings and coefficient rows of CIS vectors, as
    pywfo/dets.py:6-39 (``get_dets``) emits them: one (alpha excited, beta excited) pair per
    (state, from, to) hit of ``|c| > thresh``, coefficients ``+c/sqrt2`` and ``-c/sqrt2`` of
    ALL states."""
    ci = np.asarray(ci, dtype=np.float64)
    n_states, occ, virt = ci.shape
    _, fr, to = np.nonzero(np.abs(ci) > ci_thresh)
    base = ["d"] * occ + ["e"] * virt
    dets, coeffs = [], []
    for f, t in zip(fr, to):
        col = ci[:, f, t] / 2 ** 0.5
        ab = base.copy()
        ab[f], ab[occ + t] = "b", "a"
        dets.append("".join(ab))
        coeffs.append(col)
        ba = base.copy()
        ba[f], ba[occ + t] = "a", "b"
        dets.append("".join(ba))
        coeffs.append(-col)
    return dets, np.array(coeffs).reshape(len(dets), n_states)


def write_dets(path, dets, coeffs):
    """Determinant file of the wfoverlap program (pywfo/dets.py:99-110)."""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    with open(path, "w") as fh:
        fh.write(f"{coeffs.shape[1]} {len(dets[0])} {len(dets)}\n")
        fh.write("\n".join(d + " ".join(f"{c: >14.10f}" for c in row) for d, row in zip(dets, coeffs)))


def read_dets(path):
    """-> (list of strings, (ndets, nstates) coefficients)."""
    with open(path) as fh:
        n_states, n_mo, n_dets = (int(x) for x in fh.readline().split())
        dets, coeffs = [], []
        for line in fh:
            if not line.strip():
                continue
            dets.append(line[:n_mo])
            coeffs.append([float(x) for x in line[n_mo:].split()])
    coeffs = np.array(coeffs, dtype=np.float64).reshape(len(dets), n_states)
    if len(dets) != n_dets or any(len(d) != n_mo for d in dets):
        raise ValueError(f"{path}: header announces {n_dets} determinants of length {n_mo}")
    return dets, coeffs


# ------------------------------------------------------------------ MOs and AO overlaps
def read_turbomole_mos(path):
    """``$scfmo ... format(4d20.14)`` -> (n_mo, n_ao) array, one MO per row."""
    text = open(path).read()
    m = re.search(r"format\((\d+)d(\d+)\.(\d+)\)", text)
    width = int(m.group(2)) if m else 20
    mos, cur, nsaos = [], [], None
    for line in text.splitlines():
        if line.startswith("$end"):
            break
        if line.startswith("$") or line.startswith("#"):
            continue
        head = re.match(r"\s*\d+\s+\S+\s+eigenvalue=\S+\s+nsaos=(\d+)", line)
        if head:
            if cur:
                mos.append(cur)
            cur, nsaos = [], int(head.group(1))
            continue
        for i in range(0, len(line.rstrip()), width):
            cur.append(float(line[i:i + width].replace("D", "E").replace("d", "e")))
    if cur:
        mos.append(cur)
    out = np.array(mos, dtype=np.float64)
    if nsaos is not None and out.shape[1] != nsaos:
        raise ValueError(f"{path}: expected {nsaos} AOs per MO, got {out.shape[1]}")
    return out


def write_turbomole_mos(path, mos, per_line=4):
    mos = np.asarray(mos, dtype=np.float64)

    def d20(x):
        mant, exp = f"{x:+.13E}".split("E")
        return f"{mant}D{int(exp):+03d}"

    with open(path, "w") as fh:
        fh.write(f"$scfmo    scfconv=7  format({per_line}d20.14)\n# from pywfo_b200\n")
        for i, mo in enumerate(mos, 1):
            fh.write(f"{i:6d}  a      eigenvalue=-.00000000000000D+00   nsaos={mos.shape[1]}\n")
            for j in range(0, len(mo), per_line):
                fh.write("".join(d20(x) for x in mo[j:j + per_line]) + "\n")
        fh.write("$end")


def read_ao_overlap(path):
    with open(path) as fh:
        n, m = (int(x) for x in fh.readline().split())
        data = np.array(fh.read().split(), dtype=np.float64)
    return data.reshape(n, m)


def write_ao_overlap(path, S):
    S = np.asarray(S, dtype=np.float64)
    with open(path, "w") as fh:
        fh.write(f"{S.shape[0]} {S.shape[1]}\n")
        for row in S:
            fh.write(" ".join(f"{x: .18e}" for x in row) + "\n")


def read_ciovl_input(path):
    """``key=value`` input deck of the Fortran program -> dict."""
    out = {}
    for line in open(path):
        line = line.strip()
        if line and "=" in line:
            k, v = line.split("=", 1)
            out[k.strip()] = v.strip()
    return out


# ------------------------------------------------------------------ general determinant lists
def _blocks(index_lists):
    """Unique occupation lists ("blocks", pywfo/dets.py:119-156) and the block of every determinant."""
    uniq = {}
    which = np.empty(len(index_lists), dtype=np.int64)
    for i, lst in enumerate(index_lists):
        which[i] = uniq.setdefault(tuple(lst), len(uniq))
    return np.array(list(uniq.keys()), dtype=np.int32).reshape(len(uniq), -1), which


def overlap_from_determinants(bra_mos, ket_mos, S_AO, bra_dets, bra_coeffs, ket_dets, ket_coeffs, device=None,
                              return_info=False):
    """State overlaps ``<bra_i|ket_j>`` of two determinant lists (strings + (ndets, nstates)
    coefficients).  Closed- or open-shell: alpha and beta blocks are treated separately."""
    from . import _lib

    ctx = _lib.context(device)
    n_alpha = bra_dets[0].count("a") + bra_dets[0].count("d")
    n_beta = bra_dets[0].count("b") + bra_dets[0].count("d")

    def prep(dets, coeffs):
        pref, al, be = zip(*(expand_det_string(d, n_alpha) for d in dets))
        if any(len(a) != n_alpha for a in al) or any(len(b) != n_beta for b in be):
            raise ValueError("all determinants must have the same number of alpha and beta electrons")
        return np.asarray(coeffs, dtype=np.float64) * np.array(pref, dtype=np.float64)[:, None], _blocks(al), _blocks(be)

    bc, (ba_blk, ba_of), (bb_blk, bb_of) = prep(bra_dets, bra_coeffs)
    kc, (ka_blk, ka_of), (kb_blk, kb_of) = prep(ket_dets, ket_coeffs)
    s_mo = ctx.smo_host(bra_mos, S_AO, ket_mos)                       # dets.py:264
    # block determinants (dets.py:191-198, with super-block reuse: dets.py:119-156) and the contraction
    # (dets.py:279-283) in one library call: neither table nor the determinant x determinant matrix
    # of dets.py:280 comes back to the host
    alpha = (ba_blk, ka_blk, ba_of, ka_of) if n_alpha else None
    beta = (bb_blk, kb_blk, bb_of, kb_of) if n_beta else None
    out, info = ctx.overlap_from_lists_host(s_mo, alpha, beta, bc, kc)
    if return_info:
        return out, {"super_blocks": int(info[0]), "blocks_via_super": int(info[1]), "blocks_redone": int(info[2])}
    return out


def overlap_from_deck(directory, a_mo="a_mo", b_mo="b_mo", ao_ovlp="ao_ovlp", a_det="dets_a_spin_adapted",
                      b_det="dets_b_spin_adapted", device=None):
    """Run a wfoverlap-style input deck from disk (file names as in pywfo/tests/ref/)."""
    j = lambda f: os.path.join(directory, f)
    bra_dets, bra_c = read_dets(j(a_det))
    ket_dets, ket_c = read_dets(j(b_det))
    return overlap_from_determinants(read_turbomole_mos(j(a_mo)), read_turbomole_mos(j(b_mo)),
                                     read_ao_overlap(j(ao_ovlp)), bra_dets, bra_c, ket_dets, ket_c, device=device)
