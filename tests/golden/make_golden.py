"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run from the repo root:  python tests/golden/make_golden.py
It imports pywfo from /root/reference (read-only; never copied), silences its
prints and logger, runs the reference entry points on
  * the two numeric goldens of pywfo/tests/test_pywfo.py (RNG recipes of
    SURVEY.md section 4),
  * the bundled H2O2 fixture (read without h5py: old-format HDF5 with contiguous
    little-endian float64 datasets at fixed offsets),
  * seeded random cases (odd/even occ, unequal state counts, thresholds that
    separate ``>=`` from ``>``),
and stores inputs + reference outputs.  /root/reference does not exist on the
GPU box, so tests only ever read the committed .npz files.
"""
import contextlib
import io
import logging
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())  # pywfo opens ./pywfo.log lazily
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import pywfo.main as main
        import pywfo.dets as dets
        import pywfo.helpers as helpers
    logging.getLogger("pywfo").handlers.clear()
    os.chdir(cwd)
    return main, dets, helpers


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def read_h2o2():
    raw = open(os.path.join(REF, "pywfo/tests/h2o2_overlap_data.h5"), "rb").read()
    mo = np.frombuffer(raw, dtype="<f8", count=2 * 38 * 38, offset=2048).reshape(2, 38, 38)
    ci = np.frombuffer(raw, dtype="<f8", count=2 * 4 * 9 * 29, offset=25152).reshape(2, 4, 9, 29)
    # sanity: rows are MOs, orthonormal w.r.t. S_AO = C^-1 C^-T by construction;
    # CI vectors are close to normalised
    assert np.all(np.abs(np.linalg.norm(ci.reshape(2, 4, -1), axis=2) - 1) < 0.1)
    return mo.copy(), ci.copy()


def random_ci(rng, n_states, occ, virt, sparsity=0.0):
    c = rng.standard_normal((n_states, occ, virt))
    if sparsity:
        c[rng.random(c.shape) < sparsity] = 0.0
    c /= np.linalg.norm(c.reshape(n_states, -1), axis=1)[:, None, None]
    return c


def main():
    m, d, h = load_reference()
    out = {}

    # --- golden 1: test_simple (test_pywfo.py:60-94), value 0.97008177 --------
    np.random.seed(20180325)
    virt, bra, ket = h.get_bra_ket_dummy_mos(2, 1)
    ci = np.ones((1, 1, 1))
    g = dict(bra_mos=bra, ket_mos=ket, bra_ci=ci, ket_ci=ci, occ=1, ci_thresh=1e-4)
    for name in ("overlaps_naive", "overlaps_most_naive", "overlaps_cache"):
        g["ref_" + name] = quiet(getattr(m, name), bra, ket, ci, ci, occ=1)
    assert abs(g["ref_overlaps_cache"][0, 0] - 0.97008177) < 1e-6
    np.savez(os.path.join(HERE, "simple.npz"), **g)

    # --- golden 2: test_more_virtual_mos (test_pywfo.py:97-125) ---------------
    # RNG history of a full pytest run of the reference's module
    h.get_dummy_mos(10, 20180325)
    np.random.rand(10, 10)
    h.get_bra_ket_dummy_mos(2, 1)
    h.get_bra_ket_dummy_mos(2, 1)
    virt, bra, ket = h.get_bra_ket_dummy_mos(4, 2)
    cis = np.zeros((2, 2, 2, 2))
    cis[0, 0, 1, 0] = 1
    cis[0, 1, 1, 1] = 1
    cis[1, 0, 1, 1] = 1
    cis[1, 1, 1, 0] = 1
    g = dict(bra_mos=bra, ket_mos=ket, bra_ci=cis[0], ket_ci=cis[1], occ=2, ci_thresh=1e-4)
    for name in ("overlaps_naive", "overlaps_most_naive", "overlaps_cache", "overlaps"):
        g["ref_" + name] = quiet(getattr(m, name), bra, ket, cis[0], cis[1], occ=2)
    want = np.array((0.11025145, 0.97114924, 0.97595963, -0.09630209)).reshape(2, 2)
    assert np.allclose(g["ref_overlaps_most_naive"], want, atol=1e-7), g["ref_overlaps_most_naive"]
    np.savez(os.path.join(HERE, "more_virtual_mos.npz"), **g)

    # --- H2O2 fixture (test_pywfo.py:128-156) ---------------------------------
    mo, ci = read_h2o2()
    bra, ket = mo
    g = dict(bra_mos=bra, ket_mos=ket, bra_ci=ci[0], ket_ci=ci[1], occ=9)
    b1, k1 = ci[0][:1], ci[1][:1]
    g["ref_cache_state0_3e-2"] = quiet(m.overlaps_cache, bra, ket, b1, k1, 9, ci_thresh=3e-2)
    g["ref_cache_4x4_1e-2"] = quiet(m.overlaps_cache, bra, ket, ci[0], ci[1], 9, ci_thresh=1e-2)
    g["ref_minus_4x4_1e-2"] = quiet(m.overlaps, bra, ket, ci[0], ci[1], 9, ci_thresh=1e-2)
    g["ref_cache_4x4_1e-4"] = quiet(m.overlaps_cache, bra, ket, ci[0], ci[1], 9, ci_thresh=1e-4)
    S_AO = quiet(m.get_S_AO, bra, ket, "ket")
    g["ref_wfoverlap_4x4_1e-2"] = quiet(d.wfoverlap, bra, ket, ci[0], ci[1], 1e-2, S_AO, debug=False)
    g["ref_wfoverlap_1x1_3e-2"] = quiet(d.wfoverlap, bra, ket, b1, k1, 3e-2, S_AO, debug=False)
    np.savez(os.path.join(HERE, "h2o2.npz"), **g)

    # --- seeded random cases ---------------------------------------------------
    cases = [
        # name, seed, n_mo, occ, n_bra, n_ket, thresh, sparsity
        ("r_odd", 11, 7, 3, 2, 3, 1e-4, 0.0),
        ("r_even", 12, 8, 4, 3, 2, 1e-4, 0.0),
        ("r_sparse", 13, 12, 5, 4, 4, 0.12, 0.5),
        ("r_water", 14, 24, 5, 3, 3, 0.05, 0.0),
        ("r_n9", 15, 20, 9, 2, 2, 0.08, 0.0),
        ("r_n17", 16, 40, 17, 2, 2, 0.05, 0.0),
        ("r_n34", 17, 60, 34, 2, 2, 0.04, 0.0),
    ]
    for name, seed, n_mo, occ, nb, nk, thr, sp in cases:
        np.random.seed(seed)
        virt, bra, ket = h.get_bra_ket_dummy_mos(n_mo, occ)
        rng = np.random.default_rng(seed)
        bci = random_ci(rng, nb, occ, virt, sp)
        kci = random_ci(rng, nk, occ, virt, sp)
        # put one coefficient exactly on the threshold to separate >= from >
        bci[0, 0, 0] = thr
        kci[0, occ - 1, virt - 1] = -thr
        g = dict(bra_mos=bra, ket_mos=ket, bra_ci=bci, ket_ci=kci, occ=occ, ci_thresh=thr)
        g["ref_overlaps_cache"] = quiet(m.overlaps_cache, bra, ket, bci, kci, occ, ci_thresh=thr)
        if n_mo <= 12:
            g["ref_overlaps_naive"] = quiet(m.overlaps_naive, bra, ket, bci, kci, occ, ci_thresh=thr)
            g["ref_overlaps_most_naive"] = quiet(m.overlaps_most_naive, bra, ket, bci, kci, occ, ci_thresh=thr)
        # overlaps() needs equal CI shapes
        nmin = min(nb, nk)
        g["ref_overlaps"] = quiet(m.overlaps, bra, ket, bci[:nmin], kci[:nmin], occ, ci_thresh=thr)
        S_AO = quiet(m.get_S_AO, bra, ket, "bra")
        g["ref_overlaps_cache_bra"] = quiet(m.overlaps_cache, bra, ket, bci, kci, occ, ci_thresh=thr, ao_ovlps="bra")
        g["ref_wfoverlap"] = quiet(d.wfoverlap, bra, ket, bci, kci, thr, S_AO, debug=False)
        g["ref_moovlp"] = m.moovlp(bra, ket, S_AO)
        np.savez(os.path.join(HERE, name + ".npz"), **g)
        print(name, g["ref_overlaps_cache"].ravel()[:3])


if __name__ == "__main__":
    main()
