"""On-disk formats (pywfo_b200/formats.py): parsers against the reference's bundled deck,
writers by round trip, string expansion against the oracle.  CPU except the last tests."""
import os

import numpy as np
import pytest

from oracle import pywfo_oracle as orc
from pywfo_b200 import formats as fm
from conftest import GOLDEN, load_golden

DECK = os.path.join(GOLDEN, "ref_deck")


def test_read_reference_deck():
    a, b, S = fm.read_turbomole_mos(os.path.join(DECK, "a_mo")), fm.read_turbomole_mos(os.path.join(DECK, "b_mo")), \
        fm.read_ao_overlap(os.path.join(DECK, "ao_ovlp"))
    assert a.shape == b.shape == S.shape == (4, 4)
    assert a[0, 0] == pytest.approx(-3.5264951801226e-01, abs=1e-15)
    np.testing.assert_allclose(a @ S @ a.T, np.eye(4), atol=1e-12)      # rows are orthonormal MOs
    dets, coeffs = fm.read_dets(os.path.join(DECK, "dets_a_spin_adapted"))
    assert dets == ["ddee", "dabe", "dbae", "daeb", "dbea"] and coeffs.shape == (5, 3)
    assert coeffs[0, 0] == 1.0 and coeffs[1, 1] == pytest.approx(0.70711)


def test_round_trips(tmp_path):
    rng = np.random.default_rng(0)
    mos = rng.standard_normal((7, 7))
    fm.write_turbomole_mos(tmp_path / "mos", mos)
    np.testing.assert_allclose(fm.read_turbomole_mos(tmp_path / "mos"), mos, rtol=1e-13)
    S = rng.standard_normal((7, 7))
    fm.write_ao_overlap(tmp_path / "ao", S)
    np.testing.assert_array_equal(fm.read_ao_overlap(tmp_path / "ao"), S)
    ci = rng.standard_normal((2, 3, 4))
    dets, coeffs = fm.dets_from_ci(ci, 0.3)
    fm.write_dets(tmp_path / "dets", dets, coeffs)
    d2, c2 = fm.read_dets(tmp_path / "dets")
    assert d2 == dets
    np.testing.assert_allclose(c2, coeffs, atol=1e-10)
    (tmp_path / "in").write_text("a_mo=mos.1\nncore=0\n\nb_det=dets.2\n")
    assert fm.read_ciovl_input(tmp_path / "in") == {"a_mo": "mos.1", "ncore": "0", "b_det": "dets.2"}


def test_string_expansion_matches_reference_semantics():
    """dets_from_ci + expand_det_string reproduce the oracle's restatement of dets.py:6-96."""
    g = load_golden("r_sparse")
    ci, thr, occ = g["bra_ci"], float(g["ci_thresh"]), int(g["occ"])
    want_c, want_a, want_b = orc._wfoverlap_dets(ci, occ, thr)
    dets, coeffs = fm.dets_from_ci(ci, thr)
    pref, al, be = zip(*(fm.expand_det_string(d, occ) for d in dets))
    np.testing.assert_array_equal(np.array(al), want_a)
    np.testing.assert_array_equal(np.array(be), want_b)
    np.testing.assert_allclose(coeffs * np.array(pref)[:, None], want_c, rtol=1e-15)
    with pytest.raises(Exception, match="Invalid det string"):
        fm.expand_det_string("ddxe", 2)


def _oracle_from_strings(bra_mos, ket_mos, S_AO, bd, bc, kd, kc):
    n_alpha = bd[0].count("a") + bd[0].count("d")
    S = orc.mo_overlap(bra_mos, ket_mos, S_AO)
    pb, ba, bb = zip(*(fm.expand_det_string(d, n_alpha) for d in bd))
    pk, ka, kb = zip(*(fm.expand_det_string(d, n_alpha) for d in kd))
    return orc.ref_overlap_from_determinants(S, ba, bb, bc * np.array(pb)[:, None], ka, kb, kc * np.array(pk)[:, None])


@pytest.mark.gpu
def test_deck_on_gpu():
    """The bundled 4-MO deck end to end from disk (3x3 states incl. the ground state)."""
    got = fm.overlap_from_deck(DECK)
    a, b, S = (fm.read_turbomole_mos(os.path.join(DECK, "a_mo")), fm.read_turbomole_mos(os.path.join(DECK, "b_mo")),
               fm.read_ao_overlap(os.path.join(DECK, "ao_ovlp")))
    bd, bc = fm.read_dets(os.path.join(DECK, "dets_a_spin_adapted"))
    kd, kc = fm.read_dets(os.path.join(DECK, "dets_b_spin_adapted"))
    want = _oracle_from_strings(a, b, S, bd, bc, kd, kc)
    assert got.shape == (3, 3)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["r_sparse", "r_n9", "r_n17"])
def test_general_determinant_lists_reproduce_wfoverlap(name):
    """Strings -> blocks -> batched LU tables -> contraction == pywfo.dets.wfoverlap (golden)."""
    g = load_golden(name)
    thr = float(g["ci_thresh"])
    S_AO = orc.get_S_AO(g["bra_mos"], g["ket_mos"], "bra")
    bd, bc = fm.dets_from_ci(g["bra_ci"], thr)
    kd, kc = fm.dets_from_ci(g["ket_ci"], thr)
    got = fm.overlap_from_determinants(g["bra_mos"], g["ket_mos"], S_AO, bd, bc, kd, kc)
    want = g["ref_wfoverlap"]
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
