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


CYT = os.path.join(GOLDEN, "ref_cytosin")


def test_read_cytosine_deck():
    deck = fm.read_ciovl_input(os.path.join(CYT, "ciovl_no_gs_loose_zeroed.in"))
    assert deck["a_mo"] == "mos.1" and deck["b_det"] == "dets_no_gs_loose_zeroed.2" and deck["mix_aoovl"] == "ao_ovl"
    mos = fm.read_turbomole_mos(os.path.join(CYT, deck["a_mo"]))
    S = fm.read_ao_overlap(os.path.join(CYT, deck["mix_aoovl"]))
    assert mos.shape == (137, 137) and S.shape == (137, 137)
    dets, coeffs = fm.read_dets(os.path.join(CYT, deck["a_det"]))
    assert len(dets) == 6 and coeffs.shape == (6, 2) and all(len(d) == 137 for d in dets)
    # MOs are orthonormal in the AO metric of their own geometry only approximately for a mixed overlap
    assert np.isfinite(mos).all() and np.isfinite(S).all()


@pytest.mark.gpu
def test_cytosine_deck_from_disk_on_gpu():
    """pywfo/tests/ref_cytosin/ciovl_no_gs_loose_zeroed.in end to end from disk: 137 MOs, 29 electrons per spin,
    open-shell strings.  Checker: oracle (LAPACK determinants per block pair) on the same parsed inputs."""
    deck = fm.read_ciovl_input(os.path.join(CYT, "ciovl_no_gs_loose_zeroed.in"))
    got = fm.overlap_from_deck(CYT, a_mo=deck["a_mo"], b_mo=deck["b_mo"], ao_ovlp=deck["mix_aoovl"],
                               a_det=deck["a_det"], b_det=deck["b_det"])
    a, b = fm.read_turbomole_mos(os.path.join(CYT, deck["a_mo"])), fm.read_turbomole_mos(os.path.join(CYT, deck["b_mo"]))
    S = fm.read_ao_overlap(os.path.join(CYT, deck["mix_aoovl"]))
    bd, bc = fm.read_dets(os.path.join(CYT, deck["a_det"]))
    kd, kc = fm.read_dets(os.path.join(CYT, deck["b_det"]))
    n_alpha = bd[0].count("a") + bd[0].count("d")
    Smo = orc.mo_overlap(a, b, S)
    pb, ba, bb = zip(*(fm.expand_det_string(d, n_alpha) for d in bd))
    pk, ka, kb = zip(*(fm.expand_det_string(d, n_alpha) for d in kd))
    want = orc.ref_overlap_from_determinants(Smo, ba, bb, bc * np.array(pb)[:, None], ka, kb, kc * np.array(pk)[:, None])
    assert got.shape == (2, 2)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
    assert np.abs(want).max() > 0.1   # neighbouring geometries: the states overlap strongly


def _cis_strings(n_mo, occ, excitations):
    """Closed-shell reference with single alpha excitations (f -> occ + t) as occupation strings."""
    out = []
    for f, t in excitations:
        s = ["d"] * occ + ["e"] * (n_mo - occ)
        s[f] = "b"
        s[occ + t] = "a"
        out.append("".join(s))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("occ,n_mo", [(4, 30), (9, 40), (33, 60), (64, 90), (100, 125), (128, 150)])
def test_super_blocks_share_one_factorisation(occ, n_mo):
    """Lists whose blocks differ only in the LAST orbital (pywfo/dets.py:119-156): the bra excitations out of the
    two highest occupied orbitals into every virtual give two super-blocks of n_virt members each; they must
    be served by the shared factorisation and equal the brute-force / LAPACK values."""
    from pywfo_b200 import helpers
    np.random.seed(occ)
    _, bra_mos, ket_mos = helpers.get_bra_ket_dummy_mos(n_mo, occ)
    S_AO = orc.get_S_AO(bra_mos, ket_mos, "ket")
    virt = n_mo - occ
    rng = np.random.default_rng(occ)
    bra_exc = [(occ - 1, t) for t in range(virt)] + [(occ - 2, t) for t in range(virt)] + [(0, 1), (1, 0)]
    ket_exc = [(int(rng.integers(occ)), int(rng.integers(virt))) for _ in range(9)]
    ket_exc = list(dict.fromkeys(ket_exc))
    bd, kd = _cis_strings(n_mo, occ, bra_exc), _cis_strings(n_mo, occ, ket_exc)
    bc, kc = rng.standard_normal((len(bd), 3)), rng.standard_normal((len(kd), 2))
    got, info = fm.overlap_from_determinants(bra_mos, ket_mos, S_AO, bd, bc, kd, kc, return_info=True)
    want = _oracle_from_strings(bra_mos, ket_mos, S_AO, bd, bc, kd, kc)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
    assert info["super_blocks"] >= 2 and info["blocks_via_super"] >= 2 * virt and info["blocks_redone"] == 0


@pytest.mark.gpu
def test_super_block_with_singular_representative_is_redone():
    """The shared factorisation uses the first member's matrix; when that one is singular (here: its last row
    equals one of the shared rows) the super-block is recomputed by brute force, members stay correct."""
    occ, n_mo = 6, 20
    rng = np.random.default_rng(5)
    q, _ = np.linalg.qr(rng.standard_normal((n_mo, n_mo)))
    S = q + 0.05 * rng.standard_normal((n_mo, n_mo))
    S[occ + 0] = S[1]                     # virtual 0 == occupied 1: every determinant holding both is zero
    virt = n_mo - occ
    bra_exc = [(occ - 1, t) for t in range(virt)]     # alpha block: [0..occ-2, occ+t]: representative t = 0
    ket_exc = [(2, 3), (0, 0), (5, 7)]
    bd, kd = _cis_strings(n_mo, occ, bra_exc), _cis_strings(n_mo, occ, ket_exc)
    bc, kc = rng.standard_normal((len(bd), 2)), rng.standard_normal((len(kd), 2))
    eye = np.eye(n_mo)
    got, info = fm.overlap_from_determinants(eye, eye, S, bd, bc, kd, kc, return_info=True)   # S_MO = S
    want = _oracle_from_strings(eye, eye, S, bd, bc, kd, kc)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
    assert info["super_blocks"] >= 1 and info["blocks_redone"] >= virt
