"""Host-side index/coefficient builder (pywfo_b200/excitations.py) against the
golden vectors: the tables it builds, fed to the ORACLE's kernel-form sum,
must reproduce every reference entry point.  CPU only."""
import numpy as np
import pytest

from oracle import pywfo_oracle as orc
from pywfo_b200 import excitations as ex
from conftest import RANDOM_CASES, load_golden


def _kernel_form(g, semantics, sigma, closed=False, nmin=None, ao="ket"):
    occ = int(g["occ"])
    thr = float(g["ci_thresh"])
    bci, kci = g["bra_ci"], g["ket_ci"]
    if nmin:
        bci, kci = bci[:nmin], kci[:nmin]
    S = orc.mo_overlap(g["bra_mos"], g["ket_mos"], orc.get_S_AO(g["bra_mos"], g["ket_mos"], ao))
    bt = ex.build_tables(bci, thr, semantics)
    kt = ex.build_tables(kci, thr, semantics, order="locality")
    bra = [tuple(map(int, e)) for e in bt.exc]
    ket = [tuple(map(int, e)) for e in kt.exc]
    fn = orc.state_overlaps_closed_form if closed else orc.state_overlaps_bruteforce
    return fn(S, occ, bra, bt.coeff, ket, kt.coeff, sigma)


@pytest.mark.parametrize("name", RANDOM_CASES)
@pytest.mark.parametrize("closed", [False, True])
def test_tables_reproduce_reference(name, closed):
    g = load_golden(name)
    scale = np.abs(g["ref_overlaps_cache"]).max()
    np.testing.assert_allclose(_kernel_form(g, "cache", 1.0, closed), g["ref_overlaps_cache"],
                               rtol=1e-10, atol=1e-12 * scale)
    nmin = min(g["bra_ci"].shape[0], g["ket_ci"].shape[0])
    np.testing.assert_allclose(_kernel_form(g, "minus", -1.0, closed, nmin), g["ref_overlaps"],
                               rtol=1e-10, atol=1e-12 * scale)
    np.testing.assert_allclose(_kernel_form(g, "wfoverlap", 1.0, closed, ao="bra"), g["ref_wfoverlap"],
                               rtol=1e-10, atol=1e-12 * np.abs(g["ref_wfoverlap"]).max())


def test_threshold_conventions():
    """>= for the cache family (main.py:539-540), > for overlaps (main.py:122-123)."""
    ci = np.zeros((2, 2, 3))
    ci[0, 0, 1] = 0.5
    ci[0, 1, 2] = 0.1      # exactly on the threshold
    ci[1, 1, 0] = -0.7
    t_ge = ex.build_tables(ci, 0.1, "cache")
    t_gt = ex.build_tables(ci, 0.1, "minus")
    assert [tuple(e) for e in t_ge.exc] == [(0, 1), (1, 0), (1, 2)]
    assert [tuple(e) for e in t_gt.exc] == [(0, 1), (1, 0)]
    # per-state zeros: state 1 does not keep (0,1)
    assert t_ge.coeff[1, 0] == 0.0 and t_ge.coeff[0, 0] == 0.5 and t_ge.coeff[0, 2] == 0.1


def test_minus_sign_folding():
    """det(sign*M) = sign^occ det(M) (main.py:154,176): no sign for even occ."""
    rng = np.random.default_rng(0)
    for occ in (3, 4):
        ci = rng.standard_normal((1, occ, 2))
        t = ex.build_tables(ci, 0.0, "minus")
        f = t.exc[:, 0]
        want = ci[0, t.exc[:, 0], t.exc[:, 1]] * (1.0 if occ % 2 == 0 else (-1.0) ** (occ - f + 1))
        np.testing.assert_array_equal(t.coeff[0], want)


def test_wfoverlap_duplicates():
    """dets.py:11: np.nonzero runs over the state axis -> one entry per (state, f, t) hit."""
    ci = np.zeros((2, 2, 2))
    ci[0, 1, 0] = 0.6
    ci[1, 1, 0] = 0.8
    ci[1, 0, 1] = 0.3
    t = ex.build_tables(ci, 0.1, "wfoverlap")
    assert [tuple(e) for e in t.exc] == [(1, 0), (0, 1), (1, 0)]
    np.testing.assert_allclose(t.coeff[:, 0], ci[:, 1, 0] * (-1.0) ** (2 - 1))


def test_empty_and_minus_check():
    ci = np.zeros((2, 3, 4))
    assert ex.build_tables(ci, 1e-4).K == 0
    ci[0, 0, 0] = 1.0
    with pytest.raises(ValueError):
        ex.check_minus_inputs(ci, ci, 1e-4)


def test_index_lists_match_reference_ordering():
    exc = np.array([[0, 0], [2, 1], [4, 3], [-1, -1]])
    got = ex.exc_index_lists(exc, 5)
    want = [orc.exc_index_list(tuple(e), 5) for e in exc]
    np.testing.assert_array_equal(got, np.array(want))


def test_locality_order_is_a_permutation_and_sum_invariant():
    g = load_golden("r_water")
    bt = ex.build_tables(g["ket_ci"], 0.0, "cache")
    kt = ex.build_tables(g["ket_ci"], 0.0, "cache", order="locality")
    assert sorted(map(tuple, bt.exc)) == sorted(map(tuple, kt.exc))
    lut = {tuple(e): i for i, e in enumerate(bt.exc)}
    perm = [lut[tuple(e)] for e in kt.exc]
    np.testing.assert_array_equal(bt.coeff[:, perm], kt.coeff)
    blocks = kt.exc[:, 1] // 32
    assert np.all(np.diff(blocks) >= 0)


@pytest.mark.parametrize("K,world", [(0, 2), (1, 2), (7, 2), (8, 8), (40001, 8), (5, 8)])
def test_shard_bounds(K, world):
    b = [ex.shard_bounds(K, world, r) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == K
    assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in b]
    assert max(sizes) - min(sizes) <= 1


def test_excitation_tables_are_range_checked():
    """(from, to) entries outside the occupied / virtual ranges raise IndexError on the host
    (what the reference's own indexing does, pywfo/main.py:565-576) instead of reaching the device."""
    from pywfo_b200._lib import _check_exc
    ok = np.array([[0, 0], [4, 18], [-1, -1]], dtype=np.int32)
    _check_exc("bra", ok, 5, 19)
    _check_exc("bra", ok[:0], 5, 19)
    for bad in ([5, 0], [0, 19], [-1, 3], [2, -1], [7, 40]):
        with pytest.raises(IndexError):
            _check_exc("ket", np.array([[1, 1], bad], dtype=np.int32), 5, 19)


def test_trajectory_pairs_are_renumbered_into_the_cycles_they_touch():
    """a rank uploads only the cycles its pairs reference."""
    from pywfo_b200.trajectory import consecutive_pairs, referenced_cycles
    pairs = np.array([[7, 2], [2, 3], [9, 7]], dtype=np.int32)
    cycles, local = referenced_cycles(pairs)
    assert cycles.tolist() == [2, 3, 7, 9]
    assert np.array_equal(cycles[local], pairs)
    allp = consecutive_pairs(9)
    cyc, loc = referenced_cycles(allp[np.arange(1, 8, 3)])       # the pairs rank 1 of 3 is dealt
    assert cyc.tolist() == [1, 2, 4, 5, 7, 8] and np.array_equal(cyc[loc], allp[[1, 4, 7]])
    cyc, loc = referenced_cycles(allp)
    assert cyc.size == 9 and np.array_equal(loc, allp)


def test_precompute_accepts_the_reference_block_tuples():
    """pywfo.dets.precompute takes BlockResult tuples (dets.py:114-156, 191-198); plain index arrays work too."""
    from collections import namedtuple
    from pywfo_b200.dets import _blocks_of
    BlockResult = namedtuple("BlockResult", "block_num super_block_num blocks block_map super_map ci_block_map sort_inds")
    blocks = [[0, 1, 2], [0, 1, 5], [0, 2, 7]]
    res = BlockResult(3, 2, blocks, [0, 1, 2], [0, 2], np.array([0, 1, 2]), [0, 1, 2])
    assert np.array_equal(_blocks_of(res), np.array(blocks, dtype=np.int32))
    assert np.array_equal(_blocks_of(np.array(blocks)), np.array(blocks, dtype=np.int32))
    assert _blocks_of(res).dtype == np.int32
