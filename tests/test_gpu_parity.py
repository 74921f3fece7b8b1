"""Parity of the CUDA path (through the C ABI) with the oracle and the golden
vectors.  Needs a B200: run with ``-m gpu``.

Tolerance: north_star asks for <= 1e-10 relative error in FP64 and the sign of
every state overlap; elements below 1e-10 * max|S| are compared absolutely.
"""
import numpy as np
import pytest

from oracle import pywfo_oracle as orc
from conftest import RANDOM_CASES, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def assert_parity(got, want, rtol=RTOL):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape
    scale = np.abs(want).max() if want.size else 0.0
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * scale + 1e-300)
    big = np.abs(want) > rtol * scale
    assert np.array_equal(np.sign(got[big]), np.sign(want[big]))


@pytest.fixture(scope="module")
def ctx():
    from pywfo_b200 import _lib
    return _lib.context(0)


def _smo(g, ao="ket"):
    return orc.mo_overlap(g["bra_mos"], g["ket_mos"], orc.get_S_AO(g["bra_mos"], g["ket_mos"], ao))


# ---------------------------------------------------------------- stage 1
@pytest.mark.parametrize("name", ["simple", "more_virtual_mos", "r_odd", "r_n17", "r_n34", "h2o2"])
def test_mo_overlap(ctx, name):
    from pywfo_b200 import main
    g = load_golden(name)
    S_AO = orc.get_S_AO(g["bra_mos"], g["ket_mos"], "ket")
    want = orc.mo_overlap(g["bra_mos"], g["ket_mos"], S_AO)
    for fn in (main.moovlp, main.moovlp_expl, main.moovlp_dots):
        np.testing.assert_allclose(fn(g["bra_mos"], g["ket_mos"], S_AO), want, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(main.get_S_AO(g["bra_mos"], g["ket_mos"], "ket"), S_AO, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("dim", [500, 1000])
def test_mo_overlap_identity_big(ctx, dim):
    """pywfo/tests/test_pywfo.py:39-57 (test_big_mo_overlaps: dim = 1000 upstream, same construction and the
    same three variants and tolerance; the MOs are regenerated from the seed, upstream's are unseeded)."""
    from pywfo_b200 import helpers, main
    mos = helpers.get_dummy_mos(dim, 20180325)
    inv = np.linalg.inv(mos)
    S_AO = inv.dot(inv.T)
    for fn in (main.moovlp, main.moovlp_expl, main.moovlp_dots):
        np.testing.assert_allclose(fn(mos, mos, S_AO), np.eye(dim), atol=1e-8)
    # and with the inverse / S_AO from the device too (get_S_AO, main.py:301-311)
    S_dev = main.get_S_AO(mos, mos, "ket")
    np.testing.assert_allclose(main.moovlp(mos, mos, S_dev), np.eye(dim), atol=1e-8)


def test_mo_overlap_rectangular(ctx):
    rng = np.random.default_rng(1)
    a, s, b = rng.standard_normal((37, 70)), rng.standard_normal((70, 65)), rng.standard_normal((129, 65))
    np.testing.assert_allclose(ctx.smo_host(a, s, b), a @ s @ b.T, rtol=1e-12, atol=1e-11)


# ---------------------------------------------------------------- stage 2
@pytest.mark.parametrize("n", [1, 2, 3, 5, 9, 17, 32, 33, 34, 64, 65, 100, 128])
def test_det_table_vs_lapack(ctx, n):
    rng = np.random.default_rng(n)
    n_mo = n + 7
    q, _ = np.linalg.qr(rng.standard_normal((n_mo, n_mo)))
    S = q + 0.05 * rng.standard_normal((n_mo, n_mo))
    rows = np.array([rng.permutation(n_mo)[:n] for _ in range(5)], dtype=np.int32)
    cols = np.array([rng.permutation(n_mo)[:n] for _ in range(4)], dtype=np.int32)
    want = orc.det_from_lists(S, rows, cols)
    got = ctx.det_table_host(S, rows, cols)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12 * np.abs(want).max())
    assert np.array_equal(np.sign(got), np.sign(want))


@pytest.mark.parametrize("variant", ["left", "right"])
@pytest.mark.parametrize("n", [33, 39, 40, 41, 47, 48, 56, 57, 63, 64, 65, 72, 80, 88, 96, 97, 100, 104, 105, 112, 120, 127, 128])
@pytest.mark.parametrize("kind", ["near_orthogonal", "random"])
def test_both_lu_kernels_every_size(ctx, variant, n, kind):
    """Both warp-per-determinant LU kernels (left-looking on chip, right-looking L2 scratch) at every row-tile
    count and ragged size; `random` matrices force a row exchange in nearly every column step."""
    rng = np.random.default_rng(1000 + n)
    n_mo = n + 9
    if kind == "random":
        S = rng.standard_normal((n_mo, n_mo))
    else:
        q, _ = np.linalg.qr(rng.standard_normal((n_mo, n_mo)))
        S = q + 0.02 * rng.standard_normal((n_mo, n_mo))
    rows = np.array([rng.permutation(n_mo)[:n] for _ in range(7)], dtype=np.int32)
    cols = np.array([rng.permutation(n_mo)[:n] for _ in range(6)], dtype=np.int32)
    rows[1] = np.sort(rows[1])
    cols[1] = np.sort(cols[1])
    sgn, logdet = np.linalg.slogdet(S[rows[:, None, :, None], cols[None, :, None, :]])
    ctx.set_lu_variant(variant)
    try:
        got = ctx.det_table_host(S, rows, cols)
    finally:
        ctx.set_lu_variant("auto")
    assert np.array_equal(np.sign(got), sgn)
    np.testing.assert_allclose(np.log(np.abs(got)), logdet, rtol=0, atol=1e-10)


@pytest.mark.parametrize("variant", ["left", "right"])
@pytest.mark.parametrize("n", [34, 64, 100])
def test_lu_kernels_singular_minors(ctx, variant, n):
    """Exactly singular minors (a repeated row, a zero column) give 0, not NaN, next to regular ones."""
    rng = np.random.default_rng(n)
    n_mo = n + 5
    S = rng.standard_normal((n_mo, n_mo))
    S[:, 3] = 0.0                                   # any column list holding 3 is singular
    S[7] = S[2]                                     # any row list holding 2 and 7 is singular
    base = np.array([i for i in range(n_mo) if i not in (2, 3, 7)], dtype=np.int32)
    rows = np.stack([base[:n], np.concatenate([[2, 7], base[: n - 2]])]).astype(np.int32)
    cols = np.stack([base[:n], np.concatenate([base[: n - 1], [3]])]).astype(np.int32)
    ctx.set_lu_variant(variant)
    try:
        got = ctx.det_table_host(S, rows, cols)
    finally:
        ctx.set_lu_variant("auto")
    sgn, logdet = np.linalg.slogdet(S[np.ix_(rows[0], cols[0])])
    assert np.sign(got[0, 0]) == sgn and abs(np.log(abs(got[0, 0])) - logdet) < 1e-10
    scale = abs(got[0, 0])
    assert abs(got[0, 1]) == 0.0 and np.isfinite(got).all()
    assert abs(got[1, 0]) <= 1e-9 * scale and abs(got[1, 1]) == 0.0


def test_det_table_singular_and_tiny(ctx):
    S = np.arange(36, dtype=float).reshape(6, 6)            # rank 2: every 3x3 minor is singular
    rows = np.array([[0, 1, 2], [1, 3, 5]], dtype=np.int32)
    cols = np.array([[0, 2, 4]], dtype=np.int32)
    got = ctx.det_table_host(S, rows, cols)
    assert np.all(np.abs(got) < 1e-9)
    Z = np.zeros((4, 4))
    assert np.all(ctx.det_table_host(Z, rows[:, :2] % 4, cols[:, :2] % 4) == 0.0)


def test_precompute_matches_reference_semantics(ctx):
    """pywfo/dets.py:191-198 on explicit alpha/beta block lists."""
    from pywfo_b200 import dets
    g = load_golden("r_n9")
    S = _smo(g)
    occ = int(g["occ"])
    blocks_b = [orc.exc_index_list(e, occ) for e in [orc.GS, (0, 0), (3, 2), (8, 10)]]
    blocks_k = [orc.exc_index_list(e, occ) for e in [orc.GS, (1, 1), (8, 0)]]
    want = orc.det_from_lists(S, blocks_b, blocks_k)
    np.testing.assert_allclose(dets.precompute(blocks_b, blocks_k, S), want, rtol=1e-10, atol=1e-14)


# ---------------------------------------------------------------- whole path, goldens
@pytest.mark.parametrize("tier", ["auto", "lu", "update", "closed"])
def test_golden_pins(ctx, tier):
    from pywfo_b200 import main
    g = load_golden("simple")
    for fn in (main.overlaps_naive, main.overlaps_most_naive, main.overlaps_cache):
        got = fn(g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], occ=1, tier=tier)
        assert got[0][0] == pytest.approx(0.97008177)            # pywfo/tests/test_pywfo.py:94
        assert_parity(got, g["ref_overlaps_cache"])
    g = load_golden("more_virtual_mos")
    got = main.overlaps_most_naive(g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], occ=2, tier=tier)
    want = np.array((0.11025145, 0.97114924, 0.97595963, -0.09630209)).reshape(-1, 2)
    np.testing.assert_allclose(got, want, atol=1e-7)             # pywfo/tests/test_pywfo.py:122-125
    assert_parity(got, g["ref_overlaps_most_naive"])
    assert_parity(main.overlaps(g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], occ=2, tier=tier),
                  g["ref_overlaps"])


@pytest.mark.parametrize("tier", ["auto", "lu", "update", "closed"])
def test_h2o2(ctx, tier):
    """pywfo/tests/test_pywfo.py:128-156 and the 4x4 variants, reference self-output."""
    from pywfo_b200 import dets, main
    g = load_golden("h2o2")
    bra, ket, bci, kci = g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"]
    got = main.overlaps_cache(bra, ket, bci[:1], kci[:1], 9, ci_thresh=3e-2, ao_ovlps="ket", tier=tier)
    assert_parity(got, g["ref_cache_state0_3e-2"])
    assert_parity(main.overlaps_cache(bra, ket, bci, kci, 9, ci_thresh=1e-2, tier=tier), g["ref_cache_4x4_1e-2"])
    assert_parity(main.overlaps_cache(bra, ket, bci, kci, 9, ci_thresh=1e-4, tier=tier), g["ref_cache_4x4_1e-4"])
    assert_parity(main.overlaps(bra, ket, bci, kci, 9, ci_thresh=1e-2, tier=tier), g["ref_minus_4x4_1e-2"])
    S_AO = orc.get_S_AO(bra, ket, "ket")
    assert_parity(dets.wfoverlap(bra, ket, bci, kci, 1e-2, S_AO, debug=False, tier=tier), g["ref_wfoverlap_4x4_1e-2"])
    assert_parity(dets.wfoverlap(bra, ket, bci[:1], kci[:1], 3e-2, S_AO, tier=tier), g["ref_wfoverlap_1x1_3e-2"])


@pytest.mark.parametrize("name", RANDOM_CASES)
@pytest.mark.parametrize("tier", ["lu", "update", "closed"])
def test_random_goldens(ctx, name, tier):
    from pywfo_b200 import dets, main
    g = load_golden(name)
    a = (g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], int(g["occ"]))
    thr = float(g["ci_thresh"])
    assert_parity(main.overlaps_cache(*a, ci_thresh=thr, tier=tier), g["ref_overlaps_cache"])
    assert_parity(main.overlaps_naive(*a, ci_thresh=thr, tier=tier), g["ref_overlaps_cache"])
    assert_parity(main.overlaps_cache(*a, ci_thresh=thr, ao_ovlps="bra", tier=tier), g["ref_overlaps_cache_bra"])
    nmin = min(a[2].shape[0], a[3].shape[0])
    assert_parity(main.overlaps(a[0], a[1], a[2][:nmin], a[3][:nmin], a[4], ci_thresh=thr, tier=tier),
                  g["ref_overlaps"])
    S_AO = orc.get_S_AO(a[0], a[1], "bra")
    assert_parity(dets.wfoverlap(a[0], a[1], a[2], a[3], thr, S_AO, tier=tier), g["ref_wfoverlap"])
    # ndarray AO overlaps (documented upstream, main.py:305, but raising there): accepted here
    assert_parity(main.overlaps_cache(*a, ci_thresh=thr, ao_ovlps=S_AO, tier=tier), g["ref_overlaps_cache_bra"])


# ---------------------------------------------------------------- reference error behaviour
def test_error_behaviour(ctx):
    from pywfo_b200 import main
    g = load_golden("r_even")
    a = (g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], int(g["occ"]))
    with pytest.raises(AssertionError):                      # main.py:91
        main.overlaps(*a)
    with pytest.raises(Exception, match="Invalid AO overlaps"):   # main.py:308
        main.overlaps_cache(*a, ao_ovlps=1.5)
    bci = g["bra_ci"][:2].copy()
    bci[1] = 0
    with pytest.raises(ValueError):                          # main.py:186
        main.overlaps(g["bra_mos"], g["ket_mos"], bci, g["ket_ci"][:2], a[4])
    # thresholds that drop everything: zeros (the reference sums empty lists)
    z = main.overlaps_cache(*a, ci_thresh=10.0)
    assert z.shape == (3, 2) and not z.any()


# ---------------------------------------------------------------- kernel-form calls, edge cases
def _random_tables(rng, occ, virt, kb, kk, n_bra, n_ket, with_gs=False):
    bra = np.stack([rng.integers(0, occ, kb), rng.integers(0, virt, kb)], 1).astype(np.int32)
    ket = np.stack([rng.integers(0, occ, kk), rng.integers(0, virt, kk)], 1).astype(np.int32)
    if with_gs:
        bra[kb // 2] = (-1, -1)
        ket[0] = (-1, -1)
    return bra, rng.standard_normal((n_bra, kb)), ket, rng.standard_normal((n_ket, kk))


@pytest.mark.parametrize("occ,n_mo,kb,kk", [(1, 4, 3, 3), (2, 3, 2, 2), (5, 24, 95, 95), (9, 38, 130, 70),
                                           (34, 60, 40, 33), (64, 80, 20, 9), (100, 120, 12, 10), (128, 140, 6, 5)])
@pytest.mark.parametrize("sigma", [1.0, -1.0])
def test_kernel_form_tiers_agree_with_oracle(ctx, occ, n_mo, kb, kk, sigma):
    rng = np.random.default_rng(occ * 7 + kb)
    np.random.seed(occ)
    _, bra_mos, ket_mos = orc.bra_ket_dummy_mos(n_mo, occ)
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    bra, Cb, ket, Ck = _random_tables(rng, occ, n_mo - occ, kb, kk, 3, 4, with_gs=True)
    want = orc.state_overlaps_bruteforce(S, occ, [tuple(e) for e in bra], Cb, [tuple(e) for e in ket], Ck, sigma)
    for tier, code in (("lu", 0), ("update", 1)):
        got = ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, sigma=sigma, tier=tier)
        assert ctx.last_tier() == code
        assert_parity(got, want, rtol=1e-9)


@pytest.mark.parametrize("tier", ["lu", "update", "closed"])
def test_many_states(ctx, tier):
    """More ket states than one pass of the fused kernels holds (62): passes over state blocks."""
    from pywfo_b200 import helpers, main
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(31, 20, 6, 70, 131)
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    want = orc.state_overlaps_factored(S, 6, orc.signed_ci(bci, 6), orc.signed_ci(kci, 6), 1.0)
    got = main.overlaps_cache(bra_mos, ket_mos, bci, kci, 6, ci_thresh=0.0, tier=tier)
    assert got.shape == (70, 131)
    assert_parity(got, want, rtol=1e-9)


def test_closed_form_tier(ctx):
    """T2: GEMM-only closed form == brute force; refuses ground-state entries; shards add up."""
    from pywfo_b200._lib import WfoError
    rng = np.random.default_rng(21)
    occ, n_mo = 17, 50
    np.random.seed(21)
    _, bra_mos, ket_mos = orc.bra_ket_dummy_mos(n_mo, occ)
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    bra, Cb, ket, Ck = _random_tables(rng, occ, n_mo - occ, 60, 45, 3, 5)
    bra[7] = bra[3]                                  # a duplicate excitation (dets.get_dets produces them)
    for sigma in (1.0, -1.0):
        want = orc.state_overlaps_bruteforce(S, occ, [tuple(e) for e in bra], Cb, [tuple(e) for e in ket], Ck, sigma)
        got = ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, sigma=sigma, tier="closed")
        assert ctx.last_tier() == 2
        assert_parity(got, want, rtol=1e-9)
        parts = sum(ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, sigma=sigma, tier="closed", k_begin=lo, k_end=hi)
                    for lo, hi in ((0, 13), (13, 60)))
        np.testing.assert_allclose(parts, got, rtol=1e-12, atol=1e-14 * np.abs(got).max())
    bra[0] = (-1, -1)
    with pytest.raises(WfoError, match="ground-state"):
        ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier="closed")


def test_shards_add_up_on_gpu(ctx):
    rng = np.random.default_rng(3)
    occ, n_mo = 9, 38
    np.random.seed(3)
    _, bra_mos, ket_mos = orc.bra_ket_dummy_mos(n_mo, occ)
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    bra, Cb, ket, Ck = _random_tables(rng, occ, n_mo - occ, 150, 77, 4, 3)
    for tier in ("lu", "update"):
        full = ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier=tier)
        parts = sum(ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier=tier, k_begin=lo, k_end=hi)
                    for lo, hi in ((0, 19), (19, 19), (19, 100), (100, 150)))
        np.testing.assert_allclose(parts, full, rtol=1e-12, atol=1e-14 * np.abs(full).max())


def test_raw_ci_entry_point_shards_add_up(ctx):
    """wfo_overlaps_ci_host: excitation selection on the device, rank/world sharding."""
    g = load_golden("r_n17")
    occ = int(g["occ"])
    inv = np.linalg.inv(g["ket_mos"])
    for sem, key in ((0, "ref_overlaps_cache"), (1, "ref_overlaps")):
        nmin = min(g["bra_ci"].shape[0], g["ket_ci"].shape[0]) if sem else None
        bci, kci = g["bra_ci"][:nmin], g["ket_ci"][:nmin]
        full, info = ctx.overlaps_ci_host(g["bra_mos"], g["ket_mos"], inv, occ, bci, kci, float(g["ci_thresh"]), sem)
        assert_parity(full, g[key])
        from pywfo_b200 import excitations as ex
        assert info[0] == ex.build_tables(bci, float(g["ci_thresh"]), "minus" if sem else "cache").K
        parts = sum(ctx.overlaps_ci_host(g["bra_mos"], g["ket_mos"], inv, occ, bci, kci, float(g["ci_thresh"]), sem,
                                         rank=r, world=3)[0] for r in range(3))
        np.testing.assert_allclose(parts, full, rtol=1e-12, atol=1e-14 * np.abs(full).max())
    # sparse CI vectors, more ranks than some slices have survivors: each rank uploads and thresholds only
    # its slice of the bra (from, to) space (singlet semantics)
    g = load_golden("r_sparse")
    occ, thr = int(g["occ"]), float(g["ci_thresh"])
    inv = np.linalg.inv(g["ket_mos"])
    full, _ = ctx.overlaps_ci_host(g["bra_mos"], g["ket_mos"], inv, occ, g["bra_ci"], g["ket_ci"], thr, 0)
    assert_parity(full, g["ref_overlaps_cache"])
    for world in (2, 7, 40):
        parts = sum(ctx.overlaps_ci_host(g["bra_mos"], g["ket_mos"], inv, occ, g["bra_ci"], g["ket_ci"], thr, 0,
                                         rank=r, world=world)[0] for r in range(world))
        np.testing.assert_allclose(parts, full, rtol=1e-12, atol=1e-14 * np.abs(full).max())


def test_singular_base_block_falls_back_to_lu(ctx):
    """S_oo singular: the rank-update tier is not applicable, AUTO must use the brute force LU."""
    rng = np.random.default_rng(8)
    occ, n_mo = 6, 15
    q, _ = np.linalg.qr(rng.standard_normal((n_mo, n_mo)))
    S = q.copy()
    S[2, :occ] = 0.0                                   # a zero row inside the occupied block
    bra, Cb, ket, Ck = _random_tables(rng, occ, n_mo - occ, 30, 25, 2, 3)
    want = orc.state_overlaps_bruteforce(S, occ, [tuple(e) for e in bra], Cb, [tuple(e) for e in ket], Ck, 1.0)
    got = ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier="auto")
    assert ctx.last_tier() == 0
    assert_parity(got, want, rtol=1e-9)
    from pywfo_b200._lib import WfoError
    with pytest.raises(WfoError, match="singular"):
        ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier="update")


def test_device_pointer_entry_points(ctx):
    """The *_dev functions on torch-owned device memory (torch = memory + stream only)."""
    import torch
    g = load_golden("r_n17")
    occ = int(g["occ"])
    from pywfo_b200 import excitations as ex
    bt = ex.build_tables(g["bra_ci"], 0.05)
    kt = ex.build_tables(g["ket_ci"], 0.05, order="locality")
    dev = torch.device("cuda:0")
    f64 = dict(dtype=torch.float64, device=dev)
    # the golden MOs are transposed views (F-ordered): the C ABI wants dense row-major
    with pytest.raises(ValueError):
        t = torch.zeros((4, 6), **f64).t()
        ctx.sao_dev(t, torch.zeros((6, 6), **f64))
    bra = torch.tensor(np.ascontiguousarray(g["bra_mos"]), **f64).contiguous()
    ket = torch.tensor(np.ascontiguousarray(g["ket_mos"]), **f64).contiguous()
    inv = torch.tensor(np.linalg.inv(g["ket_mos"]), **f64).contiguous()
    n_mo = bra.shape[0]
    s_ao, s_mo = torch.empty((n_mo, n_mo), **f64), torch.empty((n_mo, n_mo), **f64)
    st = torch.cuda.current_stream().cuda_stream
    ctx.sao_dev(inv, s_ao, st)
    ctx.smo_dev(bra, s_ao, ket, s_mo, st)
    be = torch.tensor(bt.exc, dtype=torch.int32, device=dev)
    ke = torch.tensor(kt.exc, dtype=torch.int32, device=dev)
    cb, ck = torch.tensor(bt.coeff, **f64), torch.tensor(kt.coeff, **f64)
    out = torch.empty((cb.shape[0], ck.shape[0]), **f64)
    for tier in ("update", "lu"):
        ctx.state_overlaps_dev(s_mo, occ, be, cb, ke, ck, out, tier=tier, stream=st)
        torch.cuda.synchronize()
        want, _ = orc.ref_overlaps_cache(g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], occ, 0.05)
        assert_parity(out.cpu().numpy(), want)
    lists = torch.empty((bt.K, occ), dtype=torch.int32, device=dev)
    ctx.exc_lists_dev(be, occ, lists, st)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(lists.cpu().numpy(), ex.exc_index_lists(bt.exc, occ))
    assert ctx.launch_count() > 0


# ---------------------------------------------------------------- BASELINE.json configs
def test_config_water_full_cis(ctx):
    """configs[1]: n_occ=5, n_mo=24, 10x10 states, full CIS space (K=95); full reference restatement."""
    from pywfo_b200 import helpers, main
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(0, 24, 5, 10, 10)
    want, n_unique = orc.ref_overlaps_cache(bra_mos, ket_mos, bci, kci, 5, 0.0)
    assert n_unique == 96 * 96
    for tier in ("lu", "update"):
        assert_parity(main.overlaps_cache(bra_mos, ket_mos, bci, kci, 5, ci_thresh=0.0, tier=tier), want)


def test_config_naphthalene_triplet_truncated(ctx):
    """configs[2]: n_occ=34, n_mo=180, sigma=-1 (main.overlaps); truncated to what the CPU finishes."""
    from pywfo_b200 import helpers, main
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(1, 180, 34, 3, 3, top_k=40)
    want = orc.ref_overlaps_minus(bra_mos, ket_mos, bci, kci, 34, 1e-12)
    for tier in ("lu", "update"):
        assert_parity(main.overlaps(bra_mos, ket_mos, bci, kci, 34, ci_thresh=1e-12, tier=tier), want)


def test_config_naphthalene_triplet_full(ctx):
    """configs[2] at full size: n_occ=34, n_mo=180, 20x20 singlet->triplet states (main.overlaps, sigma=-1), the
    whole CIS space (K=4964 per side, 2.5e7 pairs).  The reference's pair loop cannot finish this; the checker
    is the factored closed form of the same sum, the brute-force tier is checked against the rank-update tier."""
    from pywfo_b200 import helpers, main
    occ, n_mo, ns = 34, 180, 20
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(1, n_mo, occ, ns, ns)
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    # main.overlaps folds sign^occ into the coefficients (pywfo/main.py:154,176): occ is even here
    want = orc.state_overlaps_factored(S, occ, orc.signed_ci(bci, occ), orc.signed_ci(kci, occ), -1.0)
    got = main.overlaps(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.0, tier="update")
    assert_parity(got, want, rtol=1e-9)
    got_lu = main.overlaps(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.0, tier="lu")
    assert_parity(got_lu, want, rtol=1e-9)
    assert_parity(got_lu, got, rtol=1e-9)


def test_config_large_full_space_properties(ctx):
    """configs[3]: n_occ=100, n_mo=500, 50x50 states, full CIS space (K=40 000 per side, 1.6e9 pairs).
    The reference cannot run this; checked against the factored closed form (O(n^2 v) per state) and
    through linearity in the bra coefficients."""
    from pywfo_b200 import helpers, main
    occ, n_mo, ns = 100, 500, 50
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(2, n_mo, occ, ns, ns)
    got = main.overlaps_cache(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.0, tier="update")
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    want = orc.state_overlaps_factored(S, occ, orc.signed_ci(bci, occ), orc.signed_ci(kci, occ), 1.0)
    assert_parity(got, want, rtol=1e-9)
    # linearity: state 0 replaced by a combination of states 1 and 2
    b2 = bci.copy()
    b2[0] = 0.25 * bci[1] - 1.5 * bci[2]
    got2 = main.overlaps_cache(bra_mos, ket_mos, b2, kci, occ, ci_thresh=0.0, tier="update")
    np.testing.assert_allclose(got2[0], 0.25 * got[1] - 1.5 * got[2], rtol=1e-9, atol=1e-12 * np.abs(got).max())
    np.testing.assert_allclose(got2[1:], got[1:], rtol=1e-13, atol=0)
    # the closed-form tier on the full problem
    got_c = main.overlaps_cache(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.0, tier="closed")
    assert_parity(got_c, want, rtol=1e-9)
    assert_parity(got_c, got, rtol=1e-9)


def test_config_large_lu_sample(ctx):
    """configs[3] shape, brute-force tier on a sample of the pair space (the full 1.6e9 LUs take
    minutes): LU tier == update tier == oracle on the same sample."""
    from pywfo_b200 import helpers
    occ, n_mo = 100, 500
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(2, n_mo, occ, 4, 5, top_k=24)
    from pywfo_b200 import main
    got_lu = main.overlaps_cache(bra_mos, ket_mos, bci, kci, occ, ci_thresh=1e-12, tier="lu")
    got_up = main.overlaps_cache(bra_mos, ket_mos, bci, kci, occ, ci_thresh=1e-12, tier="update")
    want, _ = orc.ref_overlaps_cache(bra_mos, ket_mos, bci, kci, occ, 1e-12)
    assert_parity(got_lu, want, rtol=1e-9)
    assert_parity(got_up, want, rtol=1e-9)


@pytest.mark.parametrize("variant", ["left", "right"])
def test_config_large_lu_sample_both_kernels(ctx, variant):
    """configs[3] shape (n_occ = 100): the brute-force tier through either LU kernel == oracle."""
    from pywfo_b200 import helpers, main
    occ, n_mo = 100, 500
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(2, n_mo, occ, 3, 3, top_k=12)
    want, _ = orc.ref_overlaps_cache(bra_mos, ket_mos, bci, kci, occ, 1e-12)
    ctx.set_lu_variant(variant)
    try:
        got = main.overlaps_cache(bra_mos, ket_mos, bci, kci, occ, ci_thresh=1e-12, tier="lu")
    finally:
        ctx.set_lu_variant("auto")
    assert_parity(got, want, rtol=1e-9)


@pytest.mark.parametrize("k", [32, 100, 316, 1000])
def test_config_sweep_n64(ctx, k):
    """configs[4]: n_occ=64, n_mo=320, 8x8 states, K_b=K_k=k random excitations shared by all states."""
    from pywfo_b200 import helpers
    occ, n_mo = 64, 320
    np.random.seed(3)
    _, bra_mos, ket_mos = helpers.get_bra_ket_dummy_mos(n_mo, occ)
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    rng = np.random.default_rng(3 + k)
    bra, Cb, ket, Ck = _random_tables(rng, occ, n_mo - occ, k, k, 8, 8)
    want = orc.state_overlaps_closed_form(S, occ, [tuple(e) for e in bra], Cb, [tuple(e) for e in ket], Ck, 1.0)
    for tier in ("lu", "update"):
        assert_parity(ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier=tier), want, rtol=1e-9)


def test_error_messages_are_formatted(ctx):
    """Every error path formats its message (regression: a mis-ordered printf argument crashed)."""
    from pywfo_b200._lib import WfoError
    S = np.eye(140)
    rows = np.arange(130, dtype=np.int32)[None, :]
    with pytest.raises(WfoError, match="n_occ=130"):
        ctx.det_table_host(S, rows, rows)
    bra = np.zeros((1, 2), dtype=np.int32)
    with pytest.raises(WfoError, match=r"bad shard \[0,5\)"):
        ctx.state_overlaps_host(np.eye(4), 2, bra, np.ones((1, 1)), bra, np.ones((1, 1)), k_begin=0, k_end=5)
    with pytest.raises(WfoError, match="n_occ=9 must be in 1..n_mo=4"):
        ctx.state_overlaps_host(np.eye(4), 9, bra, np.ones((1, 1)), bra, np.ones((1, 1)))


# ---------------------------------------------------------------- device inverse (np.linalg.inv, main.py:303)
@pytest.mark.parametrize("n,cond", [(3, 1.0), (38, 30.0), (137, 1e3), (500, 1.0), (500, 1e5)])
def test_device_inverse_matches_lapack(ctx, n, cond):
    rng = np.random.default_rng(n)
    u, _ = np.linalg.qr(rng.standard_normal((n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sv = np.logspace(0, -np.log10(cond), n)
    a = (u * sv) @ v.T
    got = ctx.inverse_host(a)
    want = np.linalg.inv(a)
    # forward error of any backward-stable inverse is O(eps * cond)
    assert np.abs(got - want).max() <= 1e-12 * cond * np.abs(want).max()
    assert np.abs(a @ got - np.eye(n)).max() <= 1e-11 * max(cond ** 0.5, 1.0)
    assert ctx.last_inverse_info()[1] is (n <= 128)   # small: one-launch pivoted Gauss-Jordan; else Newton-Schulz


@pytest.mark.parametrize("n,cond", [(24, 1e8), (100, 1e12), (137, 1e10), (500, 1e10), (200, 1e13), (500, 1e13), (1000, 1e9)])
def test_device_inverse_pivoted_fallback_is_lapack_grade(ctx, n, cond):
    """pywfo/main.py:303 is np.linalg.inv = LAPACK with partial pivoting, which returns a result up to
    cond ~ 1/eps.  Newton-Schulz stagnates above cond ~ 1e7 in FP64; the pivoted Gauss-Jordan fallback
    must then deliver what LAPACK delivers.  Bound: an ABSOLUTE residual.  For an inverse computed by
    Gaussian elimination with partial pivoting the LEFT residual |X A - I| is O(n eps cond) at worst and
    in practice far smaller; LAPACK's own residual on the same matrix is measured here and the device
    result may exceed it by at most 20x (and must stay below 1e-6 up to cond 1e10).  The element-wise
    difference to LAPACK's inverse is bounded by the forward-error bound eps * cond * |X|."""
    rng = np.random.default_rng(n + int(np.log10(cond)))
    u, _ = np.linalg.qr(rng.standard_normal((n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sv = np.logspace(0, -np.log10(cond), n)
    a = (u * sv) @ v.T
    got = ctx.inverse_host(a)
    its, pivoted = ctx.last_inverse_info()
    assert pivoted, f"expected the pivoted fallback at cond={cond:g} (Newton-Schulz ran {its} iterations)"
    want = np.linalg.inv(a)
    eye = np.eye(n)
    res_dev = min(np.abs(got @ a - eye).max(), np.abs(a @ got - eye).max())
    res_lap = min(np.abs(want @ a - eye).max(), np.abs(a @ want - eye).max())
    assert res_dev <= max(20.0 * res_lap, 1e-9), (res_dev, res_lap)
    if cond <= 1e10:
        assert res_dev <= 1e-6, res_dev
    assert np.abs(got - want).max() <= 50 * np.finfo(float).eps * cond * np.abs(want).max()


def test_ill_conditioned_mos_through_the_whole_path(ctx):
    """Ill-conditioned ket MOs (cond 1e9: Newton-Schulz gives up, the pivoted elimination takes over) through
    overlaps_cache.  No reference golden can pin this regime: upstream forms S_MO = (bra S_AO) ket^T with
    S_AO = inv inv^T, whose entries are ~cond^2 and cancel to O(1), so its own answer carries an error of
    ~eps cond^2 (measured with the unmodified reference: max|S| = 6e8 instead of O(1) at cond 1e10).  What
    can be pinned: the device-inverse path equals the same path fed with LAPACK's inverse, and both equal
    the state overlaps computed from the exact S_MO = rot (bra = rot ket by construction)."""
    from pywfo_b200 import helpers, main
    from pywfo_b200 import _lib
    rng = np.random.default_rng(7)
    n_mo, occ, ns, cond = 40, 6, 3, 1e9
    u, _ = np.linalg.qr(rng.standard_normal((n_mo, n_mo)))
    v, _ = np.linalg.qr(rng.standard_normal((n_mo, n_mo)))
    ket_mos = (u * np.logspace(0, -np.log10(cond), n_mo)) @ v.T
    k = 0.02 * rng.standard_normal((n_mo, n_mo))
    rot, _ = np.linalg.qr(np.eye(n_mo) + (k - k.T))
    bra_mos = rot @ ket_mos
    _, _, bci, kci = helpers.synthetic_case(3, n_mo, occ, ns, ns)
    got = main.overlaps_cache(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.0, ao_ovlps="ket")
    assert _lib.context().last_inverse_info()[1]
    got_lapack = main.overlaps_cache(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.0, ao_ovlps="ket",
                                     c_inv=np.linalg.inv(ket_mos))
    want = orc.state_overlaps_factored(rot, occ, orc.signed_ci(bci, occ), orc.signed_ci(kci, occ), 1.0)
    tol = 1e3 * np.finfo(float).eps * cond   # eps * cond: the error of ket Cinv - I
    assert np.abs(got - want).max() <= tol * np.abs(want).max()
    assert np.abs(got_lapack - want).max() <= tol * np.abs(want).max()


def test_device_inverse_reports_singular(ctx):
    from pywfo_b200 import _lib
    a = np.ones((8, 8))
    with pytest.raises(_lib.WfoError) as ei:
        ctx.inverse_host(a)
    assert ei.value.code == _lib.ERR_NOCONV


@pytest.mark.parametrize("name", ["h2o2", "r_n17", "r_odd"])
@pytest.mark.parametrize("ao", ["ket", "bra"])
def test_device_inverse_path_equals_host_inverse_path(ctx, name, ao):
    """ao_ovlps in {"bra","ket"}: inverse formed on the device (default) vs an inverse supplied by the
    caller (here numpy's, i.e. the number upstream's LAPACK call produces)."""
    from pywfo_b200 import main
    g = load_golden(name)
    occ = int(g["occ"])
    args = (g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], occ)
    thr = float(g["ci_thresh"]) if "ci_thresh" in g else 1e-2
    dev = main.overlaps_cache(*args, ci_thresh=thr, ao_ovlps=ao)
    host = main.overlaps_cache(*args, ci_thresh=thr, ao_ovlps=ao,
                               c_inv=np.linalg.inv(g["bra_mos"] if ao == "bra" else g["ket_mos"]))
    assert_parity(dev, host)
    want = {("h2o2", "ket"): "ref_cache_4x4_1e-2", ("r_n17", "ket"): "ref_overlaps_cache", ("r_odd", "ket"): "ref_overlaps_cache",
            ("r_n17", "bra"): "ref_overlaps_cache_bra", ("r_odd", "bra"): "ref_overlaps_cache_bra"}.get((name, ao))
    if want:
        assert_parity(dev, g[want])


def test_singular_mo_matrix_raises_like_numpy(ctx):
    from pywfo_b200 import main
    mos = np.ones((6, 6))
    ci = np.full((1, 2, 4), 0.3)
    with pytest.raises(np.linalg.LinAlgError):
        main.overlaps_cache(mos, mos, ci, ci, 2)


# ---------------------------------------------------------------- trajectory driver
@pytest.mark.parametrize("tier", ["auto", "lu"])
def test_trajectory_matches_pairwise_reference(ctx, tier):
    """One library call for the whole cytosine trajectory == the reference called pair by pair."""
    from pywfo_b200 import main, trajectory
    g = load_golden("cytosin_traj")
    occ, thr = int(g["occ"]), float(g["ci_thresh"])
    mo, ci, pairs = g["mo_coeffs"], g["ci_coeffs"], g["pairs"]
    got = trajectory.trajectory_overlaps(mo, ci, occ, ci_thresh=thr, pairs=pairs, tier=tier)
    assert_parity(got, g["ref_overlaps_cache"])
    if tier == "auto":
        got_m = trajectory.trajectory_overlaps(mo, ci, occ, ci_thresh=thr, pairs=pairs, semantics="minus")
        assert_parity(got_m, g["ref_overlaps"])
        got_b = trajectory.trajectory_overlaps(mo, ci, occ, ci_thresh=thr, pairs=pairs[:2], ao_ovlps="bra")
        assert_parity(got_b, g["ref_overlaps_cache_bra"])
        # default pairs = consecutive cycles; same numbers as single calls through the public entry point
        cons = trajectory.trajectory_overlaps(mo, ci, occ, ci_thresh=thr)
        assert cons.shape == (3, 2, 2)
        for p in range(3):
            one = main.overlaps_cache(mo[p], mo[p + 1], ci[p], ci[p + 1], occ, ci_thresh=thr)
            np.testing.assert_allclose(cons[p], one, rtol=1e-12, atol=1e-14)
            assert_parity(cons[p], g["ref_overlaps_cache"][p])


def test_trajectory_excitation_shards_add_up(ctx):
    g = load_golden("cytosin_traj")
    occ, thr = int(g["occ"]), float(g["ci_thresh"])
    full, info = ctx.trajectory_ci_host(g["mo_coeffs"], g["ci_coeffs"], occ, g["pairs"][:2], thr)
    parts = sum(ctx.trajectory_ci_host(g["mo_coeffs"], g["ci_coeffs"], occ, g["pairs"][:2], thr, rank=r, world=3)[0]
                for r in range(3))
    np.testing.assert_allclose(parts, full, rtol=1e-12, atol=1e-14 * np.abs(full).max())
    assert info[0] > 0 and info[1] > 0


def test_smo_from_inverse_two_launch_chain(ctx):
    """S_MO = (bra Cinv)(ket Cinv)^T (wfo_smo_inv_dev) == bra (Cinv Cinv^T) ket^T (main.py:304, 532)."""
    import torch
    for name in ("h2o2", "r_n34", "r_odd"):
        g = load_golden(name)
        for ao in ("ket", "bra"):
            inv = np.linalg.inv(g["ket_mos"] if ao == "ket" else g["bra_mos"])
            want = orc.mo_overlap(g["bra_mos"], g["ket_mos"], orc.get_S_AO(g["bra_mos"], g["ket_mos"], ao))
            f64 = dict(dtype=torch.float64, device="cuda")
            d_bra, d_ket, d_inv = (torch.tensor(np.ascontiguousarray(x), **f64) for x in (g["bra_mos"], g["ket_mos"], inv))
            n = d_bra.shape[0]
            out, work = torch.empty((n, n), **f64), torch.empty((2, n, n), **f64)
            ctx.smo_inv_dev(d_bra, d_ket, d_inv, out, work)
            torch.cuda.synchronize()
            np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-11, atol=1e-12)
            ctx.smo_inv_dev(d_bra, d_bra, d_inv, out, work)      # same matrix on both sides: single left factor
            torch.cuda.synchronize()
            want2 = orc.mo_overlap(g["bra_mos"], g["bra_mos"], inv @ inv.T)
            np.testing.assert_allclose(out.cpu().numpy(), want2, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("n_ket", [1, 2, 7, 8, 9, 15, 16, 17, 23, 24, 25, 32, 39, 40, 41, 47, 48, 49, 55, 56, 57, 111])
def test_every_rank_update_kernel_variant(ctx, n_ket):
    """All (row stride, DMMA tile count) variants of the rank-update kernel, and the multi-pass path for more
    than 55 ket states; ragged bra/ket tile edges (kb, kk not multiples of the 256 x 32 tiles); shard offsets."""
    rng = np.random.default_rng(1000 + n_ket)
    occ, n_mo = 7, 30
    q, _ = np.linalg.qr(rng.standard_normal((n_mo, n_mo)))
    S = q + 0.05 * rng.standard_normal((n_mo, n_mo))
    bra, Cb, ket, Ck = _random_tables(rng, occ, n_mo - occ, 301, 77, 5, n_ket, with_gs=True)
    want = orc.state_overlaps_bruteforce(S, occ, [tuple(e) for e in bra], Cb, [tuple(e) for e in ket], Ck, 1.0)
    got = ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier="update")
    assert ctx.last_tier() == 1
    assert_parity(got, want)
    parts = sum(ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier="update", k_begin=lo, k_end=hi)
                for lo, hi in ((0, 1), (1, 257), (257, 301)))
    np.testing.assert_allclose(parts, got, rtol=1e-11, atol=1e-13 * np.abs(got).max())


def test_results_are_run_to_run_identical(ctx):
    """no floating-point atomics, static stream-K split, fixed-order reductions -> bitwise
    reproducible on a fixed GPU count (both tiers, and the whole path from raw inputs)."""
    from pywfo_b200 import helpers, main
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(5, 60, 20, 7, 9)
    for tier in ("update", "lu"):
        a = main.overlaps_cache(bra_mos, ket_mos, bci, kci, 20, ci_thresh=0.02, tier=tier)
        for _ in range(3):
            b = main.overlaps_cache(bra_mos, ket_mos, bci, kci, 20, ci_thresh=0.02, tier=tier)
            assert np.array_equal(a, b)


@pytest.mark.parametrize("occ,n_mo", [(129, 150), (200, 260), (300, 340)])
def test_base_blocks_larger_than_128(ctx, occ, n_mo):
    """n_occ > 128: the base block is inverted blockwise (Schur complement on the DMMA GEMM around the
    register-resident 128-kernels); rank-update and closed-form tiers against the oracle."""
    rng = np.random.default_rng(occ)
    q, _ = np.linalg.qr(rng.standard_normal((n_mo, n_mo)))
    S = q + 0.02 * rng.standard_normal((n_mo, n_mo))
    bra, Cb, ket, Ck = _random_tables(rng, occ, n_mo - occ, 70, 45, 3, 4)
    want = orc.state_overlaps_closed_form(S, occ, [tuple(e) for e in bra], Cb, [tuple(e) for e in ket], Ck, 1.0)
    for tier in ("update", "closed", "auto"):
        got = ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier=tier)
        assert_parity(got, want, rtol=1e-9)
    # a few pair determinants directly against LAPACK, to pin the oracle's closed form at this size too
    from pywfo_b200 import _lib
    with pytest.raises(_lib.WfoError):
        ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier="lu")      # brute-force tier stops at 128
    k, l = tuple(bra[3]), tuple(ket[5])
    one = orc.state_overlaps_bruteforce(S, occ, [k], np.ones((1, 1)), [l], np.ones((1, 1)), 1.0)
    got1 = ctx.state_overlaps_host(S, occ, bra[3:4], np.ones((1, 1)), ket[5:6], np.ones((1, 1)), tier="update")
    assert_parity(got1, one, rtol=1e-9)


@pytest.mark.parametrize("occ,n_mo", [(150, 190), (260, 300)])
def test_base_blocks_larger_than_128_with_reordered_orbitals(ctx, occ, n_mo):
    """occupied orbitals reordered between the two geometries make the LEADING half of
    S_MO[:occ,:occ] singular although the block itself is perfectly conditioned; LAPACK (the reference) does
    not care.  The blockwise Schur inverse does (it pivots inside 128-blocks only), so the call falls back to
    the pivoted Gauss-Jordan inverse of the whole block and still takes the rank-update tier."""
    rng = np.random.default_rng(occ)
    S = np.eye(n_mo) + 0.02 * rng.standard_normal((n_mo, n_mo))   # MO overlaps of two close geometries
    perm = np.arange(n_mo)
    perm[:occ] = np.roll(perm[:occ], occ // 2 + 3)       # ket occupied orbitals rotated: A = (near identity) x permutation
    S = S[:, perm]
    h = occ // 2
    S[:h - 4, :h - 4] = 0.0                              # ... whose leading half is exactly rank deficient
    assert np.linalg.cond(S[:occ, :occ]) < 50 and np.linalg.matrix_rank(S[:h, :h]) < h
    bra, Cb, ket, Ck = _random_tables(rng, occ, n_mo - occ, 60, 40, 3, 4)
    want = orc.state_overlaps_closed_form(S, occ, [tuple(e) for e in bra], Cb, [tuple(e) for e in ket], Ck, 1.0)
    for tier in ("update", "auto"):
        got = ctx.state_overlaps_host(S, occ, bra, Cb, ket, Ck, tier=tier)
        assert ctx.last_tier() == 1
        assert_parity(got, want, rtol=1e-9)
    k, l = tuple(bra[2]), tuple(ket[7])                   # one pair directly against LAPACK's determinant
    one = orc.state_overlaps_bruteforce(S, occ, [k], np.ones((1, 1)), [l], np.ones((1, 1)), 1.0)
    got1 = ctx.state_overlaps_host(S, occ, bra[2:3], np.ones((1, 1)), ket[7:8], np.ones((1, 1)), tier="update")
    assert_parity(got1, one, rtol=1e-9)
    # a truly singular base block is still reported
    from pywfo_b200 import _lib
    S2 = S.copy()
    S2[5, :occ] = S2[9, :occ]
    with pytest.raises(_lib.WfoError):
        ctx.state_overlaps_host(S2, occ, bra, Cb, ket, Ck, tier="update")


def test_raw_ci_entry_point_from_device_buffers(ctx):
    """wfo_overlaps_ci_dev: orbitals and CI tensors already in HBM (torch tensors as plain memory)."""
    import torch
    for name, key, sem in (("r_n17", "ref_overlaps_cache", 0), ("r_odd", "ref_overlaps_cache", 0), ("r_n17", "ref_overlaps", 1)):
        g = load_golden(name)
        occ, thr = int(g["occ"]), float(g["ci_thresh"])
        nmin = min(g["bra_ci"].shape[0], g["ket_ci"].shape[0]) if sem else None
        f64 = dict(dtype=torch.float64, device="cuda")
        d = [torch.tensor(np.ascontiguousarray(x), **f64) for x in (g["bra_mos"], g["ket_mos"], g["bra_ci"][:nmin], g["ket_ci"][:nmin])]
        out = torch.full((d[2].shape[0], d[3].shape[0]), float("nan"), **f64)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            info = ctx.overlaps_ci_dev(d[0], d[1], occ, d[2], d[3], out, thr, semantics=sem, stream=s.cuda_stream)
        s.synchronize()
        assert info[0] > 0 and info[1] > 0
        assert_parity(out.cpu().numpy(), g[key])


@pytest.mark.parametrize("n_mo,occ", [(96, 20), (130, 37), (200, 100), (150, 128)])
def test_early_base_block_hint(ctx, n_mo, occ):
    """wfo_set_occ_hint: wfo_smo_inv_dev multiplies and inverts S_MO[:occ,:occ] (pywfo/main.py:548) on a side
    stream ahead of the full S_MO; the state-overlap call on that buffer picks it up.  Same result as the
    unhinted path (to rounding: the base block is multiplied with another tile shape), stale hints are ignored
    (other n_occ, other buffer), and a hint that is never consumed does no harm."""
    import torch
    from pywfo_b200 import excitations as ex, helpers
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(11 + occ, n_mo, occ, 3, 4, top_k=60)
    bt, kt = ex.build_tables(bci, 1e-12, "cache"), ex.build_tables(kci, 1e-12, "cache", order="locality")
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    want = orc.state_overlaps_closed_form(S, occ, [tuple(e) for e in bt.exc], bt.coeff, [tuple(e) for e in kt.exc],
                                          kt.coeff, 1.0)
    f64 = dict(dtype=torch.float64, device="cuda")
    d_bra, d_ket, d_inv = (torch.tensor(np.ascontiguousarray(x), **f64) for x in (bra_mos, ket_mos, np.linalg.inv(ket_mos)))
    d_smo, d_work = torch.empty((n_mo, n_mo), **f64), torch.empty((2, n_mo, n_mo), **f64)
    d_be, d_ke = (torch.tensor(t.exc, dtype=torch.int32, device="cuda") for t in (bt, kt))
    d_cb, d_ck = torch.tensor(bt.coeff, **f64), torch.tensor(kt.coeff, **f64)
    d_out = torch.empty((3, 4), **f64)

    def run(hint, occ_call=occ):
        ctx.set_occ_hint(hint)
        ctx.smo_inv_dev(d_bra, d_ket, d_inv, d_smo, d_work)
        ctx.state_overlaps_dev(d_smo, occ_call, d_be, d_cb, d_ke, d_ck, d_out, tier="update")
        torch.cuda.synchronize()
        return d_out.cpu().numpy().copy()

    try:
        plain = run(0)
        assert_parity(plain, want)
        hinted = run(occ)
        assert_parity(hinted, want)
        np.testing.assert_allclose(hinted, plain, rtol=1e-11, atol=1e-13 * np.abs(plain).max())
        assert np.array_equal(run(occ), hinted)                 # and reproducible
        stale = run(occ - 1)                                    # hint for another n_occ: ignored by the consumer
        assert np.array_equal(stale, plain)
        ctx.set_occ_hint(occ)                                   # a hint nobody consumes, then a fresh S_MO elsewhere
        ctx.smo_inv_dev(d_bra, d_ket, d_inv, d_smo, d_work)
        other = d_smo.clone()
        ctx.state_overlaps_dev(other, occ, d_be, d_cb, d_ke, d_ck, d_out, tier="update")
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), plain)
    finally:
        ctx.set_occ_hint(0)


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 16, 33, 63, 64, 65, 100, 121, 122, 127, 128])
@pytest.mark.parametrize("kind", ["near_identity", "random", "permuted"])
def test_base_inverse_every_panel_shape(ctx, n, kind):
    """The pipelined Gauss-Jordan inverse behind the rank-update tier (csrc/lu_base.cuh: panels of 8 columns per
    warp, ragged last panel, row pivoting across panels) against numpy.linalg.inv / det -- what the reference
    takes from LAPACK (pywfo/main.py:303, 548)."""
    rng = np.random.default_rng(300 + n)
    if kind == "near_identity":
        a = np.eye(n) + 0.05 * rng.standard_normal((n, n))
    elif kind == "random":
        a = rng.standard_normal((n, n))
    else:   # every pivot far from the diagonal
        a = (np.eye(n) + 0.01 * rng.standard_normal((n, n)))[rng.permutation(n)]
    got = ctx.inverse_host(a)
    resid = np.abs(a @ got - np.eye(n)).max()
    assert resid < 1e-9 * max(1.0, np.linalg.cond(a)), (n, kind, resid)
    np.testing.assert_allclose(got, np.linalg.inv(a), rtol=0, atol=1e-9 * np.abs(np.linalg.inv(a)).max() * max(1.0, np.linalg.cond(a) * 1e-3))


@pytest.mark.parametrize("fn_name,ref_name", [("overlaps_cache", "ref_overlaps_cache"), ("overlaps", "ref_overlaps_minus")])
def test_ci_tensors_with_fewer_virtuals(ctx, fn_name, ref_name):
    """the reference indexes S_MO with occ + to only (pywfo/main.py:565-576), so CI tensors that cover
    fewer virtual orbitals than the MO matrices hold are accepted upstream; here too (zero-widened), string and
    ndarray AO overlaps alike."""
    from pywfo_b200 import helpers, main
    n_mo, occ, vci = 40, 8, 20
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(77, n_mo, occ, 3, 3)
    bci, kci = np.ascontiguousarray(bci[:, :, :vci]), np.ascontiguousarray(kci[:, :, :vci])
    want = getattr(orc, ref_name)(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.02)
    want = want[0] if isinstance(want, tuple) else want
    got = getattr(main, fn_name)(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.02)
    assert_parity(got, want)
    inv = np.linalg.inv(ket_mos)
    got2 = getattr(main, fn_name)(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.02, ao_ovlps=inv @ inv.T)
    assert_parity(got2, want)
