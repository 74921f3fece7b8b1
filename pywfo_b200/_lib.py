"""ctypes binding of libwfo_b200.so (include/wfo_b200.h).

There is no CPU fallback: if the library is missing or no B200 is visible the
compute entry points raise.  The library is built in-tree by
``pywfo_b200/build.py`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WFO_B200_LIB") or os.path.join(HERE, "libwfo_b200.so")   # env: tuning builds only

TIER_AUTO, TIER_LU, TIER_UPDATE, TIER_CLOSED = -1, 0, 1, 2
TIERS = {"auto": TIER_AUTO, "lu": TIER_LU, "update": TIER_UPDATE, "closed": TIER_CLOSED,
         None: TIER_AUTO, -1: TIER_AUTO, 0: TIER_LU, 1: TIER_UPDATE, 2: TIER_CLOSED}

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/wfo_b200.h declares
SIGNATURES = {
    "wfo_version": (C.c_int, []),
    "wfo_last_error": (C.c_char_p, []),
    "wfo_device_count": (C.c_int, []),
    "wfo_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "wfo_destroy": (C.c_int, [c_vp]),
    "wfo_launch_count": (C.c_longlong, [c_vp]),
    "wfo_smo_dev": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "wfo_smo_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "wfo_sao_dev": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "wfo_gemm_dev": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_vp, C.c_int,
                               c_vp, C.c_int, C.c_double, c_vp, C.c_int, c_vp]),
    "wfo_det_table_dev": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp]),
    "wfo_det_table_host": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_vp]),
    "wfo_exc_lists_dev": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp]),
    "wfo_state_overlaps_dev": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_vp, C.c_int,
                                         c_vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "wfo_state_overlaps_host": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_vp, C.c_int,
                                          c_vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, c_vp]),
    "wfo_overlaps_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_vp,
                                    C.c_int, c_vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, c_vp]),
    "wfo_smo_inv_dev": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "wfo_inverse_dev": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "wfo_inverse_host": (C.c_int, [c_vp, c_vp, C.c_int, c_vp]),
    "wfo_last_inverse_info": (C.c_int, [c_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "wfo_overlaps_ci_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int,
                                       C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "wfo_overlaps_ci_dev": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int,
                                      C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "wfo_trajectory_ci_host": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                         C.c_int, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "wfo_overlap_from_lists_host": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int, C.c_int,
                                              c_vp, C.c_int, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp, C.c_int, C.c_int,
                                              c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "wfo_comm_init": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_char_p, C.c_longlong]),
    "wfo_comm_destroy": (C.c_int, [c_vp]),
    "wfo_comm_world": (C.c_int, [c_vp]),
    "wfo_allreduce_dev": (C.c_int, [c_vp, c_vp, C.c_longlong, c_vp]),
    "wfo_comm_status": (C.c_int, [c_vp]),
    "wfo_set_lu_variant": (C.c_int, [c_vp, C.c_int]),
    "wfo_last_tier": (C.c_int, [c_vp]),
    "wfo_set_occ_hint": (C.c_int, [c_vp, C.c_int]),
    "wfo_set_timing": (C.c_int, [c_vp, C.c_int]),
    "wfo_last_kernel_ms": (C.c_double, [c_vp]),
    "wfo_last_phases": (C.c_int, [c_vp, c_dp, C.c_char_p, C.c_int]),
    "wfo_fp64_peak": (C.c_int, [c_vp, C.c_int, C.c_int, c_dp]),
}


AO_GIVEN, AO_BRA, AO_KET = 0, 1, 2
ERR_NOCONV = 5
ERR_COMM = 6


class WfoError(RuntimeError):
    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


_lib = None
_lock = threading.Lock()
_contexts: dict[int, "Context"] = {}


def load():
    """dlopen the in-tree library and attach the prototypes.  Raises if absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise WfoError(
                    f"{LIB_PATH} is missing: build it with `python -m pywfo_b200.build` "
                    "(there is no CPU fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().wfo_last_error().decode("utf-8", "replace")
        raise WfoError(f"wfo_b200 error {rc}: {msg}", code=rc)


def _hptr(a: np.ndarray):
    return a.ctypes.data_as(c_vp)


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)



def _check_exc(name, exc, n_occ, n_virt):
    """(from, to) tables index S_MO on the device: out-of-range entries must fail here, as the reference's own
    indexing would (IndexError, pywfo/main.py:565-576), not read past the matrix.  (-1, -1) is the ground state."""
    if exc.size == 0 or n_occ < 1 or n_virt < 0:   # (bad sizes: the library's own message)
        return
    f, t = exc[:, 0], exc[:, 1]
    gs = (f < 0) & (t < 0)
    bad = ~gs & ((f < 0) | (f >= n_occ) | (t < 0) | (t >= n_virt))
    if bad.any():
        k = int(np.argmax(bad))
        raise IndexError(f"{name} excitation {k} = ({int(f[k])}, {int(t[k])}) is outside occ={n_occ}, virt={n_virt}")


class Context:
    """One per (process, device); owns the device scratch arena."""

    def __init__(self, device: int = 0):
        lib = load()
        if lib.wfo_device_count() <= 0:
            raise WfoError("no CUDA device visible: pywfo_b200 has no CPU fallback")
        h = c_vp()
        check(lib.wfo_create(device, C.byref(h)))
        self.lib, self.handle, self.device = lib, h, device

    def close(self):
        if self.handle:
            self.lib.wfo_destroy(self.handle)
            self.handle = None

    # ---- host-pointer entry points -------------------------------------------
    def smo_host(self, bra_mos, s_ao, ket_mos):
        bra, s, ket = as_f64(bra_mos), as_f64(s_ao), as_f64(ket_mos)
        p, u = bra.shape
        q, v = ket.shape
        if s.shape != (u, v):
            raise ValueError(f"S_AO has shape {s.shape}, expected {(u, v)}")
        out = np.empty((p, q))
        check(self.lib.wfo_smo_host(self.handle, _hptr(bra), _hptr(s), _hptr(ket), p, u, v, q, _hptr(out)))
        return out

    def det_table_host(self, s_mo, rows, cols):
        s = as_f64(s_mo)
        r, c = as_i32(rows), as_i32(cols)
        n = r.shape[1]
        if c.shape[1] != n or s.shape[0] != s.shape[1]:
            raise ValueError("rows/cols must be (K, n) lists and S_MO square")
        out = np.empty((r.shape[0], c.shape[0]))
        check(self.lib.wfo_det_table_host(self.handle, _hptr(s), s.shape[0], n, _hptr(r), r.shape[0],
                                          _hptr(c), c.shape[0], _hptr(out)))
        return out

    def state_overlaps_host(self, s_mo, n_occ, bra_exc, Cb, ket_exc, Ck, sigma=1.0, tier="auto",
                            k_begin=0, k_end=None):
        s = as_f64(s_mo)
        be, ke = as_i32(bra_exc).reshape(-1, 2), as_i32(ket_exc).reshape(-1, 2)
        cb, ck = as_f64(Cb), as_f64(Ck)
        kb, kk = be.shape[0], ke.shape[0]
        if cb.shape[1] != kb or ck.shape[1] != kk:
            raise ValueError("coefficient matrices do not match the excitation tables")
        k_end = kb if k_end is None else k_end
        _check_exc("bra", be, int(n_occ), s.shape[0] - int(n_occ))
        _check_exc("ket", ke, int(n_occ), s.shape[0] - int(n_occ))
        out = np.empty((cb.shape[0], ck.shape[0]))
        check(self.lib.wfo_state_overlaps_host(self.handle, _hptr(s), s.shape[0], int(n_occ), _hptr(be), kb,
                                               _hptr(cb), cb.shape[0], _hptr(ke), kk, _hptr(ck), ck.shape[0],
                                               float(sigma), TIERS[tier], int(k_begin), int(k_end), _hptr(out)))
        return out

    def overlaps_host(self, bra_mos, ket_mos, c_inv, n_occ, bra_exc, Cb, ket_exc, Ck, sigma=1.0, tier="auto",
                      k_begin=0, k_end=None, out=None):
        bra, ket, inv = as_f64(bra_mos), as_f64(ket_mos), as_f64(c_inv)
        n_mo = bra.shape[0]
        if bra.shape != (n_mo, n_mo) or ket.shape != (n_mo, n_mo) or inv.shape != (n_mo, n_mo):
            raise ValueError("MO matrices must be square and of equal shape")
        be, ke = as_i32(bra_exc).reshape(-1, 2), as_i32(ket_exc).reshape(-1, 2)
        cb, ck = as_f64(Cb), as_f64(Ck)
        kb, kk = be.shape[0], ke.shape[0]
        if cb.shape[1] != kb or ck.shape[1] != kk:
            raise ValueError("coefficient matrices do not match the excitation tables")
        k_end = kb if k_end is None else k_end
        _check_exc("bra", be, int(n_occ), n_mo - int(n_occ))
        _check_exc("ket", ke, int(n_occ), n_mo - int(n_occ))
        if out is None:
            out = np.empty((cb.shape[0], ck.shape[0]))
        check(self.lib.wfo_overlaps_host(self.handle, _hptr(bra), _hptr(ket), _hptr(inv), n_mo, int(n_occ),
                                         _hptr(be), kb, _hptr(cb), cb.shape[0], _hptr(ke), kk, _hptr(ck),
                                         ck.shape[0], float(sigma), TIERS[tier], int(k_begin), int(k_end),
                                         _hptr(out)))
        return out

    def inverse_host(self, a):
        """a^-1 on the device (Newton-Schulz on the DMMA GEMM); raises WfoError(code=ERR_NOCONV)
        when the iteration does not converge."""
        a = as_f64(a)
        n = a.shape[0]
        if a.shape != (n, n):
            raise ValueError("inverse needs a square matrix")
        out = np.empty((n, n))
        check(self.lib.wfo_inverse_host(self.handle, _hptr(a), n, _hptr(out)))
        return out

    def overlaps_ci_host(self, bra_mos, ket_mos, c_inv, n_occ, bra_ci, ket_ci, ci_thresh, semantics=0,
                         tier="auto", rank=0, world=1, ao_side=AO_GIVEN):
        """Whole path from raw CI tensors; excitation selection on the device.  ``c_inv`` is the
        inverse of the MO matrix behind S_AO (``ao_side=AO_GIVEN``) or None with ``ao_side`` =
        AO_BRA / AO_KET: the inverse is then computed on the device.
        Returns (partial matrix, info = [K_bra, K_ket, empty_state, tier_used])."""
        bra, ket = as_f64(bra_mos), as_f64(ket_mos)
        inv = as_f64(c_inv) if c_inv is not None else None
        if inv is None and ao_side == AO_GIVEN:
            raise ValueError("c_inv is None: pass ao_side=AO_BRA or AO_KET")
        bci, kci = as_f64(bra_ci), as_f64(ket_ci)
        n_mo = bra.shape[0]
        if bra.shape != (n_mo, n_mo) or ket.shape != (n_mo, n_mo) or (inv is not None and inv.shape != (n_mo, n_mo)):
            raise ValueError("MO matrices must be square and of equal shape")
        if bci.ndim != 3 or kci.ndim != 3 or bci.shape[1:] != (n_occ, n_mo - n_occ) or kci.shape[1:] != bci.shape[1:]:
            raise ValueError("CI coefficients must have shape (n_states, occ, n_mo - occ)")
        out = np.empty((bci.shape[0], kci.shape[0]))
        info = np.zeros(4, dtype=np.int32)
        rc = self.lib.wfo_overlaps_ci_host(self.handle, _hptr(bra), _hptr(ket), _hptr(inv) if inv is not None else None,
                                           int(ao_side if inv is None else AO_GIVEN), n_mo, int(n_occ),
                                           _hptr(bci), bci.shape[0], _hptr(kci), kci.shape[0], float(ci_thresh),
                                           int(semantics), TIERS[tier], int(rank), int(world), _hptr(out),
                                           _hptr(info))
        if rc == 4:
            raise ValueError("not enough values to unpack (expected 2, got 0)")   # pywfo/main.py:186
        check(rc)
        return out, info

    def overlap_from_lists_host(self, s_mo, alpha, beta, bra_coeffs, ket_coeffs):
        """General determinant lists.  ``alpha`` / ``beta`` = (bra_blocks, ket_blocks, bra_of, ket_of) or None
        (no electrons of that spin); coefficients (n_dets, n_states) with the string parity folded in.
        Returns (out (ns_bra, ns_ket), info)."""
        s_mo, bc, kc = as_f64(s_mo), as_f64(bra_coeffs), as_f64(ket_coeffs)
        nd_b, ns_b = bc.shape
        nd_k, ns_k = kc.shape
        args = []
        keep = []
        for spin in (alpha, beta):
            if spin is None:
                args.append((0, None, 0, None, 0, None, None))
                continue
            bb, kb, bof, kof = (as_i32(x) for x in spin)
            bb, kb = np.atleast_2d(bb), np.atleast_2d(kb)
            if bof.shape != (nd_b,) or kof.shape != (nd_k,) or bb.shape[1] != kb.shape[1]:
                raise ValueError("block maps do not match the determinant lists")
            keep += [bb, kb, bof, kof]
            args.append((bb.shape[1], _hptr(bb), bb.shape[0], _hptr(kb), kb.shape[0], _hptr(bof), _hptr(kof)))
        (na, ba, nba, ka, nka, baof, kaof), (nb, bbp, nbb, kbp, nkb, bbof, kbof) = args
        out = np.empty((ns_b, ns_k))
        info = np.zeros(4, dtype=np.int32)
        check(self.lib.wfo_overlap_from_lists_host(self.handle, _hptr(s_mo), s_mo.shape[0], s_mo.shape[1], na, ba, nba, ka,
                                                   nka, nb, bbp, nbb, kbp, nkb, nd_b, baof, bbof, _hptr(bc), ns_b, nd_k,
                                                   kaof, kbof, _hptr(kc), ns_k, _hptr(out), _hptr(info)))
        return out, info

    def trajectory_ci_host(self, mos, ci, n_occ, pairs, ci_thresh, semantics=0, tier="auto", ao_side=AO_KET,
                           rank=0, world=1):
        """State overlaps of (bra cycle, ket cycle) pairs of a trajectory; every cycle is uploaded once.
        mos (n_cycles, n_mo, n_mo), ci (n_cycles, n_states, n_occ, n_virt), pairs (n_pairs, 2).
        Returns (n_pairs, n_states, n_states) and info."""
        mos, ci = as_f64(mos), as_f64(ci)
        pr = as_i32(pairs).reshape(-1, 2)
        n_cycles, n_mo = mos.shape[0], mos.shape[1]
        if mos.ndim != 3 or mos.shape[2] != n_mo:
            raise ValueError("mos must have shape (n_cycles, n_mo, n_mo)")
        if ci.ndim != 4 or ci.shape[0] != n_cycles or ci.shape[2:] != (n_occ, n_mo - n_occ):
            raise ValueError("ci must have shape (n_cycles, n_states, occ, n_mo - occ)")
        ns = ci.shape[1]
        out = np.empty((pr.shape[0], ns, ns))
        info = np.zeros(4, dtype=np.int32)
        rc = self.lib.wfo_trajectory_ci_host(self.handle, _hptr(mos), _hptr(ci), n_cycles, n_mo, int(n_occ), ns,
                                             float(ci_thresh), int(semantics), TIERS[tier], int(ao_side), _hptr(pr),
                                             pr.shape[0], int(rank), int(world), _hptr(out), _hptr(info))
        if rc == 4:
            raise ValueError("not enough values to unpack (expected 2, got 0)")   # pywfo/main.py:186
        check(rc)
        return out, info

    # ---- device-pointer entry points (torch tensors used only as memory) ------
    @staticmethod
    def _dev(*tensors):
        """Device tensors must be dense row-major: the C ABI sees raw pointers."""
        for t in tensors:
            if not t.is_cuda or not t.is_contiguous():
                raise ValueError("device entry points need contiguous CUDA tensors "
                                 f"(got device={t.device}, strides={tuple(t.stride())})")

    def smo_dev(self, bra, s_ao, ket, out, stream=0):
        self._dev(bra, s_ao, ket, out)
        p, u = bra.shape
        q, v = ket.shape
        check(self.lib.wfo_smo_dev(self.handle, bra.data_ptr(), s_ao.data_ptr(), ket.data_ptr(), p, u, v, q,
                                   out.data_ptr(), None, stream))

    def smo_inv_dev(self, bra, ket, c_inv, out, work, stream=0):
        """S_MO = (bra Cinv)(ket Cinv)^T; ``work`` holds 2 n*n doubles."""
        self._dev(bra, ket, c_inv, out, work)
        if work.numel() < 2 * bra.shape[0] * bra.shape[0]:
            raise ValueError("work must hold 2 n*n doubles")
        check(self.lib.wfo_smo_inv_dev(self.handle, bra.data_ptr(), ket.data_ptr(), c_inv.data_ptr(), bra.shape[0],
                                       out.data_ptr(), work.data_ptr(), stream))

    def set_occ_hint(self, n_occ):
        """smo_inv_dev starts the inverse of S_MO[:n_occ,:n_occ] early for the next state-overlap call (0 = off)."""
        check(self.lib.wfo_set_occ_hint(self.handle, int(n_occ)))

    def sao_dev(self, c_inv, out, stream=0):
        self._dev(c_inv, out)
        check(self.lib.wfo_sao_dev(self.handle, c_inv.data_ptr(), c_inv.shape[0], out.data_ptr(), stream))

    def gemm_dev(self, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc, stream=0):
        check(self.lib.wfo_gemm_dev(self.handle, int(ta), int(tb), m, n, k, float(alpha), A.data_ptr(), lda,
                                    B.data_ptr(), ldb, float(beta), Cm.data_ptr(), ldc, stream))

    def det_table_dev(self, s_mo, n, rows, cols, out, stream=0):
        self._dev(rows, cols, out)
        check(self.lib.wfo_det_table_dev(self.handle, s_mo.data_ptr(), s_mo.shape[0], s_mo.stride(0), n,
                                         rows.data_ptr(), rows.shape[0], cols.data_ptr(), cols.shape[0],
                                         out.data_ptr(), stream))

    def exc_lists_dev(self, exc, n_occ, lists, stream=0):
        self._dev(exc, lists)
        check(self.lib.wfo_exc_lists_dev(self.handle, exc.data_ptr(), exc.shape[0], n_occ, lists.data_ptr(), stream))

    def overlaps_ci_dev(self, bra_mos, ket_mos, n_occ, bra_ci, ket_ci, out, ci_thresh, c_inv=None, ao_side=AO_KET,
                        semantics=0, tier="auto", rank=0, world=1, stream=0):
        """Whole path from raw inputs that already live on the device (contiguous CUDA tensors)."""
        self._dev(bra_mos, ket_mos, bra_ci, ket_ci, out)
        if c_inv is not None:
            self._dev(c_inv)
        info = np.zeros(4, dtype=np.int32)
        rc = self.lib.wfo_overlaps_ci_dev(self.handle, bra_mos.data_ptr(), ket_mos.data_ptr(),
                                          c_inv.data_ptr() if c_inv is not None else None,
                                          AO_GIVEN if c_inv is not None else int(ao_side), bra_mos.shape[0], int(n_occ),
                                          bra_ci.data_ptr(), bra_ci.shape[0], ket_ci.data_ptr(), ket_ci.shape[0],
                                          float(ci_thresh), int(semantics), TIERS[tier], int(rank), int(world),
                                          out.data_ptr(), _hptr(info), stream)
        if rc == 4:
            raise ValueError("not enough values to unpack (expected 2, got 0)")   # pywfo/main.py:186
        check(rc)
        return info

    def state_overlaps_dev(self, s_mo, n_occ, bra_exc, Cb, ket_exc, Ck, out, sigma=1.0, tier="auto",
                           k_begin=0, k_end=None, stream=0):
        self._dev(s_mo, bra_exc, Cb, ket_exc, Ck, out)
        kb, kk = bra_exc.shape[0], ket_exc.shape[0]
        k_end = kb if k_end is None else k_end
        check(self.lib.wfo_state_overlaps_dev(self.handle, s_mo.data_ptr(), s_mo.shape[0], int(n_occ),
                                              bra_exc.data_ptr(), kb, Cb.data_ptr(), Cb.shape[0],
                                              ket_exc.data_ptr(), kk, Ck.data_ptr(), Ck.shape[0], float(sigma),
                                              TIERS[tier], int(k_begin), int(k_end), out.data_ptr(), stream))

    # ---- bookkeeping -----------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.wfo_launch_count(self.handle))

    def last_tier(self) -> int:
        return int(self.lib.wfo_last_tier(self.handle))

    # ---- the in-library collective (K5) -----------------------------------------
    def comm_init(self, rank: int, world: int, tag: str, gather_bytes: int = 0):
        """Collective: every rank of the node calls this with the same ``world`` and job-unique ``tag``."""
        check(self.lib.wfo_comm_init(self.handle, int(rank), int(world), tag.encode(), int(gather_bytes)))

    def comm_destroy(self):
        check(self.lib.wfo_comm_destroy(self.handle))

    def comm_world(self) -> int:
        return int(self.lib.wfo_comm_world(self.handle))

    def allreduce_dev(self, t, stream=0):
        """Sum a float64 CUDA tensor over the ranks in place (rank order, bitwise identical everywhere)."""
        self._dev(t)
        check(self.lib.wfo_allreduce_dev(self.handle, t.data_ptr(), t.numel(), stream))

    def comm_status(self):
        check(self.lib.wfo_comm_status(self.handle))

    def last_inverse_info(self):
        """(Newton-Schulz iterations, pivoted fallback used) of the last inverse on this context."""
        it, pv = C.c_int(), C.c_int()
        check(self.lib.wfo_last_inverse_info(self.handle, C.byref(it), C.byref(pv)))
        return it.value, bool(pv.value)

    def set_lu_variant(self, variant):
        """Brute-force kernel for 32 < n_occ <= 128: 0/"auto", 1/"left" (on chip), 2/"right" (L2 scratch)."""
        v = {"auto": 0, "left": 1, "right": 2}.get(variant, variant)
        check(self.lib.wfo_set_lu_variant(self.handle, int(v)))

    def set_timing(self, on):
        """0 = off, 1 = dominant kernel only, 2 = phase breakdown (``last_phases``)."""
        check(self.lib.wfo_set_timing(self.handle, int(on)))

    def last_phases(self):
        """[(label, ms), ...] of the last state-overlap call (timing level 2)."""
        ms = (C.c_double * 24)()
        names = C.create_string_buffer(1024)
        k = int(self.lib.wfo_last_phases(self.handle, ms, names, 1024))
        labels = names.value.decode().split(",") if k else []
        return [(labels[i], ms[i]) for i in range(k)]

    def last_kernel_ms(self) -> float:
        return float(self.lib.wfo_last_kernel_ms(self.handle))

    def fp64_peak(self, kind: int, iters: int = 4096) -> float:
        t = C.c_double()
        check(self.lib.wfo_fp64_peak(self.handle, kind, iters, C.byref(t)))
        return t.value


def context(device: int | None = None) -> Context:
    """Process-wide context cache.  ``device=None`` -> LOCAL_RANK (torchrun) or 0."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx
