import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WFO_B200_LIB"] = os.path.join(ROOT, "pywfo_b200", "libwfo_dbg.so")
import bench
from pywfo_b200 import _lib, excitations as ex
for wl in ("c4", "c5"):
    w = bench.WORKLOADS[wl]; ctx = _lib.context(0); n = w["n_occ"]
    bra_mos, ket_mos, bci, kci = bench.make_inputs(w)
    S = bra_mos @ np.linalg.inv(ket_mos)
    rng = np.random.default_rng(0)
    v = w["n_mo"] - n
    def lists(k):
        out = []
        for _ in range(k):
            f, t = rng.integers(n), rng.integers(v)
            l = [i for i in range(n) if i != f] + [n + t]
            out.append(l)
        return np.array(out, dtype=np.int32)
    rows, cols = lists(148 * 8), lists(16)
    lib = ctx.lib
    buf = np.zeros(8, dtype=np.int64)
    for it in range(3):
        lib.wfo_debug_wpm_clocks(None, 1)
        D = ctx.det_table_host(S, rows, cols)
    lib.wfo_debug_wpm_clocks(buf.ctypes.data_as(C.c_void_p), 0)
    names = ["panel load", "16 column steps", "sign", "M = L21 L11^-1", "M write + rowmap", "trailing update"]
    tot = buf[:6].sum()
    nmat = (148 * 8 * 16) / (296 * 4)   # matrices per warp
    print(wl, "n =", n, "clk per matrix per warp:", int(tot / nmat))
    for nm, x in zip(names, buf): print(f"   {nm:18s} {100*x/tot:5.1f} %  {int(x/nmat):8d} clk/matrix")
