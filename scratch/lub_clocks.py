import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WFO_B200_LIB"] = os.path.join(ROOT, "pywfo_b200", "libwfo_dbg.so")
import bench
from pywfo_b200 import _lib, excitations as ex
w = bench.WORKLOADS["c4"]; ctx = _lib.context(0); n = w["n_occ"]
bra_mos, ket_mos, bci, kci = bench.make_inputs(w)
bt = ex.build_tables(bci[:2], 0.0, "cache"); kt = ex.build_tables(kci[:2], 0.0, "cache", order="locality")
S = bra_mos @ np.linalg.inv(ket_mos)
f64 = dict(dtype=torch.float64, device="cuda")
d_smo = torch.tensor(S, **f64)
d_be = torch.tensor(bt.exc[:512], dtype=torch.int32, device="cuda"); d_ke = torch.tensor(kt.exc[:512], dtype=torch.int32, device="cuda")
d_cb, d_ck = torch.tensor(bt.coeff[:, :512].copy(), **f64), torch.tensor(kt.coeff[:, :512].copy(), **f64)
d_out = torch.empty((2, 2), **f64)
lib = ctx.lib
buf = np.zeros(16, dtype=np.int64)
for it in range(4):
    lib.wfo_debug_lub_clocks(None, 1)
    ctx.state_overlaps_dev(d_smo, n, d_be, d_cb, d_ke, d_ck, d_out, tier="update", stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    lib.wfo_debug_lub_clocks(buf.ctypes.data_as(C.c_void_p), 0)
names = ["load", "first publish", "loop top/after bar", "pp/pd + pr shuffles", "f + done", "look-ahead", "update", "barrier wait", "exit", "store+perm"]
for nm, v in zip(names, buf): print(f"{nm:22s} {v:9d} clk  ({v/n:.0f}/step)")
print("total", buf.sum(), "clk =", buf.sum() / 1.965e3, "us")
