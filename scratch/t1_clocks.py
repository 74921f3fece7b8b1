"""Per-CTA elapsed clocks of the rank-update kernel (debug build libwfo_dbg.so, -DWFO_T1_CLOCKS)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WFO_B200_LIB"] = os.path.join(ROOT, "pywfo_b200", "libwfo_dbg.so")
import bench  # noqa: E402
from pywfo_b200 import _lib, excitations as ex  # noqa: E402

w = bench.WORKLOADS["c4"]
dev = torch.device("cuda", 0)
ctx = _lib.context(0)
n, n_mo = w["n_occ"], w["n_mo"]
bra_mos, ket_mos, bci, kci = bench.make_inputs(w)
bt = ex.build_tables(bci, 0.0, "cache")
kt = ex.build_tables(kci, 0.0, "cache", order="locality")
S = bra_mos @ np.linalg.inv(ket_mos)
f64 = dict(dtype=torch.float64, device=dev)
d_smo = torch.tensor(S, **f64)
d_be = torch.tensor(bt.exc, dtype=torch.int32, device=dev)
d_ke = torch.tensor(kt.exc, dtype=torch.int32, device=dev)
d_cb, d_ck = torch.tensor(bt.coeff, **f64), torch.tensor(kt.coeff, **f64)
d_out = torch.empty((50, 50), **f64)
stream = torch.cuda.current_stream().cuda_stream
lib = ctx.lib
for world in (1, 8):
    lo, hi = ex.shard_bounds(bt.K, world, 0)
    for it in range(3):
        ctx.state_overlaps_dev(d_smo, n, d_be, d_cb, d_ke, d_ck, d_out, tier="update", k_begin=lo, k_end=hi, stream=stream)
        torch.cuda.synchronize()
    buf = np.zeros((148, 4), dtype=np.int64)
    lib.wfo_debug_t1_clocks(buf.ctypes.data_as(C.c_void_p), 148)
    clk, sm, t0, t1 = buf.T
    print(f"world {world}: clocks min {clk.min()} med {np.median(clk):.0f} max {clk.max()}  "
          f"span ns: first start {0} last start {t0.max()-t0.min()} first end {t1.min()-t0.min()} last end {t1.max()-t0.min()}")
    # per SM
    per_sm = {}
    for c, s_ in zip(clk, sm):
        per_sm.setdefault(int(s_), []).append(int(c))
    cnt = np.bincount([len(v) for v in per_sm.values()])
    print("   CTAs per SM histogram:", cnt.tolist(), " SMs used:", len(per_sm))
    ends = np.sort(t1 - t0.min())
    print("   end-time percentiles (us):", [round(float(np.percentile(ends, q)) / 1e3, 1) for q in (0, 10, 50, 90, 99, 100)])
    slow = np.argsort(clk)[-8:]
    print("   slowest CTAs (cta, sm, clk):", [(int(i), int(sm[i]), int(clk[i])) for i in slow])
