"""One warm C4 state-overlap call (for ncu -k captures)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pywfo_b200 import _lib, excitations as ex
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
tier = sys.argv[2] if len(sys.argv) > 2 else "auto"
world = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = _lib.context(0)
n, n_mo = w["n_occ"], w["n_mo"]
bra_mos, ket_mos, bci, kci = bench.make_inputs(w)
thr = 0.0 if w["K"] is None else 1e-12
bt = ex.build_tables(bci, thr, "cache"); kt = ex.build_tables(kci, thr, "cache", order="locality")
S = bra_mos @ np.linalg.inv(ket_mos)
f64 = dict(dtype=torch.float64, device="cuda")
d_smo = torch.tensor(S, **f64)
d_be = torch.tensor(bt.exc, dtype=torch.int32, device="cuda"); d_ke = torch.tensor(kt.exc, dtype=torch.int32, device="cuda")
d_cb, d_ck = torch.tensor(bt.coeff, **f64), torch.tensor(kt.coeff, **f64)
d_out = torch.empty((d_cb.shape[0], d_ck.shape[0]), **f64)
lo, hi = ex.shard_bounds(bt.K, world, 0)
if tier == "lu":   # 4096 pairs: what the round-1 captures used (the DRAM traffic is compared per 4096 pairs)
    hi = min(hi, 64); d_ke = d_ke[:64].contiguous(); d_ck = d_ck[:, :64].contiguous()
for _ in range(3):
    ctx.state_overlaps_dev(d_smo, n, d_be, d_cb, d_ke, d_ck, d_out, tier=tier, k_begin=lo, k_end=hi,
                           stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
print("ok", float(d_out.abs().max()))
