"""Brute-force tier on samples of a workload: TFLOP/s of the fused pair kernel (kernel time from the library's events)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pywfo_b200 import _lib, excitations as ex
wname = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = bench.WORKLOADS[wname]
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
ctx.set_timing(True)
peak = max(ctx.fp64_peak(0, 8192), ctx.fp64_peak(1, 8192))
for ks, kl in ((64, 64), (256, 64), (1024, 64), (2048, 64), (1024, 256)):
    ks, kl = min(ks, bt.K), min(kl, kt.K)
    ke, ck = d_ke[:kl].contiguous(), d_ck[:, :kl].contiguous()
    ts = []
    for i in range(5):
        ctx.state_overlaps_dev(d_smo, n, d_be, d_cb, ke, ck, d_out, tier="lu", k_begin=0, k_end=ks,
                               stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        if i >= 2: ts.append(ctx.last_kernel_ms())
    ms = float(np.mean(ts)); fl = ks * kl * (bench.f_det(n) + 2 * w["n_ket"])
    print(f"{wname} n={n} sample {ks}x{kl}={ks*kl} pairs: {ms:.3f} ms  {fl/ms/1e9:.2f} TFLOP/s  frac {fl/ms/1e9/peak:.4f}")
