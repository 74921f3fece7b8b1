"""Warm per-phase device times (CUDA events inside the library, timing level 2) of one C4 step,
for the full bra shard (N=1) and a 1/8 shard (what one rank of 8 sees)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pywfo_b200 import _lib, excitations as ex  # noqa: E402

wname = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = bench.WORKLOADS[wname]
dev = torch.device("cuda", 0)
ctx = _lib.context(0)
n, n_mo = w["n_occ"], w["n_mo"]
bra_mos, ket_mos, bci, kci = bench.make_inputs(w)
thr = 0.0 if w["K"] is None else 1e-12
bt = ex.build_tables(bci, thr, "cache")
kt = ex.build_tables(kci, thr, "cache", order="locality")
c_inv = np.linalg.inv(ket_mos)
f64 = dict(dtype=torch.float64, device=dev)
d_bra, d_ket, d_inv = (torch.tensor(np.ascontiguousarray(x), **f64) for x in (bra_mos, ket_mos, c_inv))
d_work, d_smo = torch.empty((2, n_mo, n_mo), **f64), torch.empty((n_mo, n_mo), **f64)
d_be = torch.tensor(bt.exc, dtype=torch.int32, device=dev)
d_ke = torch.tensor(kt.exc, dtype=torch.int32, device=dev)
d_cb, d_ck = torch.tensor(bt.coeff, **f64), torch.tensor(kt.coeff, **f64)
d_out = torch.empty((bt.coeff.shape[0], kt.coeff.shape[0]), **f64)
stream = torch.cuda.current_stream().cuda_stream
ctx.set_timing(2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

for world, hint in ((1, 0), (1, n), (8, 0), (8, n)):
    ctx.set_occ_hint(hint)
    lo, hi = ex.shard_bounds(bt.K, world, 0)
    acc = {}
    order = []
    tot = []
    pre = []
    for it in range(8):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        ctx.smo_inv_dev(d_bra, d_ket, d_inv, d_smo, d_work, stream)
        e1.record()
        ctx.state_overlaps_dev(d_smo, n, d_be, d_cb, d_ke, d_ck, d_out, tier="auto", k_begin=lo, k_end=hi, stream=stream)
        e2.record()
        torch.cuda.synchronize()
        if it < 3:
            continue
        pre.append(e0.elapsed_time(e1))
        tot.append(e0.elapsed_time(e2))
        for name, ms in ctx.last_phases():
            if name not in acc:
                acc[name] = []
                order.append(name)
            acc[name].append(ms)
    print(f"--- {wname} shard 1/{world} occ_hint={hint}: step {np.mean(tot)*1e3:.1f} us, sao+smo {np.mean(pre)*1e3:.1f} us")
    for name in order:
        print(f"   {name:16s} {np.mean(acc[name])*1e3:9.1f} us")
