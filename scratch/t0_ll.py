"""Brute-force determinant kernel: correctness vs LAPACK and throughput.
python scratch/t0_ll.py  [n ...]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pywfo_b200 import _lib, helpers, excitations as ex
ns = [int(x) for x in sys.argv[1:]] or [32, 34, 40, 64, 100, 128]
ctx = _lib.context(0)
dev = torch.device("cuda:0"); f64 = dict(dtype=torch.float64, device=dev)
PEAK = 37.17
for n in ns:
    n_mo = 5 * n
    np.random.seed(2)
    _, bra_mos, ket_mos = helpers.get_bra_ket_dummy_mos(n_mo, n)
    S_id = bra_mos @ np.linalg.inv(ket_mos)
    rng = np.random.default_rng(n)
    S_rand = rng.standard_normal((n_mo, n_mo))
    for name, S in (("near-identity", S_id), ("random", S_rand)):
        ksh, kl = 256, 512
        v = n_mo - n
        bra = np.stack([rng.integers(0, n, ksh), rng.integers(0, v, ksh)], 1).astype(np.int32)
        ket = np.stack([rng.integers(0, n, kl), rng.integers(0, v, kl)], 1).astype(np.int32)
        rl, cl = ex.exc_index_lists(bra, n), ex.exc_index_lists(ket, n)
        if name == "random":   # arbitrary index lists
            rl = np.stack([rng.permutation(n_mo)[:n] for _ in range(ksh)]).astype(np.int32)
            cl = np.stack([rng.permutation(n_mo)[:n] for _ in range(kl)]).astype(np.int32)
        dS = torch.tensor(np.ascontiguousarray(S), **f64)
        rows = torch.tensor(rl, dtype=torch.int32, device=dev)
        cols = torch.tensor(cl, dtype=torch.int32, device=dev)
        D = torch.empty((ksh, kl), **f64)
        import os
        ctx.set_lu_variant(os.environ.get("LU", "auto"))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for r in range(8):
            e0.record(); ctx.det_table_dev(dS, n, rows, cols, D, torch.cuda.current_stream().cuda_stream); e1.record(); torch.cuda.synchronize()
            if r >= 3: ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts))
        fdet = n*(n-1)//2 + n*(n-1)*(2*n-1)//3 + (n-1)
        Dh = D.cpu().numpy()
        worst = 0.0
        for k in range(0, ksh, 17):
            sub = S[rl[k]][:, cl[::7]] if False else None
            for l in range(0, kl, 23):
                sgn, ld = np.linalg.slogdet(S[rl[k]][:, cl[l]])
                ref = sgn * np.exp(ld)
                worst = max(worst, abs(Dh[k, l] - ref) / abs(ref))
        tf = ksh*kl*fdet/ms/1e9
        print(f"n={n:4d} {name:14s} pairs={ksh*kl} ms={ms:.3f} TF/s={tf:.2f} frac={tf/PEAK:.3f} max_rel_err={worst:.2e}", flush=True)
