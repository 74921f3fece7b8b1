"""T0 kernel timing on a sample: python scratch/t0_prof.py n n_mo ksh kl reps"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pywfo_b200 import _lib, helpers, excitations as ex
n, n_mo, ksh, kl, reps = (int(x) for x in sys.argv[1:6])
ctx = _lib.context(0)
np.random.seed(2)
_, bra_mos, ket_mos = helpers.get_bra_ket_dummy_mos(n_mo, n)
S = bra_mos @ np.linalg.inv(ket_mos)
rng = np.random.default_rng(0)
v = n_mo - n
bra = np.stack([rng.integers(0, n, ksh), rng.integers(0, v, ksh)], 1).astype(np.int32)
ket = np.stack([rng.integers(0, n, kl), rng.integers(0, v, kl)], 1).astype(np.int32)
dev = torch.device("cuda:0"); f64 = dict(dtype=torch.float64, device=dev)
dS = torch.tensor(np.ascontiguousarray(S), **f64)
rows = torch.tensor(ex.exc_index_lists(bra, n), dtype=torch.int32, device=dev)
cols = torch.tensor(ex.exc_index_lists(ket, n), dtype=torch.int32, device=dev)
D = torch.empty((ksh, kl), **f64)
ctx.set_timing(True)
ts = []
for r in range(reps + 2):
    ctx.det_table_dev(dS, n, rows, cols, D, 0); torch.cuda.synchronize()
    if r >= 2: ts.append(ctx.last_kernel_ms())
ms = float(np.mean(ts))
fdet = n*(n-1)//2 + n*(n-1)*(2*n-1)//3 + (n-1)
print(f"n={n} pairs={ksh*kl} ms={ms:.3f} pairs/s={ksh*kl/ms*1e3:.3e} TF/s={ksh*kl*fdet/ms/1e9:.2f} clk/matrix/CTA~{ms*1e-3*1.965e9/(ksh*kl/ (148*2)):.0f}")
# check vs numpy on a few
Dh = D.cpu().numpy()
for (k,l) in [(0,0),(ksh-1,kl-1)]:
    ref = np.linalg.det(S[ex.exc_index_lists(bra[k:k+1], n)[0]][:, ex.exc_index_lists(ket[l:l+1], n)[0]])
    print(k, l, Dh[k,l], ref, abs(Dh[k,l]-ref)/abs(ref))
import ctypes, os
if os.environ.get("WFO_B200_LIB"):
    lib = _lib.load()
    if hasattr(lib, "wfo_debug_phases"):
        buf = (ctypes.c_longlong * 16)()
        lib.wfo_debug_phases(buf, 1)
        ctx.det_table_dev(dS, n, rows, cols, D, 0); torch.cuda.synchronize()
        lib.wfo_debug_phases(buf, 0)
        names = ["gather", "panel", "permute", "trsm", "upd8", "upd16", "sign", "matrices"]
        nm = max(1, buf[7])
        tot = sum(buf[i] for i in range(7))
        print("phase cycles per matrix (CTA 0):", {names[i]: int(buf[i] / nm) for i in range(7)}, "total", int(tot / nm), "matrices", nm)
