"""Small pass through every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pywfo_b200 import _lib, helpers, main, trajectory
ctx = _lib.context(0)
rng = np.random.default_rng(0)
for (n_mo, occ, nb, nk) in ((30, 9, 3, 4), (70, 40, 2, 9), (116, 100, 2, 3)):
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(7, n_mo, occ, nb, nk)
    for tier in ("update", "lu", "closed"):
        out = main.overlaps_cache(bra_mos, ket_mos, bci, kci, occ, ci_thresh=0.12 if tier == "lu" else 1e-4, tier=tier)
        print(n_mo, occ, tier, float(np.abs(out).max()))
mos = np.stack([helpers.synthetic_case(s, 24, 5, 2, 2)[0] for s in range(3)])
cis = np.stack([helpers.synthetic_case(s, 24, 5, 2, 2)[2] for s in range(3)])
print(trajectory.trajectory_overlaps(mos, cis, 5).shape)
print(ctx.inverse_host(rng.standard_normal((50, 50))).shape)
