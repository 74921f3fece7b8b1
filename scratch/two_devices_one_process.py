"""One process, two devices: every kernel attribute must be set per device."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from pywfo_b200 import main, helpers
g = load_golden("r_n34")
for dev in (0, 1, 0, 1):
    for tier in ("update", "lu", "closed"):
        got = main.overlaps_cache(g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], int(g["occ"]),
                                  ci_thresh=float(g["ci_thresh"]), tier=tier, device=dev)
        err = np.abs(got - g["ref_overlaps_cache"]).max() / np.abs(g["ref_overlaps_cache"]).max()
        assert err < 1e-10, (dev, tier, err)
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(2, 500, 100, 4, 4, top_k=300)
    a = main.overlaps_cache(bra_mos, ket_mos, bci, kci, 100, ci_thresh=1e-12, device=dev)
    print("device", dev, "ok", float(np.abs(a).max()))
