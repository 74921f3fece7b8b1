"""Run the Option-B stub of INTEGRATION.md verbatim against a golden case."""
import os, re, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
block = re.findall(r"```python\n(import ctypes as C, numpy as np.*?)```", md, flags=re.S)[0]
block = block.replace('C.CDLL("libwfo_b200.so")', f'C.CDLL("{ROOT}/pywfo_b200/libwfo_b200.so")')
ns = {}
exec(block, ns)
g = load_golden("r_n17")
for ao, key in (("ket", "ref_overlaps_cache"), ("bra", "ref_overlaps_cache_bra")):
    got = ns["overlaps_cache"](g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], int(g["occ"]),
                               ci_thresh=float(g["ci_thresh"]), ao_ovlps=ao)
    err = np.abs(got - g[key]).max() / np.abs(g[key]).max()
    print(ao, "stub max rel err", err)
    assert err < 1e-10
