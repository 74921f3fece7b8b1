"""e2e time of the public entry point on pinned host arrays (median of 15), one workload."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pywfo_b200 import main as wmain
wname = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = bench.WORKLOADS[wname]
bra_mos, ket_mos, bci, kci = bench.make_inputs(w)
n = w["n_occ"]
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
pb, pk, pbc, pkc = (pin(x) for x in (bra_mos, ket_mos, bci, kci))
fn = wmain.overlaps_cache if w["sigma"] > 0 else wmain.overlaps
thr = 0.0 if w["K"] is None else 1e-12
ts = []
for i in range(20):
    torch.cuda.synchronize(); t = time.perf_counter()
    fn(pb, pk, pbc, pkc, n, ci_thresh=thr, ao_ovlps="ket")
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
ts = np.array(ts[5:]) * 1e3
print(f"{wname} e2e median {np.median(ts):.4f} ms  min {ts.min():.4f}  max {ts.max():.4f}  (WFO_NO_EARLY_BASE={os.environ.get('WFO_NO_EARLY_BASE')})")
