"""Where the end-to-end time of pywfo_b200.main.overlaps_cache goes (C4, host numpy inputs)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pywfo_b200 import _lib, main as wmain

w = bench.WORKLOADS["c4"]
bra_mos, ket_mos, bci, kci = bench.make_inputs(w)
n = w["n_occ"]
ctx = _lib.context(0)

def pin(x):
    t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    return t.numpy()

def timeit(f, reps=5):
    f(); ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return np.median(ts) * 1e3

print("host cores", os.cpu_count())
print("np.linalg.inv(500x500): %.2f ms" % timeit(lambda: np.linalg.inv(ket_mos)))
c_inv = np.linalg.inv(ket_mos)
print("overlaps_ci_host pageable: %.2f ms" % timeit(lambda: ctx.overlaps_ci_host(bra_mos, ket_mos, c_inv, n, bci, kci, 0.0)))
pb, pk, pi, pbc, pkc = (pin(x) for x in (bra_mos, ket_mos, c_inv, bci, kci))
print("overlaps_ci_host pinned:   %.2f ms" % timeit(lambda: ctx.overlaps_ci_host(pb, pk, pi, n, pbc, pkc, 0.0)))
print("overlaps_cache pageable:   %.2f ms" % timeit(lambda: wmain.overlaps_cache(bra_mos, ket_mos, bci, kci, n, ci_thresh=0.0)))
print("overlaps_cache pinned:     %.2f ms" % timeit(lambda: wmain.overlaps_cache(pb, pk, pbc, pkc, n, ci_thresh=0.0)))
# raw copies
d = torch.empty(bci.shape, dtype=torch.float64, device="cuda")
tb = torch.from_numpy(bci); tp = torch.from_numpy(pbc)
print("H2D 16 MB pageable: %.2f ms   pinned: %.2f ms" % (timeit(lambda: d.copy_(tb)), timeit(lambda: d.copy_(tp, non_blocking=True))))
