"""torchrun --nproc-per-node 2 scratch/traj_2gpu.py : trajectory driver and overlaps_cache under NCCL."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from pywfo_b200 import dist as wdist, main, trajectory
rank, world = wdist.init_from_env()
g = load_golden("cytosin_traj")
occ, thr = int(g["occ"]), float(g["ci_thresh"])
for pairs in (g["pairs"], g["pairs"][:1]):          # dealt pairs / one pair sharded by excitations
    got = trajectory.trajectory_overlaps(g["mo_coeffs"], g["ci_coeffs"], occ, ci_thresh=thr, pairs=pairs)
    want = g["ref_overlaps_cache"][: len(pairs)]
    err = np.abs(got - want).max() / np.abs(want).max()
    print(f"rank {rank}/{world} pairs {len(pairs)} max rel err {err:.2e}")
    assert err < 1e-10
h = load_golden("r_n34")
got = main.overlaps_cache(h["bra_mos"], h["ket_mos"], h["bra_ci"], h["ket_ci"], int(h["occ"]), ci_thresh=float(h["ci_thresh"]))
err = np.abs(got - h["ref_overlaps_cache"]).max() / np.abs(h["ref_overlaps_cache"]).max()
print(f"rank {rank} overlaps_cache err {err:.2e}")
assert err < 1e-10
import torch.distributed as dist
dist.destroy_process_group()
