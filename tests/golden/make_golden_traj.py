"""Generate tests/golden/cytosin_traj.npz from the UNMODIFIED reference (build container only).

Run from the repo root:  python tests/golden/make_golden_traj.py
Reads the first cycles of the reference's bundled cytosine trajectory
(pywfo/tests/ref_cytosin/cytosin_overlap_data.h5; old-format HDF5, contiguous little-endian
float64 datasets: mo_coeffs (9,137,137) at byte 2048, ci_coeffs (9,2,29,108) right behind it,
checked against the Turbomole text files mos.1 / mos.2 next to it) and stores, for pairs of
cycles, what pywfo.main.overlaps_cache / overlaps return when called pair by pair -- the loop
that pywfo_b200.trajectory.trajectory_overlaps replaces with one library call.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import REF, load_reference, quiet  # noqa: E402

N_CYC, N_MO, N_ST, OCC = 9, 137, 2, 29
VIRT = N_MO - OCC
KEEP = 4          # cycles stored
THR = 2e-2


def main():
    m, d, h = load_reference()
    raw = open(os.path.join(REF, "pywfo/tests/ref_cytosin/cytosin_overlap_data.h5"), "rb").read()
    mo = np.frombuffer(raw, dtype="<f8", count=N_CYC * N_MO * N_MO, offset=2048).reshape(N_CYC, N_MO, N_MO)
    off = 2048 + N_CYC * N_MO * N_MO * 8
    ci = np.frombuffer(raw, dtype="<f8", count=N_CYC * N_ST * OCC * VIRT, offset=off).reshape(N_CYC, N_ST, OCC, VIRT)
    from pywfo_b200 import formats
    assert np.allclose(mo[0], formats.read_turbomole_mos(os.path.join(REF, "pywfo/tests/ref_cytosin/mos.1")), atol=1e-12)
    assert np.allclose(mo[2], formats.read_turbomole_mos(os.path.join(REF, "pywfo/tests/ref_cytosin/mos.2")), atol=1e-12)
    nrm = np.linalg.norm(ci.reshape(N_CYC, N_ST, -1), axis=2)
    assert np.all((nrm > 0.9) & (nrm < 1.01)), nrm
    mo, ci = mo[:KEEP].copy(), ci[:KEEP].copy()
    pairs = np.array([(0, 1), (1, 2), (2, 3), (0, 2), (3, 3)], dtype=np.int32)
    g = dict(mo_coeffs=mo, ci_coeffs=ci, occ=OCC, ci_thresh=THR, pairs=pairs)
    g["ref_overlaps_cache"] = np.stack([quiet(m.overlaps_cache, mo[i], mo[j], ci[i], ci[j], OCC, ci_thresh=THR)
                                        for i, j in pairs])
    g["ref_overlaps"] = np.stack([quiet(m.overlaps, mo[i], mo[j], ci[i], ci[j], OCC, ci_thresh=THR) for i, j in pairs])
    g["ref_overlaps_cache_bra"] = np.stack([quiet(m.overlaps_cache, mo[i], mo[j], ci[i], ci[j], OCC, ci_thresh=THR,
                                                  ao_ovlps="bra") for i, j in pairs[:2]])
    np.savez_compressed(os.path.join(HERE, "cytosin_traj.npz"), **g)
    print(np.round(g["ref_overlaps_cache"], 4))


if __name__ == "__main__":
    main()
