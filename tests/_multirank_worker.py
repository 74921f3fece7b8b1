"""Worker of tests/test_gpu_multirank.py (launched with torch.distributed.run, one rank per GPU).
Every rank computes its shard through the C ABI; the sum must equal the golden values that the
unmodified reference produced (tests/golden/*.npz), for both reduction semantics and for both
collectives (in-library peer-memory all-reduce, NCCL fallback)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_golden
    from pywfo_b200 import _lib, dist as wdist, main as wmain, helpers
    from oracle import pywfo_oracle as orc

    in_library = sys.argv[1] == "peer"
    rank, world = wdist.init_from_env(in_library=in_library)
    local = int(os.environ["LOCAL_RANK"])
    ctx = _lib.context(local)
    assert wdist.has_comm(ctx) == in_library
    checked = []

    def same_everywhere(a):
        t = torch.tensor(np.ascontiguousarray(a), device=f"cuda:{local}")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks hold different results"

    # goldens from the unmodified reference, both semantics
    for name in ("r_n17", "r_odd", "r_water", "r_n34", "r_sparse"):
        g = load_golden(name)
        occ, thr = int(g["occ"]), float(g["ci_thresh"])
        a = (g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"])
        nmin = min(a[2].shape[0], a[3].shape[0])
        for tier in ("update", "lu"):
            got = wmain.overlaps_cache(*a, occ, ci_thresh=thr, tier=tier)
            want = g["ref_overlaps_cache"]
            np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
            same_everywhere(got)
            got = wmain.overlaps(a[0], a[1], a[2][:nmin], a[3][:nmin], occ, ci_thresh=thr, tier=tier)
            want = g["ref_overlaps"]
            np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
            same_everywhere(got)
            checked.append((name, tier))
    g = load_golden("h2o2")
    a = (g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"])
    for fn, key in ((wmain.overlaps_cache, "ref_cache_4x4_1e-2"), (wmain.overlaps, "ref_minus_4x4_1e-2")):
        got = fn(*a, 9, ci_thresh=1e-2)
        np.testing.assert_allclose(got, g[key], rtol=1e-10, atol=1e-10 * np.abs(g[key]).max())
        checked.append(("h2o2", key))
    # a larger case: more excitations than ranks x tiles, ragged shards; checker = oracle closed form
    bra_mos, ket_mos, bci, kci = helpers.synthetic_case(5, 60, 11, 6, 7)
    S = orc.mo_overlap(bra_mos, ket_mos, orc.get_S_AO(bra_mos, ket_mos, "ket"))
    want = orc.state_overlaps_factored(S, 11, orc.signed_ci(bci, 11), orc.signed_ci(kci, 11), 1.0)
    got = wmain.overlaps_cache(bra_mos, ket_mos, bci, kci, 11, ci_thresh=0.0)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
    same_everywhere(got)
    # explicit AO overlaps (host tables, sharded state_overlaps + all-reduce)
    S_AO = orc.get_S_AO(bra_mos, ket_mos, "ket")
    got2 = wmain.overlaps_cache(bra_mos, ket_mos, bci, kci, 11, ci_thresh=0.0, ao_ovlps=S_AO)
    np.testing.assert_allclose(got2, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
    # the collective itself: rank-dependent data, result = analytic sum, identical bits on all ranks
    if in_library:
        x = torch.arange(5000, dtype=torch.float64, device=f"cuda:{local}") * (rank + 1) + 0.1 * rank
        for _ in range(3):   # consecutive collectives reuse the double-buffered slots
            y = x.clone()
            ctx.allreduce_dev(y, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ctx.comm_status()
            want_y = sum(torch.arange(5000, dtype=torch.float64) * (r + 1) + 0.1 * r for r in range(world))
            assert torch.allclose(y.cpu(), want_y, rtol=1e-15, atol=0)
            same_everywhere(y.cpu().numpy())
    dist.barrier()
    wdist.disable()
    if rank == 0:
        print(f"MULTIRANK_OK world={world} collective={sys.argv[1]} cases={len(checked)}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
