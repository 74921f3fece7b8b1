"""The N>1 path on CPU: world_size-2 gloo group, each rank computes the partial
matrix of its shard of the bra excitations (the ORACLE stands in for the CUDA
entry point, which cannot run here) and one all-reduce sums them."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import pywfo_oracle as orc
    from pywfo_b200 import dist as wdist
    from pywfo_b200 import excitations as ex

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # sharding is opt-in: a process group alone must not change what the entry points do
        assert wdist.world() == (0, 1)
        assert wdist.enable() == (rank, world)
        g = load_golden("r_water")
        occ = int(g["occ"])
        S = orc.mo_overlap(g["bra_mos"], g["ket_mos"], orc.get_S_AO(g["bra_mos"], g["ket_mos"], "ket"))
        bt = ex.build_tables(g["bra_ci"], 0.0, "cache")
        kt = ex.build_tables(g["ket_ci"], 0.0, "cache", order="locality")
        bra = [tuple(map(int, e)) for e in bt.exc]
        ket = [tuple(map(int, e)) for e in kt.exc]
        seen = []

        def partial(lo, hi):
            seen.append((lo, hi))
            return orc.state_overlaps_closed_form(S, occ, bra[lo:hi], bt.coeff[:, lo:hi], ket, kt.coeff, 1.0)

        out = wdist.sharded_state_overlaps(partial, bt.K, bt.coeff.shape[0], kt.coeff.shape[0])
        q.put((rank, seen[0], out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort(key=lambda x: x[0])
    (lo0, hi0), (lo1, hi1) = res[0][1], res[1][1]
    assert lo0 == 0 and hi0 == lo1 and hi1 == 95          # 5 occ x 19 virt, split 48/47
    np.testing.assert_array_equal(res[0][2], res[1][2])   # every rank holds the full sum
    g = load_golden("r_water")
    from oracle import pywfo_oracle as orc
    want, _ = orc.ref_overlaps_cache(g["bra_mos"], g["ket_mos"], g["bra_ci"], g["ket_ci"], int(g["occ"]), 0.0)
    np.testing.assert_allclose(res[0][2], want, rtol=1e-10, atol=1e-13)


# ---------------------------------------------------------------- trajectory pairs over ranks
def _traj_worker(rank, world, port, q, n_use):
    import torch.distributed as dist
    from oracle import pywfo_oracle as orc
    from pywfo_b200 import dist as wdist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wdist.enable()
        g = load_golden("cytosin_traj")
        occ, thr = int(g["occ"]), float(g["ci_thresh"])
        pairs = g["pairs"][:n_use]
        mo, ci = g["mo_coeffs"], g["ci_coeffs"]
        calls = []

        def compute(idx, r, w):
            calls.append((list(map(int, idx)), r, w))
            outs = []
            for i, j in pairs[idx]:
                full, _ = orc.ref_overlaps_cache(mo[i], mo[j], ci[i], ci[j], occ, thr)
                outs.append(full if w == 1 else full / w)   # stand-in for an excitation shard: partials add up
            return np.stack(outs)

        out = wdist.sharded_trajectory(compute, len(pairs), ci.shape[1])
        q.put((rank, calls, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("n_use", [3, 1])
def test_two_rank_trajectory(n_use):
    """3 pairs on 2 ranks: dealt round-robin (0,2 | 1); 1 pair on 2 ranks: both ranks take it with
    an excitation shard each.  Either way the all-reduce returns the per-pair reference values."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_traj_worker, args=(r, 2, port, q, n_use)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort(key=lambda x: x[0])
    if n_use == 3:
        assert res[0][1] == [([0, 2], 0, 1)] and res[1][1] == [([1], 0, 1)]
    else:
        assert res[0][1] == [([0], 0, 2)] and res[1][1] == [([0], 1, 2)]
    np.testing.assert_array_equal(res[0][2], res[1][2])
    g = load_golden("cytosin_traj")
    np.testing.assert_allclose(res[0][2], g["ref_overlaps_cache"][:n_use], rtol=1e-10, atol=1e-13)


# ---------------------------------------------------------------- a failure on one rank is raised on all
def _failing_worker(rank, world, port, q):
    import torch.distributed as dist
    from pywfo_b200 import dist as wdist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wdist.enable()

        def compute(idx, r, w):
            if rank == 1:
                raise ValueError("a state keeps no excitation")   # e.g. WFO_ERR_EMPTY_STATE of one pair
            return np.zeros((len(idx), 2, 2))

        try:
            wdist.sharded_trajectory(compute, 4, 2)
            q.put((rank, "returned"))
        except ValueError as e:
            q.put((rank, "ValueError: " + str(e)))
        except RuntimeError as e:
            q.put((rank, "RuntimeError: " + str(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_rank_local_failure_raises_everywhere():
    """a rank-local exception used to leave the other ranks blocked in the all-reduce."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_failing_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[1].startswith("ValueError")
    assert res[0].startswith("RuntimeError") and "another rank failed" in res[0]
