"""N > 1 on hardware: one process per GPU (torch.distributed.run, NCCL rendezvous), skipped on a box
with a single GPU.  Asserts sum-of-shards == the reference's golden values for both semantics, with
the in-library peer-memory collective (csrc/comm.cuh) and with the NCCL fallback; and, in ONE process,
that two contexts on two devices given rank/world = 0/2 and 1/2 produce partials that add up."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(900)
@pytest.mark.parametrize("collective", ["peer", "nccl"])
def test_torchrun_two_ranks_match_reference_goldens(collective):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_multirank_worker.py"), collective]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=800, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTIRANK_OK world=2" in res.stdout


@pytest.mark.parametrize("semantics,key", [(0, "ref_overlaps_cache"), (1, "ref_overlaps")])
def test_two_devices_one_process_shards_add_up(semantics, key):
    """rank/world through the C ABI without any process group: device 0 takes shard 0/2, device 1 shard 1/2."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from pywfo_b200 import _lib
    g = load_golden("r_n17")
    occ, thr = int(g["occ"]), float(g["ci_thresh"])
    parts = []
    for dev in (0, 1):
        ctx = _lib.context(dev)
        part, _ = ctx.overlaps_ci_host(g["bra_mos"], g["ket_mos"], None, occ, g["bra_ci"], g["ket_ci"], thr,
                                       semantics=semantics, rank=dev, world=2, ao_side=_lib.AO_KET)
        parts.append(part)
    want = g[key]
    np.testing.assert_allclose(parts[0] + parts[1], want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
    assert np.abs(parts[0] - want).max() > 1e-6 * np.abs(want).max()   # really partial sums
