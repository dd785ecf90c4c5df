"""Multi-GPU path on real devices (needs >= 2 GPUs; the CPU-side logic is covered with gloo in
tests/test_host_api.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_two_rank_sharded_runs_match_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(here, "mp_sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "photon shard check: ok" in r.stdout, r.stdout[-3000:]
    for mode, used in (("p2p", "GravityExchangeP2P"), ("nccl", "GravityExchange")):  # both exchanges, and the one asked for ran
        assert "gravity shard check (%s -> %s): ok" % (mode, used) in r.stdout, r.stdout[-3000:]
