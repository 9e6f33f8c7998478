"""Two ranks on two GPUs of one box (skipped on single-GPU boxes): the sharded facade gives the same joint
fit and predictions as the single-process object.  Host-side sharding logic is covered on CPU by
tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_facade_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(here, "multigpu_sharded_script.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sharded facade ok" in out.stdout and "empty shard and mixed large/small shards ok" in out.stdout
