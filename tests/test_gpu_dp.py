"""Multi-GPU parity of the fused data-parallel exchange (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_fused_exchange_matches_all_reduce():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = int(os.environ.get("NK_DP_TEST_WORLD", "2"))
    if world > n:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(root, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=170, cwd=root)
    assert r.returncode == 0 and "DP_WORKER OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
