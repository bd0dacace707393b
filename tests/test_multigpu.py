"""Hardware parity of the multi-GPU exchange paths (symmetric-memory peer push / peer stores / NCCL fallback, run(),
run_cyclic(), run_from_host(), consecutive steps) against the single-GPU result: tests/multigpu_worker.py under
torch.distributed.run with one process per visible GPU (2, 4 or 8).  Skipped on a single-GPU box; the gloo tests of
tests/test_pipeline.py cover the host logic there."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_multigpu_exchange_paths_match_single_gpu():
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    n = 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 0
    if n == 0:
        pytest.skip("needs at least 2 GPUs")
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-6000:])
    sys.stderr.write(res.stderr[-3000:])
    assert res.returncode == 0 and "MULTIGPU PARITY OK" in res.stdout
