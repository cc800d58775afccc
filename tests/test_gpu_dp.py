"""Data parallelism on real hardware (NCCL): runs tests/dp_worker.py under torch.distributed.run with 2 ranks.
Skipped when fewer than 2 GPUs are visible (the driver's single-GPU box); run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_parity():
    env = dict(os.environ)
    env.pop("UNETK_DP_GRAPH", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "dp_worker.py")],
                       capture_output=True, text=True, timeout=420, cwd=ROOT, env=env)
    tail = (r.stdout + "\n" + r.stderr)[-4000:]
    assert r.returncode == 0 and "DP_WORKER_OK" in r.stdout, tail
    print([ln for ln in r.stdout.splitlines() if "DP_WORKER_OK" in ln][-1])
