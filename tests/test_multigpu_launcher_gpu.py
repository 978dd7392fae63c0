"""Runs the torchrun-launched multi-GPU tests (tests/test_syncbn_multigpu.py: SyncBN NVLink exchange, data-parallel static step,
overlapped gradient all-reduce) from the ordinary `pytest -m gpu` session whenever the box exposes at least two GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(int(os.environ.get("WORLD_SIZE", "1")) > 1, reason="already inside torchrun")
def test_multigpu_suite_under_torchrun():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the data-parallel tests need two (tools/run_multigpu_tests.sh 2)")
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "run_multigpu_tests.sh"), "2"], cwd=ROOT, env=env, timeout=900,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]
