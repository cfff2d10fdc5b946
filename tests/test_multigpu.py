"""Spawns tests/multigpu_worker.py under torchrun when the box has at least two GPUs (the driver's 1-GPU test box skips
it; `gpurun --gpus 2` runs it).  The N > 1 HOST logic is covered without GPUs by tests/test_multirank_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_multigpu_worker():
    from freud_b200 import _capi

    n = _capi.lib().fgpu_device_count()
    if n < 2:
        pytest.skip(f"{n} GPU(s) visible: the multi-GPU worker needs 2")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multigpu_worker.py")]
    proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-6000:]
    assert proc.stdout.count("MULTIGPU OK") == world, proc.stdout[-6000:]
