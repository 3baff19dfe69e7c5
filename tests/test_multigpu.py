"""Sort-last on real GPUs (needs >= 2): one process per GPU under torchrun, NCCL + CUDA IPC peer memory; the worker
(tests/mgpu_worker.py) compares the multi-GPU frame with the single-GPU frame bit for bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["tiny", "c1"])
def test_sort_last_frame_equals_single_gpu_frame(scene):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    port = 29500 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), "--scene", scene]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert r.returncode == 0 and "MGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
