"""Multi-GPU parity (-m gpu, needs >= 2 devices): launches tests/multigpu_worker.py under torchrun, one process per GPU
over NCCL, and requires decomposition independence against the single-lattice CPU checker."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_decomposition(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-6000:])
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
