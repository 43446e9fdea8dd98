"""world_size-2 (and 4) CPU run of the slab decomposition's host-side logic over torch.distributed / gloo
(tests/gloo_worker.py); needs no GPU"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_slab_host_logic_over_gloo(world, gevb):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world), os.path.join(ROOT, "tests", "gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "GLOO SLAB OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
