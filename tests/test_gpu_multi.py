"""GPU, >= 2 devices: row-sharded run under torchrun vs a single-GPU run (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    sys.path.insert(0, ROOT)
    from gcmapexplorer_b200 import _lib
    return _lib.device_count()


@pytest.mark.parametrize("n,partition", [(6001, "rows"), (9000, "equal_area"), (9000, "balanced"), (12345, "balanced")])
def test_sharded_matches_single_gpu(n, partition):
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if g < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + n % 97), os.path.join(ROOT, "tools", "multi_gpu_check.py"), str(n), partition]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
