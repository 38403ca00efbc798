"""Real multi-GPU checks (skipped unless the box exposes >= 2 GPUs): tools/dist_check.py under
torchrun — distributed framebuffer over peer memory (every partition / owner layout, batches back to
back), NCCL gathers, sort-last through NCCL and peer memory, data-parallel gradients — each against
single-GPU renders."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def test_dist_check_on_real_gpus(cuda):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, this box has {n}")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tools" / "dist_check.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0 and "ALL OK" in r.stdout, tail
