"""Multi-GPU parity inside the GPU suite (was builder-run only): spawns one NCCL rank per GPU (2, and 4 / 8 when
the box has them) running tests/multigpu_worker.py; skipped on single-GPU boxes."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_classify_bit_identical_across_ranks(dev, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    port = 29600 + (os.getpid() + world) % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multigpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == world
    assert out["two_stage_bit_identical_vs_single_rank"] and out["two_stage_ranks_agree"]
    assert out["fast_mode_bit_identical_vs_rank0_seed"] and out["fast_mode_ranks_agree"]
