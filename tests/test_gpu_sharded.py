"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): launches benchmarks/sharded_check.py with
one rank per GPU over NCCL and expects every shard to equal the oracle's single-filter bit array."""

import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_sharded_bloom_and_cms_match_oracle():
    import pyprobables_b200 as pb

    n = pb.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "benchmarks" / "sharded_check.py"), "--keys", "500000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert "SHARDED PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
