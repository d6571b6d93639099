"""pytest configuration: registers the `gpu` marker and shared fixtures.

CPU suite:  python -m pytest tests/ -x -q -m "not gpu"   (oracle vs golden vectors, host logic, C-ABI symbols)
GPU suite:  python -m pytest tests/ -x -q -m gpu         (parity of the CUDA path through the C ABI)
"""

import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    g = json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
    g.update(json.loads((ROOT / "tests" / "golden" / "golden_r2.json").read_text()))  # round-2 additions (make_golden_r2.py)
    g.update(json.loads((ROOT / "tests" / "golden" / "golden_r2_wrappers.json").read_text()))  # make_golden_r2_wrappers.py
    return g


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.build()
    return oracle
