import gzip
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def fusion_golden():
    """Vectors produced by the reference's own code (tests/golden/make_golden.py)."""
    with gzip.open(ROOT / "tests" / "golden" / "fusion_golden.json.gz", "rb") as fh:
        return json.loads(fh.read().decode())


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from triple_hybrid_rag_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
