import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "cuda-efficient-features_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import efo
    efo.build()
    o = efo.Oracle()
    o.set_threads(min(o.max_threads(), 16))
    return o


@pytest.fixture(scope="session")
def reference():
    import efo
    if not efo.Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return efo.Reference()
