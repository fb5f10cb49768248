import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    import oracle

    return oracle.load()


@pytest.fixture(scope="session")
def kats():
    return json.loads((ROOT / "tests" / "golden" / "reference_kats.json").read_text())


@pytest.fixture(scope="session")
def pb():
    """The product package, with the CUDA library built in-tree."""
    import build_native

    build_native.build()
    import pbrt_b200

    return pbrt_b200


@pytest.fixture(scope="session")
def gpu(pb):
    pb.init(0)
    return pb
