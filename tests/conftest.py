import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run with -m gpu on the B200 box)")
    # The product library and the CPU oracle are built in-tree; (re)build them when sources are newer.
    from cudecomp_b200.build import build_library
    build_library()
    from oracle import oracle as orc
    orc.build()


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "api_golden.json")) as f:
        return json.load(f)
