import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# staging copy of the reference's own tests (tools/stage_reference_tests.sh): only run through
# tests/test_reference_suite_gpu.py, in their own pytest process
collect_ignore = ["_reference_tests"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    import oracle
    oracle.build()
