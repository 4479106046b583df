import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


from dpa_testutil import GOLDEN_CASES, load_golden  # noqa: E402


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
