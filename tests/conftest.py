import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def readme_mcp():
    from mcp_b200 import problems
    return problems.readme_qp()


@pytest.fixture(scope="session")
def lane_game():
    from mcp_b200 import problems
    return problems.lane_change_game()


@pytest.fixture(scope="session")
def clamp_game():
    from mcp_b200 import problems
    return problems.clamp_game()
