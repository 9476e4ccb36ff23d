import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cache():
    """One HBM-resident LiquidCache on cuda:0 for the whole GPU test session."""
    from liquid_cache_b200 import LiquidCacheBuilder

    c = LiquidCacheBuilder.new().build()
    yield c
    c.close()
