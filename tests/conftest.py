import os
import sys

import pytest

# small upload markers so that the streaming scan (sgz_db_finalize_async) sees many ranges on the small test DBs;
# read once by the library, before the first database is built
os.environ.setdefault("SGZ_CHUNK_FRAMES", "3000")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One engine context on cuda:0.  GPU tests must FAIL, not skip, when the native path is missing."""
    from strugatzki_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()
