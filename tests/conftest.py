import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The reference's OpenMP loops (schedule(dynamic,1)) collapse when a many-core host is oversubscribed
# (bench.py's thread scan on a 128-thread box: 24 threads 150 ms per 256^3 cycle, 128 threads 20 s), and
# every test that runs oracle/_ref inherits this process's environment.
os.environ.setdefault("OMP_NUM_THREADS", str(min(24, os.cpu_count() or 1)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """make sure the in-tree shared libraries exist (they travel with the snapshot)"""
    import __graft_entry__ as g
    g.build()
    return True
