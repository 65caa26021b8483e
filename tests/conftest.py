import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def ref_mods(oracle):
    """The reference's own code compiled by oracle/build_ref.py (may be partly empty)."""
    try:
        from oracle import build_ref
        build_ref.build()          # no-op when /root/reference is absent (GPU box)
    except Exception:
        pass
    return oracle.ref_modules()
