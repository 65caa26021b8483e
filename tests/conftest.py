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


_PARITY_LOG = {}


@pytest.fixture(scope="session")
def parity_log():
    """Measured parity errors, recorded by the GPU tests and written to gpurun_out/r02_parity.json at the end
    of the session (committed as profiles/r02_parity.json)."""
    return _PARITY_LOG


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY_LOG:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "r02_parity.json")
        old = {}
        if os.path.exists(path):
            try:
                old = json.load(open(path))
            except Exception:
                old = {}
        old.update(_PARITY_LOG)
        with open(path, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)
    except OSError:
        pass
