import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def ref(orc):
    r = orc.ref()
    if r is None:
        pytest.skip("oracle/_ref/libref_host.so not present (built only where /root/reference exists)")
    return r


@pytest.fixture(scope="session")
def api():
    from radiosity_b200 import api as a
    a.host_lib()
    return a


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)
