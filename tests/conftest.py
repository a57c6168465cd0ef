import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(params=["components", "fallback"])
def coupled_kernel(request):
    """Runs a GPU test once with each of the two implementations of the coupled pass (sapio_set_coupled_pack_from): connected
    components one per CTA (what every ordinary tick runs), and the cooperative fallback kernel that only takes over when a
    component exceeds the largest CTA tier -- which test worlds never do on their own."""
    from sapiogenesis_b200._lib import load
    lib = load()
    old = lib.sapio_set_coupled_pack_from(-1 if request.param == "components" else 2 ** 31 - 1)
    yield request.param
    lib.sapio_set_coupled_pack_from(old)
