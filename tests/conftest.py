import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name: str) -> bytes:
    return open(os.path.join(GOLDEN, name), "rb").read()


@pytest.fixture(scope="session")
def state_proof():
    from oracle import wire

    return wire.decode_state_proof(golden("mina_state.proof"))


@pytest.fixture(scope="session")
def native():
    """The native library, built if needed (CPU tests only use the host-only hooks)."""
    import __graft_entry__ as entry

    entry.build()
    import mina_bridge_b200 as mb

    mb.load()
    return mb


@pytest.fixture(scope="session")
def gpu(native):
    """Initialised device context.  Fails loudly (never falls back) if there is no GPU."""
    native.init(0)
    return native
