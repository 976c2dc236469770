import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_NAME = "scanerf-scalable-bundle-adjusting-neural-radiance-fields-for-large-scale-scene-rendering_b200"
PKG_DIR = os.path.join(ROOT, PKG_NAME)
GOLDEN = os.path.join(ROOT, "tests", "golden")

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """Import the (hyphen-named) product package and install its drop-in packages
    (hashgrid / cuda / fastMesh / vdbAdam) at the front of sys.path."""
    pkg = importlib.import_module(PKG_NAME)
    pkg.install()
    return pkg


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


def ref_module(name):
    """One of the reference extension modules rebuilt into oracle/_ref/, loaded BY PATH (oracle.ref_module), or None."""
    import oracle
    return oracle.ref_module(name)
