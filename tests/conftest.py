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


_REF_CACHE = {}


def ref_module(name):
    """Load one of the reference extension modules rebuilt by oracle/build_ref.py BY PATH
    (oracle/_ref/<name>.so) or return None when it was not built.  Never goes through sys.path /
    sys.modules: after `install()` the names `fastMesh`, `hashgrid`, `cuda` resolve to the drop-in
    packages, and a by-name import would hand the test the product instead of the reference."""
    d = os.path.join(ROOT, "oracle", "_ref")
    so = os.path.join(d, name + ".so")
    if not os.path.exists(so):
        return None
    if name in _REF_CACHE:
        return _REF_CACHE[name]
    import importlib.machinery
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    loader = importlib.machinery.ExtensionFileLoader(name, so)
    spec = importlib.util.spec_from_file_location(name, so, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    assert mod.__file__.endswith(".so") and os.path.dirname(os.path.abspath(mod.__file__)) == d, mod.__file__
    _REF_CACHE[name] = mod
    return mod
