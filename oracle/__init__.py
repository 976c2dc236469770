"""CPU oracle for the scanerf_b200 hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product package never does.
"""
import os

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_REF_CACHE = {}


def ref_module(name):
    """Load one of the UNMODIFIED reference extension modules rebuilt by oracle/build_ref.py BY PATH
    (oracle/_ref/<name>.so: CUDA_EXT, HASHGRID_EMBED, HASHGRID, fastMesh) or return None when it was not built.
    Never goes through sys.path / sys.modules: once the drop-in is installed the names `fastMesh`, `hashgrid`, `cuda`
    resolve to the product's packages, and a by-name import would hand the checker the product instead of the reference."""
    so = os.path.join(_REF_DIR, name + ".so")
    if not os.path.exists(so):
        return None
    if name not in _REF_CACHE:
        import importlib.machinery
        import importlib.util
        import torch  # noqa: F401  (libtorch must be loaded first)
        loader = importlib.machinery.ExtensionFileLoader(name, so)
        spec = importlib.util.spec_from_file_location(name, so, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        assert mod.__file__.endswith(".so") and os.path.dirname(os.path.abspath(mod.__file__)) == _REF_DIR, mod.__file__
        _REF_CACHE[name] = mod
    return _REF_CACHE[name]
