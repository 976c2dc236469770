"""Test shim: import-only presence of `plyfile` for the reference's tools/tools.py:4."""


class PlyData:
    def __init__(self, *a, **k):
        raise RuntimeError("plyfile is a test shim: not available in this image")

    @staticmethod
    def read(*a, **k):
        raise RuntimeError("plyfile is a test shim: not available in this image")


class PlyElement:
    @staticmethod
    def describe(*a, **k):
        raise RuntimeError("plyfile is a test shim: not available in this image")
