"""Test shim: import-only presence of `imageio` for the reference's tools/utils.py:7 and tools/tools.py:8.
Nothing on the tested path calls it."""


def _absent(*a, **k):
    raise RuntimeError("imageio is a test shim: not available in this image")


imread = imwrite = imsave = mimsave = mimwrite = get_writer = get_reader = _absent
