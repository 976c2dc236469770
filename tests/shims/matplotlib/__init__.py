"""Test shim: import-only presence of `matplotlib` (tools/tools.py:9,12 of the reference; its setup.py files
import matplotlib.style, which the oracle build does not use)."""
