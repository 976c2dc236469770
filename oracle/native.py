"""ctypes/numpy face of oracle/scanerf_oracle.c (TEST INFRASTRUCTURE ONLY).

Every function takes and returns numpy arrays; sizes are small enough for the
C oracle to finish in seconds.  See scanerf_oracle.c for the reference
file:line each function restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libscanerf_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "scanerf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(0)


def hash_encode_fwd(points, table, res, corner=None, size=None, want_idx=False):
    points, table, res = _f32(points), _f32(table), _i32(res)
    B, (L, T) = points.shape[0], table.shape[:2]
    out = np.zeros((B, L, 2), np.float32)
    idx = np.zeros((B, L, 8), np.uint32) if want_idx else None
    c = _f32(corner) if corner is not None else None
    s = _f32(size) if size is not None else None
    lib().oracle_hash_encode_fwd(_p(points), _p(table), _p(res), _p(c), _p(s), _p(out), _p(idx),
                                 ctypes.c_int(B), ctypes.c_int(L), ctypes.c_int(T))
    return (out, idx) if want_idx else out


def hash_encode_bwd(points, grad_in, table, res, corner=None, size=None):
    points, grad_in, table, res = _f32(points), _f32(grad_in), _f32(table), _i32(res)
    B, (L, T) = points.shape[0], table.shape[:2]
    gp = np.zeros((B, 3), np.float32)
    gt = np.zeros_like(table)
    c = _f32(corner) if corner is not None else None
    s = _f32(size) if size is not None else None
    lib().oracle_hash_encode_bwd(_p(points), _p(grad_in), _p(table), _p(res), _p(c), _p(s), _p(gp), _p(gt),
                                 ctypes.c_int(B), ctypes.c_int(L), ctypes.c_int(T))
    return gp, gt
