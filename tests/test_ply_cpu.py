"""CPU: the host-side PLY reader behind snrf_voxelize_mesh_host / snrf_mesh_create treats the file as untrusted
(ADVICE r1): a header that lies about counts, an unknown type, a truncated body or an out-of-range face index come
back as an error code + message through the C ABI -- nothing throws across `extern "C"`."""
import ctypes
import os
import struct

import pytest

from conftest import PKG_DIR

HEADER = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
          "element face %d\nproperty list %s int vertex_indices\nend_header\n")
TRI = struct.pack("9f", 0, 0, 0, 1, 0, 0, 0, 1, 0)


def _voxelize(lib, path):
    l2, c, s = (ctypes.c_int * 3)(2, 2, 2), (ctypes.c_float * 3)(0, 0, 0), (ctypes.c_float * 3)(1, 1, 1)
    vis, out = (ctypes.c_ubyte * 64)(), (ctypes.c_ubyte * 64)()
    rc = lib.snrf_voxelize_mesh_host(l2, c, s, path.encode(), vis, 1, out)
    return rc, (lib.snrf_last_error() or b"").decode(), bytes(vis)


@pytest.fixture(scope="module")
def lib():
    cdll = ctypes.CDLL(os.path.join(PKG_DIR, "lib", "libscanerf_b200.so"))
    cdll.snrf_last_error.restype = ctypes.c_char_p
    return cdll


@pytest.mark.parametrize("name,header,body,needle", [
    ("huge", HEADER % (2 ** 40, 1, "uchar"), b"", "exceeds the file size"),
    ("badtype", HEADER % (3, 1, "quux"), TRI, "unknown PLY list count type"),
    ("truncated", HEADER % (3, 1, "uchar"), TRI + b"\x03" + struct.pack("i", 0), "unexpected end of file"),
    ("index", HEADER % (3, 1, "uchar"), TRI + b"\x03" + struct.pack("3i", 0, 1, 7), "face index out of range"),
    ("quad", HEADER % (3, 1, "uchar"), TRI + b"\x04" + struct.pack("4i", 0, 1, 2, 0), "only triangle faces"),
])
def test_malformed_ply_is_an_error_not_a_crash(lib, tmp_path, name, header, body, needle):
    p = tmp_path / (name + ".ply")
    p.write_bytes(header.encode() + body)
    rc, msg, _ = _voxelize(lib, str(p))
    assert rc != 0 and needle in msg, (rc, msg)


def test_well_formed_ply_voxelizes(lib, tmp_path):
    p = tmp_path / "ok.ply"
    p.write_bytes((HEADER % (3, 1, "uchar")).encode() + TRI + b"\x03" + struct.pack("3i", 0, 1, 2))
    rc, _, vis = _voxelize(lib, str(p))
    assert rc == 0 and any(vis)
