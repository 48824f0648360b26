"""ctypes binding of include/mina_b200.h -- the stub a Python caller (or the parity tests) uses.

This plays the role of the reference's cgo shims (AL/operator/mina/mina.go:27-32,
AL/operator/mina_account/mina_account.go:27-32): pass plain buffers and lengths, get a bool / bytes.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libmina_b200.so")
_lib = None


class MinaB200Error(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


def load():
    """Load the native library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise MinaB200Error(
                "native library missing: %s (run `python -c 'import __graft_entry__ as g; g.build()'`)" % _LIB_PATH
            )
        lib = ctypes.CDLL(_LIB_PATH)
        lib.mina_b200_last_error.restype = ctypes.c_char_p
        lib.mina_b200_launch_count.restype = ctypes.c_uint64
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        raise MinaB200Error(load().mina_b200_last_error().decode() or "error %d" % rc)


def init(device: int = 0, cache_dir: str | None = None):
    if cache_dir is None:
        cache_dir = os.path.join(_HERE, "data")
    _check(load().mina_b200_init(int(device), cache_dir.encode()))


def shutdown():
    load().mina_b200_shutdown()


def launch_count() -> int:
    return int(load().mina_b200_launch_count())


def srs_points(curve: int, first: int, count: int, want_h: bool = False):
    out = ctypes.create_string_buffer(64 * count)
    h = ctypes.create_string_buffer(64)
    _check(load().mina_b200_srs_points(curve, ctypes.c_uint32(first), ctypes.c_uint32(count), out, h if want_h else None))
    return (out.raw, h.raw) if want_h else out.raw


def msm_srs(curve: int, scalars: bytes, n: int) -> list[bytes]:
    """nmsm MSMs over the resident SRS prefix g[0..n); returns canonical affine results."""
    assert n == 0 or len(scalars) % (32 * n) == 0
    nmsm = len(scalars) // (32 * n) if n else 0
    out = ctypes.create_string_buffer(64 * max(nmsm, 1))
    _check(load().mina_b200_msm_srs(curve, ctypes.c_uint32(nmsm), ctypes.c_uint32(n), scalars, out))
    return [out.raw[64 * i : 64 * i + 64] for i in range(nmsm)]


def msm(curve: int, scalars: bytes, points: bytes, window_bits: int = 0) -> bytes:
    n = len(scalars) // 32
    assert len(points) == 64 * n
    out = ctypes.create_string_buffer(64)
    _check(load().mina_b200_msm(curve, ctypes.c_uint32(n), scalars, points, int(window_bits), out))
    return out.raw


def msm_srs_device(curve: int, nmsm: int, n: int, d_scalars: int, d_out: int, stream: int, want_ms: bool = False):
    ms = ctypes.c_float(0.0)
    _check(
        load().mina_b200_msm_srs_device(
            curve, ctypes.c_uint32(nmsm), ctypes.c_uint32(n), ctypes.c_void_p(d_scalars), ctypes.c_void_p(d_out),
            ctypes.c_void_p(stream), ctypes.byref(ms) if want_ms else None,
        )
    )
    return ms.value if want_ms else None


def msm_configure(curve: int, window_bits: int, precompute: bool = True, leaf: int = 8):
    _check(load().mina_b200_msm_configure(curve, window_bits, int(precompute), leaf))


def field_op(field: int, op: int, a: bytes, b: bytes | None = None) -> bytes:
    n = len(a) // 32
    out = ctypes.create_string_buffer(32 * max(n, 1))
    _check(load().mina_b200_field_op(field, op, ctypes.c_uint32(n), a, b, out))
    return out.raw[: 32 * n]


def point_add(curve: int, a: bytes, b: bytes) -> bytes:
    n = len(a) // 64
    out = ctypes.create_string_buffer(64 * max(n, 1))
    _check(load().mina_b200_point_add(curve, ctypes.c_uint32(n), a, b, out))
    return out.raw[: 64 * n]


def host_field_op(field: int, op: int, a: bytes, b: bytes | None = None) -> bytes:
    n = len(a) // 32
    out = ctypes.create_string_buffer(32 * max(n, 1))
    rc = load().mina_b200_host_field_op(field, op, ctypes.c_uint32(n), a, b, out)
    if rc != 0:
        raise MinaB200Error("host_field_op failed: %d" % rc)
    return out.raw[: 32 * n]


def host_srs_derive(curve: int, first: int, count: int, want_h: bool = False):
    out = ctypes.create_string_buffer(64 * max(count, 1))
    h = ctypes.create_string_buffer(64)
    rc = load().mina_b200_host_srs_derive(curve, ctypes.c_uint32(first), ctypes.c_uint32(count), out, h if want_h else None)
    if rc != 0:
        raise MinaB200Error("host_srs_derive failed: %d" % rc)
    return (out.raw[: 64 * count], h.raw) if want_h else out.raw[: 64 * count]


def host_blake2b512(data: bytes) -> bytes:
    out = ctypes.create_string_buffer(64)
    load().mina_b200_host_blake2b512(data, ctypes.c_size_t(len(data)), out)
    return out.raw
