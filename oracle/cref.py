"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- ctypes loader for oracle/pasta_ref.c.

Builds oracle/build/libpasta_oracle.so with gcc on first use (or via __graft_entry__.build()).
Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pasta_ref.c")
_OUT = os.path.join(_HERE, "build", "libpasta_oracle.so")
_lib = None

FP, FQ = 0, 1  # field ids: coordinates of Pallas live in Fp, of Vesta in Fq


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O3", "-march=native", "-funroll-loops", "-shared", "-fPIC", "-pthread", "-o", _OUT, _SRC]
        )
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_msm.restype = ctypes.c_int
        _lib.oracle_msm.argtypes = [
            ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p,
            ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
        ]
        _lib.oracle_decompress.restype = ctypes.c_int
        _lib.oracle_ark_window_bits.restype = ctypes.c_int
        _lib.oracle_ark_window_bits.argtypes = [ctypes.c_size_t]
        _lib.oracle_modmul_ns.restype = ctypes.c_double
    return _lib


def msm(field_id: int, scalars: bytes, points: bytes, nthreads: int = 1, c: int = 0):
    """Return (affine 64-byte result, is_identity)."""
    n = len(scalars) // 32
    assert len(points) >= 64 * n
    out = ctypes.create_string_buffer(64)
    inf = lib().oracle_msm(field_id, n, scalars, points, out, nthreads, c)
    return out.raw, bool(inf)


def decompress(field_id: int, comp: bytes) -> bytes:
    n = len(comp) // 33
    out = ctypes.create_string_buffer(64 * n)
    rc = lib().oracle_decompress(field_id, ctypes.c_size_t(n), comp, out)
    if rc != 0:
        raise ValueError("decompression failed")
    return out.raw


def srs_derive(field_id: int, start: int, count: int, want_h: bool = False):
    out = ctypes.create_string_buffer(64 * count)
    h = ctypes.create_string_buffer(64)
    lib().oracle_srs_derive(field_id, ctypes.c_uint32(start), ctypes.c_uint32(count), out, int(want_h), h)
    return out.raw, (h.raw if want_h else None)


def to_group(field_id: int, t: int) -> bytes:
    out = ctypes.create_string_buffer(64)
    lib().oracle_to_group(field_id, t.to_bytes(32, "little"), out)
    return out.raw


def endo_to_field(scalar_field_id: int, pre: bytes, endo: int) -> bytes:
    n = len(pre) // 16
    out = ctypes.create_string_buffer(32 * n)
    lib().oracle_endo_to_field(scalar_field_id, ctypes.c_size_t(n), pre, endo.to_bytes(32, "little"), out)
    return out.raw


def bpoly_coeffs(scalar_field_id: int, chals: bytes) -> bytes:
    k = len(chals) // 32
    out = ctypes.create_string_buffer(32 << k)
    lib().oracle_bpoly_coeffs(scalar_field_id, k, chals, out)
    return out.raw


def blake2b512(data: bytes) -> bytes:
    out = ctypes.create_string_buffer(64)
    lib().oracle_blake2b512(data, ctypes.c_size_t(len(data)), out)
    return out.raw


def poseidon_permute(field_id: int, params: bytes, states: bytes) -> bytes:
    n = len(states) // 96
    buf = ctypes.create_string_buffer(states, len(states))
    lib().oracle_poseidon_permute(field_id, params, ctypes.c_size_t(n), buf)
    return buf.raw


def modmul_ns(field_id: int = 0, iters: int = 2_000_000) -> float:
    """ns per Montgomery multiplication of the port on one core (dependent chain)."""
    return float(lib().oracle_modmul_ns(field_id, iters))


def ints_to_bytes(xs) -> bytes:
    return b"".join(int(x).to_bytes(32, "little") for x in xs)


def points_to_bytes(pts) -> bytes:
    return b"".join(
        (b"\0" * 64 if p is None else int(p[0]).to_bytes(32, "little") + int(p[1]).to_bytes(32, "little"))
        for p in pts
    )


def bytes_to_point(b: bytes):
    x = int.from_bytes(b[:32], "little")
    y = int.from_bytes(b[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)
