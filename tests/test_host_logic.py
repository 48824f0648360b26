"""CPU tier: the shipped library loads, exports its ABI, and its host arithmetic matches the oracle."""
import ctypes
import hashlib
import json
import os
import random
import re

from oracle import cref, pasta
from conftest import GOLDEN, ROOT


def _declared_symbols():
    syms = []
    for name in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not name.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms += re.findall(r"\b(mina_b200_\w+|verify_\w+_ffi)\s*\(", text)
    return sorted(set(syms))


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    syms = _declared_symbols()
    assert "mina_b200_init" in syms and "mina_b200_msm_srs" in syms
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_compute_entry_points_fail_loudly_without_init(native):
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box another test may already have initialised the context
    out = ctypes.create_string_buffer(64)
    rc = native.load().mina_b200_msm_srs(1, 1, 4, b"\0" * 128, out)
    assert rc != 0
    assert b"not initialised" in native.load().mina_b200_last_error()


def test_host_field_ops_match_python(native):
    random.seed(3)
    for fid, m in ((0, pasta.P), (1, pasta.Q)):
        a = [random.randrange(m) for _ in range(300)] + [0, 1, m - 1, 2]
        b = [random.randrange(m) for _ in range(300)] + [m - 1, m - 1, m - 1, (m + 1) // 2]
        A, B = cref.ints_to_bytes(a), cref.ints_to_bytes(b)
        assert native.host_field_op(fid, 0, A, B) == cref.ints_to_bytes([x * y % m for x, y in zip(a, b)])
        assert native.host_field_op(fid, 1, A, B) == cref.ints_to_bytes([(x + y) % m for x, y in zip(a, b)])
        assert native.host_field_op(fid, 2, A, B) == cref.ints_to_bytes([(x - y) % m for x, y in zip(a, b)])
        assert native.host_field_op(fid, 4, A) == cref.ints_to_bytes([x * x % m for x in a])
        assert native.host_field_op(fid, 3, A) == cref.ints_to_bytes([pow(x, -1, m) if x else 0 for x in a])
        sq = [x * x % m for x in a]
        assert native.host_field_op(fid, 5, cref.ints_to_bytes(sq)) == cref.ints_to_bytes(
            [pasta.sqrt(x, m) for x in sq]
        )


def test_host_blake2b(native):
    for msg in (b"", b"abc", b"x" * 128, b"y" * 129, b"z" * 1000):
        assert native.host_blake2b512(msg) == hashlib.blake2b(msg).digest()


def test_host_srs_derivation_pinned(native):
    pins = json.load(open(os.path.join(GOLDEN, "srs_sha256.json")))
    for cid, name in ((0, "pallas"), (1, "vesta")):
        g, h = native.host_srs_derive(cid, 0, 1024, True)
        assert hashlib.sha256(g).hexdigest() == pins[name]["sha256_g_1024"]
        assert h.hex() == pins[name]["h"]
        g2, _ = cref.srs_derive(cid, 5000, 16, False)
        assert native.host_srs_derive(cid, 5000, 16) == g2
