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


def _write_srs_file(path, curve_id, points, h, array32=False):
    """Producer side of srs/*.srs: MessagePack [[bin33 x n], bin33], compressed points."""
    m = pasta.P if curve_id == 0 else pasta.Q

    def rec(pt):
        x, y = pt
        return b"\xc4\x21" + x.to_bytes(32, "little") + (b"\x80" if y > (m - 1) // 2 else b"\x00")

    n = len(points)
    hdr = b"\x92" + (b"\xdd" + n.to_bytes(4, "big") if array32 or n > 0xFFFF else b"\xdc" + n.to_bytes(2, "big"))
    open(path, "wb").write(hdr + b"".join(rec(p) for p in points) + rec(h))


def test_srs_file_loader_round_trips_the_committed_format(native, tmp_path):
    import pytest

    for cid in (0, 1):
        g, h = native.host_srs_derive(cid, 0, 700, True)
        pts = [cref.bytes_to_point(g[64 * i : 64 * i + 64]) for i in range(700)]
        for a32 in (False, True):
            p = str(tmp_path / ("c%d_%d.srs" % (cid, a32)))
            _write_srs_file(p, cid, pts, cref.bytes_to_point(h), a32)
            got, got_h = native.host_srs_load_file(cid, p, 700)
            assert got == g and got_h == h
            got, _ = native.host_srs_load_file(cid, p, 300)  # a prefix, like Pallas' 2^15 of 2^16
            assert got == g[: 64 * 300]
            with pytest.raises(native.MinaB200Error):
                native.host_srs_load_file(cid, p, 701)
        raw = bytearray(open(p, "rb").read())
        raw[10] ^= 1  # a corrupted x is (almost surely) not on the curve, or decompresses to another point
        open(p, "wb").write(bytes(raw))
        try:
            got, _ = native.host_srs_load_file(cid, p, 700)
            assert got != g
        except native.MinaB200Error:
            pass
        open(p, "wb").write(bytes(raw[:1000]))
        with pytest.raises(native.MinaB200Error):
            native.host_srs_load_file(cid, p, 700)


def test_srs_file_loader_on_the_reference_files_when_present(native):
    """Here (not on the GPU box) the reference tree is mounted: its committed files must load to the pinned SRS."""
    import pytest

    pins = json.load(open(os.path.join(GOLDEN, "srs_sha256.json")))
    for cid, name, depth in ((1, "vesta", 65536), (0, "pallas", 32768)):
        path = "/root/reference/srs/%s.srs" % name
        if not os.path.exists(path):
            pytest.skip("reference tree not mounted")
        g, h = native.host_srs_load_file(cid, path, depth)
        assert hashlib.sha256(g).hexdigest() == pins[name]["sha256_g_%d" % depth] and h.hex() == pins[name]["h"]


# ---- group testing planner (csrc/group_testing.hpp) against a simulated device -------------------------------------
def test_group_testing_planner_resolves_every_pattern(native):
    """The host logic that drives the batched accumulator / IPA checks: with the device replaced by the truth (a group
    reports "all good", "exactly one bad: index j" or "two or more bad"), every index must come out with the right bit,
    for every batch size and corruption pattern, within a bounded number of levels and MSMs.  The simulator also checks
    the structure the device path relies on (children tile their parent, sibling ranges, derived child closes it)."""
    rng = random.Random(0x4D494E41)
    worst = {}
    for m in [1, 2, 3, 4, 5, 7, 8, 63, 64, 65, 100, 127, 128, 129, 130, 200, 333, 512, 1000, 1024, 1500, 2048]:
        patterns = [set(), {0}, {m - 1}, {m // 2}, set(range(m)), set(range(0, m, 2)), set(range(min(m, 64))), {0, m - 1}]
        patterns += [set(rng.sample(range(m), min(m, k))) for k in (2, 3, 10, 30) for _ in range(3)]
        patterns += [{i for i in range(m) if 60 <= i % 64 or i % 64 < 3}]  # clusters across every slice boundary
        for bad in patterns:
            flags = bytes(1 if i in bad else 0 for i in range(m))
            ok, levels, msms = native.host_group_testing_sim(flags)
            assert ok == bytes(0 if i in bad else 1 for i in range(m)), (m, sorted(bad)[:8])
            k = len(bad)
            if k == 0:
                assert (levels, msms) == (1, 1)
            elif k == 1 and m > 1:
                assert (levels, msms) == (2, 3)  # one failing level 0, one locator level over the whole batch
            # never worse than testing every item on its own twice over, and logarithmic depth
            assert msms <= 1 + 2 * max(m, 2) and levels <= 2 + 2 * max(1, m).bit_length()
            worst[(m, k)] = max(worst.get((m, k), 0), msms)
    assert worst[(1024, 10)] <= 60  # the bench's shape costs ~28 on its seeded pattern
