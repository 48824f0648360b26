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


def lagrange_commitments(curve: int, log_n: int, first: int, count: int) -> list[bytes]:
    """Commitments of the Lagrange polynomials L_first.. of the 2^log_n domain over the resident SRS (canonical affine)."""
    out = ctypes.create_string_buffer(64 * max(count, 1))
    _check(load().mina_b200_lagrange_commitments(curve, ctypes.c_uint32(log_n), ctypes.c_uint32(first), ctypes.c_uint32(count), out))
    return [out.raw[64 * i : 64 * i + 64] for i in range(count)]


def public_commitments(curve: int, log_n: int, n_pub: int, pub: bytes) -> list[bytes]:
    """-sum_i pub_i L_i + h for each vector of n_pub public inputs (canonical scalars)."""
    nproofs = len(pub) // (32 * n_pub)
    out = ctypes.create_string_buffer(64 * max(nproofs, 1))
    _check(load().mina_b200_public_commitments(curve, ctypes.c_uint32(log_n), ctypes.c_uint32(n_pub), ctypes.c_uint32(nproofs), pub, out))
    return [out.raw[64 * i : 64 * i + 64] for i in range(nproofs)]


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


# ---- verifier boundary (include/mina_verifier.h, include/mina_account_verifier.h) ---------------------
MAX_STATE_PROOF_SIZE = 48 * 1024
MAX_ACCOUNT_PROOF_SIZE = 16 * 1024
MAX_PUB_INPUT_SIZE = 6 * 1024
MODE_PER_PROOF, MODE_RLC = 0, 1

STAGES = {
    "lengths": 1 << 0, "decode_proof": 1 << 1, "decode_pub": 1 << 2, "pub_structure": 1 << 3, "pub_hashes": 1 << 4,
    "consensus": 1 << 5, "accumulator": 1 << 6, "step_accumulators": 1 << 7, "kimchi": 1 << 8,
    "account_abi": 1 << 9, "account_leaf": 1 << 10, "merkle": 1 << 11, "internal_error": 1 << 31,
}


class StageReport(ctypes.Structure):
    _fields_ = [("passed", ctypes.c_uint32), ("failed", ctypes.c_uint32), ("unavailable", ctypes.c_uint32)]

    def names(self, mask):
        return sorted(k for k, v in STAGES.items() if mask & v)

    def as_dict(self):
        return {"passed": self.names(self.passed), "failed": self.names(self.failed), "unavailable": self.names(self.unavailable)}


def _padded(data: bytes, size: int):
    """The fixed-size zero-padded array the Go operator / Rust batcher pass (operator.go:534-541)."""
    buf = (ctypes.c_ubyte * max(size, len(data)))()
    ctypes.memmove(buf, data, len(data))
    return buf


def verify_mina_state(proof: bytes, pub: bytes, proof_len: int | None = None, pub_len: int | None = None) -> bool:
    """Mirror of AL/operator/mina/mina.go:27-32 VerifyMinaState."""
    lib = load()
    lib.verify_mina_state_ffi.restype = ctypes.c_bool
    lib.verify_mina_state_ffi.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    pb, ib = _padded(proof, MAX_STATE_PROOF_SIZE), _padded(pub, MAX_PUB_INPUT_SIZE)
    return bool(lib.verify_mina_state_ffi(pb, len(proof) if proof_len is None else proof_len, ib, len(pub) if pub_len is None else pub_len))


def verify_account_inclusion(proof: bytes, pub: bytes, proof_len: int | None = None, pub_len: int | None = None) -> bool:
    """Mirror of AL/operator/mina_account/mina_account.go:27-32 VerifyAccountInclusion."""
    lib = load()
    lib.verify_account_inclusion_ffi.restype = ctypes.c_bool
    lib.verify_account_inclusion_ffi.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    pb, ib = _padded(proof, MAX_ACCOUNT_PROOF_SIZE), _padded(pub, MAX_PUB_INPUT_SIZE)
    return bool(lib.verify_account_inclusion_ffi(pb, len(proof) if proof_len is None else proof_len, ib, len(pub) if pub_len is None else pub_len))


def last_stages() -> StageReport:
    rep = StageReport()
    load().mina_b200_last_stages(ctypes.byref(rep))
    return rep


class Batch:
    """Pointer / length arrays over a list of byte strings (kept alive with the object)."""

    def __init__(self, items):
        self.n = len(items)
        self.keep = [ctypes.create_string_buffer(bytes(x), max(len(x), 1)) for x in items]
        self.ptrs = (ctypes.c_void_p * max(self.n, 1))(*[ctypes.addressof(b) for b in self.keep])
        self.lens = (ctypes.c_size_t * max(self.n, 1))(*[len(x) for x in items])


def verify_state_stages(proofs, pubs, mode: int = MODE_RLC):
    """Batch verifier with per-proof stage reports: returns (accept bits, [StageReport])."""
    p, q = (proofs if isinstance(proofs, Batch) else Batch(proofs)), (pubs if isinstance(pubs, Batch) else Batch(pubs))
    reports = (StageReport * max(p.n, 1))()
    accept = (ctypes.c_uint8 * max(p.n, 1))()
    lib = load()
    lib.mina_b200_verify_state_stages.argtypes = [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    _check(lib.mina_b200_verify_state_stages(p.n, p.ptrs, p.lens, q.ptrs, q.lens, mode, reports, accept))
    return list(accept[: p.n]), list(reports[: p.n])


def verify_state_stage_masks(proofs, pubs, mode: int = MODE_RLC):
    """Same call as verify_state_stages, returning the reports as an (n, 3) uint32 numpy array (passed, failed,
    unavailable) without building n Python objects: what a throughput-sensitive caller reads."""
    import numpy as np

    p, q = (proofs if isinstance(proofs, Batch) else Batch(proofs)), (pubs if isinstance(pubs, Batch) else Batch(pubs))
    reports = np.zeros((max(p.n, 1), 3), dtype=np.uint32)
    lib = load()
    lib.mina_b200_verify_state_stages.argtypes = [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    _check(lib.mina_b200_verify_state_stages(p.n, p.ptrs, p.lens, q.ptrs, q.lens, mode, reports.ctypes.data_as(ctypes.c_void_p), None))
    return reports[: p.n]


def verify_state_batch(proofs, pubs):
    p, q = (proofs if isinstance(proofs, Batch) else Batch(proofs)), (pubs if isinstance(pubs, Batch) else Batch(pubs))
    accept = (ctypes.c_uint8 * max(p.n, 1))()
    lib = load()
    lib.verify_mina_state_batch_ffi.argtypes = [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    rc = lib.verify_mina_state_batch_ffi(p.n, p.ptrs, p.lens, q.ptrs, q.lens, accept)
    if rc != 0:
        raise MinaB200Error(load().mina_b200_last_error().decode() or "batch verify failed")
    return list(accept[: p.n])


def verify_account_stages(proofs, pubs):
    p, q = (proofs if isinstance(proofs, Batch) else Batch(proofs)), (pubs if isinstance(pubs, Batch) else Batch(pubs))
    reports = (StageReport * max(p.n, 1))()
    accept = (ctypes.c_uint8 * max(p.n, 1))()
    lib = load()
    lib.mina_b200_verify_account_stages.argtypes = [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_void_p]
    _check(lib.mina_b200_verify_account_stages(p.n, p.ptrs, p.lens, q.ptrs, q.lens, reports, accept))
    return list(accept[: p.n]), list(reports[: p.n])


def accumulator_check(proofs, mode: int = MODE_PER_PROOF):
    """accumulator_check from raw proof bytes: [(wrap_ok, step0_ok, step1_ok)] per proof."""
    p = proofs if isinstance(proofs, Batch) else Batch(proofs)
    ok = (ctypes.c_uint8 * max(3 * p.n, 1))()
    lib = load()
    lib.mina_b200_accumulator_check_batch.argtypes = [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    _check(lib.mina_b200_accumulator_check_batch(p.n, p.ptrs, p.lens, mode, ok))
    return [tuple(ok[3 * i : 3 * i + 3]) for i in range(p.n)]


# ---- K4 / K2 / K5 / K3 hooks ----------------------------------------------------------------------------
def endo_to_field(field: int, pre: bytes) -> bytes:
    n = len(pre) // 16
    out = ctypes.create_string_buffer(32 * max(n, 1))
    _check(load().mina_b200_endo_to_field(field, ctypes.c_uint32(n), pre, out))
    return out.raw[: 32 * n]


def bpoly_coeffs(field: int, chals: bytes, k: int) -> bytes:
    nproofs = len(chals) // (32 * k)
    out = ctypes.create_string_buffer((32 << k) * max(nproofs, 1))
    _check(load().mina_b200_bpoly_coeffs(field, ctypes.c_uint32(nproofs), k, chals, out))
    return out.raw[: (32 << k) * nproofs]


def bpoly_combine(field: int, chals: bytes, r: bytes, k: int) -> bytes:
    nproofs = len(r) // 32
    out = ctypes.create_string_buffer(32 << k)
    _check(load().mina_b200_bpoly_combine(field, ctypes.c_uint32(nproofs), k, chals, r, out))
    return out.raw


def bpoly_eval(field: int, chals: bytes, xs: bytes, k: int, npts: int) -> bytes:
    nproofs = len(chals) // (32 * k)
    out = ctypes.create_string_buffer(32 * max(nproofs * npts, 1))
    _check(load().mina_b200_bpoly_eval(field, ctypes.c_uint32(nproofs), ctypes.c_uint32(npts), k, chals, xs, out))
    return out.raw[: 32 * nproofs * npts]


def poseidon_permute(field: int, table: bytes, states: bytes) -> bytes:
    n = len(states) // 96
    buf = ctypes.create_string_buffer(states, max(len(states), 1))
    _check(load().mina_b200_poseidon_permute(field, table, ctypes.c_uint32(n), buf))
    return buf.raw[: 96 * n]


def merkle_fold(table: bytes, paths, leaves, roots):
    """paths: list of [(tag, sibling_int)]; returns ([ok], [folded root int])."""
    n = len(paths)
    max_depth = max([len(p) for p in paths] + [1])
    depths = (ctypes.c_uint32 * n)(*[len(p) for p in paths])
    tags = bytearray(n * max_depth)
    sib = bytearray(32 * n * max_depth)
    for i, p in enumerate(paths):
        for d, (tag, h) in enumerate(p):
            tags[i * max_depth + d] = tag
            sib[32 * (i * max_depth + d) : 32 * (i * max_depth + d) + 32] = int(h).to_bytes(32, "little")
    lv = b"".join(int(x).to_bytes(32, "little") for x in leaves)
    rt = b"".join(int(x).to_bytes(32, "little") for x in roots)
    ok = ctypes.create_string_buffer(n)
    folded = ctypes.create_string_buffer(32 * n)
    _check(load().mina_b200_merkle_fold(table, ctypes.c_uint32(n), ctypes.c_uint32(max_depth), depths, bytes(tags), bytes(sib), lv, rt, ok, folded))
    return [b for b in ok.raw], [int.from_bytes(folded.raw[32 * i : 32 * i + 32], "little") for i in range(n)]


class IpaBatch(ctypes.Structure):
    _fields_ = [("n", ctypes.c_uint32), ("rounds", ctypes.c_uint32), ("n_comm", ctypes.c_uint32), ("n_points", ctypes.c_uint32),
                ("sponge_mode", ctypes.c_uint32), ("sponge_count", ctypes.c_uint32),
                ("sponge_state96", ctypes.c_char_p), ("cip32", ctypes.c_char_p), ("polyscale32", ctypes.c_char_p),
                ("evalscale32", ctypes.c_char_p), ("z1_32", ctypes.c_char_p), ("z2_32", ctypes.c_char_p),
                ("eval_points32", ctypes.c_char_p), ("delta64", ctypes.c_char_p), ("sg64", ctypes.c_char_p),
                ("commitments64", ctypes.c_char_p), ("lr64", ctypes.c_char_p)]


def ipa_pack(openings, sponge_mode: int, sponge_count: int):
    """Pack a list of opening dicts (ints / (x, y) points: state[3], cip, polyscale, evalscale, z1, z2, elm[], delta, sg,
    commitments[], lr[(L, R)]) into the flat host buffers of mina_b200_ipa_batch.  Returns (struct, keep-alive list)."""
    n = len(openings)
    f32 = lambda x: int(x).to_bytes(32, "little")
    pt = lambda p: (b"\0" * 64 if p is None else f32(p[0]) + f32(p[1]))
    o0 = openings[0]
    b = IpaBatch()
    b.n, b.rounds, b.n_comm, b.n_points = n, len(o0["lr"]), len(o0["commitments"]), len(o0["elm"])
    b.sponge_mode, b.sponge_count = sponge_mode, sponge_count
    keep = []  # keep the byte strings alive as long as the struct

    def field(name, data):
        keep.append(data)
        setattr(b, name, data)

    field("sponge_state96", b"".join(f32(x) for o in openings for x in o["state"]))
    for name, key in (("cip32", "cip"), ("polyscale32", "polyscale"), ("evalscale32", "evalscale"), ("z1_32", "z1"), ("z2_32", "z2")):
        field(name, b"".join(f32(o[key]) for o in openings))
    field("eval_points32", b"".join(f32(x) for o in openings for x in o["elm"]))
    field("delta64", b"".join(pt(o["delta"]) for o in openings))
    field("sg64", b"".join(pt(o["sg"]) for o in openings))
    field("commitments64", b"".join(pt(p) for o in openings for p in o["commitments"]))
    field("lr64", b"".join(pt(l) + pt(r) for o in openings for (l, r) in o["lr"]))
    return b, keep


def ipa_verify_packed(curve: int, table: bytes, packed):
    b = packed[0]
    ok = ctypes.create_string_buffer(b.n)
    lib = load()
    lib.mina_b200_ipa_verify.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p]
    _check(lib.mina_b200_ipa_verify(curve, table, ctypes.byref(b), ok))
    return ok.raw


def ipa_verify(curve: int, table: bytes, openings, sponge_mode: int, sponge_count: int):
    """Batched IPA final check (SRS::verify).  Returns [ok] per opening."""
    if not openings:
        return []
    return [x for x in ipa_verify_packed(curve, table, ipa_pack(openings, sponge_mode, sponge_count))]


def load_ipa_fixture(path: str):
    """tests/golden/ipa_*.json (tools/make_ipa_fixture.py) -> (curve, table bytes, opening dict, sponge mode, count)"""
    import json

    d = json.load(open(path))
    num = lambda h: int(h, 16)
    pt = lambda p: (num(p[0]), num(p[1]))
    opening = {"state": [num(x) for x in d["state"]], "cip": num(d["cip"]), "polyscale": num(d["polyscale"]), "evalscale": num(d["evalscale"]),
               "z1": num(d["z1"]), "z2": num(d["z2"]), "elm": [num(x) for x in d["elm"]], "delta": pt(d["delta"]), "sg": pt(d["sg"]),
               "commitments": [pt(p) for p in d["commitments"]], "lr": [(pt(l), pt(r)) for l, r in d["lr"]]}
    table = b"".join(num(x).to_bytes(32, "little") for x in d["table"])
    return d["curve"], table, opening, d["sponge_mode"], d["sponge_count"]


def poseidon_trusted() -> bool:
    return bool(load().mina_b200_poseidon_trusted())


# ---- host-only hooks --------------------------------------------------------------------------------------
class WireSummary(ctypes.Structure):
    _fields_ = [
        ("consumed", ctypes.c_uint64), ("proof_end", ctypes.c_uint64),
        ("n_step_comms", ctypes.c_uint32), ("n_lr", ctypes.c_uint32), ("merkle_depth", ctypes.c_uint32), ("is_devnet", ctypes.c_uint32),
        ("blockchain_length", ctypes.c_uint32 * 17), ("curr_global_slot", ctypes.c_uint32 * 17),
        ("epoch_count", ctypes.c_uint32 * 17), ("min_window_density", ctypes.c_uint32 * 17),
        ("state_begin", ctypes.c_uint64 * 17), ("state_end", ctypes.c_uint64 * 17),
        ("previous_state_hash", (ctypes.c_uint8 * 32) * 17), ("first_pass_ledger", (ctypes.c_uint8 * 32) * 17),
        ("wrap_sg", ctypes.c_uint8 * 64), ("step_sg", (ctypes.c_uint8 * 64) * 2), ("hash0", ctypes.c_uint8 * 32),
        ("encoded_account_len", ctypes.c_uint64), ("balance", ctypes.c_uint64), ("nonce", ctypes.c_uint32), ("has_zkapp", ctypes.c_uint32),
    ]


def host_decode(kind: int, data: bytes):
    """kind: 0 state proof, 1 state pub, 2 account proof, 3 account pub.  None on a decode error."""
    s = WireSummary()
    lib = load()
    lib.mina_b200_host_decode.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
    rc = lib.mina_b200_host_decode(kind, data, len(data), ctypes.byref(s))
    return s if rc == 0 else None


def _host_bytes_out(fn, *args, cap=1 << 17):
    buf = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t(cap)
    rc = fn(*args, buf, ctypes.byref(n))
    return (buf.raw[:n.value] if rc == 0 else None), rc


def host_group_testing_sim(bad: bytes):
    """The group-testing planner against a simulated device.  bad: one byte per item (non-zero = bad).
    Returns (ok bytes, levels, msms)."""
    m = len(bad)
    ok = ctypes.create_string_buffer(max(m, 1))
    lv, ms = ctypes.c_uint32(0), ctypes.c_uint32(0)
    _check(load().mina_b200_host_group_testing_sim(ctypes.c_uint32(m), bad, ok, ctypes.byref(lv), ctypes.byref(ms)))
    return ok.raw[:m], lv.value, ms.value


def host_reencode(kind: int, data: bytes):
    """decode + encode through the C++ wire writers (csrc/wire_write.hpp).  None on a decode error."""
    lib = load()
    lib.mina_b200_host_reencode.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    return _host_bytes_out(lib.mina_b200_host_reencode, kind, data, len(data))[0]


def host_account_abi_encode(account_proof: bytes):
    """Solidity ABI encoding of the account inside an account proof (the verifier's expected_encoded_account).
    None when the proof does not decode or the conversion fails like the reference's TryFrom."""
    lib = load()
    lib.mina_b200_host_account_abi_encode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    return _host_bytes_out(lib.mina_b200_host_account_abi_encode, account_proof, len(account_proof))[0]


def host_select_secure_chain(candidate: bytes, tip: bytes) -> int:
    """1 = Candidate, 0 = Bridge; raises on the reference's Err / undecodable input."""
    res = ctypes.c_int(-1)
    lib = load()
    lib.mina_b200_host_select_secure_chain.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
    rc = lib.mina_b200_host_select_secure_chain(candidate, len(candidate), tip, len(tip), ctypes.byref(res))
    if rc != 0:
        raise MinaB200Error({-1: "decode error", -2: "constants differ", -3: "needs state hash"}.get(rc, "error %d" % rc))
    return res.value


def host_vk_load(path: str):
    out = ctypes.create_string_buffer(28 * 64 + 14 * 32)
    meta = (ctypes.c_uint32 * 4)()
    rc = load().mina_b200_host_vk_load(path.encode(), out, meta)
    if rc != 0:
        raise MinaB200Error(load().mina_b200_last_error().decode())
    raw = out.raw
    ints = lambda off, n: [int.from_bytes(raw[off + 32 * i : off + 32 * i + 32], "little") for i in range(n)]
    comms = [(int.from_bytes(raw[64 * i : 64 * i + 32], "little"), int.from_bytes(raw[64 * i + 32 : 64 * i + 64], "little")) for i in range(28)]
    o = 28 * 64
    return {
        "commitments": comms, "shifts": ints(o, 7), "group_gen": ints(o + 7 * 32, 1)[0], "zk_w3": ints(o + 8 * 32, 1)[0],
        "zkpm": ints(o + 9 * 32, 4), "endo": ints(o + 13 * 32, 1)[0],
        "log_size_of_group": meta[0], "max_poly_size": meta[1], "public": meta[2], "prev_challenges": meta[3],
    }


def host_hash_with_kimchi(table: bytes, prefix: str, xs) -> int:
    out = ctypes.create_string_buffer(32)
    data = b"".join(int(x).to_bytes(32, "little") for x in xs)
    rc = load().mina_b200_host_hash_with_kimchi(table, prefix.encode(), data, ctypes.c_uint32(len(xs)), out)
    if rc != 0:
        raise MinaB200Error("host_hash_with_kimchi failed: %d" % rc)
    return int.from_bytes(out.raw, "little")


def host_poseidon_permute(field: int, table: bytes, states: bytes) -> bytes:
    n = len(states) // 96
    buf = ctypes.create_string_buffer(states, max(len(states), 1))
    rc = load().mina_b200_host_poseidon_permute(field, table, ctypes.c_uint32(n), buf)
    if rc != 0:
        raise MinaB200Error("host_poseidon_permute failed: %d" % rc)
    return buf.raw[: 96 * n]


# ---- device-resident accumulator batches and the resident user base set (bench legs, config 2) ------------
class KernelStats(ctypes.Structure):
    _fields_ = [("accumulate_ms", ctypes.c_float), ("combine_ms", ctypes.c_float), ("msm_points", ctypes.c_uint64),
                ("msm_count", ctypes.c_uint64), ("combine_proofs", ctypes.c_uint64), ("combine_vectors", ctypes.c_uint64)]


def accumulators_device(curve: int, m: int, d_pre: int, d_pts: int, mode: int = MODE_RLC, want_stats: bool = False):
    ok = ctypes.create_string_buffer(max(m, 1))
    st = KernelStats()
    _check(load().mina_b200_accumulators_device(curve, ctypes.c_uint32(m), ctypes.c_void_p(d_pre), ctypes.c_void_p(d_pts), mode, ok,
                                                ctypes.byref(st) if want_stats else None))
    return (ok.raw[:m], st) if want_stats else ok.raw[:m]


def fixed_base_load(curve: int, points: bytes, window_bits: int = 0):
    _check(load().mina_b200_fixed_base_load(curve, ctypes.c_uint32(len(points) // 64), points, window_bits))


def fixed_base_msm(curve: int, scalars: bytes, n: int) -> list[bytes]:
    nmsm = len(scalars) // (32 * n)
    out = ctypes.create_string_buffer(64 * max(nmsm, 1))
    _check(load().mina_b200_fixed_base_msm(curve, ctypes.c_uint32(nmsm), scalars, out))
    return [out.raw[64 * i : 64 * i + 64] for i in range(nmsm)]


def fixed_base_msm_device(curve: int, nmsm: int, d_scalars: int, d_out: int, stream: int, want_ms: bool = False):
    ms = ctypes.c_float(0.0)
    _check(load().mina_b200_fixed_base_msm_device(curve, ctypes.c_uint32(nmsm), ctypes.c_void_p(d_scalars), ctypes.c_void_p(d_out),
                                                  ctypes.c_void_p(stream), ctypes.byref(ms) if want_ms else None))
    return ms.value if want_ms else None


def host_srs_load_file(curve: int, path: str, count: int):
    out = ctypes.create_string_buffer(64 * max(count, 1))
    h = ctypes.create_string_buffer(64)
    rc = load().mina_b200_host_srs_load_file(curve, path.encode(), ctypes.c_uint32(count), out, h)
    if rc != 0:
        raise MinaB200Error(load().mina_b200_last_error().decode())
    return out.raw[: 64 * count], h.raw


def msm_srs_plus(curve: int, scalars_srs: bytes, scalars_extra: bytes, points_extra: bytes) -> bytes:
    out = ctypes.create_string_buffer(64)
    _check(load().mina_b200_msm_srs_plus(curve, ctypes.c_uint32(len(scalars_srs) // 32), scalars_srs, ctypes.c_uint32(len(scalars_extra) // 32),
                                         scalars_extra, points_extra, out))
    return out.raw


def combined_inner_product(field: int, evals: bytes, scales: bytes, npolys: int, npts: int) -> bytes:
    nproofs = len(scales) // 64
    out = ctypes.create_string_buffer(32 * max(nproofs, 1))
    _check(load().mina_b200_combined_inner_product(field, ctypes.c_uint32(nproofs), ctypes.c_uint32(npolys), ctypes.c_uint32(npts), evals, scales, out))
    return out.raw[: 32 * nproofs]


def state_accumulators_device(m: int, d_pre_w: int, d_pts_w: int, d_pre_s: int, d_pts_s: int, mode: int = MODE_RLC, want_stats: bool = False):
    """Both accumulator families of m state proofs, the two pipelines overlapped on two streams."""
    ok = ctypes.create_string_buffer(max(3 * m, 1))
    st = (KernelStats * 2)()
    _check(load().mina_b200_state_accumulators_device(ctypes.c_uint32(m), ctypes.c_void_p(d_pre_w), ctypes.c_void_p(d_pts_w),
                                                      ctypes.c_void_p(d_pre_s), ctypes.c_void_p(d_pts_s), mode, ok, st if want_stats else None))
    return (ok.raw[: 3 * m], (st[0], st[1])) if want_stats else ok.raw[: 3 * m]
