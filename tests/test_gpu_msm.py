"""GPU tier: the CUDA path, called through the C ABI, against the oracle (bit-exact)."""
import hashlib
import json
import os
import random

import pytest

from oracle import cref, pasta
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

MOD = {0: pasta.P, 1: pasta.Q}          # coordinate field of curve id
SCALAR_MOD = {0: pasta.Q, 1: pasta.P}   # scalar field of curve id
DEPTH = {0: 32768, 1: 65536}


def rand_scalars(rng, n, m):
    return cref.ints_to_bytes([rng.randrange(m) for _ in range(n)])


def test_device_field_ops(gpu):
    rng = random.Random(1)
    for fid, m in ((0, pasta.P), (1, pasta.Q)):
        a = [rng.randrange(m) for _ in range(2000)] + [0, 1, m - 1, m - 1, 2]
        b = [rng.randrange(m) for _ in range(2000)] + [m - 1, m - 1, m - 1, 1, (m + 1) // 2]
        A, B = cref.ints_to_bytes(a), cref.ints_to_bytes(b)
        assert gpu.field_op(fid, 0, A, B) == cref.ints_to_bytes([x * y % m for x, y in zip(a, b)])
        assert gpu.field_op(fid, 1, A, B) == cref.ints_to_bytes([(x + y) % m for x, y in zip(a, b)])
        assert gpu.field_op(fid, 2, A, B) == cref.ints_to_bytes([(x - y) % m for x, y in zip(a, b)])
        assert gpu.field_op(fid, 4, A) == cref.ints_to_bytes([x * x % m for x in a])
        assert gpu.field_op(fid, 3, A[: 32 * 64]) == cref.ints_to_bytes([pow(x, -1, m) for x in a[:64]])


def test_both_device_multipliers_match_the_portable_product(gpu):
    """raw Montgomery products (no conversion): op 11 = portable CIOS, 16 = mul_ptx, 17 = mul_ptx2."""
    rng = random.Random(2)
    pats = [0, 1, 0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 0xFFFFFFFE]
    for fid, m in ((0, pasta.P), (1, pasta.Q)):
        vals = [rng.randrange(m) for _ in range(3000)]
        # structured limbs: zero / all-ones limbs exercise the (v == 0) carry and the deferred-carry paths
        for _ in range(3000):
            v = 0
            for k in range(8):
                v |= (rng.choice(pats) if rng.random() < 0.7 else rng.getrandbits(32)) << (32 * k)
            vals.append(v % m)
        vals += [0, 1, m - 1, m - 2, (1 << 254) % m, (1 << 255) % m, 2**32 - 1, 2**64 - 1, 2**224]
        a, b = vals, vals[::-1]
        A, B = cref.ints_to_bytes(a), cref.ints_to_bytes(b)
        want = gpu.field_op(fid, 11, A, B)
        rinv = pow(1 << 256, -1, m)
        assert want == cref.ints_to_bytes([x * y * rinv % m for x, y in zip(a, b)])
        assert gpu.field_op(fid, 16, A, B) == want
        assert gpu.field_op(fid, 17, A, B) == want
        assert gpu.field_op(fid, 10, A, B) == want
        # sums of two / three products with ONE reduction (lazy reduction, Fd::dot2 / Fd::dot3): the worst case is
        # three products of values near p, which is what the tail of `vals` reversed against itself provides
        n = len(a)
        p2 = [(a[i] * b[i] + a[(i + 1) % n] * b[(i + 1) % n]) * rinv % m for i in range(n)]
        p3 = [(a[i] * b[i] + a[(i + 1) % n] * b[(i + 1) % n] + a[(i + 2) % n] * b[(i + 2) % n]) * rinv % m for i in range(n)]
        assert gpu.field_op(fid, 18, A, B) == cref.ints_to_bytes(p2)
        assert gpu.field_op(fid, 19, A, B) == cref.ints_to_bytes(p3)
        big = [m - 1, m - 2, m - 3, m - 1, m - 5, m - 1]
        Bg = cref.ints_to_bytes(big)
        nb = len(big)
        assert gpu.field_op(fid, 19, Bg, Bg) == cref.ints_to_bytes(
            [(big[i] ** 2 + big[(i + 1) % nb] ** 2 + big[(i + 2) % nb] ** 2) * rinv % m for i in range(nb)])
        # dedicated squaring (op 4 converts in and out of Montgomery form around Fd::sqr)
        assert gpu.field_op(fid, 4, A) == cref.ints_to_bytes([x * x % m for x in a])


def test_device_srs_is_the_reference_srs(gpu):
    pins = json.load(open(os.path.join(GOLDEN, "srs_sha256.json")))
    g, h = gpu.srs_points(1, 0, 65536, True)
    assert hashlib.sha256(g).hexdigest() == pins["vesta"]["sha256_g_65536"] and h.hex() == pins["vesta"]["h"]
    g, h = gpu.srs_points(0, 0, 32768, True)
    assert hashlib.sha256(g).hexdigest() == pins["pallas"]["sha256_g_32768"] and h.hex() == pins["pallas"]["h"]


def test_device_point_add_including_special_cases(gpu):
    for cid in (0, 1):
        m = MOD[cid]
        pts = gpu.srs_points(cid, 0, 64)
        P = [cref.bytes_to_point(pts[64 * i : 64 * i + 64]) for i in range(64)]
        a = list(P)
        b = P[1:] + P[:1]
        # doubling, cancellation and identities (even index = mixed formula, odd = full formula)
        a += [P[3], P[3], P[4], P[4], None, None, P[5], P[5], None, None]
        b += [P[3], P[3], pasta.neg(P[4], m), pasta.neg(P[4], m), P[6], P[6], None, None, None, None]
        got = gpu.point_add(cid, cref.points_to_bytes(a), cref.points_to_bytes(b))
        want = cref.points_to_bytes([pasta.add(x, y, m) for x, y in zip(a, b)])
        assert got == want


@pytest.mark.parametrize("cid", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 33, 1000, 4097])
def test_msm_srs_small_random(gpu, cid, n):
    rng = random.Random(100 + n)
    sc = rand_scalars(rng, n, SCALAR_MOD[cid])
    pts = gpu.srs_points(cid, 0, n)
    want, inf = cref.msm(cid, sc, pts, 4)
    (got,) = gpu.msm_srs(cid, sc, n)
    assert got == want and not inf


@pytest.mark.parametrize("cid", [0, 1])
def test_msm_srs_edge_scalars(gpu, cid):
    m = SCALAR_MOD[cid]
    n = 513
    pts = gpu.srs_points(cid, 0, n)
    rng = random.Random(9)
    cases = {
        "zeros": [0] * n,
        "ones": [1] * n,
        "minus_one": [m - 1] * n,
        "window_boundaries": [(1 << (16 * (i % 16))) % m for i in range(n)],
        "half_digits": [int("8000" * 15, 16) % m] * n,   # every 16-bit digit = 2^15 (signed-digit edge)
        "ffff_digits": [((1 << 254) - 1) % m] * n,
        "mixed": [rng.choice([0, 1, m - 1, rng.randrange(m)]) for _ in range(n)],
    }
    for name, sc in cases.items():
        scb = cref.ints_to_bytes(sc)
        want, inf = cref.msm(cid, scb, pts, 4)
        (got,) = gpu.msm_srs(cid, scb, n)
        assert got == want, name
    # cancelling pair -> identity: s*G0 + (m-s)*G0 needs generic bases
    s = rng.randrange(m)
    two = pts[:64] * 2
    got = gpu.msm(cid, cref.ints_to_bytes([s, m - s]), two)
    assert got == b"\0" * 64


def test_msm_empty(gpu):
    assert gpu.msm(1, b"", b"") == b"\0" * 64


@pytest.mark.parametrize("cid", [0, 1])
def test_msm_batch_of_independent_msms(gpu, cid):
    rng = random.Random(77)
    n, nmsm = 2048, 5
    pts = gpu.srs_points(cid, 0, n)
    sc = [rand_scalars(rng, n, SCALAR_MOD[cid]) for _ in range(nmsm)]
    got = gpu.msm_srs(cid, b"".join(sc), n)
    for k in range(nmsm):
        want, _ = cref.msm(cid, sc[k], pts, 4)
        assert got[k] == want


@pytest.mark.parametrize("cid", [0, 1])
@pytest.mark.parametrize("c", [0, 7, 13, 16])
def test_msm_generic_bases_windows(gpu, cid, c):
    rng = random.Random(31 + c)
    n = 3000
    # bases with repeats (forces the doubling branch inside buckets) and an identity
    base = gpu.srs_points(cid, 100, 500)
    pts = b"".join(base[64 * (i % 500) : 64 * (i % 500) + 64] for i in range(n - 1)) + b"\0" * 64
    sc = rand_scalars(rng, n, SCALAR_MOD[cid])
    want, _ = cref.msm(cid, sc, pts, 4)
    assert gpu.msm(cid, sc, pts, c) == want


@pytest.mark.parametrize("c", [2, 3, 4, 5, 6, 9, 10, 11, 12, 14, 15, 17, 18, 19, 20])
def test_msm_every_window_width(gpu, c):
    """The reduction tail is planned per window width (running-sum passes, bit sums, partial sums, the flat and the
    level-by-level finalize): every width from 2 to 20 must give the oracle's point, including the widths whose top
    window only carries (c = 5, 15, 17) and the ones with more than 32 slots per group (c >= 17)."""
    rng = random.Random(900 + c)
    n = 257
    pts = gpu.srs_points(1, 7, n)
    sc = rand_scalars(rng, n - 3, SCALAR_MOD[1]) + cref.ints_to_bytes([SCALAR_MOD[1] - 1, 1, 0])
    want, _ = cref.msm(1, sc, pts, 2)
    assert gpu.msm(1, sc, pts, c) == want


def test_accumulator_kats_on_gpu(gpu, state_proof):
    # K-A, K-B, K-C with ORACLE-prepared scalars: pins the MSM engine alone.  The full row (device endo +
    # b_poly + MSM from raw proof bytes) is tests/test_gpu_verifier.py::test_accumulator_check_*.
    pr = state_proof["candidate_tip_proof"]
    pre = b"".join(x.to_bytes(16, "little") for x in pr["bulletproof_challenges"])
    s = cref.bpoly_coeffs(cref.FP, cref.endo_to_field(cref.FP, pre, pasta.ENDO_FP))
    (got,) = gpu.msm_srs(1, s, 65536)
    assert cref.bytes_to_point(got) == pr["wrap_challenge_polynomial_commitment"]
    for k in range(2):
        pre = b"".join(x.to_bytes(16, "little") for x in pr["wrap_old_bulletproof_challenges"][k])
        s = cref.bpoly_coeffs(cref.FQ, cref.endo_to_field(cref.FQ, pre, pasta.ENDO_FQ))
        (got,) = gpu.msm_srs(0, s, 32768)
        assert cref.bytes_to_point(got) == pr["step_challenge_polynomial_commitments"][k]


def test_msm_full_depth_random_and_linearity(gpu):
    # full SRS depth vs the oracle, plus a size-independent property: MSM(a) + MSM(b) == MSM(a+b)
    rng = random.Random(2024)
    for cid in (0, 1):
        n, m, fm = DEPTH[cid], SCALAR_MOD[cid], MOD[cid]
        a = [rng.randrange(m) for _ in range(n)]
        b = [rng.randrange(m) for _ in range(n)]
        ab = [(x + y) % m for x, y in zip(a, b)]
        ra, rb, rab = gpu.msm_srs(cid, cref.ints_to_bytes(a) + cref.ints_to_bytes(b) + cref.ints_to_bytes(ab), n)
        want, _ = cref.msm(cid, cref.ints_to_bytes(a), gpu.srs_points(cid, 0, n), 8)
        assert ra == want
        assert pasta.add(cref.bytes_to_point(ra), cref.bytes_to_point(rb), fm) == cref.bytes_to_point(rab)


def test_lagrange_commitments_match_the_oracle(gpu):
    """a5: kimchi `add_lagrange_basis` over the wrap domain (2^14, Pallas; verifier_index.rs:204-208).  L_i has
    coefficients omega^(-i j) / n with omega the domain generator of the VK (K-G): the device commitment must equal the
    oracle MSM of those coefficients, and the basis must behave like one: sum_i L_i commits the constant 1 = g[0]."""
    log_n, n, q = 14, 1 << 14, pasta.Q
    omega = int("1E5587687024253BB079B38D9C5371594958E496C605D3BD898B34D068AFBEE7", 16)  # devnet_vk.json index.domain.group_gen
    assert pow(omega, n, q) == 1 and pow(omega, n // 2, q) != 1
    w_inv, n_inv = pow(omega, q - 2, q), pow(n, q - 2, q)
    g = gpu.srs_points(0, 0, n)
    idx = [0, 1, 39, n - 1]
    got = [gpu.lagrange_commitments(0, log_n, i, 1)[0] for i in idx]
    for i, pt in zip(idx, got):
        coeffs, cur, step = [], n_inv, pow(w_inv, i, q)
        for _ in range(n):
            coeffs.append(cur)
            cur = cur * step % q
        want, inf = cref.msm(cref.FP, cref.ints_to_bytes(coeffs), g, 4)
        assert not inf and pt == want, i
    # a small domain in full: the commitments of all L_i add up to the commitment of the constant polynomial 1
    small = gpu.lagrange_commitments(0, 4, 0, 16)
    acc = None
    for pt in small:
        acc = pasta.add(acc, cref.bytes_to_point(pt), pasta.P)
    assert acc == cref.bytes_to_point(g[:64])
    # and the 40 the verifier uses come out in one call, equal to the single calls
    first40 = gpu.lagrange_commitments(0, log_n, 0, 40)
    assert first40[0] == got[0] and first40[1] == got[1] and first40[39] == got[2]


def test_public_input_commitment_matches_the_oracle(gpu):
    """kimchi `public_comm` (part of a8): -sum_i pub_i L_i + h over the 40 Lagrange commitments of the wrap domain,
    for a batch of public-input vectors incl. edge values, against the oracle MSM over the same bases."""
    log_n, n_pub, q = 14, 40, pasta.Q
    lag = gpu.lagrange_commitments(0, log_n, 0, n_pub)
    h = gpu.srs_points(0, 0, 1, want_h=True)[1]
    bases = b"".join(lag) + h
    rng = random.Random(99)
    vecs = [[rng.randrange(q) for _ in range(n_pub)] for _ in range(5)]
    vecs.append([0] * n_pub)                      # all zero: the commitment is h itself
    vecs.append([1] + [0] * (n_pub - 1))
    vecs.append([q - 1] * n_pub)
    got = gpu.public_commitments(0, log_n, n_pub, b"".join(cref.ints_to_bytes(v) for v in vecs))
    for v, pt in zip(vecs, got):
        want, inf = cref.msm(cref.FP, cref.ints_to_bytes([(-x) % q for x in v] + [1]), bases, 2)
        assert not inf and pt == want
    assert got[5] == h
