"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- Pasta fields, curves and the IPA helper arithmetic.

This file is a plain-Python-integer restatement of arithmetic whose production home is the
un-vendored Rust crates the reference links:

  * ark-ff 0.3 Fp256 / ark-ec 0.3 short Weierstrass (lambdaclass/openmina_algebra @ 017531e)
  * poly-commitment `SRS::create`, `b_poly_coefficients`, `b_poly`, `endos`
    (lambdaclass/openmina-proof-systems @ 44e0d3b), called from
    AL/operator/mina/lib/src/lib.rs:34 and AL/operator/mina/lib/src/verifier_index.rs:169,204-208
  * kimchi `ScalarChallenge::to_field`
  * groupmap `BWParameters::setup` / `to_group`

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The shipped verifier (mina_bridge_b200/) never does.

Parity status: pinned.  Every routine here is checked by tests/test_oracle_kats.py against
data the reference committed: the three accumulator MSM known-answer points inside
AL/scripts/test_files/mina/mina_state.proof and the srs/{vesta,pallas}.srs files.
"""
from __future__ import annotations

import hashlib
import struct

# --- moduli (mina-curves pasta) --------------------------------------------
P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001  # Fp: Pallas base / Vesta scalar
Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001  # Fq: Vesta base / Pallas scalar
B_COEFF = 5  # y^2 = x^3 + 5 on both curves
TWO_ADICITY = 32
NONRESIDUE = 5  # multiplicative generator of both fields


class Curve:
    """Short-Weierstrass curve y^2 = x^3 + 5 over `base`, prime order `scalar`."""

    def __init__(self, name: str, base: int, scalar: int):
        self.name, self.base, self.scalar = name, base, scalar
        # endo_q (base field cube root of unity) and endo_r (scalar field cube root of unity)
        # follow poly-commitment `endos::<G>()`: endo_q = g^((p-1)/3) for the field generator,
        # endo_r is the matching eigenvalue.  The reference stores `endos::<Vesta>().0` in the
        # verifier index (verifier_index.rs:169).


PALLAS = Curve("pallas", P, Q)
VESTA = Curve("vesta", Q, P)


def inv(a: int, m: int) -> int:
    return pow(a, -1, m)


# --- square roots: the exact Tonelli-Shanks ark-ff 0.3 runs -----------------
def sqrt(a: int, m: int):
    """Return the root ark-ff's `SquareRootField::sqrt` returns, or None for a non-residue."""
    a %= m
    if a == 0:
        return 0
    if pow(a, (m - 1) // 2, m) != 1:
        return None
    s = TWO_ADICITY
    t_odd = (m - 1) >> s
    z = pow(NONRESIDUE, t_odd, m)  # 2^s-th root of unity
    w = pow(a, (t_odd - 1) // 2, m)
    x = a * w % m  # a^((t+1)/2)
    b = x * w % m  # a^t
    v = s
    while b != 1:
        k = 0
        b2k = b
        while b2k != 1:
            b2k = b2k * b2k % m
            k += 1
        j = v - k - 1
        w = z
        for _ in range(j):
            w = w * w % m
        z = w * w % m
        b = b * z % m
        x = x * w % m
        v = k
    return x


# --- curve arithmetic (affine tuples, None = identity) -------------------------------------------
def is_on_curve(pt, m: int) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % m == 0


def add(p1, p2, m: int):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % m == 0:
            return None
        lam = 3 * x1 * x1 * inv(2 * y1, m) % m
    else:
        lam = (y2 - y1) * inv(x2 - x1, m) % m
    x3 = (lam * lam - x1 - x2) % m
    return (x3, (lam * (x1 - x3) - y1) % m)


def neg(pt, m: int):
    return None if pt is None else (pt[0], (-pt[1]) % m)


# Jacobian (X, Y, Z) arithmetic for the bulk loops (a = 0).
def jdouble(pt, m):
    X, Y, Z = pt
    if Z == 0:
        return pt
    A = X * X % m
    Bq = Y * Y % m
    C = Bq * Bq % m
    D = 2 * ((X + Bq) * (X + Bq) - A - C) % m
    E = 3 * A % m
    F = E * E % m
    X3 = (F - 2 * D) % m
    Y3 = (E * (D - X3) - 8 * C) % m
    Z3 = 2 * Y * Z % m
    return (X3, Y3, Z3)


def jadd_mixed(pt, q, m):
    """Jacobian += affine (q is an (x, y) tuple or None)."""
    if q is None:
        return pt
    X1, Y1, Z1 = pt
    x2, y2 = q
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % m
    U2 = x2 * Z1Z1 % m
    S2 = y2 * Z1 * Z1Z1 % m
    if U2 == X1:
        if S2 == Y1:
            return jdouble(pt, m)
        return (1, 1, 0)
    H = (U2 - X1) % m
    HH = H * H % m
    I = 4 * HH % m
    J = H * I % m
    r = 2 * (S2 - Y1) % m
    V = X1 * I % m
    X3 = (r * r - J - 2 * V) % m
    Y3 = (r * (V - X3) - 2 * Y1 * J) % m
    Z3 = ((Z1 + H) * (Z1 + H) - Z1Z1 - HH) % m
    return (X3, Y3, Z3)


def jadd(p1, p2, m):
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if Z1 == 0:
        return p2
    if Z2 == 0:
        return p1
    Z1Z1 = Z1 * Z1 % m
    Z2Z2 = Z2 * Z2 % m
    U1 = X1 * Z2Z2 % m
    U2 = X2 * Z1Z1 % m
    S1 = Y1 * Z2 * Z2Z2 % m
    S2 = Y2 * Z1 * Z1Z1 % m
    if U1 == U2:
        if S1 == S2:
            return jdouble(p1, m)
        return (1, 1, 0)
    H = (U2 - U1) % m
    I = 4 * H * H % m
    J = H * I % m
    r = 2 * (S2 - S1) % m
    V = U1 * I % m
    X3 = (r * r - J - 2 * V) % m
    Y3 = (r * (V - X3) - 2 * S1 * J) % m
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % m
    return (X3, Y3, Z3)


JZERO = (1, 1, 0)


def to_affine(pt, m):
    X, Y, Z = pt
    if Z == 0:
        return None
    zi = inv(Z, m)
    zi2 = zi * zi % m
    return (X * zi2 % m, Y * zi2 * zi % m)


def scalar_mul(k: int, pt, m: int):
    acc = JZERO
    if pt is None:
        return None
    for bit in bin(k)[2:] if k else "":
        acc = jdouble(acc, m)
        if bit == "1":
            acc = jadd_mixed(acc, pt, m)
    return to_affine(acc, m)


def msm_naive(scalars, points, m):
    """Sum s_i * P_i by double-and-add.  Only for tiny inputs."""
    acc = None
    for s, pt in zip(scalars, points):
        acc = add(acc, scalar_mul(s, pt, m), m)
    return acc


def ark_window_bits(n: int) -> int:
    """ark-ec 0.3 `VariableBaseMSM` window choice: 3 if n < 32 else ln(n)+2
    with ln computed as log2(n) * 69 / 100 on integers."""
    if n < 32:
        return 3
    log2 = (n - 1).bit_length() if n & (n - 1) else n.bit_length() - 1
    # ark_std::log2 is ceil(log2(n)).
    return log2 * 69 // 100 + 2


def msm_pippenger(scalars, points, m, scalar_bits: int = 255, c: int | None = None):
    """Bucket-method MSM following ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul`
    (openmina_algebra @ 017531e; the inner loop of SURVEY rows a7/a9/a10).
    Returns an affine tuple or None."""
    n = min(len(scalars), len(points))
    if c is None:
        c = ark_window_bits(n)
    window_sums = []
    for w_start in range(0, scalar_bits, c):
        res = JZERO
        buckets = [JZERO] * ((1 << c) - 1)
        for s, pt in zip(scalars[:n], points[:n]):
            if s == 0 or pt is None:
                continue
            if s == 1:
                if w_start == 0:
                    res = jadd_mixed(res, pt, m)
                continue
            d = (s >> w_start) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = jadd_mixed(buckets[d - 1], pt, m)
        running = JZERO
        for b in reversed(buckets):
            running = jadd(running, b, m)
            res = jadd(res, running, m)
        window_sums.append(res)
    total = window_sums[-1]
    for ws in reversed(window_sums[:-1]):
        for _ in range(c):
            total = jdouble(total, m)
        total = jadd(total, ws, m)
    return to_affine(total, m)


# --- endomorphism constants: poly-commitment `endos::<G>()` -------------------------------------
def _cube_root_of_unity(m: int) -> int:
    return pow(NONRESIDUE, (m - 1) // 3, m)


OMEGA_P = _cube_root_of_unity(P)
OMEGA_Q = _cube_root_of_unity(Q)
# endo_r used when decoding challenges INTO each field (verified by KATs K-A / K-B / K-C):
ENDO_FP = OMEGA_P * OMEGA_P % P
ENDO_FQ = OMEGA_Q * OMEGA_Q % Q


def endo_to_field(limbs128: int, endo: int, m: int) -> int:
    """kimchi `ScalarChallenge::to_field_with_length(128, endo)`."""
    a = b = 2
    for i in range(63, -1, -1):
        a = 2 * a % m
        b = 2 * b % m
        r_2i = (limbs128 >> (2 * i)) & 1
        s = 1 if r_2i else m - 1
        if (limbs128 >> (2 * i + 1)) & 1 == 0:
            b = (b + s) % m
        else:
            a = (a + s) % m
    return (a * endo + b) % m


def b_poly_coefficients(chals, m: int):
    """poly-commitment `b_poly_coefficients`:
    s[i] = prod_j chals[k-1-j]^{bit_j(i)}."""
    k = len(chals)
    s = [1] * (1 << k)
    pw = 1
    kk = 0
    for i in range(1, 1 << k):
        if i == pw << 1:
            pw <<= 1
            kk += 1
        s[i] = s[i - pw] * chals[k - 1 - kk] % m
    return s


def b_poly(chals, x: int, m: int) -> int:
    """poly-commitment `b_poly`: prod_i (1 + chals[i] * x^(2^(k-1-i)))."""
    k = len(chals)
    pow_twos = [x % m]
    for _ in range(1, k):
        pow_twos.append(pow_twos[-1] * pow_twos[-1] % m)
    r = 1
    for i in range(k):
        r = r * (1 + chals[i] * pow_twos[k - 1 - i]) % m
    return r


# --- group map (groupmap `BWParameters`) -----------------------------------
class GroupMap:
    def __init__(self, m: int):
        self.m = m
        self.u = 1
        self.fu = (1 + B_COEFF) % m
        three_u2 = 3 % m
        self.inv_three_u_squared = inv(three_u2, m)
        self.sqrt_neg_three_u_squared = sqrt((-three_u2) % m, m)
        self.sqrt_neg_three_u_squared_minus_u_over_2 = (
            (self.sqrt_neg_three_u_squared - self.u) * inv(2, m) % m
        )

    def potential_xs(self, t: int):
        m = self.m
        t2 = t * t % m
        alpha_inv = (t2 + self.fu) * t2 % m
        alpha = inv(alpha_inv, m) if alpha_inv else 0
        temp = t2 * t2 % m * alpha % m * self.sqrt_neg_three_u_squared % m
        x1 = (self.sqrt_neg_three_u_squared_minus_u_over_2 - temp) % m
        x2 = (-self.u - x1) % m
        t2_plus_fu = (t2 + self.fu) % m
        t2_inv = alpha * t2_plus_fu % m
        x3 = (self.u - t2_plus_fu * t2_plus_fu % m * t2_inv % m * self.inv_three_u_squared) % m
        return (x1, x2, x3)

    def to_group(self, t: int):
        for x in self.potential_xs(t):
            y = sqrt((x * x * x + B_COEFF) % self.m, self.m)
            if y is not None:
                return (x, y)
        raise ValueError("group map failed")


def _srs_hash_to_field(data: bytes, m: int) -> int:
    """poly-commitment srs.rs `point_of_random_bytes`: first 31 digest bytes, bits LSB-first in each
    byte, read as a big-endian bit string."""
    digest = hashlib.blake2b(data, digest_size=64).digest()
    v = 0
    for byte in digest[:31]:
        for j in range(8):
            v = (v << 1) | ((byte >> j) & 1)
    return v % m


def srs_point(i: int, gm: GroupMap):
    """`SRS::create`: g[i] = to_group(hash(be32(i)))."""
    return gm.to_group(_srs_hash_to_field(struct.pack(">I", i), gm.m))


def srs_blinding(gm: GroupMap):
    """`SRS::create`: h = to_group(hash("srs_misc" || be32(0)))."""
    return gm.to_group(_srs_hash_to_field(b"srs_misc" + struct.pack(">I", 0), gm.m))


# --- committed SRS files --------------------------------------------------
def decompress_point(buf33: bytes, m: int):
    x = int.from_bytes(buf33[:32], "little")
    flag = buf33[32]
    if flag & 0x40:
        return None
    y = sqrt((x * x * x + B_COEFF) % m, m)
    if y is None:
        raise ValueError("x not on curve")
    y_is_big = y > (m - 1) // 2
    if bool(flag & 0x80) != y_is_big:
        y = m - y
    return (x, y)


def read_srs_file(path: str, m: int, limit: int | None = None):
    """Parse srs/{pallas,vesta}.srs -> (g list, h).  MessagePack: array2[array32[bin33...], bin33]."""
    data = open(path, "rb").read()
    assert data[0] == 0x92 and data[1] == 0xDD
    n = struct.unpack(">I", data[2:6])[0]
    off = 6
    g = []
    for i in range(n):
        assert data[off] == 0xC4 and data[off + 1] == 33
        if limit is None or i < limit:
            g.append(decompress_point(data[off + 2 : off + 35], m))
        off += 35
    assert data[off] == 0xC4 and data[off + 1] == 33
    h = decompress_point(data[off + 2 : off + 35], m)
    assert off + 35 == len(data)
    return g, h


def srs_compressed_bytes(path: str):
    """Raw 33-byte compressed encodings from an .srs file: (list of g encodings, h encoding)."""
    data = open(path, "rb").read()
    n = struct.unpack(">I", data[2:6])[0]
    g = [data[6 + 35 * i + 2 : 6 + 35 * i + 35] for i in range(n)]
    off = 6 + 35 * n
    return g, data[off + 2 : off + 35]


def compress_point(pt, m: int) -> bytes:
    x, y = pt
    return x.to_bytes(32, "little") + bytes([0x80 if y > (m - 1) // 2 else 0])
