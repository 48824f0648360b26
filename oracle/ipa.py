"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- the IPA opening proof of poly-commitment, prover and verifier.

Restates `SRS::open` / `SRS::verify`, `OpeningProof::challenges`, `combine_commitments`, `shift_scalar` and the
`DefaultFqSponge` wrapper (absorb_g / absorb_fr / challenge / challenge_fq) of poly-commitment and mina-poseidon
@ lambdaclass/openmina-proof-systems 44e0d3b -- un-vendored, restated from the published algorithm
[UPSTREAM-KNOWLEDGE].  Reference call site: kimchi `verify` under `verify_block`,
AL/operator/mina/lib/src/lib.rs:99-111.

PARITY UNPINNED: the reference holds no vector for the opening proof alone, and the sponge needs the kimchi
Poseidon table, which is unavailable (oracle/poseidon.py).  What this module gives the tests is
SELF-CONSISTENCY: a prover that follows `SRS::open` produces openings under an arbitrary table, and the device
verifier must accept exactly those and reject every mutation, agreeing bit for bit with `verify_one` below.
Only tests/ may import this module.
"""
from __future__ import annotations

from . import cref, pasta
from .poseidon import Sponge


class CurveCtx:
    """curve id 0 = Pallas (coordinates in Fp, scalars in Fq), 1 = Vesta (coordinates in Fq, scalars in Fp)"""

    def __init__(self, curve_id: int):
        self.id = curve_id
        self.base = pasta.P if curve_id == 0 else pasta.Q
        self.scalar = pasta.Q if curve_id == 0 else pasta.P
        self.base_fid = curve_id            # cref field id of the coordinates
        self.scalar_fid = 1 - curve_id      # cref field id of the scalars
        self.endo_r = pasta.ENDO_FQ if curve_id == 0 else pasta.ENDO_FP


def shift_scalar(cv: CurveCtx, x: int) -> int:
    """poly-commitment `shift_scalar::<G>`"""
    two_pow = pow(2, 255, cv.scalar)
    if cv.scalar < cv.base:
        return (x - (two_pow + 1)) * pasta.inv(2, cv.scalar) % cv.scalar
    return (x - two_pow) % cv.scalar


class FqSponge:
    """mina-poseidon `DefaultFqSponge`: a sponge over the curve's BASE field that also takes scalars"""

    def __init__(self, cv: CurveCtx, table):
        self.cv = cv
        self.s = Sponge(table, cv.base)
        self.last_squeezed = []

    def absorb_g(self, pt):
        self.last_squeezed = []
        x, y = (0, 0) if pt is None else pt
        self.s.absorb(x)
        self.s.absorb(y)

    def absorb_fq(self, x):
        self.last_squeezed = []
        self.s.absorb(x)

    def absorb_fr(self, x):
        self.last_squeezed = []
        if self.cv.scalar < self.cv.base:
            self.s.absorb(x)
        else:
            self.s.absorb(x >> 1)
            self.s.absorb(x & 1)

    def challenge(self) -> int:
        """128 bits: the two low limbs of one squeezed element (CHALLENGE_LENGTH_IN_LIMBS = HIGH_ENTROPY_LIMBS = 2)"""
        if len(self.last_squeezed) >= 2:
            lo, hi = self.last_squeezed[:2]
            self.last_squeezed = self.last_squeezed[2:]
        else:
            x = self.s.squeeze()
            lo, hi = x & (2**64 - 1), (x >> 64) & (2**64 - 1)
        return lo | hi << 64

    def challenge_fq(self) -> int:
        self.last_squeezed = []
        return self.s.squeeze()

    def export(self):
        """(state[3], mode, count): what the device kernel resumes from"""
        return list(self.s.state), (0 if self.s.absorbing else 1), self.s.count


def to_field(cv: CurveCtx, pre: int) -> int:
    return pasta.endo_to_field(pre, cv.endo_r, cv.scalar)


def commit(cv: CurveCtx, g_bytes: bytes, coeffs) -> tuple:
    out, inf = cref.msm(cv.base_fid, cref.ints_to_bytes(coeffs), g_bytes[: 64 * len(coeffs)], 4)
    return None if inf else cref.bytes_to_point(out)


def _fold_coeffs(chals, j, k, m):
    """coefficient of original base h * 2^(k-j) + i inside the j-times folded base i: prod_{r<j} u_r^{bit_(j-1-r)(h)}"""
    coef = [1]
    for r in range(j):  # round r is the next lower bit of h
        coef = [c * f % m for c in coef for f in (1, chals[r])]
    return coef


def open_proof(cv: CurveCtx, g_bytes: bytes, h_pt, k: int, polys, elm, polyscale: int, evalscale: int, sponge: FqSponge, rng):
    """`SRS::open` for unblinded-commitment polynomials of one chunk each (coefficient lists of length <= 2^k).
    Returns the opening and the combined inner product the verifier is handed."""
    m, n = cv.scalar, 1 << k
    p = [0] * n
    xi = 1
    for f in polys:
        for i, c in enumerate(f):
            p[i] = (p[i] + xi * c) % m
        xi = xi * polyscale % m
    b = [0] * n
    scale = 1
    for e in elm:
        res = scale
        for i in range(n):
            b[i] = (b[i] + res) % m
            res = res * e % m
        scale = scale * evalscale % m
    cip = sum(x * y for x, y in zip(p, b)) % m
    sponge.absorb_fr(shift_scalar(cv, cip))
    t = sponge.challenge_fq()
    u_pt = cref.bytes_to_point(cref.to_group(cv.base_fid, t))
    hu = cref.points_to_bytes([h_pt, u_pt])
    a = p
    lr, blinders, chals, chal_invs = [], [], [], []
    for j in range(k):
        half = len(a) // 2
        a_lo, a_hi, b_lo, b_hi = a[:half], a[half:], b[:half], b[half:]
        rand_l, rand_r = rng.randrange(m), rng.randrange(m)
        coef = _fold_coeffs(chals, j, k, m)
        size = len(a)  # = 2^(k-j): folded base i is sum_h coef[h] * g[h * size + i]
        sc_l, sc_r = [0] * n, [0] * n
        for hh, cf in enumerate(coef):
            for i in range(half):
                sc_l[hh * size + i] = a_hi[i] * cf % m          # <a_hi, g_lo>
                sc_r[hh * size + half + i] = a_lo[i] * cf % m   # <a_lo, g_hi>
        ip_l = sum(x * y for x, y in zip(a_hi, b_lo)) % m
        ip_r = sum(x * y for x, y in zip(a_lo, b_hi)) % m
        out, inf = cref.msm(cv.base_fid, cref.ints_to_bytes(sc_l + [rand_l, ip_l]), g_bytes[: 64 * n] + hu, 4)
        l_pt = None if inf else cref.bytes_to_point(out)
        out, inf = cref.msm(cv.base_fid, cref.ints_to_bytes(sc_r + [rand_r, ip_r]), g_bytes[: 64 * n] + hu, 4)
        r_pt = None if inf else cref.bytes_to_point(out)
        lr.append((l_pt, r_pt))
        blinders.append((rand_l, rand_r))
        sponge.absorb_g(l_pt)
        sponge.absorb_g(r_pt)
        u = to_field(cv, sponge.challenge())
        u_inv = pasta.inv(u, m)
        chals.append(u)
        chal_invs.append(u_inv)
        a = [(hi * u_inv + lo) % m for hi, lo in zip(a_hi, a_lo)]
        b = [(hi * u + lo) % m for hi, lo in zip(b_hi, b_lo)]
    a0, b0 = a[0], b[0]
    s = pasta.b_poly_coefficients(chals, m)
    out, inf = cref.msm(cv.base_fid, cref.ints_to_bytes(s), g_bytes[: 64 * n], 4)
    g0 = cref.bytes_to_point(out)
    r_prime = sum(l * ui + r * u for (l, r), u, ui in zip(blinders, chals, chal_invs)) % m  # commitments are unblinded
    d, r_delta = rng.randrange(m), rng.randrange(m)
    out, inf = cref.msm(cv.base_fid, cref.ints_to_bytes([d, d * b0 % m, r_delta]), cref.points_to_bytes([g0, u_pt, h_pt]), 1)
    delta = cref.bytes_to_point(out)
    sponge.absorb_g(delta)
    c = to_field(cv, sponge.challenge())
    return {"lr": lr, "delta": delta, "z1": (a0 * c + d) % m, "z2": (c * r_prime + r_delta) % m, "sg": g0}, cip


def verify_one(cv: CurveCtx, g_bytes: bytes, h_pt, k: int, commitments, elm, polyscale, evalscale, sponge: FqSponge, opening, cip) -> bool:
    """`SRS::verify` for one opening (rand_base = sg_rand_base = 1): the MSM must be the identity"""
    m, n = cv.scalar, 1 << k
    sponge.absorb_fr(shift_scalar(cv, cip))
    t = sponge.challenge_fq()
    u_pt = cref.bytes_to_point(cref.to_group(cv.base_fid, t))
    chal = []
    for l_pt, r_pt in opening["lr"]:
        sponge.absorb_g(l_pt)
        sponge.absorb_g(r_pt)
        chal.append(to_field(cv, sponge.challenge()))
    chal_inv = [pasta.inv(x, m) for x in chal]
    sponge.absorb_g(opening["delta"])
    c = to_field(cv, sponge.challenge())
    b0, scale = 0, 1
    for e in elm:
        b0 = (b0 + scale * pasta.b_poly(chal, e, m)) % m
        scale = scale * evalscale % m
    s = pasta.b_poly_coefficients(chal, m)
    z1, z2 = opening["z1"], opening["z2"]
    points, scalars = [h_pt], [-z2 % m]
    scalars += s  # over g[0..n)
    extra_pts, extra_sc = [opening["sg"]], [(-z1 - 1) % m]
    extra_pts.append(u_pt)
    extra_sc.append(-z1 * b0 % m)
    for (l_pt, r_pt), x, xi in zip(opening["lr"], chal, chal_inv):
        extra_pts += [l_pt, r_pt]
        extra_sc += [c * xi % m, c * x % m]
    xi_i = 1
    for cm in commitments:
        extra_pts.append(cm)
        extra_sc.append(c * xi_i % m)
        xi_i = xi_i * polyscale % m
    extra_pts += [u_pt, opening["delta"]]
    extra_sc += [c * cip % m, 1]
    out, inf = cref.msm(cv.base_fid, cref.ints_to_bytes(scalars + extra_sc),
                        cref.points_to_bytes(points) + g_bytes[: 64 * n] + cref.points_to_bytes(extra_pts), 4)
    return inf
