// Host-side Pasta field / curve arithmetic (4 x 64-bit limbs, Montgomery R = 2^256) used by the
// Fiat-Shamir driver, the loaders and the SRS derivation.  Same in-memory representation as the
// device `fe` (8 x u32 little-endian == 4 x u64 little-endian), so values move to the GPU by memcpy.
//
// Host-side equivalent of ark-ff 0.3 `Fp256<FpParameters>` / `Fp256<FqParameters>` (mina-curves,
// lambdaclass/openmina-proof-systems @ 44e0d3b) as used throughout AL/operator/mina/lib/src/*.rs.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <string>

#include "pasta_params.h"

namespace pasta {
namespace host {

typedef unsigned __int128 u128;

template <class F>
struct Fe {
    uint64_t l[4];

    static Fe zero() { return Fe{{0, 0, 0, 0}}; }
    static Fe one() { return Fe{{F::R1_64(0), F::R1_64(1), F::R1_64(2), F::R1_64(3)}}; }
    static Fe raw(uint64_t a, uint64_t b, uint64_t c, uint64_t d) { return Fe{{a, b, c, d}}; }
    static Fe modulus_raw() { return Fe{{F::MOD_64(0), F::MOD_64(1), F::MOD_64(2), F::MOD_64(3)}}; }

    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fe &o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
    bool operator!=(const Fe &o) const { return !(*this == o); }

    static bool geq_raw(const Fe &a, const Fe &b) {
        for (int i = 3; i >= 0; i--) {
            if (a.l[i] > b.l[i]) return true;
            if (a.l[i] < b.l[i]) return false;
        }
        return true;
    }
    static uint64_t add_raw(Fe &o, const Fe &a, const Fe &b) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)a.l[i] + b.l[i];
            o.l[i] = (uint64_t)c;
            c >>= 64;
        }
        return (uint64_t)c;
    }
    static uint64_t sub_raw(Fe &o, const Fe &a, const Fe &b) {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)a.l[i] - b.l[i] - borrow;
            o.l[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 127);
        }
        return borrow;
    }

    Fe operator+(const Fe &b) const {
        Fe t;
        add_raw(t, *this, b);
        Fe p = modulus_raw();
        if (geq_raw(t, p)) sub_raw(t, t, p);
        return t;
    }
    Fe operator-(const Fe &b) const {
        Fe t;
        if (sub_raw(t, *this, b)) {
            Fe p = modulus_raw();
            add_raw(t, t, p);
        }
        return t;
    }
    Fe operator-() const { return zero() - *this; }
    Fe dbl() const { return *this + *this; }

    Fe operator*(const Fe &b) const {
        // Montgomery CIOS specialised to the Pasta shape: MOD limb 2 is zero.
        uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
        for (int i = 0; i < 4; i++) {
            u128 c = (u128)l[0] * b.l[i] + t0;
            t0 = (uint64_t)c;
            c >>= 64;
            c += (u128)l[1] * b.l[i] + t1;
            t1 = (uint64_t)c;
            c >>= 64;
            c += (u128)l[2] * b.l[i] + t2;
            t2 = (uint64_t)c;
            c >>= 64;
            c += (u128)l[3] * b.l[i] + t3;
            t3 = (uint64_t)c;
            c >>= 64;
            c += t4;
            t4 = (uint64_t)c;
            t5 = (uint64_t)(c >> 64);
            uint64_t m = t0 * F::NINV64;
            c = (u128)m * F::MOD_64(0) + t0;
            c >>= 64;
            c += (u128)m * F::MOD_64(1) + t1;
            t0 = (uint64_t)c;
            c >>= 64;
            c += t2;  // MOD_64(2) == 0
            t1 = (uint64_t)c;
            c >>= 64;
            c += (u128)m * F::MOD_64(3) + t3;
            t2 = (uint64_t)c;
            c >>= 64;
            c += t4;
            t3 = (uint64_t)c;
            t4 = t5 + (uint64_t)(c >> 64);
        }
        Fe r{{t0, t1, t2, t3}};
        Fe p = modulus_raw();
        if (t4 || geq_raw(r, p)) sub_raw(r, r, p);
        return r;
    }
    Fe sqr() const { return *this * *this; }
    Fe &operator+=(const Fe &b) { return *this = *this + b; }
    Fe &operator-=(const Fe &b) { return *this = *this - b; }
    Fe &operator*=(const Fe &b) { return *this = *this * b; }

    // plain integer (canonical) <-> Montgomery
    static Fe from_canonical(const Fe &c) { return c * Fe{{F::R2_64(0), F::R2_64(1), F::R2_64(2), F::R2_64(3)}}; }
    Fe to_canonical() const { return *this * Fe{{1, 0, 0, 0}}; }
    static Fe from_u64(uint64_t v) { return from_canonical(Fe{{v, 0, 0, 0}}); }

    // 32 little-endian bytes; returns false if the integer is >= p
    static bool from_bytes_le(const uint8_t *b, Fe &out) {
        Fe c;
        std::memcpy(c.l, b, 32);
        if (geq_raw(c, modulus_raw())) return false;
        out = from_canonical(c);
        return true;
    }
    // reduce an arbitrary 256-bit little-endian integer mod p (values < 2^256 < 4p + ...)
    static Fe from_bytes_le_mod(const uint8_t *b) {
        Fe c;
        std::memcpy(c.l, b, 32);
        Fe p = modulus_raw();
        while (geq_raw(c, p)) sub_raw(c, c, p);
        return from_canonical(c);
    }
    void to_bytes_le(uint8_t *b) const {
        Fe c = to_canonical();
        std::memcpy(b, c.l, 32);
    }

    Fe pow(const uint64_t e[4]) const {
        Fe acc = one();
        bool started = false;
        for (int i = 255; i >= 0; i--) {
            if (started) acc = acc.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) {
                acc = acc * *this;
                started = true;
            }
        }
        return acc;
    }
    Fe pow_u64(uint64_t e) const {
        uint64_t ee[4] = {e, 0, 0, 0};
        return pow(ee);
    }
    Fe inv() const {
        const uint64_t e[4] = {F::MOD_MINUS_2_64(0), F::MOD_MINUS_2_64(1), F::MOD_MINUS_2_64(2), F::MOD_MINUS_2_64(3)};
        return pow(e);
    }
    // The square root ark-ff 0.3 returns (Tonelli-Shanks with the 2^32-th root of unity 5^t).
    bool sqrt(Fe &out) const {
        if (is_zero()) {
            out = *this;
            return true;
        }
        const uint64_t half[4] = {F::HALF_64(0), F::HALF_64(1), F::HALF_64(2), F::HALF_64(3)};
        if (pow(half) != one()) return false;
        const uint64_t tm[4] = {F::T_MINUS1_DIV2_64(0), F::T_MINUS1_DIV2_64(1), F::T_MINUS1_DIV2_64(2), F::T_MINUS1_DIV2_64(3)};
        Fe z{{F::ROOT_OF_UNITY_64(0), F::ROOT_OF_UNITY_64(1), F::ROOT_OF_UNITY_64(2), F::ROOT_OF_UNITY_64(3)}};
        Fe w = pow(tm);
        Fe x = *this * w;
        Fe b = x * w;
        int v = 32;
        while (b != one()) {
            int k = 0;
            Fe b2k = b;
            while (b2k != one()) {
                b2k = b2k.sqr();
                k++;
            }
            int j = v - k - 1;
            w = z;
            for (int i = 0; i < j; i++) w = w.sqr();
            z = w.sqr();
            b = b * z;
            x = x * w;
            v = k;
        }
        out = x;
        return true;
    }
    // canonical integer comparison helper: is the canonical value > (p-1)/2 ?
    bool is_lexicographically_large() const {
        Fe c = to_canonical();
        Fe h{{F::HALF_64(0), F::HALF_64(1), F::HALF_64(2), F::HALF_64(3)}};
        return !geq_raw(h, c);
    }
};

using Fp = Fe<FpParams>;
using Fq = Fe<FqParams>;

// Affine point on y^2 = x^3 + 5 over the base field B.
template <class B>
struct Affine {
    Fe<B> x, y;
    bool inf = false;
    static Affine identity() {
        Affine a;
        a.x = Fe<B>::zero();
        a.y = Fe<B>::zero();
        a.inf = true;
        return a;
    }
    bool on_curve() const {
        if (inf) return true;
        return y.sqr() == x.sqr() * x + Fe<B>::from_u64(5);
    }
    Affine neg() const {
        Affine r = *this;
        r.y = -y;
        return r;
    }
    bool operator==(const Affine &o) const { return inf == o.inf && (inf || (x == o.x && y == o.y)); }
};

// Jacobian arithmetic for the handful of host-side group operations (small MSMs stay on the GPU).
template <class B>
struct Jac {
    Fe<B> x, y, z;
    static Jac identity() { return Jac{Fe<B>::one(), Fe<B>::one(), Fe<B>::zero()}; }
    static Jac from_affine(const Affine<B> &a) { return a.inf ? identity() : Jac{a.x, a.y, Fe<B>::one()}; }
    bool is_identity() const { return z.is_zero(); }
    Jac dbl() const {
        if (is_identity()) return *this;
        Fe<B> A = x.sqr(), Bq = y.sqr(), C = Bq.sqr();
        Fe<B> D = ((x + Bq).sqr() - A - C).dbl();
        Fe<B> E = A.dbl() + A, Fv = E.sqr();
        Jac r;
        r.x = Fv - D.dbl();
        r.y = E * (D - r.x) - C.dbl().dbl().dbl();
        r.z = (y * z).dbl();
        return r;
    }
    Jac add(const Jac &q) const {
        if (is_identity()) return q;
        if (q.is_identity()) return *this;
        Fe<B> z1z1 = z.sqr(), z2z2 = q.z.sqr();
        Fe<B> u1 = x * z2z2, u2 = q.x * z1z1;
        Fe<B> s1 = y * q.z * z2z2, s2 = q.y * z * z1z1;
        if (u1 == u2) {
            if (s1 == s2) return dbl();
            return identity();
        }
        Fe<B> h = u2 - u1, i = h.dbl().sqr(), j = h * i, r = (s2 - s1).dbl(), v = u1 * i;
        Jac o;
        o.x = r.sqr() - j - v.dbl();
        o.y = r * (v - o.x) - (s1 * j).dbl();
        o.z = ((z + q.z).sqr() - z1z1 - z2z2) * h;
        return o;
    }
    Jac add_affine(const Affine<B> &q) const { return add(from_affine(q)); }
    Affine<B> to_affine() const {
        if (is_identity()) return Affine<B>::identity();
        Fe<B> zi = z.inv(), zi2 = zi.sqr();
        Affine<B> a;
        a.x = x * zi2;
        a.y = y * zi2 * zi;
        a.inf = false;
        return a;
    }
    // scalar given as canonical little-endian 4 x u64
    static Jac mul(const Affine<B> &p, const uint64_t k[4]) {
        Jac acc = identity();
        for (int i = 255; i >= 0; i--) {
            acc = acc.dbl();
            if ((k[i >> 6] >> (i & 63)) & 1) acc = acc.add_affine(p);
        }
        return acc;
    }
};

}  // namespace host
}  // namespace pasta
