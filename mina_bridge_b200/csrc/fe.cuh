// Pasta field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// Replaces (on the GPU) what the reference gets from ark-ff 0.3 `Fp256` with the x86 `asm` feature
// (AL/operator/mina/lib/Cargo.toml:17-19; lambdaclass/openmina_algebra @ 017531e).  Values are
// always fully reduced to [0, p), so results are bit-identical to the CPU representation.
//
// Both Pasta moduli have the shape p = 2^254 + t with t < 2^126 and p = 1 (mod 2^32):
//   * the Montgomery quotient digit is simply -T[i] (no multiply), and
//   * p has only three "interesting" 32-bit limbs (1..3); limb 0 is 1 and limb 7 is 2^30.
// `mont_mul` exploits that: 64 multiply-adds for the product and 24 for the reduction.
#pragma once
#include <cstdint>
#include "pasta_params.h"

namespace pasta {

struct alignas(16) fe {
    uint32_t v[8];
};

PASTA_HD bool fe_is_zero(const fe &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
PASTA_HD bool fe_eq(const fe &a, const fe &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
PASTA_HD fe fe_zero() {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}

template <class F>
struct Fd {
    PASTA_HD static fe one() {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = F::R1(i);
        return r;
    }
    PASTA_HD static fe constant_r2() {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = F::R2(i);
        return r;
    }
    PASTA_HD static fe five() {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = F::FIVE(i);
        return r;
    }
    PASTA_HD static fe endo_r() {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = F::ENDO_R(i);
        return r;
    }

    // r = a - p if a >= p else a   (a < 2p assumed, `carry` = bit 256 of a)
    PASTA_HD static fe reduce_once(const fe &a, uint32_t carry = 0) {
        fe d;
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)a.v[i] - F::MOD(i) - borrow;
            d.v[i] = (uint32_t)t;
            borrow = (uint32_t)(t >> 63);
        }
        bool ge = carry || !borrow;
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = ge ? d.v[i] : a.v[i];
        return r;
    }

    PASTA_HD static fe add(const fe &a, const fe &b) {
        fe s;
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)a.v[i] + b.v[i] + c;
            s.v[i] = (uint32_t)t;
            c = (uint32_t)(t >> 32);
        }
        return reduce_once(s, c);  // p < 2^255 so c is always 0; kept for clarity
    }
    PASTA_HD static fe dbl(const fe &a) { return add(a, a); }

    PASTA_HD static fe sub(const fe &a, const fe &b) {
        fe d;
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)a.v[i] - b.v[i] - borrow;
            d.v[i] = (uint32_t)t;
            borrow = (uint32_t)(t >> 63);
        }
        uint32_t mask = 0u - borrow, c = 0;
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)d.v[i] + (F::MOD(i) & mask) + c;
            r.v[i] = (uint32_t)t;
            c = (uint32_t)(t >> 32);
        }
        return r;
    }
    PASTA_HD static fe neg(const fe &a) { return sub(fe_zero(), a); }

    // Portable CIOS Montgomery product (host and device); the reference implementation the
    // PTX path below is tested against.
    PASTA_HD static fe mul_portable(const fe &a, const fe &b) {
        uint32_t t[10];
#pragma unroll
        for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                c += (uint64_t)a.v[j] * b.v[i] + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[8] = (uint32_t)c;
            t[9] = (uint32_t)(c >> 32);
            uint32_t m = t[0] * F::NINV32;
            c = (uint64_t)m * F::MOD(0) + t[0];
            c >>= 32;
#pragma unroll
            for (int j = 1; j < 8; j++) {
                c += (uint64_t)m * F::MOD(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[7] = (uint32_t)c;
            t[8] = t[9] + (uint32_t)(c >> 32);
        }
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = t[i];
        return reduce_once(r, t[8]);
    }


    PASTA_HD static fe mul(const fe &a, const fe &b) { return mul_portable(a, b); }
    PASTA_HD static fe sqr(const fe &a) { return mul(a, a); }

    PASTA_HD static fe to_mont(const fe &a) { return mul(a, constant_r2()); }
    PASTA_HD static fe from_mont(const fe &a) {
        fe o = fe_zero();
        o.v[0] = 1;
        return mul(a, o);
    }

    // a^e for a 256-bit plain exponent given as 4 x u64 little-endian
    PASTA_HD static fe pow_u256(const fe &a, const uint64_t e[4]) {
        fe acc = one();
        for (int i = 255; i >= 0; i--) {
            acc = sqr(acc);
            if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    PASTA_HD static fe inv(const fe &a) {
        const uint64_t e[4] = {F::MOD_MINUS_2_64(0), F::MOD_MINUS_2_64(1), F::MOD_MINUS_2_64(2), F::MOD_MINUS_2_64(3)};
        return pow_u256(a, e);
    }
};

}  // namespace pasta
