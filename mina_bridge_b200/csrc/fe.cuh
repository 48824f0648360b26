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

#ifdef __CUDACC__
// Values ptxas must not see through.  -p^-1 mod 2^32 is 0xffffffff for both Pasta primes, so the
// Montgomery quotient digit is just -v; written as `0 - v`, ptxas rewrites m*t as v*(-t) for the low
// halves only, which splits every lo/hi pair and blocks the IMAD.WIDE fusion mul_ptx2 is built on.
// Reading the factor from constant memory costs one IMAD instead of one IADD3 and keeps the pairs.
// The zero is used to order the product before the reduction (see mul_ptx2).
static __constant__ uint32_t PASTA_NINV32 = 0xffffffffu;
static __constant__ uint32_t PASTA_ZERO = 0u;
#endif

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

    PASTA_HD static fe add_portable(const fe &a, const fe &b) {
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

    PASTA_HD static fe sub_portable(const fe &a, const fe &b) {
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

#ifdef __CUDA_ARCH__
    // ---- sm_100a path ---------------------------------------------------------------------------
    // acc[0..8] += (a0, a1, a2, a3) * b: four 64-bit products on aligned register pairs, one carry
    // chain (each mad.lo.cc/madc.hi.cc pair fuses into one IMAD.WIDE.U32.X), carry into acc[8].
    static __device__ __forceinline__ void mad_row(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                   uint32_t b) {
        asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
              "+r"(acc[7]), "+r"(acc[8])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
    }
    // same, into fresh registers (acc[0..8) = products, no carry limb)
    static __device__ __forceinline__ void mul_row(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                   uint32_t b) {
        asm("mul.lo.u32 %0, %8, %12;\n\t"
            "mul.hi.u32 %1, %8, %12;\n\t"
            "mul.lo.u32 %2, %9, %12;\n\t"
            "mul.hi.u32 %3, %9, %12;\n\t"
            "mul.lo.u32 %4, %10, %12;\n\t"
            "mul.hi.u32 %5, %10, %12;\n\t"
            "mul.lo.u32 %6, %11, %12;\n\t"
            "mul.hi.u32 %7, %11, %12;"
            : "=&r"(acc[0]), "=&r"(acc[1]), "=&r"(acc[2]), "=&r"(acc[3]), "=&r"(acc[4]), "=&r"(acc[5]), "=&r"(acc[6]),
              "=&r"(acc[7])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
    }

    // One Montgomery round on the window U[i..i+4]:  m = -U[i];  U += m * (1, t1, t2, t3) << 32i.
    // `d` carries the deferred carry that belongs to limb i+4 in and the one for limb i+5 out.
    static __device__ __forceinline__ uint32_t redc_round(uint32_t u0, uint32_t &u1, uint32_t &u2, uint32_t &u3,
                                                          uint32_t &u4, uint32_t &d) {
        uint32_t m, scratch;
        asm("sub.u32 %0, 0, %7;\n\t"
            "add.cc.u32 %1, %7, %0;\n\t"       // CF = (u0 != 0)
            "madc.lo.cc.u32 %2, %0, %8, %2;\n\t"
            "madc.lo.cc.u32 %3, %0, %9, %3;\n\t"
            "madc.lo.cc.u32 %4, %0, %10, %4;\n\t"
            "addc.cc.u32 %5, %5, %6;\n\t"
            "addc.u32 %6, 0, 0;\n\t"
            "mad.hi.cc.u32 %3, %0, %8, %3;\n\t"
            "madc.hi.cc.u32 %4, %0, %9, %4;\n\t"
            "madc.hi.cc.u32 %5, %0, %10, %5;\n\t"
            "addc.u32 %6, %6, 0;"
            : "=&r"(m), "=&r"(scratch), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(d)
            : "r"(u0), "r"(F::MOD(1)), "r"(F::MOD(2)), "r"(F::MOD(3)));
        return m;
    }

    static __device__ __forceinline__ fe cond_sub_p(const uint32_t *r) {
        // r < 2p < 2^256: subtract p once if r >= p
        uint32_t d[8], borrow;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=&r"(d[0]), "=&r"(d[1]), "=&r"(d[2]), "=&r"(d[3]), "=&r"(d[4]), "=&r"(d[5]), "=&r"(d[6]), "=&r"(d[7]),
              "=&r"(borrow)
            : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
              "r"(F::MOD(0)), "r"(F::MOD(1)), "r"(F::MOD(2)), "r"(F::MOD(3)), "r"(F::MOD(4)), "r"(F::MOD(5)),
              "r"(F::MOD(6)), "r"(F::MOD(7)));
        fe o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.v[i] = borrow ? r[i] : d[i];
        return o;
    }

    static __device__ __forceinline__ fe mul_ptx(const fe &a, const fe &b) {
        uint32_t U[16];
        mul_wide(U, a, b);
        return redc(U);
    }
    static __device__ __forceinline__ void mul_wide(uint32_t *U, const fe &a, const fe &b) {
        // 512-bit product: X collects a_j*b_i with i+j even, Y those with i+j odd (stored one limb down)
        uint32_t X[17], Y[17];
#pragma unroll
        for (int i = 8; i < 17; i++) X[i] = Y[i] = 0;
        mul_row(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
        mul_row(Y, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            if (i & 1) {
                mad_row(&Y[i - 1], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
                mad_row(&X[i + 1], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
            } else {
                mad_row(&X[i], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
                mad_row(&Y[i], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
            }
        }
        U[0] = X[0];
        asm("add.cc.u32 %0, %15, %30;\n\t"
            "addc.cc.u32 %1, %16, %31;\n\t"
            "addc.cc.u32 %2, %17, %32;\n\t"
            "addc.cc.u32 %3, %18, %33;\n\t"
            "addc.cc.u32 %4, %19, %34;\n\t"
            "addc.cc.u32 %5, %20, %35;\n\t"
            "addc.cc.u32 %6, %21, %36;\n\t"
            "addc.cc.u32 %7, %22, %37;\n\t"
            "addc.cc.u32 %8, %23, %38;\n\t"
            "addc.cc.u32 %9, %24, %39;\n\t"
            "addc.cc.u32 %10, %25, %40;\n\t"
            "addc.cc.u32 %11, %26, %41;\n\t"
            "addc.cc.u32 %12, %27, %42;\n\t"
            "addc.cc.u32 %13, %28, %43;\n\t"
            "addc.u32 %14, %29, %44;"
            : "=&r"(U[1]), "=&r"(U[2]), "=&r"(U[3]), "=&r"(U[4]), "=&r"(U[5]), "=&r"(U[6]), "=&r"(U[7]), "=&r"(U[8]),
              "=&r"(U[9]), "=&r"(U[10]), "=&r"(U[11]), "=&r"(U[12]), "=&r"(U[13]), "=&r"(U[14]), "=&r"(U[15])
            : "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]), "r"(X[8]), "r"(X[9]),
              "r"(X[10]), "r"(X[11]), "r"(X[12]), "r"(X[13]), "r"(X[14]), "r"(X[15]), "r"(Y[0]), "r"(Y[1]),
              "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]), "r"(Y[8]), "r"(Y[9]), "r"(Y[10]),
              "r"(Y[11]), "r"(Y[12]), "r"(Y[13]), "r"(Y[14]));
    }

    // Montgomery reduction of a 512-bit value U < p * 2^256 (consumes U).
    static __device__ __forceinline__ fe redc(uint32_t *U) {
        uint32_t m[8], d = 0, d7;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            m[i] = redc_round(U[i], U[i + 1], U[i + 2], U[i + 3], U[i + 4], d);
            if (i == 0) {
                // The 2^254 part of p: (m0 << 30) lands in limb 7 and must be there before round 7.
                // The carry is taken by comparison on purpose: with add.cc here, ptxas 12.9 fuses
                // neg + shl + add.cc into `LEA RZ, P, -R, R, 0x1e`, whose carry-out is wrong
                // whenever m0 = 0 (mod 4).
                uint32_t s30 = m[0] << 30;
                U[7] += s30;
                d7 = U[7] < s30 ? 1u : 0u;
            }
        }
        // limbs 0..7 are now zero.  result = U[8..16) + (m >> 2) + d7 (limb 8) + d (limb 12)
        uint32_t sh[8];
#pragma unroll
        for (int k = 0; k < 7; k++) sh[k] = __funnelshift_r(m[k], m[k + 1], 2);
        sh[7] = m[7] >> 2;
        uint32_t r[8];
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;\n\t"
            "add.cc.u32 %0, %0, %24;\n\t"
            "addc.cc.u32 %1, %1, 0;\n\t"
            "addc.cc.u32 %2, %2, 0;\n\t"
            "addc.cc.u32 %3, %3, 0;\n\t"
            "addc.cc.u32 %4, %4, %25;\n\t"
            "addc.cc.u32 %5, %5, 0;\n\t"
            "addc.cc.u32 %6, %6, 0;\n\t"
            "addc.u32 %7, %7, 0;"
            : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
            : "r"(U[8]), "r"(U[9]), "r"(U[10]), "r"(U[11]), "r"(U[12]), "r"(U[13]), "r"(U[14]), "r"(U[15]),
              "r"(sh[0]), "r"(sh[1]), "r"(sh[2]), "r"(sh[3]), "r"(sh[4]), "r"(sh[5]), "r"(sh[6]), "r"(sh[7]),
              "r"(d7), "r"(d));
        return cond_sub_p(r);
    }

    // ---- second-generation multiplication: reduction on IMAD.WIDE carry chains -------------------------
    // The product is left in its two parity accumulators (X: columns with i+j even, Y: odd, stored one
    // limb down; true value = X + (Y << 32)) and the eight Montgomery rounds add m_i * (t1, t2, t3)
    // straight into them: m_i*t1 and m_i*t3 land on register pairs of one accumulator, m_i*t2 on a pair
    // of the other, so every product is ONE IMAD.WIDE.U32(.X) with carry-in/out instead of an IMAD +
    // IADD3.X + IMAD.HI + IADD3.X quartet.  Only limb i itself is merged per round (to get m_i); carry
    // bits that fall off a chain are counted in K[limb] and folded back in when that limb is merged.
    static __device__ __forceinline__ void redc_chain4(uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t &a3, uint32_t &k,
                                                       uint32_t m) {
        asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
            "madc.lo.cc.u32 %2, %5, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %5, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(k)
            : "r"(m), "r"(F::MOD(1)), "r"(F::MOD(3)));
    }
    static __device__ __forceinline__ void redc_chain2(uint32_t &b0, uint32_t &b1, uint32_t &k, uint32_t m) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, %2, 0;"
            : "+r"(b0), "+r"(b1), "+r"(k)
            : "r"(m), "r"(F::MOD(2)));
    }
    // Reduction of a product held in the two parity accumulators (true value = X + (Y << 32), X[16] and
    // Y[14] are the top carry limbs).  Consumes X and Y.
    static __device__ __forceinline__ fe redc2(uint32_t *X, uint32_t *Y) {
        // Ordering only: without this data dependency ptxas interleaves the reduction rounds with the
        // product rows, keeps a dozen carry chains alive at once and spills predicates (35 P2R + 43 ISETP
        // per multiplication measured); with it at most three chains are live.
        X[0] += (X[16] | Y[14]) & PASTA_ZERO;
        uint32_t K[13], m[8], c = 0, d7 = 0;
#pragma unroll
        for (int i = 0; i < 13; i++) K[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i == 7) {
                // the 2^254 part of p: (m0 << 30) belongs to limb 7 and must be there before it is merged;
                // carry by comparison (ptxas 12.9 mis-fuses neg + shl + add.cc into LEA, see redc())
                uint32_t s30 = m[0] << 30;
                X[7] += s30;
                d7 = X[7] < s30 ? 1u : 0u;
            }
            uint64_t sum = (uint64_t)X[i] + (uint64_t)(i ? Y[i - 1] : 0u) + (uint64_t)(c + K[i]);
            uint32_t v = (uint32_t)sum, scratch;
            m[i] = v * PASTA_NINV32;  // = -v
            // limb i becomes v + m_i = 0 or 2^32: the carry of that addition is (v != 0)
            asm("add.cc.u32 %0, %2, %3;\n\t"
                "addc.u32 %1, %4, 0;"
                : "=&r"(scratch), "=r"(c)
                : "r"(v), "r"(m[i]), "r"((uint32_t)(sum >> 32)));
            if (i & 1) {
                redc_chain4(X[i + 1], X[i + 2], X[i + 3], X[i + 4], K[i + 5], m[i]);  // true limbs i+1 .. i+4
                redc_chain2(Y[i + 1], Y[i + 2], K[i + 4], m[i]);                      // true limbs i+2, i+3
            } else {
                redc_chain4(Y[i], Y[i + 1], Y[i + 2], Y[i + 3], K[i + 5], m[i]);
                redc_chain2(X[i + 2], X[i + 3], K[i + 4], m[i]);
            }
        }
        // true limbs 8..15 = X[8..16) + Y[7..15) + (m >> 2 funnel) + pending carries
        uint32_t sh[8];
#pragma unroll
        for (int k = 0; k < 7; k++) sh[k] = __funnelshift_r(m[k], m[k + 1], 2);
        sh[7] = m[7] >> 2;
        uint32_t k8 = c + d7 + K[8];
        uint32_t r[8];
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
            : "r"(X[8]), "r"(X[9]), "r"(X[10]), "r"(X[11]), "r"(X[12]), "r"(X[13]), "r"(X[14]), "r"(X[15]),
              "r"(Y[7]), "r"(Y[8]), "r"(Y[9]), "r"(Y[10]), "r"(Y[11]), "r"(Y[12]), "r"(Y[13]), "r"(Y[14]));
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32 %7, %7, %15;\n\t"
            "add.cc.u32 %0, %0, %16;\n\t"
            "addc.cc.u32 %1, %1, %17;\n\t"
            "addc.cc.u32 %2, %2, %18;\n\t"
            "addc.cc.u32 %3, %3, %19;\n\t"
            "addc.cc.u32 %4, %4, %20;\n\t"
            "addc.cc.u32 %5, %5, 0;\n\t"
            "addc.cc.u32 %6, %6, 0;\n\t"
            "addc.u32 %7, %7, 0;"
            : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
            : "r"(sh[0]), "r"(sh[1]), "r"(sh[2]), "r"(sh[3]), "r"(sh[4]), "r"(sh[5]), "r"(sh[6]), "r"(sh[7]),
              "r"(k8), "r"(K[9]), "r"(K[10]), "r"(K[11]), "r"(K[12]));
        return cond_sub_p(r);
    }

    // the 512-bit product in the two parity accumulators (no reduction)
    static __device__ __forceinline__ void mul_xy(uint32_t *X, uint32_t *Y, const fe &a, const fe &b) {
#pragma unroll
        for (int i = 8; i < 17; i++) X[i] = Y[i] = 0;
        mul_row(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
        mul_row(Y, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            if (i & 1) {
                mad_row(&Y[i - 1], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
                mad_row(&X[i + 1], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
            } else {
                mad_row(&X[i], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
                mad_row(&Y[i], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
            }
        }
    }
    static __device__ __forceinline__ fe mul_ptx2(const fe &a, const fe &b) {
        uint32_t X[17], Y[17];
        mul_xy(X, Y, a, b);
        return redc2(X, Y);
    }
    // ---- lazy reduction: a sum of up to THREE products reduced once -------------------------------------------------
    // p is 2^254 + t, so p^2 / 2^256 < p / 4 + eps: the Montgomery reduction of a sum of k products is below
    // (k / 4 + 1) p, i.e. still under 2 p for k <= 3 and the single conditional subtraction at the end of redc2 is
    // enough.  The reduction is 26 of the 100 multiply-pipe instructions of a product, so sums of products
    // (k_bpoly_combine) save about a quarter of them.  (X, Y) += (X1, Y1): two full carry chains.
    static __device__ __forceinline__ void xy_add(uint32_t *X, uint32_t *Y, const uint32_t *X1, const uint32_t *Y1) {
        asm("add.cc.u32 %0, %0, %17;\n\t"
            "addc.cc.u32 %1, %1, %18;\n\t"
            "addc.cc.u32 %2, %2, %19;\n\t"
            "addc.cc.u32 %3, %3, %20;\n\t"
            "addc.cc.u32 %4, %4, %21;\n\t"
            "addc.cc.u32 %5, %5, %22;\n\t"
            "addc.cc.u32 %6, %6, %23;\n\t"
            "addc.cc.u32 %7, %7, %24;\n\t"
            "addc.cc.u32 %8, %8, %25;\n\t"
            "addc.cc.u32 %9, %9, %26;\n\t"
            "addc.cc.u32 %10, %10, %27;\n\t"
            "addc.cc.u32 %11, %11, %28;\n\t"
            "addc.cc.u32 %12, %12, %29;\n\t"
            "addc.cc.u32 %13, %13, %30;\n\t"
            "addc.cc.u32 %14, %14, %31;\n\t"
            "addc.cc.u32 %15, %15, %32;\n\t"
            "addc.u32 %16, %16, %33;"
            : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8]), "+r"(X[9]),
              "+r"(X[10]), "+r"(X[11]), "+r"(X[12]), "+r"(X[13]), "+r"(X[14]), "+r"(X[15]), "+r"(X[16])
            : "r"(X1[0]), "r"(X1[1]), "r"(X1[2]), "r"(X1[3]), "r"(X1[4]), "r"(X1[5]), "r"(X1[6]), "r"(X1[7]), "r"(X1[8]), "r"(X1[9]),
              "r"(X1[10]), "r"(X1[11]), "r"(X1[12]), "r"(X1[13]), "r"(X1[14]), "r"(X1[15]), "r"(X1[16]));
        asm("add.cc.u32 %0, %0, %15;\n\t"
            "addc.cc.u32 %1, %1, %16;\n\t"
            "addc.cc.u32 %2, %2, %17;\n\t"
            "addc.cc.u32 %3, %3, %18;\n\t"
            "addc.cc.u32 %4, %4, %19;\n\t"
            "addc.cc.u32 %5, %5, %20;\n\t"
            "addc.cc.u32 %6, %6, %21;\n\t"
            "addc.cc.u32 %7, %7, %22;\n\t"
            "addc.cc.u32 %8, %8, %23;\n\t"
            "addc.cc.u32 %9, %9, %24;\n\t"
            "addc.cc.u32 %10, %10, %25;\n\t"
            "addc.cc.u32 %11, %11, %26;\n\t"
            "addc.cc.u32 %12, %12, %27;\n\t"
            "addc.cc.u32 %13, %13, %28;\n\t"
            "addc.u32 %14, %14, %29;"
            : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7]), "+r"(Y[8]), "+r"(Y[9]),
              "+r"(Y[10]), "+r"(Y[11]), "+r"(Y[12]), "+r"(Y[13]), "+r"(Y[14])
            : "r"(Y1[0]), "r"(Y1[1]), "r"(Y1[2]), "r"(Y1[3]), "r"(Y1[4]), "r"(Y1[5]), "r"(Y1[6]), "r"(Y1[7]), "r"(Y1[8]), "r"(Y1[9]),
              "r"(Y1[10]), "r"(Y1[11]), "r"(Y1[12]), "r"(Y1[13]), "r"(Y1[14]));
    }
    // a0 b0 + a1 b1 (+ a2 b2) in Montgomery form; pass n = 2 or 3
    static __device__ __forceinline__ fe dot_ptx2(const fe &a0, const fe &b0, const fe &a1, const fe &b1, const fe &a2, const fe &b2, int n) {
        uint32_t X[17], Y[17], X1[17], Y1[17];
        mul_xy(X, Y, a0, b0);
        mul_xy(X1, Y1, a1, b1);
        xy_add(X, Y, X1, Y1);
        if (n == 3) {
            mul_xy(X1, Y1, a2, b2);
            xy_add(X, Y, X1, Y1);
        }
        return redc2(X, Y);
    }

    // acc[0..4) += (a0, a1) * b, carry into acc[4]; acc[0..6) += (a0, a1, a2) * b, carry into acc[6]
    static __device__ __forceinline__ void mad_row2(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t b) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
            : "r"(a0), "r"(a1), "r"(b));
    }
    static __device__ __forceinline__ void mad_row3(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
            "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
            "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32 %6, %6, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
            : "r"(a0), "r"(a1), "r"(a2), "r"(b));
    }
    // Dedicated squaring: the 28 off-diagonal products a_i*a_j (i < j) are computed once and doubled
    // (36 IMAD.WIDE for the product instead of 64; the doubling is 30 funnel shifts on the ALU pipe,
    // which has slack while the multiply pipe is the bottleneck).  Same parity accumulators as
    // mul_ptx2: Y collects the 16 (even, odd) pairs, X the 6 (even, even) pairs, Z the 6 (odd, odd)
    // pairs (rows must move upward so a row's carry limb is always untouched above; the two X-parity
    // families interleave, hence the third accumulator, merged with one add chain).
    static __device__ __forceinline__ fe sqr_ptx2(const fe &a) {
        uint32_t X[17], Y[17], Z[15];
#pragma unroll
        for (int i = 0; i < 17; i++) X[i] = Y[i] = 0;
#pragma unroll
        for (int i = 0; i < 15; i++) Z[i] = 0;
        // (even, odd) pairs: true limb e + o = Y index e + o - 1
        mul_row(Y, a.v[0], a.v[2], a.v[4], a.v[6], a.v[1]);
        mad_row(&Y[2], a.v[0], a.v[2], a.v[4], a.v[6], a.v[3]);
        mad_row(&Y[4], a.v[0], a.v[2], a.v[4], a.v[6], a.v[5]);
        mad_row(&Y[6], a.v[0], a.v[2], a.v[4], a.v[6], a.v[7]);
        // (even, even) pairs into X, (odd, odd) pairs into Z
        {
            uint64_t p = (uint64_t)a.v[0] * a.v[2];
            X[2] = (uint32_t)p;
            X[3] = (uint32_t)(p >> 32);
            uint64_t q = (uint64_t)a.v[1] * a.v[3];
            Z[4] = (uint32_t)q;
            Z[5] = (uint32_t)(q >> 32);
        }
        mad_row2(&X[4], a.v[0], a.v[2], a.v[4]);
        mad_row3(&X[6], a.v[0], a.v[2], a.v[4], a.v[6]);
        mad_row2(&Z[6], a.v[1], a.v[3], a.v[5]);
        mad_row3(&Z[8], a.v[1], a.v[3], a.v[5], a.v[7]);
        asm("add.cc.u32 %0, %0, %12;\n\t"
            "addc.cc.u32 %1, %1, %13;\n\t"
            "addc.cc.u32 %2, %2, %14;\n\t"
            "addc.cc.u32 %3, %3, %15;\n\t"
            "addc.cc.u32 %4, %4, %16;\n\t"
            "addc.cc.u32 %5, %5, %17;\n\t"
            "addc.cc.u32 %6, %6, %18;\n\t"
            "addc.cc.u32 %7, %7, %19;\n\t"
            "addc.cc.u32 %8, %8, %20;\n\t"
            "addc.cc.u32 %9, %9, %21;\n\t"
            "addc.cc.u32 %10, %10, %22;\n\t"
            "addc.u32 %11, %11, 0;"
            : "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8]), "+r"(X[9]), "+r"(X[10]), "+r"(X[11]), "+r"(X[12]),
              "+r"(X[13]), "+r"(X[14]), "+r"(X[15])
            : "r"(Z[4]), "r"(Z[5]), "r"(Z[6]), "r"(Z[7]), "r"(Z[8]), "r"(Z[9]), "r"(Z[10]), "r"(Z[11]), "r"(Z[12]), "r"(Z[13]),
              "r"(Z[14]));
        // double the off-diagonal sums
#pragma unroll
        for (int k = 15; k >= 3; k--) X[k] = __funnelshift_l(X[k - 1], X[k], 1);
        X[2] <<= 1;
#pragma unroll
        for (int k = 15; k >= 1; k--) Y[k] = __funnelshift_l(Y[k - 1], Y[k], 1);
        Y[0] <<= 1;
        // diagonal a_i^2 at true limbs 2i, 2i+1: one 16-limb chain
        asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
            "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
            "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
            "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
            "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
            "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
            "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
            "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
            "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
            "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
            "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
            "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
            "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
            "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
            "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
            "madc.hi.u32 %15, %23, %23, %15;"
            : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8]), "+r"(X[9]),
              "+r"(X[10]), "+r"(X[11]), "+r"(X[12]), "+r"(X[13]), "+r"(X[14]), "+r"(X[15])
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
        return redc2(X, Y);
    }

    static __device__ __forceinline__ fe add_ptx(const fe &a, const fe &b) {
        uint32_t r[8];
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
              "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
        return cond_sub_p(r);
    }
    static __device__ __forceinline__ fe sub_ptx(const fe &a, const fe &b) {
        uint32_t d[8], borrow;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=&r"(d[0]), "=&r"(d[1]), "=&r"(d[2]), "=&r"(d[3]), "=&r"(d[4]), "=&r"(d[5]), "=&r"(d[6]), "=&r"(d[7]),
              "=&r"(borrow)
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
              "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
        // borrow is 0 or 0xffffffff: add p back under the mask
        fe o;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=&r"(o.v[0]), "=&r"(o.v[1]), "=&r"(o.v[2]), "=&r"(o.v[3]), "=&r"(o.v[4]), "=&r"(o.v[5]), "=&r"(o.v[6]),
              "=&r"(o.v[7])
            : "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
              "r"(F::MOD(0) & borrow), "r"(F::MOD(1) & borrow), "r"(F::MOD(2) & borrow), "r"(F::MOD(3) & borrow),
              "r"(F::MOD(4) & borrow), "r"(F::MOD(5) & borrow), "r"(F::MOD(6) & borrow), "r"(F::MOD(7) & borrow));
        return o;
    }
#endif

    PASTA_HD static fe mul(const fe &a, const fe &b) {
#ifdef __CUDA_ARCH__
        return mul_ptx2(a, b);
#else
        return mul_portable(a, b);
#endif
    }
    // a0 b0 + a1 b1 + a2 b2 with one reduction (device: lazy reduction; host: three products)
    PASTA_HD static fe dot3(const fe &a0, const fe &b0, const fe &a1, const fe &b1, const fe &a2, const fe &b2) {
#ifdef __CUDA_ARCH__
        return dot_ptx2(a0, b0, a1, b1, a2, b2, 3);
#else
        return add_portable(add_portable(mul_portable(a0, b0), mul_portable(a1, b1)), mul_portable(a2, b2));
#endif
    }
    PASTA_HD static fe dot2(const fe &a0, const fe &b0, const fe &a1, const fe &b1) {
#ifdef __CUDA_ARCH__
        return dot_ptx2(a0, b0, a1, b1, a1, b1, 2);
#else
        return add_portable(mul_portable(a0, b0), mul_portable(a1, b1));
#endif
    }
    PASTA_HD static fe sqr(const fe &a) {
#ifdef __CUDA_ARCH__
        return sqr_ptx2(a);
#else
        return mul_portable(a, a);
#endif
    }
    PASTA_HD static fe add(const fe &a, const fe &b) {
#ifdef __CUDA_ARCH__
        return add_ptx(a, b);
#else
        return add_portable(a, b);
#endif
    }
    PASTA_HD static fe sub(const fe &a, const fe &b) {
#ifdef __CUDA_ARCH__
        return sub_ptx(a, b);
#else
        return sub_portable(a, b);
#endif
    }
    PASTA_HD static fe dbl(const fe &a) { return add(a, a); }
    PASTA_HD static fe neg(const fe &a) { return sub(fe_zero(), a); }

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
