// Kernels of the batched IPA final check: poly-commitment `SRS::verify`,
// `OpeningProof::challenges`, `combine_commitments`, `shift_scalar` and mina-poseidon's `DefaultFqSponge`
// (lambdaclass/openmina-proof-systems @ 44e0d3b, un-vendored; restated from the published algorithm).
// Reference call site: kimchi `verify` under `verify_block`, AL/operator/mina/lib/src/lib.rs:99-111.
//
// Per opening i the verifier equation is (rand_base -> rb_i, sg_rand_base -> sgrb_i, both drawn per opening here)
//     0 == sgrb_i * <s_i, G[0..2^k)>  +  B_i
//     B_i = (-rb z1 - sgrb) sg + rb (c cip - z1 b0) U + rb c sum_j (chal_j^-1 L_j + chal_j R_j)
//           + rb c sum_m polyscale^m C_m + rb delta - rb z2 H
// with U, chal_j, c from the Fq-sponge, b0 = sum_t evalscale^t b_poly(chal, pt_t), s_i = b_poly_coefficients(chal).
// The g side is exactly the accumulator relation, so the same group testing runs it (verifier.cu: rlc_levels
// with r_i = sgrb_i and P_i = -B_i); this file computes the transcript, the scalars and the B_i.
//
// PARITY: the arithmetic after the transcript is checked against an independent oracle prover + verifier
// (oracle/ipa.py); the transcript needs the kimchi Poseidon table, which is unavailable => UNPINNED against the
// reference (the sponge is table-driven like every other Poseidon user here).
#pragma once
#include <cuda_runtime.h>

#include "ec.cuh"
#include "poseidon.cuh"

namespace pasta {

// One opening per 4 lanes.  Inputs canonical; outputs: t (canonical base-field element for to_group) and k + 1
// 128-bit prechallenges (k rounds, then the one for c).
//   absorb_fr(shift_scalar(cip)); t = challenge_fq(); for each (L, R): absorb_g(L), absorb_g(R), challenge();
//   absorb_g(delta); challenge().
// SCALAR_LARGER: the curve's scalar modulus exceeds its base modulus (Pallas): shift_scalar = x - 2^255 and the
// element is absorbed as (x >> 1, x & 1); otherwise (Vesta) shift_scalar = (x - 2^255 - 1) / 2, absorbed whole.
template <class F, class S, bool SCALAR_LARGER>
__global__ void __launch_bounds__(128) k_ipa_transcript(const fe *__restrict__ state0, uint32_t mode, uint32_t count0,
                                                        const fe *__restrict__ cip, const fe *__restrict__ lr /* [n][k][2][2] */,
                                                        const fe *__restrict__ delta /* [n][2] */, uint32_t n, int k,
                                                        const fe *__restrict__ tab, fe *__restrict__ t_out, uint4 *__restrict__ pre_rounds /* [n][k] */,
                                                        uint4 *__restrict__ pre_c /* [n] */) {
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const uint32_t i = gid < n ? gid : n - 1;  // tail groups recompute the last opening (uniform control flow, no stores)
    LaneSponge<F> sp;
    sp.lane = threadIdx.x & 3u;
    sp.base = threadIdx.x & 28u;
    sp.tab = tab;
    sp.absorbing = mode == 0;
    sp.count = (int)count0;
    sp.st = Fd<F>::to_mont(state0[(size_t)3 * i + (sp.lane < 3 ? sp.lane : 0)]);
    // shift_scalar in the scalar field, then read the canonical integer as base-field element(s)
    fe x = Fd<S>::to_mont(cip[i]);
    fe two_pow = Fd<S>::one();
    for (int b = 0; b < 255; b++) two_pow = Fd<S>::dbl(two_pow);
    if (SCALAR_LARGER) {
        fe v = Fd<S>::from_mont(Fd<S>::sub(x, two_pow));
        fe hi, lo = fe_zero();
        lo.v[0] = v.v[0] & 1u;
#pragma unroll
        for (int w = 0; w < 8; w++) hi.v[w] = (v.v[w] >> 1) | (w < 7 ? v.v[w + 1] << 31 : 0u);
        sp.absorb(Fd<F>::to_mont(hi));
        sp.absorb(Fd<F>::to_mont(lo));
    } else {
        fe v = Fd<S>::from_mont(Fd<S>::sub(x, Fd<S>::add(two_pow, Fd<S>::one())));
        // divide by two: add the (odd) modulus when odd, then shift
        uint32_t carry = 0;
        if (v.v[0] & 1u) {
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint64_t sum = (uint64_t)v.v[w] + S::MOD(w) + carry;
                v.v[w] = (uint32_t)sum;
                carry = (uint32_t)(sum >> 32);
            }
        }
        fe h;
#pragma unroll
        for (int w = 0; w < 8; w++) h.v[w] = (v.v[w] >> 1) | ((w < 7 ? v.v[w + 1] : carry) << 31);
        sp.absorb(Fd<F>::to_mont(h));
    }
    fe t = sp.squeeze();
    if (gid < n && sp.lane == 0) t_out[i] = Fd<F>::from_mont(t);
    const fe *pts = lr + (size_t)i * k * 4;
    for (int j = 0; j <= k; j++) {
        if (j < k) {
            for (int q = 0; q < 4; q++) sp.absorb(Fd<F>::to_mont(pts[4 * j + q]));  // L.x, L.y, R.x, R.y
        } else {
            sp.absorb(Fd<F>::to_mont(delta[2 * (size_t)i]));
            sp.absorb(Fd<F>::to_mont(delta[2 * (size_t)i + 1]));
        }
        fe c = Fd<F>::from_mont(sp.squeeze());  // the two low limbs are the challenge
        if (gid < n && sp.lane == 0) (j < k ? pre_rounds[(size_t)i * k + j] : pre_c[i]) = make_uint4(c.v[0], c.v[1], c.v[2], c.v[3]);
    }
}

// ---- U = to_group(t): the Shallue-van de Woestijne map of groupmap `BWParameters` with u = 1 -----------
// Square root exactly as ark-ff 0.3 returns it (Tonelli-Shanks with the 2^32-th root of unity 5^t): the SIGN of
// y matters here because no file supplies it.  Same algorithm as host_field.hpp `Fe::sqrt`, which the derivation
// of the committed SRS pins.
template <class F>
__device__ bool fe_sqrt(const fe &a, fe &out) {
    if (fe_is_zero(a)) {
        out = a;
        return true;
    }
    const uint64_t half[4] = {F::HALF_64(0), F::HALF_64(1), F::HALF_64(2), F::HALF_64(3)};
    const fe one = Fd<F>::one();
    if (!fe_eq(Fd<F>::pow_u256(a, half), one)) return false;
    const uint64_t tm[4] = {F::T_MINUS1_DIV2_64(0), F::T_MINUS1_DIV2_64(1), F::T_MINUS1_DIV2_64(2), F::T_MINUS1_DIV2_64(3)};
    fe z;
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = F::ROOT_OF_UNITY(i);
    fe w = Fd<F>::pow_u256(a, tm);
    fe x = Fd<F>::mul(a, w);
    fe b = Fd<F>::mul(x, w);
    int v = 32;
    while (!fe_eq(b, one)) {
        int k = 0;
        fe b2k = b;
        while (!fe_eq(b2k, one)) {
            b2k = Fd<F>::sqr(b2k);
            k++;
        }
        w = z;
        for (int i = 0; i < v - k - 1; i++) w = Fd<F>::sqr(w);
        z = Fd<F>::sqr(w);
        b = Fd<F>::mul(b, z);
        x = Fd<F>::mul(x, w);
        v = k;
    }
    out = x;
    return true;
}
struct GroupMapConsts {  // Montgomery, computed once on the host (srs.hpp GroupMap)
    fe sqrt_neg_three, sqrt_neg_three_minus_one_over_two, inv_three;
};
// out[i * stride + slot] = to_group(t[i]) as a Montgomery affine point
template <class F>
__global__ void __launch_bounds__(64) k_to_group(const fe *__restrict__ t_can, uint32_t n, GroupMapConsts gm, affine *__restrict__ out,
                                                 uint32_t stride, uint32_t slot) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fe one = Fd<F>::one(), five = Fd<F>::five();
    const fe fu = Fd<F>::add(one, five);  // u^3 + b with u = 1
    fe t = Fd<F>::to_mont(t_can[i]);
    fe t2 = Fd<F>::sqr(t);
    fe t2_fu = Fd<F>::add(t2, fu);
    fe alpha_inv = Fd<F>::mul(t2_fu, t2);
    fe alpha = fe_is_zero(alpha_inv) ? alpha_inv : Fd<F>::inv(alpha_inv);
    fe xs[3];
    xs[0] = Fd<F>::sub(gm.sqrt_neg_three_minus_one_over_two, Fd<F>::mul(Fd<F>::mul(Fd<F>::sqr(t2), alpha), gm.sqrt_neg_three));
    xs[1] = Fd<F>::sub(Fd<F>::neg(one), xs[0]);
    xs[2] = Fd<F>::sub(one, Fd<F>::mul(Fd<F>::mul(Fd<F>::sqr(t2_fu), Fd<F>::mul(alpha, t2_fu)), gm.inv_three));
    affine p;
    p.x = fe_zero();
    p.y = fe_zero();
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        fe y;
        if (fe_sqrt<F>(Fd<F>::add(Fd<F>::mul(Fd<F>::sqr(xs[k]), xs[k]), five), y)) {
            p.x = xs[k];
            p.y = y;
            break;
        }
    }
    out[(size_t)i * stride + slot] = p;
}

// Scalars of the per-opening points, one thread per opening.  All inputs canonical except the challenges
// (Montgomery: k round challenges per opening, and c) and the two randomisers (Montgomery).  Output: canonical scalars in the order
//   sg, U, L_0, R_0, ..., L_{k-1}, R_{k-1}, C_0 .. C_{nc-1}, delta, H          (2k + nc + 4 per opening)
template <class S>
__global__ void __launch_bounds__(64) k_ipa_scalars(const fe *__restrict__ chal, const fe *__restrict__ chal_c, const fe *__restrict__ z1, const fe *__restrict__ z2,
                                                    const fe *__restrict__ cip, const fe *__restrict__ polyscale,
                                                    const fe *__restrict__ evalscale, const fe *__restrict__ elm /* [n][npts] */,
                                                    const fe *__restrict__ rb, const fe *__restrict__ sgrb, uint32_t n, int k, uint32_t nc,
                                                    uint32_t npts, fe *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fe *ch = chal + (size_t)i * k;
    const fe c = chal_c[i];
    // b0 = sum_t evalscale^t * b_poly(chal, pt_t);  b_poly(chal, x) = prod_i (1 + chal[i] x^(2^(k-1-i)))
    const fe es = Fd<S>::to_mont(evalscale[i]);
    fe b0 = fe_zero(), scale = Fd<S>::one();
    for (uint32_t t = 0; t < npts; t++) {
        fe pw = Fd<S>::to_mont(elm[(size_t)i * npts + t]);
        fe acc = Fd<S>::one();
        for (int r = k - 1; r >= 0; r--) {
            acc = Fd<S>::mul(acc, Fd<S>::add(Fd<S>::one(), Fd<S>::mul(ch[r], pw)));
            pw = Fd<S>::sqr(pw);
        }
        b0 = Fd<S>::add(b0, Fd<S>::mul(scale, acc));
        scale = Fd<S>::mul(scale, es);
    }
    const fe Z1 = Fd<S>::to_mont(z1[i]), Z2 = Fd<S>::to_mont(z2[i]), CIP = Fd<S>::to_mont(cip[i]);
    const fe RB = rb[i], SG = sgrb[i];
    const fe rbc = Fd<S>::mul(RB, c);
    const fe rbz1 = Fd<S>::mul(RB, Z1);
    fe *o = out + (size_t)i * (2 * k + nc + 4);
    o[0] = Fd<S>::from_mont(Fd<S>::neg(Fd<S>::add(rbz1, SG)));
    o[1] = Fd<S>::from_mont(Fd<S>::sub(Fd<S>::mul(rbc, CIP), Fd<S>::mul(rbz1, b0)));
    // inverses of the round challenges with one inversion: inv_all = 1 / prod; walk back
    fe prod = Fd<S>::one();
    for (int r = 0; r < k; r++) prod = Fd<S>::mul(prod, ch[r]);
    fe inv_run = Fd<S>::inv(prod);  // a zero challenge gives 0 here; arkworks' batch_inversion leaves zeros too
    for (int r = k - 1; r >= 0; r--) {
        // prefix = prod_{q<r} ch[q]
        fe prefix = Fd<S>::one();
        for (int q = 0; q < r; q++) prefix = Fd<S>::mul(prefix, ch[q]);
        fe inv_r = Fd<S>::mul(inv_run, prefix);
        inv_run = Fd<S>::mul(inv_run, ch[r]);
        o[2 + 2 * r] = Fd<S>::from_mont(Fd<S>::mul(rbc, inv_r));
        o[3 + 2 * r] = Fd<S>::from_mont(Fd<S>::mul(rbc, ch[r]));
    }
    const fe ps = Fd<S>::to_mont(polyscale[i]);
    fe xi = rbc;
    for (uint32_t m = 0; m < nc; m++) {
        o[2 + 2 * k + m] = Fd<S>::from_mont(xi);
        xi = Fd<S>::mul(xi, ps);
    }
    o[2 + 2 * k + nc] = Fd<S>::from_mont(RB);
    o[3 + 2 * k + nc] = Fd<S>::from_mont(Fd<S>::neg(Fd<S>::mul(RB, Z2)));
}

// term[t] = scalar[t] * point[t]: one thread per (opening, point), plain double-and-add on a 255-bit canonical
// scalar and an affine Montgomery base ((0, 0) = identity).
template <class F>
__global__ void __launch_bounds__(64) k_ipa_point_terms(const affine *__restrict__ pts, const fe *__restrict__ scalars, uint32_t total,
                                                        xyzz *__restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const affine q = pts[t];
    const fe s = scalars[t];
    xyzz acc = Ec<F>::identity();
    int top = 255;
    while (top >= 0 && !((s.v[top >> 5] >> (top & 31)) & 1u)) top--;
#pragma unroll 1
    for (int b = top; b >= 0; b--) {
        acc = Ec<F>::dbl(acc);
        if ((s.v[b >> 5] >> (b & 31)) & 1u) Ec<F>::add_mixed(acc, q);
    }
    out[t] = acc;
}

// P_i = -(sum of the npp terms of opening i): one warp per opening
template <class F>
__global__ void __launch_bounds__(32) k_ipa_sum_terms(const xyzz *__restrict__ terms, uint32_t npp, xyzz *__restrict__ out) {
    const xyzz *src = terms + (size_t)blockIdx.x * npp;
    xyzz acc = Ec<F>::identity();
    for (uint32_t t = threadIdx.x; t < npp; t += 32) Ec<F>::add(acc, src[t]);
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        xyzz o;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            o.x.v[i] = __shfl_down_sync(0xffffffffu, acc.x.v[i], d);
            o.y.v[i] = __shfl_down_sync(0xffffffffu, acc.y.v[i], d);
            o.zz.v[i] = __shfl_down_sync(0xffffffffu, acc.zz.v[i], d);
            o.zzz.v[i] = __shfl_down_sync(0xffffffffu, acc.zzz.v[i], d);
        }
        Ec<F>::add(acc, o);
    }
    if (threadIdx.x == 0) {
        acc.y = Fd<F>::neg(acc.y);
        out[blockIdx.x] = acc;
    }
}

}  // namespace pasta
