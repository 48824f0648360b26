// IPA scalar-side kernels for sm_100a: K4 `endo_to_field`, K2 `bpoly` tables / coefficients / batch
// combination, K5 `b_poly` evaluation.
//
// What they replace (all un-vendored, lambdaclass/openmina-proof-systems @ 44e0d3b):
//   * kimchi `ScalarChallenge::to_field(endo_r)`                       -> k_endo_to_field
//   * poly-commitment `b_poly_coefficients(chals)`                     -> k_bpoly_tables (+ the
//     on-the-fly product inside the MSM digit kernels, msm_impl.cuh) and k_bpoly_materialize
//   * the scalar side of `batch_dlog_accumulator_check` (sum_j r_j * b_poly_coefficients(chals_j))
//                                                                      -> k_bpoly_combine
//   * poly-commitment `b_poly(chals, x)`                               -> k_bpoly_eval
// Reference call site: `verify_block`, AL/operator/mina/lib/src/lib.rs:99-111.
//
// b_poly_coefficients has product structure: s[i] = prod_{j in bits(i)} chal[k-1-j].  Splitting
// i = 256*ih + il gives s[i] = hi[ih] * lo[il] with two 256-entry tables per proof (16 KiB), so a
// coefficient costs ONE field multiplication and the 2 MiB coefficient vector never has to exist:
// the MSM digit kernels rebuild each scalar from the tables (L1/L2 resident) when they need it.
#pragma once
#include <cuda_runtime.h>

#include "fe.cuh"

namespace pasta {

static constexpr int BPOLY_LO_BITS = 8;
static constexpr int BPOLY_TABLE = 512;  // fe per proof: lo[256] then hi[256]

// 128-bit prechallenge (two little-endian u64 limbs) -> field element, Montgomery form.
//   a = b = 2; for i = 63..0: a = 2a, b = 2b, s = bit(2i) ? +1 : -1, bit(2i+1) ? a += s : b += s
//   result = a * endo_r + b                                                     
template <class S>
__global__ void __launch_bounds__(128) k_endo_to_field(const uint4 *__restrict__ pre, fe *__restrict__ out, uint32_t n) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    uint4 r = __ldg(pre + idx);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    fe a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = b.v[i] = S::TWO(i);
    const fe one = Fd<S>::one();
    const fe neg_one = Fd<S>::neg(one);
    for (int i = 63; i >= 0; i--) {
        a = Fd<S>::dbl(a);
        b = Fd<S>::dbl(b);
        uint32_t two = (w[i >> 4] >> ((2 * i) & 31)) & 3u;  // bit 0 = sign, bit 1 = which accumulator
        fe s = (two & 1u) ? one : neg_one;
        if (two & 2u)
            a = Fd<S>::add(a, s);
        else
            b = Fd<S>::add(b, s);
    }
    out[idx] = Fd<S>::add(Fd<S>::mul(a, Fd<S>::endo_r()), b);
}

// Build the two product tables of every proof.  chals: [nproofs][k] Montgomery.  tables:
// [nproofs][512].  `scale` (optional, [nproofs] Montgomery) multiplies the hi table (the random
// r_j of the batched check).  lo_plain != 0 stores lo without the Montgomery factor, so that
// mont_mul(hi, lo) is directly the canonical integer the MSM digit extraction needs.
template <class S>
__global__ void __launch_bounds__(256) k_bpoly_tables(const fe *__restrict__ chals, fe *__restrict__ tables, uint32_t nproofs,
                                                      int k, const fe *__restrict__ scale, int lo_plain) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t proof = idx / BPOLY_TABLE, t = idx % BPOLY_TABLE;
    if (proof >= nproofs) return;
    const fe *c = chals + (size_t)proof * k;
    const int lo_bits = k < BPOLY_LO_BITS ? k : BPOLY_LO_BITS;
    fe acc = Fd<S>::one();
    if (t < 256) {
        for (int j = 0; j < lo_bits; j++)
            if ((t >> j) & 1u) acc = Fd<S>::mul(acc, c[k - 1 - j]);
        if (t >> lo_bits) acc = fe_zero();  // unused slots
        if (lo_plain) acc = Fd<S>::from_mont(acc);
    } else {
        uint32_t ih = t - 256;
        const int hi_bits = k - lo_bits;
        for (int j = 0; j < hi_bits; j++)
            if ((ih >> j) & 1u) acc = Fd<S>::mul(acc, c[k - 1 - lo_bits - j]);
        if (ih >> hi_bits)
            acc = fe_zero();
        else if (scale)
            acc = Fd<S>::mul(acc, scale[proof]);
    }
    tables[idx] = acc;
}

// The coefficient the tables encode (canonical if lo is plain, Montgomery otherwise).
template <class S>
__device__ __forceinline__ fe bpoly_coeff(const fe *__restrict__ table, uint32_t i) {
    fe lo = table[i & 255u];
    fe hi = table[256u + (i >> BPOLY_LO_BITS)];
    return Fd<S>::mul(hi, lo);
}

// Materialise b_poly_coefficients (canonical, 32 B each) -- parity tests and the K2 bench only; the
// verifier path never writes this vector.
template <class S>
__global__ void __launch_bounds__(256) k_bpoly_materialize(const fe *__restrict__ tables, fe *__restrict__ out, uint32_t nproofs, int k) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t total = (uint64_t)nproofs << k;
    if (idx >= total) return;
    uint32_t proof = (uint32_t)(idx >> k), i = (uint32_t)(idx & ((1u << k) - 1u));
    out[idx] = bpoly_coeff<S>(tables + (size_t)proof * BPOLY_TABLE, i);
}

// out[i] = sum_{j in subset} hi_j[i >> 8] * lo_j[i & 255]   (canonical), i < 2^k.
// Tables are Montgomery x Montgomery here and hi_j carries r_j, so this is the g-side scalar vector
// of batch_dlog_accumulator_check (sign handled by the caller: the commitments go on the other side
// of the equation).  One thread owns ITEMS coefficients that share `il` so its lo value is loaded
// once per proof; the 32 lanes of a warp read 32 consecutive lo entries (1 KiB, coalesced) and one
// broadcast hi entry per proof.
static constexpr int COMBINE_ITEMS = 4;
// Several groups per launch (blockIdx.y = group): group g sums the proofs subset[group_off[g] ..
// group_off[g+1]) into out + g * 2^k.  group_off == nullptr: one group of `nsub` proofs.
template <class S>
__global__ void __launch_bounds__(128) k_bpoly_combine(const fe *__restrict__ tables, const uint32_t *__restrict__ subset,
                                                       const uint32_t *__restrict__ group_off, uint32_t nsub, int k,
                                                       fe *__restrict__ out) {
    // thread -> (il, ih0): ih = ih0 * ITEMS + e
    const uint32_t n_hi = 1u << (k - BPOLY_LO_BITS);
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t il = tid & 255u, ih0 = (tid >> 8) * COMBINE_ITEMS;
    if (ih0 >= n_hi) return;
    uint32_t begin = 0, end = nsub;
    if (group_off) {
        begin = group_off[blockIdx.y];
        end = group_off[blockIdx.y + 1];
        out += (size_t)blockIdx.y << k;
    }
    fe acc[COMBINE_ITEMS];
#pragma unroll
    for (int e = 0; e < COMBINE_ITEMS; e++) acc[e] = fe_zero();
    // three proofs per step: the three products of a coefficient are summed unreduced and reduced ONCE (Fd::dot3,
    // fe.cuh: p^2 / 2^256 < p / 4, so the sum still reduces below 2 p) -- a quarter fewer multiply-pipe instructions
    uint32_t j = begin;
    for (; j + 3 <= end; j += 3) {
        const fe *t0 = tables + (size_t)(subset ? subset[j] : j) * BPOLY_TABLE;
        const fe *t1 = tables + (size_t)(subset ? subset[j + 1] : j + 1) * BPOLY_TABLE;
        const fe *t2 = tables + (size_t)(subset ? subset[j + 2] : j + 2) * BPOLY_TABLE;
        const fe lo0 = t0[il], lo1 = t1[il], lo2 = t2[il];
#pragma unroll
        for (int e = 0; e < COMBINE_ITEMS; e++) {
            if (ih0 + e < n_hi)
                acc[e] = Fd<S>::add(acc[e], Fd<S>::dot3(t0[256u + ih0 + e], lo0, t1[256u + ih0 + e], lo1, t2[256u + ih0 + e], lo2));
        }
    }
    for (; j < end; j++) {
        const fe *t = tables + (size_t)(subset ? subset[j] : j) * BPOLY_TABLE;
        fe lo = t[il];
#pragma unroll
        for (int e = 0; e < COMBINE_ITEMS; e++) {
            if (ih0 + e < n_hi) acc[e] = Fd<S>::add(acc[e], Fd<S>::mul(t[256u + ih0 + e], lo));
        }
    }
#pragma unroll
    for (int e = 0; e < COMBINE_ITEMS; e++)
        if (ih0 + e < n_hi) out[((size_t)(ih0 + e) << BPOLY_LO_BITS) + il] = Fd<S>::from_mont(acc[e]);
}

// b_poly(chals, x) = prod_{i<k} (1 + chals[i] * x^(2^(k-1-i)))  for npts points per proof.
// chals: [nproofs][k] Montgomery; x: [nproofs][npts] Montgomery; out likewise.
template <class S>
__global__ void __launch_bounds__(128) k_bpoly_eval(const fe *__restrict__ chals, const fe *__restrict__ x, fe *__restrict__ out,
                                                    uint32_t nproofs, uint32_t npts, int k) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nproofs * npts) return;
    const fe *c = chals + (size_t)(idx / npts) * k;
    fe pw = x[idx];
    fe acc = Fd<S>::one();
    for (int i = k - 1; i >= 0; i--) {
        acc = Fd<S>::mul(acc, Fd<S>::add(Fd<S>::one(), Fd<S>::mul(c[i], pw)));
        pw = Fd<S>::sqr(pw);
    }
    out[idx] = acc;
}

// kimchi `combined_inner_product` (un-vendored poly-commitment @ 44e0d3b, every polynomial
// has ONE chunk in the blockchain circuit and no degree-bound shift):
//   cip = sum_i polyscale^i * ( sum_j evalscale^j * evals[i][j] ),  i < npolys, j < npts
// evals: [nproofs][npolys][npts], scales: [nproofs][2] = (polyscale, evalscale); all Montgomery.
// One thread per proof (47 x 2 terms: a serial Horner is the whole job).
template <class S>
__global__ void __launch_bounds__(128) k_combined_inner_product(const fe *__restrict__ evals, const fe *__restrict__ scales,
                                                                fe *__restrict__ out, uint32_t nproofs, uint32_t npolys, uint32_t npts) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nproofs) return;
    const fe polyscale = scales[2 * (size_t)p], evalscale = scales[2 * (size_t)p + 1];
    const fe *e = evals + (size_t)p * npolys * npts;
    fe acc = fe_zero();
    for (int i = (int)npolys - 1; i >= 0; i--) {  // Horner in polyscale, inner Horner in evalscale
        fe inner = fe_zero();
        for (int j = (int)npts - 1; j >= 0; j--) inner = Fd<S>::add(Fd<S>::mul(inner, evalscale), e[(size_t)i * npts + j]);
        acc = Fd<S>::add(Fd<S>::mul(acc, polyscale), inner);
    }
    out[p] = acc;
}

// Coefficients of the Lagrange polynomials of the radix-2 domain of size n = 2^log_n (kimchi `add_lagrange_basis`,
// AL/operator/mina/lib/src/verifier_index.rs:204-208):  L_i(x) = (1/n) sum_j omega^(-i j) x^j, so the commitment of
// L_i is the MSM of row i below over g[0..n).  out: [count][n] canonical; omega_inv, n_inv: Montgomery.
template <class S>
__global__ void __launch_bounds__(256) k_lagrange_scalars(fe omega_inv, fe n_inv, int log_n, uint32_t first, uint32_t count,
                                                          fe *__restrict__ out) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = 1u << log_n;
    if (idx >= (uint64_t)count * n) return;
    const uint32_t i = first + (uint32_t)(idx >> log_n), j = (uint32_t)(idx & (n - 1));
    // omega^(-i j): the exponent only matters mod n
    const uint32_t e = (uint32_t)(((uint64_t)i * j) & (n - 1));
    fe acc = Fd<S>::one(), base = omega_inv;
    for (int b = 0; b < log_n; b++) {
        if ((e >> b) & 1u) acc = Fd<S>::mul(acc, base);
        base = Fd<S>::sqr(base);
    }
    out[idx] = Fd<S>::from_mont(Fd<S>::mul(acc, n_inv));
}

// canonical <-> Montgomery for flat arrays of field elements
template <class S>
__global__ void __launch_bounds__(256) k_fe_to_mont(const fe *__restrict__ in, fe *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = Fd<S>::to_mont(in[i]);
}
template <class S>
__global__ void __launch_bounds__(256) k_fe_from_mont(const fe *__restrict__ in, fe *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = Fd<S>::from_mont(in[i]);
}

// ---- launchers (scalar field chosen by id: 0 = Fp, 1 = Fq); implemented in ipa.cu -------------------
void launch_endo_to_field(int field, const void *d_pre16, fe *d_out, uint32_t n, cudaStream_t s);
void launch_bpoly_tables(int field, const fe *d_chals, fe *d_tables, uint32_t nproofs, int k, const fe *d_scale, bool lo_plain,
                         cudaStream_t s);
void launch_bpoly_materialize(int field, const fe *d_tables, fe *d_out, uint32_t nproofs, int k, cudaStream_t s);
// ngroups == 0: one group of nsub proofs (d_group_off ignored); else d_group_off[ngroups + 1] indexes d_subset
void launch_bpoly_combine(int field, const fe *d_tables, const uint32_t *d_subset, const uint32_t *d_group_off, uint32_t ngroups,
                          uint32_t nsub, int k, fe *d_out, cudaStream_t s);
void launch_bpoly_eval(int field, const fe *d_chals, const fe *d_x, fe *d_out, uint32_t nproofs, uint32_t npts, int k,
                       cudaStream_t s);
void launch_combined_inner_product(int field, const fe *d_evals, const fe *d_scales, fe *d_out, uint32_t nproofs, uint32_t npolys,
                                   uint32_t npts, cudaStream_t s);
void launch_lagrange_scalars(int field, const fe &omega_inv_mont, const fe &n_inv_mont, int log_n, uint32_t first, uint32_t count, fe *d_out,
                             cudaStream_t s);
void launch_fe_to_mont(int field, const fe *d_in, fe *d_out, uint32_t n, cudaStream_t s);
void launch_fe_from_mont(int field, const fe *d_in, fe *d_out, uint32_t n, cudaStream_t s);

}  // namespace pasta
