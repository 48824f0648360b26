// Pippenger bucket MSM over Pallas / Vesta -- hand-written kernels for sm_100a.
//
// What it replaces: ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul` (lambdaclass/openmina_algebra
// @ 017531e), the inner loop of `accumulator_check` and of the IPA final check under
// `verify_block` (AL/operator/mina/lib/src/lib.rs:99-111).
//
// Design (B200-first, not a translation of the CPU algorithm):
//   * The SRS is fixed, HBM is 180 GB: keep a table T[w][i] = 2^(c*w) * G_i of affine points resident
//     (16 x 65536 x 64 B = 64 MiB for Vesta -- it also fits the 126 MB L2).  Every window of every
//     scalar then lands in ONE shared set of 2^(c-1) signed-digit buckets: no per-window reduction,
//     no doublings at run time.
//   * Scalars -> signed c-bit digits -> counting sort by bucket (histogram, scan, scatter) -> one
//     thread per bucket sums its points with mixed XYZZ additions -> running-sum reduction in
//     log_m(#buckets) levels -> one affine point.
//   * A batch of independent MSMs over the same bases (one per proof) runs as one launch set:
//     bucket id = msm * buckets_per_msm + bucket.
//   * Arbitrary (non-resident) bases use the same kernels with one bucket set per window and a
//     final Horner pass (precompute = false).
// The group law is exact integer arithmetic; the affine output is bit-identical to arkworks'.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#pragma once
#include "ipa.cuh"
#include "msm.cuh"

namespace pasta {

// The scalar field of the curve whose coordinates live in F (Pallas: Fp -> Fq, Vesta: Fq -> Fp).
template <class F>
struct ScalarFieldOf;
template <>
struct ScalarFieldOf<FpParams> {
    using type = FqParams;
};
template <>
struct ScalarFieldOf<FqParams> {
    using type = FpParams;
};

#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));                   \
    } while (0)

static constexpr int MAX_LEVELS = 8;

struct DigitParams {
    uint32_t n_used;   // scalars per MSM
    uint32_t n_bases;  // stride of one table window
    uint32_t nmsm;
    int c;
    int W;             // windows per scalar
    int wsep;          // 1: one bucket set per window (generic bases); 0: shared (precomputed table)
    uint32_t nbw;      // buckets per window = 2^(c-1)
    int bpoly_k;       // SRC_BPOLY: log2 of the coefficients per proof (n_used == 1 << bpoly_k)
    uint32_t *err;     // device flag: bit 0 = a scalar did not fit the signed-digit windows (>= 2^255)
};

// Where a digit kernel takes its scalars from.
enum { SRC_MEMORY = 0, SRC_BPOLY = 1 };

// bits [pos, pos+c) of a 256-bit little-endian integer (c <= 24)
__device__ __forceinline__ uint32_t window_bits(const uint32_t s[8], int pos, int c) {
    int limb = pos >> 5, sh = pos & 31;
    if (limb >= 8) return 0;
    uint64_t v = s[limb];
    if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

// Pass 1 (count) and pass 3 (scatter) walk the digits the same way.  With SRC_BPOLY the scalar is
// b_poly_coefficients(chals_m)[i], rebuilt from the proof's two product tables (one field
// multiplication, ipa.cuh) instead of being read from a materialised 2 MiB vector.
template <bool SCATTER, int SRC, class S>
static __global__ void __launch_bounds__(256) k_digits(const uint32_t *__restrict__ scalars, DigitParams p,
                                                uint32_t *__restrict__ counters, uint32_t *__restrict__ pairs) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t total = (uint64_t)p.nmsm * p.n_used;
    if (idx >= total) return;
    uint32_t m = (uint32_t)(idx / p.n_used);
    uint32_t i = (uint32_t)(idx % p.n_used);
    uint32_t s[8];
    if (SRC == SRC_BPOLY) {
        fe v = bpoly_coeff<S>(reinterpret_cast<const fe *>(scalars) + (size_t)m * BPOLY_TABLE, i);
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = v.v[k];
    } else {
        const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + idx * 8);
        uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
        s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
        s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
    }
    const uint32_t half = 1u << (p.c - 1);
    uint32_t carry = 0;
    uint32_t group_base = m * (p.wsep ? (uint32_t)p.W : 1u);
    for (int w = 0; w < p.W; w++) {
        uint32_t raw = window_bits(s, w * p.c, p.c) + carry;
        uint32_t neg = raw > half;
        uint32_t mag = neg ? (1u << p.c) - raw : raw;
        carry = neg;
        if (mag == 0) continue;
        uint32_t bucket = (group_base + (p.wsep ? (uint32_t)w : 0u)) * p.nbw + (mag - 1u);
        if (!SCATTER) {
            atomicAdd(&counters[bucket], 1u);
        } else {
            uint32_t pos = atomicAdd(&counters[bucket], 1u);
            uint32_t entry = p.wsep ? i : (uint32_t)w * p.n_bases + i;
            pairs[pos] = entry | (neg << 31);
        }
    }
    // a carry out of the top window means the scalar needs more than W*c signed-digit bits
    if (!SCATTER && carry) atomicOr(p.err, 1u);
}

// ---- exclusive scan over u32 (three small kernels; inputs are a few 10^4 .. 10^7 counters) --------
static constexpr int SCAN_THREADS = 1024;
static constexpr int SCAN_ITEMS = 4;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                             uint32_t *__restrict__ tile_sums, uint32_t n) {
    __shared__ uint32_t warp_sums[32];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        local += v[k];
    }
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += t;
        }
        warp_sums[lane] = wi - ws;  // exclusive
        if (lane == 31) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + incl - local;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}
// single block: exclusive scan of tile sums in place; writes grand total to tile_sums[ntiles]
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_top(uint32_t *tile_sums, uint32_t ntiles) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t start = 0; start < ntiles; start += SCAN_THREADS) {
        uint32_t i = start + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0u, incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= (uint32_t)d) wi += t;
            }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_sums[wid] + incl - v;
        if (i < ntiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry_s;
}
// offsets[i] += tile offset; also writes offsets[n] = total and copies to the scatter cursors
static __global__ void __launch_bounds__(256) k_scan_finish(uint32_t *__restrict__ offsets, uint32_t *__restrict__ cursors,
                                                     const uint32_t *__restrict__ tile_sums, uint32_t n, uint32_t ntiles) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t v = offsets[i] + tile_sums[i / SCAN_TILE];
        offsets[i] = v;
        cursors[i] = v;
    } else if (i == n) {
        offsets[n] = tile_sums[ntiles];
    }
}

// ---- bucket accumulation: the dominant kernel ----------------------------------------------------
// One thread per bucket; the bucket's points are a contiguous run of `pairs`.  Each step gathers one
// 64-byte affine point (4 x LDG.128, mostly L2 hits: the whole table is 64 MiB) and does one mixed
// XYZZ addition (10 field multiplications).  The next point is fetched before the current addition
// so the gather latency hides behind ~2.5k integer instructions.
//
// Round-2 changes, each backed by profiles/r2_k_accumulate_{before,after}.md:
//   * buckets are visited in order of decreasing population (counting sort of the bucket sizes,
//     k_order_*): with one bucket per thread and Poisson(32) populations a warp used to wait for its
//     longest lane (22.8 of 32 lanes active on average); sorted, all 32 lanes of a warp have
//     (nearly) the same trip count;
//   * the field multiplication is a real call (`mul_call`) instead of ten inlined copies: the loop
//     body shrinks from ~85 KB to ~12 KB of SASS and stops missing the instruction cache (the top
//     stall reason before was `no_instruction`);
//   * a bucket contributes at most ACC_SEG points here; the rest of an over-full bucket (adversarial
//     scalars: all-equal digits put up to n*W points in ONE bucket) is summed by a whole block in
//     k_accumulate_overflow, so no thread ever walks more than ACC_SEG points.
static constexpr uint32_t ACC_SEG = 256;

template <class F>
__device__ __forceinline__ affine load_point(const affine *__restrict__ table, uint32_t entry) {
    const uint4 *p = reinterpret_cast<const uint4 *>(table + (entry & 0x7fffffffu));
    uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    affine q;
    q.x.v[0] = a.x; q.x.v[1] = a.y; q.x.v[2] = a.z; q.x.v[3] = a.w;
    q.x.v[4] = b.x; q.x.v[5] = b.y; q.x.v[6] = b.z; q.x.v[7] = b.w;
    q.y.v[0] = c.x; q.y.v[1] = c.y; q.y.v[2] = c.z; q.y.v[3] = c.w;
    q.y.v[4] = d.x; q.y.v[5] = d.y; q.y.v[6] = d.z; q.y.v[7] = d.w;
    return q;
}

template <class F>
__device__ __noinline__ fe mul_call(const fe a, const fe b) {
    return Fd<F>::mul(a, b);
}
template <class F>
__device__ __noinline__ fe sqr_call(const fe a) {
    return Fd<F>::sqr(a);
}
// a0 b0 + a1 b1 with ONE reduction (Fd::dot2): Y3 of the mixed addition is a difference of two products
template <class F>
__device__ __noinline__ fe dot2_call(const fe a0, const fe b0, const fe a1, const fe b1) {
    return Fd<F>::dot2(a0, b0, a1, b1);
}
// Ec<F>::add_mixed with the multiplications out of line (same formulas, "madd-2008-s")
template <class F>
__device__ __forceinline__ void add_mixed_compact(xyzz &p, const affine &q) {
    using fd = Fd<F>;
    if (Ec<F>::is_identity(q)) return;
    if (Ec<F>::is_identity(p)) {
        p = Ec<F>::from_affine(q);
        return;
    }
    fe U2 = mul_call<F>(q.x, p.zz);
    fe S2 = mul_call<F>(q.y, p.zzz);
    fe P = fd::sub(U2, p.x);
    fe R = fd::sub(S2, p.y);
    if (fe_is_zero(P)) {
        if (fe_is_zero(R))
            p = Ec<F>::dbl_affine(q);
        else
            p = Ec<F>::identity();
        return;
    }
    fe PP = sqr_call<F>(P);
    fe PPP = mul_call<F>(P, PP);
    fe Q = mul_call<F>(p.x, PP);
    fe X3 = fd::sub(fd::sub(sqr_call<F>(R), PPP), fd::dbl(Q));
    fe Y3 = dot2_call<F>(R, fd::sub(Q, X3), fd::neg(p.y), PPP);
    p.x = X3;
    p.y = Y3;
    p.zz = mul_call<F>(p.zz, PP);
    p.zzz = mul_call<F>(p.zzz, PPP);
}

// -- bucket visiting order: counting sort of min(population, ACC_SEG), longest first -------------------
static constexpr int ORDER_BINS = ACC_SEG + 1;
static constexpr int ORDER_BLOCK = 1024;
// pass 1: histogram of bucket sizes (per-block shared histogram, one global atomic per non-empty bin),
// plus the list of over-full buckets
static __global__ void __launch_bounds__(ORDER_BLOCK) k_order_hist(const uint32_t *__restrict__ offsets, uint32_t nb,
                                                            uint32_t *__restrict__ hist, uint32_t *__restrict__ over_list,
                                                            uint32_t *__restrict__ over_count) {
    __shared__ uint32_t sh[ORDER_BINS];
    for (int i = threadIdx.x; i < ORDER_BINS; i += ORDER_BLOCK) sh[i] = 0;
    __syncthreads();
    uint32_t b = blockIdx.x * ORDER_BLOCK + threadIdx.x;
    if (b < nb) {
        uint32_t cnt = offsets[b + 1] - offsets[b];
        if (cnt > ACC_SEG) {
            over_list[atomicAdd(over_count, 1u)] = b;
            cnt = ACC_SEG;
        }
        atomicAdd(&sh[cnt], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ORDER_BINS; i += ORDER_BLOCK)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// pass 2 (one block): start[c] = number of buckets with a larger size (descending order)
static __global__ void __launch_bounds__(32) k_order_scan(uint32_t *__restrict__ hist) {
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int c = ORDER_BINS - 1; c >= 0; c--) {
            uint32_t h = hist[c];
            hist[c] = run;
            run += h;
        }
    }
}
// pass 3: scatter bucket ids; a block reserves a range per bin with one global atomic
static __global__ void __launch_bounds__(ORDER_BLOCK) k_order_scatter(const uint32_t *__restrict__ offsets, uint32_t nb,
                                                               uint32_t *__restrict__ cursors, uint32_t *__restrict__ order) {
    __shared__ uint32_t sh[ORDER_BINS];
    __shared__ uint32_t base[ORDER_BINS];
    for (int i = threadIdx.x; i < ORDER_BINS; i += ORDER_BLOCK) sh[i] = 0;
    __syncthreads();
    uint32_t b = blockIdx.x * ORDER_BLOCK + threadIdx.x;
    uint32_t cnt = 0, local = 0;
    if (b < nb) {
        cnt = offsets[b + 1] - offsets[b];
        cnt = cnt > ACC_SEG ? ACC_SEG : cnt;
        local = atomicAdd(&sh[cnt], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ORDER_BINS; i += ORDER_BLOCK)
        if (sh[i]) base[i] = atomicAdd(&cursors[i], sh[i]);
    __syncthreads();
    if (b < nb) order[base[cnt] + local] = b;
}

template <class F>
__global__ void __launch_bounds__(128, 4) k_accumulate(const uint32_t *__restrict__ order, const uint32_t *__restrict__ offsets,
                                                       const uint32_t *__restrict__ pairs, const affine *__restrict__ table,
                                                       xyzz *__restrict__ buckets, uint32_t nbuckets) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbuckets) return;
    uint32_t b = order[t];
    uint32_t k = offsets[b], end = offsets[b + 1];
    if (end - k > ACC_SEG) end = k + ACC_SEG;
    xyzz acc = Ec<F>::identity();
    if (k < end) {
        // software pipeline: the point for step k+1 and the table index for step k+2 are in flight while
        // step k's ~2.3k integer instructions run (index -> point is a dependent pair of loads)
        uint32_t e = pairs[k];
        uint32_t e_next = (k + 1 < end) ? pairs[k + 1] : 0u;
        affine q = load_point<F>(table, e);
        for (;;) {
            uint32_t e_next2 = 0;
            affine q_next;
            bool more = (k + 1 < end);
            if (more) {
                q_next = load_point<F>(table, e_next);
                if (k + 2 < end) e_next2 = pairs[k + 2];
            }
            if (e >> 31) q.y = Fd<F>::neg(q.y);
            add_mixed_compact<F>(acc, q);
            if (!more) break;
            q = q_next;
            e = e_next;
            e_next = e_next2;
            k++;
        }
    }
    buckets[b] = acc;
}

template <class F>
__device__ __forceinline__ xyzz shfl_down_point(const xyzz &p, int delta);

// Few buckets (a single MSM, the first levels of the group testing): one thread per bucket leaves most of the GPU idle
// and the launch lasts as long as the fullest bucket's ~50 dependent additions.  Here TPB adjacent lanes share a bucket:
// each walks a contiguous part of its point list, then the parts are added with a shuffle tree (log2 TPB additions).
template <class F, int TPB>
__global__ void __launch_bounds__(128) k_accumulate_split(const uint32_t *__restrict__ order, const uint32_t *__restrict__ offsets,
                                                          const uint32_t *__restrict__ pairs, const affine *__restrict__ table,
                                                          xyzz *__restrict__ buckets, uint32_t nbuckets) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t slot = t / TPB, part = t % TPB;
    const bool live = slot < nbuckets;  // whole warps stay for the shuffles (nbuckets * TPB need not fill the last warp)
    uint32_t b = 0, k = 0, end = 0;
    if (live) {
        b = order[slot];
        k = offsets[b];
        end = offsets[b + 1];
        if (end - k > ACC_SEG) end = k + ACC_SEG;
        const uint32_t cnt = end - k, per = (cnt + TPB - 1) / TPB;
        const uint32_t lo = k + part * per;
        end = lo + per < end ? lo + per : end;
        k = lo < end ? lo : end;
    }
    xyzz acc = Ec<F>::identity();
    for (; k < end; k++) {
        uint32_t e = pairs[k];
        affine q = load_point<F>(table, e);
        if (e >> 31) q.y = Fd<F>::neg(q.y);
        add_mixed_compact<F>(acc, q);
    }
#pragma unroll 1
    for (int d = TPB / 2; d >= 1; d >>= 1) {
        xyzz o = shfl_down_point<F>(acc, d);
        if (part + d < TPB) Ec<F>::add(acc, o);
    }
    if (live && part == 0) buckets[b] = acc;
}

// Over-full buckets: one block per listed bucket (block-stride), threads stride over the points past
// ACC_SEG, then a shuffle / shared-memory tree; the block's sum is added into the bucket.
static constexpr int OVER_THREADS = 256;
template <class F>
__global__ void __launch_bounds__(OVER_THREADS) k_accumulate_overflow(const uint32_t *__restrict__ over_list,
                                                                      const uint32_t *__restrict__ over_count,
                                                                      const uint32_t *__restrict__ offsets,
                                                                      const uint32_t *__restrict__ pairs, const affine *__restrict__ table,
                                                                      xyzz *__restrict__ buckets) {
    __shared__ xyzz warp_part[OVER_THREADS / 32];
    const uint32_t nover = *over_count;
    for (uint32_t it = blockIdx.x; it < nover; it += gridDim.x) {
        const uint32_t b = over_list[it];
        const uint32_t begin = offsets[b] + ACC_SEG, end = offsets[b + 1];
        xyzz acc = Ec<F>::identity();
        for (uint32_t k = begin + threadIdx.x; k < end; k += OVER_THREADS) {
            uint32_t e = pairs[k];
            affine q = load_point<F>(table, e);
            if (e >> 31) q.y = Fd<F>::neg(q.y);
            add_mixed_compact<F>(acc, q);
        }
#pragma unroll 1
        for (int d = 16; d >= 1; d >>= 1) {
            xyzz o = shfl_down_point<F>(acc, d);
            Ec<F>::add(acc, o);
        }
        const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) warp_part[wid] = acc;
        __syncthreads();
        if (wid == 0) {
            acc = lane < OVER_THREADS / 32 ? warp_part[lane] : Ec<F>::identity();
#pragma unroll 1
            for (int d = 4; d >= 1; d >>= 1) {
                xyzz o = shfl_down_point<F>(acc, d);
                Ec<F>::add(acc, o);
            }
            if (lane == 0) {
                xyzz cur = buckets[b];
                Ec<F>::add(cur, acc);
                buckets[b] = cur;
            }
        }
        __syncthreads();
    }
}

// ---- running-sum reduction -------------------------------------------------------------------------
// in: groups x N points.  Thread (g, t) walks m consecutive points from the top:
//   S = sum_j in[t*m+j],   Wt = sum_j (j+1) * in[t*m+j]
template <class F>
__global__ void __launch_bounds__(128) k_reduce_level(const xyzz *__restrict__ in, xyzz *__restrict__ outS,
                                                      xyzz *__restrict__ outW, uint32_t total_out, int m) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total_out) return;
    const xyzz *src = in + (size_t)id * m;
    xyzz run = Ec<F>::identity(), acc = Ec<F>::identity();
    for (int j = m - 1; j >= 0; j--) {
        xyzz b = src[j];
        Ec<F>::add(run, b);
        Ec<F>::add(acc, run);
    }
    outS[id] = run;
    outW[id] = acc;
}

template <class F>
__device__ __forceinline__ xyzz shfl_down_point(const xyzz &p, int delta) {
    xyzz r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], delta);
        r.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], delta);
        r.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], delta);
        r.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], delta);
    }
    return r;
}
// sum over the warp of one point per lane (valid in lane 0); every lane must call it
template <class F>
__device__ __forceinline__ xyzz warp_sum_points(xyzz acc) {
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        xyzz o = shfl_down_point<F>(acc, d);
        Ec<F>::add(acc, o);
    }
    return acc;
}

// ---- the tail of the bucket reduction: bit sums ------------------------------------------------------
// After L running-sum passes a group is down to N = 2^nbits points B_t that still carry the weight t:
//   Z = sum_t t * B_t = sum_k 2^k * U_k,   U_k = sum of the B_t whose index has bit k set.
// The U_k are plain sums, all independent, so the latency-bound chain of further running-sum levels (5 x 2
// launches, ~1.2 ms for one MSM in round 1) becomes ONE launch of independent warps plus a short Horner in
// k_finalize.  The same launch also produces, as partial sums over <= 256 points each, T = sum_t B_t and
// the sum of every pass's W values (which the old code took one launch per level for).
// One warp per output: a lane walks its strided share, then a shuffle tree (a block-wide tree would spend
// most of its additions on the tree itself: measured 4x slower at 77 groups).
static constexpr uint32_t TAIL_ITEMS = 256;   // points per partial-sum warp
static constexpr uint32_t TAIL_MAX_PARTS = 32;  // partials of one source (k_finalize adds them in one warp)
static constexpr uint32_t TAIL_BITMAX = 512;  // bit sums start at <= 2^9 points per group
struct TailParams {
    int L;                       // running-sum passes done before the bit sums
    int log_m[MAX_LEVELS];       // log2 of the chunk size of pass l
    int nbits;                   // log2 of the points per group left after the passes
    int W;                       // windows to Horner-combine per MSM (1 with the precomputed table)
    int c;
    uint32_t groups;             // nmsm * W
    uint32_t slots;              // outputs per group
    uint32_t parts_last;         // partials of T (slots nbits .. nbits + parts_last)
    uint32_t w_n[MAX_LEVELS];    // W values per group written by pass l
    uint32_t w_parts[MAX_LEVELS], w_slot[MAX_LEVELS];  // partials of sum W^l and their first slot
    uint64_t w_off[MAX_LEVELS];  // where pass l's W values start inside the W buffer (in points)
};
static inline uint32_t tail_parts(uint32_t n) {
    uint32_t p = (n + TAIL_ITEMS - 1) / TAIL_ITEMS;
    return p < 1 ? 1 : (p > TAIL_MAX_PARTS ? TAIL_MAX_PARTS : p);
}
// out[g][slot]: slot < nbits -> U_slot; then the partials of T; then the partials of sum W^l per pass.
template <class F>
__global__ void __launch_bounds__(32) k_bit_sums(const xyzz *__restrict__ last, const xyzz *__restrict__ wbuf, TailParams tp,
                                                 xyzz *__restrict__ out) {
    const uint32_t g = blockIdx.x, slot = blockIdx.y, lane = threadIdx.x;
    const uint32_t N = 1u << tp.nbits;
    xyzz acc = Ec<F>::identity();
    if (slot < (uint32_t)tp.nbits) {
        const xyzz *src = last + (size_t)g * N;
        const uint32_t mask = 1u << slot;
        for (uint32_t i = lane; i < N / 2; i += 32) {
            uint32_t t = ((i >> slot) << (slot + 1)) | mask | (i & (mask - 1u));
            Ec<F>::add(acc, src[t]);
        }
    } else {
        const xyzz *src = last + (size_t)g * N;
        uint32_t n = N, parts = tp.parts_last, q = slot - tp.nbits;
        for (int l = 0; l < tp.L; l++)
            if (slot >= tp.w_slot[l]) {
                n = tp.w_n[l];
                parts = tp.w_parts[l];
                q = slot - tp.w_slot[l];
                src = wbuf + tp.w_off[l] + (size_t)g * n;
            }
        const uint32_t per = (n + parts - 1) / parts, begin = q * per, end = min(n, begin + per);
        for (uint32_t t = begin + lane; t < end; t += 32) Ec<F>::add(acc, src[t]);
    }
    acc = warp_sum_points<F>(acc);
    if (lane == 0) out[(size_t)g * tp.slots + slot] = acc;
}

// One warp per group (= one window of one MSM): Horner over the bit sums (lane k doubles U_k k times, then a shuffle
// tree) and unwinding of the running-sum passes.  Valid in lane 0.
//   pass l turned points B with weights (b+1) into S (weights t) and W:  sum = sum W + m_l * sum_t t*S_t
//   below the last pass the weights are t, above it they are (t+1):  D_l = sumW_l + m_l * (D_{l+1} - T).
// s * p for a small non-negative scalar (binary method; a power of two costs only its doublings)
template <class F>
__device__ __forceinline__ xyzz small_scalar_mul(const xyzz &p, uint64_t s) {
    xyzz acc = Ec<F>::identity();
    if (s == 0) return acc;
#pragma unroll 1
    for (int b = 63 - __clzll((long long)s); b >= 0; b--) {
        acc = Ec<F>::dbl(acc);
        if ((s >> b) & 1ull) Ec<F>::add(acc, p);
    }
    return acc;
}
// Unwound, the result of a group is ONE linear combination of its slots with small coefficients:
//   D = sum_k 2^k M_L U_k  +  sum_l M_l (sum of W^l partials)  -  (M_1 + .. + M_{L-1}) T      (L >= 1;  M_l = m_0 .. m_{l-1})
//   D = sum_k 2^k U_k + T                                                                      (L == 0)
// With at most 32 slots every lane scales its own slot (the deepest does nbits + log2 M_L doublings) and one shuffle
// tree adds them: ~95 us instead of ~150 us for the level-by-level version below.
template <class F>
__device__ __forceinline__ xyzz group_result_flat(const xyzz *__restrict__ base, const TailParams &tp, uint32_t lane) {
    int sh[MAX_LEVELS + 1];
    sh[0] = 0;
    for (int l = 0; l < tp.L; l++) sh[l + 1] = sh[l] + tp.log_m[l];
    uint64_t coef = 0;
    bool negate = false;
    if (lane < (uint32_t)tp.nbits) {
        coef = 1ull << (lane + sh[tp.L]);
    } else if (lane < (uint32_t)tp.nbits + tp.parts_last) {
        if (tp.L == 0) {
            coef = 1;
        } else {
            for (int l = 1; l < tp.L; l++) coef += 1ull << sh[l];
            negate = true;
        }
    } else if (lane < tp.slots) {
        for (int l = 0; l < tp.L; l++)
            if (lane >= tp.w_slot[l]) coef = 1ull << sh[l];
    }
    xyzz V = (lane < tp.slots && coef) ? small_scalar_mul<F>(base[lane], coef) : Ec<F>::identity();
    if (negate) V.y = Fd<F>::neg(V.y);
    return warp_sum_points<F>(V);
}
template <class F>
__device__ __forceinline__ xyzz group_result(const xyzz *__restrict__ base, const TailParams &tp, uint32_t lane) {
    if (tp.slots <= 32 && tp.nbits + (tp.L ? tp.log_m[0] * tp.L : 0) < 60) return group_result_flat<F>(base, tp, lane);
    xyzz V = lane < (uint32_t)tp.nbits ? base[lane] : Ec<F>::identity();
    for (uint32_t k = 0; k < lane && lane < (uint32_t)tp.nbits; k++) V = Ec<F>::dbl(V);
    xyzz D = warp_sum_points<F>(V);  // Z = sum_t t * B_t
    xyzz T = warp_sum_points<F>(lane < tp.parts_last ? base[tp.nbits + lane] : Ec<F>::identity());
    if (tp.L == 0) {
        Ec<F>::add(D, T);  // bucket b weighs b + 1
    } else {
        xyzz negT = T;
        negT.y = Fd<F>::neg(negT.y);
        for (int l = tp.L - 1; l >= 0; l--) {
            if (l != tp.L - 1) Ec<F>::add(D, negT);
            for (int k = 0; k < tp.log_m[l]; k++) D = Ec<F>::dbl(D);
            xyzz Wl = warp_sum_points<F>(lane < tp.w_parts[l] ? base[tp.w_slot[l] + lane] : Ec<F>::identity());
            Ec<F>::add(D, Wl);
        }
    }
    return D;
}
// Fixed-base engine (one group per MSM): the group result IS the MSM.  AFFINE = false leaves the result in XYZZ
// form (no field inversion): callers that only COMPARE results (the accumulator checks) cross-multiply instead.
template <class F, bool AFFINE>
__global__ void __launch_bounds__(32) k_finalize(const xyzz *__restrict__ sums /* [groups][slots] */, TailParams tp,
                                                 void *__restrict__ out_any) {
    const uint32_t msm = blockIdx.x, lane = threadIdx.x;
    xyzz res = group_result<F>(sums + (size_t)msm * tp.slots, tp, lane);
    if (lane != 0) return;
    if (AFFINE)
        reinterpret_cast<affine *>(out_any)[msm] = Ec<F>::to_affine(res);
    else
        reinterpret_cast<xyzz *>(out_any)[msm] = res;
}
// Generic bases: every window of every MSM gets its own warp (the tails of W windows in ONE warp cost W x 0.2 ms) ...
template <class F>
__global__ void __launch_bounds__(32) k_window_results(const xyzz *__restrict__ sums, TailParams tp, xyzz *__restrict__ out) {
    xyzz D = group_result<F>(sums + (size_t)blockIdx.x * tp.slots, tp, threadIdx.x);
    if (threadIdx.x == 0) out[blockIdx.x] = D;
}
// ... and one thread per MSM combines them: res = sum_w 2^(c w) D_w by Horner (255 doublings, the latency floor of any
// variable-base MSM).
template <class F, bool AFFINE>
__global__ void __launch_bounds__(32) k_combine_windows(const xyzz *__restrict__ win /* [nmsm][W] */, TailParams tp,
                                                        void *__restrict__ out_any, uint32_t nmsm) {
    const uint32_t msm = blockIdx.x * blockDim.x + threadIdx.x;
    if (msm >= nmsm) return;
    xyzz res = Ec<F>::identity();
    for (int w = tp.W - 1; w >= 0; w--) {
        if (w != tp.W - 1)
            for (int k = 0; k < tp.c; k++) res = Ec<F>::dbl(res);
        Ec<F>::add(res, win[(size_t)msm * tp.W + w]);
    }
    if (AFFINE)
        reinterpret_cast<affine *>(out_any)[msm] = Ec<F>::to_affine(res);
    else
        reinterpret_cast<xyzz *>(out_any)[msm] = res;
}

// ---- fixed-base table: T[w] = 2^c * T[w-1], normalised to affine (one Fermat inversion per point) ----
template <class F>
__global__ void __launch_bounds__(128) k_table_next(const affine *__restrict__ prev, affine *__restrict__ next, uint32_t n, int c) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz p = Ec<F>::from_affine(prev[i]);
    for (int k = 0; k < c; k++) p = Ec<F>::dbl(p);
    next[i] = Ec<F>::to_affine(p);
}

template <class F>
__global__ void __launch_bounds__(256) k_affine_to_mont(const uint32_t *__restrict__ in, affine *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        x.v[k] = in[(size_t)i * 16 + k];
        y.v[k] = in[(size_t)i * 16 + 8 + k];
    }
    affine q;
    q.x = Fd<F>::to_mont(x);
    q.y = Fd<F>::to_mont(y);
    out[i] = q;
}
// Same, for untrusted input: flags (bad |= 1) any point whose coordinates are not canonical (>= p)
// or that is neither (0,0) nor on y^2 = x^3 + 5 -- arkworks' `of_coordinates` does not check, but a
// verifier fed attacker-chosen points must never run the group law on an off-curve point.
template <class F>
__device__ __forceinline__ bool fe_is_canonical(const fe &a) {
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a.v[i] - F::MOD(i) - borrow;
        borrow = (uint32_t)(t >> 63);
    }
    return borrow != 0;  // a < p
}
template <class F>
__global__ void __launch_bounds__(256) k_affine_to_mont_checked(const uint32_t *__restrict__ in, affine *__restrict__ out, uint32_t n,
                                                                uint32_t *__restrict__ bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        x.v[k] = in[(size_t)i * 16 + k];
        y.v[k] = in[(size_t)i * 16 + 8 + k];
    }
    bool ok = fe_is_canonical<F>(x) && fe_is_canonical<F>(y);
    affine q;
    q.x = Fd<F>::to_mont(x);
    q.y = Fd<F>::to_mont(y);
    if (ok && !(fe_is_zero(x) && fe_is_zero(y))) {
        fe rhs = Fd<F>::add(Fd<F>::mul(Fd<F>::sqr(q.x), q.x), Fd<F>::five());
        ok = fe_eq(Fd<F>::sqr(q.y), rhs);
    }
    if (!ok) {
        atomicOr(bad, 1u);
        q.x = fe_zero();
        q.y = fe_zero();
    }
    out[i] = q;
}
template <class F>
void launch_affine_to_mont_checked_t(const uint32_t *d_in, affine *d_out, uint32_t n, uint32_t *d_bad, cudaStream_t s) {
    if (n == 0) return;
    k_affine_to_mont_checked<F><<<(n + 255) / 256, 256, 0, s>>>(d_in, d_out, n, d_bad);
}

template <class F>
__global__ void __launch_bounds__(256) k_affine_from_mont(const affine *__restrict__ in, uint32_t *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine q = in[i];
    fe x = Fd<F>::from_mont(q.x), y = Fd<F>::from_mont(q.y);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        out[(size_t)i * 16 + k] = x.v[k];
        out[(size_t)i * 16 + 8 + k] = y.v[k];
    }
}

template <class F>
void launch_affine_to_mont_t(const uint32_t *d_in, affine *d_out, uint32_t n, cudaStream_t s) {
    if (n == 0) return;
    k_affine_to_mont<F><<<(n + 255) / 256, 256, 0, s>>>(d_in, d_out, n);
}
template <class F>
void launch_affine_from_mont_t(const affine *d_in, uint32_t *d_out, uint32_t n, cudaStream_t s) {
    if (n == 0) return;
    k_affine_from_mont<F><<<(n + 255) / 256, 256, 0, s>>>(d_in, d_out, n);
}

// ---------------------------------------------------------------------------------------------------
template <class F>
class MsmEngine : public MsmEngineBase {
    using S = typename ScalarFieldOf<F>::type;

   public:
    ~MsmEngine() override { release(); }

    void set_bases(const affine *d_bases, uint32_t n, const MsmConfig &cfg, cudaStream_t s) override {
        // validate and plan into locals first; members change only once nothing can throw any more
        if (cfg.c < 2 || cfg.c > 20) throw std::runtime_error("msm: window bits out of range");
        if (cfg.leaf < 2 || (cfg.leaf & (cfg.leaf - 1))) throw std::runtime_error("msm: leaf must be a power of two");
        const int W = 255 / cfg.c + 1;
        const uint32_t nbw = 1u << (cfg.c - 1);
        affine *table_new = nullptr;
        if (cfg.precompute) {
            if ((uint64_t)W * n >= (1ull << 31)) throw std::runtime_error("msm: table too large for 31-bit entries");
            CUDA_OK(cudaMalloc(&table_new, (size_t)W * n * sizeof(affine)));
            cudaError_t e = cudaMemcpyAsync(table_new, d_bases, (size_t)n * sizeof(affine), cudaMemcpyDeviceToDevice, s);
            for (int w = 1; w < W && e == cudaSuccess; w++) {
                k_table_next<F><<<(n + 127) / 128, 128, 0, s>>>(table_new + (size_t)(w - 1) * n, table_new + (size_t)w * n, n, cfg.c);
                launches_++;
                e = cudaGetLastError();
            }
            if (e != cudaSuccess) {
                cudaFree(table_new);
                throw std::runtime_error(std::string("msm: table build failed: ") + cudaGetErrorString(e));
            }
        } else if (n >= (1u << 31)) {
            throw std::runtime_error("msm: too many bases");
        }
        // commit
        const bool same_plan = W == W_ && nbw == nbw_ && cfg.leaf == cfg_.leaf && cfg.precompute == cfg_.precompute;
        free_dev(table_own_);
        table_own_ = table_new;
        table_ = cfg.precompute ? table_new : d_bases;
        cfg_ = cfg;
        n_bases_ = n;
        W_ = W;
        nbw_ = nbw;
        if (!same_plan) drop_workspace();  // the reduction plan sizes the level buffers
    }

    void run(const uint32_t *d_scalars, uint32_t nmsm, uint32_t n_used, affine *d_out, cudaStream_t s) override {
        run_src(SRC_MEMORY, d_scalars, 0, nmsm, n_used, d_out, true, s);
    }
    void run_bpoly(const fe *d_tables, uint32_t nmsm, int k, affine *d_out, cudaStream_t s) override {
        if (k < BPOLY_LO_BITS || k > 2 * BPOLY_LO_BITS) throw std::runtime_error("msm: bpoly rounds must be in [8, 16]");
        run_src(SRC_BPOLY, reinterpret_cast<const uint32_t *>(d_tables), k, nmsm, 1u << k, d_out, true, s);
    }
    void run_xyzz(const uint32_t *d_scalars, uint32_t nmsm, uint32_t n_used, xyzz *d_out, cudaStream_t s) override {
        run_src(SRC_MEMORY, d_scalars, 0, nmsm, n_used, d_out, false, s);
    }
    void run_bpoly_xyzz(const fe *d_tables, uint32_t nmsm, int k, xyzz *d_out, cudaStream_t s) override {
        if (k < BPOLY_LO_BITS || k > 2 * BPOLY_LO_BITS) throw std::runtime_error("msm: bpoly rounds must be in [8, 16]");
        run_src(SRC_BPOLY, reinterpret_cast<const uint32_t *>(d_tables), k, nmsm, 1u << k, d_out, false, s);
    }

    size_t workspace_bytes() const override { return ws_bytes_ + (table_own_ ? (size_t)W_ * n_bases_ * sizeof(affine) : 0); }
    uint64_t launches() const override { return launches_; }
    uint32_t take_error(cudaStream_t s) override {
        if (!err_) return 0;
        uint32_t h = 0;
        CUDA_OK(cudaMemcpyAsync(&h, err_, sizeof h, cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaStreamSynchronize(s));
        if (h) CUDA_OK(cudaMemsetAsync(err_, 0, sizeof h, s));
        return h;
    }
    void enable_kernel_timing(bool on) override { timing_ = on; }
    // sum over every chunk of the last run()/run_bpoly() call
    float last_accumulate_ms() override {
        float total = 0.f;
        for (size_t i = 0; i < ev_used_; i++) {
            float ms = 0.f;
            CUDA_OK(cudaEventSynchronize(ev_[2 * i + 1]));
            CUDA_OK(cudaEventElapsedTime(&ms, ev_[2 * i], ev_[2 * i + 1]));
            total += ms;
        }
        return total;
    }

   private:
    template <class T>
    static void free_dev(T *&p) {
        if (p) cudaFree(p);
        p = nullptr;
    }
    void drop_workspace() {
        cap_buckets_ = cap_pairs_ = 0;
        cap_groups_ = 0;
        ws_bytes_ = 0;
        free_dev(counters_);
        free_dev(offsets_);
        free_dev(tile_sums_);
        free_dev(pairs_);
        free_dev(order_);
        free_dev(over_list_);
        free_dev(order_hist_);
        free_dev(buckets_);
        free_dev(lvlS_[0]);
        free_dev(lvlS_[1]);
        free_dev(lvlW_);
        free_dev(sums_);
    }
    void release() {
        free_dev(table_own_);
        drop_workspace();
        free_dev(err_);
        for (cudaEvent_t e : ev_) cudaEventDestroy(e);
        ev_.clear();
    }

    // chunk size of a running-sum pass over N points per group
    int tail_log_m(uint32_t N) const {
        int lg = 0;
        while ((1 << (lg + 1)) <= cfg_.leaf && (2u << lg) <= N) lg++;
        return lg;
    }
    TailParams plan_tail(uint32_t groups) const {
        TailParams tp{};
        uint32_t N = nbw_;
        uint64_t off = 0;
        while (N > TAIL_BITMAX) {
            if (tp.L >= MAX_LEVELS) throw std::runtime_error("msm: too many reduction levels");
            const int lg = tail_log_m(N);
            tp.log_m[tp.L] = lg;
            N >>= lg;
            tp.w_n[tp.L] = N;
            tp.w_off[tp.L] = off;
            off += (uint64_t)groups * N;
            tp.L++;
        }
        int nbits = 0;
        while ((1u << nbits) < N) nbits++;
        tp.nbits = nbits;
        tp.parts_last = tail_parts(N);
        tp.slots = (uint32_t)nbits + tp.parts_last;
        for (int l = 0; l < tp.L; l++) {
            tp.w_parts[l] = tail_parts(tp.w_n[l]);
            tp.w_slot[l] = tp.slots;
            tp.slots += tp.w_parts[l];
        }
        return tp;
    }
    uint32_t max_slots() const { return plan_tail(1).slots; }

    void run_src(int src, const uint32_t *d_src, int bpoly_k, uint32_t nmsm, uint32_t n_used, void *d_out, bool affine_out, cudaStream_t s) {
        if (!table_) throw std::runtime_error("msm: no bases set");
        if (n_used > n_bases_) throw std::runtime_error("msm: n_used exceeds resident bases");
        if (nmsm == 0) return;
        const uint32_t groups_per_msm = cfg_.precompute ? 1u : (uint32_t)W_;
        const uint64_t buckets_per_msm = (uint64_t)groups_per_msm * nbw_;
        const uint64_t pairs_per_msm = (uint64_t)n_used * W_;
        // chunk the batch so the workspace stays bounded
        uint64_t chunk = nmsm;
        const uint64_t max_buckets = 1ull << 23, max_pairs = 1ull << 28;
        if (chunk * buckets_per_msm > max_buckets) chunk = max_buckets / buckets_per_msm;
        if (pairs_per_msm && chunk * pairs_per_msm > max_pairs) chunk = max_pairs / pairs_per_msm;
        if (chunk == 0) chunk = 1;
        if (chunk * buckets_per_msm >= (1ull << 32) || chunk * pairs_per_msm >= (1ull << 32))
            throw std::runtime_error("msm: problem too large for 32-bit indices");
        ensure_workspace((uint32_t)chunk, n_used);
        ev_used_ = 0;
        for (uint32_t done = 0; done < nmsm; done += (uint32_t)chunk) {
            uint32_t cur = (uint32_t)std::min<uint64_t>(chunk, nmsm - done);
            const uint32_t *src_ptr = src == SRC_BPOLY ? d_src + (size_t)done * BPOLY_TABLE * 8 : d_src + (size_t)done * n_used * 8;
            void *out_ptr = affine_out ? (void *)(reinterpret_cast<affine *>(d_out) + done) : (void *)(reinterpret_cast<xyzz *>(d_out) + done);
            run_chunk(src, src_ptr, bpoly_k, cur, n_used, out_ptr, affine_out, s);
        }
    }

    void ensure_workspace(uint32_t nmsm, uint32_t n_used) {
        const uint32_t groups = nmsm * (cfg_.precompute ? 1u : (uint32_t)W_);
        const uint64_t nb = (uint64_t)groups * nbw_;
        const uint64_t np = (uint64_t)nmsm * n_used * W_;
        if (!err_) {
            CUDA_OK(cudaMalloc(&err_, sizeof(uint32_t)));
            CUDA_OK(cudaMemset(err_, 0, sizeof(uint32_t)));
        }
        if (nb <= cap_buckets_ && np <= cap_pairs_ && groups <= cap_groups_) return;
        const uint64_t want_b = std::max<uint64_t>(nb, cap_buckets_), want_p = std::max<uint64_t>(np, cap_pairs_);
        const uint32_t want_g = std::max<uint32_t>(groups, cap_groups_);
        drop_workspace();  // caps are zero from here until every allocation below has succeeded
        size_t ntiles = (want_b + SCAN_TILE - 1) / SCAN_TILE;
        // level buffers: pass 0 writes want_b / m points, later passes a factor m fewer each (m >= 2)
        const int lg0 = tail_log_m(nbw_);
        size_t first = (want_b >> lg0) + 1;
        size_t bytes = 0;
        auto alloc = [&](auto &ptr, size_t b) {
            cudaError_t e = cudaMalloc(&ptr, b);
            if (e != cudaSuccess) {
                drop_workspace();
                throw std::runtime_error(std::string("msm: workspace allocation failed: ") + cudaGetErrorString(e));
            }
            bytes += b;
        };
        alloc(counters_, (want_b + 1) * sizeof(uint32_t));
        alloc(offsets_, (want_b + 1) * sizeof(uint32_t));
        alloc(tile_sums_, (ntiles + 1) * sizeof(uint32_t));
        alloc(pairs_, std::max<uint64_t>(want_p, 1) * sizeof(uint32_t));
        alloc(order_, want_b * sizeof(uint32_t));
        alloc(over_list_, (want_p / ACC_SEG + 1) * sizeof(uint32_t));
        alloc(order_hist_, (ORDER_BINS + 1) * sizeof(uint32_t));
        alloc(buckets_, want_b * sizeof(xyzz));
        alloc(lvlS_[0], first * sizeof(xyzz));
        alloc(lvlS_[1], first * sizeof(xyzz));
        alloc(lvlW_, 2 * first * sizeof(xyzz));
        alloc(sums_, (size_t)max_slots() * want_g * sizeof(xyzz));
        cap_buckets_ = want_b;
        cap_pairs_ = want_p;
        cap_groups_ = want_g;
        ws_bytes_ = bytes;
    }

    template <bool SCATTER>
    void launch_digits(int src, const uint32_t *d_src, const DigitParams &dp, uint32_t dblocks, uint32_t *counters, uint32_t *pairs,
                       cudaStream_t s) {
        if (!dblocks) return;
        if (src == SRC_BPOLY)
            k_digits<SCATTER, SRC_BPOLY, S><<<dblocks, 256, 0, s>>>(d_src, dp, counters, pairs);
        else
            k_digits<SCATTER, SRC_MEMORY, S><<<dblocks, 256, 0, s>>>(d_src, dp, counters, pairs);
        launches_++;
    }

    void run_chunk(int src, const uint32_t *d_src, int bpoly_k, uint32_t nmsm, uint32_t n_used, void *d_out, bool affine_out, cudaStream_t s) {
        DigitParams dp;
        dp.n_used = n_used;
        dp.n_bases = n_bases_;
        dp.nmsm = nmsm;
        dp.c = cfg_.c;
        dp.W = W_;
        dp.wsep = cfg_.precompute ? 0 : 1;
        dp.nbw = nbw_;
        dp.bpoly_k = bpoly_k;
        dp.err = err_;
        const uint32_t groups = nmsm * (cfg_.precompute ? 1u : (uint32_t)W_);
        const uint32_t nb = groups * nbw_;
        const uint64_t nscal = (uint64_t)nmsm * n_used;
        const uint32_t dblocks = (uint32_t)((nscal + 255) / 256);
        const uint32_t ntiles = (nb + SCAN_TILE - 1) / SCAN_TILE;

        CUDA_OK(cudaMemsetAsync(counters_, 0, (size_t)(nb + 1) * sizeof(uint32_t), s));
        launch_digits<false>(src, d_src, dp, dblocks, counters_, nullptr, s);
        k_scan_tiles<<<ntiles, SCAN_THREADS, 0, s>>>(counters_, offsets_, tile_sums_, nb);
        k_scan_top<<<1, SCAN_THREADS, 0, s>>>(tile_sums_, ntiles);
        k_scan_finish<<<(nb + 1 + 255) / 256, 256, 0, s>>>(offsets_, counters_, tile_sums_, nb, ntiles);
        launches_ += 3;
        launch_digits<true>(src, d_src, dp, dblocks, counters_, pairs_, s);
        // visiting order (longest bucket first) and the list of over-full buckets
        CUDA_OK(cudaMemsetAsync(order_hist_, 0, (ORDER_BINS + 1) * sizeof(uint32_t), s));
        k_order_hist<<<(nb + ORDER_BLOCK - 1) / ORDER_BLOCK, ORDER_BLOCK, 0, s>>>(offsets_, nb, order_hist_, over_list_, order_hist_ + ORDER_BINS);
        k_order_scan<<<1, 32, 0, s>>>(order_hist_);
        k_order_scatter<<<(nb + ORDER_BLOCK - 1) / ORDER_BLOCK, ORDER_BLOCK, 0, s>>>(offsets_, nb, order_hist_, order_);
        launches_ += 3;
        if (timing_) {
            while (ev_.size() < 2 * (ev_used_ + 1)) {
                cudaEvent_t e;
                CUDA_OK(cudaEventCreate(&e));
                ev_.push_back(e);
            }
            CUDA_OK(cudaEventRecord(ev_[2 * ev_used_], s));
        }
        // enough buckets to fill the GPU (148 SMs x 4 blocks x 128 threads): one thread per bucket; fewer: share buckets
        if (nb > 4u * 32768u)
            k_accumulate<F><<<(nb + 127) / 128, 128, 0, s>>>(order_, offsets_, pairs_, table_, buckets_, nb);
        else if (nb > 32768u)
            k_accumulate_split<F, 2><<<(2 * nb + 127) / 128, 128, 0, s>>>(order_, offsets_, pairs_, table_, buckets_, nb);
        else
            k_accumulate_split<F, 4><<<(4 * nb + 127) / 128, 128, 0, s>>>(order_, offsets_, pairs_, table_, buckets_, nb);
        launches_++;
        if (timing_) {
            CUDA_OK(cudaEventRecord(ev_[2 * ev_used_ + 1], s));
            ev_used_++;
        }
        k_accumulate_overflow<F><<<296, OVER_THREADS, 0, s>>>(over_list_, order_hist_ + ORDER_BINS, offsets_, pairs_, table_, buckets_);
        launches_++;

        // running-sum passes while a group has more points than the bit sums should take, then the bit sums
        TailParams tp = plan_tail(groups);
        tp.W = cfg_.precompute ? 1 : W_;
        tp.c = cfg_.c;
        tp.groups = groups;
        const xyzz *in = buckets_;
        uint32_t N = nbw_;
        for (int l = 0; l < tp.L; l++) {
            const int m = 1 << tp.log_m[l];
            const uint32_t total_out = groups * (N / m);
            xyzz *Sl = lvlS_[l & 1];
            k_reduce_level<F><<<(total_out + 127) / 128, 128, 0, s>>>(in, Sl, lvlW_ + tp.w_off[l], total_out, m);
            launches_++;
            in = Sl;
            N /= m;
        }
        k_bit_sums<F><<<dim3(groups, tp.slots), 32, 0, s>>>(in, lvlW_, tp, sums_);
        if (tp.W == 1) {
            if (affine_out)
                k_finalize<F, true><<<nmsm, 32, 0, s>>>(sums_, tp, d_out);
            else
                k_finalize<F, false><<<nmsm, 32, 0, s>>>(sums_, tp, d_out);
            launches_ += 2;
        } else {
            xyzz *win = lvlS_[0];  // free again: the bit sums have consumed the last pass (>= groups entries: nbw >= 2)
            k_window_results<F><<<groups, 32, 0, s>>>(sums_, tp, win);
            if (affine_out)
                k_combine_windows<F, true><<<(nmsm + 31) / 32, 32, 0, s>>>(win, tp, d_out, nmsm);
            else
                k_combine_windows<F, false><<<(nmsm + 31) / 32, 32, 0, s>>>(win, tp, d_out, nmsm);
            launches_ += 3;
        }
        CUDA_OK(cudaGetLastError());
    }

    MsmConfig cfg_;
    uint32_t n_bases_ = 0, nbw_ = 0;
    int W_ = 0;
    const affine *table_ = nullptr;
    affine *table_own_ = nullptr;
    uint32_t *counters_ = nullptr, *offsets_ = nullptr, *tile_sums_ = nullptr, *pairs_ = nullptr, *err_ = nullptr;
    uint32_t *order_ = nullptr, *over_list_ = nullptr, *order_hist_ = nullptr;
    xyzz *buckets_ = nullptr, *lvlS_[2] = {nullptr, nullptr}, *lvlW_ = nullptr, *sums_ = nullptr;
    uint64_t cap_buckets_ = 0, cap_pairs_ = 0;
    uint32_t cap_groups_ = 0;
    size_t ws_bytes_ = 0;
    uint64_t launches_ = 0;
    bool timing_ = false;
    std::vector<cudaEvent_t> ev_;  // (start, stop) per chunk of the last run
    size_t ev_used_ = 0;
};

}  // namespace pasta
