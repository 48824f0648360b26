// The two drop-in entry points, their batch twins, and the staged pipeline behind them.
//
//   verify_mina_state_ffi         replaces AL/operator/mina/lib/src/lib.rs:41-113
//   verify_account_inclusion_ffi  replaces AL/operator/mina_account/lib/src/lib.rs:16-78
//
// A proof is accepted only if EVERY stage of the reference's check has run here and passed.  Stages
// that are not built yet (or whose constants are unavailable) are reported as `unavailable` and force
// a reject: the functions never return true on a partial check.  The per-stage outcome is exposed
// through mina_b200_last_stages() / mina_b200_verify_state_stages() so tests and the bench can say
// exactly what was verified.
//
// Device work per batch of state proofs:
//   16 + 30 prechallenges/proof --k_endo_to_field--> challenges --k_bpoly_tables--> 16 KiB tables/proof
//   mode PER_PROOF: one MSM per accumulator, scalars rebuilt from the tables inside the digit kernels
//   mode RLC:       S = sum_j r_j * b_poly_coefficients(chals_j)  (k_bpoly_combine), ONE MSM <S, G>,
//                   compared with sum_j r_j * C_j; bisection recovers per-proof bits on a mismatch
//                   (what poly-commitment's batch_dlog_accumulator_check does for a batch).
// Only ~1 KiB per proof crosses PCIe (prechallenges + accumulator points); the 2 MiB coefficient
// vectors never exist on the host.
#include <sys/random.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>

#include "../../include/mina_account_verifier.h"
#include "../../include/mina_b200.h"
#include "../../include/mina_verifier.h"
#include "consensus.hpp"
#include "context.cuh"
#include "group_testing.hpp"
#include "ipa.cuh"
#include "ipa_verify.cuh"
#include "poseidon.cuh"
#include "sol_account.hpp"
#include "wire.hpp"

namespace pasta {
using gt::COMBINE_SLICE;
using gt::LevelGroup;
using gt::LevelPlan;

// ---- persistent staging ---------------------------------------------------------------------------
struct SideBuffers {  // one accumulator family (Vesta k=16 or Pallas k=15)
    PinnedBuf<uint8_t> h_pre, h_pts, h_r, h_out;
    PinnedBuf<uint32_t> h_subset, h_bad, h_soff, h_status;
    DevBuf<uint8_t> d_pre;
    DevBuf<fe> d_chal, d_r_can, d_r, d_tab, d_S, d_partial;
    DevBuf<uint32_t> d_subset, d_pts_can, d_out_can, d_bad, d_sc, d_soff;
    DevBuf<affine> d_pts, d_res;
    DevBuf<xyzz> d_xyzz, d_scaled;
};
struct IpaBuffers {  // the batched IPA final check (one per curve)
    PinnedBuf<uint8_t> h_in, h_pts;
    DevBuf<uint8_t> d_in;
    DevBuf<fe> d_tab, d_t, d_chal, d_chal_c, d_rand, d_scalars;
    DevBuf<uint4> d_pre;
    DevBuf<uint32_t> d_pts_can;
    DevBuf<affine> d_pts;
    DevBuf<xyzz> d_terms;
};
struct VerifierState {
    SideBuffers side[2];  // index = curve id: 0 Pallas (step accumulators), 1 Vesta (wrap accumulator)
    IpaBuffers ipa[2];
    // Merkle fold
    DevBuf<MerkleNodeDev> d_nodes;
    DevBuf<uint32_t> d_depths;
    DevBuf<fe> d_leaves, d_roots, d_prefix, d_folded;
    DevBuf<uint8_t> d_ok;
    uint32_t prefix_depth = 0;
    cudaEvent_t fork_join[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // per pipeline: fork, join (timing disabled)
    ~VerifierState() {
        for (auto &pair : fork_join)
            for (cudaEvent_t &e : pair)
                if (e) cudaEventDestroy(e);
    }
};
void verifier_release(Context &c) {
    delete c.verifier;
    c.verifier = nullptr;
}
static VerifierState &vstate() {
    Context &c = ctx();
    if (!c.verifier) c.verifier = new VerifierState();
    return *c.verifier;
}

// ---- small host helpers ------------------------------------------------------------------------------
static void parallel_for(size_t n, const std::function<void(size_t)> &fn);
template <class F>
static bool canonical(const wire::B32 &b, host::Fe<F> &out) {
    return host::Fe<F>::from_bytes_le(b.data(), out);
}
template <class B>
static bool point_on_curve(const wire::Point &p) {
    host::Affine<B> a;
    if (!canonical<B>(p.x, a.x) || !canonical<B>(p.y, a.y)) return false;
    a.inf = false;
    return a.on_curve();  // (0,0) is not on y^2 = x^3 + 5, so the identity encoding is rejected too
}
static void random_bytes(uint8_t *out, size_t n) {
    size_t got = 0;
    while (got < n) {
        ssize_t r = getrandom(out + got, std::min<size_t>(n - got, 256), 0);  // <= 256 bytes never returns short
        if (r <= 0) throw std::runtime_error("getrandom failed");
        got += (size_t)r;
    }
}
// one non-zero 128-bit value, zero-extended to a 32-byte little-endian field element
static void random_128(uint8_t *out32) {
    std::memset(out32, 0, 32);
    for (;;) {
        random_bytes(out32, 16);
        uint64_t lo, hi;
        std::memcpy(&lo, out32, 8);
        std::memcpy(&hi, out32 + 8, 8);
        if (lo | hi) return;
    }
}
// n of them, one system call per 16 values
static void random_128_array(uint8_t *out32, size_t n) {
    uint8_t buf[256];
    for (size_t i = 0; i < n; i += 16) {
        const size_t cnt = std::min<size_t>(16, n - i);
        random_bytes(buf, 16 * cnt);
        for (size_t k = 0; k < cnt; k++) {
            uint8_t *o = out32 + 32 * (i + k);
            std::memset(o, 0, 32);
            std::memcpy(o, buf + 16 * k, 16);
            uint64_t lo, hi;
            std::memcpy(&lo, o, 8);
            std::memcpy(&hi, o + 8, 8);
            if (!(lo | hi)) random_128(o);
        }
    }
}

// ---- one accumulator family on the device ------------------------------------------------------------
// items: m accumulators, each = k prechallenges (16 B each) + one claimed commitment C (canonical, already
// validated on-curve).  ok[i] = ( <b_poly_coefficients(to_field(pre_i)), G[0..2^k)> == C_i ).
struct AccumulatorBatch {
    int curve;       // 0 Pallas / 1 Vesta
    int k;           // 15 / 16
    uint32_t m = 0;
    std::vector<uint8_t> pre;   // m * k * 16   (host input) ...
    std::vector<uint8_t> pts;   // m * 64
    const uint8_t *d_pre_ext = nullptr;  // ... or the same two arrays already resident in HBM
    const uint8_t *d_pts_ext = nullptr;
    std::vector<uint8_t> ok;    // m
};

static __global__ void __launch_bounds__(256) k_points_equal(const uint4 *__restrict__ a, const uint4 *__restrict__ b, uint32_t m,
                                                             uint8_t *__restrict__ ok) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t diff = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint4 x = a[4 * (size_t)i + k], y = b[4 * (size_t)i + k];
        diff |= (x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w);
    }
    ok[i] = diff == 0;
}

// res[i] (XYZZ) == claimed affine point (canonical bytes), without normalising res:
//   x = X/ZZ, y = Y/ZZZ  <=>  X == x*ZZ and Y == y*ZZZ   (the claimed point is never the identity)
template <class F>
static __global__ void __launch_bounds__(128) k_xyzz_equals_affine(const xyzz *__restrict__ res, const uint32_t *__restrict__ pts_can,
                                                                   uint32_t m, uint8_t *__restrict__ ok) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    xyzz r = res[i];
    fe x, y;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        x.v[k] = pts_can[(size_t)i * 16 + k];
        y.v[k] = pts_can[(size_t)i * 16 + 8 + k];
    }
    x = Fd<F>::to_mont(x);
    y = Fd<F>::to_mont(y);
    bool good = !fe_is_zero(r.zz) && fe_eq(r.x, Fd<F>::mul(x, r.zz)) && fe_eq(r.y, Fd<F>::mul(y, r.zzz));
    ok[i] = good ? 1 : 0;
}
// a[g] == b[g] for two XYZZ arrays (cross-multiplication; identity == identity)
template <class F>
static __global__ void __launch_bounds__(128) k_xyzz_pairs_equal(const xyzz *__restrict__ a, const xyzz *__restrict__ b, uint32_t n,
                                                                 uint8_t *__restrict__ ok) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz p = a[i], q = b[i];
    bool pi = fe_is_zero(p.zz), qi = fe_is_zero(q.zz);
    bool good;
    if (pi || qi)
        good = pi && qi;
    else
        good = fe_eq(Fd<F>::mul(p.x, q.zz), Fd<F>::mul(q.x, p.zz)) && fe_eq(Fd<F>::mul(p.y, q.zzz), Fd<F>::mul(q.y, p.zzz));
    ok[i] = good ? 1 : 0;
}

// One accumulator family runs on its own stream with its own statistics, so the Vesta and the Pallas
// pipelines of a batch can be driven by two host threads and overlap on the device (their MSM tails are
// latency-bound and leave most SMs idle).
struct AccRun {
    cudaStream_t s = nullptr, aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;  // owned by VerifierState (created once)
    bool timing = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    mina_b200_kernel_stats stats{0.f, 0.f, 0, 0, 0, 0};
};

struct AccDevice {  // device views of one prepared batch
    const uint8_t *d_pre = nullptr;
    const uint32_t *d_pts_can = nullptr;
};

// pipeline 0 = the caller's thread on the compute stream, 1 = the helper thread on the second stream
static void setup_run(Context &c, AccRun &rs, int pipeline, bool timing) {
    VerifierState &vs = vstate();
    for (cudaEvent_t &e : vs.fork_join[pipeline])
        if (!e) CTX_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    rs.s = pipeline == 0 ? c.stream : c.copy_stream;
    rs.aux = c.aux_stream[pipeline];
    rs.fork = vs.fork_join[pipeline][0];
    rs.join = vs.fork_join[pipeline][1];
    rs.timing = timing;
}

static AccDevice acc_prepare(Context &c, AccRun &rs, SideBuffers &sb, const AccumulatorBatch &ab) {
    const int field = ab.curve == 1 ? 0 : 1;  // scalar field of the curve
    const size_t npre = (size_t)ab.m * ab.k;
    AccDevice dv;
    if (ab.d_pre_ext) {
        dv.d_pre = ab.d_pre_ext;
        dv.d_pts_can = reinterpret_cast<const uint32_t *>(ab.d_pts_ext);
    } else {
        uint8_t *h_pre = sb.h_pre.reserve(npre * 16 + (size_t)ab.m * 64);
        std::memcpy(h_pre, ab.pre.data(), npre * 16);
        std::memcpy(h_pre + npre * 16, ab.pts.data(), (size_t)ab.m * 64);
        uint8_t *d_pre = sb.d_pre.reserve(npre * 16 + (size_t)ab.m * 64);
        CTX_CUDA_OK(cudaMemcpyAsync(d_pre, h_pre, npre * 16 + (size_t)ab.m * 64, cudaMemcpyHostToDevice, rs.s));
        dv.d_pre = d_pre;
        dv.d_pts_can = reinterpret_cast<const uint32_t *>(d_pre + npre * 16);
    }
    fe *d_chal = sb.d_chal.reserve(npre);
    launch_endo_to_field(field, dv.d_pre, d_chal, (uint32_t)npre, rs.s);
    c.launches += 1;
    return dv;
}

static void acc_per_proof(Context &c, AccRun &rs, SideBuffers &sb, AccumulatorBatch &ab) {
    const int field = ab.curve == 1 ? 0 : 1;
    CurveCtx &cc = c.curve[ab.curve];
    AccDevice dv = acc_prepare(c, rs, sb, ab);
    fe *d_tab = sb.d_tab.reserve((size_t)ab.m * BPOLY_TABLE);
    xyzz *d_res = sb.d_xyzz.reserve(ab.m);
    uint8_t *d_ok = reinterpret_cast<uint8_t *>(sb.d_subset.reserve((ab.m + 3) / 4));
    uint8_t *h_out = sb.h_out.reserve(ab.m);
    launch_bpoly_tables(field, sb.d_chal.p, d_tab, ab.m, ab.k, nullptr, true, rs.s);
    cc.fixed->enable_kernel_timing(rs.timing);
    cc.fixed->run_bpoly_xyzz(d_tab, ab.m, ab.k, d_res, rs.s);
    if (ab.curve == 0)
        k_xyzz_equals_affine<FpParams><<<(ab.m + 127) / 128, 128, 0, rs.s>>>(d_res, dv.d_pts_can, ab.m, d_ok);
    else
        k_xyzz_equals_affine<FqParams><<<(ab.m + 127) / 128, 128, 0, rs.s>>>(d_res, dv.d_pts_can, ab.m, d_ok);
    c.launches += 2;
    CTX_CUDA_OK(cudaMemcpyAsync(h_out, d_ok, ab.m, cudaMemcpyDeviceToHost, rs.s));
    if (cc.fixed->take_error(rs.s)) throw std::runtime_error("accumulator check: scalar overflow flagged by the MSM engine");
    if (rs.timing) {
        rs.stats.accumulate_ms += cc.fixed->last_accumulate_ms();
        rs.stats.msm_points += (uint64_t)ab.m << ab.k;
        rs.stats.msm_count += ab.m;
    }
    for (uint32_t i = 0; i < ab.m; i++) ab.ok[i] = h_out[i];
}

// sum of the kept slices of each group: out[g][i] = sum_{s in [gso[2g], gso[2g+1])} partial[s][i]  (canonical)
template <class S>
static __global__ void __launch_bounds__(256) k_sum_slices(const fe *__restrict__ partial, const uint32_t *__restrict__ gso, int k,
                                                           fe *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> k) return;
    uint32_t s0 = gso[2 * blockIdx.y], s1 = gso[2 * blockIdx.y + 1];
    fe acc = fe_zero();
    for (uint32_t s = s0; s < s1; s++) acc = Fd<S>::add(acc, partial[((size_t)s << k) + i]);
    out[((size_t)blockIdx.y << k) + i] = acc;
}

// ---- commitment side of the random linear combination ----------------------------------------------
// P_j = r_j * C_j once per batch (r_j < 2^128: plain double-and-add, one thread per point, left in XYZZ).
// It runs on the pipeline's auxiliary stream beside the table / combine / MSM work of level 0; every
// level then only needs sums of P_j over its groups (k_range_sums), not another MSM.
template <class F>
static __global__ void __launch_bounds__(64) k_scale_points_128(const affine *__restrict__ pts, const fe *__restrict__ r_can, uint32_t m,
                                                                xyzz *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const affine q = pts[i];
    const fe r = r_can[i];
    xyzz acc = Ec<F>::identity();
#pragma unroll 1
    for (int b = 127; b >= 0; b--) {
        acc = Ec<F>::dbl(acc);
        if ((r.v[b >> 5] >> (b & 31)) & 1u) Ec<F>::add_mixed(acc, q);
    }
    out[i] = acc;
}
// w * p for a small scalar (w < 2^31), XYZZ in and out
template <class F>
static __device__ __forceinline__ xyzz small_mul(const xyzz &p, uint32_t w) {
    xyzz acc = Ec<F>::identity();
#pragma unroll 1
    for (int b = 31 - __clz(w | 1u); b >= 0; b--) {
        acc = Ec<F>::dbl(acc);
        if ((w >> b) & 1u) Ec<F>::add(acc, p);
    }
    return acc;
}
// The locator weights: proof j of the batch weighs j + 1.  Pw_j = (j + 1) * P_j,  rw_j = (j + 1) * r_j.
template <class F>
static __global__ void __launch_bounds__(64) k_weight_points(const xyzz *__restrict__ P, uint32_t m, xyzz *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = small_mul<F>(P[i], i + 1u);
}
template <class S>
static __global__ void __launch_bounds__(128) k_weight_scalars(const fe *__restrict__ r_mont, uint32_t m, fe *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    fe w = fe_zero();
    w.v[0] = i + 1u;
    out[i] = Fd<S>::mul(r_mont[i], Fd<S>::to_mont(w));
}
template <class F>
static __device__ __forceinline__ xyzz shfl_down_xyzz(const xyzz &p, int delta) {
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
// block-wide sum of one XYZZ point per thread (result valid in thread 0)
static constexpr int SUBSET_THREADS = 128;
template <class F>
static __device__ __forceinline__ xyzz block_sum_xyzz(xyzz acc, xyzz *warp_part) {
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        xyzz o = shfl_down_xyzz<F>(acc, d);
        Ec<F>::add(acc, o);
    }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_part[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        acc = lane < SUBSET_THREADS / 32 ? warp_part[lane] : Ec<F>::identity();
#pragma unroll 1
        for (int d = SUBSET_THREADS / 64; d >= 1; d >>= 1) {
            xyzz o = shfl_down_xyzz<F>(acc, d);
            Ec<F>::add(acc, o);
        }
    }
    return acc;
}
// D[which][g] = sum_{j in [rng[2g], rng[2g+1])} P_which[j]: one block per (group, which); which = blockIdx.y
// selects the plain points (0) or the weighted ones (1); the outputs are `stride` apart.
template <class F>
static __global__ void __launch_bounds__(SUBSET_THREADS) k_range_sums(const xyzz *__restrict__ P0, const xyzz *__restrict__ P1,
                                                                      const uint32_t *__restrict__ rng, xyzz *__restrict__ out, uint32_t stride) {
    __shared__ xyzz warp_part[SUBSET_THREADS / 32];
    const xyzz *P = blockIdx.y ? P1 : P0;
    const uint32_t t0 = rng[2 * blockIdx.x], t1 = rng[2 * blockIdx.x + 1];
    xyzz acc = Ec<F>::identity();
    for (uint32_t t = t0 + threadIdx.x; t < t1; t += SUBSET_THREADS) Ec<F>::add(acc, P[t]);
    acc = block_sum_xyzz<F>(acc, warp_part);
    if (threadIdx.x == 0) out[(size_t)blockIdx.y * stride + blockIdx.x] = acc;
}
// The last child of every split group needs no MSM:  A_last = A_parent - sum of its siblings' A.
// derived[3*d + {0,1,2}] = parent index in `A_prev`, [sib_begin, sib_end) in `A` (this level's MSM results);
// the result goes to A[n_msm + d].  One block per (derived group, which); the plain and the weighted arrays
// are `stride` apart.
template <class F>
static __global__ void __launch_bounds__(SUBSET_THREADS) k_derive_last_child(const xyzz *__restrict__ A_prev, xyzz *__restrict__ A,
                                                                             const uint32_t *__restrict__ derived, uint32_t n_msm, uint32_t stride) {
    __shared__ xyzz warp_part[SUBSET_THREADS / 32];
    A_prev += (size_t)blockIdx.y * stride;
    A += (size_t)blockIdx.y * stride;
    const uint32_t parent = derived[3 * blockIdx.x], s0 = derived[3 * blockIdx.x + 1], s1 = derived[3 * blockIdx.x + 2];
    xyzz acc = Ec<F>::identity();
    for (uint32_t t = s0 + threadIdx.x; t < s1; t += SUBSET_THREADS) Ec<F>::add(acc, A[t]);
    acc = block_sum_xyzz<F>(acc, warp_part);
    if (threadIdx.x == 0) {
        acc.y = Fd<F>::neg(acc.y);
        xyzz p = A_prev[parent];
        Ec<F>::add(p, acc);
        A[n_msm + blockIdx.x] = p;
    }
}
template <class F>
static __global__ void k_negate_xyzz(xyzz *p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i].y = Fd<F>::neg(p[i].y);
}
template <class F>
static __device__ __forceinline__ bool xyzz_equal(const xyzz &p, const xyzz &q) {
    bool pi = fe_is_zero(p.zz), qi = fe_is_zero(q.zz);
    if (pi || qi) return pi && qi;
    return fe_eq(Fd<F>::mul(p.x, q.zz), Fd<F>::mul(q.x, p.zz)) && fe_eq(Fd<F>::mul(p.y, q.zzz), Fd<F>::mul(q.y, p.zzz));
}
// Verdict of one group from its two error sums
//   T0 = A0 - D0 = sum_{j in g} r_j e_j,   T1 = A1 - D1 = sum_{j in g} (j + 1) r_j e_j,   e_j = <s_j, G> - C_j:
//   T0 == 0                        -> every proof of the group is good                          (status 1)
//   T1 == w * T0, w - 1 in [a, b)  -> proof w - 1 is the ONLY bad one of the group               (status 2 + w - 1)
//   otherwise                      -> at least two bad proofs (or no locator at this level)      (status 0)
// Why the middle case is sound: with two or more e_j != 0 the relation sum_j (j + 1 - w) r_j e_j = 0 is, for each
// fixed w, a non-trivial linear equation in independent 128-bit r_j chosen after the proofs: probability
// <= 2^-128 per w, <= 2^-118 over the <= 1024 candidates.  (Single-error location in a batch, Law & Matt 2007.)
static constexpr int LOCATE_THREADS = 128;
template <class F>
static __global__ void __launch_bounds__(LOCATE_THREADS) k_locate(const xyzz *__restrict__ A, const xyzz *__restrict__ D, uint32_t stride,
                                                                  const uint32_t *__restrict__ rng, int with_locator,
                                                                  uint32_t *__restrict__ status) {
    __shared__ uint32_t found;
    const uint32_t g = blockIdx.x, a = rng[2 * g], b = rng[2 * g + 1];
    if (threadIdx.x == 0) found = 0;
    __syncthreads();
    xyzz T0 = A[g], nD = D[g];
    nD.y = Fd<F>::neg(nD.y);
    Ec<F>::add(T0, nD);
    if (Ec<F>::is_identity(T0)) {
        if (threadIdx.x == 0) status[g] = 1;
        return;
    }
    if (!with_locator) {
        if (threadIdx.x == 0) status[g] = 0;
        return;
    }
    xyzz T1 = A[stride + g];
    nD = D[stride + g];
    nD.y = Fd<F>::neg(nD.y);
    Ec<F>::add(T1, nD);
    // thread t tries w = a + 1 + t, then strides by the block size (one addition per further candidate)
    uint32_t w = a + 1u + threadIdx.x;
    if (w <= b) {
        xyzz X = small_mul<F>(T0, w);
        const xyzz step = (w + LOCATE_THREADS <= b) ? small_mul<F>(T0, LOCATE_THREADS) : Ec<F>::identity();
        for (;;) {
            if (xyzz_equal<F>(X, T1)) atomicMax(&found, w);
            w += LOCATE_THREADS;
            if (w > b) break;
            Ec<F>::add(X, step);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) status[g] = found ? 2u + (found - 1u) : 0u;
}

// ---- group testing over index ranges ----------------------------------------------------------------
// Every group is a contiguous range [a, b) of the batch.  Level 0 is the whole batch, combined in slices of
// COMBINE_SLICE proofs whose partial vectors are KEPT: a later group that is a union of whole slices gets its
// scalar vector by adding slices (k_sum_slices) instead of combining tables again.  If level 0 fails, every
// later level computes TWO sums per group (plain and locator-weighted, see k_locate), so a group with a single
// bad proof is resolved by one test instead of log2(size) halvings; a group with more is split (wide while few
// groups are open, so the launch set fills the GPU; at slice boundaries while groups are large).  The last
// child of every split needs no MSM (A_last = A_parent - siblings).
// What stays on the device for the whole batch
struct RlcBatch {
    fe *d_tab = nullptr, *d_tab_w = nullptr;          // product tables: hi scaled by r_j / by (j + 1) r_j
    fe *d_partial = nullptr, *d_partial_w = nullptr;  // level-0 slices of the two scalar vectors
    xyzz *d_scaled = nullptr, *d_scaled_w = nullptr;  // P_j, (j + 1) P_j
    uint32_t n_slices0 = 0;
    bool weighted = false;  // the *_w members are built (only after level 0 failed)
    // IPA: level 0 may take the commitment-side sum of the WHOLE batch from one Pippenger MSM over all per-opening
    // points (d_D0) and produce the per-item P_j only if level 0 fails (materialize enqueues that work on rs.s)
    const xyzz *d_D0 = nullptr;
    std::function<void()> materialize;
};

// One level of checks, ALL groups in one launch set.  status[g] as in k_locate.
static std::vector<uint32_t> acc_check_groups(Context &c, AccRun &rs, SideBuffers &sb, const AccumulatorBatch &ab, const RlcBatch &rb,
                                              const LevelPlan &plan, int level, bool locator) {
    const int field = ab.curve == 1 ? 0 : 1;
    CurveCtx &cc = c.curve[ab.curve];
    const uint32_t Gs = (uint32_t)plan.sliced.size(), Gc = (uint32_t)plan.combined.size(), Gm = Gs + Gc;
    const uint32_t Gd = (uint32_t)plan.derived.size(), G = Gm + Gd;
    const uint32_t nv = locator ? 2u : 1u;  // sums per group
    // one upload: slice ranges of the sliced groups | combine subset | combine group offsets | ranges of all groups |
    // derived triples
    std::vector<uint32_t> meta;
    const size_t o_sl = meta.size();
    for (auto &g : plan.sliced) {
        meta.push_back(g.a / COMBINE_SLICE);
        meta.push_back((g.b + COMBINE_SLICE - 1) / COMBINE_SLICE);
    }
    const size_t o_sub = meta.size();
    std::vector<uint32_t> coff{0};
    uint32_t nsub = 0;
    for (auto &g : plan.combined) {
        for (uint32_t j = g.a; j < g.b; j++) meta.push_back(j);
        nsub += g.size();
        coff.push_back(nsub);
    }
    const size_t o_coff = meta.size();
    meta.insert(meta.end(), coff.begin(), coff.end());
    const size_t o_rng = meta.size();
    for (uint32_t g = 0; g < G; g++) {
        meta.push_back(plan.at(g).a);
        meta.push_back(plan.at(g).b);
    }
    const size_t o_der = meta.size();
    for (auto &d : plan.derived_meta) meta.insert(meta.end(), d.begin(), d.end());
    uint32_t *h_meta = sb.h_subset.reserve(meta.size());
    std::memcpy(h_meta, meta.data(), meta.size() * 4);
    uint32_t *d_meta = sb.d_subset.reserve(meta.size());
    // XYZZ scratch, each array ab.m long (groups are disjoint and non-empty: G <= m):
    //   A even [plain | weighted], A odd [plain | weighted], D [plain | weighted], raw MSM output [2 m]
    const uint32_t stride = ab.m;
    xyzz *d_x = sb.d_xyzz.reserve(8 * (size_t)ab.m);
    xyzz *d_A = d_x + (size_t)(level & 1) * 2 * ab.m, *d_A_prev = d_x + (size_t)((level & 1) ^ 1) * 2 * ab.m;
    xyzz *d_D = d_x + 4 * (size_t)ab.m, *d_raw = d_x + 6 * (size_t)ab.m;
    uint32_t *d_status = sb.d_out_can.reserve(G);
    uint32_t *h_status = sb.h_status.reserve(G);
    fe *d_S = sb.d_S.reserve((size_t)std::max<uint32_t>(nv * Gm, 1) << ab.k);
    CTX_CUDA_OK(cudaMemcpyAsync(d_meta, h_meta, meta.size() * 4, cudaMemcpyHostToDevice, rs.s));
    // commitment side on the auxiliary stream, beside the MSM
    CTX_CUDA_OK(cudaEventRecord(rs.fork, rs.s));
    CTX_CUDA_OK(cudaStreamWaitEvent(rs.aux, rs.fork, 0));
    if (!locator && rb.d_D0)
        CTX_CUDA_OK(cudaMemcpyAsync(d_D, rb.d_D0, sizeof(xyzz), cudaMemcpyDeviceToDevice, rs.aux));
    else if (ab.curve == 0)
        k_range_sums<FpParams><<<dim3(G, nv), SUBSET_THREADS, 0, rs.aux>>>(rb.d_scaled, rb.d_scaled_w, d_meta + o_rng, d_D, stride);
    else
        k_range_sums<FqParams><<<dim3(G, nv), SUBSET_THREADS, 0, rs.aux>>>(rb.d_scaled, rb.d_scaled_w, d_meta + o_rng, d_D, stride);
    CTX_CUDA_OK(cudaEventRecord(rs.join, rs.aux));
    c.launches += 1;
    if (rs.timing) {
        if (!rs.ev[0]) {
            CTX_CUDA_OK(cudaEventCreate(&rs.ev[0]));
            CTX_CUDA_OK(cudaEventCreate(&rs.ev[1]));
        }
        CTX_CUDA_OK(cudaEventRecord(rs.ev[0], rs.s));
    }
    uint32_t combine_proofs = 0, combine_vectors = 0;
    if (Gc)
        for (uint32_t v = 0; v < nv; v++) {
            launch_bpoly_combine(field, v ? rb.d_tab_w : rb.d_tab, d_meta + o_sub, d_meta + o_coff, Gc, nsub, ab.k,
                                 d_S + ((size_t)(v * Gm + Gs) << ab.k), rs.s);
            c.launches += 1;
            combine_proofs += nsub;
            combine_vectors += Gc;
        }
    if (rs.timing) CTX_CUDA_OK(cudaEventRecord(rs.ev[1], rs.s));
    if (Gs)
        for (uint32_t v = 0; v < nv; v++) {
            dim3 grid(((1u << ab.k) + 255) / 256, Gs);
            const fe *src = v ? rb.d_partial_w : rb.d_partial;
            fe *dst = d_S + ((size_t)(v * Gm) << ab.k);
            if (field == 0)
                k_sum_slices<FpParams><<<grid, 256, 0, rs.s>>>(src, d_meta + o_sl, ab.k, dst);
            else
                k_sum_slices<FqParams><<<grid, 256, 0, rs.s>>>(src, d_meta + o_sl, ab.k, dst);
            c.launches += 1;
        }
    if (Gm) {
        cc.fixed->enable_kernel_timing(rs.timing);
        cc.fixed->run_xyzz(reinterpret_cast<const uint32_t *>(d_S), nv * Gm, 1u << ab.k, d_raw, rs.s);
        for (uint32_t v = 0; v < nv; v++)
            CTX_CUDA_OK(cudaMemcpyAsync(d_A + (size_t)v * stride, d_raw + (size_t)v * Gm, (size_t)Gm * sizeof(xyzz), cudaMemcpyDeviceToDevice, rs.s));
    }
    if (Gd) {
        if (ab.curve == 0)
            k_derive_last_child<FpParams><<<dim3(Gd, nv), SUBSET_THREADS, 0, rs.s>>>(d_A_prev, d_A, d_meta + o_der, Gm, stride);
        else
            k_derive_last_child<FqParams><<<dim3(Gd, nv), SUBSET_THREADS, 0, rs.s>>>(d_A_prev, d_A, d_meta + o_der, Gm, stride);
        c.launches += 1;
    }
    CTX_CUDA_OK(cudaStreamWaitEvent(rs.s, rs.join, 0));
    if (ab.curve == 0)
        k_locate<FpParams><<<G, LOCATE_THREADS, 0, rs.s>>>(d_A, d_D, stride, d_meta + o_rng, locator ? 1 : 0, d_status);
    else
        k_locate<FqParams><<<G, LOCATE_THREADS, 0, rs.s>>>(d_A, d_D, stride, d_meta + o_rng, locator ? 1 : 0, d_status);
    c.launches += 1;
    CTX_CUDA_OK(cudaMemcpyAsync(h_status, d_status, (size_t)G * 4, cudaMemcpyDeviceToHost, rs.s));
    uint32_t e = cc.fixed->take_error(rs.s);  // synchronises
    if (e) throw std::runtime_error("accumulator check: scalar overflow flagged by the MSM engine");
    if (rs.timing) {
        float ms = 0.f;
        CTX_CUDA_OK(cudaEventElapsedTime(&ms, rs.ev[0], rs.ev[1]));
        rs.stats.combine_ms += ms;
        if (Gm) rs.stats.accumulate_ms += cc.fixed->last_accumulate_ms();
        rs.stats.msm_points += (uint64_t)(nv * Gm) << ab.k;
        rs.stats.msm_count += nv * Gm;
        rs.stats.combine_proofs += combine_proofs;
        rs.stats.combine_vectors += combine_vectors;
    }
    return std::vector<uint32_t>(h_status, h_status + G);
}

// Level-0 slices of one scalar vector (subset == nullptr: proof j is table j)
static void combine_slices(Context &c, AccRun &rs, SideBuffers &sb, const AccumulatorBatch &ab, const fe *d_tab, fe *d_partial,
                           uint32_t n_slices0) {
    const int field = ab.curve == 1 ? 0 : 1;
    uint32_t *h_soff = sb.h_soff.reserve(n_slices0 + 1);
    for (uint32_t i = 0; i <= n_slices0; i++) h_soff[i] = std::min(ab.m, i * COMBINE_SLICE);
    uint32_t *d_soff = sb.d_soff.reserve(n_slices0 + 1);
    CTX_CUDA_OK(cudaMemcpyAsync(d_soff, h_soff, (size_t)(n_slices0 + 1) * 4, cudaMemcpyHostToDevice, rs.s));
    if (rs.timing) {
        if (!rs.ev[0]) {
            CTX_CUDA_OK(cudaEventCreate(&rs.ev[0]));
            CTX_CUDA_OK(cudaEventCreate(&rs.ev[1]));
        }
        CTX_CUDA_OK(cudaEventRecord(rs.ev[0], rs.s));
    }
    launch_bpoly_combine(field, d_tab, nullptr, d_soff, n_slices0, ab.m, ab.k, d_partial, rs.s);
    c.launches += 1;
    if (rs.timing) {
        CTX_CUDA_OK(cudaEventRecord(rs.ev[1], rs.s));
        CTX_CUDA_OK(cudaEventSynchronize(rs.ev[1]));
        float ms = 0.f;
        CTX_CUDA_OK(cudaEventElapsedTime(&ms, rs.ev[0], rs.ev[1]));
        rs.stats.combine_ms += ms;
        rs.stats.combine_proofs += ab.m;
        rs.stats.combine_vectors += n_slices0;
    }
}

static void rlc_levels(Context &c, AccRun &rs, SideBuffers &sb, AccumulatorBatch &ab, RlcBatch &rb, const fe *d_chal, fe *d_r,
                       const uint32_t *h_bad);
static constexpr uint32_t EAGER_LOCATOR_MAX = 256;  // batches up to this size compute the locator sums from level 0 on
static void rlc_buffers(SideBuffers &sb, const AccumulatorBatch &ab, RlcBatch &rb) {
    rb.n_slices0 = (ab.m + COMBINE_SLICE - 1) / COMBINE_SLICE;
    rb.d_tab = sb.d_tab.reserve(2 * (size_t)ab.m * BPOLY_TABLE);
    rb.d_tab_w = rb.d_tab + (size_t)ab.m * BPOLY_TABLE;
    rb.d_partial = sb.d_partial.reserve((size_t)2 * rb.n_slices0 << ab.k);
    rb.d_partial_w = rb.d_partial + ((size_t)rb.n_slices0 << ab.k);
    rb.d_scaled = sb.d_scaled.reserve(2 * (size_t)ab.m);
    rb.d_scaled_w = rb.d_scaled + ab.m;
}

// Group testing in levels (see the block comment above LevelGroup).  The common case -- every proof good -- ends
// after level 0: one combine + one MSM.
static void acc_rlc(Context &c, AccRun &rs, SideBuffers &sb, AccumulatorBatch &ab) {
    const int field = ab.curve == 1 ? 0 : 1;
    AccDevice dv = acc_prepare(c, rs, sb, ab);
    uint8_t *h_r = sb.h_r.reserve((size_t)ab.m * 32);
    random_128_array(h_r, ab.m);
    fe *d_r_can = sb.d_r_can.reserve(ab.m), *d_r = sb.d_r.reserve(2 * (size_t)ab.m);
    RlcBatch rb;
    rlc_buffers(sb, ab, rb);
    affine *d_pts = sb.d_pts.reserve(ab.m);
    uint32_t *d_bad = sb.d_bad.reserve(1);
    uint32_t *h_bad = sb.h_bad.reserve(1);
    CTX_CUDA_OK(cudaMemcpyAsync(d_r_can, h_r, (size_t)ab.m * 32, cudaMemcpyHostToDevice, rs.s));
    CTX_CUDA_OK(cudaMemsetAsync(d_bad, 0, 4, rs.s));
    launch_affine_to_mont_checked(ab.curve, dv.d_pts_can, d_pts, ab.m, d_bad, rs.s);
    CTX_CUDA_OK(cudaMemcpyAsync(h_bad, d_bad, 4, cudaMemcpyDeviceToHost, rs.s));
    // P_j = r_j C_j beside the tables and the combine of level 0 (the first k_range_sums is queued behind it)
    CTX_CUDA_OK(cudaEventRecord(rs.fork, rs.s));
    CTX_CUDA_OK(cudaStreamWaitEvent(rs.aux, rs.fork, 0));
    if (ab.curve == 0)
        k_scale_points_128<FpParams><<<(ab.m + 63) / 64, 64, 0, rs.aux>>>(d_pts, d_r_can, ab.m, rb.d_scaled);
    else
        k_scale_points_128<FqParams><<<(ab.m + 63) / 64, 64, 0, rs.aux>>>(d_pts, d_r_can, ab.m, rb.d_scaled);
    launch_fe_to_mont(field, d_r_can, d_r, ab.m, rs.s);
    c.launches += 3;
    rlc_levels(c, rs, sb, ab, rb, sb.d_chal.p, d_r, h_bad);
}

// The levels themselves, shared by the accumulator checks and the IPA final checks: item j contributes
// r_j * <b_poly_coefficients(chal_j), G[0..2^k)> on the g side (d_r: Montgomery, 2 m entries, the second half is
// scratch for the locator weights) and P_j on the other (rb.d_scaled, complete or in flight on rs.aux).
static void rlc_levels(Context &c, AccRun &rs, SideBuffers &sb, AccumulatorBatch &ab, RlcBatch &rb, const fe *d_chal, fe *d_r,
                       const uint32_t *h_bad) {
    const int field = ab.curve == 1 ? 0 : 1;
    launch_bpoly_tables(field, d_chal, rb.d_tab, ab.m, ab.k, d_r, false, rs.s);
    c.launches += 1;
    combine_slices(c, rs, sb, ab, rb.d_tab, rb.d_partial, rb.n_slices0);

    // The locator-weighted twins of the tables, slices and points.  Built after a failed level 0 -- or at once for a
    // small batch (a shard of a multi-GPU job), where they cost ~0.2 ms and save the whole extra level (~1.2 ms)
    // whenever something is bad.
    auto build_weighted = [&]() {
        if (rb.materialize) rb.materialize();  // per-item P_j that level 0 did without
        CTX_CUDA_OK(cudaEventRecord(rs.fork, rs.s));
        CTX_CUDA_OK(cudaStreamWaitEvent(rs.aux, rs.fork, 0));
        if (ab.curve == 0) {
            k_weight_points<FpParams><<<(ab.m + 63) / 64, 64, 0, rs.aux>>>(rb.d_scaled, ab.m, rb.d_scaled_w);
            k_weight_scalars<FqParams><<<(ab.m + 127) / 128, 128, 0, rs.s>>>(d_r, ab.m, d_r + ab.m);
        } else {
            k_weight_points<FqParams><<<(ab.m + 63) / 64, 64, 0, rs.aux>>>(rb.d_scaled, ab.m, rb.d_scaled_w);
            k_weight_scalars<FpParams><<<(ab.m + 127) / 128, 128, 0, rs.s>>>(d_r, ab.m, d_r + ab.m);
        }
        launch_bpoly_tables(field, d_chal, rb.d_tab_w, ab.m, ab.k, d_r + ab.m, false, rs.s);
        c.launches += 3;
        combine_slices(c, rs, sb, ab, rb.d_tab_w, rb.d_partial_w, rb.n_slices0);
        rb.weighted = true;
    };
    LevelPlan plan;
    plan.sliced.push_back(LevelGroup{0, ab.m, 0});
    bool locator = false;
    if (ab.m <= EAGER_LOCATOR_MAX && !rb.materialize && !rb.d_D0) {
        build_weighted();
        locator = true;
    }
    for (int level = 0; plan.size(); level++) {
        std::vector<uint32_t> status = acc_check_groups(c, rs, sb, ab, rb, plan, level, locator);
        // the stream has drained: the on-curve flag of the batch's points is on the host
        if (level == 0 && h_bad && *h_bad) throw std::runtime_error("accumulator check: a commitment is not a canonical curve point (callers validate first)");
        LevelPlan next;
        if (!locator) {  // level 0
            if (status[0] == 1) {
                for (uint32_t i = 0; i < ab.m; i++) ab.ok[i] = 1;
                return;
            }
            // something is bad: build the locator-weighted twins once, then test the whole batch again with both sums
            build_weighted();
            locator = true;
            next.sliced.push_back(LevelGroup{0, ab.m, 0});
            plan = std::move(next);
            continue;
        }
        // verdicts; unresolved groups are split next (group_testing.hpp)
        if (!gt::plan_next_level(plan, status, ab.m, ab.ok, next)) break;
        plan = std::move(next);
    }
}

// ---- batched IPA final check ------------------------------------------------------------------
// Host: validation, packing, U = to_group(t) (group map with the exact Tonelli-Shanks root, srs.hpp).  Device: the
// Fq-sponge transcript, endo challenges, scalars, the per-opening point sums B_i, and the group testing of the
// accumulator checks on   sgrb_i * <b_poly_coefficients(chal_i), G[0..2^k)> + B_i == 0.
template <class F, class S, bool SCALAR_LARGER>
static void ipa_verify_t(Context &c, int curve, const poseidon::Params<F> &table, const mina_b200_ipa_batch &b, uint8_t *ok) {
    using E = host::Fe<F>;
    using ES = host::Fe<S>;
    const uint32_t n = b.n, k = b.rounds, nc = b.n_comm, npts = b.n_points;
    const uint32_t npp = 2 * k + nc + 4;
    std::memset(ok, 0, n);
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    AccRun rs;
    setup_run(c, rs, 0, false);
    IpaBuffers &ib = vstate().ipa[curve];
    SideBuffers &sb = vstate().side[curve];
    CurveCtx &cc = c.curve[curve];
    const int sfield = curve == 1 ? 0 : 1;  // scalar field id
    // (1) everything the transcript and the scalar kernel read goes up unvalidated, so the sponge (the longest
    // dependent chain of the whole check) starts at once; validation runs on the host meanwhile and neutralises bad
    // openings through their randomisers (zero) and points (identity) before anything is summed across openings.
    // packed 32-byte words: state[3n] | cip[n] | lr[4kn] | delta[2n] | z1 | z2 | polyscale | evalscale | elm[npts n]
    const size_t w_state = 0, w_cip = w_state + 3 * (size_t)n, w_lr = w_cip + n, w_delta = w_lr + 4 * (size_t)k * n, w_z1 = w_delta + 2 * (size_t)n,
                 w_z2 = w_z1 + n, w_ps = w_z2 + n, w_es = w_ps + n, w_elm = w_es + n, w_end = w_elm + (size_t)npts * n;
    uint8_t *h_in = ib.h_in.reserve(32 * w_end);
    std::memcpy(h_in + 32 * w_state, b.sponge_state96, 96 * (size_t)n);
    std::memcpy(h_in + 32 * w_cip, b.cip32, 32 * (size_t)n);
    std::memcpy(h_in + 32 * w_lr, b.lr64, 128 * (size_t)k * n);
    std::memcpy(h_in + 32 * w_delta, b.delta64, 64 * (size_t)n);
    std::memcpy(h_in + 32 * w_z1, b.z1_32, 32 * (size_t)n);
    std::memcpy(h_in + 32 * w_z2, b.z2_32, 32 * (size_t)n);
    std::memcpy(h_in + 32 * w_ps, b.polyscale32, 32 * (size_t)n);
    std::memcpy(h_in + 32 * w_es, b.evalscale32, 32 * (size_t)n);
    std::memcpy(h_in + 32 * w_elm, b.eval_points32, 32 * (size_t)npts * n);
    const fe *d_in = reinterpret_cast<const fe *>(ib.d_in.reserve(32 * w_end));
    CTX_CUDA_OK(cudaMemcpyAsync(ib.d_in.p, h_in, 32 * w_end, cudaMemcpyHostToDevice, rs.s));
    auto tab = table.device_table();
    fe *d_tab = ib.d_tab.reserve(POSEIDON_TABLE_WORDS);
    CTX_CUDA_OK(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, rs.s));
    fe *d_t = ib.d_t.reserve(n);
    uint4 *d_pre = ib.d_pre.reserve((size_t)n * (k + 1));
    k_ipa_transcript<F, S, SCALAR_LARGER><<<(4 * n + 127) / 128, 128, 0, rs.s>>>(d_in + w_state, b.sponge_mode, b.sponge_count, d_in + w_cip, d_in + w_lr,
                                                                             d_in + w_delta, n, (int)k, d_tab, d_t, d_pre, d_pre + (size_t)n * k);
    fe *d_chal = ib.d_chal.reserve((size_t)n * k), *d_chal_c = ib.d_chal_c.reserve(n);
    launch_endo_to_field(sfield, d_pre, d_chal, n * k, rs.s);
    launch_endo_to_field(sfield, d_pre + (size_t)n * k, d_chal_c, n, rs.s);
    c.launches += 3;
    // (2) host validation: canonical field elements, points on the curve ((0,0) = identity allowed for commitments only)
    std::vector<uint8_t> valid(n, 1);
    auto point_ok = [](const uint8_t *p, bool allow_identity) {
        host::Affine<F> a;
        if (!E::from_bytes_le(p, a.x) || !E::from_bytes_le(p + 32, a.y)) return false;
        if (a.x.is_zero() && a.y.is_zero()) return allow_identity;
        a.inf = false;
        return a.on_curve();
    };
    // points: sg, U (written on the device), L_0, R_0, ..., C_0 .., delta, H; randomisers rb | sgrb
    uint8_t *h_pts = ib.h_pts.reserve(64 * (size_t)n * npp + 64 * (size_t)n);
    uint8_t *h_rand = h_pts + 64 * (size_t)n * npp;
    const uint8_t *h_can = cc.host_canonical.data() + 64 * (size_t)cc.depth;  // h
    parallel_for(n, [&](size_t i) {
        bool good = true;
        E tmp;
        ES ts;
        for (int q = 0; q < 3; q++) good = good && E::from_bytes_le(b.sponge_state96 + 96 * i + 32 * q, tmp);
        for (const uint8_t *arr : {b.cip32, b.polyscale32, b.evalscale32, b.z1_32, b.z2_32}) good = good && ES::from_bytes_le(arr + 32 * i, ts);
        for (uint32_t t = 0; t < npts; t++) good = good && ES::from_bytes_le(b.eval_points32 + 32 * (i * npts + t), ts);
        good = good && point_ok(b.delta64 + 64 * i, false) && point_ok(b.sg64 + 64 * i, false);
        for (uint32_t j = 0; j < 2 * k && good; j++) good = point_ok(b.lr64 + 64 * (i * 2 * k + j), false);
        for (uint32_t m = 0; m < nc && good; m++) good = point_ok(b.commitments64 + 64 * (i * nc + m), true);
        valid[i] = good;
        uint8_t *o = h_pts + 64 * i * npp;
        std::memset(h_rand + 32 * i, 0, 32);
        std::memset(h_rand + 32 * (n + i), 0, 32);
        if (!good) {
            std::memset(o, 0, 64 * (size_t)npp);
            return;
        }
        std::memcpy(o, b.sg64 + 64 * i, 64);
        std::memset(o + 64, 0, 64);
        std::memcpy(o + 64 * 2, b.lr64 + 64 * (i * 2 * k), 128 * (size_t)k);
        std::memcpy(o + 64 * (2 + 2 * (size_t)k), b.commitments64 + 64 * (i * nc), 64 * (size_t)nc);
        std::memcpy(o + 64 * (2 + 2 * (size_t)k + nc), b.delta64 + 64 * i, 64);
        std::memcpy(o + 64 * (3 + 2 * (size_t)k + nc), h_can, 64);
        random_128(h_rand + 32 * i);
        random_128(h_rand + 32 * (n + i));
    });
    uint32_t *d_pts_can = ib.d_pts_can.reserve(16 * (size_t)n * npp + 16 * (size_t)n);
    affine *d_pts = ib.d_pts.reserve((size_t)n * npp);
    xyzz *d_terms = ib.d_terms.reserve((size_t)n * npp);
    CTX_CUDA_OK(cudaMemcpyAsync(d_pts_can, h_pts, 64 * (size_t)n * npp + 64 * (size_t)n, cudaMemcpyHostToDevice, rs.s));
    launch_affine_to_mont(curve, d_pts_can, d_pts, n * npp, rs.s);
    {
        host::GroupMap<F> hm;
        GroupMapConsts gm;
        std::memcpy(&gm.sqrt_neg_three, hm.sqrt_neg_three_u2.l, 32);
        std::memcpy(&gm.sqrt_neg_three_minus_one_over_two, hm.sqrt_neg_three_u2_minus_u_over_2.l, 32);
        std::memcpy(&gm.inv_three, hm.inv_three_u2.l, 32);
        k_to_group<F><<<(n + 63) / 64, 64, 0, rs.s>>>(d_t, n, gm, d_pts, npp, 1);
    }
    fe *d_rand = ib.d_rand.reserve(2 * (size_t)n);  // rb | sgrb, Montgomery (zero for an invalid opening)
    launch_fe_to_mont(sfield, reinterpret_cast<const fe *>(d_pts_can + 16 * (size_t)n * npp), d_rand, 2 * n, rs.s);
    fe *d_scalars = ib.d_scalars.reserve((size_t)n * npp);
    k_ipa_scalars<S><<<(n + 63) / 64, 64, 0, rs.s>>>(d_chal, d_chal_c, d_in + w_z1, d_in + w_z2, d_in + w_cip, d_in + w_ps, d_in + w_es, d_in + w_elm,
                                                     d_rand, d_rand + n, n, (int)k, nc, npts, d_scalars);
    c.launches += 4;
    AccumulatorBatch ab;
    ab.curve = curve;
    ab.k = (int)k;
    ab.m = n;
    ab.ok.assign(n, 0);
    RlcBatch rbt;
    rlc_buffers(sb, ab, rbt);
    // (3) Level 0 needs only sum_i B_i: ONE Pippenger MSM over all n * npp points (a few 10^5 additions).  The
    // per-opening B_i (npp full-width scalar multiplications each, ~30x the work) are produced only if level 0 fails.
    // Window width: the signed-digit recoding leaves only 254 - c (W - 1) bits (+ a carry) for the top window, so for
    // most c nearly every scalar hits a handful of top-window buckets; c = 16 (14 bits left, no carry window) and
    // c = 8 (6 bits) are the widths where the top window stays spread out.
    MsmConfig vcfg;
    vcfg.precompute = false;
    vcfg.c = n * npp >= (1u << 14) ? 16 : 8;
    // on the auxiliary stream, beside the tables / combine / MSM of the g side
    CTX_CUDA_OK(cudaEventRecord(rs.fork, rs.s));
    CTX_CUDA_OK(cudaStreamWaitEvent(rs.aux, rs.fork, 0));
    cc.var->set_bases(d_pts, n * npp, vcfg, rs.aux);
    xyzz *d_total = d_terms;  // slot 0 of the term scratch; the terms themselves only exist after a failed level 0
    cc.var->run_xyzz(reinterpret_cast<const uint32_t *>(d_scalars), 1, n * npp, d_total, rs.aux);
    k_negate_xyzz<F><<<1, 32, 0, rs.aux>>>(d_total, 1);
    c.launches += 2;
    rbt.d_D0 = d_total;
    xyzz *d_scaled = rbt.d_scaled;
    rbt.materialize = [&c, &rs, d_pts, d_scalars, d_terms, d_scaled, n, npp]() {
        k_ipa_point_terms<F><<<(n * npp + 63) / 64, 64, 0, rs.s>>>(d_pts, d_scalars, n * npp, d_terms);
        k_ipa_sum_terms<F><<<n, 32, 0, rs.s>>>(d_terms, npp, d_scaled);
        c.launches += 2;
    };
    fe *d_r = sb.d_r.reserve(2 * (size_t)n);
    CTX_CUDA_OK(cudaMemcpyAsync(d_r, d_rand + n, 32 * (size_t)n, cudaMemcpyDeviceToDevice, rs.s));
    rlc_levels(c, rs, sb, ab, rbt, d_chal, d_r, nullptr);
    if (cc.var->take_error(rs.aux)) throw std::runtime_error("ipa_verify: scalar overflow flagged by the MSM engine");
    for (uint32_t i = 0; i < n; i++) ok[i] = valid[i] ? ab.ok[i] : 0;
}

static void run_accumulators(Context &c, AccRun &rs, AccumulatorBatch &ab, int mode) {
    ab.ok.assign(ab.m, 0);
    if (ab.m == 0) return;
    SideBuffers &sb = vstate().side[ab.curve];
    if (mode == MINA_B200_MODE_RLC && ab.m > 1)
        acc_rlc(c, rs, sb, ab);
    else
        acc_per_proof(c, rs, sb, ab);
}

// Both accumulator families of a batch, concurrently: the caller's thread drives the wrap (Vesta) side on the
// compute stream, a helper thread drives the step (Pallas) side on the second stream.  Holds the device lock.
static void run_both_sides(Context &c, AccumulatorBatch &wrap, AccumulatorBatch &step, int mode, bool timing,
                           mina_b200_kernel_stats *stats_wrap, mina_b200_kernel_stats *stats_step) {
    AccRun rw, rp;
    setup_run(c, rw, 0, timing);  // also creates the staging object before two threads race for it
    setup_run(c, rp, 1, timing);
    std::exception_ptr err;
    std::thread helper([&]() {
        try {
            if (cudaSetDevice(c.device) != cudaSuccess) throw std::runtime_error("cudaSetDevice failed in helper thread");
            run_accumulators(c, rp, step, mode);
        } catch (...) {
            err = std::current_exception();
        }
    });
    try {
        run_accumulators(c, rw, wrap, mode);
    } catch (...) {
        helper.join();
        for (int i = 0; i < 2; i++) {
            if (rw.ev[i]) cudaEventDestroy(rw.ev[i]);
            if (rp.ev[i]) cudaEventDestroy(rp.ev[i]);
        }
        throw;
    }
    helper.join();
    for (int i = 0; i < 2; i++) {
        if (rw.ev[i]) cudaEventDestroy(rw.ev[i]);
        if (rp.ev[i]) cudaEventDestroy(rp.ev[i]);
    }
    if (err) std::rethrow_exception(err);
    if (stats_wrap) *stats_wrap = rw.stats;
    if (stats_step) *stats_step = rp.stats;
}

// ---- state proofs -------------------------------------------------------------------------------------
struct StateJob {
    const uint8_t *proof = nullptr, *pub = nullptr;
    size_t proof_len = 0, pub_len = 0;
    mina_b200_stage_report rep{0, 0, 0};
    bool accept = false;
    bool host_done = false;
    // filled by the host pass for the device pass
    bool want_wrap_acc = false, want_step_acc = false;
    uint8_t pre_wrap[16 * 16];
    uint8_t pre_step[2][15 * 16];
    uint8_t pt_wrap[64], pt_step[2][64];
};

static const uint32_t STATE_ALL = MINA_B200_STAGE_LENGTHS | MINA_B200_STAGE_DECODE_PROOF | MINA_B200_STAGE_DECODE_PUB |
                                  MINA_B200_STAGE_PUB_STRUCTURE | MINA_B200_STAGE_PUB_HASHES | MINA_B200_STAGE_CONSENSUS |
                                  MINA_B200_STAGE_ACCUMULATOR | MINA_B200_STAGE_STEP_ACCUMULATORS | MINA_B200_STAGE_KIMCHI;

static void put_u128(uint8_t *dst, const wire::U128 &v) {
    std::memcpy(dst, &v.lo, 8);
    std::memcpy(dst + 8, &v.hi, 8);
}

// Everything that needs no device: lengths, bincode, the structural half of check_pub_inputs, consensus.
static void state_host_pass(StateJob &j) {
    if (j.host_done) return;
    j.host_done = true;
    auto pass = [&](uint32_t s) { j.rep.passed |= s; };
    auto fail = [&](uint32_t s) { j.rep.failed |= s; };
    auto unavailable = [&](uint32_t s) { j.rep.unavailable |= s; };
    // lib.rs:48-56
    if (j.proof_len > wire::MAX_STATE_PROOF_SIZE || j.pub_len > wire::MAX_PUB_INPUT_SIZE || !j.proof || !j.pub) return fail(MINA_B200_STAGE_LENGTHS);
    pass(MINA_B200_STAGE_LENGTHS);
    // lib.rs:58-71
    auto proof = std::make_unique<wire::StateProof>();
    wire::StatePubInputs pub;
    std::string err;
    if (!wire::decode_state_proof(j.proof, j.proof_len, *proof, err)) return fail(MINA_B200_STAGE_DECODE_PROOF);
    pass(MINA_B200_STAGE_DECODE_PROOF);
    if (!wire::decode_state_pub(j.pub, j.pub_len, pub, err)) return fail(MINA_B200_STAGE_DECODE_PUB);
    pass(MINA_B200_STAGE_DECODE_PUB);

    // check_pub_inputs (lib.rs:117-216).  The 17 Poseidon state hashes (lib.rs:128-160,183-193) need the
    // protocol-state ROInput packing and a trusted Poseidon table: not built.
    unavailable(MINA_B200_STAGE_PUB_HASHES);
    {   // the parts that are plain comparisons: ledger hashes (lib.rs:163-180) and the two to_fp()
        // conversions (lib.rs:183-186, 202-209), which fail on a non-canonical field element
        bool ok = true;
        for (int i = 0; i < wire::FRONTIER_LEN; i++)
            ok = ok && pub.candidate_chain_ledger_hashes[i] ==
                           proof->candidate_chain_states[i].blockchain_state.target.first_pass_ledger;
        host::Fp tmp;
        ok = ok && canonical<FpParams>(pub.bridge_tip_state_hash, tmp);
        ok = ok && canonical<FpParams>(pub.candidate_chain_state_hashes[wire::FRONTIER_LEN - 1], tmp);
        ok ? pass(MINA_B200_STAGE_PUB_STRUCTURE) : fail(MINA_B200_STAGE_PUB_STRUCTURE);
    }
    // consensus (lib.rs:83-94): candidate tip = last chain state, against the bridge's tip
    {
        consensus::ChainResult res = consensus::ChainResult::Bridge;
        consensus::Status st = consensus::select_secure_chain(proof->candidate_chain_states[wire::FRONTIER_LEN - 1],
                                                              proof->bridge_tip_state, consensus::StateHashCmp(), res);
        if (st == consensus::Status::NeedStateHash)
            unavailable(MINA_B200_STAGE_CONSENSUS);  // exact tie on height and VRF digest: needs Poseidon state hashes
        else if (st == consensus::Status::Ok && res == consensus::ChainResult::Candidate)
            pass(MINA_B200_STAGE_CONSENSUS);
        else
            fail(MINA_B200_STAGE_CONSENSUS);
    }
    // verify_block (lib.rs:99-111): accumulator_check is built; kimchi verify (to_batch + IPA) is not.
    unavailable(MINA_B200_STAGE_KIMCHI);
    const wire::PicklesProof &p = proof->candidate_tip_proof;
    if (point_on_curve<FqParams>(p.wrap_challenge_polynomial_commitment)) {
        j.want_wrap_acc = true;
        for (int i = 0; i < 16; i++) put_u128(j.pre_wrap + 16 * i, p.bulletproof_challenges[i]);
        std::memcpy(j.pt_wrap, p.wrap_challenge_polynomial_commitment.x.data(), 32);
        std::memcpy(j.pt_wrap + 32, p.wrap_challenge_polynomial_commitment.y.data(), 32);
    } else {
        fail(MINA_B200_STAGE_ACCUMULATOR);
    }
    // the wrap proof's two previous-challenge accumulators (Pallas side); blockchain proofs are N2
    if (p.step_challenge_polynomial_commitments.size() == 2 && point_on_curve<FpParams>(p.step_challenge_polynomial_commitments[0]) &&
        point_on_curve<FpParams>(p.step_challenge_polynomial_commitments[1])) {
        j.want_step_acc = true;
        for (int k = 0; k < 2; k++) {
            for (int i = 0; i < 15; i++) put_u128(j.pre_step[k] + 16 * i, p.wrap_old_bulletproof_challenges[k][i]);
            std::memcpy(j.pt_step[k], p.step_challenge_polynomial_commitments[k].x.data(), 32);
            std::memcpy(j.pt_step[k] + 32, p.step_challenge_polynomial_commitments[k].y.data(), 32);
        }
    } else {
        fail(MINA_B200_STAGE_STEP_ACCUMULATORS);
    }
}

// Host passes (bincode decode, on-curve checks, packing) run on a persistent pool: spawning 16 threads per batch cost
// about as much as decoding it.  One batch at a time uses the pool; a second concurrent caller (rare: the coalescer
// already merges callers) falls back to its own short-lived threads.  Exceptions are carried back to the caller.
class HostPool {
   public:
    static HostPool &get() {
        static HostPool *p = new HostPool;  // never destroyed: its detached workers outlive static destruction
        return *p;
    }
    bool try_run(size_t n, const std::function<void(size_t)> &fn) {
        std::unique_lock<std::mutex> own(owner_, std::try_to_lock);
        if (!own.owns_lock() || workers_.empty()) return false;
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            pending_ = workers_.size();
            err_ = nullptr;
            generation_++;
        }
        cv_.notify_all();
        work();  // the caller takes its share
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
        if (err_) std::rethrow_exception(err_);
        return true;
    }

   private:
    HostPool() {
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned nw = std::min(hw, 32u);
        for (unsigned t = 1; t < nw; t++) workers_.emplace_back([this] { loop(); });
        for (auto &w : workers_) w.detach();  // process-lifetime workers
    }
    void work() {
        try {
            for (size_t i; (i = next_.fetch_add(1)) < n_;) (*fn_)(i);
        } catch (...) {
            std::lock_guard<std::mutex> lk(m_);
            if (!err_) err_ = std::current_exception();
            next_.store(n_);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
            }
            work();
            std::lock_guard<std::mutex> lk(m_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::mutex owner_, m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t n_ = 0, pending_ = 0;
    std::atomic<size_t> next_{0};
    uint64_t generation_ = 0;
    std::exception_ptr err_;
};

static void parallel_for(size_t n, const std::function<void(size_t)> &fn) {
    if (n < 8) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    if (HostPool::get().try_run(n, fn)) return;
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    size_t nthreads = std::min<size_t>(hw, (n + 3) / 4);
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex em;
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nthreads; t++)
        pool.emplace_back([&]() {
            try {
                for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
            } catch (...) {
                std::lock_guard<std::mutex> lk(em);
                if (!err) err = std::current_exception();
                next.store(n);
            }
        });
    for (auto &th : pool) th.join();
    if (err) std::rethrow_exception(err);
}

// `accumulators_only`: skip the pub-input / consensus stages (mina_b200_accumulator_check*)
static void verify_state_jobs(StateJob *jobs, size_t n, int mode) {
    parallel_for(n, [&](size_t i) { state_host_pass(jobs[i]); });
    AccumulatorBatch wrap, step;
    wrap.curve = 1;
    wrap.k = 16;
    step.curve = 0;
    step.k = 15;
    std::vector<size_t> wrap_owner, step_owner;
    for (size_t i = 0; i < n; i++) {
        StateJob &j = jobs[i];
        if (j.want_wrap_acc) {
            wrap.pre.insert(wrap.pre.end(), j.pre_wrap, j.pre_wrap + sizeof j.pre_wrap);
            wrap.pts.insert(wrap.pts.end(), j.pt_wrap, j.pt_wrap + 64);
            wrap_owner.push_back(i);
        }
        if (j.want_step_acc)
            for (int k = 0; k < 2; k++) {
                step.pre.insert(step.pre.end(), j.pre_step[k], j.pre_step[k] + sizeof j.pre_step[k]);
                step.pts.insert(step.pts.end(), j.pt_step[k], j.pt_step[k] + 64);
                step_owner.push_back(i);
            }
    }
    wrap.m = (uint32_t)wrap_owner.size();
    step.m = (uint32_t)step_owner.size();
    if (wrap.m || step.m) {
        require_ready();
        Context &c = ctx();
        std::lock_guard<std::mutex> lk(c.mu);
        CTX_CUDA_OK(cudaSetDevice(c.device));
        run_both_sides(c, wrap, step, mode, false, nullptr, nullptr);
    }
    for (size_t t = 0; t < wrap_owner.size(); t++) {
        StateJob &j = jobs[wrap_owner[t]];
        (wrap.ok[t] ? j.rep.passed : j.rep.failed) |= MINA_B200_STAGE_ACCUMULATOR;
    }
    for (size_t t = 0; t + 1 < step_owner.size(); t += 2) {
        StateJob &j = jobs[step_owner[t]];
        ((step.ok[t] && step.ok[t + 1]) ? j.rep.passed : j.rep.failed) |= MINA_B200_STAGE_STEP_ACCUMULATORS;
    }
    for (size_t i = 0; i < n; i++)
        jobs[i].accept = jobs[i].rep.failed == 0 && jobs[i].rep.unavailable == 0 && jobs[i].rep.passed == STATE_ALL;
}

// ---- account proofs -------------------------------------------------------------------------------------
struct AccountJob {
    const uint8_t *proof = nullptr, *pub = nullptr;
    size_t proof_len = 0, pub_len = 0;
    mina_b200_stage_report rep{0, 0, 0};
    bool accept = false;
};
static const uint32_t ACCOUNT_ALL = MINA_B200_STAGE_LENGTHS | MINA_B200_STAGE_DECODE_PROOF | MINA_B200_STAGE_DECODE_PUB |
                                    MINA_B200_STAGE_ACCOUNT_ABI | MINA_B200_STAGE_ACCOUNT_LEAF | MINA_B200_STAGE_MERKLE;

static void account_host_pass(AccountJob &j) {
    auto pass = [&](uint32_t s) { j.rep.passed |= s; };
    auto fail = [&](uint32_t s) { j.rep.failed |= s; };
    auto unavailable = [&](uint32_t s) { j.rep.unavailable |= s; };
    if (j.proof_len > wire::MAX_ACCOUNT_PROOF_SIZE || j.pub_len > wire::MAX_PUB_INPUT_SIZE || !j.proof || !j.pub) return fail(MINA_B200_STAGE_LENGTHS);
    pass(MINA_B200_STAGE_LENGTHS);
    wire::AccountProof proof;
    wire::AccountPubInputs pub;
    std::string err;
    if (!wire::decode_account_proof(j.proof, j.proof_len, proof, err)) return fail(MINA_B200_STAGE_DECODE_PROOF);
    // o1_utils SerdeAs / ark-serialize reject field elements >= p while deserialising
    host::Fp tmp;
    for (auto &node : proof.merkle_path)
        if (!canonical<FpParams>(node.hash, tmp)) return fail(MINA_B200_STAGE_DECODE_PROOF);
    pass(MINA_B200_STAGE_DECODE_PROOF);
    if (!wire::decode_account_pub(j.pub, j.pub_len, pub, err) || !canonical<FpParams>(pub.ledger_hash, tmp)) return fail(MINA_B200_STAGE_DECODE_PUB);
    pass(MINA_B200_STAGE_DECODE_PUB);
    // mina_account/lib/src/lib.rs:54-66: the Solidity ABI encoding of the proof's account must equal the
    // public input's `encoded_account` byte for byte (conversion failure -> reject)
    {
        std::vector<uint8_t> expected;
        if (!sol::abi_encode_account(proof.account, expected) || expected != pub.encoded_account) return fail(MINA_B200_STAGE_ACCOUNT_ABI);
        pass(MINA_B200_STAGE_ACCOUNT_ABI);
    }
    // :70 Account::hash (mina-tree ROInput packing + Poseidon, un-vendored) is not built.  The Merkle fold
    // (merkle_verifier.rs:9-35) exists as a kernel (k_merkle_fold) but needs the leaf hash and a trusted
    // Poseidon table.
    unavailable(MINA_B200_STAGE_ACCOUNT_LEAF);
    unavailable(MINA_B200_STAGE_MERKLE);
}

static void verify_account_jobs(AccountJob *jobs, size_t n) {
    parallel_for(n, [&](size_t i) { account_host_pass(jobs[i]); });
    for (size_t i = 0; i < n; i++)
        jobs[i].accept = jobs[i].rep.failed == 0 && jobs[i].rep.unavailable == 0 && jobs[i].rep.passed == ACCOUNT_ALL;
}

// ---- coalescing of concurrent single-proof callers (leader / follower) ---------------------------------
// The operator calls the FFI from one goroutine per proof (AL/operator/pkg/operator.go:448-454).  The
// first caller to arrive becomes the leader and verifies everything queued so far as ONE batch; callers
// that arrive meanwhile queue up and are served by the next leader.
template <class Job>
class Coalescer {
   public:
    using BatchFn = void (*)(Job *, size_t);
    explicit Coalescer(BatchFn fn) : fn_(fn) {}
    void submit(Job &job) {
        std::unique_lock<std::mutex> lk(m_);
        Ticket t{&job, false};
        pending_.push_back(&t);
        while (!t.done) {
            if (!leader_) {
                leader_ = true;
                std::vector<Ticket *> batch;
                batch.swap(pending_);
                lk.unlock();
                std::vector<Job> jobs;
                jobs.reserve(batch.size());
                for (Ticket *b : batch) jobs.push_back(*b->job);
                bool failed = false;
                try {
                    fn_(jobs.data(), jobs.size());
                } catch (...) {
                    failed = true;
                }
                lk.lock();
                for (size_t i = 0; i < batch.size(); i++) {
                    if (failed) {
                        jobs[i].accept = false;
                        jobs[i].rep.failed |= MINA_B200_STAGE_INTERNAL_ERROR;
                    }
                    *batch[i]->job = jobs[i];
                    batch[i]->done = true;
                }
                leader_ = false;
                cv_.notify_all();
            } else {
                cv_.wait(lk);
            }
        }
    }

   private:
    struct Ticket {
        Job *job;
        bool done;
    };
    BatchFn fn_;
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<Ticket *> pending_;
    bool leader_ = false;
};

static int default_mode() {
    const char *e = std::getenv("MINA_B200_MODE");
    if (e && std::string(e) == "per_proof") return MINA_B200_MODE_PER_PROOF;
    return MINA_B200_MODE_RLC;
}
static void state_batch_default(StateJob *jobs, size_t n) { verify_state_jobs(jobs, n, default_mode()); }
static Coalescer<StateJob> g_state_queue(state_batch_default);
static Coalescer<AccountJob> g_account_queue(verify_account_jobs);

static thread_local mina_b200_stage_report g_last_report{0, 0, 0};

// Lazy initialisation, like the reference's lazy_static (lib.rs:23-35).  Returns false when no device
// context can be created; the caller then rejects (there is no CPU path).
static bool ensure_init() {
    if (ctx().ready) return true;
    int device = 0;
    if (const char *e = std::getenv("MINA_B200_DEVICE")) device = std::atoi(e);
    return mina_b200_init(device, nullptr) == 0;
}

}  // namespace pasta

using namespace pasta;

extern "C" {

int mina_verifier_init(const char *data_dir, int device) { return mina_b200_init(device, data_dir); }
void mina_verifier_shutdown(void) { mina_b200_shutdown(); }

bool verify_mina_state_ffi(const unsigned char *proof_buffer, size_t proof_len, const unsigned char *pub_input_buffer,
                           size_t pub_input_len) {
    StateJob j;
    j.proof = proof_buffer;
    j.proof_len = proof_len;
    j.pub = pub_input_buffer;
    j.pub_len = pub_input_len;
    try {
        // Oversize / undecodable inputs are rejected before (and without) touching the device, exactly
        // like the reference's early returns (lib.rs:48-71).
        state_host_pass(j);
        if (j.rep.failed & (MINA_B200_STAGE_LENGTHS | MINA_B200_STAGE_DECODE_PROOF | MINA_B200_STAGE_DECODE_PUB)) {
            g_last_report = j.rep;
            return false;
        }
        if (!ensure_init()) {
            j.rep.failed |= MINA_B200_STAGE_INTERNAL_ERROR;
            g_last_report = j.rep;
            return false;
        }
        g_state_queue.submit(j);
    } catch (...) {
        j.accept = false;
        j.rep.failed |= MINA_B200_STAGE_INTERNAL_ERROR;
    }
    g_last_report = j.rep;
    return j.accept;
}

bool verify_account_inclusion_ffi(const unsigned char *proof_buffer, size_t proof_len, const unsigned char *public_input_buffer,
                                  size_t public_input_len) {
    AccountJob j;
    j.proof = proof_buffer;
    j.proof_len = proof_len;
    j.pub = public_input_buffer;
    j.pub_len = public_input_len;
    try {
        g_account_queue.submit(j);
    } catch (...) {
        j.accept = false;
        j.rep.failed |= MINA_B200_STAGE_INTERNAL_ERROR;
    }
    g_last_report = j.rep;
    return j.accept;
}

void mina_b200_last_stages(mina_b200_stage_report *out) {
    if (out) *out = g_last_report;
}

int mina_b200_verify_state_stages(size_t n, const unsigned char *const *proofs, const size_t *proof_lens,
                                  const unsigned char *const *pub_inputs, const size_t *pub_input_lens, int mode,
                                  mina_b200_stage_report *reports, uint8_t *accept_out) {
    try {
        std::vector<StateJob> jobs(n);
        for (size_t i = 0; i < n; i++) {
            jobs[i].proof = proofs[i];
            jobs[i].proof_len = proof_lens[i];
            jobs[i].pub = pub_inputs[i];
            jobs[i].pub_len = pub_input_lens[i];
        }
        verify_state_jobs(jobs.data(), n, mode);
        for (size_t i = 0; i < n; i++) {
            if (reports) reports[i] = jobs[i].rep;
            if (accept_out) accept_out[i] = jobs[i].accept ? 1 : 0;
        }
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
    } catch (...) {
        set_error("unknown error");
    }
    for (size_t i = 0; i < n; i++) {
        if (reports) reports[i] = mina_b200_stage_report{0, MINA_B200_STAGE_INTERNAL_ERROR, 0};
        if (accept_out) accept_out[i] = 0;
    }
    return -1;
}

int verify_mina_state_batch_ffi(size_t n, const unsigned char *const *proofs, const size_t *proof_lens,
                                const unsigned char *const *pub_inputs, const size_t *pub_input_lens, uint8_t *accept_out) {
    if (n && !ensure_init()) {
        for (size_t i = 0; i < n; i++) accept_out[i] = 0;
        return -1;
    }
    return mina_b200_verify_state_stages(n, proofs, proof_lens, pub_inputs, pub_input_lens, default_mode(), nullptr, accept_out);
}

int mina_b200_verify_account_stages(size_t n, const unsigned char *const *proofs, const size_t *proof_lens,
                                    const unsigned char *const *pub_inputs, const size_t *pub_input_lens,
                                    mina_b200_stage_report *reports, uint8_t *accept_out) {
    try {
        std::vector<AccountJob> jobs(n);
        for (size_t i = 0; i < n; i++) {
            jobs[i].proof = proofs[i];
            jobs[i].proof_len = proof_lens[i];
            jobs[i].pub = pub_inputs[i];
            jobs[i].pub_len = pub_input_lens[i];
        }
        verify_account_jobs(jobs.data(), n);
        for (size_t i = 0; i < n; i++) {
            if (reports) reports[i] = jobs[i].rep;
            if (accept_out) accept_out[i] = jobs[i].accept ? 1 : 0;
        }
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
    } catch (...) {
        set_error("unknown error");
    }
    for (size_t i = 0; i < n; i++) {
        if (reports) reports[i] = mina_b200_stage_report{0, MINA_B200_STAGE_INTERNAL_ERROR, 0};
        if (accept_out) accept_out[i] = 0;
    }
    return -1;
}

int verify_account_inclusion_batch_ffi(size_t n, const unsigned char *const *proofs, const size_t *proof_lens,
                                       const unsigned char *const *pub_inputs, const size_t *pub_input_lens, uint8_t *accept_out) {
    return mina_b200_verify_account_stages(n, proofs, proof_lens, pub_inputs, pub_input_lens, nullptr, accept_out);
}

// accumulator_check alone, from raw proof bytes: ok3[3*i + {0,1,2}] = wrap (Vesta) accumulator, step
// (Pallas) accumulators 0 and 1.  A proof that does not decode gives 0,0,0.
int mina_b200_accumulator_check_batch(size_t n, const unsigned char *const *proofs, const size_t *proof_lens, int mode, uint8_t *ok3) {
    try {
        require_ready();
        std::memset(ok3, 0, 3 * n);
        AccumulatorBatch wrap, step;
        wrap.curve = 1;
        wrap.k = 16;
        step.curve = 0;
        step.k = 15;
        std::vector<size_t> wrap_owner, step_owner;
        std::vector<std::unique_ptr<wire::PicklesProof>> dec(n);
        parallel_for(n, [&](size_t i) {
            if (!proofs[i] || proof_lens[i] > wire::MAX_STATE_PROOF_SIZE) return;
            auto p = std::make_unique<wire::PicklesProof>();
            wire::Reader r(proofs[i], proof_lens[i]);
            wire::read_pickles_proof(r, *p);
            if (r.ok()) dec[i] = std::move(p);
        });
        for (size_t i = 0; i < n; i++) {
            if (!dec[i]) continue;
            const wire::PicklesProof &p = *dec[i];
            uint8_t buf[16 * 16];
            if (point_on_curve<FqParams>(p.wrap_challenge_polynomial_commitment)) {
                for (int t = 0; t < 16; t++) put_u128(buf + 16 * t, p.bulletproof_challenges[t]);
                wrap.pre.insert(wrap.pre.end(), buf, buf + 256);
                wrap.pts.insert(wrap.pts.end(), p.wrap_challenge_polynomial_commitment.x.begin(), p.wrap_challenge_polynomial_commitment.x.end());
                wrap.pts.insert(wrap.pts.end(), p.wrap_challenge_polynomial_commitment.y.begin(), p.wrap_challenge_polynomial_commitment.y.end());
                wrap_owner.push_back(i);
            }
            for (size_t k = 0; k < 2 && p.step_challenge_polynomial_commitments.size() == 2; k++) {
                const wire::Point &pt = p.step_challenge_polynomial_commitments[k];
                if (!point_on_curve<FpParams>(pt)) continue;
                for (int t = 0; t < 15; t++) put_u128(buf + 16 * t, p.wrap_old_bulletproof_challenges[k][t]);
                step.pre.insert(step.pre.end(), buf, buf + 240);
                step.pts.insert(step.pts.end(), pt.x.begin(), pt.x.end());
                step.pts.insert(step.pts.end(), pt.y.begin(), pt.y.end());
                step_owner.push_back(3 * i + 1 + k);
            }
        }
        wrap.m = (uint32_t)wrap_owner.size();
        step.m = (uint32_t)step_owner.size();
        {
            Context &c = ctx();
            std::lock_guard<std::mutex> lk(c.mu);
            CTX_CUDA_OK(cudaSetDevice(c.device));
            run_both_sides(c, wrap, step, mode, false, nullptr, nullptr);
        }
        for (size_t t = 0; t < wrap_owner.size(); t++) ok3[3 * wrap_owner[t]] = wrap.ok[t];
        for (size_t t = 0; t < step_owner.size(); t++) ok3[step_owner[t]] = step.ok[t];
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
    } catch (...) {
        set_error("unknown error");
    }
    return -1;
}

// The same device pipeline over inputs that already sit in HBM (bench `value` leg): d_pre16 = m*k
// 16-byte prechallenges, d_pts64 = m canonical affine points (validated by the caller), k = 16 for
// Vesta (curve 1) and 15 for Pallas (curve 0).  ok_host receives m bytes.  The call synchronises.
int mina_b200_accumulators_device(int curve, uint32_t m, const void *d_pre16, const void *d_pts64, int mode, uint8_t *ok_host,
                                  mina_b200_kernel_stats *stats) {
    try {
        require_ready();
        if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
        AccumulatorBatch ab;
        ab.curve = curve;
        ab.k = curve == 1 ? 16 : 15;
        ab.m = m;
        ab.d_pre_ext = (const uint8_t *)d_pre16;
        ab.d_pts_ext = (const uint8_t *)d_pts64;
        Context &c = ctx();
        std::lock_guard<std::mutex> lk(c.mu);
        CTX_CUDA_OK(cudaSetDevice(c.device));
        AccRun rs;
        setup_run(c, rs, 0, stats != nullptr);
        try {
            run_accumulators(c, rs, ab, mode);
        } catch (...) {
            for (int i = 0; i < 2; i++)
                if (rs.ev[i]) cudaEventDestroy(rs.ev[i]);
            throw;
        }
        for (int i = 0; i < 2; i++)
            if (rs.ev[i]) cudaEventDestroy(rs.ev[i]);
        if (stats) *stats = rs.stats;
        if (m) std::memcpy(ok_host, ab.ok.data(), m);
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
    } catch (...) {
        set_error("unknown error");
    }
    return -1;
}

// Both families of m state proofs at once (Vesta wrap accumulator + the two Pallas step accumulators per proof),
// driven concurrently on two streams.  ok3[3*i + {0,1,2}] like mina_b200_accumulator_check_batch.
int mina_b200_state_accumulators_device(uint32_t m, const void *d_pre_wrap, const void *d_pts_wrap, const void *d_pre_step,
                                        const void *d_pts_step, int mode, uint8_t *ok3, mina_b200_kernel_stats *stats2) {
    try {
        require_ready();
        AccumulatorBatch wrap, step;
        wrap.curve = 1;
        wrap.k = 16;
        wrap.m = m;
        wrap.d_pre_ext = (const uint8_t *)d_pre_wrap;
        wrap.d_pts_ext = (const uint8_t *)d_pts_wrap;
        step.curve = 0;
        step.k = 15;
        step.m = 2 * m;
        step.d_pre_ext = (const uint8_t *)d_pre_step;
        step.d_pts_ext = (const uint8_t *)d_pts_step;
        Context &c = ctx();
        std::lock_guard<std::mutex> lk(c.mu);
        CTX_CUDA_OK(cudaSetDevice(c.device));
        run_both_sides(c, wrap, step, mode, stats2 != nullptr, stats2, stats2 ? stats2 + 1 : nullptr);
        for (uint32_t i = 0; i < m; i++) {
            ok3[3 * i] = wrap.ok[i];
            ok3[3 * i + 1] = step.ok[2 * i];
            ok3[3 * i + 2] = step.ok[2 * i + 1];
        }
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
    } catch (...) {
        set_error("unknown error");
    }
    return -1;
}

int mina_b200_accumulator_check(const unsigned char *proof, size_t proof_len, uint8_t ok3[3]) {
    return mina_b200_accumulator_check_batch(1, &proof, &proof_len, MINA_B200_MODE_PER_PROOF, ok3);
}

int mina_b200_ipa_verify(int curve, const uint8_t *poseidon_table, const mina_b200_ipa_batch *batch, uint8_t *ok) {
    try {
        require_ready();
        if (curve < 0 || curve > 1 || !batch || !ok) throw std::runtime_error("ipa_verify: bad arguments");
        const mina_b200_ipa_batch &b = *batch;
        if (b.n == 0) return 0;
        Context &c = ctx();
        if (b.rounds < (uint32_t)BPOLY_LO_BITS || b.rounds > 16 || (1u << b.rounds) > c.curve[curve].depth)
            throw std::runtime_error("ipa_verify: rounds must be in [8, 16] and 2^rounds must not exceed the resident SRS");
        if (b.n_points == 0 || b.n_points > 8 || b.n_comm > 4096) throw std::runtime_error("ipa_verify: bad shape");
        if (b.sponge_mode > 1 || b.sponge_count > 2) throw std::runtime_error("ipa_verify: bad sponge mode");
        if (curve == 0) {
            poseidon::Params<FpParams> t;
            if (!t.from_bytes(poseidon_table, (size_t)poseidon::TABLE_WORDS * 32)) throw std::runtime_error("ipa_verify: bad Poseidon table");
            ipa_verify_t<FpParams, FqParams, true>(c, 0, t, b, ok);
        } else {
            poseidon::Params<FqParams> t;
            if (!t.from_bytes(poseidon_table, (size_t)poseidon::TABLE_WORDS * 32)) throw std::runtime_error("ipa_verify: bad Poseidon table");
            ipa_verify_t<FqParams, FpParams, false>(c, 1, t, b, ok);
        }
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
    } catch (...) {
        set_error("unknown error");
    }
    return -1;
}

// Merkle fold over caller-supplied leaves with a caller-supplied Poseidon table (parity hook for K3 and
// merkle_verifier.rs:9-35; the product path will use the trusted resident table).
int mina_b200_merkle_fold(const uint8_t *table, uint32_t nproofs, uint32_t max_depth, const uint32_t *depths, const uint8_t *tags,
                          const uint8_t *siblings32, const uint8_t *leaves32, const uint8_t *roots32, uint8_t *ok, uint8_t *folded32) {
    try {
        require_ready();
        if (!nproofs) return 0;
        poseidon::Params<FpParams> params;
        if (!params.from_bytes(table, (size_t)poseidon::TABLE_WORDS * 32)) throw std::runtime_error("merkle_fold: bad Poseidon table");
        for (uint32_t p = 0; p < nproofs; p++)
            if (depths[p] > max_depth) throw std::runtime_error("merkle_fold: depth exceeds max_depth");
        // per-depth prefix states (host, once per call: max_depth permutations)
        std::vector<uint64_t> prefix((size_t)std::max<uint32_t>(max_depth, 1) * 12);
        for (uint32_t d = 0; d < max_depth; d++) {
            host::Fp st[3];
            if (!poseidon::prefix_state<FpParams>(params, poseidon::merkle_prefix(d), st)) throw std::runtime_error("merkle_fold: bad prefix");
            for (int k = 0; k < 3; k++) std::memcpy(&prefix[(size_t)d * 12 + 4 * k], st[k].l, 32);
        }
        std::vector<MerkleNodeDev> nodes((size_t)nproofs * std::max<uint32_t>(max_depth, 1));
        for (uint32_t p = 0; p < nproofs; p++)
            for (uint32_t d = 0; d < depths[p]; d++) {
                MerkleNodeDev &nd = nodes[(size_t)p * max_depth + d];
                std::memcpy(nd.hash, siblings32 + 32 * ((size_t)p * max_depth + d), 32);
                nd.tag = tags[(size_t)p * max_depth + d];
            }
        auto tab = params.device_table();
        Context &c = ctx();
        std::lock_guard<std::mutex> lk(c.mu);
        CTX_CUDA_OK(cudaSetDevice(c.device));
        VerifierState &vs = vstate();
        fe *d_tab = vs.d_prefix.reserve((size_t)POSEIDON_TABLE_WORDS + prefix.size() / 4);
        fe *d_prefix = d_tab + POSEIDON_TABLE_WORDS;
        MerkleNodeDev *d_nodes = vs.d_nodes.reserve(nodes.size());
        uint32_t *d_depths = vs.d_depths.reserve(nproofs);
        fe *d_leaves = vs.d_leaves.reserve(nproofs), *d_roots = vs.d_roots.reserve(nproofs), *d_folded = vs.d_folded.reserve(nproofs);
        uint8_t *d_ok = vs.d_ok.reserve(nproofs);
        CTX_CUDA_OK(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, c.stream));
        CTX_CUDA_OK(cudaMemcpyAsync(d_prefix, prefix.data(), prefix.size() * 8, cudaMemcpyHostToDevice, c.stream));
        CTX_CUDA_OK(cudaMemcpyAsync(d_nodes, nodes.data(), nodes.size() * sizeof(MerkleNodeDev), cudaMemcpyHostToDevice, c.stream));
        CTX_CUDA_OK(cudaMemcpyAsync(d_depths, depths, (size_t)nproofs * 4, cudaMemcpyHostToDevice, c.stream));
        CTX_CUDA_OK(cudaMemcpyAsync(d_leaves, leaves32, (size_t)nproofs * 32, cudaMemcpyHostToDevice, c.stream));
        CTX_CUDA_OK(cudaMemcpyAsync(d_roots, roots32, (size_t)nproofs * 32, cudaMemcpyHostToDevice, c.stream));
        launch_merkle_fold(0, d_tab, d_prefix, d_nodes, d_depths, max_depth, d_leaves, d_roots, d_ok, d_folded, nproofs, c.stream);
        c.launches += 1;
        CTX_CUDA_OK(cudaMemcpyAsync(ok, d_ok, nproofs, cudaMemcpyDeviceToHost, c.stream));
        if (folded32) CTX_CUDA_OK(cudaMemcpyAsync(folded32, d_folded, (size_t)nproofs * 32, cudaMemcpyDeviceToHost, c.stream));
        CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
    } catch (...) {
        set_error("unknown error");
    }
    return -1;
}

}  // extern "C"
