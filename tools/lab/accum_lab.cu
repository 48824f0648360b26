// Kernel lab (not part of the library): times k_accumulate variants and the raw IMAD issue rates on
// the engine's real intermediate data (64 x 2^16-point MSMs, c = 16, fixed-base table).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo
//             -I mina_bridge_b200/csrc tools/lab/accum_lab.cu mina_bridge_b200/csrc/ipa.cu -o gpurun_out/accum_lab
#include <cstdio>
#include <random>
#include <vector>
#define private public
#include "msm_impl.cuh"
#undef private

using namespace pasta;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// ---- raw issue-rate probes -------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) k_rate(uint32_t *out, uint32_t a, uint32_t b, int iters) {
    uint32_t x0 = threadIdx.x, x1 = a, x2 = b, x3 = a ^ b, x4 = a + 1, x5 = b + 2, x6 = a + 3, x7 = b + 5;
    uint64_t w0 = x0, w1 = x1, w2 = x2, w3 = x3, w4 = x4, w5 = x5, w6 = x6, w7 = x7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (MODE == 0) {  // IMAD.WIDE.U32: 64-bit accumulate
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w4) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w5) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w6) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w7) : "r"(a), "r"(b));
            } else if (MODE == 1) {  // IMAD (lo)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x0) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x1) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x2) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x3) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x4) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x5) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x6) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x7) : "r"(a), "r"(b));
            } else if (MODE == 2) {  // IMAD.HI
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x0) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x1) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x2) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x3) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x4) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x5) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x6) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x7) : "r"(a), "r"(b));
            } else {  // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x0) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x1) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x2) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x3) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x4) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x5) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x6) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x7) : "r"(a));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (uint32_t)(w0 + w1 + w2 + w3 + w4 + w5 + w6 + w7);
}

// modmul throughput: 4 independent chains per thread
template <class F, int V>
__device__ __forceinline__ fe mulv(const fe &a, const fe &b) { return V == 2 ? Fd<F>::mul_ptx2(a, b) : Fd<F>::mul_ptx(a, b); }
template <class F, int V>
__global__ void __launch_bounds__(128) k_modmul_rate(fe *io, int iters) {
    fe a = io[threadIdx.x], b = io[threadIdx.x + 128], c = io[threadIdx.x + 256], d = io[threadIdx.x + 384];
    for (int i = 0; i < iters; i++) {
        a = mulv<F, V>(a, b); b = mulv<F, V>(b, c); c = mulv<F, V>(c, d); d = mulv<F, V>(d, a);
    }
    io[blockIdx.x * 128 + threadIdx.x] = Fd<F>::add(Fd<F>::add(a, b), Fd<F>::add(c, d));
}
// correctness: mul_ptx2 against the portable CIOS product on (pseudo)random and structured inputs
template <class F>
__global__ void k_mul_check(const fe *a, const fe *b, uint32_t n, uint32_t *bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x = a[i], y = b[i];
    for (int it = 0; it < 8; it++) {
        fe r1 = Fd<F>::mul_portable(x, y), r2 = Fd<F>::mul_ptx2(x, y);
        if (!fe_eq(r1, r2)) atomicAdd(bad, 1u);
        x = y;
        y = r1;  // walk: covers values produced by the arithmetic itself
    }
}

// ---- accumulate variants ---------------------------------------------------------------------------------------
template <class F, int V>
__device__ __noinline__ fe mul_call_v(const fe a, const fe b) { return mulv<F, V>(a, b); }
template <class F, int V>
__device__ __forceinline__ void add_mixed_v(xyzz &p, const affine &q) {
    using fd = Fd<F>;
    if (Ec<F>::is_identity(q)) return;
    if (Ec<F>::is_identity(p)) { p = Ec<F>::from_affine(q); return; }
    fe U2 = mul_call_v<F, V>(q.x, p.zz);
    fe S2 = mul_call_v<F, V>(q.y, p.zzz);
    fe P = fd::sub(U2, p.x);
    fe R = fd::sub(S2, p.y);
    if (fe_is_zero(P)) {
        if (fe_is_zero(R)) p = Ec<F>::dbl_affine(q); else p = Ec<F>::identity();
        return;
    }
    fe PP = mul_call_v<F, V>(P, P);
    fe PPP = mul_call_v<F, V>(P, PP);
    fe Q = mul_call_v<F, V>(p.x, PP);
    fe X3 = fd::sub(fd::sub(mul_call_v<F, V>(R, R), PPP), fd::dbl(Q));
    fe Y3 = fd::sub(mul_call_v<F, V>(R, fd::sub(Q, X3)), mul_call_v<F, V>(p.y, PPP));
    p.x = X3; p.y = Y3;
    p.zz = mul_call_v<F, V>(p.zz, PP);
    p.zzz = mul_call_v<F, V>(p.zzz, PPP);
}
template <class F, int V, int MINB>
__global__ void __launch_bounds__(128, MINB) k_acc_v(const uint32_t *__restrict__ order, const uint32_t *__restrict__ offsets,
                                                     const uint32_t *__restrict__ pairs, const affine *__restrict__ table,
                                                     xyzz *__restrict__ buckets, uint32_t nbuckets) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbuckets) return;
    uint32_t b = order[t];
    uint32_t k = offsets[b], end = offsets[b + 1];
    xyzz acc = Ec<F>::identity();
    if (k < end) {
        uint32_t e = pairs[k];
        affine q = load_point<F>(table, e);
        for (;;) {
            uint32_t e_next = 0;
            affine q_next;
            bool more = (k + 1 < end);
            if (more) { e_next = pairs[k + 1]; q_next = load_point<F>(table, e_next); }
            if (e >> 31) q.y = Fd<F>::neg(q.y);
            add_mixed_v<F, V>(acc, q);
            if (!more) break;
            q = q_next; e = e_next; k++;
        }
    }
    buckets[b] = acc;
}

template <class K>
float time_it(K launch, int reps = 3) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, clk);
    // ---- issue rates: 148*4 blocks of 128 threads = 1 warp per SMSP x 4 ... use 8 warps/SMSP to saturate
    {
        uint32_t *out; CK(cudaMalloc(&out, 148 * 32 * 128 * 4));
        const int iters = 2000;
        const char *names[4] = {"IMAD.WIDE.U32", "IMAD.LO", "IMAD.HI", "IADD"};
        for (int mode = 0; mode < 4; mode++) {
            auto launch = [&]() {
                if (mode == 0) k_rate<0><<<148 * 8, 128>>>(out, 12345, 67891, iters);
                if (mode == 1) k_rate<1><<<148 * 8, 128>>>(out, 12345, 67891, iters);
                if (mode == 2) k_rate<2><<<148 * 8, 128>>>(out, 12345, 67891, iters);
                if (mode == 3) k_rate<3><<<148 * 8, 128>>>(out, 12345, 67891, iters);
            };
            float ms = time_it(launch);
            double warp_insts = 148.0 * 8 * 4 * iters * 16 * 8;  // blocks * warps * ...
            double per_smsp = warp_insts / (148.0 * 4);
            printf("rate %-14s: %.3f ms, %.2f Tthread-op/s, %.3f warp-inst/ns/SMSP\n", names[mode], ms, warp_insts * 32 / ms / 1e9,
                   per_smsp / (ms * 1e6));
        }
        cudaFree(out);
    }
    {
        fe *io; CK(cudaMalloc(&io, 148 * 16 * 128 * sizeof(fe) + 512 * sizeof(fe)));
        std::vector<uint32_t> h(512 * 8);
        std::mt19937 rng(1);
        for (auto &x : h) x = rng();
        for (int i = 0; i < 512; i++) h[i * 8 + 7] &= 0x3fffffff;
        CK(cudaMemcpy(io, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        const int iters = 2000;
        for (int nb = 4; nb <= 16; nb *= 2) {
            float ms = time_it([&]() { k_modmul_rate<FqParams, 1><<<148 * nb, 128>>>(io, iters); });
            printf("modmul v1 chains, %2d blocks/SM: %.3f ms -> %.1f G modmul/s\n", nb, ms, 148.0 * nb * 128 * 4 * iters / ms / 1e6);
            ms = time_it([&]() { k_modmul_rate<FqParams, 2><<<148 * nb, 128>>>(io, iters); });
            printf("modmul v2 chains, %2d blocks/SM: %.3f ms -> %.1f G modmul/s\n", nb, ms, 148.0 * nb * 128 * 4 * iters / ms / 1e6);
        }
        cudaFree(io);
        // correctness of v2 on both fields
        const uint32_t n = 1 << 20;
        std::vector<uint32_t> ha((size_t)n * 8), hb((size_t)n * 8);
        std::mt19937 r2(99);
        const uint32_t pat[6] = {0u, 1u, 0xffffffffu, 0x80000000u, 0x7fffffffu, 0xfffffffeu};
        for (uint32_t i = 0; i < n; i++)
            for (int k = 0; k < 8; k++) {
                bool structured = i < (n / 4);
                ha[(size_t)i * 8 + k] = structured ? pat[r2() % 6] : r2();
                hb[(size_t)i * 8 + k] = structured ? pat[r2() % 6] : r2();
                if (k == 7) { ha[(size_t)i * 8 + 7] &= 0x3fffffff; hb[(size_t)i * 8 + 7] &= 0x3fffffff; }
            }
        fe *da, *db; uint32_t *dbad;
        CK(cudaMalloc(&da, (size_t)n * 32)); CK(cudaMalloc(&db, (size_t)n * 32)); CK(cudaMalloc(&dbad, 4));
        CK(cudaMemcpy(da, ha.data(), (size_t)n * 32, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, hb.data(), (size_t)n * 32, cudaMemcpyHostToDevice));
        for (int f = 0; f < 2; f++) {
            CK(cudaMemset(dbad, 0, 4));
            if (f == 0) k_mul_check<FpParams><<<n / 128, 128>>>(da, db, n, dbad); else k_mul_check<FqParams><<<n / 128, 128>>>(da, db, n, dbad);
            uint32_t bad = 0; CK(cudaMemcpy(&bad, dbad, 4, cudaMemcpyDeviceToHost));
            printf("mul_ptx2 vs portable, field %d: %u mismatches of %u\n", f, bad, n * 8);
        }
        cudaFree(da); cudaFree(db); cudaFree(dbad);
    }

    // ---- engine data -------------------------------------------------------------------------------------------
    const uint32_t n = 65536, nmsm = 64;
    std::vector<uint32_t> pts((size_t)n * 16), sc((size_t)nmsm * n * 8);
    std::mt19937 rng(7);
    for (auto &x : pts) x = rng();
    for (size_t i = 0; i < (size_t)n * 2; i++) pts[i * 8 + 7] &= 0x3fffffff;
    for (auto &x : sc) x = rng();
    for (size_t i = 0; i < (size_t)nmsm * n; i++) sc[i * 8 + 7] &= 0x1fffffff;
    affine *d_pts, *d_out; uint32_t *d_sc;
    CK(cudaMalloc(&d_pts, (size_t)n * 64)); CK(cudaMalloc(&d_sc, sc.size() * 4)); CK(cudaMalloc(&d_out, nmsm * 64));
    CK(cudaMemcpy(d_pts, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sc, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice));
    MsmEngine<FqParams> eng;
    MsmConfig cfg;
    eng.set_bases(d_pts, n, cfg, 0);
    eng.enable_kernel_timing(true);
    eng.run(d_sc, nmsm, n, d_out, 0);
    CK(cudaDeviceSynchronize());
    printf("engine k_accumulate (in pipeline): %.3f ms\n", eng.last_accumulate_ms());
    const uint32_t nb = nmsm * eng.nbw_;
    dim3 g((nb + 127) / 128);
#define RUN(label, ...) { float ms = time_it([&]() { __VA_ARGS__; }); printf("%-46s %.3f ms\n", label, ms); }
    RUN("library k_accumulate (sorted, call, mul v1)", (k_accumulate<FqParams><<<g, 128>>>(eng.order_, eng.offsets_, eng.pairs_, eng.table_, eng.buckets_, nb)));
    RUN("k_acc_v mul v1 minb=4", (k_acc_v<FqParams, 1, 4><<<g, 128>>>(eng.order_, eng.offsets_, eng.pairs_, eng.table_, eng.buckets_, nb)));
    RUN("k_acc_v mul v2 minb=3", (k_acc_v<FqParams, 2, 3><<<g, 128>>>(eng.order_, eng.offsets_, eng.pairs_, eng.table_, eng.buckets_, nb)));
    RUN("k_acc_v mul v2 minb=4", (k_acc_v<FqParams, 2, 4><<<g, 128>>>(eng.order_, eng.offsets_, eng.pairs_, eng.table_, eng.buckets_, nb)));
    RUN("k_acc_v mul v2 minb=5", (k_acc_v<FqParams, 2, 5><<<g, 128>>>(eng.order_, eng.offsets_, eng.pairs_, eng.table_, eng.buckets_, nb)));
    // whole MSM batch through the engine (all kernels), for the share of the non-accumulate stages
    RUN("engine.run 64 x 2^16 (all kernels)", (eng.run(d_sc, nmsm, n, d_out, 0)));
    RUN("engine.run  1 x 2^16 (all kernels)", (eng.run(d_sc, 1, n, d_out, 0)));
    return 0;
}
