// C ABI: lifecycle, MSM entry points and device self-test hooks (include/mina_b200.h).
#include <cstdio>
#include <cstring>

#include "../../include/mina_b200.h"
#include "context.cuh"

namespace pasta {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

Context &ctx() {
    static Context c;
    return c;
}
void require_ready() {
    if (!ctx().ready) throw std::runtime_error("mina_b200: not initialised (call mina_b200_init; a CUDA device is required)");
}

template <class B>
static void upload_srs(CurveCtx &cc, const host::Srs<B> &srs, uint32_t depth, cudaStream_t s) {
    cc.depth = depth;
    std::vector<uint64_t> flat((size_t)(depth + 1) * 8);
    cc.host_canonical.resize((size_t)(depth + 1) * 64);
    for (uint32_t i = 0; i <= depth; i++) {
        const host::Affine<B> &p = i < depth ? srs.g[i] : srs.h;
        std::memcpy(&flat[(size_t)i * 8], p.x.l, 32);
        std::memcpy(&flat[(size_t)i * 8 + 4], p.y.l, 32);
        p.x.to_bytes_le(&cc.host_canonical[(size_t)i * 64]);
        p.y.to_bytes_le(&cc.host_canonical[(size_t)i * 64 + 32]);
    }
    CTX_CUDA_OK(cudaMalloc(&cc.d_srs, flat.size() * 8));
    CTX_CUDA_OK(cudaMemcpyAsync(cc.d_srs, flat.data(), flat.size() * 8, cudaMemcpyHostToDevice, s));
    CTX_CUDA_OK(cudaStreamSynchronize(s));
}

template <class B>
static host::Srs<B> load_or_create_srs(const char *cache_dir, const char *name, uint32_t depth) {
    host::Srs<B> srs;
    std::string path;
    if (cache_dir && *cache_dir) {
        path = std::string(cache_dir) + "/" + name + "_" + std::to_string(depth) + ".srsbin";
        if (host::srs_load_cache<B>(path, depth, srs)) return srs;
    }
    srs = host::srs_create<B>(depth);
    if (!path.empty()) host::srs_store_cache<B>(path, srs);  // best effort
    return srs;
}

}  // namespace pasta

using namespace pasta;

#define ABI_TRY try {
#define ABI_CATCH                          \
    }                                      \
    catch (const std::exception &e) {      \
        set_error(e.what());               \
        return -1;                         \
    }                                      \
    catch (...) {                          \
        set_error("unknown error");        \
        return -1;                         \
    }

extern "C" {

const char *mina_b200_last_error(void) { return g_last_error.c_str(); }
uint64_t mina_b200_launch_count(void) { return ctx().launches.load(); }

int mina_b200_msm_configure(int curve, int window_bits, int precompute, int leaf) {
    ABI_TRY
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    MsmConfig cfg;
    cfg.c = window_bits;
    cfg.precompute = precompute != 0;
    cfg.leaf = leaf;
    c.curve[curve].cfg = cfg;
    if (c.ready) {
        c.curve[curve].fixed->set_bases(c.curve[curve].d_srs, c.curve[curve].depth, cfg, c.stream);
        CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    }
    return 0;
    ABI_CATCH
}

int mina_b200_init(int device, const char *cache_dir) {
    ABI_TRY
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.ready) return 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error("mina_b200_init: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) throw std::runtime_error("mina_b200_init: bad device index");
    CTX_CUDA_OK(cudaSetDevice(device));
    c.device = device;
    CTX_CUDA_OK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    // Both files in the reference hold 65536 points; the Pallas side only ever uses the first 2^15.
    c.srs_vesta = load_or_create_srs<FqParams>(cache_dir, "vesta", VESTA_SRS_DEPTH);
    c.srs_pallas = load_or_create_srs<FpParams>(cache_dir, "pallas", PALLAS_SRS_DEPTH);
    upload_srs<FpParams>(c.curve[0], c.srs_pallas, PALLAS_SRS_DEPTH, c.stream);
    upload_srs<FqParams>(c.curve[1], c.srs_vesta, VESTA_SRS_DEPTH, c.stream);
    for (int k = 0; k < 2; k++) {
        c.curve[k].fixed.reset(make_msm_engine(k));
        c.curve[k].var.reset(make_msm_engine(k));
        c.curve[k].fixed->set_bases(c.curve[k].d_srs, c.curve[k].depth, c.curve[k].cfg, c.stream);
    }
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    c.ready = true;
    return 0;
    ABI_CATCH
}

void mina_b200_shutdown(void) {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.ready) return;
    cudaSetDevice(c.device);
    for (int k = 0; k < 2; k++) {
        c.curve[k].fixed.reset();
        c.curve[k].var.reset();
        if (c.curve[k].d_srs) cudaFree(c.curve[k].d_srs);
        c.curve[k].d_srs = nullptr;
    }
    cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.ready = false;
}

int mina_b200_srs_points(int curve, uint32_t first, uint32_t count, uint8_t *out64, uint8_t *h64) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    CurveCtx &cc = ctx().curve[curve];
    if ((uint64_t)first + count > cc.depth) throw std::runtime_error("srs range out of bounds");
    if (count) std::memcpy(out64, &cc.host_canonical[(size_t)first * 64], (size_t)count * 64);
    if (h64) std::memcpy(h64, &cc.host_canonical[(size_t)cc.depth * 64], 64);
    return 0;
    ABI_CATCH
}

int mina_b200_msm_srs_device(int curve, uint32_t nmsm, uint32_t n, const void *d_scalars, void *d_out64,
                             void *cuda_stream, float *accumulate_ms) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    CurveCtx &cc = c.curve[curve];
    cudaStream_t s = (cudaStream_t)cuda_stream;
    DevBuf<affine> out(nmsm);
    cc.fixed->enable_kernel_timing(accumulate_ms != nullptr);
    cc.fixed->run((const uint32_t *)d_scalars, nmsm, n, out.p, s);
    launch_affine_from_mont(curve, out.p, (uint32_t *)d_out64, nmsm, s);
    c.launches += (uint64_t)cc.fixed->launches_per_run() + 1;
    if (accumulate_ms) {
        CTX_CUDA_OK(cudaStreamSynchronize(s));
        *accumulate_ms = cc.fixed->last_accumulate_ms();
    } else {
        CTX_CUDA_OK(cudaStreamSynchronize(s));  // `out` is freed on return
    }
    return 0;
    ABI_CATCH
}

int mina_b200_msm_srs(int curve, uint32_t nmsm, uint32_t n, const uint8_t *scalars32, uint8_t *out64) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    CurveCtx &cc = c.curve[curve];
    if (n > cc.depth) throw std::runtime_error("msm_srs: n exceeds the resident SRS depth");
    if (nmsm == 0) return 0;
    size_t nsc = (size_t)nmsm * n;
    DevBuf<uint32_t> d_sc(std::max<size_t>(nsc * 8, 8));
    DevBuf<affine> d_out(nmsm);
    DevBuf<uint32_t> d_can((size_t)nmsm * 16);
    if (nsc) CTX_CUDA_OK(cudaMemcpyAsync(d_sc.p, scalars32, nsc * 32, cudaMemcpyHostToDevice, c.stream));
    cc.fixed->enable_kernel_timing(false);
    cc.fixed->run(d_sc.p, nmsm, n, d_out.p, c.stream);
    launch_affine_from_mont(curve, d_out.p, d_can.p, nmsm, c.stream);
    CTX_CUDA_OK(cudaMemcpyAsync(out64, d_can.p, (size_t)nmsm * 64, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    c.launches += (uint64_t)cc.fixed->launches_per_run() + 1;
    return 0;
    ABI_CATCH
}

int mina_b200_msm(int curve, uint32_t n, const uint8_t *scalars32, const uint8_t *points64, int window_bits,
                  uint8_t *out64) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    CurveCtx &cc = c.curve[curve];
    if (n == 0) {
        std::memset(out64, 0, 64);
        return 0;
    }
    DevBuf<uint32_t> d_sc((size_t)n * 8), d_pts_can((size_t)n * 16), d_can(16);
    DevBuf<affine> d_pts(n), d_out(1);
    CTX_CUDA_OK(cudaMemcpyAsync(d_sc.p, scalars32, (size_t)n * 32, cudaMemcpyHostToDevice, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(d_pts_can.p, points64, (size_t)n * 64, cudaMemcpyHostToDevice, c.stream));
    launch_affine_to_mont(curve, d_pts_can.p, d_pts.p, n, c.stream);
    MsmConfig cfg;
    cfg.precompute = false;
    if (window_bits > 0)
        cfg.c = window_bits;
    else {
        // balance n*W mixed adds against W * 2^c reduction adds
        int lg = 0;
        while ((1u << lg) < n) lg++;
        cfg.c = lg <= 6 ? 4 : (lg - 2 > 16 ? 16 : lg - 2);
    }
    cc.var->set_bases(d_pts.p, n, cfg, c.stream);
    cc.var->run(d_sc.p, 1, n, d_out.p, c.stream);
    launch_affine_from_mont(curve, d_out.p, d_can.p, 1, c.stream);
    CTX_CUDA_OK(cudaMemcpyAsync(out64, d_can.p, 64, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    c.launches += (uint64_t)cc.var->launches_per_run() + 2;
    return 0;
    ABI_CATCH
}

}  // extern "C"

// ---- self-test kernels ---------------------------------------------------------------------------
namespace pasta {
template <class F>
__global__ void k_field_op(int op, uint32_t n, const fe *a, const fe *b, fe *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op >= 10) {  // raw (no Montgomery conversion) variants for debugging the PTX path
        fe r0;
        switch (op) {
            case 10: r0 = Fd<F>::mul(a[i], b[i]); break;
            case 11: r0 = Fd<F>::mul_portable(a[i], b[i]); break;
            case 12: r0 = Fd<F>::add(a[i], b[i]); break;
            case 13: r0 = Fd<F>::sub(a[i], b[i]); break;
            case 14: r0 = Fd<F>::add_portable(a[i], b[i]); break;
            case 15: r0 = Fd<F>::sub_portable(a[i], b[i]); break;
#ifdef __CUDA_ARCH__
            case 20: case 21: {
                uint32_t U[16];
                Fd<F>::mul_wide(U, a[i], b[i]);
                for (int k = 0; k < 8; k++) r0.v[k] = U[(op == 20 ? 0 : 8) + k];
                break;
            }
            case 22: {
                uint32_t U[16];
                for (int k = 0; k < 8; k++) { U[k] = a[i].v[k]; U[8 + k] = b[i].v[k]; }
                r0 = Fd<F>::redc(U);
                break;
            }
#endif
            default: r0 = fe_zero(); break;
        }
        out[i] = r0;
        return;
    }
    fe x = Fd<F>::to_mont(a[i]);
    fe y = Fd<F>::to_mont(b[i]);
    fe r;
    switch (op) {
        case 0: r = Fd<F>::mul(x, y); break;
        case 1: r = Fd<F>::add(x, y); break;
        case 2: r = Fd<F>::sub(x, y); break;
        case 3: r = Fd<F>::inv(x); break;
        default: r = Fd<F>::sqr(x); break;
    }
    out[i] = Fd<F>::from_mont(r);
}
template <class F>
__global__ void k_point_add(uint32_t n, const affine *a, const affine *b, affine *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz p = Ec<F>::from_affine(a[i]);
    // exercise both addition formulas: mixed for even i, full for odd i
    if (i & 1) {
        xyzz q = Ec<F>::from_affine(b[i]);
        q = Ec<F>::dbl(q);            // 2b in non-trivial XYZZ form
        xyzz nb = Ec<F>::from_affine(Ec<F>::neg(b[i]));
        Ec<F>::add(q, nb);            // 2b - b = b, ZZ != 1
        Ec<F>::add(p, q);
    } else {
        Ec<F>::add_mixed(p, b[i]);
    }
    out[i] = Ec<F>::to_affine(p);
}
}  // namespace pasta

extern "C" {

int mina_b200_field_op(int field, int op, uint32_t n, const uint8_t *a32, const uint8_t *b32, uint8_t *out32) {
    ABI_TRY
    require_ready();
    if (field < 0 || field > 1 || op < 0 || op > 22) throw std::runtime_error("bad field/op");
    if (n == 0) return 0;
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    DevBuf<fe> a(n), b(n), o(n);
    CTX_CUDA_OK(cudaMemcpyAsync(a.p, a32, (size_t)n * 32, cudaMemcpyHostToDevice, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(b.p, b32 ? b32 : a32, (size_t)n * 32, cudaMemcpyHostToDevice, c.stream));
    if (field == 0)
        k_field_op<FpParams><<<(n + 127) / 128, 128, 0, c.stream>>>(op, n, a.p, b.p, o.p);
    else
        k_field_op<FqParams><<<(n + 127) / 128, 128, 0, c.stream>>>(op, n, a.p, b.p, o.p);
    CTX_CUDA_OK(cudaGetLastError());
    CTX_CUDA_OK(cudaMemcpyAsync(out32, o.p, (size_t)n * 32, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    c.launches += 1;
    return 0;
    ABI_CATCH
}

int mina_b200_point_add(int curve, uint32_t n, const uint8_t *a64, const uint8_t *b64, uint8_t *out64) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    if (n == 0) return 0;
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    DevBuf<uint32_t> ca((size_t)n * 16), cb((size_t)n * 16), co((size_t)n * 16);
    DevBuf<affine> a(n), b(n), o(n);
    CTX_CUDA_OK(cudaMemcpyAsync(ca.p, a64, (size_t)n * 64, cudaMemcpyHostToDevice, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(cb.p, b64, (size_t)n * 64, cudaMemcpyHostToDevice, c.stream));
    launch_affine_to_mont(curve, ca.p, a.p, n, c.stream);
    launch_affine_to_mont(curve, cb.p, b.p, n, c.stream);
    if (curve == 0)
        k_point_add<FpParams><<<(n + 63) / 64, 64, 0, c.stream>>>(n, a.p, b.p, o.p);
    else
        k_point_add<FqParams><<<(n + 63) / 64, 64, 0, c.stream>>>(n, a.p, b.p, o.p);
    CTX_CUDA_OK(cudaGetLastError());
    launch_affine_from_mont(curve, o.p, co.p, n, c.stream);
    CTX_CUDA_OK(cudaMemcpyAsync(out64, co.p, (size_t)n * 64, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    c.launches += 4;
    return 0;
    ABI_CATCH
}

}  // extern "C"
