// C ABI: lifecycle, MSM entry points, IPA / Poseidon kernel hooks and device self-test hooks
// (include/mina_b200.h).  The verifier entry points live in verifier.cu.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mina_b200.h"
#include "context.cuh"
#include "ipa.cuh"
#include "poseidon.cuh"

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
uint64_t engine_launches() {
    uint64_t n = 0;
    Context &c = ctx();
    for (int k = 0; k < 2; k++) {
        if (c.curve[k].fixed) n += c.curve[k].fixed->launches();
        if (c.curve[k].user) n += c.curve[k].user->launches();
        if (c.curve[k].var) n += c.curve[k].var->launches();
    }
    return n;
}

// Persistent scratch of the low-level entry points (grow-only; guarded by Context::mu).
struct AbiScratch {
    DevBuf<uint32_t> scalars[2];  // double-buffered H2D staging of scalar chunks
    DevBuf<affine> out;
    DevBuf<uint32_t> out_can, pts_can;
    DevBuf<affine> pts;
    DevBuf<fe> a, b, c, d;
    DevBuf<uint8_t> bytes;
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    ~AbiScratch() {
        for (int i = 0; i < 2; i++) {
            if (copied[i]) cudaEventDestroy(copied[i]);
            if (consumed[i]) cudaEventDestroy(consumed[i]);
        }
    }
};
static AbiScratch *g_scratch = nullptr;
static AbiScratch &scratch() {
    if (!g_scratch) {
        g_scratch = new AbiScratch();
        for (int i = 0; i < 2; i++) {
            CTX_CUDA_OK(cudaEventCreateWithFlags(&g_scratch->copied[i], cudaEventDisableTiming));
            CTX_CUDA_OK(cudaEventCreateWithFlags(&g_scratch->consumed[i], cudaEventDisableTiming));
        }
    }
    return *g_scratch;
}

template <class B>
static void upload_srs(CurveCtx &cc, const host::Srs<B> &srs, uint32_t depth, cudaStream_t s) {
    cc.depth = depth;
    std::vector<uint64_t> flat = host::srs_payload(srs);
    cc.host_canonical.resize((size_t)(depth + 1) * 64);
    for (uint32_t i = 0; i <= depth; i++) {
        const host::Affine<B> &p = i < depth ? srs.g[i] : srs.h;
        p.x.to_bytes_le(&cc.host_canonical[(size_t)i * 64]);
        p.y.to_bytes_le(&cc.host_canonical[(size_t)i * 64 + 32]);
    }
    CTX_CUDA_OK(cudaMalloc(&cc.d_srs, flat.size() * 8));
    CTX_CUDA_OK(cudaMemcpyAsync(cc.d_srs, flat.data(), flat.size() * 8, cudaMemcpyHostToDevice, s));
    CTX_CUDA_OK(cudaStreamSynchronize(s));
}

// Load the SRS from the cache when its payload matches the compiled-in pin; otherwise derive it by
// hash-to-curve exactly as the reference does at first use (lib.rs:34, verifier_index.rs:204-208),
// check the derivation against the same pin, and store it for the next process.
template <class B>
static host::Srs<B> load_or_create_srs(const std::string &dir, const char *name, uint32_t depth) {
    host::Srs<B> srs;
    std::string path;
    if (!dir.empty()) {
        path = dir + "/" + name + "_" + std::to_string(depth) + ".srsbin";
        if (host::srs_load_cache<B>(path, depth, srs)) return srs;
    }
    // the committed file the bridge ships (srs/<name>.srs), if someone dropped it into the data directory
    std::string err;
    if (!dir.empty() && host::srs_load_file<B>(dir + "/" + name + ".srs", depth, srs, err) && host::srs_matches_pin<B>(srs)) {
        host::srs_store_cache<B>(path, srs);
        return srs;
    }
    srs = host::srs_create<B>(depth);
    if (!host::srs_matches_pin<B>(srs)) throw std::runtime_error(std::string("SRS derivation does not match the pinned digest: ") + name);
    if (!path.empty()) host::srs_store_cache<B>(path, srs);  // best effort
    return srs;
}

static std::string default_data_dir() {
    if (const char *e = std::getenv("MINA_B200_DATA_DIR")) return e;
    Dl_info info;
    if (dladdr((void *)&default_data_dir, &info) && info.dli_fname) {
        std::string so = info.dli_fname;
        size_t slash = so.rfind('/');
        std::string dir = slash == std::string::npos ? "." : so.substr(0, slash);
        return dir + "/../data";  // mina_bridge_b200/lib/libmina_b200.so -> mina_bridge_b200/data
    }
    return "";
}

static void load_keys_and_tables(Context &c) {
    // verification keys: failure is remembered, not fatal -- only state verification needs them
    try {
        c.vk[0] = vk::load_verifier_index(c.data_dir + "/mainnet_vk.json");
        c.vk[1] = vk::load_verifier_index(c.data_dir + "/devnet_vk.json");
        c.vk_loaded = true;
    } catch (const std::exception &e) {
        c.vk_loaded = false;
        c.vk_error = e.what();
    }
    // Poseidon tables: optional data; trusted only when the reference's known-answer test passes
    bool fp = c.poseidon_fp.from_file(c.data_dir + "/poseidon_fp_kimchi.bin");
    bool fq = c.poseidon_fq.from_file(c.data_dir + "/poseidon_fq_kimchi.bin");
    c.poseidon_trusted = fp && poseidon::passes_reference_kat(c.poseidon_fp);
    if (!c.poseidon_trusted) c.poseidon_fp.loaded = false;
    if (!fq || !c.poseidon_trusted) c.poseidon_fq.loaded = false;
}

static void destroy_device_state(Context &c) {
    verifier_release(c);
    delete g_scratch;
    g_scratch = nullptr;
    for (int k = 0; k < 2; k++) {
        c.curve[k].fixed.reset();
        c.curve[k].var.reset();
        c.curve[k].lagr.reset();
        if (c.curve[k].d_lagr) cudaFree(c.curve[k].d_lagr);
        c.curve[k].d_lagr = nullptr;
        c.curve[k].lagr_n = c.curve[k].lagr_log_n = 0;
        c.curve[k].user.reset();
        if (c.curve[k].d_user) cudaFree(c.curve[k].d_user);
        c.curve[k].d_user = nullptr;
        c.curve[k].user_n = 0;
        if (c.curve[k].d_srs) cudaFree(c.curve[k].d_srs);
        c.curve[k].d_srs = nullptr;
        if (c.d_poseidon_tab[k]) cudaFree(c.d_poseidon_tab[k]);
        c.d_poseidon_tab[k] = nullptr;
    }
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    for (cudaStream_t &a : c.aux_stream) {
        if (a) cudaStreamDestroy(a);
        a = nullptr;
    }
    c.stream = c.copy_stream = nullptr;
    c.ready = false;
    c.device = -1;
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

template <class S>
static void lagrange_consts(uint32_t log_n, fe &omega_inv, fe &n_inv) {
    using E = host::Fe<S>;
    E w{{S::ROOT_OF_UNITY_64(0), S::ROOT_OF_UNITY_64(1), S::ROOT_OF_UNITY_64(2), S::ROOT_OF_UNITY_64(3)}};  // order 2^32
    for (uint32_t i = log_n; i < 32; i++) w = w.sqr();  // order 2^log_n: the domain generator (K-G for log_n = 14 over Fq)
    E wi = w.inv(), ni = E::from_u64(1ull << log_n).inv();
    std::memcpy(&omega_inv, wi.l, 32);
    std::memcpy(&n_inv, ni.l, 32);
}

extern "C" {

const char *mina_b200_last_error(void) { return g_last_error.c_str(); }
uint64_t mina_b200_launch_count(void) { return ctx().launches.load() + engine_launches(); }

int mina_b200_msm_configure(int curve, int window_bits, int precompute, int leaf) {
    ABI_TRY
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    MsmConfig cfg;
    cfg.c = window_bits;
    cfg.precompute = precompute != 0;
    cfg.leaf = leaf;
    if (c.ready) {
        CTX_CUDA_OK(cudaSetDevice(c.device));
        c.curve[curve].fixed->set_bases(c.curve[curve].d_srs, c.curve[curve].depth, cfg, c.stream);  // throws before mutating
        CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    }
    c.curve[curve].cfg = cfg;
    return 0;
    ABI_CATCH
}

int mina_b200_init(int device, const char *data_dir) {
    ABI_TRY
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.ready) {
        if (device != c.device) throw std::runtime_error("mina_b200_init: already initialised on device " + std::to_string(c.device));
        return 0;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error("mina_b200_init: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) throw std::runtime_error("mina_b200_init: bad device index");
    try {
        CTX_CUDA_OK(cudaSetDevice(device));
        c.device = device;
        c.data_dir = (data_dir && *data_dir) ? std::string(data_dir) : default_data_dir();
        CTX_CUDA_OK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        CTX_CUDA_OK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        for (cudaStream_t &a : c.aux_stream) CTX_CUDA_OK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        // Both files in the reference hold 65536 points; the Pallas side only ever uses the first 2^15.
        c.srs_vesta = load_or_create_srs<FqParams>(c.data_dir, "vesta", VESTA_SRS_DEPTH);
        c.srs_pallas = load_or_create_srs<FpParams>(c.data_dir, "pallas", PALLAS_SRS_DEPTH);
        upload_srs<FpParams>(c.curve[0], c.srs_pallas, PALLAS_SRS_DEPTH, c.stream);
        upload_srs<FqParams>(c.curve[1], c.srs_vesta, VESTA_SRS_DEPTH, c.stream);
        for (int k = 0; k < 2; k++) {
            c.curve[k].fixed.reset(make_msm_engine(k));
            c.curve[k].var.reset(make_msm_engine(k));
            c.curve[k].fixed->set_bases(c.curve[k].d_srs, c.curve[k].depth, c.curve[k].cfg, c.stream);
        }
        load_keys_and_tables(c);
        if (c.poseidon_fp.loaded) {
            auto t = c.poseidon_fp.device_table();
            CTX_CUDA_OK(cudaMalloc(&c.d_poseidon_tab[0], t.size() * 8));
            CTX_CUDA_OK(cudaMemcpy(c.d_poseidon_tab[0], t.data(), t.size() * 8, cudaMemcpyHostToDevice));
        }
        if (c.poseidon_fq.loaded) {
            auto t = c.poseidon_fq.device_table();
            CTX_CUDA_OK(cudaMalloc(&c.d_poseidon_tab[1], t.size() * 8));
            CTX_CUDA_OK(cudaMemcpy(c.d_poseidon_tab[1], t.data(), t.size() * 8, cudaMemcpyHostToDevice));
        }
        CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    } catch (...) {
        destroy_device_state(c);
        throw;
    }
    c.ready = true;
    return 0;
    ABI_CATCH
}

void mina_b200_shutdown(void) {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.ready) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    destroy_device_state(c);
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

static void throw_on_engine_error(MsmEngineBase &e, cudaStream_t s) {
    if (e.take_error(s) & 1u) throw std::runtime_error("msm: a scalar is >= 2^255 (not a canonical field element); result discarded");
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
    AbiScratch &sc = scratch();
    affine *out = sc.out.reserve(std::max<uint32_t>(nmsm, 1));
    cc.fixed->enable_kernel_timing(accumulate_ms != nullptr);
    cc.fixed->run((const uint32_t *)d_scalars, nmsm, n, out, s);
    launch_affine_from_mont(curve, out, (uint32_t *)d_out64, nmsm, s);
    c.launches += 1;
    if (accumulate_ms) {  // the only case that synchronises
        CTX_CUDA_OK(cudaStreamSynchronize(s));
        *accumulate_ms = cc.fixed->last_accumulate_ms();
    }
    return 0;
    ABI_CATCH
}

// device side of the Lagrange commitments: `count` Montgomery affine points into d_out (c.mu held)
static void lagrange_commitments_device(Context &c, int curve, uint32_t log_n, uint32_t first, uint32_t count, affine *d_out) {
    CurveCtx &cc = c.curve[curve];
    if (log_n == 0 || log_n > 30 || (1u << log_n) > cc.depth) throw std::runtime_error("lagrange: domain larger than the resident SRS");
    if ((uint64_t)first + count > (1ull << log_n)) throw std::runtime_error("lagrange: index out of the domain");
    const int sfield = curve == 1 ? 0 : 1;
    fe omega_inv, n_inv;
    if (sfield == 0)
        lagrange_consts<FpParams>(log_n, omega_inv, n_inv);
    else
        lagrange_consts<FqParams>(log_n, omega_inv, n_inv);
    AbiScratch &sc = scratch();
    const uint32_t n = 1u << log_n;
    // rows of scalars in chunks of at most 64 commitments (64 x 2^14 x 32 B = 32 MiB)
    const uint32_t chunk = std::max<uint32_t>(1, std::min<uint32_t>(count, (1u << 20) / n ? (1u << 20) / n : 1));
    fe *d_sc = reinterpret_cast<fe *>(sc.scalars[0].reserve((size_t)chunk * n * 8));
    cc.fixed->enable_kernel_timing(false);
    for (uint32_t done = 0; done < count; done += chunk) {
        const uint32_t cur = std::min(chunk, count - done);
        launch_lagrange_scalars(sfield, omega_inv, n_inv, (int)log_n, first + done, cur, d_sc, c.stream);
        cc.fixed->run(reinterpret_cast<const uint32_t *>(d_sc), cur, n, d_out + done, c.stream);
        c.launches += 1;
        CTX_CUDA_OK(cudaStreamSynchronize(c.stream));  // d_sc is reused by the next chunk
    }
    if (cc.fixed->take_error(c.stream)) throw std::runtime_error("lagrange: scalar overflow flagged by the MSM engine");
}

int mina_b200_lagrange_commitments(int curve, uint32_t log_n, uint32_t first, uint32_t count, uint8_t *out64) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    if (!count) return 0;
    AbiScratch &sc = scratch();
    affine *out = sc.out.reserve(count);
    uint32_t *can = sc.out_can.reserve((size_t)count * 16);
    lagrange_commitments_device(c, curve, log_n, first, count, out);
    launch_affine_from_mont(curve, out, can, count, c.stream);
    c.launches += 1;
    CTX_CUDA_OK(cudaMemcpyAsync(out64, can, 64 * (size_t)count, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    return 0;
    ABI_CATCH
}

// kimchi's public-input commitment (verifier.rs `public_comm`): -sum_i pub_i L_i + h for nproofs vectors
// of n_pub public inputs.  The n_pub Lagrange commitments and h form a small fixed base set (window table, c = 8).
int mina_b200_public_commitments(int curve, uint32_t log_n, uint32_t n_pub, uint32_t nproofs, const uint8_t *pub32, uint8_t *out64) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    if (n_pub == 0 || n_pub > 4096) throw std::runtime_error("public_commitments: bad number of public inputs");
    if (!nproofs) return 0;
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    CurveCtx &cc = c.curve[curve];
    if (!cc.lagr || cc.lagr_n != n_pub || cc.lagr_log_n != log_n) {
        cc.lagr.reset();
        if (cc.d_lagr) cudaFree(cc.d_lagr);
        cc.d_lagr = nullptr;
        cc.lagr_n = cc.lagr_log_n = 0;
        CTX_CUDA_OK(cudaMalloc(&cc.d_lagr, (size_t)(n_pub + 1) * sizeof(affine)));
        lagrange_commitments_device(c, curve, log_n, 0, n_pub, cc.d_lagr);
        CTX_CUDA_OK(cudaMemcpyAsync(cc.d_lagr + n_pub, cc.d_srs + cc.depth, sizeof(affine), cudaMemcpyDeviceToDevice, c.stream));  // h
        MsmConfig cfg;
        cfg.precompute = true;
        cfg.c = 8;
        cc.lagr.reset(make_msm_engine(curve));
        cc.lagr->set_bases(cc.d_lagr, n_pub + 1, cfg, c.stream);
        CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
        cc.lagr_n = n_pub;
        cc.lagr_log_n = log_n;
    }
    // scalars: -pub_i (host, canonical) then 1 for h
    const int sfield = curve == 1 ? 0 : 1;
    std::vector<uint8_t> sc_host((size_t)nproofs * (n_pub + 1) * 32);
    auto negate = [&](auto tag) {
        using E = host::Fe<decltype(tag)>;
        for (uint32_t p = 0; p < nproofs; p++) {
            for (uint32_t i = 0; i < n_pub; i++) {
                E x;
                if (!E::from_bytes_le(pub32 + 32 * ((size_t)p * n_pub + i), x)) throw std::runtime_error("public_commitments: a public input is not canonical");
                (-x).to_bytes_le(&sc_host[32 * ((size_t)p * (n_pub + 1) + i)]);
            }
            E::one().to_bytes_le(&sc_host[32 * ((size_t)p * (n_pub + 1) + n_pub)]);
        }
    };
    if (sfield == 0)
        negate(FpParams{});
    else
        negate(FqParams{});
    AbiScratch &sc = scratch();
    uint32_t *d_sc = sc.scalars[0].reserve((size_t)nproofs * (n_pub + 1) * 8);
    affine *out = sc.out.reserve(nproofs);
    uint32_t *can = sc.out_can.reserve((size_t)nproofs * 16);
    CTX_CUDA_OK(cudaMemcpyAsync(d_sc, sc_host.data(), sc_host.size(), cudaMemcpyHostToDevice, c.stream));
    cc.lagr->enable_kernel_timing(false);
    cc.lagr->run(d_sc, nproofs, n_pub + 1, out, c.stream);
    launch_affine_from_mont(curve, out, can, nproofs, c.stream);
    c.launches += 1;
    CTX_CUDA_OK(cudaMemcpyAsync(out64, can, 64 * (size_t)nproofs, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    if (cc.lagr->take_error(c.stream)) throw std::runtime_error("public_commitments: scalar overflow flagged by the MSM engine");
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
    AbiScratch &sc = scratch();
    affine *out = sc.out.reserve(nmsm);
    uint32_t *can = sc.out_can.reserve((size_t)nmsm * 16);
    cc.fixed->enable_kernel_timing(false);
    // Chunked, double-buffered: the H2D copy of chunk k+1 (copy stream) overlaps the MSMs of chunk k.
    const size_t per = (size_t)n * 8;
    uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(nmsm, (16u << 20) / std::max<size_t>(per * 4, 1)));
    for (int b = 0; b < 2; b++) sc.scalars[b].reserve(std::max<size_t>((size_t)chunk * per, 8));
    uint32_t idx = 0;
    for (uint32_t done = 0; done < nmsm; done += chunk, idx++) {
        uint32_t cur = std::min(chunk, nmsm - done);
        int b = idx & 1;
        if (idx >= 2) CTX_CUDA_OK(cudaStreamWaitEvent(c.copy_stream, sc.consumed[b], 0));
        if (per) CTX_CUDA_OK(cudaMemcpyAsync(sc.scalars[b].p, scalars32 + (size_t)done * per * 4, (size_t)cur * per * 4, cudaMemcpyHostToDevice, c.copy_stream));
        CTX_CUDA_OK(cudaEventRecord(sc.copied[b], c.copy_stream));
        CTX_CUDA_OK(cudaStreamWaitEvent(c.stream, sc.copied[b], 0));
        cc.fixed->run(sc.scalars[b].p, cur, n, out + done, c.stream);
        CTX_CUDA_OK(cudaEventRecord(sc.consumed[b], c.stream));
    }
    launch_affine_from_mont(curve, out, can, nmsm, c.stream);
    c.launches += 1;
    CTX_CUDA_OK(cudaMemcpyAsync(out64, can, (size_t)nmsm * 64, cudaMemcpyDeviceToHost, c.stream));
    throw_on_engine_error(*cc.fixed, c.stream);  // synchronises
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
    AbiScratch &sc = scratch();
    uint32_t *d_sc = sc.scalars[0].reserve((size_t)n * 8);
    uint32_t *d_pts_can = sc.pts_can.reserve((size_t)n * 16);
    affine *d_pts = sc.pts.reserve(n);
    affine *d_out = sc.out.reserve(1);
    uint32_t *d_can = sc.out_can.reserve(16);
    uint8_t *d_bad = sc.bytes.reserve(4);
    CTX_CUDA_OK(cudaMemsetAsync(d_bad, 0, 4, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(d_sc, scalars32, (size_t)n * 32, cudaMemcpyHostToDevice, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(d_pts_can, points64, (size_t)n * 64, cudaMemcpyHostToDevice, c.stream));
    launch_affine_to_mont_checked(curve, d_pts_can, d_pts, n, (uint32_t *)d_bad, c.stream);
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
    cc.var->set_bases(d_pts, n, cfg, c.stream);
    cc.var->run(d_sc, 1, n, d_out, c.stream);
    launch_affine_from_mont(curve, d_out, d_can, 1, c.stream);
    c.launches += 2;
    CTX_CUDA_OK(cudaMemcpyAsync(out64, d_can, 64, cudaMemcpyDeviceToHost, c.stream));
    uint32_t bad = 0;
    CTX_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, c.stream));
    throw_on_engine_error(*cc.var, c.stream);  // synchronises
    if (bad) throw std::runtime_error("msm: a base point is non-canonical or not on the curve");
    return 0;
    ABI_CATCH
}

int mina_b200_fixed_base_load(int curve, uint32_t n, const uint8_t *points64, int window_bits) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    if (n == 0) throw std::runtime_error("fixed_base_load: empty base set");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    CurveCtx &cc = c.curve[curve];
    AbiScratch &sc = scratch();
    cc.user.reset();
    if (cc.d_user) cudaFree(cc.d_user);
    cc.d_user = nullptr;
    cc.user_n = 0;
    CTX_CUDA_OK(cudaMalloc(&cc.d_user, (size_t)n * sizeof(affine)));
    uint32_t *d_can = sc.pts_can.reserve((size_t)n * 16);
    uint8_t *d_bad = sc.bytes.reserve(4);
    CTX_CUDA_OK(cudaMemsetAsync(d_bad, 0, 4, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(d_can, points64, (size_t)n * 64, cudaMemcpyHostToDevice, c.stream));
    launch_affine_to_mont_checked(curve, d_can, cc.d_user, n, (uint32_t *)d_bad, c.stream);
    c.launches += 1;
    uint32_t bad = 0;
    CTX_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    if (bad) throw std::runtime_error("fixed_base_load: a base point is non-canonical or not on the curve");
    MsmConfig cfg;
    cfg.precompute = true;
    cfg.c = window_bits > 0 ? window_bits : 16;
    cc.user.reset(make_msm_engine(curve));
    cc.user->set_bases(cc.d_user, n, cfg, c.stream);
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    cc.user_n = n;
    return 0;
    ABI_CATCH
}

int mina_b200_fixed_base_msm_device(int curve, uint32_t nmsm, const void *d_scalars, void *d_out64, void *cuda_stream,
                                    float *accumulate_ms) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    CurveCtx &cc = c.curve[curve];
    if (!cc.user) throw std::runtime_error("fixed_base_msm: no base set loaded");
    cudaStream_t s = (cudaStream_t)cuda_stream;
    AbiScratch &sc = scratch();
    affine *out = sc.out.reserve(std::max<uint32_t>(nmsm, 1));
    cc.user->enable_kernel_timing(accumulate_ms != nullptr);
    cc.user->run((const uint32_t *)d_scalars, nmsm, cc.user_n, out, s);
    launch_affine_from_mont(curve, out, (uint32_t *)d_out64, nmsm, s);
    c.launches += 1;
    if (accumulate_ms) {
        CTX_CUDA_OK(cudaStreamSynchronize(s));
        *accumulate_ms = cc.user->last_accumulate_ms();
    }
    return 0;
    ABI_CATCH
}

int mina_b200_fixed_base_msm(int curve, uint32_t nmsm, const uint8_t *scalars32, uint8_t *out64) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    CurveCtx &cc = c.curve[curve];
    if (!cc.user) throw std::runtime_error("fixed_base_msm: no base set loaded");
    if (nmsm == 0) return 0;
    AbiScratch &sc = scratch();
    const size_t per = (size_t)cc.user_n * 8;
    affine *out = sc.out.reserve(nmsm);
    uint32_t *can = sc.out_can.reserve((size_t)nmsm * 16);
    cc.user->enable_kernel_timing(false);
    for (uint32_t k = 0; k < nmsm; k++) {  // one MSM per chunk, double-buffered like mina_b200_msm_srs
        int b = k & 1;
        uint32_t *d_sc = sc.scalars[b].reserve(per);
        if (k >= 2) CTX_CUDA_OK(cudaStreamWaitEvent(c.copy_stream, sc.consumed[b], 0));
        CTX_CUDA_OK(cudaMemcpyAsync(d_sc, scalars32 + (size_t)k * per * 4, per * 4, cudaMemcpyHostToDevice, c.copy_stream));
        CTX_CUDA_OK(cudaEventRecord(sc.copied[b], c.copy_stream));
        CTX_CUDA_OK(cudaStreamWaitEvent(c.stream, sc.copied[b], 0));
        cc.user->run(d_sc, 1, cc.user_n, out + k, c.stream);
        CTX_CUDA_OK(cudaEventRecord(sc.consumed[b], c.stream));
    }
    launch_affine_from_mont(curve, out, can, nmsm, c.stream);
    c.launches += 1;
    CTX_CUDA_OK(cudaMemcpyAsync(out64, can, (size_t)nmsm * 64, cudaMemcpyDeviceToHost, c.stream));
    if (cc.user->take_error(c.stream) & 1u) throw std::runtime_error("msm: a scalar is >= 2^255 (not a canonical field element); result discarded");
    return 0;
    ABI_CATCH
}

// ---- K4 / K2 / K5 / K3 hooks: host buffers in, host buffers out, canonical field elements --------------
int mina_b200_endo_to_field(int field, uint32_t n, const uint8_t *pre16, uint8_t *out32) {
    ABI_TRY
    require_ready();
    if (!n) return 0;
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    AbiScratch &sc = scratch();
    uint8_t *d_pre = sc.bytes.reserve((size_t)n * 16);
    fe *d_m = sc.a.reserve(n), *d_o = sc.b.reserve(n);
    CTX_CUDA_OK(cudaMemcpyAsync(d_pre, pre16, (size_t)n * 16, cudaMemcpyHostToDevice, c.stream));
    launch_endo_to_field(field, d_pre, d_m, n, c.stream);
    launch_fe_from_mont(field, d_m, d_o, n, c.stream);
    c.launches += 2;
    CTX_CUDA_OK(cudaMemcpyAsync(out32, d_o, (size_t)n * 32, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    return 0;
    ABI_CATCH
}

int mina_b200_bpoly_coeffs(int field, uint32_t nproofs, int k, const uint8_t *chals32, uint8_t *out32) {
    ABI_TRY
    require_ready();
    if (!nproofs) return 0;
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    AbiScratch &sc = scratch();
    size_t nch = (size_t)nproofs * k, total = (size_t)nproofs << k;
    fe *d_c = sc.a.reserve(nch), *d_cm = sc.b.reserve(nch), *d_t = sc.c.reserve((size_t)nproofs * BPOLY_TABLE), *d_o = sc.d.reserve(total);
    CTX_CUDA_OK(cudaMemcpyAsync(d_c, chals32, nch * 32, cudaMemcpyHostToDevice, c.stream));
    launch_fe_to_mont(field, d_c, d_cm, (uint32_t)nch, c.stream);
    launch_bpoly_tables(field, d_cm, d_t, nproofs, k, nullptr, true, c.stream);
    launch_bpoly_materialize(field, d_t, d_o, nproofs, k, c.stream);
    c.launches += 3;
    CTX_CUDA_OK(cudaMemcpyAsync(out32, d_o, total * 32, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    return 0;
    ABI_CATCH
}

int mina_b200_bpoly_combine(int field, uint32_t nproofs, int k, const uint8_t *chals32, const uint8_t *r32, uint8_t *out32) {
    ABI_TRY
    require_ready();
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    AbiScratch &sc = scratch();
    size_t nch = (size_t)nproofs * k, total = (size_t)1 << k;
    fe *d_in = sc.a.reserve(nch + nproofs + 1), *d_m = sc.b.reserve(nch + nproofs + 1);
    fe *d_t = sc.c.reserve((size_t)std::max<uint32_t>(nproofs, 1) * BPOLY_TABLE), *d_o = sc.d.reserve(total);
    if (nproofs) {
        CTX_CUDA_OK(cudaMemcpyAsync(d_in, chals32, nch * 32, cudaMemcpyHostToDevice, c.stream));
        CTX_CUDA_OK(cudaMemcpyAsync(d_in + nch, r32, (size_t)nproofs * 32, cudaMemcpyHostToDevice, c.stream));
        launch_fe_to_mont(field, d_in, d_m, (uint32_t)(nch + nproofs), c.stream);
        launch_bpoly_tables(field, d_m, d_t, nproofs, k, d_m + nch, false, c.stream);
    }
    launch_bpoly_combine(field, d_t, nullptr, nullptr, 0, nproofs, k, d_o, c.stream);
    c.launches += 3;
    CTX_CUDA_OK(cudaMemcpyAsync(out32, d_o, total * 32, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    return 0;
    ABI_CATCH
}

int mina_b200_bpoly_eval(int field, uint32_t nproofs, uint32_t npts, int k, const uint8_t *chals32, const uint8_t *x32, uint8_t *out32) {
    ABI_TRY
    require_ready();
    size_t nch = (size_t)nproofs * k, nx = (size_t)nproofs * npts;
    if (!nx) return 0;
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    AbiScratch &sc = scratch();
    fe *d_in = sc.a.reserve(nch + nx), *d_m = sc.b.reserve(nch + nx), *d_o = sc.c.reserve(nx), *d_oc = sc.d.reserve(nx);
    CTX_CUDA_OK(cudaMemcpyAsync(d_in, chals32, nch * 32, cudaMemcpyHostToDevice, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(d_in + nch, x32, nx * 32, cudaMemcpyHostToDevice, c.stream));
    launch_fe_to_mont(field, d_in, d_m, (uint32_t)(nch + nx), c.stream);
    launch_bpoly_eval(field, d_m, d_m + nch, d_o, nproofs, npts, k, c.stream);
    launch_fe_from_mont(field, d_o, d_oc, (uint32_t)nx, c.stream);
    c.launches += 3;
    CTX_CUDA_OK(cudaMemcpyAsync(out32, d_oc, nx * 32, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    return 0;
    ABI_CATCH
}

int mina_b200_combined_inner_product(int field, uint32_t nproofs, uint32_t npolys, uint32_t npts, const uint8_t *evals32,
                                      const uint8_t *scales32, uint8_t *out32) {
    ABI_TRY
    require_ready();
    if (!nproofs) return 0;
    size_t ne = (size_t)nproofs * npolys * npts, ns = (size_t)nproofs * 2;
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    AbiScratch &sc = scratch();
    fe *d_in = sc.a.reserve(ne + ns), *d_m = sc.b.reserve(ne + ns), *d_o = sc.c.reserve(nproofs), *d_oc = sc.d.reserve(nproofs);
    if (ne) CTX_CUDA_OK(cudaMemcpyAsync(d_in, evals32, ne * 32, cudaMemcpyHostToDevice, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(d_in + ne, scales32, ns * 32, cudaMemcpyHostToDevice, c.stream));
    launch_fe_to_mont(field, d_in, d_m, (uint32_t)(ne + ns), c.stream);
    launch_combined_inner_product(field, d_m, d_m + ne, d_o, nproofs, npolys, npts, c.stream);
    launch_fe_from_mont(field, d_o, d_oc, nproofs, c.stream);
    c.launches += 3;
    CTX_CUDA_OK(cudaMemcpyAsync(out32, d_oc, (size_t)nproofs * 32, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    return 0;
    ABI_CATCH
}

// Shape of the IPA final-check MSM: scalars over the resident SRS prefix g[0..n_srs) PLUS a
// handful of per-proof points ({h, sg, U, C_k, L_j, R_j, delta}: ~80) in one result.
int mina_b200_msm_srs_plus(int curve, uint32_t n_srs, const uint8_t *scalars_srs32, uint32_t n_extra, const uint8_t *scalars_extra32,
                           const uint8_t *points_extra64, uint8_t *out64) {
    ABI_TRY
    require_ready();
    if (curve < 0 || curve > 1) throw std::runtime_error("bad curve id");
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    CurveCtx &cc = c.curve[curve];
    if (n_srs > cc.depth) throw std::runtime_error("msm_srs_plus: n_srs exceeds the resident SRS depth");
    AbiScratch &sc = scratch();
    affine *d_res = sc.out.reserve(2);
    uint32_t *d_can = sc.out_can.reserve(32);
    uint8_t *d_bad = sc.bytes.reserve(4);
    CTX_CUDA_OK(cudaMemsetAsync(d_res, 0, 2 * sizeof(affine), c.stream));
    CTX_CUDA_OK(cudaMemsetAsync(d_bad, 0, 4, c.stream));
    if (n_srs) {
        uint32_t *d_s = sc.scalars[0].reserve((size_t)n_srs * 8);
        CTX_CUDA_OK(cudaMemcpyAsync(d_s, scalars_srs32, (size_t)n_srs * 32, cudaMemcpyHostToDevice, c.stream));
        cc.fixed->enable_kernel_timing(false);
        cc.fixed->run(d_s, 1, n_srs, d_res, c.stream);
    }
    if (n_extra) {
        uint32_t *d_t = sc.scalars[1].reserve((size_t)n_extra * 8);
        uint32_t *d_pc = sc.pts_can.reserve((size_t)n_extra * 16);
        affine *d_p = sc.pts.reserve(n_extra);
        CTX_CUDA_OK(cudaMemcpyAsync(d_t, scalars_extra32, (size_t)n_extra * 32, cudaMemcpyHostToDevice, c.stream));
        CTX_CUDA_OK(cudaMemcpyAsync(d_pc, points_extra64, (size_t)n_extra * 64, cudaMemcpyHostToDevice, c.stream));
        launch_affine_to_mont_checked(curve, d_pc, d_p, n_extra, (uint32_t *)d_bad, c.stream);
        MsmConfig cfg;
        cfg.precompute = false;
        cfg.c = 8;
        cc.var->set_bases(d_p, n_extra, cfg, c.stream);
        cc.var->run(d_t, 1, n_extra, d_res + 1, c.stream);
    }
    launch_affine_from_mont(curve, d_res, d_can, 2, c.stream);
    c.launches += 2;
    uint8_t parts[128];
    uint32_t bad = 0;
    CTX_CUDA_OK(cudaMemcpyAsync(parts, d_can, 128, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, c.stream));
    uint32_t e = (n_srs ? cc.fixed->take_error(c.stream) : 0) | (n_extra ? cc.var->take_error(c.stream) : 0);
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    if (e & 1u) throw std::runtime_error("msm: a scalar is >= 2^255 (not a canonical field element); result discarded");
    if (bad) throw std::runtime_error("msm: a base point is non-canonical or not on the curve");
    // the last addition of two points happens on the host
    auto add2 = [&](auto tag) {
        using B = decltype(tag);
        auto load = [&](const uint8_t *b) {
            host::Affine<B> a;
            bool zero = true;
            for (int i = 0; i < 64; i++) zero = zero && b[i] == 0;
            if (zero) return host::Affine<B>::identity();
            host::Fe<B>::from_bytes_le(b, a.x);
            host::Fe<B>::from_bytes_le(b + 32, a.y);
            a.inf = false;
            return a;
        };
        host::Affine<B> r = host::Jac<B>::from_affine(load(parts)).add_affine(load(parts + 64)).to_affine();
        std::memset(out64, 0, 64);
        if (!r.inf) {
            r.x.to_bytes_le(out64);
            r.y.to_bytes_le(out64 + 32);
        }
    };
    if (curve == 0)
        add2(FpParams{});
    else
        add2(FqParams{});
    return 0;
    ABI_CATCH
}

static void table_to_device(int field, const uint8_t *table, fe *d_raw, fe *d_mont, cudaStream_t s) {
    CTX_CUDA_OK(cudaMemcpyAsync(d_raw, table, (size_t)POSEIDON_TABLE_WORDS * 32, cudaMemcpyHostToDevice, s));
    launch_fe_to_mont(field, d_raw, d_mont, POSEIDON_TABLE_WORDS, s);
}

int mina_b200_poseidon_permute(int field, const uint8_t *table, uint32_t n, uint8_t *states96) {
    ABI_TRY
    require_ready();
    if (!n) return 0;
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    CTX_CUDA_OK(cudaSetDevice(c.device));
    AbiScratch &sc = scratch();
    fe *d_raw = sc.a.reserve(POSEIDON_TABLE_WORDS), *d_tab = sc.b.reserve(POSEIDON_TABLE_WORDS), *d_st = sc.c.reserve((size_t)n * 3);
    table_to_device(field, table, d_raw, d_tab, c.stream);
    CTX_CUDA_OK(cudaMemcpyAsync(d_st, states96, (size_t)n * 96, cudaMemcpyHostToDevice, c.stream));
    launch_poseidon_permute(field, d_tab, d_st, n, c.stream);
    c.launches += 2;
    CTX_CUDA_OK(cudaMemcpyAsync(states96, d_st, (size_t)n * 96, cudaMemcpyDeviceToHost, c.stream));
    CTX_CUDA_OK(cudaStreamSynchronize(c.stream));
    return 0;
    ABI_CATCH
}

int mina_b200_poseidon_trusted(void) { return ctx().ready && ctx().poseidon_trusted ? 1 : 0; }

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
            case 16: r0 = Fd<F>::mul_ptx(a[i], b[i]); break;   // first-generation multiplier
            case 17: r0 = Fd<F>::mul_ptx2(a[i], b[i]); break;  // IMAD.WIDE-chain reduction (the default)
            // sums of products with one reduction (lazy reduction): elements i, i+1, i+2 (cyclically)
            case 18: r0 = Fd<F>::dot2(a[i], b[i], a[(i + 1) % n], b[(i + 1) % n]); break;
            case 19: r0 = Fd<F>::dot3(a[i], b[i], a[(i + 1) % n], b[(i + 1) % n], a[(i + 2) % n], b[(i + 2) % n]); break;
#endif
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
