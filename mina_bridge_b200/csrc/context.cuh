// Process-wide device context: resident SRS, MSM engines, verification keys, Poseidon tables,
// persistent staging.  The CUDA analogue of the reference's lazy statics MINA_SRS,
// DEVNET_VERIFIER_INDEX and MAINNET_VERIFIER_INDEX (AL/operator/mina/lib/src/lib.rs:23-35).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "host_field.hpp"
#include "msm.cuh"
#include "poseidon.hpp"
#include "srs.hpp"
#include "vk.hpp"

namespace pasta {

#define CTX_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));                   \
    } while (0)

static constexpr uint32_t VESTA_SRS_DEPTH = 1u << 16;   // Fq::SRS_DEPTH, lib.rs:34
static constexpr uint32_t PALLAS_SRS_DEPTH = 1u << 15;  // index.max_poly_size, devnet_vk.json

struct CurveCtx {
    uint32_t depth = 0;
    affine *d_srs = nullptr;  // depth + 1 points (last = h), Montgomery
    std::unique_ptr<MsmEngineBase> fixed;  // over the resident SRS
    std::unique_ptr<MsmEngineBase> var;    // caller-supplied bases
    std::unique_ptr<MsmEngineBase> user;   // caller-supplied bases that stay resident (fixed-base table)
    affine *d_user = nullptr;
    uint32_t user_n = 0;
    // public-input commitment: fixed-base engine over the first `lagr_n` Lagrange commitments of the 2^lagr_log_n domain + h
    std::unique_ptr<MsmEngineBase> lagr;
    affine *d_lagr = nullptr;
    uint32_t lagr_n = 0, lagr_log_n = 0;
    MsmConfig cfg;
    std::vector<uint8_t> host_canonical;  // (depth + 1) x 64 bytes canonical, for tests / host logic
};

// Grow-only device / pinned-host buffers: nothing on a hot entry point calls cudaMalloc once warm.
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { reserve(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    T *reserve(size_t count) {
        if (count > n) {
            release();
            size_t want = count + count / 4 + 16;
            CTX_CUDA_OK(cudaMalloc(&p, want * sizeof(T)));
            n = want;
        }
        return p;
    }
    void alloc(size_t count) { reserve(count); }
};
template <class T>
struct PinnedBuf {
    T *p = nullptr;
    size_t n = 0;
    PinnedBuf() {}
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() { release(); }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
    T *reserve(size_t count) {
        if (count > n) {
            release();
            size_t want = count + count / 4 + 16;
            CTX_CUDA_OK(cudaHostAlloc(&p, want * sizeof(T), cudaHostAllocDefault));
            n = want;
        }
        return p;
    }
};

struct VerifierState;  // verifier.cu

struct Context {
    int device = -1;
    bool ready = false;
    cudaStream_t stream = nullptr;   // compute
    cudaStream_t copy_stream = nullptr;  // H2D staging that overlaps compute; second pipeline of a state batch
    cudaStream_t aux_stream[2] = {nullptr, nullptr};  // per pipeline: work that runs beside the MSM (r_j * C_j)
    CurveCtx curve[2];
    host::Srs<FpParams> srs_pallas;  // coordinates in Fp
    host::Srs<FqParams> srs_vesta;   // coordinates in Fq
    std::string data_dir;
    // verification keys (index 0 = mainnet, 1 = devnet: `is_state_proof_from_devnet`)
    vk::VerifierIndex vk[2];
    bool vk_loaded = false;
    std::string vk_error;
    // Poseidon tables (0 = Fp, 1 = Fq); `trusted` only if the Fp table passes the reference's KAT
    poseidon::Params<FpParams> poseidon_fp;
    poseidon::Params<FqParams> poseidon_fq;
    bool poseidon_trusted = false;
    fe *d_poseidon_tab[2] = {nullptr, nullptr};
    std::mutex mu;  // the device lock: one GPU work item (a whole coalesced batch) at a time
    std::atomic<uint64_t> launches{0};
    VerifierState *verifier = nullptr;  // owned; created / released by verifier.cu
};

Context &ctx();
void set_error(const std::string &msg);
void require_ready();
uint64_t engine_launches();
void verifier_release(Context &c);

}  // namespace pasta
