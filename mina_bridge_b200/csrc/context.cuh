// Process-wide device context: resident SRS, MSM engines, scratch.  The CUDA analogue of the
// reference's lazy statics (AL/operator/mina/lib/src/lib.rs:23-35).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "host_field.hpp"
#include "msm.cuh"
#include "srs.hpp"

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
    MsmConfig cfg;
    std::vector<uint8_t> host_canonical;  // (depth + 1) x 64 bytes canonical, for tests / host logic
};

struct Context {
    int device = -1;
    bool ready = false;
    cudaStream_t stream = nullptr;
    CurveCtx curve[2];
    host::Srs<FpParams> srs_pallas;  // coordinates in Fp
    host::Srs<FqParams> srs_vesta;   // coordinates in Fq
    std::mutex mu;                   // serialises GPU work issued through the C ABI
    std::atomic<uint64_t> launches{0};
};

Context &ctx();
void set_error(const std::string &msg);
void require_ready();

// RAII device buffer
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    void alloc(size_t count) {
        if (p) cudaFree(p);
        p = nullptr;
        n = count;
        if (count) CTX_CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
    }
};

}  // namespace pasta
