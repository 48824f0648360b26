// Pippenger bucket MSM over Pallas / Vesta for sm_100a -- engine interface.
//
// Replaces ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul` (called under
// `verify_block`, AL/operator/mina/lib/src/lib.rs:99-111).  See msm.cu for the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "ec.cuh"

namespace pasta {

struct MsmConfig {
    int c = 16;              // window bits (signed digits, 2^(c-1) buckets per window)
    bool precompute = true;  // fixed-base table T[w][i] = 2^(c*w) * G_i  (all windows share buckets)
    int leaf = 8;            // buckets per thread in the running-sum reduction
};

// One engine = one curve + one resident base set.  Not thread-safe; callers serialise per engine.
class MsmEngineBase {
   public:
    virtual ~MsmEngineBase() {}
    // bases: device pointer to n affine points (Montgomery).  The engine keeps the pointer (and, with
    // precompute, builds its own table of W*n points).
    virtual void set_bases(const affine *d_bases, uint32_t n, const MsmConfig &cfg, cudaStream_t s) = 0;
    // scalars: device pointer, nmsm * n * 8 u32 (canonical little-endian integers < 2^255).
    // out: device pointer to nmsm affine points (Montgomery; (0,0) = identity).
    // n_used <= n lets a caller run over a prefix of the bases.
    virtual void run(const uint32_t *d_scalars, uint32_t nmsm, uint32_t n_used, affine *d_out, cudaStream_t s) = 0;
    // Same, but MSM m takes its 2^k scalars from b_poly_coefficients of proof m, given as the proof's
    // two product tables (ipa.cuh: lo plain, hi Montgomery); the coefficient vector is never stored.
    virtual void run_bpoly(const fe *d_tables, uint32_t nmsm, int k, affine *d_out, cudaStream_t s) = 0;
    // Same two entry points without the final normalisation: results stay in XYZZ form (no inversion).
    virtual void run_xyzz(const uint32_t *d_scalars, uint32_t nmsm, uint32_t n_used, xyzz *d_out, cudaStream_t s) = 0;
    virtual void run_bpoly_xyzz(const fe *d_tables, uint32_t nmsm, int k, xyzz *d_out, cudaStream_t s) = 0;
    virtual size_t workspace_bytes() const = 0;
    // kernels launched by this engine since construction (counted where they are issued)
    virtual uint64_t launches() const = 0;
    // Reads and clears the device error flag (synchronises `s`): bit 0 = a scalar was >= 2^255 and
    // does not fit the signed-digit windows, the affected result is invalid.
    virtual uint32_t take_error(cudaStream_t s) = 0;
    // duration of the dominant kernel (bucket accumulation) of the last run, if timing was enabled
    virtual void enable_kernel_timing(bool on) = 0;
    virtual float last_accumulate_ms() = 0;
};

MsmEngineBase *make_msm_engine(int curve);  // 0 = Pallas (coords in Fp), 1 = Vesta (coords in Fq)

// Small helpers shared with the rest of the library (implemented in msm.cu)
void launch_affine_to_mont(int curve, const uint32_t *d_canonical_xy, affine *d_out, uint32_t n, cudaStream_t s);
// untrusted points: *d_bad |= 1 if any is non-canonical or off-curve (such points are replaced by the identity)
void launch_affine_to_mont_checked(int curve, const uint32_t *d_canonical_xy, affine *d_out, uint32_t n, uint32_t *d_bad, cudaStream_t s);
void launch_affine_from_mont(int curve, const affine *d_in, uint32_t *d_canonical_xy, uint32_t n, cudaStream_t s);

}  // namespace pasta
