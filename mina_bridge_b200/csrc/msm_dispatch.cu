// Curve-id dispatch for the MSM engine (0 = Pallas / Fp coordinates, 1 = Vesta / Fq coordinates).
#include <stdexcept>

#include "msm.cuh"

namespace pasta {
MsmEngineBase *make_msm_engine_fp();
MsmEngineBase *make_msm_engine_fq();
void launch_affine_to_mont_fp(const uint32_t *, affine *, uint32_t, cudaStream_t);
void launch_affine_to_mont_fq(const uint32_t *, affine *, uint32_t, cudaStream_t);
void launch_affine_to_mont_checked_fp(const uint32_t *, affine *, uint32_t, uint32_t *, cudaStream_t);
void launch_affine_to_mont_checked_fq(const uint32_t *, affine *, uint32_t, uint32_t *, cudaStream_t);
void launch_affine_from_mont_fp(const affine *, uint32_t *, uint32_t, cudaStream_t);
void launch_affine_from_mont_fq(const affine *, uint32_t *, uint32_t, cudaStream_t);

MsmEngineBase *make_msm_engine(int curve) {
    if (curve == 0) return make_msm_engine_fp();
    if (curve == 1) return make_msm_engine_fq();
    throw std::runtime_error("msm: unknown curve id");
}
void launch_affine_to_mont(int curve, const uint32_t *d_in, affine *d_out, uint32_t n, cudaStream_t s) {
    curve == 0 ? launch_affine_to_mont_fp(d_in, d_out, n, s) : launch_affine_to_mont_fq(d_in, d_out, n, s);
}
void launch_affine_to_mont_checked(int curve, const uint32_t *d_in, affine *d_out, uint32_t n, uint32_t *d_bad, cudaStream_t s) {
    curve == 0 ? launch_affine_to_mont_checked_fp(d_in, d_out, n, d_bad, s) : launch_affine_to_mont_checked_fq(d_in, d_out, n, d_bad, s);
}
void launch_affine_from_mont(int curve, const affine *d_in, uint32_t *d_out, uint32_t n, cudaStream_t s) {
    curve == 0 ? launch_affine_from_mont_fp(d_in, d_out, n, s) : launch_affine_from_mont_fq(d_in, d_out, n, s);
}
}  // namespace pasta
