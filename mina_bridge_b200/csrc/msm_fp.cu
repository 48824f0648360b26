// MSM engine instantiation for curves whose coordinates live in FpParams (see msm_impl.cuh).
#include "msm_impl.cuh"

namespace pasta {
MsmEngineBase *make_msm_engine_fp() { return new MsmEngine<FpParams>(); }
void launch_affine_to_mont_fp(const uint32_t *d_in, affine *d_out, uint32_t n, cudaStream_t s) {
    launch_affine_to_mont_t<FpParams>(d_in, d_out, n, s);
}
void launch_affine_from_mont_fp(const affine *d_in, uint32_t *d_out, uint32_t n, cudaStream_t s) {
    launch_affine_from_mont_t<FpParams>(d_in, d_out, n, s);
}
void launch_affine_to_mont_checked_fp(const uint32_t *d_in, affine *d_out, uint32_t n, uint32_t *d_bad, cudaStream_t s) {
    launch_affine_to_mont_checked_t<FpParams>(d_in, d_out, n, d_bad, s);
}
}  // namespace pasta
