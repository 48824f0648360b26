// MSM engine instantiation for curves whose coordinates live in FqParams (see msm_impl.cuh).
#include "msm_impl.cuh"

namespace pasta {
MsmEngineBase *make_msm_engine_fq() { return new MsmEngine<FqParams>(); }
void launch_affine_to_mont_fq(const uint32_t *d_in, affine *d_out, uint32_t n, cudaStream_t s) {
    launch_affine_to_mont_t<FqParams>(d_in, d_out, n, s);
}
void launch_affine_from_mont_fq(const affine *d_in, uint32_t *d_out, uint32_t n, cudaStream_t s) {
    launch_affine_from_mont_t<FqParams>(d_in, d_out, n, s);
}
void launch_affine_to_mont_checked_fq(const uint32_t *d_in, affine *d_out, uint32_t n, uint32_t *d_bad, cudaStream_t s) {
    launch_affine_to_mont_checked_t<FqParams>(d_in, d_out, n, d_bad, s);
}
}  // namespace pasta
