// Launchers for the IPA scalar-side kernels (ipa.cuh), dispatched on the scalar field id.
#include <stdexcept>

#include "ipa.cuh"

namespace pasta {

static void check_k(int k) {
    if (k < BPOLY_LO_BITS || k > 2 * BPOLY_LO_BITS) throw std::runtime_error("bpoly: rounds must be in [8, 16]");
}
static void check_field(int field) {
    if (field != 0 && field != 1) throw std::runtime_error("bad field id");
}

void launch_endo_to_field(int field, const void *d_pre16, fe *d_out, uint32_t n, cudaStream_t s) {
    check_field(field);
    if (!n) return;
    dim3 g((n + 127) / 128);
    if (field == 0)
        k_endo_to_field<FpParams><<<g, 128, 0, s>>>((const uint4 *)d_pre16, d_out, n);
    else
        k_endo_to_field<FqParams><<<g, 128, 0, s>>>((const uint4 *)d_pre16, d_out, n);
}

void launch_bpoly_tables(int field, const fe *d_chals, fe *d_tables, uint32_t nproofs, int k, const fe *d_scale, bool lo_plain,
                         cudaStream_t s) {
    check_field(field);
    check_k(k);
    if (!nproofs) return;
    dim3 g((nproofs * BPOLY_TABLE + 255) / 256);
    if (field == 0)
        k_bpoly_tables<FpParams><<<g, 256, 0, s>>>(d_chals, d_tables, nproofs, k, d_scale, lo_plain ? 1 : 0);
    else
        k_bpoly_tables<FqParams><<<g, 256, 0, s>>>(d_chals, d_tables, nproofs, k, d_scale, lo_plain ? 1 : 0);
}

void launch_bpoly_materialize(int field, const fe *d_tables, fe *d_out, uint32_t nproofs, int k, cudaStream_t s) {
    check_field(field);
    check_k(k);
    if (!nproofs) return;
    uint64_t total = (uint64_t)nproofs << k;
    dim3 g((unsigned)((total + 255) / 256));
    if (field == 0)
        k_bpoly_materialize<FpParams><<<g, 256, 0, s>>>(d_tables, d_out, nproofs, k);
    else
        k_bpoly_materialize<FqParams><<<g, 256, 0, s>>>(d_tables, d_out, nproofs, k);
}

void launch_bpoly_combine(int field, const fe *d_tables, const uint32_t *d_subset, const uint32_t *d_group_off, uint32_t ngroups,
                          uint32_t nsub, int k, fe *d_out, cudaStream_t s) {
    check_field(field);
    check_k(k);
    uint32_t n_hi = 1u << (k - BPOLY_LO_BITS);
    uint32_t threads = 256u * ((n_hi + COMBINE_ITEMS - 1) / COMBINE_ITEMS);
    dim3 g((threads + 127) / 128, ngroups ? ngroups : 1);
    const uint32_t *goff = ngroups ? d_group_off : nullptr;
    if (field == 0)
        k_bpoly_combine<FpParams><<<g, 128, 0, s>>>(d_tables, d_subset, goff, nsub, k, d_out);
    else
        k_bpoly_combine<FqParams><<<g, 128, 0, s>>>(d_tables, d_subset, goff, nsub, k, d_out);
}

void launch_bpoly_eval(int field, const fe *d_chals, const fe *d_x, fe *d_out, uint32_t nproofs, uint32_t npts, int k,
                       cudaStream_t s) {
    check_field(field);
    if (k < 1 || k > 32) throw std::runtime_error("bpoly_eval: bad round count");
    uint32_t n = nproofs * npts;
    if (!n) return;
    dim3 g((n + 127) / 128);
    if (field == 0)
        k_bpoly_eval<FpParams><<<g, 128, 0, s>>>(d_chals, d_x, d_out, nproofs, npts, k);
    else
        k_bpoly_eval<FqParams><<<g, 128, 0, s>>>(d_chals, d_x, d_out, nproofs, npts, k);
}

void launch_combined_inner_product(int field, const fe *d_evals, const fe *d_scales, fe *d_out, uint32_t nproofs, uint32_t npolys,
                                   uint32_t npts, cudaStream_t s) {
    check_field(field);
    if (!nproofs) return;
    dim3 g((nproofs + 127) / 128);
    if (field == 0)
        k_combined_inner_product<FpParams><<<g, 128, 0, s>>>(d_evals, d_scales, d_out, nproofs, npolys, npts);
    else
        k_combined_inner_product<FqParams><<<g, 128, 0, s>>>(d_evals, d_scales, d_out, nproofs, npolys, npts);
}

void launch_fe_to_mont(int field, const fe *d_in, fe *d_out, uint32_t n, cudaStream_t s) {
    check_field(field);
    if (!n) return;
    dim3 g((n + 255) / 256);
    if (field == 0)
        k_fe_to_mont<FpParams><<<g, 256, 0, s>>>(d_in, d_out, n);
    else
        k_fe_to_mont<FqParams><<<g, 256, 0, s>>>(d_in, d_out, n);
}
void launch_fe_from_mont(int field, const fe *d_in, fe *d_out, uint32_t n, cudaStream_t s) {
    check_field(field);
    if (!n) return;
    dim3 g((n + 255) / 256);
    if (field == 0)
        k_fe_from_mont<FpParams><<<g, 256, 0, s>>>(d_in, d_out, n);
    else
        k_fe_from_mont<FqParams><<<g, 256, 0, s>>>(d_in, d_out, n);
}

}  // namespace pasta

namespace pasta {
void launch_lagrange_scalars(int field, const fe &omega_inv_mont, const fe &n_inv_mont, int log_n, uint32_t first, uint32_t count, fe *d_out,
                             cudaStream_t s) {
    if (field != 0 && field != 1) throw std::runtime_error("bad field id");
    const uint64_t total = (uint64_t)count << log_n;
    if (!total) return;
    dim3 g((unsigned)((total + 255) / 256));
    if (field == 0)
        k_lagrange_scalars<FpParams><<<g, 256, 0, s>>>(omega_inv_mont, n_inv_mont, log_n, first, count, d_out);
    else
        k_lagrange_scalars<FqParams><<<g, 256, 0, s>>>(omega_inv_mont, n_inv_mont, log_n, first, count, d_out);
}
}  // namespace pasta
