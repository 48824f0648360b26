/*
 * mina_b200.h -- C ABI of the B200-native Mina proof verifier.
 *
 * Drop-in boundary: the two entry points the Aligned operator / batcher bind today
 *   verify_mina_state_ffi        (AL/operator/mina/lib/src/lib.rs:41-113,
 *                                 header AL/operator/mina/lib/mina_verifier.h:3-6,
 *                                 cgo caller AL/operator/mina/mina.go:27-32)
 *   verify_account_inclusion_ffi (AL/operator/mina_account/lib/src/lib.rs:16-78,
 *                                 header AL/operator/mina_account/lib/mina_account_verifier.h:3-6,
 *                                 cgo caller AL/operator/mina_account/mina_account.go:27-32)
 * are declared in mina_verifier.h / mina_account_verifier.h next to this file with the reference's
 * exact signatures.  Everything below is ADDITIVE: lifecycle, batch entry points and the individual
 * hot-path kernels (MSM, IPA scalar helpers, Poseidon) exposed for parity tests and benchmarks.
 *
 * Conventions: plain pointers and sizes only.  Field elements are 32-byte little-endian canonical
 * integers; affine points are x || y (64 bytes), all-zero = point at infinity.  Curve ids:
 * 0 = Pallas (coordinates in Fp, scalars in Fq), 1 = Vesta (coordinates in Fq, scalars in Fp).
 * Field ids: 0 = Fp, 1 = Fq.  Functions returning int give 0 on success, negative on error
 * (message via mina_b200_last_error()).  No CPU fallback exists: without a CUDA device every
 * compute entry point fails.
 */
#ifndef MINA_B200_H
#define MINA_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MINA_B200_CURVE_PALLAS 0
#define MINA_B200_CURVE_VESTA 1
#define MINA_B200_FIELD_FP 0
#define MINA_B200_FIELD_FQ 1

/* ---- lifecycle ------------------------------------------------------------------------------- */
/* Select `device`, build (or load from `cache_dir`, may be NULL) the two SRS, upload them and build
 * the fixed-base MSM tables.  Mirrors the reference's lazy statics MINA_SRS / *_VERIFIER_INDEX
 * (AL/operator/mina/lib/src/lib.rs:23-35).  Idempotent; thread-safe. */
int mina_b200_init(int device, const char *cache_dir);
void mina_b200_shutdown(void);
const char *mina_b200_last_error(void);
/* Number of kernels launched by this library since init (the `gpu_launches` claim of bench.py). */
uint64_t mina_b200_launch_count(void);

/* ---- SRS access (for tests: feed the same bases to the oracle) ---------------------------------- */
/* Copies g[first .. first+count) of the resident SRS as canonical affine points. `h64` may be NULL. */
int mina_b200_srs_points(int curve, uint32_t first, uint32_t count, uint8_t *out64, uint8_t *h64);

/* ---- K1: multi-scalar multiplication ---------------------------------------------------------- */
/* nmsm independent MSMs over the resident SRS prefix g[0..n): scalars is nmsm*n*32 bytes,
 * out64 is nmsm*64 bytes.  Host buffers; copies are part of the call.
 * Replaces VariableBaseMSM::multi_scalar_mul over the SRS. */
int mina_b200_msm_srs(int curve, uint32_t nmsm, uint32_t n, const uint8_t *scalars32, uint8_t *out64);
/* One MSM over caller-supplied bases (host buffers, canonical affine).  window_bits 0 = default. */
int mina_b200_msm(int curve, uint32_t n, const uint8_t *scalars32, const uint8_t *points64, int window_bits,
                  uint8_t *out64);
/* Device-resident variant used by bench.py's `value` leg: d_scalars is a CUDA device pointer to
 * nmsm*n*8 uint32 (canonical), d_out64 a device pointer to nmsm*16 uint32 (canonical affine).
 * Work is enqueued on `cuda_stream` (a cudaStream_t passed as void*); no synchronisation.
 * If `accumulate_ms` is non-NULL the call synchronises and stores the CUDA-event duration of the
 * dominant kernel (bucket accumulation) of the LAST chunk. */
int mina_b200_msm_srs_device(int curve, uint32_t nmsm, uint32_t n, const void *d_scalars, void *d_out64,
                             void *cuda_stream, float *accumulate_ms);
/* MSM engine tuning (takes effect at the next init / table rebuild): window bits and running-sum
 * chunk.  Returns 0 on success. */
int mina_b200_msm_configure(int curve, int window_bits, int precompute, int leaf);

/* ---- field self-test hook (parity tests of the device arithmetic) ------------------------------- */
/* op: 0 mul, 1 add, 2 sub, 3 inverse (b ignored), 4 square (b ignored).  n elements each. */
int mina_b200_field_op(int field, int op, uint32_t n, const uint8_t *a32, const uint8_t *b32, uint8_t *out32);
/* Group self-test: out[i] = a[i] + b[i] on `curve` through the XYZZ formulas (canonical affine). */
int mina_b200_point_add(int curve, uint32_t n, const uint8_t *a64, const uint8_t *b64, uint8_t *out64);

/* ---- host-only hooks (no GPU needed): the C++ host arithmetic under the Fiat-Shamir driver -------- */
/* op: 0 mul, 1 add, 2 sub, 3 inverse, 4 square, 5 sqrt (0 when non-residue).  Returns -2 on a
 * non-canonical input. */
int mina_b200_host_field_op(int field, int op, uint32_t n, const uint8_t *a32, const uint8_t *b32, uint8_t *out32);
/* SRS::create restated on the host: g[first..first+count) and optionally h, canonical affine. */
int mina_b200_host_srs_derive(int curve, uint32_t first, uint32_t count, uint8_t *out64, uint8_t *h64);
/* Derive both SRS on the host and store them under cache_dir (what mina_b200_init loads). */
int mina_b200_host_build_srs_cache(const char *cache_dir);
int mina_b200_host_blake2b512(const uint8_t *data, size_t len, uint8_t out[64]);

#ifdef __cplusplus
}
#endif
#endif /* MINA_B200_H */
