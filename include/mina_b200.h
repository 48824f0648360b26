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
 * are declared in include/mina_verifier.h and include/mina_account_verifier.h with the reference's
 * signatures.  Everything below is ADDITIVE: lifecycle, per-stage reports, and the individual
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
/* Select `device`; load the two SRS from `data_dir` (derive them by hash-to-curve when the cache is
 * missing or does not match the digest compiled into the library), upload them, build the fixed-base
 * MSM tables, load the verification keys (<data_dir>/{devnet,mainnet}_vk.json) and, when present and
 * passing the reference's known-answer test, the Poseidon tables.  Mirrors the reference's lazy
 * statics MINA_SRS / *_VERIFIER_INDEX (AL/operator/mina/lib/src/lib.rs:23-35).  `data_dir` NULL =
 * $MINA_B200_DATA_DIR, else the `data` directory next to the shared object.  Idempotent for the same
 * device; an error for a different one.  Thread-safe. */
int mina_b200_init(int device, const char *data_dir);
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
/* Device-resident variant: d_scalars is a CUDA device pointer to nmsm*n*8 uint32 (canonical), d_out64
 * a device pointer to nmsm*16 uint32 (canonical affine).  Work is enqueued on `cuda_stream` (a
 * cudaStream_t passed as void*) and the call returns without synchronising; results and the engine's
 * scalar-range flag are valid once the stream has drained.  If `accumulate_ms` is non-NULL the call
 * DOES synchronise and stores the CUDA-event duration of the dominant kernel (bucket accumulation)
 * of the last chunk. */
int mina_b200_msm_srs_device(int curve, uint32_t nmsm, uint32_t n, const void *d_scalars, void *d_out64,
                             void *cuda_stream, float *accumulate_ms);
/* Shape of the IPA final-check MSM: sum_i s_i g[i] over the resident SRS prefix plus
 * sum_j t_j P_j over a few caller-supplied points (canonical affine, validated).  Either part may be empty. */
int mina_b200_msm_srs_plus(int curve, uint32_t n_srs, const uint8_t *scalars_srs32, uint32_t n_extra, const uint8_t *scalars_extra32,
                           const uint8_t *points_extra64, uint8_t *out64);
/* Fixed-base MSM over a caller-supplied base set that stays resident (BASELINE config 2: 2^20 Vesta
 * points): load once (builds the window table: ceil(255/c)+... x n x 64 B), then run any number of MSMs.
 * Points are validated (canonical, on curve).  window_bits 0 = 16. */
int mina_b200_fixed_base_load(int curve, uint32_t n, const uint8_t *points64, int window_bits);
int mina_b200_fixed_base_msm(int curve, uint32_t nmsm, const uint8_t *scalars32, uint8_t *out64);
int mina_b200_fixed_base_msm_device(int curve, uint32_t nmsm, const void *d_scalars, void *d_out64, void *cuda_stream,
                                    float *accumulate_ms);
/* Commitments of the Lagrange polynomials L_first .. L_{first+count-1} of the radix-2 domain of size 2^log_n over the
 * resident SRS (kimchi `SRS::add_lagrange_basis`, AL/operator/mina/lib/src/verifier_index.rs:204-208; the verifier's
 * public-input commitment is -sum_i pub_i L_i + h over the first `public` = 40 of them).  2^log_n <= SRS depth. */
int mina_b200_lagrange_commitments(int curve, uint32_t log_n, uint32_t first, uint32_t count, uint8_t *out64);
/* kimchi's public-input commitment (`public_comm` in kimchi's verifier; part of SURVEY row a8) for nproofs vectors of
 * n_pub public inputs (pub32: [nproofs][n_pub] canonical scalars): out = -sum_i pub_i L_i + h.  The Lagrange commitments
 * are computed once per (log_n, n_pub) and kept as a small fixed-base table. */
int mina_b200_public_commitments(int curve, uint32_t log_n, uint32_t n_pub, uint32_t nproofs, const uint8_t *pub32, uint8_t *out64);
/* MSM engine tuning (takes effect at the next init / table rebuild): window bits and running-sum
 * chunk.  Returns 0 on success. */
int mina_b200_msm_configure(int curve, int window_bits, int precompute, int leaf);

/* ---- verifier stages ------------------------------------------------------------------------------ */
/* One bit per stage of the reference's check.  A proof is accepted iff every stage of its kind is in
 * `passed`; `unavailable` lists stages this build cannot run yet (they force a reject). */
#define MINA_B200_STAGE_LENGTHS (1u << 0)           /* len <= MAX (lib.rs:48-56) */
#define MINA_B200_STAGE_DECODE_PROOF (1u << 1)      /* bincode (lib.rs:58-64) */
#define MINA_B200_STAGE_DECODE_PUB (1u << 2)        /* bincode (lib.rs:65-71) */
#define MINA_B200_STAGE_PUB_STRUCTURE (1u << 3)     /* ledger-hash comparisons + to_fp (lib.rs:163-186,202-209) */
#define MINA_B200_STAGE_PUB_HASHES (1u << 4)        /* 17 Poseidon state hashes (lib.rs:128-160,188) */
#define MINA_B200_STAGE_CONSENSUS (1u << 5)         /* select_secure_chain == Candidate (lib.rs:83-94) */
#define MINA_B200_STAGE_ACCUMULATOR (1u << 6)       /* accumulator_check: Vesta MSM 2^16 (verify_block) */
#define MINA_B200_STAGE_STEP_ACCUMULATORS (1u << 7) /* the wrap proof's two previous-challenge accumulators (Pallas 2^15) */
#define MINA_B200_STAGE_KIMCHI (1u << 8)            /* kimchi to_batch + IPA final check (verify_block) */
#define MINA_B200_STAGE_ACCOUNT_ABI (1u << 9)       /* Solidity ABI re-encode + compare (mina_account lib.rs:54-66) */
#define MINA_B200_STAGE_ACCOUNT_LEAF (1u << 10)     /* Account::hash (mina_account lib.rs:70) */
#define MINA_B200_STAGE_MERKLE (1u << 11)           /* verify_merkle_proof (merkle_verifier.rs:9-35) */
#define MINA_B200_STAGE_INTERNAL_ERROR (1u << 31)   /* device / allocation failure: rejected */
typedef struct {
    uint32_t passed, failed, unavailable;
} mina_b200_stage_report;

#define MINA_B200_MODE_PER_PROOF 0 /* one MSM per accumulator, like the reference */
#define MINA_B200_MODE_RLC 1       /* random linear combination over the batch + bisection (default) */

/* Stage report of the last verify_mina_state_ffi / verify_account_inclusion_ffi call on this thread. */
void mina_b200_last_stages(mina_b200_stage_report *out);
/* The batch verifier with per-proof reports.  reports / accept_out may be NULL. */
int mina_b200_verify_state_stages(size_t n, const unsigned char *const *proofs, const size_t *proof_lens,
                                  const unsigned char *const *pub_inputs, const size_t *pub_input_lens, int mode,
                                  mina_b200_stage_report *reports, uint8_t *accept_out);
int mina_b200_verify_account_stages(size_t n, const unsigned char *const *proofs, const size_t *proof_lens,
                                    const unsigned char *const *pub_inputs, const size_t *pub_input_lens,
                                    mina_b200_stage_report *reports, uint8_t *accept_out);
/* accumulator_check from raw proof bytes: ok3 = {wrap/Vesta, step/Pallas #0, #1}. */
int mina_b200_accumulator_check(const unsigned char *proof, size_t proof_len, uint8_t ok3[3]);
int mina_b200_accumulator_check_batch(size_t n, const unsigned char *const *proofs, const size_t *proof_lens, int mode, uint8_t *ok3);

/* Device-resident variant (bench `value` leg): d_pre16 = m*k 16-byte prechallenges and d_pts64 = m
 * canonical affine points already in HBM (k = 16 on Vesta, 15 on Pallas); ok_host gets m bytes.
 * stats (may be NULL): CUDA-event time of the two dominant kernels summed over every launch of the call,
 * and how much work those launches covered (for the roofline arithmetic in bench.py). */
typedef struct {
    float accumulate_ms;      /* k_accumulate, all chunks of all MSM batches */
    float combine_ms;         /* k_bpoly_combine, all levels (RLC mode) */
    uint64_t msm_points;      /* points summed by k_accumulate launches = sum over MSMs of n */
    uint64_t msm_count;       /* MSMs over the resident SRS */
    uint64_t combine_proofs;  /* proofs read by k_bpoly_combine launches (16 KiB of tables each) */
    uint64_t combine_vectors; /* 2^k-element vectors written by k_bpoly_combine launches */
} mina_b200_kernel_stats;
int mina_b200_accumulators_device(int curve, uint32_t m, const void *d_pre16, const void *d_pts64, int mode, uint8_t *ok_host,
                                  mina_b200_kernel_stats *stats);

/* Both accumulator families of m state proofs at once, driven concurrently on two streams: d_pre_wrap =
 * m x 16 prechallenges, d_pts_wrap = m points (Vesta); d_pre_step = 2m x 15 prechallenges, d_pts_step = 2m
 * points (Pallas).  ok3 = 3 bytes per proof; stats2 (may be NULL) = {wrap, step}. */
int mina_b200_state_accumulators_device(uint32_t m, const void *d_pre_wrap, const void *d_pts_wrap, const void *d_pre_step,
                                        const void *d_pts_step, int mode, uint8_t *ok3, mina_b200_kernel_stats *stats2);

/* ---- K4 / K2 / K5: IPA scalar helpers (host buffers, canonical 32-byte field elements) -------------- */
/* ScalarChallenge::to_field for n 16-byte prechallenges landing in `field` (endo = that field's endo_r). */
int mina_b200_endo_to_field(int field, uint32_t n, const uint8_t *pre16, uint8_t *out32);
/* b_poly_coefficients for nproofs x k challenges (8 <= k <= 16): out is nproofs * 2^k elements. */
int mina_b200_bpoly_coeffs(int field, uint32_t nproofs, int k, const uint8_t *chals32, uint8_t *out32);
/* sum_j r_j * b_poly_coefficients(chals_j): out is 2^k elements. */
int mina_b200_bpoly_combine(int field, uint32_t nproofs, int k, const uint8_t *chals32, const uint8_t *r32, uint8_t *out32);
/* b_poly(chals_j, x_{j,t}) for npts points per proof. */
int mina_b200_bpoly_eval(int field, uint32_t nproofs, uint32_t npts, int k, const uint8_t *chals32, const uint8_t *x32, uint8_t *out32);

/* kimchi combined_inner_product: sum_i polyscale^i sum_j evalscale^j evals[p][i][j]; scales32 = nproofs x
 * (polyscale, evalscale). */
int mina_b200_combined_inner_product(int field, uint32_t nproofs, uint32_t npolys, uint32_t npts, const uint8_t *evals32,
                                      const uint8_t *scales32, uint8_t *out32);

/* ---- a9: batched IPA final check (poly-commitment SRS::verify) -------------------------------------------- */
/* n openings over `curve`, all with the same number of rounds k (8 <= k <= 16, 2^k <= the resident SRS depth: the
 * bases are g[0..2^k) and h), commitments and evaluation points.  Host buffers, canonical little-endian:
 *   sponge_state96 [n][3] base-field elements + the sponge mode shared by the batch (0 = Absorbed(count), 1 =
 *   Squeezed(count)): the Fq-sponge exactly as kimchi hands it over (`fq_sponge_before_evaluations`);
 *   cip32, polyscale32, evalscale32, z1_32, z2_32 [n] scalars; eval_points32 [n][n_points]; delta64, sg64 [n];
 *   commitments64 [n][n_comm]; lr64 [n][k][2] (L then R).
 * poseidon_table: 174 x 32 bytes for the curve's BASE field (see the K3 section; table-driven, unpinned).
 * ok[i] = 1 iff opening i verifies.  Openings are batched with per-opening 128-bit randomisers and the group
 * testing of the accumulator checks; a non-canonical input or an off-curve point rejects that opening only. */
typedef struct {
    uint32_t n, rounds, n_comm, n_points;
    uint32_t sponge_mode, sponge_count;
    const uint8_t *sponge_state96;
    const uint8_t *cip32, *polyscale32, *evalscale32, *z1_32, *z2_32;
    const uint8_t *eval_points32;
    const uint8_t *delta64, *sg64;
    const uint8_t *commitments64;
    const uint8_t *lr64;
} mina_b200_ipa_batch;
int mina_b200_ipa_verify(int curve, const uint8_t *poseidon_table, const mina_b200_ipa_batch *batch, uint8_t *ok);

/* ---- K3: Poseidon (table-driven; see csrc/poseidon.hpp for the parity status of the constants) ------- */
/* table = 174 x 32 bytes (MDS row-major, then 55 x 3 round constants).  states: n x 96 bytes, in place. */
int mina_b200_poseidon_permute(int field, const uint8_t *table, uint32_t n, uint8_t *states96);
/* Merkle fold (merkle_verifier.rs:9-35) of nproofs paths of up to max_depth nodes: tags[p*max_depth+d]
 * (0 = Left, 1 = Right), siblings32 likewise, Fp only.  ok[p] = folded root == roots32[p]. */
int mina_b200_merkle_fold(const uint8_t *table, uint32_t nproofs, uint32_t max_depth, const uint32_t *depths, const uint8_t *tags,
                          const uint8_t *siblings32, const uint8_t *leaves32, const uint8_t *roots32, uint8_t *ok, uint8_t *folded32);
/* 1 iff a table from <data_dir>/poseidon_fp_kimchi.bin passed the reference's known-answer test. */
int mina_b200_poseidon_trusted(void);

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
/* Loader for the committed srs/{vesta,pallas}.srs files (MessagePack + compressed points):
 * the first `count` points as canonical affine, and `h`.  -1 on a malformed file (see last_error). */
int mina_b200_host_srs_load_file(int curve, const char *path, uint32_t count, uint8_t *out64, uint8_t *h64);
int mina_b200_host_blake2b512(const uint8_t *data, size_t len, uint8_t out[64]);

/* Wire decoders (csrc/wire.hpp).  kind: 0 state proof, 1 state pub, 2 account proof, 3 account pub.
 * Returns 0 and fills `summary` (see the field list in csrc/abi_host.cu), or -1 on a decode error. */
typedef struct {
    uint64_t consumed;       /* bytes read */
    uint64_t proof_end;      /* state proof: end of the Pickles proof (13 849 in the fixture) */
    uint32_t n_step_comms, n_lr, merkle_depth, is_devnet;
    uint32_t blockchain_length[17], curr_global_slot[17], epoch_count[17], min_window_density[17];
    uint64_t state_begin[17], state_end[17];
    uint8_t previous_state_hash[17][32], first_pass_ledger[17][32];
    uint8_t wrap_sg[64], step_sg[2][64];
    uint8_t hash0[32];       /* state pub: bridge tip hash; account pub: ledger hash */
    uint64_t encoded_account_len, balance;
    uint32_t nonce, has_zkapp;
} mina_b200_wire_summary;
int mina_b200_host_decode(int kind, const uint8_t *data, size_t len, mina_b200_wire_summary *summary);
/* select_secure_chain (consensus_state.rs:22-37) on two bincode-encoded protocol states.
 * *result: 0 = Bridge, 1 = Candidate.  Returns 0, -1 decode error, -2 constants differ (the reference's
 * Err), -3 the tie needs Poseidon state hashes. */
int mina_b200_host_select_secure_chain(const uint8_t *candidate, size_t candidate_len, const uint8_t *tip, size_t tip_len, int *result);
/* Verification-key loader (verifier_index.rs:109-276).  out32: 28 x 64 bytes of commitments (canonical),
 * then shifts[7], group_gen, zk_w3, zkpm[4], endo as 32-byte canonical Fq = 28*64 + 14*32 bytes;
 * meta: log_size_of_group, max_poly_size, public, prev_challenges.  -1 on error (see last_error). */
int mina_b200_host_vk_load(const char *path, uint8_t *out, uint32_t meta[4]);
/* Host Poseidon sponge: hash_with_kimchi(prefix, xs[0..n)) with the given table (Fp). */
int mina_b200_host_hash_with_kimchi(const uint8_t *table, const char *prefix, const uint8_t *xs32, uint32_t n, uint8_t out32[32]);
int mina_b200_host_poseidon_permute(int field, const uint8_t *table, uint32_t n, uint8_t *states96);
/* The planner of the batched checks' group testing (csrc/group_testing.hpp) against a simulated device: bad[i] != 0
 * marks the bad items of a batch of m.  ok_out[i] receives the bit the batched check would report; *levels and *msms the
 * number of sequential levels and of MSM evaluations over the resident SRS it would cost.  -1 on an internal
 * inconsistency (see last_error). */
int mina_b200_host_group_testing_sim(uint32_t m, const uint8_t *bad, uint8_t *ok_out, uint32_t *levels, uint32_t *msms);
/* Wire encoders (csrc/wire_write.hpp; the producer side, core/src/aligned.rs:33-49): decode `data` as `kind`
 * (same ids as mina_b200_host_decode) and encode it again.  *out_len in = capacity, out = bytes written.
 * Returns 0, -1 decode error, -2 not re-encodable / buffer too small. */
int mina_b200_host_reencode(int kind, const uint8_t *data, size_t len, uint8_t *out, size_t *out_len);
/* Solidity ABI encoding of the account inside an account proof (core/src/sol/account.rs:25-314 +
 * `abi_encode()`, mina_account lib.rs:54-62): what the verifier compares with the public input's
 * `encoded_account`.  Returns 0, -1 decode error, -2 conversion failure (non-UTF-8 token symbol) / buffer too small. */
int mina_b200_host_account_abi_encode(const uint8_t *account_proof, size_t len, uint8_t *out, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif /* MINA_B200_H */
