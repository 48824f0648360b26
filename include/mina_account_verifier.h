/*
 * mina_account_verifier.h -- drop-in C ABI of the Mina proof-of-account (Merkle inclusion) verifier.
 *
 * Replaces the Rust cdylib `mina-account-verifier-ffi`:
 *   definition   AL/operator/mina_account/lib/src/lib.rs:16-78
 *   C header     AL/operator/mina_account/lib/mina_account_verifier.h:3-6 (lengths `unsigned int` there,
 *                `usize` in the definition; size_t here)
 *   callers      AL/operator/mina_account/mina_account.go:27-32 (cgo),
 *                AL/batcher/aligned-batcher/src/zk_utils/mod.rs:87-109 (rlib)
 *
 * Contract (lib.rs:23-52): fixed 16 KiB / 6 KiB zero-padded caller-owned arrays, first `len` bytes read,
 * oversize lengths / decode failures / failed checks return false, re-entrant, never unwinds.
 */
#ifndef MINA_ACCOUNT_VERIFIER_H
#define MINA_ACCOUNT_VERIFIER_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MINA_ACCOUNT_MAX_PROOF_SIZE (16 * 1024)    /* mina_account/lib/src/lib.rs:13 */
#define MINA_ACCOUNT_MAX_PUB_INPUT_SIZE (6 * 1024) /* mina_account/lib/src/lib.rs:14 */

bool verify_account_inclusion_ffi(const unsigned char *proof_buffer, size_t proof_len, const unsigned char *public_input_buffer,
                                  size_t public_input_len);

/* ADDITIVE batch twin (see mina_verifier.h). */
int verify_account_inclusion_batch_ffi(size_t n, const unsigned char *const *proofs, const size_t *proof_lens,
                                       const unsigned char *const *pub_inputs, const size_t *pub_input_lens, uint8_t *accept_out);

#ifdef __cplusplus
}
#endif
#endif /* MINA_ACCOUNT_VERIFIER_H */
