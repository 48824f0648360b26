/*
 * mina_verifier.h -- drop-in C ABI of the Mina proof-of-state verifier.
 *
 * Replaces the Rust cdylib `mina-state-verifier-ffi`:
 *   definition   AL/operator/mina/lib/src/lib.rs:41-113   (#[no_mangle] extern "C", lengths are `usize`)
 *   C header     AL/operator/mina/lib/mina_verifier.h:3-6 (declares the lengths as `unsigned int`;
 *                identical on x86-64 SysV because cgo zero-extends -- this header uses size_t, the
 *                width the Rust definition actually reads)
 *   callers      AL/operator/mina/mina.go:27-32 (cgo), AL/batcher/aligned-batcher/src/zk_utils/mod.rs:64-86 (rlib)
 *
 * Contract (lib.rs:48-94): the caller owns both buffers (fixed 48 KiB / 6 KiB zero-padded arrays);
 * only the first `len` bytes are read; `len` greater than the maximum, any decoding failure and any
 * failed check return false; nothing is returned by pointer; the call never unwinds.  It may be
 * called concurrently from any number of threads: concurrent callers are coalesced into GPU batches.
 */
#ifndef MINA_VERIFIER_H
#define MINA_VERIFIER_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MINA_STATE_MAX_PROOF_SIZE (48 * 1024)    /* lib.rs:38 */
#define MINA_STATE_MAX_PUB_INPUT_SIZE (6 * 1024) /* lib.rs:39 */

bool verify_mina_state_ffi(const unsigned char *proof_buffer, size_t proof_len, const unsigned char *pub_input_buffer,
                           size_t pub_input_len);

/* ADDITIVE (not in the reference): one call for a whole Aligned batch -- what the operator's
 * one-goroutine-per-proof fan-out (AL/operator/pkg/operator.go:448-465) becomes when the Mina items
 * of a batch are routed together.  accept_out[i] = 1/0 with exactly the per-proof semantics above.
 * Returns 0, or a negative value if the batch could not be processed at all (accept_out is then all 0). */
int verify_mina_state_batch_ffi(size_t n, const unsigned char *const *proofs, const size_t *proof_lens,
                                const unsigned char *const *pub_inputs, const size_t *pub_input_lens, uint8_t *accept_out);

/* ADDITIVE: explicit lifecycle.  The reference initialises lazily at first use (lazy_static, lib.rs:23-35);
 * so does this library (device 0, data next to the shared object, overridable with MINA_B200_DEVICE /
 * MINA_B200_DATA_DIR).  Calling init first moves the one-time cost (SRS upload, MSM tables, VK load)
 * out of the first verification.  `data_dir` may be NULL. */
int mina_verifier_init(const char *data_dir, int device);
void mina_verifier_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* MINA_VERIFIER_H */
