//! Thin Rust face of the B200 verifier: same symbol, same signature, same semantics as
//! `AL/operator/mina/lib/src/lib.rs:41-47`; the body lives in `libmina_b200.so`
//! (mina_bridge_b200/csrc/verifier.cu).  UNCOMPILED in this repository (no Rust toolchain in the image).

pub const MAX_PROOF_SIZE: usize = 48 * 1024; // lib.rs:38
pub const MAX_PUB_INPUT_SIZE: usize = 6 * 1024; // lib.rs:39

mod sys {
    extern "C" {
        // include/mina_verifier.h -- exported by libmina_b200.so under the reference's own name
        #[link_name = "verify_mina_state_ffi"]
        pub fn b200_verify_mina_state_ffi(proof: *const u8, proof_len: usize, pub_input: *const u8, pub_input_len: usize) -> bool;
        pub fn verify_mina_state_batch_ffi(
            n: usize,
            proofs: *const *const u8,
            proof_lens: *const usize,
            pub_inputs: *const *const u8,
            pub_input_lens: *const usize,
            accept_out: *mut u8,
        ) -> i32;
        pub fn mina_verifier_init(data_dir: *const std::os::raw::c_char, device: i32) -> i32;
    }
}

/// What the batcher calls today (`zk_utils/mod.rs:85`): identical signature to the reference's rlib function.
/// When this crate is built as `lib` the batcher links it; the cdylib/staticlib builds re-export the C symbol
/// from libmina_b200.so directly (no Rust definition with the same `#[no_mangle]` name is emitted, to avoid a
/// duplicate symbol at link time).
pub fn verify_mina_state_ffi(
    proof_buffer: &[u8; MAX_PROOF_SIZE],
    proof_len: usize,
    pub_input_buffer: &[u8; MAX_PUB_INPUT_SIZE],
    pub_input_len: usize,
) -> bool {
    unsafe { sys::b200_verify_mina_state_ffi(proof_buffer.as_ptr(), proof_len, pub_input_buffer.as_ptr(), pub_input_len) }
}

/// Additive: verify all Mina items of an Aligned batch in one call.
pub fn verify_mina_state_batch(items: &[(&[u8], &[u8])]) -> Vec<bool> {
    let proofs: Vec<*const u8> = items.iter().map(|(p, _)| p.as_ptr()).collect();
    let proof_lens: Vec<usize> = items.iter().map(|(p, _)| p.len()).collect();
    let pubs: Vec<*const u8> = items.iter().map(|(_, q)| q.as_ptr()).collect();
    let pub_lens: Vec<usize> = items.iter().map(|(_, q)| q.len()).collect();
    let mut out = vec![0u8; items.len()];
    let rc = unsafe {
        sys::verify_mina_state_batch_ffi(items.len(), proofs.as_ptr(), proof_lens.as_ptr(), pubs.as_ptr(), pub_lens.as_ptr(), out.as_mut_ptr())
    };
    if rc != 0 {
        return vec![false; items.len()];
    }
    out.into_iter().map(|b| b == 1).collect()
}

/// Optional: move the one-time cost (SRS upload, MSM tables, VK load) out of the first verification.
pub fn init(device: i32) -> bool {
    unsafe { sys::mina_verifier_init(std::ptr::null(), device) == 0 }
}
