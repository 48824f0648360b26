//! Thin Rust face of the B200 account-inclusion verifier: same symbol and signature as
//! `AL/operator/mina_account/lib/src/lib.rs:16-22`; body in `libmina_b200.so`.  UNCOMPILED here.

pub const MAX_PROOF_SIZE: usize = 16 * 1024; // mina_account/lib/src/lib.rs:13
pub const MAX_PUB_INPUT_SIZE: usize = 6 * 1024; // mina_account/lib/src/lib.rs:14

mod sys {
    extern "C" {
        #[link_name = "verify_account_inclusion_ffi"]
        pub fn b200_verify_account_inclusion_ffi(proof: *const u8, proof_len: usize, pub_input: *const u8, pub_input_len: usize) -> bool;
    }
}

/// `zk_utils/mod.rs:108` calls this exactly as it calls the reference's rlib function.
pub fn verify_account_inclusion_ffi(
    proof_buffer: &[u8; MAX_PROOF_SIZE],
    proof_len: usize,
    pub_input_buffer: &[u8; MAX_PUB_INPUT_SIZE],
    pub_input_len: usize,
) -> bool {
    unsafe { sys::b200_verify_account_inclusion_ffi(proof_buffer.as_ptr(), proof_len, pub_input_buffer.as_ptr(), pub_input_len) }
}
