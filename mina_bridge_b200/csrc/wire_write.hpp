// bincode 1.3 encoders: the producer side of the four wire buffers.
//
// Replaces, for synthetic batch generation and proof mutation in the exact wire format, what the bridge
// CLI does with `bincode::serialize(&proof)` / `bincode::serialize(&pub_input)` before it submits to
// Aligned or dumps `--save-proof` files (core/src/aligned.rs:33-49, 60-69).  Field order and integer
// widths are the mirror image of wire.hpp's decoders (core/src/proof/state_proof.rs:10-41,
// core/src/proof/account_proof.rs:9-35, `SolSerialize` core/src/sol/serialization.rs:11-86).
//
// Parity: decode(fixture) -> encode reproduces each of the reference's five fixtures byte for byte
// (tests/test_boundary_cpu.py::test_writers_reproduce_the_fixtures).
#pragma once
#include "wire.hpp"

namespace pasta {
namespace wire {

class Writer {
   public:
    std::vector<uint8_t> out;
    void u8(uint8_t v) { out.push_back(v); }
    void u32(uint32_t v) {
        for (int i = 0; i < 4; i++) out.push_back((uint8_t)(v >> (8 * i)));
    }
    void u64(uint64_t v) {
        for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i)));
    }
    void boolean(bool v) { u8(v ? 1 : 0); }
    void raw32(const B32 &b) { out.insert(out.end(), b.begin(), b.end()); }
    void bigint(const B32 &b) {  // mina_p2p_messages BigInt: serialize_bytes of 32 bytes
        u64(32);
        raw32(b);
    }
    void bytes(const std::vector<uint8_t> &b) {
        u64(b.size());
        out.insert(out.end(), b.begin(), b.end());
    }
    void u128(const U128 &v) {
        u64(v.lo);
        u64(v.hi);
    }
    void point(const Point &p) {
        bigint(p.x);
        bigint(p.y);
    }
};

inline void write_signed_amount(Writer &w, const SignedAmount &a) {
    w.u64(a.magnitude);
    w.u32(a.sgn);
}
inline void write_pubkey(Writer &w, const PubKey &k) {
    w.bigint(k.x);
    w.boolean(k.is_odd);
}
inline void write_registers(Writer &w, const Registers &g) {
    w.bigint(g.first_pass_ledger);
    w.bigint(g.second_pass_ledger);
    w.bigint(g.pc_data);
    w.bigint(g.pc_state_init);
    w.bigint(g.pc_state_curr);
    const LocalState &ls = g.local_state;
    w.bigint(ls.stack_frame);
    w.bigint(ls.call_stack);
    w.bigint(ls.transaction_commitment);
    w.bigint(ls.full_transaction_commitment);
    write_signed_amount(w, ls.excess);
    write_signed_amount(w, ls.supply_increase);
    w.bigint(ls.ledger);
    w.boolean(ls.success);
    w.u32(ls.account_update_index);
    w.u64(ls.failure_status_tbl.size());
    for (auto &row : ls.failure_status_tbl) {
        w.u64(row.size());
        for (uint32_t tag : row) w.u32(tag);
    }
    w.boolean(ls.will_succeed);
}
inline void write_epoch_data(Writer &w, const EpochData &e) {
    w.bigint(e.ledger_hash);
    w.u64(e.ledger_total_currency);
    w.bigint(e.seed);
    w.bigint(e.start_checkpoint);
    w.bigint(e.lock_checkpoint);
    w.u32(e.epoch_length);
}
inline void write_protocol_state(Writer &w, const ProtocolState &st) {
    w.bigint(st.previous_state_hash);
    w.bigint(st.genesis_state_hash);
    const BlockchainState &b = st.blockchain_state;
    w.bigint(b.staged_ledger_hash);
    w.bytes(b.aux_hash);
    w.bytes(b.pending_coinbase_aux);
    w.bigint(b.pending_coinbase_hash);
    w.bigint(b.genesis_ledger_hash);
    write_registers(w, b.source);
    write_registers(w, b.target);
    w.bigint(b.connecting_ledger_left);
    w.bigint(b.connecting_ledger_right);
    write_signed_amount(w, b.supply_increase);
    w.bigint(b.fee_token_l);
    write_signed_amount(w, b.fee_excess_l);
    w.bigint(b.fee_token_r);
    write_signed_amount(w, b.fee_excess_r);
    w.u64(b.timestamp);
    w.bytes(b.body_reference);
    const ConsensusState &c = st.consensus_state;
    w.u32(c.blockchain_length);
    w.u32(c.epoch_count);
    w.u32(c.min_window_density);
    w.u64(c.sub_window_densities.size());
    for (uint32_t d : c.sub_window_densities) w.u32(d);
    w.bytes(c.last_vrf_output);
    w.u64(c.total_currency);
    w.u32(0);  // SinceHardFork
    w.u32(c.curr_global_slot);
    w.u32(c.slots_per_epoch);
    w.u32(0);  // SinceGenesis
    w.u32(c.global_slot_since_genesis);
    write_epoch_data(w, c.staking_epoch_data);
    write_epoch_data(w, c.next_epoch_data);
    w.boolean(c.has_ancestor_in_same_checkpoint_window);
    write_pubkey(w, c.block_stake_winner);
    write_pubkey(w, c.block_creator);
    write_pubkey(w, c.coinbase_receiver);
    w.boolean(c.supercharge_coinbase);
    const ProtocolConstants &k = st.constants;
    w.u32(k.k);
    w.u32(k.slots_per_epoch);
    w.u32(k.slots_per_sub_window);
    w.u32(k.grace_period_slots);
    w.u32(k.delta);
    w.u64(k.genesis_state_timestamp);
}
inline void write_bigint_vec(Writer &w, const std::vector<B32> &v) {
    w.u64(v.size());
    for (auto &x : v) w.bigint(x);
}
inline void write_pickles_proof(Writer &w, const PicklesProof &p) {
    w.u128(p.alpha);
    w.u128(p.beta);
    w.u128(p.gamma);
    w.u128(p.zeta);
    w.boolean(p.has_joint_combiner);
    if (p.has_joint_combiner) w.u128(p.joint_combiner);
    for (int i = 0; i < 8; i++) w.boolean(p.feature_flags[i]);
    for (int i = 0; i < 16; i++) w.u128(p.bulletproof_challenges[i]);
    w.u32(p.proofs_verified);
    w.u8(p.domain_log2);
    for (int i = 0; i < 4; i++) w.u64(p.sponge_digest_before_evaluations[i]);
    w.point(p.wrap_challenge_polynomial_commitment);
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < 15; i++) w.u128(p.wrap_old_bulletproof_challenges[k][i]);
    w.u64(p.step_challenge_polynomial_commitments.size());
    for (auto &pt : p.step_challenge_polynomial_commitments) w.point(pt);
    w.u64(p.step_old_bulletproof_challenges.size());
    for (auto &row : p.step_old_bulletproof_challenges)
        for (int j = 0; j < 16; j++) w.u128(row[j]);
    const PrevEvals &pe = p.prev_evals;
    w.bigint(pe.public_input[0]);
    w.bigint(pe.public_input[1]);
    for (int e = 0; e < N_EVALS; e++) {
        write_bigint_vec(w, pe.evals[e][0]);
        write_bigint_vec(w, pe.evals[e][1]);
    }
    for (int i = 0; i < 19; i++) w.u8(0);  // the 19 optional (lookup / range-check ...) evaluations: None
    w.bigint(pe.ft_eval1);
    const WireProof &q = p.proof;
    for (int i = 0; i < 15; i++) w.point(q.w_comm[i]);
    w.point(q.z_comm);
    for (int i = 0; i < 7; i++) w.point(q.t_comm[i]);
    for (int e = 0; e < N_EVALS; e++) {
        w.bigint(q.evals[e][0]);
        w.bigint(q.evals[e][1]);
    }
    w.bigint(q.ft_eval1);
    w.u64(q.lr.size());
    for (auto &pr : q.lr) {
        w.point(pr[0]);
        w.point(pr[1]);
    }
    w.bigint(q.z_1);
    w.bigint(q.z_2);
    w.point(q.delta);
    w.point(q.sg);
}

// MinaStateProof (core/src/proof/state_proof.rs:28-41).  A proof whose prev_evals carried optional
// entries cannot be re-encoded (the decoder drops them): returns false.
inline bool encode_state_proof(const StateProof &sp, std::vector<uint8_t> &out) {
    if (sp.candidate_tip_proof.prev_evals.has_optional) return false;
    Writer w;
    write_pickles_proof(w, sp.candidate_tip_proof);
    for (int i = 0; i < FRONTIER_LEN; i++) write_protocol_state(w, sp.candidate_chain_states[i]);
    write_protocol_state(w, sp.bridge_tip_state);
    out.swap(w.out);
    return true;
}
// MinaStatePubInputs (state_proof.rs:10-25): bool + 33 raw 32-byte field elements (SolSerialize)
inline void encode_state_pub(const StatePubInputs &pub, std::vector<uint8_t> &out) {
    Writer w;
    w.boolean(pub.is_state_proof_from_devnet);
    w.raw32(pub.bridge_tip_state_hash);
    for (int i = 0; i < FRONTIER_LEN; i++) w.raw32(pub.candidate_chain_state_hashes[i]);
    for (int i = 0; i < FRONTIER_LEN; i++) w.raw32(pub.candidate_chain_ledger_hashes[i]);
    out.swap(w.out);
}
inline void write_zkapp(Writer &w, const ZkappAccount &z) {
    for (int i = 0; i < 8; i++) w.bigint(z.app_state[i]);
    w.boolean(z.has_vk);
    if (z.has_vk) {
        w.u32(z.vk_max_proofs_verified);
        w.u32(z.vk_actual_wrap_domain_size);
        for (int i = 0; i < 28; i++) w.point(z.vk_wrap_index[i]);
    }
    w.u32(z.zkapp_version);
    for (int i = 0; i < 5; i++) w.bigint(z.action_state[i]);
    w.u32(z.last_action_slot);
    w.boolean(z.proved_state);
    w.bytes(z.zkapp_uri);
}
// MinaAccountProof (core/src/proof/account_proof.rs:28-35)
inline void encode_account_proof(const AccountProof &ap, std::vector<uint8_t> &out) {
    Writer w;
    w.u64(ap.merkle_path.size());
    for (auto &node : ap.merkle_path) {
        w.u32(node.tag);
        w.bigint(node.hash);  // o1_utils SerdeAs: byte string of the 32-byte ark-serialize encoding
    }
    const Account &a = ap.account;
    write_pubkey(w, a.public_key);
    w.bigint(a.token_id);
    w.bytes(a.token_symbol);
    w.u64(a.balance);
    w.u32(a.nonce);
    w.bigint(a.receipt_chain_hash);
    w.boolean(a.has_delegate);
    if (a.has_delegate) write_pubkey(w, a.delegate);
    w.bigint(a.voting_for);
    w.u32(a.timing.timed ? 1 : 0);
    if (a.timing.timed) {
        w.u64(a.timing.initial_minimum_balance);
        w.u32(a.timing.cliff_time);
        w.u64(a.timing.cliff_amount);
        w.u32(a.timing.vesting_period);
        w.u64(a.timing.vesting_increment);
    }
    for (int i = 0; i < 13; i++) {
        w.u32(a.permissions.auth[i]);
        if (i == 6) w.u32(a.permissions.set_vk_txn_version);
    }
    w.boolean(a.has_zkapp);
    if (a.has_zkapp) write_zkapp(w, a.zkapp);
    out.swap(w.out);
}
// MinaAccountPubInputs (account_proof.rs:17-25)
inline void encode_account_pub(const AccountPubInputs &pub, std::vector<uint8_t> &out) {
    Writer w;
    w.raw32(pub.ledger_hash);
    w.bytes(pub.encoded_account);
    out.swap(w.out);
}

}  // namespace wire
}  // namespace pasta
