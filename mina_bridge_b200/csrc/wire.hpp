// bincode 1.3 decoders for the four buffers the two FFI entry points receive.
//
// Host-side replacement for `bincode::deserialize::<MinaStateProof / MinaStatePubInputs /
// MinaAccountProof / MinaAccountPubInputs>` at AL/operator/mina/lib/src/lib.rs:58-71 and
// AL/operator/mina_account/lib/src/lib.rs:33-52.  Wire types: core/src/proof/state_proof.rs:10-41,
// core/src/proof/account_proof.rs:9-35, core/src/sol/serialization.rs:11-86, and the
// mina-p2p-messages 0.6.4 types they embed (layout: SURVEY Appendix A, pinned by byte-exact
// consumption of the reference's five fixtures).
//
// bincode defaults: little-endian fixed-width integers, u64 sequence lengths, u32 enum tags, one-byte
// bool / Option tags (any value other than 0/1 is an error), trailing bytes allowed.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace pasta {
namespace wire {

static constexpr size_t MAX_STATE_PROOF_SIZE = 48 * 1024;    // AL/operator/mina/lib/src/lib.rs:38
static constexpr size_t MAX_ACCOUNT_PROOF_SIZE = 16 * 1024;  // AL/operator/mina_account/lib/src/lib.rs:13
static constexpr size_t MAX_PUB_INPUT_SIZE = 6 * 1024;       // lib.rs:39, mina_account lib.rs:14
static constexpr int FRONTIER_LEN = 16;                      // core/src/utils/constants.rs:31

using B32 = std::array<uint8_t, 32>;  // a 256-bit little-endian integer exactly as it sits on the wire

struct U128 {
    uint64_t lo = 0, hi = 0;  // limb pair: two i64 little-endian, low limb first
};
struct Point {
    B32 x, y;
};
struct SignedAmount {
    uint64_t magnitude = 0;
    uint32_t sgn = 0;  // 0 = Pos, 1 = Neg
};
struct PubKey {
    B32 x;
    bool is_odd = false;
};

struct LocalState {
    B32 stack_frame, call_stack, transaction_commitment, full_transaction_commitment;
    SignedAmount excess, supply_increase;
    B32 ledger;
    bool success = false;
    uint32_t account_update_index = 0;
    std::vector<std::vector<uint32_t>> failure_status_tbl;
    bool will_succeed = false;
};
struct Registers {
    B32 first_pass_ledger, second_pass_ledger;
    B32 pc_data, pc_state_init, pc_state_curr;  // pending-coinbase stack
    LocalState local_state;
};
struct EpochData {
    B32 ledger_hash;
    uint64_t ledger_total_currency = 0;
    B32 seed, start_checkpoint, lock_checkpoint;
    uint32_t epoch_length = 0;
};
struct ConsensusState {
    uint32_t blockchain_length = 0, epoch_count = 0, min_window_density = 0;
    std::vector<uint32_t> sub_window_densities;
    std::vector<uint8_t> last_vrf_output;
    uint64_t total_currency = 0;
    uint32_t curr_global_slot = 0, slots_per_epoch = 0;  // curr_global_slot_since_hard_fork
    uint32_t global_slot_since_genesis = 0;
    EpochData staking_epoch_data, next_epoch_data;
    bool has_ancestor_in_same_checkpoint_window = false;
    PubKey block_stake_winner, block_creator, coinbase_receiver;
    bool supercharge_coinbase = false;
};
struct ProtocolConstants {
    uint32_t k = 0, slots_per_epoch = 0, slots_per_sub_window = 0, grace_period_slots = 0, delta = 0;
    uint64_t genesis_state_timestamp = 0;
    bool operator==(const ProtocolConstants &o) const {
        return k == o.k && slots_per_epoch == o.slots_per_epoch && slots_per_sub_window == o.slots_per_sub_window &&
               grace_period_slots == o.grace_period_slots && delta == o.delta &&
               genesis_state_timestamp == o.genesis_state_timestamp;
    }
};
struct BlockchainState {
    B32 staged_ledger_hash;
    std::vector<uint8_t> aux_hash, pending_coinbase_aux;
    B32 pending_coinbase_hash;
    B32 genesis_ledger_hash;
    Registers source, target;
    B32 connecting_ledger_left, connecting_ledger_right;
    SignedAmount supply_increase;
    B32 fee_token_l, fee_token_r;
    SignedAmount fee_excess_l, fee_excess_r;
    uint64_t timestamp = 0;
    std::vector<uint8_t> body_reference;
};
struct ProtocolState {
    B32 previous_state_hash;
    B32 genesis_state_hash;
    BlockchainState blockchain_state;
    ConsensusState consensus_state;
    ProtocolConstants constants;
    size_t wire_begin = 0, wire_end = 0;  // byte span inside the proof buffer
};

static constexpr int N_EVALS = 43;  // w[15], coefficients[15], z, s[6], 6 selectors (Appendix A.1)
struct PrevEvals {
    B32 public_input[2];
    std::vector<B32> evals[N_EVALS][2];  // (zeta, zeta*omega), one chunk each for this circuit
    bool has_optional = false;           // any of the 19 optional (lookup / range-check ...) entries present
    B32 ft_eval1;
};
struct WireProof {
    Point w_comm[15], z_comm, t_comm[7];
    B32 evals[N_EVALS][2];
    B32 ft_eval1;
    std::vector<std::array<Point, 2>> lr;
    B32 z_1, z_2;
    Point delta, sg;
};
struct PicklesProof {
    U128 alpha, beta, gamma, zeta;
    bool has_joint_combiner = false;
    U128 joint_combiner;
    bool feature_flags[8] = {false};
    U128 bulletproof_challenges[16];  // step ("Tick") side -> Fp via endo
    uint32_t proofs_verified = 0;
    uint8_t domain_log2 = 0;
    uint64_t sponge_digest_before_evaluations[4] = {0};
    Point wrap_challenge_polynomial_commitment;  // Vesta point: accumulator of the step proof
    U128 wrap_old_bulletproof_challenges[2][15];  // wrap ("Tock") side -> Fq via endo
    std::vector<Point> step_challenge_polynomial_commitments;            // Pallas points
    std::vector<std::array<U128, 16>> step_old_bulletproof_challenges;
    PrevEvals prev_evals;
    WireProof proof;
    size_t wire_end = 0;
};
struct StateProof {
    PicklesProof candidate_tip_proof;
    ProtocolState candidate_chain_states[FRONTIER_LEN];
    ProtocolState bridge_tip_state;
};
struct StatePubInputs {
    bool is_state_proof_from_devnet = false;
    B32 bridge_tip_state_hash;
    B32 candidate_chain_state_hashes[FRONTIER_LEN];
    B32 candidate_chain_ledger_hashes[FRONTIER_LEN];
};

struct MerkleNode {
    uint32_t tag = 0;  // 0 = Left(sibling on the right), 1 = Right(sibling on the left)
    B32 hash;
};
struct Timing {
    bool timed = false;
    uint64_t initial_minimum_balance = 0, cliff_amount = 0, vesting_increment = 0;
    uint32_t cliff_time = 0, vesting_period = 0;
};
struct Permissions {
    uint32_t auth[13] = {0};  // edit_state, access, send, receive, set_delegate, set_permissions,
                              // set_verification_key, set_zkapp_uri, edit_action_state, set_token_symbol,
                              // increment_nonce, set_voting_for, set_timing
    uint32_t set_vk_txn_version = 0;
};
struct ZkappAccount {
    B32 app_state[8];
    bool has_vk = false;
    uint32_t vk_max_proofs_verified = 0, vk_actual_wrap_domain_size = 0;
    Point vk_wrap_index[28];
    uint32_t zkapp_version = 0;
    B32 action_state[5];
    uint32_t last_action_slot = 0;
    bool proved_state = false;
    std::vector<uint8_t> zkapp_uri;
};
struct Account {
    PubKey public_key;
    B32 token_id;
    std::vector<uint8_t> token_symbol;
    uint64_t balance = 0;
    uint32_t nonce = 0;
    B32 receipt_chain_hash;
    bool has_delegate = false;
    PubKey delegate;
    B32 voting_for;
    Timing timing;
    Permissions permissions;
    bool has_zkapp = false;
    ZkappAccount zkapp;
};
struct AccountProof {
    std::vector<MerkleNode> merkle_path;
    Account account;
};
struct AccountPubInputs {
    B32 ledger_hash;
    std::vector<uint8_t> encoded_account;
};

// ---- reader --------------------------------------------------------------------------------------
class Reader {
   public:
    Reader(const uint8_t *p, size_t n) : p_(p), n_(n) {}
    bool ok() const { return err_.empty(); }
    const std::string &error() const { return err_; }
    size_t offset() const { return o_; }
    size_t remaining() const { return n_ - o_; }
    void fail(const char *what) {
        if (err_.empty()) err_ = std::string(what) + " at byte " + std::to_string(o_);
    }
    const uint8_t *take(size_t k) {
        if (!ok()) return nullptr;
        if (k > n_ - o_) {
            fail("unexpected end of input");
            return nullptr;
        }
        const uint8_t *r = p_ + o_;
        o_ += k;
        return r;
    }
    uint8_t u8() {
        const uint8_t *b = take(1);
        return b ? b[0] : 0;
    }
    uint32_t u32() {
        const uint8_t *b = take(4);
        uint32_t v = 0;
        if (b) std::memcpy(&v, b, 4);
        return v;
    }
    uint64_t u64() {
        const uint8_t *b = take(8);
        uint64_t v = 0;
        if (b) std::memcpy(&v, b, 8);
        return v;
    }
    bool boolean() {
        uint8_t v = u8();
        if (v > 1) fail("invalid bool encoding");
        return v == 1;
    }
    bool option() {
        uint8_t v = u8();
        if (v > 1) fail("invalid Option tag");
        return v == 1;
    }
    uint32_t variant(uint32_t count) {
        uint32_t v = u32();
        if (ok() && v >= count) fail("enum variant out of range");
        return v;
    }
    // sequence length, bounded by what the remaining bytes could possibly hold
    size_t len(size_t min_item_bytes) {
        uint64_t v = u64();
        if (!ok()) return 0;
        if (min_item_bytes && v > remaining() / min_item_bytes) {
            fail("sequence length exceeds input");
            return 0;
        }
        return (size_t)v;
    }
    // mina_p2p_messages::bigint::BigInt: serialize_bytes of exactly 32 bytes
    void bigint(B32 &out) {
        uint64_t l = u64();
        if (ok() && l != 32) fail("BigInt length is not 32");
        raw32(out);
    }
    void raw32(B32 &out) {
        const uint8_t *b = take(32);
        if (b)
            std::memcpy(out.data(), b, 32);
        else
            out.fill(0);
    }
    void bytes(std::vector<uint8_t> &out) {
        size_t l = len(1);
        const uint8_t *b = take(l);
        out.clear();
        if (b) out.assign(b, b + l);
    }
    void u128(U128 &out) {
        out.lo = u64();
        out.hi = u64();
    }
    void point(Point &pt) {
        bigint(pt.x);
        bigint(pt.y);
    }

   private:
    const uint8_t *p_;
    size_t n_, o_ = 0;
    std::string err_;
};

inline void read_signed_amount(Reader &r, SignedAmount &a) {
    a.magnitude = r.u64();
    a.sgn = r.variant(2);
}
inline void read_pubkey(Reader &r, PubKey &k) {
    r.bigint(k.x);
    k.is_odd = r.boolean();
}
inline void read_registers(Reader &r, Registers &g) {
    r.bigint(g.first_pass_ledger);
    r.bigint(g.second_pass_ledger);
    r.bigint(g.pc_data);
    r.bigint(g.pc_state_init);
    r.bigint(g.pc_state_curr);
    LocalState &ls = g.local_state;
    r.bigint(ls.stack_frame);
    r.bigint(ls.call_stack);
    r.bigint(ls.transaction_commitment);
    r.bigint(ls.full_transaction_commitment);
    read_signed_amount(r, ls.excess);
    read_signed_amount(r, ls.supply_increase);
    r.bigint(ls.ledger);
    ls.success = r.boolean();
    ls.account_update_index = r.u32();
    size_t rows = r.len(8);
    ls.failure_status_tbl.clear();
    for (size_t i = 0; i < rows && r.ok(); i++) {
        size_t cols = r.len(4);
        std::vector<uint32_t> row;
        for (size_t j = 0; j < cols && r.ok(); j++) row.push_back(r.u32());  // failure enum: tag only decoded
        ls.failure_status_tbl.push_back(std::move(row));
    }
    ls.will_succeed = r.boolean();
}
inline void read_epoch_data(Reader &r, EpochData &e) {
    r.bigint(e.ledger_hash);
    e.ledger_total_currency = r.u64();
    r.bigint(e.seed);
    r.bigint(e.start_checkpoint);
    r.bigint(e.lock_checkpoint);
    e.epoch_length = r.u32();
}

// MinaStateProtocolStateValueStableV2
inline void read_protocol_state(Reader &r, ProtocolState &st) {
    st.wire_begin = r.offset();
    r.bigint(st.previous_state_hash);
    r.bigint(st.genesis_state_hash);
    BlockchainState &b = st.blockchain_state;
    r.bigint(b.staged_ledger_hash);
    r.bytes(b.aux_hash);
    r.bytes(b.pending_coinbase_aux);
    r.bigint(b.pending_coinbase_hash);
    r.bigint(b.genesis_ledger_hash);
    read_registers(r, b.source);
    read_registers(r, b.target);
    r.bigint(b.connecting_ledger_left);
    r.bigint(b.connecting_ledger_right);
    read_signed_amount(r, b.supply_increase);
    r.bigint(b.fee_token_l);
    read_signed_amount(r, b.fee_excess_l);
    r.bigint(b.fee_token_r);
    read_signed_amount(r, b.fee_excess_r);
    b.timestamp = r.u64();
    r.bytes(b.body_reference);
    ConsensusState &c = st.consensus_state;
    c.blockchain_length = r.u32();
    c.epoch_count = r.u32();
    c.min_window_density = r.u32();
    size_t nsub = r.len(4);
    c.sub_window_densities.clear();
    for (size_t i = 0; i < nsub && r.ok(); i++) c.sub_window_densities.push_back(r.u32());
    r.bytes(c.last_vrf_output);
    c.total_currency = r.u64();
    r.variant(1);  // MinaNumbersGlobalSlotSinceHardForkMStableV1::SinceHardFork
    c.curr_global_slot = r.u32();
    c.slots_per_epoch = r.u32();
    r.variant(1);  // ...SinceGenesis
    c.global_slot_since_genesis = r.u32();
    read_epoch_data(r, c.staking_epoch_data);
    read_epoch_data(r, c.next_epoch_data);
    c.has_ancestor_in_same_checkpoint_window = r.boolean();
    read_pubkey(r, c.block_stake_winner);
    read_pubkey(r, c.block_creator);
    read_pubkey(r, c.coinbase_receiver);
    c.supercharge_coinbase = r.boolean();
    ProtocolConstants &k = st.constants;
    k.k = r.u32();
    k.slots_per_epoch = r.u32();
    k.slots_per_sub_window = r.u32();
    k.grace_period_slots = r.u32();
    k.delta = r.u32();
    k.genesis_state_timestamp = r.u64();
    st.wire_end = r.offset();
}

inline void read_bigint_vec(Reader &r, std::vector<B32> &v) {
    size_t n = r.len(40);
    v.clear();
    for (size_t i = 0; i < n && r.ok(); i++) {
        B32 x;
        r.bigint(x);
        v.push_back(x);
    }
}

// MinaBaseProofStableV2 (PicklesProofProofsVerified2ReprStableV2)
inline void read_pickles_proof(Reader &r, PicklesProof &p) {
    r.u128(p.alpha);
    r.u128(p.beta);
    r.u128(p.gamma);
    r.u128(p.zeta);
    p.has_joint_combiner = r.option();
    if (p.has_joint_combiner) r.u128(p.joint_combiner);
    for (int i = 0; i < 8; i++) p.feature_flags[i] = r.boolean();
    for (int i = 0; i < 16; i++) r.u128(p.bulletproof_challenges[i]);
    p.proofs_verified = r.variant(3);
    p.domain_log2 = r.u8();
    for (int i = 0; i < 4; i++) p.sponge_digest_before_evaluations[i] = r.u64();
    r.point(p.wrap_challenge_polynomial_commitment);
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < 15; i++) r.u128(p.wrap_old_bulletproof_challenges[k][i]);
    size_t n = r.len(80);
    p.step_challenge_polynomial_commitments.clear();
    for (size_t i = 0; i < n && r.ok(); i++) {
        Point pt;
        r.point(pt);
        p.step_challenge_polynomial_commitments.push_back(pt);
    }
    n = r.len(256);
    p.step_old_bulletproof_challenges.clear();
    for (size_t i = 0; i < n && r.ok(); i++) {
        std::array<U128, 16> row;
        for (int j = 0; j < 16; j++) r.u128(row[j]);
        p.step_old_bulletproof_challenges.push_back(row);
    }
    PrevEvals &pe = p.prev_evals;
    r.bigint(pe.public_input[0]);
    r.bigint(pe.public_input[1]);
    for (int e = 0; e < N_EVALS; e++) {
        read_bigint_vec(r, pe.evals[e][0]);
        read_bigint_vec(r, pe.evals[e][1]);
    }
    pe.has_optional = false;
    for (int i = 0; i < 19 && r.ok(); i++) {
        if (r.option()) {  // Some((Vec<BigInt>, Vec<BigInt>)): decoded and discarded; the verifier rejects later
            std::vector<B32> a, b;
            read_bigint_vec(r, a);
            read_bigint_vec(r, b);
            pe.has_optional = true;
        }
    }
    r.bigint(pe.ft_eval1);
    WireProof &w = p.proof;
    for (int i = 0; i < 15; i++) r.point(w.w_comm[i]);
    r.point(w.z_comm);
    for (int i = 0; i < 7; i++) r.point(w.t_comm[i]);
    for (int e = 0; e < N_EVALS; e++) {
        r.bigint(w.evals[e][0]);
        r.bigint(w.evals[e][1]);
    }
    r.bigint(w.ft_eval1);
    n = r.len(320);
    w.lr.clear();
    for (size_t i = 0; i < n && r.ok(); i++) {
        std::array<Point, 2> pr;
        r.point(pr[0]);
        r.point(pr[1]);
        w.lr.push_back(pr);
    }
    r.bigint(w.z_1);
    r.bigint(w.z_2);
    r.point(w.delta);
    r.point(w.sg);
    p.wire_end = r.offset();
}

inline bool decode_state_proof(const uint8_t *data, size_t n, StateProof &out, std::string &err) {
    Reader r(data, n);
    read_pickles_proof(r, out.candidate_tip_proof);
    for (int i = 0; i < FRONTIER_LEN && r.ok(); i++) read_protocol_state(r, out.candidate_chain_states[i]);
    if (r.ok()) read_protocol_state(r, out.bridge_tip_state);
    err = r.error();
    return r.ok();
}

inline bool decode_state_pub(const uint8_t *data, size_t n, StatePubInputs &out, std::string &err) {
    Reader r(data, n);
    out.is_state_proof_from_devnet = r.boolean();
    r.raw32(out.bridge_tip_state_hash);
    for (int i = 0; i < FRONTIER_LEN; i++) r.raw32(out.candidate_chain_state_hashes[i]);
    for (int i = 0; i < FRONTIER_LEN; i++) r.raw32(out.candidate_chain_ledger_hashes[i]);
    err = r.error();
    return r.ok();
}

inline void read_zkapp(Reader &r, ZkappAccount &z) {
    // MinaBaseZkappAccountStableV2.  Layout restated from mina-p2p-messages 0.6.4 WITHOUT a fixture to
    // pin it (the reference's only account vector has zkapp = None): "layout unpinned".
    for (int i = 0; i < 8; i++) r.bigint(z.app_state[i]);
    z.has_vk = r.option();
    if (z.has_vk) {
        z.vk_max_proofs_verified = r.variant(3);
        z.vk_actual_wrap_domain_size = r.variant(3);
        for (int i = 0; i < 28; i++) r.point(z.vk_wrap_index[i]);
    }
    z.zkapp_version = r.u32();
    for (int i = 0; i < 5; i++) r.bigint(z.action_state[i]);
    z.last_action_slot = r.u32();
    z.proved_state = r.boolean();
    r.bytes(z.zkapp_uri);
}

inline bool decode_account_proof(const uint8_t *data, size_t n, AccountProof &out, std::string &err) {
    Reader r(data, n);
    size_t plen = r.len(44);
    out.merkle_path.clear();
    for (size_t i = 0; i < plen && r.ok(); i++) {
        MerkleNode node;
        node.tag = r.variant(2);
        // o1_utils SerdeAs: a byte string holding the ark-serialize encoding; the first 32 bytes are read
        size_t l = r.len(1);
        if (r.ok() && l < 32) r.fail("field element shorter than 32 bytes");
        const uint8_t *b = r.take(l);
        if (b) std::memcpy(node.hash.data(), b, 32);
        out.merkle_path.push_back(node);
    }
    Account &a = out.account;
    read_pubkey(r, a.public_key);
    r.bigint(a.token_id);
    r.bytes(a.token_symbol);
    a.balance = r.u64();
    a.nonce = r.u32();
    r.bigint(a.receipt_chain_hash);
    a.has_delegate = r.option();
    if (a.has_delegate) read_pubkey(r, a.delegate);
    r.bigint(a.voting_for);
    a.timing.timed = r.variant(2) == 1;
    if (a.timing.timed) {
        a.timing.initial_minimum_balance = r.u64();
        a.timing.cliff_time = r.u32();
        a.timing.cliff_amount = r.u64();
        a.timing.vesting_period = r.u32();
        a.timing.vesting_increment = r.u64();
    }
    for (int i = 0; i < 13; i++) {
        a.permissions.auth[i] = r.variant(5);  // None, Either, Proof, Signature, Impossible
        if (i == 6) a.permissions.set_vk_txn_version = r.u32();
    }
    a.has_zkapp = r.option();
    if (a.has_zkapp) read_zkapp(r, a.zkapp);
    err = r.error();
    return r.ok();
}

inline bool decode_account_pub(const uint8_t *data, size_t n, AccountPubInputs &out, std::string &err) {
    Reader r(data, n);
    r.raw32(out.ledger_hash);
    r.bytes(out.encoded_account);
    err = r.error();
    return r.ok();
}

}  // namespace wire
}  // namespace pasta
