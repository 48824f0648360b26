// Solidity ABI encoding of a Mina account: the `expected_encoded_account` of the account verifier.
//
// Replaces `MinaAccountValidationExample::Account::try_from(&account)?.abi_encode()`
// (AL/operator/mina_account/lib/src/lib.rs:54-66; conversion rules core/src/sol/account.rs:25-314; struct
// layout contract/src/MinaAccountValidationExample.sol:75-163).  The verifier compares these bytes with the
// `encoded_account` carried in the public inputs, so they must be identical to alloy's `abi_encode()` of a
// single dynamic struct value: one offset word (0x20), then the tuple -- static members inline, dynamic
// members (`string tokenSymbol`, the `ZkappAccount` tuple because of its `bytes zkappUri`) as offsets
// into the tail.
//
// Parity: re-encoding the account of mina_account.proof reproduces bytes 40.. of mina_account.pub
// (3 456 B) exactly (tests/test_boundary_cpu.py::test_account_abi_encoding_matches_the_fixture).
#pragma once
#include "wire.hpp"

namespace pasta {
namespace sol {

class AbiWords {
   public:
    std::vector<uint8_t> out;
    void word_u64(uint64_t v) {
        size_t o = out.size();
        out.resize(o + 32, 0);
        for (int i = 0; i < 8; i++) out[o + 31 - i] = (uint8_t)(v >> (8 * i));
    }
    void word_bytes32(const wire::B32 &b) { out.insert(out.end(), b.begin(), b.end()); }  // FixedBytes<32>: verbatim
    void word_zero() { out.resize(out.size() + 32, 0); }
    // `bytes` / `string`: length word, then the data right-padded to a multiple of 32
    void dynamic_bytes(const std::vector<uint8_t> &b) {
        word_u64(b.size());
        out.insert(out.end(), b.begin(), b.end());
        out.resize(out.size() + (32 - b.size() % 32) % 32, 0);
    }
    void append(const AbiWords &o) { out.insert(out.end(), o.out.begin(), o.out.end()); }
    size_t size() const { return out.size(); }
};

inline bool is_utf8(const std::vector<uint8_t> &s) {  // what String::from_utf8 accepts
    size_t i = 0, n = s.size();
    while (i < n) {
        uint8_t c = s[i];
        size_t need;
        uint32_t cp;
        if (c < 0x80) {
            i++;
            continue;
        } else if ((c & 0xe0) == 0xc0) {
            need = 1;
            cp = c & 0x1f;
        } else if ((c & 0xf0) == 0xe0) {
            need = 2;
            cp = c & 0x0f;
        } else if ((c & 0xf8) == 0xf0) {
            need = 3;
            cp = c & 0x07;
        } else {
            return false;
        }
        if (i + need >= n) return false;  // truncated sequence
        for (size_t k = 1; k <= need; k++) {
            if ((s[i + k] & 0xc0) != 0x80) return false;
            cp = (cp << 6) | (s[i + k] & 0x3f);
        }
        if ((need == 1 && cp < 0x80) || (need == 2 && cp < 0x800) || (need == 3 && (cp < 0x10000 || cp > 0x10ffff))) return false;
        if (cp >= 0xd800 && cp <= 0xdfff) return false;
        i += need + 1;
    }
    return true;
}

// ZkappAccount tuple (75 head words + the zkappUri tail).  `z == nullptr` is the empty ZkappAccount the
// reference substitutes for `zkapp: None` (account.rs:266-297).
inline AbiWords encode_zkapp(const wire::ZkappAccount *z) {
    AbiWords w;
    for (int i = 0; i < 8; i++) z ? w.word_bytes32(z->app_state[i]) : w.word_zero();
    if (z && z->has_vk) {
        w.word_u64(z->vk_max_proofs_verified);       // N0/N1/N2 -> 0/1/2
        w.word_u64(z->vk_actual_wrap_domain_size);
        // sigmaComm[7], coefficientsComm[15], generic, psm, completeAdd, mul, emul, endomulScalar: (x, y) each
        for (int i = 0; i < 28; i++) {
            w.word_bytes32(z->vk_wrap_index[i].x);
            w.word_bytes32(z->vk_wrap_index[i].y);
        }
    } else {
        for (int i = 0; i < 2 + 56; i++) w.word_zero();
    }
    w.word_u64(z ? z->zkapp_version : 0);
    for (int i = 0; i < 5; i++) z ? w.word_bytes32(z->action_state[i]) : w.word_zero();
    w.word_u64(z ? z->last_action_slot : 0);
    w.word_u64(z && z->proved_state ? 1 : 0);
    w.word_u64(75 * 32);  // offset of zkappUri inside this tuple
    w.dynamic_bytes(z ? z->zkapp_uri : std::vector<uint8_t>());
    return w;
}

// Returns false where the reference's TryFrom fails (token symbol that is not UTF-8).
inline bool abi_encode_account(const wire::Account &a, std::vector<uint8_t> &out) {
    if (!is_utf8(a.token_symbol)) return false;
    AbiWords sym, zk = encode_zkapp(a.has_zkapp ? &a.zkapp : nullptr);
    sym.dynamic_bytes(a.token_symbol);
    const uint64_t HEAD = 30 * 32;
    AbiWords w;
    w.word_u64(0x20);
    w.word_bytes32(a.public_key.x);
    w.word_u64(a.public_key.is_odd ? 1 : 0);
    w.word_bytes32(a.token_id);
    w.word_u64(HEAD);  // tokenSymbol
    w.word_u64(a.balance);
    w.word_u64(a.nonce);
    w.word_bytes32(a.receipt_chain_hash);
    if (a.has_delegate) {
        w.word_bytes32(a.delegate.x);
        w.word_u64(a.delegate.is_odd ? 1 : 0);
    } else {  // account.rs:64-69
        w.word_zero();
        w.word_u64(1);
    }
    w.word_bytes32(a.voting_for);
    const wire::Timing &t = a.timing;  // Untimed -> all zero (account.rs:104-110)
    w.word_u64(t.timed ? t.initial_minimum_balance : 0);
    w.word_u64(t.timed ? t.cliff_time : 0);
    w.word_u64(t.timed ? t.cliff_amount : 0);
    w.word_u64(t.timed ? t.vesting_period : 0);
    w.word_u64(t.timed ? t.vesting_increment : 0);
    for (int i = 0; i < 13; i++) {  // None, Either, Proof, Signature, Impossible -> 0..4 (account.rs:112-120)
        w.word_u64(a.permissions.auth[i]);
        if (i == 6) w.word_u64(a.permissions.set_vk_txn_version);
    }
    w.word_u64(HEAD + sym.size());  // zkapp
    w.append(sym);
    w.append(zk);
    out.swap(w.out);
    return true;
}

}  // namespace sol
}  // namespace pasta
