// Ouroboros Samasika fork choice between the candidate tip and the bridge's tip.
//
// Host-side replacement for AL/operator/mina/lib/src/consensus_state.rs:22-146 (select_secure_chain,
// select_longer_chain, is_short_range, relative_min_window_density, hash_last_vrf).  Pure u32 logic
// plus one Blake2b-512 on height ties; bug-compatible with the reference:
//   Q1  the long-range branch evaluates BOTH densities with (candidate, tip), so they are always
//       equal and the rule degenerates to select_longer_chain (consensus_state.rs:29-35);
//   Q2  relative_min_window_density only reads the candidate's window (consensus_state.rs:91-134);
//   Q3  ties go to the bridge: the candidate must be strictly better (consensus_state.rs:43-59).
// The last tiebreak compares Poseidon state hashes (hash_state, consensus_state.rs:148-150); the
// caller supplies that comparison when it can compute it.
#pragma once
#include <algorithm>
#include <functional>
#include <string>

#include "blake2b.hpp"
#include "wire.hpp"

namespace pasta {
namespace consensus {

static constexpr uint32_t GRACE_PERIOD_END = 1440;      // consensus_state.rs:12
static constexpr uint32_t SUB_WINDOWS_PER_WINDOW = 11;  // :13
static constexpr uint32_t SLOTS_PER_SUB_WINDOW = 7;     // :14

enum class ChainResult { Bridge = 0, Candidate = 1 };
enum class Status { Ok = 0, ConstantsDiffer = 1, NeedStateHash = 2 };

// returns <0, 0, >0 like memcmp over the hex strings the reference compares (hex of the digest bytes
// orders exactly like the bytes themselves)
inline int compare_last_vrf(const wire::ProtocolState &a, const wire::ProtocolState &b) {
    auto da = host::Blake2b512::hash(a.consensus_state.last_vrf_output.data(), a.consensus_state.last_vrf_output.size());
    auto db = host::Blake2b512::hash(b.consensus_state.last_vrf_output.data(), b.consensus_state.last_vrf_output.size());
    return std::memcmp(da.data(), db.data(), 64);
}

// state_hash_cmp(candidate, tip) -> sign of hex(hash(candidate)) vs hex(hash(tip)); may be empty.
using StateHashCmp = std::function<bool(const wire::ProtocolState &, const wire::ProtocolState &, int &)>;

inline Status select_longer_chain(const wire::ProtocolState &candidate, const wire::ProtocolState &tip, const StateHashCmp &cmp,
                                  ChainResult &out) {
    uint32_t ch = candidate.consensus_state.blockchain_length, th = tip.consensus_state.blockchain_length;
    out = ChainResult::Bridge;
    if (ch > th) {
        out = ChainResult::Candidate;
    } else if (ch == th) {
        int v = compare_last_vrf(candidate, tip);
        if (v > 0) {
            out = ChainResult::Candidate;
        } else if (v == 0) {
            int s = 0;
            if (!cmp || !cmp(candidate, tip, s)) return Status::NeedStateHash;
            if (s > 0) out = ChainResult::Candidate;
        }
    }
    return Status::Ok;
}

inline bool is_short_range(const wire::ProtocolState &candidate, const wire::ProtocolState &tip, bool &out) {
    if (!(tip.constants == candidate.constants)) return false;
    const uint32_t slots_per_epoch = tip.constants.slots_per_epoch;
    const wire::ConsensusState &c = candidate.consensus_state, &t = tip.consensus_state;
    auto check = [&](const wire::ConsensusState &s1, const wire::ConsensusState &s2) {
        // a zero slots_per_epoch panics in the reference (remainder by zero); treated as "not short range"
        if (slots_per_epoch == 0) return false;
        uint32_t s2_epoch_slot = s2.curr_global_slot % slots_per_epoch;
        // u32 arithmetic as in the reference: epoch_count + 1 and slots_per_epoch * 2 wrap only on absurd inputs
        if (s1.epoch_count == s2.epoch_count + 1 && s2_epoch_slot >= slots_per_epoch * 2 / 3)
            return s1.staking_epoch_data.lock_checkpoint == s2.next_epoch_data.lock_checkpoint;
        return false;
    };
    if (c.epoch_count == t.epoch_count)
        out = c.staking_epoch_data.lock_checkpoint == t.staking_epoch_data.lock_checkpoint;
    else
        out = check(c, t) || check(t, c);
    return true;
}

inline uint32_t relative_sub_window(const wire::ConsensusState &s) {
    return (s.curr_global_slot / SLOTS_PER_SUB_WINDOW) % SUB_WINDOWS_PER_WINDOW;
}

inline uint32_t relative_min_window_density(const wire::ProtocolState &candidate, const wire::ProtocolState &tip) {
    const wire::ConsensusState &c = candidate.consensus_state, &t = tip.consensus_state;
    uint32_t max_slot = std::max(c.curr_global_slot, t.curr_global_slot);
    if (max_slot < GRACE_PERIOD_END) return c.min_window_density;
    uint32_t shift_count = 0;
    if (max_slot > c.curr_global_slot) shift_count = max_slot - c.curr_global_slot - 1;  // checked_sub twice, else 0
    shift_count = std::min(shift_count, SUB_WINDOWS_PER_WINDOW);
    std::vector<uint32_t> window(c.sub_window_densities);
    uint32_t i = relative_sub_window(c);
    for (uint32_t k = 0; k < shift_count; k++) {
        i = (i + 1) % SUB_WINDOWS_PER_WINDOW;
        if (i < window.size()) window[i] = 0;
    }
    uint32_t density = 0;
    for (uint32_t d : window) density += d;  // u32 sum (the reference's iter().sum() would panic on overflow in debug only)
    return std::min(c.min_window_density, density);
}

inline Status select_secure_chain(const wire::ProtocolState &candidate, const wire::ProtocolState &tip, const StateHashCmp &cmp,
                                  ChainResult &out) {
    bool short_range = false;
    if (!is_short_range(candidate, tip, short_range)) return Status::ConstantsDiffer;
    if (short_range) return select_longer_chain(candidate, tip, cmp, out);
    uint32_t tip_density = relative_min_window_density(candidate, tip);        // Q1: same arguments,
    uint32_t candidate_density = relative_min_window_density(candidate, tip);  // on purpose
    if (candidate_density < tip_density) {
        out = ChainResult::Bridge;
        return Status::Ok;
    }
    if (candidate_density > tip_density) {
        out = ChainResult::Candidate;
        return Status::Ok;
    }
    return select_longer_chain(candidate, tip, cmp, out);
}

}  // namespace consensus
}  // namespace pasta
