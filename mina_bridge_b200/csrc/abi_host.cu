// Host-only C ABI hooks: exercise the C++ host arithmetic (field, sqrt, group map, SRS derivation)
// without a GPU, so the CPU test tier can check it against the oracle.
#include <cstring>
#include <stdexcept>
#include <thread>

#include "../../include/mina_b200.h"
#include "consensus.hpp"
#include "context.cuh"
#include "group_testing.hpp"
#include "sol_account.hpp"
#include "wire.hpp"
#include "wire_write.hpp"

using namespace pasta;

namespace {
template <class F>
int host_field_op_t(int op, uint32_t n, const uint8_t *a32, const uint8_t *b32, uint8_t *out32) {
    using E = host::Fe<F>;
    for (uint32_t i = 0; i < n; i++) {
        E a, b;
        if (!E::from_bytes_le(a32 + 32 * (size_t)i, a)) return -2;
        if (!E::from_bytes_le((b32 ? b32 : a32) + 32 * (size_t)i, b)) return -2;
        E r;
        switch (op) {
            case 0: r = a * b; break;
            case 1: r = a + b; break;
            case 2: r = a - b; break;
            case 3: r = a.inv(); break;
            case 4: r = a.sqr(); break;
            case 5:
                if (!a.sqrt(r)) r = E::zero();
                break;
            default: return -3;
        }
        r.to_bytes_le(out32 + 32 * (size_t)i);
    }
    return 0;
}
template <class B>
void host_srs_t(uint32_t first, uint32_t count, uint8_t *out64, uint8_t *h64) {
    host::GroupMap<B> gm;
    unsigned nthreads = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 64);
    if (count < 256) nthreads = 1;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; t++)
        pool.emplace_back([&, t]() {
            for (uint32_t k = t; k < count; k += nthreads) {
                uint32_t i = first + k;
                uint8_t msg[4] = {(uint8_t)(i >> 24), (uint8_t)(i >> 16), (uint8_t)(i >> 8), (uint8_t)i};
                host::Affine<B> p = gm.to_group(host::srs_hash_to_field<B>(msg, 4));
                p.x.to_bytes_le(out64 + 64 * (size_t)k);
                p.y.to_bytes_le(out64 + 64 * (size_t)k + 32);
            }
        });
    for (auto &th : pool) th.join();
    if (h64) {
        const uint8_t misc[12] = {'s', 'r', 's', '_', 'm', 'i', 's', 'c', 0, 0, 0, 0};
        host::Affine<B> p = gm.to_group(host::srs_hash_to_field<B>(misc, 12));
        p.x.to_bytes_le(h64);
        p.y.to_bytes_le(h64 + 32);
    }
}
}  // namespace

template <class F>
static int host_permute_t(const uint8_t *table, uint32_t n, uint8_t *states96) {
    poseidon::Params<F> params;
    if (!params.from_bytes(table, (size_t)poseidon::TABLE_WORDS * 32)) return -1;
    for (uint32_t i = 0; i < n; i++) {
        host::Fe<F> st[3];
        for (int k = 0; k < 3; k++)
            if (!host::Fe<F>::from_bytes_le(states96 + 96 * (size_t)i + 32 * k, st[k])) return -2;
        poseidon::permute<F>(params, st);
        for (int k = 0; k < 3; k++) st[k].to_bytes_le(states96 + 96 * (size_t)i + 32 * k);
    }
    return 0;
}

extern "C" {

int mina_b200_host_field_op(int field, int op, uint32_t n, const uint8_t *a32, const uint8_t *b32, uint8_t *out32) {
    if (field == 0) return host_field_op_t<FpParams>(op, n, a32, b32, out32);
    if (field == 1) return host_field_op_t<FqParams>(op, n, a32, b32, out32);
    return -1;
}

int mina_b200_host_srs_derive(int curve, uint32_t first, uint32_t count, uint8_t *out64, uint8_t *h64) {
    if (curve == 0)
        host_srs_t<FpParams>(first, count, out64, h64);
    else if (curve == 1)
        host_srs_t<FqParams>(first, count, out64, h64);
    else
        return -1;
    return 0;
}

int mina_b200_host_build_srs_cache(const char *cache_dir) {
    if (!cache_dir || !*cache_dir) return -1;
    try {
        std::string dir(cache_dir);
        {
            std::string path = dir + "/vesta_" + std::to_string(VESTA_SRS_DEPTH) + ".srsbin";
            host::Srs<FqParams> s;
            if (!host::srs_load_cache<FqParams>(path, VESTA_SRS_DEPTH, s)) {
                s = host::srs_create<FqParams>(VESTA_SRS_DEPTH);
                if (!host::srs_store_cache<FqParams>(path, s)) return -2;
            }
        }
        {
            std::string path = dir + "/pallas_" + std::to_string(PALLAS_SRS_DEPTH) + ".srsbin";
            host::Srs<FpParams> s;
            if (!host::srs_load_cache<FpParams>(path, PALLAS_SRS_DEPTH, s)) {
                s = host::srs_create<FpParams>(PALLAS_SRS_DEPTH);
                if (!host::srs_store_cache<FpParams>(path, s)) return -2;
            }
        }
    } catch (...) {
        return -3;
    }
    return 0;
}

int mina_b200_host_srs_load_file(int curve, const char *path, uint32_t count, uint8_t *out64, uint8_t *h64) {
    std::string err;
    auto emit = [&](auto &srs) {
        for (uint32_t i = 0; i < count; i++) {
            srs.g[i].x.to_bytes_le(out64 + 64 * (size_t)i);
            srs.g[i].y.to_bytes_le(out64 + 64 * (size_t)i + 32);
        }
        if (h64) {
            srs.h.x.to_bytes_le(h64);
            srs.h.y.to_bytes_le(h64 + 32);
        }
    };
    bool ok = false;
    if (curve == 0) {
        host::Srs<FpParams> srs;
        if ((ok = host::srs_load_file<FpParams>(path, count, srs, err))) emit(srs);
    } else if (curve == 1) {
        host::Srs<FqParams> srs;
        if ((ok = host::srs_load_file<FqParams>(path, count, srs, err))) emit(srs);
    } else {
        err = "bad curve id";
    }
    if (!ok) set_error(err);
    return ok ? 0 : -1;
}

int mina_b200_host_blake2b512(const uint8_t *data, size_t len, uint8_t out[64]) {
    auto dg = host::Blake2b512::hash(data, len);
    std::memcpy(out, dg.data(), 64);
    return 0;
}


int mina_b200_host_decode(int kind, const uint8_t *data, size_t len, mina_b200_wire_summary *out) {
    try {
        std::memset(out, 0, sizeof *out);
        std::string err;
        if (kind == 0) {
            auto p = std::make_unique<wire::StateProof>();
            if (!wire::decode_state_proof(data, len, *p, err)) {
                set_error(err);
                return -1;
            }
            const wire::PicklesProof &pp = p->candidate_tip_proof;
            out->proof_end = pp.wire_end;
            out->consumed = p->bridge_tip_state.wire_end;
            out->n_step_comms = (uint32_t)pp.step_challenge_polynomial_commitments.size();
            out->n_lr = (uint32_t)pp.proof.lr.size();
            std::memcpy(out->wrap_sg, pp.wrap_challenge_polynomial_commitment.x.data(), 32);
            std::memcpy(out->wrap_sg + 32, pp.wrap_challenge_polynomial_commitment.y.data(), 32);
            for (size_t k = 0; k < 2 && k < pp.step_challenge_polynomial_commitments.size(); k++) {
                std::memcpy(out->step_sg[k], pp.step_challenge_polynomial_commitments[k].x.data(), 32);
                std::memcpy(out->step_sg[k] + 32, pp.step_challenge_polynomial_commitments[k].y.data(), 32);
            }
            for (int i = 0; i < 17; i++) {
                const wire::ProtocolState &st = i < 16 ? p->candidate_chain_states[i] : p->bridge_tip_state;
                out->blockchain_length[i] = st.consensus_state.blockchain_length;
                out->curr_global_slot[i] = st.consensus_state.curr_global_slot;
                out->epoch_count[i] = st.consensus_state.epoch_count;
                out->min_window_density[i] = st.consensus_state.min_window_density;
                out->state_begin[i] = st.wire_begin;
                out->state_end[i] = st.wire_end;
                std::memcpy(out->previous_state_hash[i], st.previous_state_hash.data(), 32);
                std::memcpy(out->first_pass_ledger[i], st.blockchain_state.target.first_pass_ledger.data(), 32);
            }
            return 0;
        }
        if (kind == 1) {
            wire::StatePubInputs pub;
            if (!wire::decode_state_pub(data, len, pub, err)) {
                set_error(err);
                return -1;
            }
            out->consumed = 1 + 33 * 32;
            out->is_devnet = pub.is_state_proof_from_devnet;
            std::memcpy(out->hash0, pub.bridge_tip_state_hash.data(), 32);
            for (int i = 0; i < 16; i++) {
                std::memcpy(out->previous_state_hash[i], pub.candidate_chain_state_hashes[i].data(), 32);
                std::memcpy(out->first_pass_ledger[i], pub.candidate_chain_ledger_hashes[i].data(), 32);
            }
            return 0;
        }
        if (kind == 2) {
            wire::AccountProof ap;
            wire::Reader probe(data, len);
            if (!wire::decode_account_proof(data, len, ap, err)) {
                set_error(err);
                return -1;
            }
            out->merkle_depth = (uint32_t)ap.merkle_path.size();
            out->balance = ap.account.balance;
            out->nonce = ap.account.nonce;
            out->has_zkapp = ap.account.has_zkapp;
            std::memcpy(out->hash0, ap.account.public_key.x.data(), 32);
            return 0;
        }
        if (kind == 3) {
            wire::AccountPubInputs pub;
            if (!wire::decode_account_pub(data, len, pub, err)) {
                set_error(err);
                return -1;
            }
            out->consumed = 40 + pub.encoded_account.size();
            out->encoded_account_len = pub.encoded_account.size();
            std::memcpy(out->hash0, pub.ledger_hash.data(), 32);
            return 0;
        }
        set_error("bad kind");
        return -1;
    } catch (const std::exception &e) {
        set_error(e.what());
        return -1;
    }
}

int mina_b200_host_select_secure_chain(const uint8_t *candidate, size_t candidate_len, const uint8_t *tip, size_t tip_len, int *result) {
    try {
        wire::ProtocolState c, t;
        wire::Reader rc(candidate, candidate_len), rt(tip, tip_len);
        wire::read_protocol_state(rc, c);
        wire::read_protocol_state(rt, t);
        if (!rc.ok() || !rt.ok()) {
            set_error(rc.ok() ? rt.error() : rc.error());
            return -1;
        }
        consensus::ChainResult res = consensus::ChainResult::Bridge;
        consensus::Status st = consensus::select_secure_chain(c, t, consensus::StateHashCmp(), res);
        if (st == consensus::Status::ConstantsDiffer) return -2;
        if (st == consensus::Status::NeedStateHash) return -3;
        *result = res == consensus::ChainResult::Candidate ? 1 : 0;
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
        return -1;
    }
}

int mina_b200_host_vk_load(const char *path, uint8_t *out, uint32_t meta[4]) {
    try {
        vk::VerifierIndex vi = vk::load_verifier_index(path);
        auto comms = vi.all_commitments();
        uint8_t *o = out;
        for (auto &p : comms) {
            if (!p.on_curve()) throw std::runtime_error("vk: commitment is not on Pallas");
            p.x.to_bytes_le(o);
            p.y.to_bytes_le(o + 32);
            o += 64;
        }
        for (int i = 0; i < 7; i++, o += 32) vi.shift[i].to_bytes_le(o);
        vi.group_gen.to_bytes_le(o);
        o += 32;
        vi.w.to_bytes_le(o);
        o += 32;
        for (int i = 0; i < 4; i++, o += 32) vi.zkpm[i].to_bytes_le(o);
        vi.endo.to_bytes_le(o);
        meta[0] = vi.log_size_of_group;
        meta[1] = vi.max_poly_size;
        meta[2] = vi.public_inputs;
        meta[3] = vi.prev_challenges;
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
        return -1;
    }
}

int mina_b200_host_hash_with_kimchi(const uint8_t *table, const char *prefix, const uint8_t *xs32, uint32_t n, uint8_t out32[32]) {
    poseidon::Params<FpParams> params;
    if (!params.from_bytes(table, (size_t)poseidon::TABLE_WORDS * 32)) return -1;
    std::vector<host::Fp> xs(n);
    for (uint32_t i = 0; i < n; i++)
        if (!host::Fp::from_bytes_le(xs32 + 32 * (size_t)i, xs[i])) return -2;
    host::Fp out;
    if (!poseidon::hash_with_kimchi<FpParams>(params, prefix, xs.data(), n, out)) return -3;
    out.to_bytes_le(out32);
    return 0;
}

static int emit_bytes(const std::vector<uint8_t> &enc, uint8_t *out, size_t *out_len) {
    if (!out_len || enc.size() > *out_len) {
        set_error("output buffer too small");
        return -2;
    }
    if (!enc.empty()) std::memcpy(out, enc.data(), enc.size());
    *out_len = enc.size();
    return 0;
}

int mina_b200_host_group_testing_sim(uint32_t m, const uint8_t *bad, uint8_t *ok_out, uint32_t *levels, uint32_t *msms) {
    try {
        std::vector<uint8_t> b(bad, bad + m), ok;
        uint32_t n_msm = 0;
        uint32_t lv = m ? gt::simulate(m, b, ok, n_msm) : 0;
        if (m) std::memcpy(ok_out, ok.data(), m);
        if (levels) *levels = lv;
        if (msms) *msms = n_msm;
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
        return -1;
    }
}

int mina_b200_host_reencode(int kind, const uint8_t *data, size_t len, uint8_t *out, size_t *out_len) {
    try {
        std::string err;
        std::vector<uint8_t> enc;
        if (kind == 0) {
            auto p = std::make_unique<wire::StateProof>();
            if (!wire::decode_state_proof(data, len, *p, err)) return set_error(err), -1;
            if (!wire::encode_state_proof(*p, enc)) return set_error("proof carries optional evaluations"), -2;
        } else if (kind == 1) {
            wire::StatePubInputs pub;
            if (!wire::decode_state_pub(data, len, pub, err)) return set_error(err), -1;
            wire::encode_state_pub(pub, enc);
        } else if (kind == 2) {
            wire::AccountProof ap;
            if (!wire::decode_account_proof(data, len, ap, err)) return set_error(err), -1;
            wire::encode_account_proof(ap, enc);
        } else if (kind == 3) {
            wire::AccountPubInputs pub;
            if (!wire::decode_account_pub(data, len, pub, err)) return set_error(err), -1;
            wire::encode_account_pub(pub, enc);
        } else {
            return set_error("bad kind"), -1;
        }
        return emit_bytes(enc, out, out_len);
    } catch (const std::exception &e) {
        set_error(e.what());
        return -1;
    }
}

int mina_b200_host_account_abi_encode(const uint8_t *account_proof, size_t len, uint8_t *out, size_t *out_len) {
    try {
        std::string err;
        wire::AccountProof ap;
        if (!wire::decode_account_proof(account_proof, len, ap, err)) return set_error(err), -1;
        std::vector<uint8_t> enc;
        if (!sol::abi_encode_account(ap.account, enc)) return set_error("token symbol is not UTF-8"), -2;
        return emit_bytes(enc, out, out_len);
    } catch (const std::exception &e) {
        set_error(e.what());
        return -1;
    }
}

int mina_b200_host_poseidon_permute(int field, const uint8_t *table, uint32_t n, uint8_t *states96) {
    if (field == 0) return host_permute_t<FpParams>(table, n, states96);
    if (field == 1) return host_permute_t<FqParams>(table, n, states96);
    return -1;
}

}  // extern "C"
