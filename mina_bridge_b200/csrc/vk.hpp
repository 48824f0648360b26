// Blockchain verification keys (devnet / mainnet): loader and derived constants.
//
// Host-side replacement for `deserialize_blockchain_vk`, `zk_w3` and `zk_polynomial`
// (AL/operator/mina/lib/src/verifier_index.rs:109-276) and for the lazy statics
// DEVNET_VERIFIER_INDEX / MAINNET_VERIFIER_INDEX (AL/operator/mina/lib/src/lib.rs:23-33).
// Consumes the same JSON shape the reference embeds: `commitments.{sigma_comm[7],
// coefficients_comm[15], generic_comm, psm_comm, complete_add_comm, mul_comm, emul_comm,
// endomul_scalar_comm}` as ["0x<hex BE>", "0x<hex BE>"] pairs and `index.{domain.log_size_of_group,
// max_poly_size, public, prev_challenges, shifts[7]}`; every other key is ignored
// (verifier_index.rs:28-69).  Hex strings are reduced mod the field like
// `from_be_bytes_mod_order` (verifier_index.rs:71-87); points are taken as finite (Q6).
#pragma once
#include <fstream>
#include <sstream>

#include "host_field.hpp"
#include "json.hpp"

namespace pasta {
namespace vk {

using host::Affine;
using host::Fe;
using FpE = host::Fe<FpParams>;
using FqE = host::Fe<FqParams>;

static constexpr uint64_t ZK_ROWS = 3;  // verifier_index.rs:253

struct VerifierIndex {
    // commitments on Pallas (coordinates in Fp)
    Affine<FpParams> sigma_comm[7], coefficients_comm[15];
    Affine<FpParams> generic_comm, psm_comm, complete_add_comm, mul_comm, emul_comm, endomul_scalar_comm;
    FqE shift[7];
    uint32_t log_size_of_group = 0, max_poly_size = 0, public_inputs = 0, prev_challenges = 0;
    FqE group_gen;    // generator of the 2^log_size_of_group evaluation domain in Fq
    FqE w;            // zk_w3: group_gen^(n - 3)
    FqE zkpm[4];      // zk_polynomial coefficients, low degree first
    FqE endo;         // endos::<Vesta>().0: cube root of unity in Fq (verifier_index.rs:169)

    // all 28 commitments in the order sigma[7], coefficients[15], generic, psm, complete_add, mul, emul,
    // endomul_scalar (the order the Pickles step message hashes them in)
    std::vector<Affine<FpParams>> all_commitments() const {
        std::vector<Affine<FpParams>> v;
        for (auto &p : sigma_comm) v.push_back(p);
        for (auto &p : coefficients_comm) v.push_back(p);
        v.push_back(generic_comm);
        v.push_back(psm_comm);
        v.push_back(complete_add_comm);
        v.push_back(mul_comm);
        v.push_back(emul_comm);
        v.push_back(endomul_scalar_comm);
        return v;
    }
};

inline int hex_nibble(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

// "0x..." big-endian hex of any even length, reduced mod the field
template <class F>
Fe<F> field_from_hex_be(const std::string &s) {
    size_t i = 0;
    while (i + 1 < s.size() && s[i] == '0' && s[i + 1] == 'x') i += 2;  // trim_start_matches("0x")
    if ((s.size() - i) % 2) throw std::runtime_error("vk: odd-length hex string");
    Fe<F> acc = Fe<F>::zero();
    const Fe<F> k256 = Fe<F>::from_u64(256);
    for (; i < s.size(); i += 2) {
        int hi = hex_nibble(s[i]), lo = hex_nibble(s[i + 1]);
        if (hi < 0 || lo < 0) throw std::runtime_error("vk: invalid hex digit");
        acc = acc * k256 + Fe<F>::from_u64((uint64_t)(hi * 16 + lo));
    }
    return acc;
}

inline Affine<FpParams> point_from_json(const json::Value &v) {
    if (v.kind != json::Value::Array || v.arr.size() != 2 || v.arr[0].kind != json::Value::String ||
        v.arr[1].kind != json::Value::String)
        throw std::runtime_error("vk: commitment is not a pair of hex strings");
    Affine<FpParams> p;
    p.x = field_from_hex_be<FpParams>(v.arr[0].str);
    p.y = field_from_hex_be<FpParams>(v.arr[1].str);
    p.inf = false;
    return p;
}

inline uint32_t uint_from_json(const json::Value &v, const char *what) {
    if (v.kind != json::Value::Number || v.num < 0 || v.num > 4294967295.0 || v.num != (double)(uint64_t)v.num)
        throw std::runtime_error(std::string("vk: ") + what + " is not an unsigned integer");
    return (uint32_t)v.num;
}

inline VerifierIndex parse_verifier_index(const std::string &text) {
    json::Value root = json::parse(text);
    const json::Value &c = root.at("commitments");
    const json::Value &ix = root.at("index");
    VerifierIndex vi;
    auto fixed_array = [&](const json::Value &a, size_t n, const char *what) -> const std::vector<json::Value> & {
        if (a.kind != json::Value::Array || a.arr.size() != n) throw std::runtime_error(std::string("vk: bad length of ") + what);
        return a.arr;
    };
    {
        auto &a = fixed_array(c.at("sigma_comm"), 7, "sigma_comm");
        for (int i = 0; i < 7; i++) vi.sigma_comm[i] = point_from_json(a[i]);
    }
    {
        auto &a = fixed_array(c.at("coefficients_comm"), 15, "coefficients_comm");
        for (int i = 0; i < 15; i++) vi.coefficients_comm[i] = point_from_json(a[i]);
    }
    vi.generic_comm = point_from_json(c.at("generic_comm"));
    vi.psm_comm = point_from_json(c.at("psm_comm"));
    vi.complete_add_comm = point_from_json(c.at("complete_add_comm"));
    vi.mul_comm = point_from_json(c.at("mul_comm"));
    vi.emul_comm = point_from_json(c.at("emul_comm"));
    vi.endomul_scalar_comm = point_from_json(c.at("endomul_scalar_comm"));
    vi.log_size_of_group = uint_from_json(ix.at("domain").at("log_size_of_group"), "log_size_of_group");
    vi.max_poly_size = uint_from_json(ix.at("max_poly_size"), "max_poly_size");
    vi.public_inputs = uint_from_json(ix.at("public"), "public");
    vi.prev_challenges = uint_from_json(ix.at("prev_challenges"), "prev_challenges");
    {
        auto &a = fixed_array(ix.at("shifts"), 7, "shifts");
        for (int i = 0; i < 7; i++) {
            if (a[i].kind != json::Value::String) throw std::runtime_error("vk: shift is not a hex string");
            vi.shift[i] = field_from_hex_be<FqParams>(a[i].str);
        }
    }
    // Radix2EvaluationDomain::new(1 << log): group_gen = (2^32-th root of unity)^(2^(32 - log))
    if (vi.log_size_of_group > 32) throw std::runtime_error("vk: failed to create domain");
    FqE g = FqE::raw(FqParams::ROOT_OF_UNITY_64(0), FqParams::ROOT_OF_UNITY_64(1), FqParams::ROOT_OF_UNITY_64(2),
                     FqParams::ROOT_OF_UNITY_64(3));
    for (uint32_t i = vi.log_size_of_group; i < 32; i++) g = g.sqr();
    vi.group_gen = g;
    // zk_w3 / zk_polynomial (verifier_index.rs:252-276)
    const uint64_t n = 1ull << vi.log_size_of_group;
    FqE w3 = g.pow_u64(n - ZK_ROWS);
    FqE w2 = g * w3, w1 = g * w2;
    FqE w1w2 = w1 * w2;
    vi.w = w3;
    vi.zkpm[0] = -(w1w2 * w3);
    vi.zkpm[1] = w1w2 + w1 * w3 + w3 * w2;
    vi.zkpm[2] = -w1 - w2 - w3;
    vi.zkpm[3] = FqE::one();
    // endos::<Vesta>().0 = 5^((q-1)/3) in Fq; ENDO_R of FqParams is its square (the scalar-challenge endo)
    FqE endo_r = FqE::raw(FqParams::ENDO_R_64(0), FqParams::ENDO_R_64(1), FqParams::ENDO_R_64(2), FqParams::ENDO_R_64(3));
    vi.endo = endo_r.sqr();  // (w^2)^2 = w^4 = w
    return vi;
}

inline VerifierIndex load_verifier_index(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("vk: cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse_verifier_index(ss.str());
}

}  // namespace vk
}  // namespace pasta
