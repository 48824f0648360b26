// SRS construction: g[i] = to_group(H(be32(i))), h = to_group(H("srs_misc" || be32(0))).
//
// Host-side replacement for poly-commitment `SRS::<G>::create(depth)` and groupmap `BWParameters`
// (lambdaclass/openmina-proof-systems @ 44e0d3b), which the reference evaluates lazily at first
// use: `SRS::<Vesta>::create(Fq::SRS_DEPTH)` at AL/operator/mina/lib/src/lib.rs:34 and
// `SRS::create(max_poly_size)` at AL/operator/mina/lib/src/verifier_index.rs:204-208.
// The committed srs/vesta.srs and srs/pallas.srs hold exactly these points (tests pin a SHA-256 of
// the derived arrays that was checked against those files).
#pragma once
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include "blake2b.hpp"
#include "host_field.hpp"

namespace pasta {
namespace host {

// Shallue-van de Woestijne map with u = 1 (groupmap `BWParameters::setup`).
template <class B>
struct GroupMap {
    Fe<B> fu, inv_three_u2, sqrt_neg_three_u2, sqrt_neg_three_u2_minus_u_over_2, five;
    GroupMap() {
        five = Fe<B>::from_u64(5);
        fu = Fe<B>::one() + five;  // u^3 + b with u = 1
        Fe<B> three = Fe<B>::from_u64(3);
        inv_three_u2 = three.inv();
        bool ok = (-three).sqrt(sqrt_neg_three_u2);
        (void)ok;
        sqrt_neg_three_u2_minus_u_over_2 = (sqrt_neg_three_u2 - Fe<B>::one()) * Fe<B>::from_u64(2).inv();
    }
    Affine<B> to_group(const Fe<B> &t) const {
        Fe<B> t2 = t.sqr();
        Fe<B> t2_fu = t2 + fu;
        Fe<B> alpha_inv = t2_fu * t2;
        Fe<B> alpha = alpha_inv.is_zero() ? alpha_inv : alpha_inv.inv();
        Fe<B> xs[3];
        xs[0] = sqrt_neg_three_u2_minus_u_over_2 - t2.sqr() * alpha * sqrt_neg_three_u2;
        xs[1] = -Fe<B>::one() - xs[0];
        xs[2] = Fe<B>::one() - t2_fu.sqr() * (alpha * t2_fu) * inv_three_u2;
        for (int k = 0; k < 3; k++) {
            Fe<B> y;
            if ((xs[k].sqr() * xs[k] + five).sqrt(y)) {
                Affine<B> p;
                p.x = xs[k];
                p.y = y;
                p.inf = false;
                return p;
            }
        }
        return Affine<B>::identity();  // unreachable for a valid map
    }
};

// poly-commitment `point_of_random_bytes`: first 31 digest bytes, bits LSB-first per byte, read as
// a big-endian bit string.
template <class B>
Fe<B> srs_hash_to_field(const uint8_t *msg, size_t len) {
    auto dg = Blake2b512::hash(msg, len);
    uint64_t limbs[4] = {0, 0, 0, 0};
    for (int i = 0; i < 31; i++)
        for (int j = 0; j < 8; j++)
            if ((dg[i] >> j) & 1) {
                int pos = 247 - (i * 8 + j);
                limbs[pos >> 6] |= 1ull << (pos & 63);
            }
    return Fe<B>::from_canonical(Fe<B>::raw(limbs[0], limbs[1], limbs[2], limbs[3]));
}

template <class B>
struct Srs {
    std::vector<Affine<B>> g;
    Affine<B> h;
};

template <class B>
Srs<B> srs_create(uint32_t depth, unsigned nthreads = 0) {
    Srs<B> srs;
    srs.g.resize(depth);
    GroupMap<B> gm;
    if (nthreads == 0) nthreads = std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min<unsigned>(nthreads, 64);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; t++)
        pool.emplace_back([&, t]() {
            for (uint32_t i = t; i < depth; i += nthreads) {
                uint8_t msg[4] = {(uint8_t)(i >> 24), (uint8_t)(i >> 16), (uint8_t)(i >> 8), (uint8_t)i};
                srs.g[i] = gm.to_group(srs_hash_to_field<B>(msg, 4));
            }
        });
    for (auto &th : pool) th.join();
    const uint8_t misc[12] = {'s', 'r', 's', '_', 'm', 'i', 's', 'c', 0, 0, 0, 0};
    srs.h = gm.to_group(srs_hash_to_field<B>(misc, 12));
    return srs;
}

// Flat cache: magic, depth, then (depth + 1) x 64 bytes of Montgomery-form (x, y); last entry is h.
template <class B>
bool srs_load_cache(const std::string &path, uint32_t depth, Srs<B> &out) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    uint32_t hdr[4];
    bool ok = std::fread(hdr, sizeof hdr, 1, f) == 1 && hdr[0] == 0x53525342u && hdr[1] == (uint32_t)B::ID && hdr[2] == depth;
    if (ok) {
        out.g.resize(depth);
        std::vector<uint64_t> buf((size_t)(depth + 1) * 8);
        ok = std::fread(buf.data(), 64, depth + 1, f) == depth + 1;
        if (ok) {
            for (uint32_t i = 0; i <= depth; i++) {
                Affine<B> &p = i < depth ? out.g[i] : out.h;
                std::memcpy(p.x.l, &buf[(size_t)i * 8], 32);
                std::memcpy(p.y.l, &buf[(size_t)i * 8 + 4], 32);
                p.inf = false;
            }
        }
    }
    std::fclose(f);
    return ok;
}
template <class B>
bool srs_store_cache(const std::string &path, const Srs<B> &srs) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    uint32_t hdr[4] = {0x53525342u, (uint32_t)B::ID, (uint32_t)srs.g.size(), 0};
    bool ok = std::fwrite(hdr, sizeof hdr, 1, f) == 1;
    for (size_t i = 0; ok && i <= srs.g.size(); i++) {
        const Affine<B> &p = i < srs.g.size() ? srs.g[i] : srs.h;
        ok = std::fwrite(p.x.l, 32, 1, f) == 1 && std::fwrite(p.y.l, 32, 1, f) == 1;
    }
    std::fclose(f);
    return ok;
}

}  // namespace host
}  // namespace pasta
